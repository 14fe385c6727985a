#!/usr/bin/env python
"""Developer tool: per-block timelines of the four phase kernels (duet_debug_timers).
    python tools/kernel_timeline.py [c2|c4|c1|c5] [flags,flags,...]
Per DUET_FLAGS variant (developer switches of csrc/phase_kernels.cuh; default 0): the device time of a
graph replay (CUDA events, L2 flushed, instrumentation off) and the per-kernel event times, then, with the
instrumentation on, per kernel: launch ramp (first/last block start), per-mark mean and max times since the
block started, and when the last block ended -- i.e. where a kernel's microseconds actually go."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_sample  # noqa: E402
from duet_b200.columnar import from_synth  # noqa: E402
from duet_b200.engine import PhaseEngine, pin_batch  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
do_flush = "--no-flush" not in sys.argv
clean = "--clean-flush" in sys.argv      # after the 512 MiB memset, READ another 512 MiB: cold but clean L2 (no write-back storm)
batch = pin_batch(from_synth(make_sample(wl, 0), with_text=False))
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
flush2 = torch.zeros(128 << 20, dtype=torch.int32, device="cuda")


def do_flush_now():
    flush.zero_()
    if clean:
        flush2.sum()

names = ["k_table (1: k_init done, 2: end)", "k_probe (1: k_table done; 2: filter in shared memory; 3: stream done; 4: candidates resolved)", "k_reduce (1: gathered, 2: end)",
         "k_tail (1: candidates in the set, 2: contig list sorted, 3: decided, 4: order written)"]
ref = None
for flags in variants:
    os.environ["DUET_FLAGS"] = str(flags)
    eng = PhaseEngine(0)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    eng.upload(batch)
    ms, kms = [], {}
    with torch.cuda.stream(stream):
        for it in range(25):
            if do_flush:
                do_flush_now()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); eng.execute(); b.record(stream)
            stream.synchronize()
            if it >= 5:
                ms.append(a.elapsed_time(b))
        for it in range(10):
            if do_flush:
                do_flush_now()
            eng.execute(per_kernel=True)
            stream.synchronize()
            for k, v in eng.timings()["kernel_ms"].items():
                kms[k] = kms.get(k, 0.0) + v / 10
    res = eng.download()
    if ref is None:
        ref = res
    same = all(np.array_equal(getattr(ref, k), getattr(res, k)) for k in ("gt", "ps", "join_row", "order", "shard_counts"))
    print(f"=== {wl} DUET_FLAGS={flags}{'' if do_flush else ' (L2 NOT flushed)'}{' (clean flush)' if clean else ''}: graph replay {np.mean(ms) * 1e3:.1f} us (min {np.min(ms) * 1e3:.1f}); serial per kernel (us): "
          + ", ".join(f"{k} {v * 1e3:.1f}" for k, v in kms.items() if v > 0.004) + f"; results equal to first variant: {same}")
    eng.lib.duet_debug_timers(eng.h, 1, None)
    eng.upload(batch)
    out = np.zeros((4, 2048, 8, 2), np.int64)
    with torch.cuda.stream(stream):
        for it in range(4):
            if do_flush:
                do_flush_now()
            stream.synchronize()
            eng.execute()
            eng.sync()
            eng.lib.duet_debug_timers(eng.h, 1, out.ctypes.data)
    t_first = None
    for k, nm in enumerate(names):
        g = out[k, :, :, 0].astype(np.float64)
        used = g[:, 0] > 0
        if not used.any():
            continue
        g = g[used]
        clk = out[k, used, :, 1].astype(np.float64)
        t0 = g[:, 0].min()
        if t_first is None:
            t_first = t0
        marks = [m for m in range(8) if (g[:, m] > 0).any()]
        end = max(g[:, m].max() for m in marks)
        print(f"{nm}: blocks {used.sum()}  kernel start +{(t0 - t_first) / 1e3:.1f} us  span {(end - t0) / 1e3:.1f} us  "
              f"last block start +{(g[:, 0].max() - t0) / 1e3:.1f} us  end +{(end - t_first) / 1e3:.1f} us")
        for m in marks[1:]:
            ok = g[:, m] > 0
            d = (clk[ok, m] - clk[ok, 0]) / 1.965e3  # SM cycles -> us at 1965 MHz
            late = (g[ok, m] - t0) / 1e3
            print(f"    mark {m}: blocks {ok.sum():5d}  since block start mean {d.mean():7.2f} us  max {d.max():7.2f} us   "
                  f"since kernel start mean {late.mean():7.2f}  max {late.max():7.2f} us")
    eng.close()
