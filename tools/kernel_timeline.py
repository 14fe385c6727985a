#!/usr/bin/env python
"""Developer tool: per-block timelines of the four phase kernels (duet_debug_timers).
    python tools/kernel_timeline.py [c2|c4|c1]
Prints, per kernel: launch ramp (first/last block start), per-mark mean and max times since the
block started, and when the last block ended -- i.e. where a kernel's microseconds actually go."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import make_sample  # noqa: E402
from duet_b200.columnar import from_synth  # noqa: E402
from duet_b200.engine import PhaseEngine, pin_batch  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
batch = pin_batch(from_synth(make_sample(wl, 0), with_text=False))
eng = PhaseEngine(0)
eng.lib.duet_debug_timers(eng.h, 1, None)
eng.upload(batch)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
out = np.zeros((4, 2048, 8, 2), np.int64)
for it in range(4):
    flush.zero_()
    torch.cuda.synchronize()
    eng.execute()
    eng.sync()
    eng.lib.duet_debug_timers(eng.h, 1, out.ctypes.data)
names = ["k_table", "k_probe (mark 2: stream done, 3: candidates resolved)", "k_reduce", "k_predict"]
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
          f"last block start +{(g[:, 0].max() - t0) / 1e3:.1f} us")
    for m in marks[1:]:
        ok = g[:, m] > 0
        base_mark = 0
        d = (clk[ok, m] - clk[ok, base_mark]) / 1.965e3  # SM cycles -> us at 1965 MHz
        late = (g[ok, m] - t0) / 1e3
        print(f"    mark {m}: blocks {ok.sum():5d}  since block start mean {d.mean():7.2f} us  max {d.max():7.2f} us   "
              f"since kernel start mean {late.mean():7.2f}  max {late.max():7.2f} us")
