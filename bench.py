#!/usr/bin/env python
"""Benchmark of the sv_phasing hot path (BASELINE.json metric: SVs phased/sec & support-read
joins/sec at 1/2/4/8 B200 vs the host-CPU reference).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5]
                    [--scaling weak|strong] [--impl b200|reference]

One "step" = one pass of the whole device path (join table build -> read stream/probe ->
per-SV reductions -> per-contig one-PS list, decision tree, emission order) over one batch.
Headline (N=1, no flags): C2 = BASELINE.json configs[1]; the same line carries the other four
configs under `configs` (C1 chr21 demo, C3 SVIM signature clustering, C4 60x dense, C5 one
GPU's share of the 32-sample cohort), each with its device value, its e2e value and its
roofline fraction, measured with the same rules.

N>1 (torchrun, one process per GPU): the line's own figures are WEAK scaling -- rank r owns
sample r of an N-sample cohort; shards are (sample, contig) pairs and never exchange data, so no
collective sits on the data path.  The same line carries `strong`: ONE C4 sample LPT-sharded by
contig over the N GPUs (duet_b200.sharding.lpt_assign), device time = max over ranks, the NCCL
counter all-gather inside the timed call, every rank's slice compared with the unsharded result.
`--scaling strong` makes that the line's headline instead.

Timing: CUDA events on the stream the kernels are launched on, one event pair per step, L2
flushed (512 MiB memset, outside the event pair) between steps; max over ranks.  `e2e` = the
public API (PhaseEngine.run) on page-locked HOST columns, wall clock, copies included.

--impl reference: the UNMODIFIED reference (oracle/_ref, pip-installed by oracle/build_ref.py from
/root/reference) on the same workload written out as the files it reads; one Python thread (its
hot path is single-threaded).  `value` = post-decode compute (join + classify + predict + sort:
the span the device path covers); `stage_wall_s` and `decode_s` are reported beside it.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "SVs phased/sec (support-read joins/sec in joins_per_sec)"
METRIC_C3 = "SV signatures clustered/sec (span-position distance, cluster_max_distance=0.9)"
WORKLOADS = {
    "c1": "C1 chr21 demo shape (70k ONT reads, 2.5k SVs)",
    "c2": "C2 synthetic GRCh37 WGS 30x ONT (4.5M reads, 25k cuteSV SVs)",
    "c3": "C3 SVIM base caller: span-position-distance clustering of 2M SV signatures, cluster_max_distance=0.9",
    "c4": "C4 synthetic WGS 60x ONT dense support lists (9M reads, 30k SVs)",
    "c5": "C5 cohort share of one GPU: 4 samples x WGS 30x (32 samples over 8 GPUs), 96 shards in one call",
}
CPU_SAMPLE_CONTIGS = ["17", "18", "19", "20", "21", "22"]      # 12.3 % of GRCh37: bounded CPU sample
C3_N = 2_000_000
C3_CPU_SAMPLE = C3_N          # the CPU restatement is vectorised numpy + scipy: the whole workload takes ~2 s
KERNEL_LABEL = {"init": "k_init", "build": "k_table", "probe": "k_probe", "reduce": "k_reduce", "tail": "k_tail", "oneps": "k_oneps",
                "predict": "k_predict", "order": "k_order"}


def make_sample(workload: str, seed: int):
    from duet_b200 import synth
    if workload == "c5":                               # the 4 samples a GPU owns in the 32-sample cohort
        return [synth.config_c2(4 * seed + k, id_base=(4 * seed + k) << 44) for k in range(4)]
    return {"c1": synth.config_c1, "c2": synth.config_c2, "c4": synth.config_c4}[workload](seed)


def sub_sample(sample, contigs):
    from duet_b200 import synth
    keep = [c for c in sample.contigs if c.name in contigs] or sample.contigs
    return synth.SynthSample(keep, chr_prefix=sample.chr_prefix, seed=sample.seed)


# ----------------------------------------------------------------------------------------------
# clocks (pynvml, sampled while the timed loops run)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.period_s = float(os.environ.get("BENCH_CLOCK_PERIOD_MS", "2")) * 1e-3
        self._stop = threading.Event()
        self._active = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        while not self._stop.is_set():
            if self.nv is not None and self._active.is_set():
                try:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.dev, self.nv.NVML_CLOCK_SM))
                    bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                    for b, name in self.REASONS.items():
                        if bits & b and name != "gpu_idle":
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(self.period_s)

    def __enter__(self):
        self._active.set()
        return self

    def __exit__(self, *a):
        self._active.clear()

    def summary(self):
        self._stop.set()
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# CPU legs: the only place bench.py executes oracle/ (the unmodified reference in oracle/_ref when it
# has been installed, else the oracle port)
# ----------------------------------------------------------------------------------------------
def reference_phase(sample, steps: int, warmup: int, full_stage_runs: int = 1) -> dict:
    """The unmodified reference on `sample` written out as files.  One Python thread."""
    from oracle import ref_runner
    run = ref_runner.ReferenceRun(sample)
    try:
        walls = []
        rows_full = None
        for _ in range(full_stage_runs):
            w, rows_full = run.stage_wall()
            walls.append(w)
        dec = run.decode()
        times, rows = run.post_decode_steps(warmup + steps)
        assert rows_full is None or rows == rows_full
        times = times[warmup:]
        return {"kind": "reference", "seconds": float(np.mean(times)), "best_seconds": float(np.min(times)),
                "n_svs": run.n_svs, "n_joins": run.n_joins, "n_tagged": run.n_tagged, "rows": len(rows),
                "stage_wall_s": float(np.min(walls)) if walls else None, **dec,
                "input_mb": round(run.input_bytes / 1e6, 1), "workdir_write_s": round(run.write_s, 1)}
    finally:
        run.close()


def port_phase(sample, steps: int, warmup: int) -> dict:
    """Fallback when oracle/_ref is absent: the oracle port (oracle/ref_port.py), post-decode compute."""
    from oracle import ref_port, synth_adapter
    names, tables, per_contig = synth_adapter.tables_and_records(sample)       # untimed: text -> fields
    rows_in = [list(t.items()) for t in tables]
    times = []
    for _ in range(warmup + steps):
        t0 = time.perf_counter()
        built = []
        for items in rows_in:                      # dict insert per kept alignment row (:26-29)
            d = {}
            for nm, tag in items:
                d[nm] = tag
            built.append(d)
        flat = ref_port.join_support_reads(per_contig, built)
        rows = ref_port.phase_records(flat, names, 50, 2)
        times.append(time.perf_counter() - t0)
    times = times[warmup:]
    return {"kind": "port", "seconds": float(np.mean(times)), "best_seconds": float(np.min(times)),
            "n_svs": sample.n_svs, "n_joins": sample.n_joins, "n_tagged": sample.n_tagged, "rows": len(rows),
            "stage_wall_s": None}


def cpu_phase(sample, steps, warmup, full_stage_runs=1) -> dict:
    from oracle import ref_runner
    if ref_runner.available():
        return reference_phase(sample, steps, warmup, full_stage_runs)
    return port_phase(sample, steps, warmup)


def cpu_cluster(n: int, repeats: int = 1) -> dict:
    """Kernel set B on the CPU: oracle/cluster_oracle.py (numpy + scipy connected components) on the first n
    signatures of the C3 workload.  SVIM parity unpinned: this restates the builder's spec, not svim."""
    from duet_b200 import synth
    from oracle import cluster_oracle
    cols = [c[:n] for c in synth.make_signatures(0, C3_N)]
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        _ids, n_clusters = cluster_oracle.cluster(*cols)
        best = min(best, time.perf_counter() - t0)
    return {"seconds": best, "n": n, "n_clusters": int(n_clusters)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "c3":
        n = C3_CPU_SAMPLE
        r = cpu_cluster(n, repeats=max(1, min(args.steps, 3)))
        val = n / r["seconds"]
        desc = (f"{n} of the {C3_N} signatures; oracle/cluster_oracle.py (numpy + scipy connected components) "
                f"on 1 core; SVIM parity unpinned (svim 1.4.2 is external to the reference)")
        line = {"impl": "reference", "metric": METRIC_C3, "value": val, "unit": "signatures/s", "n_gpus": args.gpus,
                "steps": max(1, min(args.steps, 3)), "warmup": 0, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int64/f64 (numpy)", "data": "synthetic",
                "config": {"workload": WORKLOADS["c3"], "sample": desc},
                "cpu_baseline": {"value": val, "unit": "signatures/s", "cores": 1, "kind": "port", "sample": desc},
                "e2e": {"value": val, "unit": "signatures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return
    sample = make_sample(args.workload, 0)
    if isinstance(sample, list):                               # c5: one GPU's share = 4 samples; time one of them
        sample = sample[0]
    r = cpu_phase(sample, args.steps, args.warmup)
    sec = r["seconds"]
    val = r["n_svs"] / sec
    what = ("unmodified reference (oracle/_ref = pip install of /root/reference): generate_phased_callset with its two decode "
            "calls answered from memory" if r["kind"] == "reference" else "oracle port (oracle/_ref not installed)")
    desc = (f"FULL {args.workload} seed 0: {r['n_tagged']} tagged reads, {r['n_svs']} SVs, {r['n_joins']} joins; {what}; "
            f"post-decode compute (join + classify + one-PS sets + predict + sort) on 1 core, every step the whole workload")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "SV/s", "joins_per_sec": r["n_joins"] / sec,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64/f64 (CPython)",
        "data": "synthetic", "config": {"workload": WORKLOADS[args.workload], "sample": desc},
        "cpu_baseline": {"value": val, "unit": "SV/s", "cores": 1, "kind": r["kind"], "sample": desc,
                         "host_cores_available": os.cpu_count(), "rows": r["rows"]},
        "stage_wall_s": r.get("stage_wall_s"), "decode_s": r.get("decode_s"), "read_hap_bam_s": r.get("read_hap_bam_s"),
        "parse_vcf_s": r.get("parse_vcf_s"), "post_decode_s": sec, "input_mb": r.get("input_mb"),
        "stage_note": "stage_wall_s = the unmodified generate_phased_callset, files in -> rows out (SAM text through the samtools "
                      "PATH shim + the SV VCF), one run; decode_s = read_hap_bam + parse_vcf alone; value = S / post_decode_s",
        "e2e": {"value": val, "unit": "SV/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm helpers
# ----------------------------------------------------------------------------------------------
class Ctx:
    """Per-process state shared by the measurements."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local)
        from duet_b200.sharding import bind_to_gpu_numa_node
        self.numa = bind_to_gpu_numa_node(self.local)        # before any page-locked buffer is allocated
        if self.world > 1:
            os.environ.setdefault("NCCL_DEBUG", "WARN")       # a pre-set level (the driver's) is honoured
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dev = f"cuda:{self.local}"
        self.stream = torch.cuda.Stream(device=self.local)
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=self.dev)
        self.clocks = ClockSampler(self.local)
        self.args = args
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        self.peak = float(peaks.get("hbm_gbs", 6650.0))
        self.peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        self.traffic = {}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                self.traffic = json.load(f)
        except Exception:
            pass

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def sum_over_ranks(self, *vals):
        if self.world == 1:
            return [int(v) for v in vals]
        t = self.torch.tensor(list(vals), dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(x) for x in t]


def kernel_bytes(batch, n_hits: int) -> dict:
    """Algorithmic bytes of each kernel of THIS design (kernels): what the step has to move, not
    what the implementation happens to touch (a survivor's slot lookup, the candidate list and the filter
    copies are overhead and are NOT counted)."""
    J, R, S = batch.n_joins, batch.n_reads, batch.n_svs
    return {"init": 20 * 3 * J,              # the slot range of the names (~3 slots of 16 B) + the 4-byte join result
            "build": 28 * J,                 # per name 8 key + 16 slot + 4 filter word
            "probe": 8 * R,                  # the key stream
            "reduce": 24 * J + 64 * S,       # 4 join row + 4 check + 16 tag per join; per-SV in/out
            "tail": 126 * S, "oneps": 12 * S, "predict": 96 * S, "order": 18 * S}


def measure_phase(ctx: Ctx, eng, batch, steps: int, warmup: int, *, e2e: bool = True, after_step=None) -> dict:
    """Device-resident steps (value), per-kernel event times (roofline) and the e2e loop of one batch."""
    from duet_b200.engine import pinned_outputs
    torch = ctx.torch
    stream = ctx.stream
    ksum: dict = {}

    def timed_steps(n, per_kernel=False):
        evs = []
        with torch.cuda.stream(stream):
            for _ in range(n):
                if not ctx.args.no_flush:
                    ctx.flush.zero_()                       # evict the previous step from L2
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                eng.execute(per_kernel)
                if after_step is not None:
                    after_step()
                b.record(stream)
                evs.append((a, b))
                if per_kernel:
                    stream.synchronize()
                    for k, v in eng.timings()["kernel_ms"].items():
                        ksum[k] = ksum.get(k, 0.0) + v
        stream.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    eng.upload(batch)
    eng.sync()
    timed_steps(warmup)
    launches0 = eng.launch_count()
    ctx.barrier()
    with ctx.clocks:
        step_ms = timed_steps(steps)
    ctx.barrier()
    launches = eng.launch_count() - launches0
    res = eng.download()
    out = {"total_ms": float(np.sum(step_ms)), "launches": int(launches), "res": res}
    ksum.clear()
    timed_steps(steps, per_kernel=True)
    out["kernel_ms"] = {k: v / steps for k, v in ksum.items() if v > 0}
    if e2e:
        bufs = pinned_outputs(batch)                  # results land in page-locked host memory too

        def loop(tags_in_place: bool):
            for _ in range(warmup):
                eng.run(batch, buffers=bufs, tags_in_place=tags_in_place)
            ctx.barrier()
            with ctx.clocks:
                t0 = time.perf_counter()
                for _ in range(steps):
                    r = eng.run(batch, buffers=bufs, tags_in_place=tags_in_place)
                sec = time.perf_counter() - t0
            ctx.barrier()
            assert np.array_equal(r.gt, res.gt) and np.array_equal(r.ps, res.ps) and np.array_equal(r.order, res.order)
            return sec, eng.timings(), r

        out["copied_s"], out["copied_tm"], _ = loop(False)
        out["e2e_s"], out["e2e_tm"], r2 = loop(True)
        # cohort mode: the same calls with two in flight (PhasePipeline: sample k+1's upload under sample k's kernels
        # and download); every step still moves its own inputs and results
        from duet_b200.engine import PhasePipeline
        pipe = PhasePipeline(ctx.local, 2)
        try:
            pbufs = [pinned_outputs(batch), pinned_outputs(batch)]
            for _ in pipe.run_many([batch] * max(warmup, 2), buffers=pbufs, tags_in_place=True):
                pass
            ctx.barrier()
            t0 = time.perf_counter()
            for rp in pipe.run_many([batch] * steps, buffers=pbufs, tags_in_place=True):
                pass
            out["pipe_s"] = time.perf_counter() - t0
            ctx.barrier()
            assert np.array_equal(rp.gt, res.gt) and np.array_equal(rp.ps, res.ps) and np.array_equal(rp.order, res.order)
        finally:
            pipe.close()
        out["d2h_bytes"] = int(sum(getattr(r2, k).nbytes for k in ("gt", "ps", "cls", "hap1", "hap2", "hap0", "allhap", "totsc1",
                                                                   "totsc2", "features", "join_row", "shard_counts")) + 4 * batch.n_svs)
        out["h2d_bytes"] = int(batch.input_bytes() - batch.read_tag.nbytes + 32 * res.shard_counts[:, 7].sum())
    return out


def roofline_of(ctx: Ctx, workload: str, batch, kernel_ms: dict, n_hits: int) -> tuple[dict, dict]:
    kb = kernel_bytes(batch, n_hits)
    # dominant kernel = the longest one; kernels within 5 % of the longest count as tied and the tie goes
    # to the one that moves the most bytes (otherwise the reported kernel flips from run to run)
    longest = max(kernel_ms.values())
    dom = max((k for k in kernel_ms if kernel_ms[k] >= 0.95 * longest), key=lambda k: kb[k])
    ach = kb[dom] / (kernel_ms[dom] * 1e-3) / 1e9 if kernel_ms[dom] > 0 else 0.0
    roof = {"bound": "hbm", "kernel": KERNEL_LABEL[dom], "achieved": ach, "peak": ctx.peak, "unit": "GB/s",
            "frac": ach / ctx.peak, "traffic": ctx.traffic.get(workload, {}).get(dom), "peak_kind": ctx.peak_kind,
            "algorithmic_bytes": kb[dom], "kernel_ms": kernel_ms[dom]}
    stages = {k: {"ms": kernel_ms[k], "algorithmic_bytes": kb[k],
                  "GBps": kb[k] / (kernel_ms[k] * 1e-3) / 1e9, "frac": kb[k] / (kernel_ms[k] * 1e-3) / 1e9 / ctx.peak}
              for k in kernel_ms}
    return roof, stages


def phase_config_result(ctx: Ctx, eng, workload: str, steps: int, warmup: int, seed: int = 0) -> dict:
    """One phasing config measured on this GPU alone -> the compact record of `configs`."""
    from duet_b200.columnar import from_synth
    from duet_b200.engine import pin_batch
    batch = pin_batch(from_synth(make_sample(workload, seed), with_text=False))
    m = measure_phase(ctx, eng, batch, steps, warmup)
    ms = m["total_ms"] / steps
    alg = batch.algorithmic_bytes()
    roof, stages = roofline_of(ctx, workload, batch, m["kernel_ms"], int(m["res"].shard_counts[:, 7].sum()))
    return {"workload": WORKLOADS[workload], "value": batch.n_svs / (ms * 1e-3), "unit": "SV/s",
            "joins_per_sec": batch.n_joins / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "reads_tagged": batch.n_reads, "svs": batch.n_svs, "joins": batch.n_joins, "shards": batch.n_shards,
            "gpu_launches": m["launches"],
            "e2e": {"value": batch.n_svs / (m["e2e_s"] / steps), "unit": "SV/s", "ms_per_step": m["e2e_s"] / steps * 1e3,
                    "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": m["d2h_bytes"],
                    "two_calls_in_flight": {"value": batch.n_svs / (m["pipe_s"] / steps), "ms_per_step": m["pipe_s"] / steps * 1e3}},
            "roofline": roof, "kernel_ms": m["kernel_ms"],
            "path_roofline": {"algorithmic_bytes_8d": alg["total"], "achieved": alg["total"] / (ms * 1e-3) / 1e9,
                              "frac": alg["total"] / (ms * 1e-3) / 1e9 / ctx.peak},
            "n_emitted": int(m["res"].shard_counts[:, 2].sum())}


def cluster_config_result(ctx: Ctx, eng, steps: int, warmup: int, with_cpu: bool) -> dict:
    """C3: span-position-distance clustering of 2 M signatures (kernel set B)."""
    import ctypes as C
    from duet_b200 import _lib, synth
    from duet_b200.engine import pinned_empty
    torch = ctx.torch
    cols = synth.make_signatures(0, C3_N)
    n = int(cols[0].shape[0])
    dev = [torch.from_numpy(c).to(ctx.dev) for c in cols]
    out_dev = torch.empty(n, dtype=torch.int32, device=ctx.dev)
    par = _lib.ClusterParams()
    eng.lib.duet_default_cluster_params(C.byref(par))
    nclu, ms_c = C.c_int64(), C.c_float()

    def run(mem, ptrs, out_ptr):
        inp = _lib.ClusterInput()
        inp.mem, inp.n = mem, n
        inp.contig, inp.type, inp.start, inp.end = ptrs
        rc = eng.lib.duet_cluster_run(eng.h, C.byref(inp), C.byref(par), out_ptr, C.byref(nclu), C.byref(ms_c))
        if rc != _lib.DUET_OK:
            raise RuntimeError(eng.lib.duet_last_error(eng.h).decode())
        return float(ms_c.value)

    dptr = [int(t.data_ptr()) for t in dev]
    launches0 = eng.launch_count()
    with torch.cuda.stream(ctx.stream):
        for _ in range(warmup):
            run(_lib.MEM_DEVICE, dptr, int(out_dev.data_ptr()))
        launches1 = eng.launch_count()
        dev_ms = []
        with ctx.clocks:
            for _ in range(steps):
                if not ctx.args.no_flush:
                    ctx.flush.zero_()
                    ctx.stream.synchronize()
                dev_ms.append(run(_lib.MEM_DEVICE, dptr, int(out_dev.data_ptr())))    # library events on the launch stream
    launches = (eng.launch_count() - launches1) // max(steps, 1)
    ids_dev = out_dev.cpu().numpy()
    kms = cluster_kernel_ms(eng)
    # e2e: page-locked host columns in, cluster ids out on the host
    hcols = []
    for c in cols:
        h = pinned_empty(c.shape, c.dtype)
        h[...] = c
        hcols.append(h)
    hout = pinned_empty(n, np.int32)
    hptr = [h.ctypes.data for h in hcols]
    for _ in range(warmup):
        run(_lib.MEM_HOST, hptr, hout.ctypes.data)
    t0 = time.perf_counter()
    for _ in range(steps):
        run(_lib.MEM_HOST, hptr, hout.ctypes.data)
    e2e_s = (time.perf_counter() - t0) / steps
    assert np.array_equal(hout, ids_dev)
    ms = float(np.mean(dev_ms))
    alg = 20 * n
    rec = {"workload": WORKLOADS["c3"], "metric": METRIC_C3, "value": n / (ms * 1e-3), "unit": "signatures/s",
           "ms_per_step": ms, "steps": steps, "warmup": warmup, "signatures": n, "clusters": int(nclu.value),
           "gpu_launches": int(launches) * steps,
           "e2e": {"value": n / e2e_s, "unit": "signatures/s", "ms_per_step": e2e_s * 1e3,
                   "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 4 * n},
           "path_roofline": {"algorithmic_bytes_8d": alg, "achieved": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / ctx.peak,
                             "note": "SURVEY.md 8(d): 16 B in + 4 B out per signature over the whole device path"},
           "parity": "bit-exact vs oracle/cluster_oracle.py (the builder's spec); SVIM parity unpinned -- svim 1.4.2 is external "
                     "to the reference and clusters by average linkage, not connected components"}
    if kms:
        dom = max(kms, key=kms.get)
        kb = cluster_kernel_bytes(n)
        rec["kernel_ms"] = kms
        rec["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": kb.get(dom, alg) / (kms[dom] * 1e-3) / 1e9, "peak": ctx.peak,
                           "unit": "GB/s", "frac": kb.get(dom, alg) / (kms[dom] * 1e-3) / 1e9 / ctx.peak,
                           "traffic": ctx.traffic.get("c3", {}).get(dom), "peak_kind": ctx.peak_kind,
                           "algorithmic_bytes": kb.get(dom, alg), "kernel_ms": kms[dom]}
    else:
        rec["roofline"] = {"bound": "hbm", "kernel": "whole call", "achieved": alg / (ms * 1e-3) / 1e9, "peak": ctx.peak, "unit": "GB/s",
                           "frac": alg / (ms * 1e-3) / 1e9 / ctx.peak, "traffic": None, "peak_kind": ctx.peak_kind,
                           "algorithmic_bytes": alg, "kernel_ms": ms}
    if with_cpu:
        r = cpu_cluster(C3_CPU_SAMPLE)
        rec["cpu_baseline"] = {"value": r["n"] / r["seconds"], "unit": "signatures/s", "cores": 1, "kind": "port",
                               "seconds": r["seconds"],
                               "sample": f"{r["n"]} of the {n} signatures; oracle/cluster_oracle.py (numpy + scipy) on 1 core; "
                                         "SVIM parity unpinned"}
    return rec


def cluster_kernel_ms(eng) -> dict:
    """Per-kernel event times of the last duet_cluster_run, when the library exposes them."""
    import ctypes as C
    fn = getattr(eng.lib, "duet_cluster_timings", None)
    if fn is None:
        return {}
    names = (C.c_char_p * 16)()
    ms = (C.c_float * 16)()
    k = fn(eng.h, names, ms, 16)
    return {names[i].decode(): float(ms[i]) for i in range(max(k, 0)) if ms[i] > 0}


def cluster_kernel_bytes(n: int) -> dict:
    """Algorithmic bytes per stage of kernel set B: keys 16 B in + 16 B out; a sort
    pass moves 16 B in + 16 B out per signature and reads the keys once more for the histogram (five passes for
    a human genome: 36 key bits); k_cl_runs reads 12 B and writes the 8 B of forest state, the edge kernel reads 12 B + the 4-byte forest; labels 12 B + 8 B."""
    return {"k_cl_max + k_cl_hist + k_cl_scatter": (16 + 12) * n + 8 * n + (12 + 16) * n, "k_cl_bucket": 16 * n + 4 * n, "k_cl_fix": 4 * n,
            "k_cl_keys": 32 * n, "k_rs_hist + k_rs_scan + k_rs_scatter": 5 * 40 * n, "k_cl_runs + k_cl_edges": 20 * n + 16 * n,
            "k_cl_label + k_cl_write": 24 * n}


def strong_scaling(ctx: Ctx, eng, steps: int, warmup: int) -> dict:
    """ONE C4 sample, contigs LPT-packed over the ranks; every rank phases only its contigs; the step ends
    with the NCCL all-gather of the per-contig counters.  Each rank also runs the unsharded batch once and
    compares its slice."""
    from duet_b200 import _lib, sharding
    from duet_b200.columnar import from_synth
    from duet_b200.engine import pin_batch
    torch, dist = ctx.torch, ctx.dist
    full = from_synth(make_sample("c4", 0), with_text=False)
    plan = sharding.lpt_assign(sharding.shard_weights(full), ctx.world)
    mine = plan[ctx.rank]
    sub = pin_batch(full.select_shards(mine)) if mine else None
    width = max(1, max(len(p) for p in plan))
    counts_dev = torch.zeros((width, _lib.N_COUNTERS), dtype=torch.int64, device=ctx.dev)
    gathered = [torch.empty_like(counts_dev) for _ in range(ctx.world)]
    # unsharded result on this GPU, for the comparison
    eng.upload(pin_batch(full))
    eng.execute()
    ref = eng.download()
    ok = True
    res = None
    if sub is not None:
        eng.upload(sub)
        eng.execute()
        res = eng.download()
        s_of = np.concatenate([np.arange(full.sv_off[s], full.sv_off[s + 1]) for s in mine]) if mine else np.zeros(0, np.int64)
        ok = bool(np.array_equal(res.gt, ref.gt[s_of]) and np.array_equal(res.ps, ref.ps[s_of]) and
                  np.array_equal(res.shard_counts, ref.shard_counts[mine]))
        counts_dev[:len(mine)] = torch.from_numpy(res.shard_counts).to(ctx.dev)

    def step():
        if sub is not None:
            eng.execute()
        if ctx.world > 1:
            with torch.cuda.stream(ctx.stream):
                dist.all_gather(gathered, counts_dev)          # the path's only collective: inside the timed call

    evs = []
    with torch.cuda.stream(ctx.stream):
        for _ in range(warmup):
            ctx.flush.zero_()
            step()
    ctx.barrier()
    with torch.cuda.stream(ctx.stream):
        for _ in range(steps):
            ctx.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(ctx.stream)
            step()
            b.record(ctx.stream)
            evs.append((a, b))
    ctx.stream.synchronize()
    ctx.barrier()
    total_ms = float(sum(a.elapsed_time(b) for a, b in evs))
    (total_ms,) = ctx.max_over_ranks(total_ms)
    (bad,) = ctx.sum_over_ranks(0 if ok else 1)
    ms = total_ms / steps
    loads = [int(sum(sharding.shard_weights(full)[i] for i in p)) for p in plan]
    n_emit = int(sum(int(g[:len(p), 2].sum()) for g, p in zip(gathered, plan))) if ctx.world > 1 else int(ref.shard_counts[:, 2].sum())
    return {"scaling": "strong", "workload": WORKLOADS["c4"] + f", ONE sample, its 24 contigs LPT-packed over {ctx.world} GPU(s)",
            "value": full.n_svs / (ms * 1e-3), "unit": "SV/s", "joins_per_sec": full.n_joins / (ms * 1e-3), "ms_per_step": ms,
            "steps": steps, "warmup": warmup, "svs": full.n_svs, "joins": full.n_joins, "reads_tagged": full.n_reads,
            "contigs_per_rank": [len(p) for p in plan], "load_share_max": max(loads) / max(sum(loads), 1),
            "slices_equal_unsharded": bad == 0, "n_emitted_gathered": n_emit,
            "n_emitted_unsharded": int(ref.shard_counts[:, 2].sum()),
            "collective": "dist.all_gather of 8 int64 counters per contig (NCCL), inside the timed call",
            "bounds": "speed-up is capped by the largest bin of the LPT plan (load_share_max; chr1 alone is 8 % of the genome) "
                      "and by the launch floor of the chain, which does not shrink with the shard"}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other configs' sub-results (N=1 headline only)")
    ap.add_argument("--no-stage-wall", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic only: keep L2 warm between steps (not a valid bench line)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
        return

    from duet_b200.columnar import from_synth
    from duet_b200.engine import PhaseEngine, pin_batch

    ctx = Ctx(args)
    rank, world = ctx.rank, ctx.world
    eng = PhaseEngine(ctx.local)
    eng.set_stream(ctx.stream.cuda_stream)
    eng.set_thresholds(50, 2)
    sub_steps, sub_warm = max(3, min(args.steps, 10)), 3

    if args.workload == "c3":
        rec = cluster_config_result(ctx, eng, args.steps, args.warmup, with_cpu=not args.no_cpu_baseline and rank == 0)
        if rank == 0:
            line = {"metric": rec.pop("metric"), "value": rec["value"] * world, "unit": rec["unit"], "n_gpus": world,
                    "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "u64/i32/f64", "data": "synthetic",
                    "config": {"workload": WORKLOADS["c3"] + (f" x {world} replicas (one per GPU)" if world > 1 else ""),
                               "l2": "flushed between steps (512 MiB memset outside the timed region)"},
                    "clocks": ctx.clocks.summary(), **{k: v for k, v in rec.items() if k not in ("value", "unit", "ms_per_step", "steps", "warmup", "workload")}}
            print(json.dumps(line))
        if world > 1:
            ctx.dist.barrier()
            ctx.dist.destroy_process_group()
        return

    t0 = time.perf_counter()
    sample = make_sample(args.workload, seed=rank)
    batch = pin_batch(from_synth(sample, with_text=False))
    gen_s = time.perf_counter() - t0
    m = measure_phase(ctx, eng, batch, args.steps, args.warmup)
    res = m["res"]
    total_ms, e2e_s, copied_s, pipe_s = ctx.max_over_ranks(m["total_ms"], m["e2e_s"], m["copied_s"], m["pipe_s"])
    per_rank = None
    if world > 1:                                  # where the e2e time of a multi-GPU run goes: every rank's own figures
        tm_r = m["e2e_tm"]
        torch = ctx.torch
        mine = torch.tensor([m["e2e_s"] / args.steps * 1e3, tm_r["h2d_ms"], tm_r["device_ms"], tm_r["d2h_ms"]], dtype=torch.float64, device=ctx.dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        ctx.dist.all_gather(allr, mine)
        per_rank = {"call_ms": [round(float(t[0]), 3) for t in allr], "h2d_ms_last_step": [round(float(t[1]), 3) for t in allr],
                    "device_ms_last_step": [round(float(t[2]), 3) for t in allr], "d2h_ms_last_step": [round(float(t[3]), 3) for t in allr],
                    "note": "wall clock per call of each rank's e2e loop, and the event-timed copies / kernels of its last step"}
    n_svs, n_joins, n_reads = ctx.sum_over_ranks(batch.n_svs, batch.n_joins, batch.n_reads)

    # counters gathered once (not on the weak-scaling data path: shards never exchange data)
    gather_ms = None
    if world > 1:
        torch, dist = ctx.torch, ctx.dist
        counts = torch.from_numpy(res.shard_counts.sum(axis=0)).to(ctx.dev)
        outl = [torch.empty_like(counts) for _ in range(world)]
        dist.all_gather(outl, counts)                        # warm
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dist.all_gather(outl, counts)
        b.record()
        torch.cuda.synchronize()
        gather_ms = a.elapsed_time(b)
        all_counts = torch.stack(outl).sum(0).tolist()
    else:
        all_counts = res.shard_counts.sum(axis=0).tolist()

    strong = strong_scaling(ctx, eng, max(5, min(args.steps, 20)), 3) if world > 1 else None

    configs = {}
    if world == 1 and args.workload == "c2" and not args.no_configs:
        for w in ("c1", "c4", "c5"):
            configs[w] = phase_config_result(ctx, eng, w, sub_steps, sub_warm)
        configs["c3"] = cluster_config_result(ctx, eng, sub_steps, sub_warm, with_cpu=not args.no_cpu_baseline)

    if rank == 0:
        ms = total_ms / args.steps
        alg = batch.algorithmic_bytes()
        roof, stages = roofline_of(ctx, args.workload, batch, m["kernel_ms"], int(res.shard_counts[:, 7].sum()))
        roof["note"] = ("dominant = longest kernel (ties within 5 % go to the one moving most bytes); every kernel's own figure is under "
                        "`stages`, the whole path's under `path_roofline`")
        tm, copied_tm = m["e2e_tm"], m["copied_tm"]
        line = {
            "metric": METRIC, "value": n_svs / (ms * 1e-3), "unit": "SV/s",
            "joins_per_sec": n_joins / (ms * 1e-3), "reads_per_sec": n_reads / (ms * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/i32/f64",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload] + (f" x {world} samples (one per GPU)" if world > 1 else ""),
                       "reads_tagged_per_gpu": batch.n_reads, "svs_per_gpu": batch.n_svs,
                       "joins_per_gpu": batch.n_joins, "shards_per_gpu": batch.n_shards,
                       "parallelism": f"shard=(sample,contig); {world} GPU(s), no data-path collective",
                       "l2": "NOT FLUSHED (diagnostic run, invalid as a bench line)" if args.no_flush else
                             "flushed between steps (512 MiB memset outside the event pair)",
                       "thresholds": "svlen>=50, support>=2 (reference defaults)"},
            "clocks": None,
            "e2e": {"value": n_svs / (e2e_s / args.steps), "unit": "SV/s", "ms_per_step": e2e_s / args.steps * 1e3,
                    "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": m["d2h_bytes"],
                    "mode": "PhaseEngine.run(tags_in_place=True): all columns copied from page-locked memory except the "
                            "16-byte tag records, which k_reduce gathers over the bus (one 32-byte sector per joined "
                            "read, counted in h2d_bytes_per_step)",
                    "last_step": {k: tm[k] for k in ("h2d_ms", "device_ms", "d2h_ms")},
                    "per_rank": per_rank,
                    "two_calls_in_flight": {"value": n_svs / (pipe_s / args.steps), "ms_per_step": pipe_s / args.steps * 1e3,
                                            "mode": "duet_b200.engine.PhasePipeline(depth=2): the same calls, sample k+1's upload and "
                                                    "kernels enqueued before sample k's results are waited for (cohort mode); every "
                                                    "step moves its own inputs and results"},
                    "all_columns_copied": {"value": n_svs / (copied_s / args.steps), "ms_per_step": copied_s / args.steps * 1e3,
                                           "h2d_bytes_per_step": batch.input_bytes(),
                                           "last_step": {k: copied_tm[k] for k in ("h2d_ms", "device_ms", "d2h_ms")}}},
            "gpu_launches": m["launches"],
            "roofline": roof, "kernel_ms": m["kernel_ms"], "stages": stages,
            "path_roofline": {"algorithmic_bytes_8d": alg["total"], "device_ms": ms,
                              "achieved": alg["total"] * world / (ms * 1e-3) / 1e9 / world,
                              "frac": alg["total"] / (ms * 1e-3) / 1e9 / ctx.peak,
                              "note": "SURVEY.md 8(d): 33 B/tagged read + 24 B/join + 64 B/SV over the whole device path, per GPU"},
            "counters": dict(zip(("n_sv", "n_kept", "n_emitted", "n_1|0", "n_0|1", "n_1|1", "n_joins", "n_hits"),
                                 [int(x) for x in all_counts])),
            "gather_ms": gather_ms, "synth_seconds": gen_s, "numa": ctx.numa,
        }
        if strong is not None:
            line["strong"] = strong
            if args.scaling == "strong":
                line["weak"] = {k: line[k] for k in ("value", "joins_per_sec", "ms_per_step", "e2e", "config")}
                line.update({"scaling": "strong", "value": strong["value"], "joins_per_sec": strong["joins_per_sec"],
                             "ms_per_step": strong["ms_per_step"], "steps": strong["steps"], "warmup": strong["warmup"]})
                line["config"] = {"workload": strong["workload"], "parallelism": "contig shards LPT-packed over the GPUs; NCCL counter "
                                  "all-gather inside the timed call", "l2": "flushed between steps"}
        if configs:
            line["configs"] = configs
        if world == 1 and not args.no_stage_wall and args.workload in ("c1", "c2"):
            line.update(stage_wall(sample, eng))
        if not args.no_cpu_baseline and args.workload != "c5":
            sub = sub_sample(make_sample(args.workload, 0), CPU_SAMPLE_CONTIGS)
            r = cpu_phase(sub, 3, 1)
            desc = (f"contigs {','.join(c.name for c in sub.contigs)} of {args.workload} seed 0: {r['n_tagged']} tagged reads, {r['n_svs']} SVs, "
                    f"{r['n_joins']} joins; " + ("the unmodified reference (oracle/_ref)" if r["kind"] == "reference" else "oracle port")
                    + " on 1 core, post-decode compute (join + classify + predict + sort), mean of 3")
            line["cpu_baseline"] = {"value": r["n_svs"] / r["seconds"], "unit": "SV/s", "cores": 1, "kind": r["kind"],
                                    "joins_per_sec": r["n_joins"] / r["seconds"], "seconds": r["seconds"],
                                    "host_cores_available": os.cpu_count(), "sample": desc,
                                    "stage_wall_s": r.get("stage_wall_s"), "decode_s": r.get("decode_s")}
        line["clocks"] = ctx.clocks.summary()
        print(json.dumps(line))
    if world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def stage_wall(sample, eng) -> dict:
    """The whole drop-in stage -- duet_b200.sv_phasing.sv_phasing(home, ...), files in -> phased_sv.vcf out -- on the
    workload written out as the files the reference reads.  Host decode is reported apart from the device call."""
    import shutil
    import tempfile
    from duet_b200 import sv_phasing, sv_phasing_fn, synth
    home = tempfile.mkdtemp(prefix="duet_stage_")
    try:
        synth.write_workdir(sample, home)
        sv_phasing_fn._ENGINES.setdefault(eng.device, eng)          # context creation is not part of the stage
        threads = min(8, os.cpu_count() or 1)
        runs = []
        for _ in range(3):
            t0 = time.perf_counter()
            sv_phasing.sv_phasing(home, 50, 2, threads, False)
            runs.append((time.perf_counter() - t0, dict(sv_phasing_fn.last_timings)))
        best, tm = min(runs, key=lambda r: r[0])
        return {"stage_wall_s": best, "host_decode_s": tm["host_decode_s"], "device_call_s": tm["device_call_s"],
                "rows_s": tm["rows_s"], "stage_threads": threads,
                "stage_note": "duet_b200.sv_phasing.sv_phasing on the workload's files (per-contig SAM text + cuteSV VCF) -> phased_sv.vcf, "
                              "best of 3; the reference arm reports the same stage as its stage_wall_s"}
    finally:
        shutil.rmtree(home, ignore_errors=True)


if __name__ == "__main__":
    main()
