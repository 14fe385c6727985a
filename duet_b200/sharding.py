"""Multi-GPU execution: one process per GPU, shards (contigs, or (sample, contig) pairs)
distributed over the ranks, NO collective on the data path.

Every dict, set and loop of the reference is per contig
(/root/reference/src/duet/sv_phasing_fn.py:15-18,195-212) and the final sort key starts with the
CHROM string (:229), so a rank can decode, phase and format its contigs alone.  The only
exchange is one small all-gather of per-shard counters (8 x int64 per shard), which gives
every rank the `Duet.<idx>` offset of its row slices (write_file.py:10-16 numbers rows
1..N in output order).  torch.distributed carries it: NCCL on the GPUs, gloo in CPU tests.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (and so, by first touch, its page-locked
    buffers to that node's memory).  With one process per GPU and nothing bound, eight ranks' uploads and in-place
    tag gathers cross the socket interconnect and contend for the same root complexes (round 1: e2e efficiency
    0.59 at 8 GPUs).  Best effort: returns what it did; never raises."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                     # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed))
    except Exception as e:                                  # noqa: BLE001 -- a machine without sysfs / nvml: leave it
        info["error"] = type(e).__name__
    return info


def lpt_assign(weights, n_bins: int) -> list[list[int]]:
    """Longest-processing-time-first bin packing; deterministic on every rank.  Items of equal load
    (in particular the zero-weight ones: contigs without a haplotagged BAM) go to the bin holding the
    fewest items, so they spread over the ranks instead of piling up on the first idle one."""
    order = sorted(range(len(weights)), key=lambda i: (-int(weights[i]), i))
    loads = [0] * n_bins
    bins: list[list[int]] = [[] for _ in range(n_bins)]
    for i in order:
        k = min(range(n_bins), key=lambda b: (loads[b], len(bins[b]), b))
        bins[k].append(i)
        loads[k] += int(weights[i])
    return [sorted(b) for b in bins]


def shard_weights(batch) -> np.ndarray:
    """Device work of a shard ~ streamed reads + gathered joins (+ per-SV work)."""
    reads = np.diff(batch.read_off)
    svs = np.diff(batch.sv_off)
    joins = batch.csr_off[batch.sv_off[1:]] - batch.csr_off[batch.sv_off[:-1]]
    return reads + 4 * joins + 8 * svs


def gather_counters(local_ids, local_counts: np.ndarray, plan: list[list[int]], device=None) -> np.ndarray:
    """All-gather the per-shard counter rows.  `plan[r]` lists rank r's shards, `local_counts[k]`
    belongs to shard `local_ids[k]`.  Returns [n_shards_total, N_COUNTERS] on every rank.
    One collective of max_k x 8 int64 per rank; with world size 1 nothing is sent."""
    import torch
    import torch.distributed as dist
    n_total = sum(len(p) for p in plan)
    out = np.zeros((n_total, _lib.N_COUNTERS), np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        for k, s in enumerate(local_ids):
            out[s] = local_counts[k]
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    assert list(local_ids) == plan[rank]
    width = max(1, max(len(p) for p in plan))
    buf = torch.zeros((width, _lib.N_COUNTERS), dtype=torch.int64)
    if len(local_ids):
        buf[:len(local_ids)] = torch.from_numpy(np.ascontiguousarray(local_counts, dtype=np.int64))
    if device is not None:
        buf = buf.to(device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    for r, p in enumerate(parts):
        arr = p.cpu().numpy()
        for k, s in enumerate(plan[r]):
            out[s] = arr[k]
    return out


def slice_first_ids(shard_keys: list, n_emitted) -> list[int] | None:
    """First `Duet.<idx>` of each shard's row slice when the output is the shards' slices laid end
    to end in CHROM-string order.  `shard_keys[s]` = the set of CHROM strings of shard s's emitted
    rows.  Returns None when two shards share a CHROM string or a shard has two (then slices
    interleave and rank 0 has to merge -- only possible with exotic contig lists)."""
    seen = {}
    for s, ks in enumerate(shard_keys):
        ks = sorted(ks)
        if len(ks) > 1:
            return None
        for k in ks:
            if k in seen:
                return None
            seen[k] = s
    first = [0] * len(shard_keys)
    nxt = 1
    for k in sorted(seen):
        first[seen[k]] = nxt
        nxt += int(n_emitted[seen[k]])
    return first


def bam_weights(sam_home: str, chrom_list) -> list[int]:
    """Decode cost of a contig before anything is read: the size of its haplotagged BAM."""
    out = []
    for ctg in chrom_list:
        size = 0
        for cand in (sam_home + "chr" + ctg + ".bam", sam_home + ctg + ".bam"):
            if os.path.exists(cand):
                size = os.path.getsize(cand)
                break
        out.append(size)
    return out


def sv_phasing_sharded(home, svlen_thres, suppread_thres, thread, include_all_ctgs, *, phase_fn=None,
                       rank: int | None = None, world: int | None = None, device=None):
    """Contig-sharded version of sv_phasing.sv_phasing(): every rank decodes and phases only its
    contigs and writes its row slices; rank 0 writes the header and stitches the slices into
    <home>/phased_sv.vcf, byte-identical to the single-process output.  Returns the global
    per-contig counter table.  `phase_fn(batch, svlen_thres, suppread_thres)` defaults to the
    device path (tests inject a stand-in to exercise the plumbing without a GPU)."""
    import torch.distributed as dist
    from . import sv_phasing_fn as fn
    from .read_file import init_chrom_list, parse_vcf
    from .write_file import format_rows, header_text

    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
    phase_fn = phase_fn or fn.phase_batch
    vcf_path, sam_home, out_path = home + "/sv_calling/variants.vcf", home + "/snp_phasing/", home + "/phased_sv.vcf"
    chrom_list = init_chrom_list(include_all_ctgs, home)
    plan = lpt_assign(bam_weights(sam_home, chrom_list), world)
    mine = plan[rank]

    comp_call = parse_vcf(vcf_path, include_all_ctgs)                  # small; every rank reads it
    sources = [next((p for p in (sam_home + "chr" + c + ".bam", sam_home + c + ".bam") if os.path.exists(p)), None)
               for c in chrom_list]
    probe = [fn.ReadColumns.empty() for _ in chrom_list]               # only the file identities, for the
    for cols, p in zip(probe, sources):                                # 'c' and 'chr'+c both listed check
        cols.source = p
    comp_call = fn.contig_records(chrom_list, comp_call, probe)
    read_hap = []
    for ch in mine:                                                    # the heavy decode: own contigs only
        cols = fn.load_hap_bam(sources[ch], thread) if sources[ch] else fn.ReadColumns.empty()
        cols.source = sources[ch]
        read_hap.append(cols)
    rows_by_shard = {}
    if mine:
        batch = fn.build_batch([chrom_list[ch] for ch in mine], read_hap, [comp_call[ch] for ch in mine])
        res = phase_fn(batch, svlen_thres, suppread_thres)
        local_counts = res.shard_counts
        shard_of = np.searchsorted(batch.sv_off, res.order, side="right") - 1
        for k, ch in enumerate(mine):
            sub = type(res)(**{**res.__dict__, "order": res.order[shard_of == k]})
            rows_by_shard[ch] = sub.rows(batch)
    else:
        # more ranks than contigs: this rank owns nothing, but it still takes part in every collective
        local_counts = np.zeros((0, _lib.N_COUNTERS), np.int64)
    counts = gather_counters(mine, local_counts, plan, device)
    keys_local = {ch: sorted({r["chrom"] for r in rows}) for ch, rows in rows_by_shard.items()}
    if world > 1:
        all_keys = [None] * world
        dist.all_gather_object(all_keys, keys_local)
        keys = {}
        for d in all_keys:
            keys.update(d)
    else:
        keys = keys_local
    first = slice_first_ids([keys.get(ch, []) for ch in range(len(chrom_list))], counts[:, 2])
    if first is not None:
        for ch, rows in rows_by_shard.items():
            if rows:
                with open(f"{out_path}.slice.{ch}", "w") as f:
                    f.write(format_rows(rows, first[ch]))
        if world > 1:
            dist.barrier()
        if rank == 0:
            order = sorted((keys[ch][0], ch) for ch in keys if keys[ch])
            with open(out_path, "w") as out:
                out.write(header_text(vcf_path, include_all_ctgs))
                for _, ch in order:
                    with open(f"{out_path}.slice.{ch}") as f:
                        out.write(f.read())
                    os.remove(f"{out_path}.slice.{ch}")
    else:                                                              # slices interleave: merge on rank 0
        gathered = [None] * world
        if world > 1:
            dist.gather_object(rows_by_shard, gathered if rank == 0 else None, dst=0)
        else:
            gathered = [rows_by_shard]
        if rank == 0:
            rows = []
            for ch in range(len(chrom_list)):
                for d in gathered:
                    rows += d.get(ch, [])
            rows.sort(key=lambda r: (r["chrom"], r["pos"]))
            with open(out_path, "w") as out:
                out.write(header_text(vcf_path, include_all_ctgs))
                out.write(format_rows(rows))
    if world > 1:
        dist.barrier()
    return counts
