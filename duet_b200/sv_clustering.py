"""Kernel set B host API: span-position-distance clustering of SV signatures on the GPU.

In the reference this step happens inside the external `svim` process that
/root/reference/src/duet/sv_calling.py:14-15 launches with `--cluster_max_distance <c>`
(`-c`, default 0.9: utils.py:27-28, README.md:63).  svim is not part of the reference tree, so what
is implemented is the spec written in csrc/cluster_kernels.cuh (connected components of the
thresholded span-position distance inside a partition window).

NOT A DROP-IN FOR SVIM.  svim 1.4.2 cuts position partitions and runs average-linkage hierarchical
clustering (scipy `linkage('average')` + `fcluster`) inside them; this module builds single-linkage
connected components (union-find), as BASELINE.json's north_star prescribes.  Single linkage chains,
so for the same threshold it yields fewer and larger clusters.  Parity with svim is unpinned: the
results are bit-exact against oracle/cluster_oracle.py, which restates THIS spec, not svim.  Do not
wire it into sv_calling in place of svim without pinning it against real svim output first.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import DuetError, PhaseEngine

SIGNATURE_TYPES = ("DEL", "INS", "INV", "DUP_TAN", "DUP_INT", "BND")


def cluster_signatures(contig, sig_type, start, end, cluster_max_distance: float = 0.9, *,
                       position_normalizer: float = 900.0, partition_window: int = 1000,
                       engine: PhaseEngine | None = None):
    """contig / sig_type / start / end: int arrays of equal length (insertions: end = start + length).
    Returns (cluster_id int32[n] -- the smallest signature index of each signature's cluster --,
    n_clusters, device_ms).  Raises without a GPU."""
    from .sv_phasing_fn import get_engine
    eng = engine or get_engine()
    cols = [np.ascontiguousarray(x, dtype=np.int32) for x in (contig, sig_type, start, end)]
    n = cols[0].shape[0]
    if any(c.shape[0] != n for c in cols):
        raise ValueError("columns differ in length")
    inp = _lib.ClusterInput()
    inp.mem, inp.n = _lib.MEM_HOST, n
    inp.contig, inp.type, inp.start, inp.end = (c.ctypes.data for c in cols)
    par = _lib.ClusterParams()
    eng.lib.duet_default_cluster_params(C.byref(par))
    par.max_distance, par.position_normalizer, par.partition_window = cluster_max_distance, position_normalizer, partition_window
    out = np.empty(n, np.int32)
    n_clusters, ms = C.c_int64(), C.c_float()
    rc = eng.lib.duet_cluster_run(eng.h, C.byref(inp), C.byref(par), out.ctypes.data, C.byref(n_clusters), C.byref(ms))
    if rc != _lib.DUET_OK:
        raise DuetError(rc, eng.lib.duet_last_error(eng.h).decode())
    return out, int(n_clusters.value), float(ms.value)


def cluster_signatures_sharded(contig, sig_type, start, end, cluster_max_distance: float = 0.9, *,
                               position_normalizer: float = 900.0, partition_window: int = 1000,
                               rank: int | None = None, world: int | None = None, cluster_fn=None, device=None):
    """Kernel set B over several GPUs: signatures only ever interact inside their (contig, type) group, so the
    groups are LPT-packed over the ranks by size (sharding.lpt_assign) and every rank clusters its own groups
    -- no exchange on the data path.  Cluster ids are ORIGINAL signature indices, so the ranks' results need no
    renumbering; the one collective is the all-reduce of the per-rank cluster counts.

    Returns (index, cluster_id, n_clusters_total): `index` = the original indices this rank owned (ascending),
    `cluster_id[k]` = the cluster of signature index[k].  `cluster_fn(contig, type, start, end, ...)` defaults to
    the device path (tests inject a stand-in to exercise the plumbing without a GPU)."""
    import torch.distributed as dist
    from .sharding import lpt_assign
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
        world = dist.get_world_size() if dist.is_initialized() else 1
    cluster_fn = cluster_fn or cluster_signatures
    contig, sig_type = np.asarray(contig, np.int64), np.asarray(sig_type, np.int64)
    start, end = np.asarray(start), np.asarray(end)
    group = contig * 256 + sig_type
    keys, inverse, counts = np.unique(group, return_inverse=True, return_counts=True)
    plan = lpt_assign(counts.tolist(), world)
    mine = np.zeros(keys.shape[0], bool)
    mine[plan[rank]] = True
    index = np.nonzero(mine[inverse])[0]
    if index.shape[0]:
        ids, n_local, _ = cluster_fn(contig[index], sig_type[index], start[index], end[index], cluster_max_distance,
                                     position_normalizer=position_normalizer, partition_window=partition_window)
        cluster_id = index[np.asarray(ids, np.int64)].astype(np.int32)      # local smallest member -> its original index
    else:
        cluster_id, n_local = np.zeros(0, np.int32), 0
    total = int(n_local)
    if world > 1:
        import torch
        t = torch.tensor([int(n_local)], dtype=torch.int64, device=device) if device is not None else torch.tensor([int(n_local)], dtype=torch.int64)
        dist.all_reduce(t)
        total = int(t.item())
    return index, cluster_id, total
