"""Kernel set B host API: span-position-distance clustering of SV signatures on the GPU.

In the reference this step happens inside the external `svim` process that
/root/reference/src/duet/sv_calling.py:14-15 launches with `--cluster_max_distance <c>`
(`-c`, default 0.9: utils.py:27-28, README.md:63).  svim is not part of the reference tree, so what
is implemented is the spec written in csrc/cluster_kernels.cuh (connected components of the
thresholded span-position distance inside a partition window); parity with svim 1.4.2 is unpinned.
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
