"""Host driver of the device path: one `PhaseEngine` per GPU wraps one C-ABI handle
(include/duet_b200.h).  All compute happens in libduet_b200.so; nothing here falls back
to the CPU -- without the library or a GPU the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import DuetError
from .columnar import PhaseBatch

GT_TEXT = {1: "1|0", 2: "0|1", 3: "1|1"}   # sv_phasing_fn.py:217-222


def _ptr(a):
    return None if a is None else a.ctypes.data


@dataclass
class PhaseResult:
    gt: np.ndarray          # uint8 [S]
    ps: np.ndarray          # int32 [S]
    cls: np.ndarray         # uint8 [S]
    hap1: np.ndarray
    hap2: np.ndarray
    hap0: np.ndarray
    allhap: np.ndarray
    totsc1: np.ndarray
    totsc2: np.ndarray
    features: np.ndarray    # float64 [6, S]
    join_row: np.ndarray    # int32 [J]
    order: np.ndarray       # int32 [n_emitted]
    shard_counts: np.ndarray  # int64 [n_shards, 8]

    def feature(self, name: str) -> np.ndarray:
        return self.features[_lib.FEATURE_NAMES.index(name)]

    def rows(self, batch: PhaseBatch, sample: int | None = None) -> list[dict]:
        """The reference's return value (sv_phasing_fn.py:213-229): one dict per phased SV,
        stably sorted by (chrom, pos).  The device order already is the reference's append order
        sorted per shard, so the final sort only has to interleave shards."""
        out = []
        order = self.order
        shard_of = (np.searchsorted(batch.sv_off, order, side="right") - 1).tolist()
        # plain Python lists of just the emitted rows: indexing numpy scalars one by one is what costs here
        ps, gt = self.ps[order].tolist(), self.gt[order].tolist()
        pos, svlen = batch.sv_pos[order].tolist(), batch.sv_svlen[order].tolist()
        for k, (i, s) in enumerate(zip(order.tolist(), shard_of)):
            if sample is not None and batch.shard_sample[s] != sample:
                continue
            svtype = batch.sv_type[i]
            ln = svlen[k]
            out.append({"ps": ps[k], "hp": GT_TEXT[gt[k]], "chrom": batch.sv_chrom[i],
                        "pos": pos[k], "svlen": ln if svtype in ("INS", "DUP") else -ln,
                        "svtype": svtype, "ref": batch.sv_ref[i], "alt": batch.sv_alt[i]})
        out.sort(key=lambda d: (d["chrom"], d["pos"]))
        return out


class PhaseEngine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.duet_create(int(device), C.byref(h))
        if rc != _lib.DUET_OK:
            raise DuetError(rc, self.lib.duet_last_error(None).decode())
        self.h = h
        self.device = device
        self._batch = None
        self._keep = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.duet_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != _lib.DUET_OK:
            raise DuetError(rc, self.lib.duet_last_error(self.h).decode())

    # -- configuration ----------------------------------------------------------------------
    def thresholds(self) -> _lib.Thresholds:
        t = _lib.Thresholds()
        self.lib.duet_default_thresholds(C.byref(t))
        return t

    def set_thresholds(self, svlen_thres: int = 50, suppread_thres: int = 2, **overrides):
        t = self.thresholds()
        t.svlen_thres, t.suppread_thres = int(svlen_thres), int(suppread_thres)
        for k, v in overrides.items():
            setattr(t, k, v)
        self._check(self.lib.duet_set_thresholds(self.h, C.byref(t)))

    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.duet_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    # -- data path --------------------------------------------------------------------------
    def upload(self, batch: PhaseBatch, *, tags_in_place: bool = False):
        """Host columns -> device (copies; pinned arrays copy at PCIe speed).  `tags_in_place`: the
        per-read tag records -- the largest column, of which only the joined rows are ever read --
        are NOT copied; the kernels read them over the bus from `batch.read_tag`, which must be
        page-locked (`pin_batch`, `pinned_empty`) and stay untouched until the results are down."""
        inp = _lib.PhaseInput()
        inp.mem = _lib.MEM_HOST_MAPPED if tags_in_place else _lib.MEM_HOST
        inp.n_shards, inp.n_reads, inp.n_svs, inp.n_joins = batch.n_shards, batch.n_reads, batch.n_svs, batch.n_joins
        for name in _lib.INPUT_COLUMNS:
            arr = getattr(batch, name)
            if arr is not None and not arr.flags["C_CONTIGUOUS"]:
                raise ValueError(f"{name} must be C-contiguous")
            setattr(inp, name, _ptr(arr))
        self._batch = batch
        self._check(self.lib.duet_phase_upload(self.h, C.byref(inp)))

    def upload_device(self, batch: PhaseBatch, dev_ptrs: dict, keep=None):
        """Columns already resident in HBM (e.g. torch tensors): `dev_ptrs` maps column name ->
        device address; the offset descriptors still come from `batch` (host)."""
        inp = _lib.PhaseInput()
        inp.mem = _lib.MEM_DEVICE
        inp.n_shards, inp.n_reads, inp.n_svs, inp.n_joins = batch.n_shards, batch.n_reads, batch.n_svs, batch.n_joins
        inp.read_off, inp.sv_off = _ptr(batch.read_off), _ptr(batch.sv_off)
        for name, p in dev_ptrs.items():
            setattr(inp, name, p)
        self._batch = batch
        self._keep = keep
        self._check(self.lib.duet_phase_upload(self.h, C.byref(inp)))

    def execute(self, per_kernel: bool = False):
        self._check(self.lib.duet_phase_execute(self.h, 1 if per_kernel else 0))

    def sync(self):
        self._check(self.lib.duet_sync(self.h))

    def download(self, *, join: bool = True, buffers: dict | None = None, only: tuple | None = None) -> PhaseResult:
        """`only`: names of the outputs to copy back (the rest come back empty) -- the drop-in stage needs just
        the genotype, the phase set, the emission order and the counters."""
        b = self._batch
        S, J, ns = b.n_svs, b.n_joins, b.n_shards
        mk = (lambda name, shape, dt: buffers[name]) if buffers else (lambda name, shape, dt: np.empty(shape, dt))
        if only is not None:
            keep = set(only)
            mk0 = mk
            mk = lambda name, shape, dt: mk0(name, shape, dt) if name in keep else None
            join = join and "join_row" in keep
        arr = {
            "gt": mk("gt", S, np.uint8), "ps": mk("ps", S, np.int32), "cls": mk("cls", S, np.uint8),
            "hap1": mk("hap1", S, np.int32), "hap2": mk("hap2", S, np.int32), "hap0": mk("hap0", S, np.int32),
            "allhap": mk("allhap", S, np.int32), "totsc1": mk("totsc1", S, np.int64),
            "totsc2": mk("totsc2", S, np.int64), "features": mk("features", (_lib.N_FEATURES, S), np.float64),
            "join_row": mk("join_row", J, np.int32) if join else None,
            "order": mk("order", S, np.int32), "shard_counts": mk("shard_counts", (ns, _lib.N_COUNTERS), np.int64),
        }
        out = _lib.PhaseOutput()
        for k, v in arr.items():
            setattr(out, k, _ptr(v))
        self._check(self.lib.duet_phase_download(self.h, C.byref(out)))
        n = int(out.n_emitted)
        for k, v in list(arr.items()):
            if v is None:
                arr[k] = np.zeros((_lib.N_FEATURES, 0) if k == "features" else 0, np.int32)
        arr["order"] = arr["order"][:n]
        return PhaseResult(**arr)

    def run(self, batch: PhaseBatch, *, join: bool = True, buffers: dict | None = None,
            tags_in_place: bool = False, only: tuple | None = None) -> PhaseResult:
        """upload + execute + download.  `buffers` (see `pinned_outputs`) lets the results land in
        page-locked memory, so the device->host copies run as plain DMA; `tags_in_place`: see upload;
        `only`: see download."""
        self.upload(batch, tags_in_place=tags_in_place)
        self.execute()
        return self.download(join=join, buffers=buffers, only=only)

    def timings(self) -> dict:
        t = _lib.Timings()
        self._check(self.lib.duet_get_timings(self.h, C.byref(t)))
        return {"h2d_ms": t.h2d_ms, "device_ms": t.device_ms, "d2h_ms": t.d2h_ms,
                "kernel_ms": dict(zip(_lib.KERNEL_NAMES, list(t.kernel_ms)))}

    def launch_count(self) -> int:
        return int(self.lib.duet_launch_count(self.h))


class PhasePipeline:
    """Cohort mode: several samples (batches) through ONE GPU with two calls in flight.  Two engines, each
    with its own stream and device buffers; sample k+1's upload and kernels are enqueued before sample k's
    results are waited for, so its host->device copy runs under sample k's kernels and device->host copy
    (the reference phases one sample per process run: sv_phasing.py:8-19; a cohort is a loop over samples).
    Every call still moves its own inputs and results -- nothing is cached between samples."""

    def __init__(self, device: int = 0, depth: int = 2):
        self.engines = [PhaseEngine(device) for _ in range(max(1, int(depth)))]

    def close(self):
        for e in self.engines:
            e.close()

    def set_thresholds(self, *a, **kw):
        for e in self.engines:
            e.set_thresholds(*a, **kw)

    def run_many(self, batches, *, buffers=None, tags_in_place: bool = True, join: bool = True, only: tuple | None = None):
        """Yields one PhaseResult per batch, in order.  `buffers`: one `pinned_outputs` dict per engine (results
        land in page-locked memory); a result that lives in such a buffer is overwritten when its engine comes
        round again, i.e. `depth` results later."""
        n_e = len(self.engines)
        pending = []                                           # (engine index) of the calls in flight, oldest first
        for k, batch in enumerate(batches):
            e = k % n_e
            if len(pending) == n_e:                            # this engine still owes a result
                j = pending.pop(0)
                yield self.engines[j].download(join=join, buffers=buffers[j] if buffers else None, only=only)
            self.engines[e].upload(batch, tags_in_place=tags_in_place)
            self.engines[e].execute()
            pending.append(e)
        for j in pending:
            yield self.engines[j].download(join=join, buffers=buffers[j] if buffers else None, only=only)

    def launch_count(self) -> int:
        return sum(e.launch_count() for e in self.engines)


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over page-locked memory from duet_host_alloc; freed when the last view dies."""
    lib = _lib.load()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    p = C.c_void_p()
    rc = lib.duet_host_alloc(C.byref(p), n * dtype.itemsize)
    if rc != _lib.DUET_OK:
        raise DuetError(rc, lib.duet_last_error(None).decode())
    buf = (C.c_uint8 * max(n * dtype.itemsize, 1)).from_address(p.value)
    weakref.finalize(buf, lib.duet_host_free, C.c_void_p(p.value))
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


_COLUMNS = _lib.INPUT_COLUMNS


_IN_ARENA = (("sv_pos", np.int32, 0), ("sv_svlen", np.int32, 0), ("sv_svread", np.int32, 0), ("sv_refread", np.int32, 0),
             ("sv_flags", np.uint8, 0), ("sv_group", np.int32, 0), ("csr_off", np.int64, 1), ("csr_key", np.uint64, 2),
             ("csr_chk", np.uint32, 2))          # name, dtype, length kind: 0 = S, 1 = S + 1, 2 = J
_OUT_ARENA = ("gt", "ps", "cls", "hap1", "hap2", "hap0", "allhap", "totsc1", "totsc2", "features", "shard_counts", "join_row")


def input_arena(n_svs: int, n_joins: int, raw_alloc=None) -> dict:
    """The small input columns as views into ONE page-locked allocation at the offsets the library keeps
    them at on the device (duet_phase_input_layout): duet_phase_upload then moves them with one copy."""
    lib = _lib.load()
    off = (C.c_int64 * 9)()
    total = int(lib.duet_phase_input_layout(n_svs, n_joins, off))
    raw = (raw_alloc or (lambda n: pinned_empty(n, np.uint8)))(max(total, 64))
    out = {}
    for k, (name, dt, kind) in enumerate(_IN_ARENA):
        n = (n_svs, n_svs + 1, n_joins)[kind]
        out[name] = raw[off[k]:off[k] + n * np.dtype(dt).itemsize].view(dt)
    return out


def pinned_outputs(batch: PhaseBatch, raw_alloc=None) -> dict:
    """Page-locked result arrays for `PhaseEngine.run(batch, buffers=...)` / `download(buffers=...)`: views into
    ONE allocation at the library's own result layout (duet_phase_output_layout), so the download is one copy."""
    lib = _lib.load()
    S, J, ns = batch.n_svs, batch.n_joins, batch.n_shards
    off = (C.c_int64 * 12)()
    total = int(lib.duet_phase_output_layout(S, J, ns, off))
    alloc = raw_alloc or (lambda n: pinned_empty(n, np.uint8))
    raw = alloc(max(total, 64))
    shapes = {"gt": (S, np.uint8), "ps": (S, np.int32), "cls": (S, np.uint8), "hap1": (S, np.int32),
              "hap2": (S, np.int32), "hap0": (S, np.int32), "allhap": (S, np.int32), "totsc1": (S, np.int64),
              "totsc2": (S, np.int64), "features": ((_lib.N_FEATURES, S), np.float64), "join_row": (J, np.int32),
              "shard_counts": ((ns, _lib.N_COUNTERS), np.int64)}
    out = {}
    for k, name in enumerate(_OUT_ARENA):
        shape, dt = shapes[name]
        n = int(np.prod(shape))
        out[name] = raw[off[k]:off[k] + n * np.dtype(dt).itemsize].view(dt).reshape(shape)
    out["order"] = alloc(max(S * 4, 64))[:S * 4].view(np.int32)
    return out


def pin_batch(batch: PhaseBatch) -> PhaseBatch:
    """Copy every column into page-locked memory (what a decoder writing into duet_host_alloc'd buffers
    produces directly); the small columns go into one arena (see input_arena)."""
    import dataclasses
    new = {}
    small = input_arena(batch.n_svs, batch.n_joins)
    for name in _COLUMNS:
        arr = getattr(batch, name)
        if name in small:
            if arr is None and name == "csr_chk":
                continue                            # absent check words mean "do not check": stays absent
            dst = small[name]
            dst[...] = 0 if arr is None else arr    # an absent sv_group column: all zero means the same
            new[name] = dst
            continue
        if arr is None:
            continue
        dst = pinned_empty(arr.shape, arr.dtype)
        dst[...] = arr
        new[name] = dst
    return dataclasses.replace(batch, **new)


class PinnedPool:
    """Page-locked host buffers that outlive a call: page-locking costs far more than the copy it speeds up,
    so the decoders of the drop-in stage write into buffers kept from call to call (grow-only, one per name).
    `get` hands out a view; what it held before is overwritten."""

    def __init__(self):
        self._raw: dict[str, np.ndarray] = {}

    def get(self, name: str, shape, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) if not isinstance(shape, int) else int(shape)
        need = n * dtype.itemsize
        raw = self._raw.get(name)
        if raw is None or raw.nbytes < need:
            raw = self._raw[name] = pinned_empty(max(need + need // 4, 64), np.uint8)
        return raw[:need].view(dtype).reshape(shape)


def is_pinned(arr: np.ndarray) -> bool:
    """True when the array's memory is page-locked host memory known to CUDA."""
    if arr is None or arr.nbytes == 0:
        return True
    return bool(_lib.load().duet_host_is_pinned(C.c_void_p(arr.ctypes.data)))
