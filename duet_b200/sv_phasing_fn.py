"""Drop-in for /root/reference/src/duet/sv_phasing_fn.py: same public functions, same
arguments, same return value of `generate_phased_callset` -- computed on the GPU.

    reference                         here
    read_hap_bam      (:11-34)   ->   read_hap_bam: `samtools view` text -> per-contig read columns (C++ scan)
    generate_callinfo (:36-68)   ->   generate_callinfo: read columns + SV columns -> one PhaseBatch
                                      (the JOIN itself happens on the device)
    get_phase_info / predict_hp  ->   device kernels (csrc/phase_kernels.cuh)
    generate_phased_callset      ->   decode, one device call, rows

There is no CPU compute path: without libduet_b200.so and a CUDA device these functions raise.
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import shlex
import subprocess
import time
from dataclasses import dataclass

import numpy as np

from . import _lib
from .columnar import TAG_DTYPE, PhaseBatch
from .engine import DuetError, PhaseEngine
from .read_file import ContigSvs, init_chrom_list, parse_vcf

_ENGINES: dict[int, PhaseEngine] = {}
last_timings: dict = {}          # host decode / device / row-building seconds of the last call


def get_engine(device: int | None = None) -> PhaseEngine:
    if device is None:
        device = int(os.environ.get("DUET_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    eng = _ENGINES.get(device)
    if eng is None:
        eng = _ENGINES[device] = PhaseEngine(device)
    return eng


@dataclass
class ReadColumns:
    """Kept rows of one per-contig haplotagged BAM, file order (sv_phasing_fn.py:26-29)."""
    key: np.ndarray          # uint64 low hash word
    tag: np.ndarray          # TAG_DTYPE: HP / PS / PC + check word
    n_lines: int = 0
    source: str | None = None    # the file the rows came from

    @property
    def hp(self):
        return self.tag["hp"]

    @property
    def ps(self):
        return self.tag["ps"]

    @property
    def pc(self):
        return self.tag["pc"]

    def __len__(self):
        return int(self.key.shape[0])

    @staticmethod
    def empty():
        return ReadColumns(np.zeros(0, np.uint64), np.zeros(0, TAG_DTYPE))


_DECODE_EXC = {
    _lib.DECODE_ERR_INDEX: (IndexError, "list index out of range"),
    _lib.DECODE_ERR_VALUE: (ValueError, "invalid literal for int() in an HP/PC/PS field"),
    _lib.DECODE_ERR_ASCII: (UnicodeDecodeError, None),
    _lib.DECODE_ERR_RANGE: (OverflowError, "HP outside 0..255 or PS/PC outside int32"),
}


def _read_only_view(data):
    """What ctypes should pass for a `const char *` parameter: `bytes` go by pointer (the C side only
    reads), writable buffers are wrapped in place."""
    if isinstance(data, bytes):
        return C.c_char_p(data) if data else C.c_char_p(b"\0")
    return (C.c_char * max(len(data), 1)).from_buffer(data)


def decode_sam_text(text: bytes) -> ReadColumns:
    """One contig's `samtools view` output -> columns (C++: csrc/decode.cpp)."""
    lib = _lib.load()
    n = len(text)
    buf = _read_only_view(text)                    # no copy: a 40 MB memcpy per contig under the GIL serialises the threads
    cap = int(lib.duet_count_lines(buf, n))
    key, tag = np.empty(cap, np.uint64), np.empty(cap, TAG_DTYPE)
    n_rows, n_lines, err_line = C.c_int64(), C.c_int64(), C.c_int64()
    rc = lib.duet_decode_sam_text(buf, n, cap, key.ctypes.data, tag.ctypes.data, C.byref(n_rows), C.byref(n_lines),
                                  C.byref(err_line))
    if rc != _lib.DUET_OK:
        exc, msg = _DECODE_EXC.get(rc, (RuntimeError, f"decode error {rc}"))
        if exc is UnicodeDecodeError:
            raise UnicodeDecodeError("ascii", bytes(text[:1]), 0, 1, f"ordinal not in range(128) (line {err_line.value})")
        raise exc(f"{msg} (alignment line {err_line.value})")
    k = n_rows.value
    return ReadColumns(key[:k], tag[:k], n_lines.value)


def decode_bam(data: bytes) -> ReadColumns:
    """A whole BGZF/BAM file -> columns, natively (csrc/bam_decode.cpp): same rows, same values as
    `samtools view` piped through decode_sam_text."""
    lib = _lib.load()
    key_p, tag_p = C.c_void_p(), C.c_void_p()
    n_rows, n_rec, err = C.c_int64(), C.c_int64(), C.c_int64()
    buf = _read_only_view(data)
    rc = lib.duet_decode_bam(buf, len(data), C.byref(key_p), C.byref(tag_p), C.byref(n_rows), C.byref(n_rec), C.byref(err))
    if rc != _lib.DUET_OK:
        exc, msg = _DECODE_EXC.get(rc, (ValueError, f"not a readable BAM (decode error {rc})"))
        if exc is UnicodeDecodeError:
            raise UnicodeDecodeError("ascii", b"\xff", 0, 1, f"ordinal not in range(128) (record {err.value})")
        raise exc(f"{msg} (alignment record {err.value})")
    k = n_rows.value
    try:
        key = np.ctypeslib.as_array(C.cast(key_p, C.POINTER(C.c_uint64)), shape=(max(k, 1),))[:k].copy()
        raw = np.ctypeslib.as_array(C.cast(tag_p, C.POINTER(C.c_uint8)), shape=(max(k, 1) * 16,))[:k * 16].copy()
    finally:
        lib.duet_free(key_p)
        lib.duet_free(tag_p)
    return ReadColumns(key, raw.view(TAG_DTYPE), n_rec.value)


def load_hap_bam(path: str, thread: int) -> ReadColumns:
    """One per-contig haplotagged BAM -> columns.  BGZF/BAM files are decoded natively; anything
    else gzip-compressed goes through `samtools view` like the reference (sv_phasing_fn.py:25);
    uncompressed files are SAM text already."""
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:2] != b"\x1f\x8b":
        return decode_sam_text(data)
    try:
        return decode_bam(data)
    except ValueError as e:
        if "not a readable BAM" not in str(e):
            raise
    return decode_sam_text(subprocess.check_output(shlex.split("samtools view -@" + str(thread) + " " + path)))


def _sam_text(path: str, thread: int) -> bytes:
    """`samtools view -@thread <bam>` (sv_phasing_fn.py:25).  A file that is not gzip/BGZF already
    is SAM text and is read directly (no samtools needed)."""
    with open(path, "rb") as fh:
        magic = fh.read(2)
        if magic != b"\x1f\x8b":
            return magic + fh.read()
    return subprocess.check_output(shlex.split("samtools view -@" + str(thread) + " " + path))


def read_hap_bam(path, thread, include_all_ctgs):
    """-> list[ReadColumns], one per contig of init_chrom_list; a missing BAM gives an empty table
    (sv_phasing_fn.py:19-24).  `path` = '<home>/snp_phasing/'."""
    logging.info("extract SNP signatures")
    chrom_list = init_chrom_list(include_all_ctgs, path[:len(path) - 13])
    paths = []
    for ctg in chrom_list:
        if os.path.exists(path + "chr" + ctg + ".bam"):
            paths.append(path + "chr" + ctg + ".bam")
        elif os.path.exists(path + ctg + ".bam"):
            paths.append(path + ctg + ".bam")
        else:
            paths.append(None)
    # the reference gives its `thread` count to samtools; here the contigs are decoded in parallel
    # (the C++ scanners release the GIL)
    from concurrent.futures import ThreadPoolExecutor
    thread = max(1, int(thread))
    n_files = max(1, sum(p is not None for p in paths))
    workers = min(thread, n_files)
    # fewer files than threads (the one-contig demo): the rest go to the BGZF blocks inside each file
    prev = _lib.load().duet_set_decode_threads(max(1, thread // workers))
    try:
        with ThreadPoolExecutor(max_workers=workers) as pool:
            read_hap = list(pool.map(lambda p: load_hap_bam(p, 1) if p else ReadColumns.empty(), paths))
    finally:
        _lib.load().duet_set_decode_threads(prev)
    for cols, p in zip(read_hap, paths):
        cols.source = p
    for ctg, cols, p in zip(chrom_list, read_hap, paths):
        if p:
            logging.info(("  signatures extracted from " if cols.n_lines else "  no signature from ") + ctg)
    return read_hap


def hash_name_lists(lists: list[list[str]]):
    """Hash every support-read name of every SV; returns (csr lengths, lo, hi)."""
    lib = _lib.load()
    lens = np.fromiter((len(l) for l in lists), np.int64, len(lists))
    flat = [n for l in lists for n in l]
    blob = "".join(flat).encode("ascii")
    off = np.zeros(len(flat) + 1, np.int64)
    if flat:
        np.cumsum(np.fromiter((len(n) for n in flat), np.int64, len(flat)), out=off[1:])
    lo, hi = np.empty(len(flat), np.uint64), np.empty(len(flat), np.uint64)
    if flat:
        lib.duet_hash_names(blob, off.ctypes.data, len(flat), lo.ctypes.data, hi.ctypes.data)
    return lens, lo, hi


def hash_name_csv(csv_lists: list[str]):
    """Same result as hash_name_lists([s.split(',') for s in csv_lists]) without creating a Python string
    per name: the C++ side splits and hashes (duet_hash_name_lists)."""
    lib = _lib.load()
    n = len(csv_lists)
    lens = np.zeros(n, np.int64)
    if n == 0:
        return lens, np.empty(0, np.uint64), np.empty(0, np.uint64)
    blob = "\n".join(csv_lists).encode("ascii")
    cap = blob.count(b",") + n
    lo, hi = np.empty(cap, np.uint64), np.empty(cap, np.uint64)
    got = lib.duet_hash_name_lists(blob, len(blob), n, cap, lens.ctypes.data, lo.ctypes.data, hi.ctypes.data)
    if got != cap:
        raise ValueError("support-read name lists contain a line break")     # cannot come from a split VCF line
    return lens, lo, hi


def generate_callinfo(caller_path, read_hap, include_all_ctgs, comp_call=None) -> PhaseBatch:
    """The reference joins here on the host (:46-48); this version only lays the two sides out as
    one columnar batch -- shard = contig -- and leaves the join to the device.  `comp_call`: the
    already parsed VCF (generate_phased_callset parses it while the BAMs are being decoded)."""
    if comp_call is None:
        logging.info("extract SV signatures")
        comp_call = parse_vcf(caller_path, include_all_ctgs)
    chrom_list = init_chrom_list(include_all_ctgs, caller_path[:len(caller_path) - 24])
    return build_batch(chrom_list, read_hap, contig_records(chrom_list, comp_call, read_hap))


def contig_records(chrom_list, comp_call: list[ContigSvs], read_hap: list[ReadColumns] | None = None):
    """SV records each contig's prediction loop sees, in the reference's order.

    The reference flattens the per-contig record lists and later selects, for contig c, every flat record
    whose CHROM is 'c' or 'chr'+c (sv_phasing_fn.py:198,208).  Normally that is contig c's own list.  With
    `include_all_ctgs` the list may name both 'c' and 'chr'+c: then a 'chrc' record was parsed (and joined)
    once per naming contig and EVERY copy is selected by each of them -- reproduced here.  A copy parsed
    under another contig was joined against that contig's BAM; that is only reproducible when both
    contigs read the same file, anything else is refused."""
    owners: dict[str, list[int]] = {}
    for ch, c in enumerate(chrom_list):
        for nm in ("chr" + c, c):
            if ch not in owners.setdefault(nm, []):
                owners[nm].append(ch)
    if all(len(v) == 1 for v in owners.values()):
        return comp_call
    out = []
    for ch, c in enumerate(chrom_list):
        accept = ("chr" + c, c)
        cs = ContigSvs()
        for o, src in enumerate(comp_call):
            idx = [i for i, nm in enumerate(src.chrom) if nm in accept]
            if not idx:
                continue
            if o != ch and read_hap is not None and read_hap[o].source != read_hap[ch].source:
                raise NotImplementedError(
                    f"contigs {chrom_list[o]!r} and {c!r} both claim CHROM {src.chrom[idx[0]]!r} but read different "
                    "haplotagged BAMs; the reference's result for such a contig list cannot be reproduced per contig")
            for f in ("chrom", "pos", "ref", "alt", "svlen", "svtype", "svread", "names_csv", "gt", "refread", "altread"):
                getattr(cs, f).extend(getattr(src, f)[i] for i in idx)
        out.append(cs)
    return out


def build_batch(chrom_list, read_hap: list[ReadColumns], comp_call: list[ContigSvs], sample: int = 0) -> PhaseBatch:
    ns = len(chrom_list)
    read_off = np.zeros(ns + 1, np.int64)
    sv_off = np.zeros(ns + 1, np.int64)
    read_off[1:] = np.cumsum([len(r) for r in read_hap])
    sv_off[1:] = np.cumsum([len(c) for c in comp_call])
    cat = lambda parts, dt: (np.concatenate(parts) if parts else np.zeros(0, dt)).astype(dt, copy=False)
    chrom, svtype, ref, alt, pos, svlen, svread, refread, flags, group, lists = [], [], [], [], [], [], [], [], [], [], []
    any_group = False
    for cs in comp_call:
        chrom += cs.chrom; svtype += cs.svtype; ref += cs.ref; alt += cs.alt
        pos += cs.pos; svlen += [abs(v) for v in cs.svlen]; svread += cs.svread; refread += cs.refread
        flags += [_lib.SV_GT_MISSING if g == "./." else 0 for g in cs.gt]
        lists += cs.names_csv
        ranks = {c: i for i, c in enumerate(sorted(set(cs.chrom)))}     # 'chr1' and '1' rows in one contig
        any_group |= len(ranks) > 1
        group += [ranks[c] for c in cs.chrom]
    lens, ck, ch = hash_name_csv(lists)
    csr_off = np.zeros(len(lists) + 1, np.int64)
    np.cumsum(lens, out=csr_off[1:])
    b = PhaseBatch(
        read_off, sv_off,
        cat([r.key for r in read_hap], np.uint64), cat([r.tag for r in read_hap], TAG_DTYPE),
        np.asarray(pos, np.int32), np.asarray(svlen, np.int32), np.asarray(svread, np.int32),
        np.asarray(refread, np.int32), np.asarray(flags, np.uint8),
        np.asarray(group, np.int32) if any_group else None, csr_off, ck,
        (ch & np.uint64(0xFFFFFFFF)).astype(np.uint32),
        [sample] * ns, list(chrom_list), chrom, svtype, ref, alt)
    b.validate()
    return b


_STATUS_EXC = {_lib.ERR_BAD_HP: KeyError, _lib.ERR_ZERO_DIVISION: ZeroDivisionError}


def phase_batch(batch: PhaseBatch, svlen_thres, suppread_thres, engine: PhaseEngine | None = None):
    eng = engine or get_engine()
    eng.set_thresholds(int(svlen_thres), int(suppread_thres))
    try:
        return eng.run(batch)
    except DuetError as e:
        exc = _STATUS_EXC.get(e.code)
        if exc is None:
            raise
        raise exc(e.msg) from e      # the exception the reference raises at :96 / :123


# ---- the fast path of the stage: both decoders write straight into page-locked columns --------------------
_POOL = None          # engine.PinnedPool of the process: page-locked column and result buffers, kept between calls
last_batch = None     # the batch of the last generate_phased_callset call (tests look at where its columns live)


def _pool():
    global _POOL
    if _POOL is None:
        from .engine import PinnedPool
        _POOL = PinnedPool()
    return _POOL


def _hap_paths(path, chrom_list):
    out = []
    for ctg in chrom_list:
        if os.path.exists(path + "chr" + ctg + ".bam"):
            out.append(path + "chr" + ctg + ".bam")
        elif os.path.exists(path + ctg + ".bam"):
            out.append(path + ctg + ".bam")
        else:
            out.append(None)
    return out


class _ReadJob:
    """One haplotagged file scanned natively (csrc: duet_decode_reads): the kept rows stay inside the library
    until `take` copies them into their slice of the page-locked columns.  BGZF files are mapped, not read:
    the inflate runs in bounded batches, so memory does not grow with the file."""

    def __init__(self, path: str):
        import mmap
        lib = _lib.load()
        self.lib, self.path, self.job, self.n_rows, self.n_records = lib, path, C.c_void_p(), 0, 0
        with open(path, "rb") as fh:
            size = os.fstat(fh.fileno()).st_size
            if size == 0:
                return
            mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            view = np.frombuffer(mm, np.uint8)
            if bytes(view[:2]) == b"\x1f\x8b":
                kind = _lib.READS_BAM
            else:
                kind = _lib.READS_SAM_TEXT
            n_rows, n_rec, err = C.c_int64(), C.c_int64(), C.c_int64()
            rc = lib.duet_decode_reads(C.c_void_p(view.ctypes.data), size, kind, C.byref(self.job), C.byref(n_rows),
                                       C.byref(n_rec), C.byref(err))
            if rc == _lib.DECODE_ERR_FORMAT and kind == _lib.READS_BAM:
                raise _NotBam()                                   # gzip, but not BGZF/BAM: samtools' business
            if rc != _lib.DUET_OK:
                exc, msg = _DECODE_EXC.get(rc, (RuntimeError, f"decode error {rc}"))
                what = "record" if kind == _lib.READS_BAM else "line"
                if exc is UnicodeDecodeError:
                    raise UnicodeDecodeError("ascii", b"\xff", 0, 1, f"ordinal not in range(128) ({what} {err.value})")
                raise exc(f"{msg} (alignment {what} {err.value})")
            self.n_rows, self.n_records = n_rows.value, n_rec.value
        finally:
            del view
            mm.close()

    def take(self, key: np.ndarray, tag: np.ndarray):
        if self.job:
            self.lib.duet_rows_take(self.job, C.c_void_p(key.ctypes.data), C.c_void_p(tag.ctypes.data))
            self.job = C.c_void_p()

    def drop(self):
        if self.job:
            self.lib.duet_rows_free(self.job)
            self.job = C.c_void_p()


class _NotBam(Exception):
    pass


def _fast_batch(vcf_path, sam_home, thread, include_all_ctgs):
    """Both decodes natively and concurrently, rows and records landing in page-locked columns.  Returns None
    when an input is outside what the native readers claim (then the general path below takes over)."""
    from concurrent.futures import ThreadPoolExecutor
    from .columnar import TextColumn
    from .engine import input_arena
    from .read_file import decode_sv_vcf
    chrom_list = init_chrom_list(include_all_ctgs, sam_home[:len(sam_home) - 13])
    paths = _hap_paths(sam_home, chrom_list)
    thread = max(1, int(thread))
    pool = _pool()
    n_files = max(1, sum(p is not None for p in paths))
    workers = min(thread, n_files)
    prev = _lib.load().duet_set_decode_threads(max(1, thread // workers))
    jobs = []
    try:
        with ThreadPoolExecutor(max_workers=workers) as ex:
            futs = [ex.submit(_ReadJob, p) if p else None for p in paths]
            logging.info("extract SV signatures")
            svs = decode_sv_vcf(vcf_path, include_all_ctgs, thread,
                                arena=lambda S, J: input_arena(S, J, raw_alloc=lambda n: pool.get("sv_arena", n, np.uint8)))
            try:
                jobs = [f.result() if f else None for f in futs]
            except _NotBam:
                svs = None
        if svs is None:
            return None
        ns = len(chrom_list)
        read_off = np.zeros(ns + 1, np.int64)
        read_off[1:] = np.cumsum([j.n_rows if j else 0 for j in jobs])
        R = int(read_off[-1])
        read_key, read_tag = pool.get("read_key", R, np.uint64), pool.get("read_tag", R, TAG_DTYPE)
        with ThreadPoolExecutor(max_workers=workers) as ex:
            list(ex.map(lambda k: jobs[k].take(read_key[read_off[k]:read_off[k + 1]], read_tag[read_off[k]:read_off[k + 1]])
                        if jobs[k] else None, range(ns)))
        for ctg, j in zip(chrom_list, jobs):
            if j:
                logging.info(("  signatures extracted from " if j.n_records else "  no signature from ") + ctg)
        np.abs(svs.svlen, out=svs.svlen)                          # the device column is |SVLEN| (:62); rows take the sign from SVTYPE
        tc = lambda k, pre="", suf="": TextColumn(svs.text, svs.str_span[:, k, :], pre, suf)
        batch = PhaseBatch(read_off, svs.sv_off, read_key, read_tag, svs.pos, svs.svlen, svs.svread, svs.refread, svs.flags,
                           svs.group, svs.csr_off, svs.csr_key, svs.csr_chk, [0] * ns, list(chrom_list),
                           tc(0), tc(3), tc(1), tc(2))
        batch.validate()
        return batch
    finally:
        for j in jobs:
            if j:
                j.drop()
        _lib.load().duet_set_decode_threads(prev)


def generate_phased_callset(vcf_path, sam_home, svlen_thres, suppread_thres, thread, include_all_ctgs):
    """Same signature and return value as the reference's (sv_phasing_fn.py:185-230)."""
    global last_batch
    t0 = time.perf_counter()
    logging.info("extract SNP signatures")
    batch = None
    if not os.environ.get("DUET_GENERAL_DECODE"):
        batch = _fast_batch(vcf_path, sam_home, thread, include_all_ctgs)
    fast = batch is not None
    if batch is None:
        # The general readers.  The two decodes are independent: the haplotagged BAMs are scanned by C++ threads
        # (GIL released) while this thread parses the SV VCF.  Errors surface in the reference's order: it reads
        # the BAMs first (:186).
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=1) as background:
            pending = background.submit(read_hap_bam, sam_home, thread, include_all_ctgs)
            vcf_error = comp_call = None
            try:
                logging.info("extract SV signatures")
                comp_call = parse_vcf(vcf_path, include_all_ctgs)
            except Exception as e:                               # noqa: BLE001 -- re-raised below, after the BAM errors
                vcf_error = e
            read_hap = pending.result()
        if vcf_error is not None:
            raise vcf_error
        batch = generate_callinfo(vcf_path, read_hap, include_all_ctgs, comp_call)
    last_batch = batch
    t1 = time.perf_counter()
    logging.info("integrate read weight information")
    logging.info("calculate read weight statistics")
    logging.info("predict SV haplotypes in the callset")
    eng = get_engine()
    eng.set_thresholds(int(svlen_thres), int(suppread_thres))
    try:
        if fast:
            # page-locked columns: the tag records stay where they are (only the joined rows cross the bus), the
            # results the rows need -- genotype, phase set, order, counters -- land in page-locked buffers too
            pool, S, ns = _pool(), batch.n_svs, batch.n_shards
            bufs = {"gt": pool.get("o_gt", S, np.uint8), "ps": pool.get("o_ps", S, np.int32), "order": pool.get("o_order", S, np.int32),
                    "shard_counts": pool.get("o_counts", (ns, _lib.N_COUNTERS), np.int64)}
            res = eng.run(batch, buffers=bufs, tags_in_place=True, only=("gt", "ps", "order", "shard_counts"))
        else:
            res = eng.run(batch)
    except DuetError as e:
        exc = _STATUS_EXC.get(e.code)
        if exc is None:
            raise
        raise exc(e.msg) from e      # the exception the reference raises at :96 / :123
    t2 = time.perf_counter()
    rows = res.rows(batch)
    t3 = time.perf_counter()
    last_timings.clear()
    last_timings.update(host_decode_s=t1 - t0, device_call_s=t2 - t1, rows_s=t3 - t2, native_decode=fast, **eng.timings())
    return rows
