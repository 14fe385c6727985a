"""Columnar batch handed to the device (the layout of include/duet_b200.h::duet_phase_input).

A *shard* is one (sample, contig) pair: the reference keeps one QNAME dict, one one-PS set and
one prediction loop per contig (/root/reference/src/duet/sv_phasing_fn.py:15-18,195-212), so
shards are independent and any number of them -- all contigs of a sample, or of a cohort --
go to the GPU in one call.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .namehash import hash128_fixed

# include/duet_b200.h :: duet_read_tag (16 bytes)
TAG_DTYPE = np.dtype([("ps", "<i4"), ("pc", "<i4"), ("chk", "<u4"), ("hp", "u1"), ("pad", "u1", (3,))])
assert TAG_DTYPE.itemsize == 16


def pack_tags(hp, ps, pc, hi=None) -> np.ndarray:
    """HP / PS / PC (+ hash high words) -> duet_read_tag records; chk = low 32 bits of `hi`."""
    n = len(hp)
    t = np.zeros(n, TAG_DTYPE)
    t["hp"], t["ps"], t["pc"] = hp, ps, pc
    if hi is not None:
        t["chk"] = (np.asarray(hi, np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return t


class TextColumn:
    """Strings of one per-SV text column (CHROM, REF, ALT, SVTYPE) kept as (offset, length) spans into the
    VCF bytes and decoded on access -- only the rows that end up in phased_sv.vcf are ever materialised."""

    def __init__(self, text, spans: np.ndarray, prefix: str = "", suffix: str = ""):
        self.text, self.spans, self.prefix, self.suffix = text, spans, prefix, suffix      # spans: int64 [n, 2]

    def __len__(self):
        return int(self.spans.shape[0])

    def _one(self, i: int) -> str:
        o, n = int(self.spans[i, 0]), int(self.spans[i, 1])
        return self.prefix + bytes(self.text[o:o + n]).decode("ascii") + self.suffix

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(k) for k in range(*i.indices(len(self)))]
        return self._one(int(i))

    def __iter__(self):
        return (self._one(i) for i in range(len(self)))

    def __eq__(self, other):
        return list(self) == list(other)


@dataclass
class PhaseBatch:
    # shard descriptors
    read_off: np.ndarray      # int64 [n_shards+1]
    sv_off: np.ndarray        # int64 [n_shards+1]
    # haplotagged reads, file order inside a shard (sv_phasing_fn.py:28-29)
    read_key: np.ndarray      # uint64 [R]  low word of the name hash
    read_tag: np.ndarray      # TAG_DTYPE [R]  HP / PS / PC + check word
    # SV records, VCF order inside a shard (read_file.py:30)
    sv_pos: np.ndarray        # int32 [S]
    sv_svlen: np.ndarray      # int32 [S]  |SVLEN|
    sv_svread: np.ndarray     # int32 [S]
    sv_refread: np.ndarray    # int32 [S]
    sv_flags: np.ndarray      # uint8 [S]
    sv_group: np.ndarray | None   # int32 [S] rank of the CHROM string inside the shard
    csr_off: np.ndarray       # int64 [S+1]
    csr_key: np.ndarray       # uint64 [J]
    csr_chk: np.ndarray | None    # uint32 [J] check word, None = no collision check
    # host-only row text (never sent to the device)
    shard_sample: list = field(default_factory=list)     # sample index per shard
    shard_contig: list = field(default_factory=list)     # chrom_list name per shard
    sv_chrom: list = field(default_factory=list)         # CHROM string per SV
    sv_type: list = field(default_factory=list)          # SVTYPE string per SV
    sv_ref: list = field(default_factory=list)
    sv_alt: list = field(default_factory=list)

    @property
    def read_hp(self) -> np.ndarray:
        return self.read_tag["hp"]

    @property
    def read_ps(self) -> np.ndarray:
        return self.read_tag["ps"]

    @property
    def read_pc(self) -> np.ndarray:
        return self.read_tag["pc"]

    @property
    def n_shards(self) -> int:
        return int(self.read_off.shape[0] - 1)

    @property
    def n_reads(self) -> int:
        return int(self.read_key.shape[0])

    @property
    def n_svs(self) -> int:
        return int(self.sv_pos.shape[0])

    @property
    def n_joins(self) -> int:
        return int(self.csr_key.shape[0])

    def input_bytes(self) -> int:
        """Bytes a host->device upload moves (the h2d_bytes_per_step of bench.py)."""
        tot = 0
        for name in _lib.INPUT_COLUMNS:
            arr = getattr(self, name)
            if arr is not None:
                tot += arr.nbytes
        return tot

    def algorithmic_bytes(self) -> dict:
        """33 B per tagged read + 24 B per join + 64 B per SV."""
        r, j, s = self.n_reads, self.n_joins, self.n_svs
        return {"reads": 33 * r, "joins": 24 * j, "svs": 64 * s, "total": 33 * r + 24 * j + 64 * s}

    def validate(self) -> None:
        ns = self.n_shards
        assert self.read_off.dtype == np.int64 and self.sv_off.dtype == np.int64 and self.csr_off.dtype == np.int64
        assert self.read_off[0] == 0 and self.read_off[ns] == self.n_reads
        assert self.sv_off[0] == 0 and self.sv_off[ns] == self.n_svs
        assert self.csr_off[0] == 0 and self.csr_off[self.n_svs] == self.n_joins
        assert self.read_key.dtype == np.uint64 and self.csr_key.dtype == np.uint64
        assert self.read_tag.dtype == TAG_DTYPE and self.read_tag.shape[0] == self.n_reads
        assert self.sv_flags.dtype == np.uint8
        assert self.csr_chk is None or self.csr_chk.dtype == np.uint32
        for a in (self.sv_pos, self.sv_svlen, self.sv_svread, self.sv_refread):
            assert a.dtype == np.int32

    def select_shards(self, idx) -> "PhaseBatch":
        """Sub-batch holding the given shards (in the given order) -- the unit of multi-GPU work."""
        idx = [int(i) for i in idx]

        def cat(parts, dtype):
            return np.ascontiguousarray(np.concatenate(parts).astype(dtype, copy=False)) if parts else np.zeros(0, dtype)

        r_sl = [slice(int(self.read_off[s]), int(self.read_off[s + 1])) for s in idx]
        v_sl = [slice(int(self.sv_off[s]), int(self.sv_off[s + 1])) for s in idx]
        j_sl = [slice(int(self.csr_off[v.start]), int(self.csr_off[v.stop])) for v in v_sl]
        read_off = np.zeros(len(idx) + 1, np.int64)
        sv_off = np.zeros(len(idx) + 1, np.int64)
        read_off[1:] = np.cumsum([r.stop - r.start for r in r_sl])
        sv_off[1:] = np.cumsum([v.stop - v.start for v in v_sl])
        lens = cat([np.diff(self.csr_off[v.start:v.stop + 1]) for v in v_sl], np.int64)
        csr_off = np.zeros(lens.shape[0] + 1, np.int64)
        csr_off[1:] = np.cumsum(lens)
        pick = lambda a, sl, dt: None if a is None else cat([a[x] for x in sl], dt)
        pick_list = lambda lst: [x for v in v_sl for x in lst[v]] if lst else []
        return PhaseBatch(
            read_off, sv_off,
            pick(self.read_key, r_sl, np.uint64), pick(self.read_tag, r_sl, TAG_DTYPE),
            pick(self.sv_pos, v_sl, np.int32), pick(self.sv_svlen, v_sl, np.int32),
            pick(self.sv_svread, v_sl, np.int32), pick(self.sv_refread, v_sl, np.int32),
            pick(self.sv_flags, v_sl, np.uint8), pick(self.sv_group, v_sl, np.int32),
            csr_off, pick(self.csr_key, j_sl, np.uint64), pick(self.csr_chk, j_sl, np.uint32),
            [self.shard_sample[s] for s in idx] if self.shard_sample else [],
            [self.shard_contig[s] for s in idx] if self.shard_contig else [],
            pick_list(self.sv_chrom), pick_list(self.sv_type), pick_list(self.sv_ref), pick_list(self.sv_alt))


def from_synth(samples, *, with_hi: bool = True, with_text: bool = True) -> PhaseBatch:
    """Build the batch straight from generator arrays (no text round trip).  Equal to what the
    text decoders produce from `synth.write_workdir` of the same samples (tests check that)."""
    from . import synth as sy
    if not isinstance(samples, (list, tuple)):
        samples = [samples]
    read_off, sv_off = [0], [0]
    cols = {k: [] for k in ("rk", "rt", "pos", "len", "svr", "ref", "flg", "ck", "ch", "cl")}
    shard_sample, shard_contig, sv_chrom, sv_type, sv_alt = [], [], [], [], []
    for si, sample in enumerate(samples):
        for c in sample.contigs:
            t = c.row_tagged
            lo, hi = hash128_fixed(sy.names_from_ids(c.row_id[t]))
            cols["rk"].append(lo)
            cols["rt"].append(pack_tags(c.row_hp[t], c.row_ps[t], c.row_pc[t], hi if with_hi else None))
            read_off.append(read_off[-1] + int(t.sum()))
            n = c.sv_pos.shape[0]
            sv_off.append(sv_off[-1] + n)
            cols["pos"].append(c.sv_pos); cols["len"].append(np.abs(c.sv_len))
            cols["svr"].append(c.sv_svread); cols["ref"].append(c.sv_refread)
            cols["flg"].append((c.sv_gt == sy.GTS.index("./.")).astype(np.uint8) * _lib.SV_GT_MISSING)
            lo, hi = hash128_fixed(sy.names_from_ids(c.sup_id))
            cols["ck"].append(lo); cols["ch"].append((hi & np.uint64(0xFFFFFFFF)).astype(np.uint32))
            cols["cl"].append(np.diff(c.sup_off))
            shard_sample.append(si); shard_contig.append(c.name)
            if with_text:
                cn = ("chr" if sample.chr_prefix else "") + c.name
                sv_chrom += [cn] * n
                types = [sy.SVTYPES[int(x)] for x in c.sv_type]
                sv_type += types
                sv_alt += ["<" + x + ">" for x in types]
    cat = lambda k, dt: (np.concatenate(cols[k]) if cols[k] else np.zeros(0, dt)).astype(dt, copy=False)
    lens = cat("cl", np.int64)
    csr_off = np.zeros(lens.shape[0] + 1, np.int64)
    csr_off[1:] = np.cumsum(lens)
    b = PhaseBatch(
        np.asarray(read_off, np.int64), np.asarray(sv_off, np.int64),
        cat("rk", np.uint64), cat("rt", TAG_DTYPE),
        cat("pos", np.int32), cat("len", np.int32), cat("svr", np.int32), cat("ref", np.int32),
        cat("flg", np.uint8), None, csr_off, cat("ck", np.uint64), cat("ch", np.uint32) if with_hi else None,
        shard_sample, shard_contig, sv_chrom, sv_type, ["N"] * len(sv_chrom), sv_alt)
    b.validate()
    return b
