"""Test helpers: hand-built device batches and oracle comparison."""
from __future__ import annotations

import numpy as np

from duet_b200 import _lib
from duet_b200.columnar import PhaseBatch, pack_tags
from duet_b200.namehash import hash_names


class BatchBuilder:
    """Assemble a PhaseBatch shard by shard from explicit (name, hp, ps, pc) rows and SV lists."""

    def __init__(self):
        self.read_off, self.sv_off = [0], [0]
        self.rnames, self.hp, self.ps, self.pc = [], [], [], []
        self.pos, self.svlen, self.svread, self.refread, self.flags = [], [], [], [], []
        self.lists = []
        self.chrom, self.svtype = [], []
        self.contig = []

    def shard(self, reads, svs, contig="1"):
        """reads: [(name, hp, ps, pc)] in file order; svs: [dict(pos, svlen, svread, refread, names,
        gt_missing=False, svtype='INS')] in VCF order."""
        for nm, hp, ps, pc in reads:
            self.rnames.append(nm); self.hp.append(hp); self.ps.append(ps); self.pc.append(pc)
        self.read_off.append(len(self.rnames))
        for sv in svs:
            self.pos.append(sv["pos"]); self.svlen.append(sv.get("svlen", 100))
            self.svread.append(sv["svread"]); self.refread.append(sv["refread"])
            self.flags.append(_lib.SV_GT_MISSING if sv.get("gt_missing") else 0)
            self.lists.append(list(sv["names"]))
            self.chrom.append(sv.get("chrom", contig)); self.svtype.append(sv.get("svtype", "INS"))
        self.sv_off.append(len(self.pos))
        self.contig.append(contig)
        return self

    def build(self, with_hi=True) -> PhaseBatch:
        rk, rh = hash_names(self.rnames)
        flat = [n for l in self.lists for n in l]
        ck, ch = hash_names(flat)
        csr_off = np.zeros(len(self.lists) + 1, np.int64)
        csr_off[1:] = np.cumsum([len(l) for l in self.lists])
        i32 = lambda x: np.asarray(x, np.int32).reshape(-1)
        tags = pack_tags(np.asarray(self.hp, np.uint8).reshape(-1), i32(self.ps), i32(self.pc), rh if with_hi else None)
        b = PhaseBatch(np.asarray(self.read_off, np.int64), np.asarray(self.sv_off, np.int64),
                       rk, tags, i32(self.pos), i32(self.svlen), i32(self.svread), i32(self.refread),
                       np.asarray(self.flags, np.uint8).reshape(-1), None, csr_off, ck,
                       (ch & np.uint64(0xFFFFFFFF)).astype(np.uint32) if with_hi else None,
                       [0] * len(self.contig), list(self.contig), list(self.chrom), list(self.svtype),
                       ["N"] * len(self.chrom), ["<" + t + ">" for t in self.svtype])
        b.validate()
        return b


def kat_class(case) -> int:
    """Class the pipeline derives for a KAT case (distinct PS over joined reads, capped at 2)."""
    return min(len({r[2] for r in case["reads"] if len(r) > 1}), 2)


def kat_shard(bb: BatchBuilder, case, tag: str):
    """One shard = the KAT's SV plus helper class-1 SVs that pin the one-PS set to case['oneps'].
    Returns the index (inside the shard) of the KAT's SV (always 0)."""
    reads, names = [], []
    for k, r in enumerate(case["reads"]):
        nm = f"{tag}.r{k}"
        names.append(nm)
        if len(r) > 1:
            reads.append((nm, r[1], r[2], r[3]))
    svs = [dict(pos=case["pos"], svread=case["svread"], refread=case["refread"], names=names)]
    for k, ps in enumerate(case["oneps"]):
        nm = f"{tag}.h{k}"
        reads.append((nm, 1, ps, 0))
        svs.append(dict(pos=1, svread=2, refread=5, names=[nm]))
    bb.shard(reads, svs)
    return 0


def assert_matches_trace(res, batch, trace, flat, index_of=None):
    """Device result vs the oracle's per-SV trace and join (all SVs of one sample/batch).
    `trace` entries: (contig idx, class, record, pred, features); `flat`: joined records."""
    index_of = index_of or (lambda r: r.index)
    seen = np.zeros(batch.n_svs, bool)
    for ci, ps_num, r, pred, f in trace:
        i = index_of(r)
        seen[i] = True
        assert res.cls[i] == ps_num, (i, res.cls[i], ps_num)
        assert res.gt[i] == pred, (i, res.gt[i], pred, f)
        assert res.ps[i] == f["ps"], (i, res.ps[i], f["ps"])
        for k, arr in (("hap1", res.hap1), ("hap2", res.hap2), ("hap0", res.hap0), ("allhap", res.allhap),
                       ("hap1_totsc", res.totsc1), ("hap2_totsc", res.totsc2)):
            assert arr[i] == f[k], (k, i, arr[i], f[k])
        for k in _lib.FEATURE_NAMES:
            got, want = float(res.feature(k)[i]), float(f[k])
            assert got == want, (k, i, got, want)          # same IEEE operations: bit-equal
    assert not res.gt[~seen].any()                          # everything else was dropped
    # the join itself
    for r in flat:
        i = index_of(r)
        b, e = int(batch.csr_off[i]), int(batch.csr_off[i + 1])
        assert e - b == len(r.reads)
        rows = res.join_row[b:e]
        for row, rd in zip(rows.tolist(), r.reads):
            if len(rd) == 1:
                assert row == -1
            else:
                assert row >= 0
                assert (int(batch.read_hp[row]), int(batch.read_ps[row]), int(batch.read_pc[row])) == tuple(rd[1:])
