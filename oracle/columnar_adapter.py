"""TEST INFRASTRUCTURE: run the oracle port on a columnar PhaseBatch (64-bit keys stand in for
the read names -- the algorithm only needs them hashable) and lay its answers out like the
device's PhaseResult, so the two can be compared array by array at any size."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from duet_b200 import _lib

from . import ref_port


def phase_batch_oracle(batch, svlen_thres=50, suppread_thres=2):
    S, J, ns = batch.n_svs, batch.n_joins, batch.n_shards
    gt = np.zeros(S, np.uint8)
    ps = np.zeros(S, np.int32)
    cls = np.full(S, _lib.CLS_FILTERED, np.uint8)
    ints = {k: np.zeros(S, np.int64) for k in ("hap1", "hap2", "hap0", "allhap", "hap1_totsc", "hap2_totsc")}
    feats = np.zeros((_lib.N_FEATURES, S), np.float64)
    join_row = np.full(J, -1, np.int32)
    counts = np.zeros((ns, _lib.N_COUNTERS), np.int64)
    traced = np.zeros(S, bool)
    order = []
    keys = batch.read_key.tolist()
    hp, pss, pc = batch.read_hp.tolist(), batch.read_ps.tolist(), batch.read_pc.tolist()
    ckeys = batch.csr_key.tolist()
    csr = batch.csr_off.tolist()
    for s in range(ns):
        r0, r1 = int(batch.read_off[s]), int(batch.read_off[s + 1])
        v0, v1 = int(batch.sv_off[s]), int(batch.sv_off[s + 1])
        table, last_row = {}, {}
        for r in range(r0, r1):                       # later rows overwrite (sv_phasing_fn.py:29)
            table[keys[r]] = (hp[r], pss[r], pc[r])
            last_row[keys[r]] = r
        recs = []
        for i in range(v0, v1):
            chrom = batch.sv_chrom[i] if batch.sv_chrom else "c"
            rec = ref_port.SvRecord([chrom, str(int(batch.sv_pos[i])), ".", "N", "<X>"])
            rec.svlen = int(batch.sv_svlen[i])
            rec.svtype = batch.sv_type[i] if batch.sv_type else "INS"
            rec.svread = int(batch.sv_svread[i])
            rec.refread = int(batch.sv_refread[i])
            rec.gt = "./." if batch.sv_flags[i] & _lib.SV_GT_MISSING else "0/1"
            rec.names = ckeys[csr[i]:csr[i + 1]]
            recs.append(rec)
            for j in range(csr[i], csr[i + 1]):
                join_row[j] = last_row.get(ckeys[j], -1)
        flat = ref_port.join_support_reads([recs], [table])
        trace = []
        chroms = sorted({r.chrom for r in recs}) or ["c"]
        # one contig whose accepted CHROM strings are exactly the shard's
        rows = _phase_one_shard(flat, chroms, svlen_thres, suppread_thres, trace)
        emitted = []
        for ci, ps_num, rec, pred, f in trace:
            i = v0 + rec.index
            traced[i] = True
            cls[i], gt[i], ps[i] = ps_num, pred, f["ps"]
            for k in ints:
                ints[k][i] = f[k]
            for k, name in enumerate(_lib.FEATURE_NAMES):
                feats[k, i] = f[name]
            if pred:
                grp = int(batch.sv_group[i]) if batch.sv_group is not None else 0
                emitted.append((grp, int(batch.sv_pos[i]), ps_num, i))
        # kept-but-skipped SVs (empty one-PS set) still have a class on the device
        kept = [r for r in flat if r.svlen >= svlen_thres and r.svread >= suppread_thres and r.gt != "./."]
        for r in kept:
            i = v0 + r.index
            if not traced[i]:
                cls[i] = min(len({rd[2] for rd in r.reads if len(rd) > 1}), 2)
        emitted.sort()
        order += [e[3] for e in emitted]
        g = gt[v0:v1]
        counts[s] = [v1 - v0, len(kept), len(emitted), int((g == 1).sum()), int((g == 2).sum()), int((g == 3).sum()),
                     csr[v1] - csr[v0], int((join_row[csr[v0]:csr[v1]] >= 0).sum())]
    return SimpleNamespace(gt=gt, ps=ps, cls=cls, hap1=ints["hap1"], hap2=ints["hap2"], hap0=ints["hap0"],
                           allhap=ints["allhap"], totsc1=ints["hap1_totsc"], totsc2=ints["hap2_totsc"],
                           features=feats, join_row=join_row, order=np.asarray(order, np.int32),
                           shard_counts=counts, traced=traced)


def _phase_one_shard(flat, chroms, svlen_thres, suppread_thres, trace):
    """ref_port.phase_records for a single contig that accepts every CHROM string in `chroms`."""
    # phase_records matches `r.chrom in ('chr'+ctg, ctg)`; give every record the same CHROM for the
    # duration of the call, then restore
    saved = [r.chrom for r in flat]
    for r in flat:
        r.chrom = "c"
    try:
        rows = ref_port.phase_records(flat, ["c"], svlen_thres, suppread_thres, trace)
    finally:
        for r, c in zip(flat, saved):
            r.chrom = c
    return rows
