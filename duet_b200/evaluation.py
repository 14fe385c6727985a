"""Regression scorer for phased SV callsets: precision / recall / F1 of calling, genotyping and
phasing against a truth set.  Host-side restatement of the reference's stand-alone script
/root/reference/src/scripts/evaluation.py (`parse_vcf` :34-97, `evaluation` :99-159, CLI :161-189):
same record filters, same nearest-truth matching and the same per-phase-set orientation rule, so the
ten numbers are identical -- SURVEY.md §8(f4).  It is not on the device path; it exists so that a
maintainer can show the phased callset scores the same before and after swapping the stage.

The matching is done on columns (one searchsorted per contig and SV type) instead of the
reference's per-call Python loops.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

CHROMS = [str(i) for i in range(1, 23)] + ["X", "Y"]
_SYMBOLIC = ("<INS>", "<DEL>", "<DUP:TANDEM>", "<DUP:INT>", "<DUP>")
_TYPES = ("INS", "DEL")


def _records(path):
    with open(path) as fh:
        return [line.split() for line in fh]


def parse_bed(path):
    """Per contig of CHROMS: the inclusive [start, end] intervals of its 'chr<name>' rows (:25-32)."""
    spans = [[] for _ in CHROMS]
    where = {"chr" + c: i for i, c in enumerate(CHROMS)}
    for f in _records(path):
        i = where.get(f[0])
        if i is not None:
            spans[i].append((int(f[1]), int(f[2])))
    return [np.asarray(s, np.int64).reshape(-1, 2) for s in spans]


def parse_vcf(vcf_path, skip_phasing, bed_path=""):
    """-> list of {'chr','pos','id','hp','ps','len','type'} (evaluation.py:34-97).

    Kept quirks: contigs must be 'chr'+name (:43); the haplotype is the first three characters of the
    LAST column with '.' read as '0' (:54-60); an unphased record survives only as '1/1' (or with
    skip_phasing) and then belongs to the phase set named after its contig (:61-70); a phased record's
    phase set is contig + '_' + the last column from its last ':' on (:72); DUP counts as INS (:77-78);
    records shorter than 50 bp, '0|0', or outside the BED intervals are dropped (:95-96)."""
    spans = parse_bed(bed_path) if bed_path != "" else None
    out = []
    for f in _records(vcf_path):
        if f[0][0] == "#":
            continue
        contig = f[0][3:]
        if contig not in CHROMS or "SVLEN=." in f[7]:
            continue
        if not any(t in f[7] or t in f[4] for t in ("INS", "DEL", "DUP")):
            continue
        last = f[-1]
        hp = last[:3]
        if hp[0] == ".":
            hp = "0" + hp[1:]
        if hp[2] == ".":
            hp = hp[:2] + "0"
        if hp[1] == "/":
            if not skip_phasing and hp != "1/1":
                continue
            hp, ps = hp[0] + "|" + hp[2], f[0]
        else:
            ps = f[0] + "_" + last[last.rfind(":"):]
        rec = {"chr": f[0], "pos": int(f[1]), "id": f[2] + f[0] + f[1], "hp": hp, "ps": ps}
        if "SVLEN" in f[7]:
            items = f[7].split(";")
            length = next(x for x in items if "SVLEN" in x)
            rec["len"] = abs(int(length[7:] if "SVLEN=>" in f[7] else length[6:]))
            kind = f[4][1:-1] if f[4] in _SYMBOLIC else next(x for x in items if "SVTYPE" in x)[7:]
            rec["type"] = "INS" if "DUP" in kind else kind
        else:
            d = len(f[3]) - len(f[4])
            if d:
                rec["len"], rec["type"] = abs(d), ("DEL" if d > 0 else "INS")
        inside = True
        if spans is not None:
            iv = spans[CHROMS.index(contig)]
            inside = bool(((iv[:, 0] <= rec["pos"]) & (rec["pos"] <= iv[:, 1])).any())
        if not inside:
            continue
        if rec["len"] < 50 or hp == "0|0":                       # KeyError('len') for equal-length alleles, as :95
            continue
        out.append(rec)
    return out


def _codes(values):
    """Small integers for hashable values (equal values -> equal codes)."""
    table = {}
    return np.fromiter((table.setdefault(v, len(table)) for v in values), np.int64, len(values)), len(table)


def evaluation(baseinfo, callinfo, threshold_tp_range, ratio):
    """-> (avg SVs per phase set, P, R, F1 of calling, of genotyping, of phasing) (evaluation.py:99-159).

    A call is matched to the truth record of its contig and type nearest in position (ties towards the
    upper neighbour as np.searchsorted + the :121-127 rule give); it is a true positive if within
    `threshold_tp_range` bp and the length ratio is >= `ratio`.  Genotype agrees if both are
    heterozygous or both homozygous (:132-135).  Phasing is scored per (contig, phase set) in
    whichever orientation -- as called, or with the two haplotypes swapped -- agrees with more
    records (:136-150).  Raises like the reference: IndexError when a contig has calls of a type but
    no truth of it, ZeroDivisionError when nothing matches or the callset is empty."""
    n_call, n_base = len(callinfo), len(baseinfo)
    ps_code, n_ps = _codes([c["ps"] for c in callinfo])
    avg_sv_num = n_call / n_ps
    contig = {"chr" + c: i for i, c in enumerate(CHROMS)}
    hp_of = {"1|0": 1, "0|1": 2, "1|1": 3}

    def columns(info):
        n = len(info)
        return (np.fromiter((contig.get(s["chr"], -1) for s in info), np.int64, n),
                np.fromiter((_TYPES.index(s["type"]) if s["type"] in _TYPES else -1 for s in info), np.int64, n),
                np.fromiter((s["pos"] for s in info), np.int64, n),
                np.fromiter((s["len"] for s in info), np.int64, n),
                np.fromiter((hp_of.get(s["hp"], 0) for s in info), np.int64, n),
                _codes([s["id"] for s in info])[0])

    b_ch, b_ty, b_pos, b_len, b_hp, b_id = columns(baseinfo)
    c_ch, c_ty, c_pos, c_len, c_hp, c_id = columns(callinfo)
    # haplotype strings outside {1|0, 0|1, 1|1} still compare equal to themselves (:136)
    hp_text, _ = _codes([s["hp"] for s in baseinfo] + [s["hp"] for s in callinfo])
    b_hps, c_hps = hp_text[:n_base], hp_text[n_base:]

    m_call, m_base = [], []                                      # matched pairs (indices into the two lists)
    for ch in range(len(CHROMS)):
        for ty in range(len(_TYPES)):
            calls = np.nonzero((c_ch == ch) & (c_ty == ty))[0]
            if calls.size == 0:
                continue
            base = np.nonzero((b_ch == ch) & (b_ty == ty))[0]
            if base.size == 0:
                raise IndexError("list index out of range")     # base[-1] on an empty list (:123)
            base = base[np.argsort(b_pos[base], kind="stable")]
            bp, cp = b_pos[base], c_pos[calls]
            idx = np.searchsorted(bp, cp)
            at_end = idx == base.size
            hi = np.minimum(idx, base.size - 1)
            lo = np.maximum(idx - 1, 0)
            prev = at_end | ((idx > 0) & (np.abs(cp - bp[hi]) > np.abs(cp - bp[lo])))
            pick = base[np.where(prev, lo, hi)]
            near = np.abs(cp - b_pos[pick]) <= threshold_tp_range
            alike = np.minimum(c_len[calls], b_len[pick]) / np.maximum(c_len[calls], b_len[pick]) >= ratio
            ok = near & alike
            m_call.append(calls[ok])
            m_base.append(pick[ok])
    mc = np.concatenate(m_call) if m_call else np.zeros(0, np.int64)
    mb = np.concatenate(m_base) if m_base else np.zeros(0, np.int64)

    def distinct(ids):
        return int(np.unique(ids).size)

    tp_call, tp_base = distinct(c_id[mc]), distinct(b_id[mb])
    het_c, het_b = (c_hp[mc] == 1) | (c_hp[mc] == 2), (b_hp[mb] == 1) | (b_hp[mb] == 2)
    gt = (het_c & het_b) | ((c_hp[mc] == 3) & (b_hp[mb] == 3))
    gt_call, gt_base = distinct(c_id[mc][gt]), distinct(b_id[mb][gt])

    same = c_hps[mc] == b_hps[mb]
    swapped = ((c_hp[mc] == 3) & (b_hp[mb] == 3)) | ((c_hp[mc] == 2) & (b_hp[mb] == 1)) | ((c_hp[mc] == 1) & (b_hp[mb] == 2))
    group = c_ch[mc] * n_ps + ps_code[mc]                        # (contig, phase set)

    def per_group(mask):
        """distinct call ids + distinct truth ids of the masked pairs, per group"""
        g = group[mask]
        score = {}
        for ids in (c_id[mc][mask], b_id[mb][mask]):
            pairs = np.unique(np.stack([g, ids], axis=1), axis=0) if g.size else np.zeros((0, 2), np.int64)
            for k, n in zip(*np.unique(pairs[:, 0], return_counts=True)):
                score[int(k)] = score.get(int(k), 0) + int(n)
        return score

    as_called, as_swapped = per_group(same), per_group(swapped)
    keep_called = np.fromiter((as_called.get(int(g), 0) > as_swapped.get(int(g), 0) for g in group), bool, group.size)
    chosen = np.where(keep_called, same, swapped)
    hp_call, hp_base = distinct(c_id[mc][chosen]), distinct(b_id[mb][chosen])

    def prf(n_c, n_b):
        p, r = n_c / n_call, n_b / n_base
        return p, r, 2 * p * r / (p + r)

    return (avg_sv_num,) + prf(tp_call, tp_base) + prf(gt_call, gt_base) + prf(hp_call, hp_base)


def parse_args(argv):
    parser = argparse.ArgumentParser(description="evaluate SV calling, genotyping and phasing performance")
    parser.add_argument("callset", type=str, help="phased SV callset in .vcf format")
    parser.add_argument("truthset", type=str, help="phased SV truthset in .vcf format")
    parser.add_argument("-r", "--refdist", type=int, default=1000,
                        help="maximum distance comparison calls must be within from base call")
    parser.add_argument("-p", "--pctsim", type=float, default=0,
                        help="minimum length ratio between base and comparison call")
    parser.add_argument("-b", "--bed_file", type=str, help="optional .bed file to confine benchmark regions")
    parser.add_argument("--skip_phasing", action="store_true", help="only benchmark on SV calling and genotyping")
    return parser.parse_args(argv)


def main(argv=None):
    args = parse_args(sys.argv[1:] if argv is None else argv)
    bed = args.bed_file or ""
    res = evaluation(parse_vcf(args.truthset, args.skip_phasing, bed), parse_vcf(args.callset, args.skip_phasing, bed),
                     args.refdist, args.pctsim)
    if not args.skip_phasing:
        print("Average SV number per phase set is", res[0])
    print("The precision, recall and F1 score of SV calling are", *res[1:4])
    print("The precision, recall and F1 score of SV genotyping are", *res[4:7])
    if not args.skip_phasing:
        print("The precision, recall and F1 score of SV phasing are", *res[7:10])
    return res


if __name__ == "__main__":
    main()
