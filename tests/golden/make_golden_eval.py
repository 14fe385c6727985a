#!/usr/bin/env python
"""Golden fixture for duet_b200/evaluation.py: truth/call VCF pairs (text) and what the UNMODIFIED
reference scorer (/root/reference/src/scripts/evaluation.py) makes of them -- only runnable in the
build container.

    python tests/golden/make_golden_eval.py      ->  tests/golden/eval_cases.json.gz

The tests read only the .json.gz file."""
import gzip
import json
import os
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src/scripts")

CONTIGS = ["chr%s" % c for c in list(range(1, 23)) + ["X", "Y"]]


def truth_and_calls(seed, n_events=400, contigs=CONTIGS[:6] + ["chrX"], extra_contigs=("chrUn_1", "7")):
    """A truth set and a perturbed call set: jittered positions and lengths, flipped / wrong haplotypes,
    misses, false calls, and the dialect quirks the scorer handles (SVLEN=>, SVLEN=., DUP -> INS,
    sequence-resolved records without SVLEN, unphased and half-missing genotypes)."""
    rng = np.random.default_rng(seed)
    truth, calls = [], []
    for k in range(n_events):
        ch = contigs[int(rng.integers(len(contigs)))]
        pos = int(rng.integers(10_000, 3_000_000))
        ln = int(rng.integers(30, 4000))
        ty = ["INS", "DEL", "DUP", "INV"][int(rng.choice(4, p=[0.42, 0.42, 0.1, 0.06]))]
        hp = ["1|0", "0|1", "1|1", "0/1", "1/1", "./1", "0|0", ".|1"][int(rng.choice(8, p=[.3, .3, .2, .06, .06, .03, .03, .02]))]
        ps = pos // 150_000 * 150_000 + 1                        # phase blocks shared by neighbouring SVs
        style = int(rng.integers(5))
        if style == 0:
            info, ref, alt = f"SVTYPE={ty};SVLEN={-ln if ty == 'DEL' else ln};END={pos + ln}", "N", f"<{ty}>"
        elif style == 1:
            info, ref, alt = f"SVLEN={ln};SVTYPE={ty}", "N", "<DUP:TANDEM>" if ty == "DUP" else f"<{ty}>"
        elif style == 2 and ty in ("INS", "DEL"):
            seq = "ACGT" * (ln // 4 + 1)
            ref, alt = ("A" + seq[:ln], "A") if ty == "DEL" else ("A", "A" + seq[:ln])
            info = f"SVTYPE={ty}"
        elif style == 3:
            info, ref, alt = f"SVTYPE={ty};SVLEN=>{ln}", "N", f"<{ty}>"
        else:
            info, ref, alt = f"PRECISE;SVTYPE={ty};SVLEN={ln}", "N", "ACGTTGCA"[: 1 + k % 7]
        if rng.random() < 0.02:
            info = f"SVTYPE={ty};SVLEN=."
        truth.append(f"{ch}\t{pos}\ttruth{k}\t{ref}\t{alt}\t.\tPASS\t{info}\tGT:PS\t{hp}:{ps}")
        if rng.random() < 0.8:                                   # called, somewhere near
            cpos = pos + int(rng.integers(-1500, 1500)) if rng.random() < 0.3 else pos + int(rng.integers(-40, 40))
            cln = max(1, int(ln * rng.uniform(0.5, 1.5))) if rng.random() < 0.3 else ln
            r = rng.random()
            flipped = zlib.crc32(f"{ch}:{ps}:{seed}".encode()) % 10 < 3          # a whole block called in the other orientation
            chp = hp[:3].replace("/", "|") if r < 0.8 else ["1|0", "0|1", "1|1"][int(rng.integers(3))]
            if flipped:
                chp = {"1|0": "0|1", "0|1": "1|0"}.get(chp, chp)
            cty = "INS" if ty == "DUP" else ty
            cps = ps if rng.random() < 0.9 else ps + 77
            calls.append(f"{ch}\t{cpos}\tDuet.{len(calls) + 1}\tN\t<{cty}>\t.\tPASS\tSVLEN={-cln if cty == 'DEL' else cln};SVTYPE={cty}\tHP:PS\t{chp}:{cps}")
    for k in range(n_events // 10):                               # false calls + contigs the scorer ignores
        ch = (list(contigs) + list(extra_contigs))[int(rng.integers(len(contigs) + len(extra_contigs)))]
        pos = int(rng.integers(10_000, 3_000_000))
        ty = ["INS", "DEL"][int(rng.integers(2))]
        calls.append(f"{ch}\t{pos}\tDuet.{len(calls) + 1}\tN\t<{ty}>\t.\tPASS\tSVLEN={int(rng.integers(30, 900))};SVTYPE={ty}\tHP:PS\t{['1|0', '0|1', '1|1'][k % 3]}:{pos}")
    header = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n"
    bed = []
    for ch in contigs:
        for _ in range(6):
            a = int(rng.integers(0, 2_800_000))
            bed.append(f"{ch}\t{a}\t{a + int(rng.integers(50_000, 400_000))}")
    return header + "\n".join(truth) + "\n", header + "\n".join(calls) + "\n", "\n".join(bed) + "\n"


def main():
    import evaluation as ref                                       # the reference scorer, unmodified
    cases = []
    work = tempfile.mkdtemp(prefix="duet_eval_")
    for seed, (skip, refdist, ratio, use_bed) in enumerate([(False, 1000, 0.0, False), (False, 500, 0.7, False),
                                                            (True, 1000, 0.0, False), (False, 1000, 0.5, True),
                                                            (True, 200, 0.9, True), (False, 1000, 0.0, False)]):
        t, c, b = truth_and_calls(seed, n_events=400 if seed < 5 else 60, contigs=CONTIGS[:6] + ["chrX"] if seed < 5 else CONTIGS[:2])
        paths = {}
        for name, text in (("truth.vcf", t), ("calls.vcf", c), ("regions.bed", b)):
            paths[name] = os.path.join(work, f"{seed}_{name}")
            with open(paths[name], "w") as f:
                f.write(text)
        bed = paths["regions.bed"] if use_bed else ""
        base = ref.parse_vcf(paths["truth.vcf"], skip, bed)
        call = ref.parse_vcf(paths["calls.vcf"], skip, bed)
        out = ref.evaluation(base, call, refdist, ratio)
        cases.append({"truth": t, "calls": c, "bed": b if use_bed else "", "skip_phasing": skip, "refdist": refdist,
                      "ratio": ratio, "base_info": base, "call_info": call, "result": [float(x) for x in out]})
        print(seed, len(base), len(call), [round(float(x), 4) for x in out])
    # what the reference raises on degenerate inputs
    errors = []
    hdr = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n"
    for name, t, c in (
            ("no_base_of_that_type", hdr + "chr1\t100\tt\tN\t<DEL>\t.\tPASS\tSVTYPE=DEL;SVLEN=-90\tGT:PS\t1|0:5\n",
             hdr + "chr1\t100\tc\tN\t<INS>\t.\tPASS\tSVTYPE=INS;SVLEN=90\tGT:PS\t1|0:5\n"),
            ("equal_length_alleles", hdr + "chr1\t100\tt\tACG\tTGA\t.\tPASS\tSVTYPE=INS\tGT:PS\t1|0:5\n", hdr),
            ("empty_callset", hdr + "chr1\t100\tt\tN\t<DEL>\t.\tPASS\tSVTYPE=DEL;SVLEN=-90\tGT:PS\t1|0:5\n", hdr),
            ("nothing_matches", hdr + "chr1\t100\tt\tN\t<DEL>\t.\tPASS\tSVTYPE=DEL;SVLEN=-90\tGT:PS\t1|0:5\n",
             hdr + "chr1\t900000\tc\tN\t<DEL>\t.\tPASS\tSVTYPE=DEL;SVLEN=-90\tGT:PS\t1|0:5\n")):
        pt, pc = os.path.join(work, name + "_t.vcf"), os.path.join(work, name + "_c.vcf")
        open(pt, "w").write(t); open(pc, "w").write(c)
        try:
            ref.evaluation(ref.parse_vcf(pt, False, ""), ref.parse_vcf(pc, False, ""), 1000, 0.0)
            raised = None
        except Exception as e:                                    # noqa: BLE001
            raised = type(e).__name__
        errors.append({"name": name, "truth": t, "calls": c, "raises": raised})
        print(name, raised)
    with gzip.open(os.path.join(HERE, "eval_cases.json.gz"), "wt") as f:
        json.dump({"cases": cases, "errors": errors}, f)


if __name__ == "__main__":
    main()
