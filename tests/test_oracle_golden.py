"""The oracle port (oracle/ref_port.py) against the outputs of the UNMODIFIED
reference recorded in tests/golden/ (see tests/golden/make_golden.py)."""
import pytest

from conftest import golden_e2e_names, load_golden
from oracle import ref_port

FLOAT_KEYS = ("hapread_ratio", "sv_ratio", "hap1_avgsc", "hap2_avgsc", "totsc_ratio", "hap_avgsc_diff")
INT_KEYS = ("hap1", "hap2", "hap0", "allhap", "nohap", "ps", "hap1_totsc", "hap2_totsc", "onehap_totsc")


def _reads(case_reads):
    return [tuple(r) for r in case_reads]


def test_kat_phase_info():
    cases = load_golden("kat_phase_info.json.gz")
    assert len(cases) > 1500
    n_raise = 0
    preds = set()
    for c in cases:
        args = (_reads(c["reads"]), c["pos"], c["svread"], c["refread"], c["ps_num"], set(c["oneps"]))
        if "raises" in c:
            n_raise += 1
            with pytest.raises(Exception) as ei:
                ref_port.predict(*args)
            assert type(ei.value).__name__ == c["raises"], c["note"]
            continue
        pred, ps, f = ref_port.predict(*args)
        assert pred == c["pred"], c
        assert ps == c["ps"], c
        preds.add((c["ps_num"], pred))
        for k, v in c["features"].items():
            assert f[k] == v, (k, c)       # floats included: same IEEE operations, bit-equal
    assert n_raise >= 3
    # every (class, genotype) outcome the tree can produce is present in the fixture
    assert preds >= {(0, 0), (0, 3), (1, 0), (1, 1), (1, 2), (1, 3), (2, 0), (2, 3)}


@pytest.mark.parametrize("name", golden_e2e_names())
def test_e2e_against_reference(name, golden_workdir):
    case, home = golden_workdir(name)
    inc = case.get("include_all_ctgs", False)
    vcf = home + "/sv_calling/variants.vcf"
    sam_home = home + "/snp_phasing/"
    # the join
    tables = ref_port.haplotag_tables(sam_home, 1, inc)
    flat = ref_port.join_support_reads(ref_port.sv_records(vcf, inc), tables)
    assert len(flat) == len(case["joined"])
    for r, g in zip(flat, case["joined"]):
        assert (r.chrom, r.pos, r.svlen, r.svtype, r.svread, r.refread, r.gt) == \
               (g["chrom"], g["pos"], g["svlen"], g["svtype"], g["svread"], g["refread"], g["callgt"])
        assert [list(x[1:]) for x in r.reads] == g["reads"]
    # per-SV features in the reference's evaluation order, then the rows
    trace = []
    rows = ref_port.generate_phased_callset(vcf, sam_home, case["svlen_thres"], case["suppread_thres"], 1,
                                            inc, trace=trace)
    assert len(trace) == len(case["trace"])
    for (ci, ps_num, r, pred, f), g in zip(trace, case["trace"]):
        assert (r.chrom, r.pos, ps_num) == (g["chrom"], g["pos"], g["ps_num"])
        for k in INT_KEYS + FLOAT_KEYS:
            assert f[k] == g["f"][k], (k, g)
    assert rows == case["rows"]
    # the output file, byte for byte
    ref_port.sv_phasing(home, case["svlen_thres"], case["suppread_thres"], 1, inc)
    with open(home + "/phased_sv.vcf") as fh:
        assert fh.read() == case["phased_sv_vcf"]
