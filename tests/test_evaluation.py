"""duet_b200/evaluation.py against what the unmodified reference scorer
(/root/reference/src/scripts/evaluation.py) produced for the same files -- fixture
tests/golden/eval_cases.json.gz, written by tests/golden/make_golden_eval.py."""
import pytest

from conftest import load_golden
from duet_b200 import evaluation


@pytest.fixture(scope="module")
def golden():
    return load_golden("eval_cases.json.gz")


def _write(tmp_path, case):
    paths = {}
    for key, name in (("truth", "truth.vcf"), ("calls", "calls.vcf"), ("bed", "regions.bed")):
        if case.get(key):
            paths[key] = str(tmp_path / name)
            with open(paths[key], "w") as f:
                f.write(case[key])
    return paths


@pytest.mark.parametrize("k", range(6))
def test_scores_equal_the_reference(golden, tmp_path, k):
    case = golden["cases"][k]
    p = _write(tmp_path, case)
    bed = p.get("bed", "")
    base = evaluation.parse_vcf(p["truth"], case["skip_phasing"], bed)
    call = evaluation.parse_vcf(p["calls"], case["skip_phasing"], bed)
    assert base == case["base_info"]                         # same records kept, same fields
    assert call == case["call_info"]
    got = evaluation.evaluation(base, call, case["refdist"], case["ratio"])
    assert [float(x) for x in got] == case["result"]         # the ten numbers, bit for bit


def test_degenerate_inputs_raise_like_the_reference(golden, tmp_path):
    import builtins
    for case in golden["errors"]:
        p = _write(tmp_path, case)
        with pytest.raises(getattr(builtins, case["raises"])):
            evaluation.evaluation(evaluation.parse_vcf(p["truth"], False, ""), evaluation.parse_vcf(p["calls"], False, ""),
                                  1000, 0.0)


def test_cli_prints_the_reference_lines(golden, tmp_path, capsys):
    case = golden["cases"][0]
    p = _write(tmp_path, case)
    res = evaluation.main([p["calls"], p["truth"], "-r", str(case["refdist"]), "-p", str(case["ratio"])])
    assert [float(x) for x in res] == case["result"]
    out = capsys.readouterr().out.splitlines()
    assert out[0].startswith("Average SV number per phase set is")
    assert out[3].startswith("The precision, recall and F1 score of SV phasing are")
