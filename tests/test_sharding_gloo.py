"""Host-side multi-GPU plumbing on CPU: world_size-2 gloo processes run the contig-sharded stage
with the oracle standing in for the device (phase_fn injection) and must produce the
byte-identical phased_sv.vcf the unmodified reference produced, plus the same counter table."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT, load_golden, materialise
from duet_b200 import sharding


def test_lpt_assign_balances_and_is_deterministic():
    w = [249, 243, 198, 191, 181, 171, 159, 146, 141, 135, 135, 133, 115, 107, 102, 90, 81, 78, 59, 63, 48, 51, 155, 59]
    for n in (1, 2, 4, 8):
        plan = sharding.lpt_assign(w, n)
        assert sorted(i for p in plan for i in p) == list(range(24))
        loads = [sum(w[i] for i in p) for p in plan]
        assert max(loads) <= sum(w) / n * 1.12 + 1
        assert plan == sharding.lpt_assign(w, n)


def test_lpt_assign_spreads_zero_weight_items():
    """One BAM present, 23 contigs without one (the chr21 demo): the idle ranks share the empty contigs
    instead of the first idle rank taking them all; with more ranks than contigs some plans are empty."""
    w = [0] * 24
    w[20] = 5000
    plan = sharding.lpt_assign(w, 4)
    assert sorted(i for p in plan for i in p) == list(range(24))
    assert [20] in plan and max(len(p) for p in plan) <= 8 and min(len(p) for p in plan) >= 1
    plan = sharding.lpt_assign([7, 3], 4)
    assert sorted(len(p) for p in plan) == [0, 0, 1, 1]


def test_slice_first_ids():
    assert sharding.slice_first_ids([["1"], ["10"], [], ["chr1"]], [5, 7, 0, 2]) == [1, 6, 0, 13]
    assert sharding.slice_first_ids([["1"], ["1"]], [1, 1]) is None
    assert sharding.slice_first_ids([["1", "chr1"]], [3]) is None


def _oracle_phase_fn(batch, svlen_thres, suppread_thres):
    from duet_b200.engine import PhaseResult
    from oracle.columnar_adapter import phase_batch_oracle
    o = phase_batch_oracle(batch, svlen_thres, suppread_thres)
    return PhaseResult(o.gt, o.ps, o.cls, o.hap1, o.hap2, o.hap0, o.allhap, o.totsc1, o.totsc2, o.features,
                       o.join_row, o.order, o.shard_counts)


def _worker(rank, world, port, home, svlen, supp, out_q, on_gpu=False, inc=False):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if on_gpu:
        import torch
        os.environ["DUET_DEVICE"] = str(rank % torch.cuda.device_count())
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        counts = sharding.sv_phasing_sharded(home, svlen, supp, 1, inc,
                                             phase_fn=None if on_gpu else _oracle_phase_fn)
        if on_gpu:
            import torch
            from duet_b200 import sv_phasing_fn
            used = sorted(sv_phasing_fn._ENGINES)
            assert used == [rank % torch.cuda.device_count()], used      # >= 2 GPUs visible: every rank on its own
        out_q.put((rank, counts))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_two_ranks(name, tmp_path, on_gpu, world=2):
    case = load_golden(f"e2e_{name}.json.gz")
    home = materialise(case, str(tmp_path))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, home, case["svlen_thres"], case["suppread_thres"], q, on_gpu,
                                               case.get("include_all_ctgs", False)))
             for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    with open(home + "/phased_sv.vcf") as f:
        assert f.read() == case["phased_sv_vcf"]
    for r in range(1, world):
        assert np.array_equal(got[0], got[r])                  # every rank holds the whole counter table
    assert got[0][:, 2].sum() == len(case["rows"])
    assert not [fn for fn in os.listdir(home) if ".slice." in fn]


@pytest.mark.parametrize("name", ["cutesv_3ctg", "mixed_prefix", "svim_shuffled", "all_ctgs"])
def test_two_rank_stage_is_byte_identical(name, tmp_path):
    _run_two_ranks(name, tmp_path, on_gpu=False)


def test_more_ranks_than_contigs(tmp_path):
    """`all_ctgs` lists three contigs (tabix shim): with four ranks one owns nothing, must not call the
    device with an empty batch, and must still take part in every collective."""
    _run_two_ranks("all_ctgs", tmp_path, on_gpu=False, world=4)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cutesv_3ctg", "dense"])
def test_two_rank_stage_on_gpu(name, tmp_path):
    """Same, with the real device path in both ranks: each rank takes its own GPU when >= 2 are visible
    (asserted inside the workers through DUET_DEVICE); on a 1-GPU box both share it."""
    _run_two_ranks(name, tmp_path, on_gpu=True)


def test_single_process_path_matches_too(tmp_path):
    case = load_golden("e2e_sniffles_chr.json.gz")
    home = materialise(case, str(tmp_path))
    counts = sharding.sv_phasing_sharded(home, case["svlen_thres"], case["suppread_thres"], 1, False,
                                         phase_fn=_oracle_phase_fn, rank=0, world=1)
    with open(home + "/phased_sv.vcf") as f:
        assert f.read() == case["phased_sv_vcf"]
    assert counts[:, 2].sum() == len(case["rows"])
