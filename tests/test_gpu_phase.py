"""Parity of the CUDA path (through the C-ABI) against the oracle port and the golden
fixtures recorded from the unmodified reference.  Bit-exact: genotype, PS, class, counts,
score sums, join rows, AND the fp64 features (same IEEE operations in the same order)."""
import numpy as np
import pytest

from conftest import load_golden
from duet_b200 import _lib, synth
from duet_b200.columnar import from_synth
from duet_b200.engine import DuetError, PhaseEngine, pin_batch
from oracle import synth_adapter
from util_batches import BatchBuilder, assert_matches_trace, kat_class, kat_shard

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    eng = PhaseEngine(0)
    yield eng
    eng.close()


def run_and_compare(engine, sample, svlen=50, supp=2, **kw):
    batch = from_synth(sample, **kw)
    engine.set_thresholds(svlen, supp)
    res = engine.run(batch)
    trace = []
    rows, flat = synth_adapter.phase_sample(sample, svlen, supp, trace)
    assert_matches_trace(res, batch, trace, flat)
    assert res.rows(batch) == rows
    # counters: n_sv, n_kept, n_emitted, genotype split, joins, hits
    c = res.shard_counts.sum(axis=0)
    assert c[0] == batch.n_svs and c[2] == len(rows) == res.order.shape[0]
    assert c[3] == sum(r["hp"] == "1|0" for r in rows) and c[4] == sum(r["hp"] == "0|1" for r in rows)
    assert c[5] == sum(r["hp"] == "1|1" for r in rows)
    assert c[6] == batch.n_joins and c[7] == int((res.join_row >= 0).sum())
    return batch, res, rows


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_multi_contig_small(engine, seed):
    s = synth.make_sample(seed, contigs=["1", "2", "7", "21", "X", "Y"], n_reads=12000, n_svs=900,
                          bp_per_read=700, block_mean=1.2e5, shuffle_vcf=(seed % 2 == 1),
                          chr_prefix=(seed >= 2))
    _, _, rows = run_and_compare(engine, s)
    assert len(rows) > 100
    assert {r["hp"] for r in rows} == {"1|0", "0|1", "1|1"}


def test_thresholds_change_result(engine):
    s = synth.make_sample(5, contigs=["3", "4"], n_reads=6000, n_svs=500, bp_per_read=700, block_mean=1e5)
    _, _, rows_a = run_and_compare(engine, s, 50, 2)
    _, _, rows_b = run_and_compare(engine, s, 500, 12)
    assert 0 < len(rows_b) < len(rows_a)


def test_dense_support_lists(engine):
    s = synth.make_sample(6, contigs=["20", "21"], n_reads=30000, n_svs=300, dense=True, bp_per_read=350,
                          block_mean=3e5, empty_oneps_contig=None)
    batch, res, _ = run_and_compare(engine, s)
    assert np.diff(batch.csr_off).max() > 300


def test_c1_shape_chr21(engine):
    """BASELINE.json configs[0]: chr21 demo shape (70 k reads, 2.5 k SVs)."""
    batch, res, rows = run_and_compare(engine, synth.config_c1(0))
    assert batch.n_svs == 2500 and len(rows) > 800


def test_without_hi_words(engine):
    s = synth.make_sample(8, contigs=["1", "2"], n_reads=4000, n_svs=300, bp_per_read=700, block_mean=1e5)
    run_and_compare(engine, s, with_hi=False)


def test_pinned_input_same_result(engine):
    s = synth.make_sample(9, contigs=["1", "2"], n_reads=4000, n_svs=300, bp_per_read=700, block_mean=1e5)
    batch = from_synth(s)
    engine.set_thresholds(50, 2)
    a = engine.run(batch)
    b = engine.run(pin_batch(batch))
    for k in ("gt", "ps", "cls", "hap1", "hap2", "join_row", "order"):
        assert np.array_equal(getattr(a, k), getattr(b, k))
    # execute is repeatable on staged columns (the bench loop relies on it)
    engine.execute(); engine.execute()
    c = engine.download()
    assert np.array_equal(a.gt, c.gt) and np.array_equal(a.ps, c.ps) and np.array_equal(a.order, c.order)


def test_tags_read_in_place_from_pinned_host_memory(engine):
    """DUET_MEM_HOST_MAPPED: the tag records are not copied, the kernels read the joined rows' records
    over the bus -- same results; pageable memory is refused."""
    s = synth.make_sample(10, contigs=["1", "2", "X"], n_reads=9000, n_svs=700, bp_per_read=700, block_mean=1e5)
    batch = from_synth(s)
    engine.set_thresholds(50, 2)
    a = engine.run(batch)
    pinned = pin_batch(batch)
    b = engine.run(pinned, tags_in_place=True)
    c = engine.run(pinned, tags_in_place=True)                  # replay with the same host buffer
    d = engine.run(batch)                                        # and back to copies
    for k in ("gt", "ps", "cls", "hap1", "hap2", "hap0", "allhap", "totsc1", "totsc2", "join_row", "order", "shard_counts"):
        for other in (b, c, d):
            assert np.array_equal(getattr(a, k), getattr(other, k)), k
    assert np.array_equal(a.features, b.features)
    with pytest.raises(DuetError) as ei:
        engine.run(batch, tags_in_place=True)                    # numpy-owned (pageable) memory
    assert ei.value.code == _lib.ERR_INVALID
    import dataclasses
    from duet_b200.engine import pinned_empty
    raw = pinned_empty(batch.read_tag.nbytes + 8, np.uint8)     # page-locked but 8 bytes off a 16-byte boundary
    skew = raw[8:8 + batch.read_tag.nbytes].view(batch.read_tag.dtype)
    skew[...] = batch.read_tag
    with pytest.raises(DuetError) as ei:
        engine.run(dataclasses.replace(pinned, read_tag=skew), tags_in_place=True)
    assert ei.value.code == _lib.ERR_INVALID and "aligned" in str(ei.value)
    e = engine.run(pinned, tags_in_place=True)                   # the handle is still usable
    assert np.array_equal(a.gt, e.gt) and np.array_equal(a.order, e.order)


def test_pipeline_two_calls_in_flight_equals_sequential(engine):
    """PhasePipeline (cohort mode): five different samples through two engines with two calls in flight give, in
    order, what the single engine gives one call at a time -- with pageable columns, with page-locked columns read
    in place, and with results landing in per-engine page-locked buffers (copied before the engine comes round)."""
    from duet_b200.engine import PhasePipeline, pinned_outputs
    samples = [synth.make_sample(20 + k, contigs=["1", "2", "X"][: 1 + k % 3], n_reads=3000 + 900 * k, n_svs=200 + 60 * k,
                                 bp_per_read=700, block_mean=1e5) for k in range(5)]
    batches = [from_synth(s) for s in samples]
    engine.set_thresholds(50, 2)
    want = [engine.run(b) for b in batches]
    keys = ("gt", "ps", "cls", "hap1", "hap2", "hap0", "allhap", "totsc1", "totsc2", "join_row", "order", "shard_counts", "features")
    pipe = PhasePipeline(0, 2)
    try:
        pipe.set_thresholds(50, 2)
        got = list(pipe.run_many(batches, tags_in_place=False))
        assert len(got) == len(want)
        for a, b in zip(want, got):
            for k in keys:
                assert np.array_equal(getattr(a, k), getattr(b, k)), k
        pinned = [pin_batch(b) for b in batches]
        i_big = max(range(len(batches)), key=lambda i: (batches[i].n_svs, batches[i].n_joins))
        for k, r in enumerate(pipe.run_many(pinned, tags_in_place=True)):
            for name in keys:
                assert np.array_equal(getattr(want[k], name), getattr(r, name)), name
        # same-shape batches may share per-engine result buffers; a result is valid until its engine's next download
        bufs = [pinned_outputs(batches[i_big]), pinned_outputs(batches[i_big])]
        same = [pinned[i_big]] * 4
        for r in pipe.run_many(same, buffers=bufs, tags_in_place=True):
            w = want[i_big]
            assert np.array_equal(w.gt, r.gt) and np.array_equal(w.ps, r.ps) and np.array_equal(w.order, r.order)
    finally:
        pipe.close()


def test_two_branch_chain_equals_serial_chain(engine, monkeypatch):
    """Calls with >= 1 M support-read names take the two-branch chain (k_bloom -> k_stream beside k_init -> k_table,
    meeting at k_resolve; the full-size C4 / C5 parity tests run it).  DUET_FLAGS=64 forces it for small calls too:
    same results as the serial chain, call after call (the filter is handed back clean by k_reduce)."""
    from duet_b200.engine import PhaseEngine
    monkeypatch.setenv("DUET_FLAGS", "64")
    forced = PhaseEngine(0)
    monkeypatch.delenv("DUET_FLAGS")
    try:
        forced.set_thresholds(50, 2)
        engine.set_thresholds(50, 2)
        keys = ("gt", "ps", "cls", "hap1", "hap2", "hap0", "allhap", "totsc1", "totsc2", "join_row", "order", "shard_counts", "features")
        for k in range(4):
            s = synth.make_sample(40 + k, contigs=["1", "2", "X", "21"][: 1 + k], n_reads=5000 + 3000 * k, n_svs=300 + 150 * k,
                                  bp_per_read=700, block_mean=1e5, dense=(k == 3))
            batch = from_synth(s)
            want = engine.run(batch)
            for _ in range(3):
                before = forced.launch_count()
                got = forced.run(batch)
                for name in keys:
                    assert np.array_equal(getattr(want, name), getattr(got, name)), name
                # k_init, k_table, k_bloom, k_stream, k_resolve, k_reduce, k_tail (+ k_reduce_heavy when a dense batch
                # has support lists of more than 256 reads)
                assert forced.launch_count() - before in ((7, 8) if k == 3 else (7,))
    finally:
        forced.close()


def test_golden_kat_on_device(engine):
    """Every known-answer case recorded from the reference's get_phase_info / predict_hp whose
    class is reachable through the pipeline, run as one shard each in ONE device call."""
    cases = [c for c in load_golden("kat_phase_info.json.gz") if "raises" not in c and kat_class(c) == c["ps_num"]]
    assert len(cases) > 900
    bb = BatchBuilder()
    for k, c in enumerate(cases):
        kat_shard(bb, c, f"k{k}")
    batch = bb.build()
    engine.set_thresholds(0, 0)
    res = engine.run(batch)
    outcomes = set()
    for k, c in enumerate(cases):
        i = int(batch.sv_off[k])
        assert res.cls[i] == c["ps_num"]
        assert res.gt[i] == c["pred"], (c, res.gt[i])
        assert res.ps[i] == c["ps"], (c, res.ps[i])
        f = c["features"]
        assert (res.hap1[i], res.hap2[i], res.hap0[i], res.allhap[i]) == (f["hap1"], f["hap2"], f["hap0"], f["allhap"])
        assert (res.totsc1[i], res.totsc2[i]) == (f["hap1_totsc"], f["hap2_totsc"])
        for name in _lib.FEATURE_NAMES:
            assert float(res.feature(name)[i]) == float(f[name]), (name, c)
        outcomes.add((c["ps_num"], c["pred"]))
    assert outcomes >= {(0, 0), (0, 3), (1, 0), (1, 1), (1, 2), (1, 3), (2, 0), (2, 3)}


def test_golden_kat_errors(engine):
    """KeyError (:96) and ZeroDivisionError (:123) of the reference surface as status codes."""
    want = {"KeyError": _lib.ERR_BAD_HP, "ZeroDivisionError": _lib.ERR_ZERO_DIVISION}
    cases = [c for c in load_golden("kat_phase_info.json.gz") if "raises" in c]
    assert {c["raises"] for c in cases} == set(want)
    engine.set_thresholds(0, 0)
    for k, c in enumerate(cases):
        bb = BatchBuilder()
        if c["raises"] == "KeyError":      # force class 2 with an extra phase set
            c = dict(c, reads=c["reads"] + [["x", 1, 4242, 0]])
        kat_shard(bb, c, f"e{k}")
        with pytest.raises(DuetError) as ei:
            engine.run(bb.build())
        assert ei.value.code == want[c["raises"]]


def test_last_row_wins_and_duplicate_names(engine):
    """duplicate QNAME -> the LAST row's tag (:29); a name listed twice counts twice (:46-48);
    the same name in another contig's BAM is a different read."""
    bb = BatchBuilder()
    bb.shard([("dup", 1, 500, 100), ("q2", 1, 500, 7), ("dup", 2, 500, 300), ("solo", 1, 500, 1)],
             [dict(pos=10, svread=5, refread=0, names=["dup", "q2", "dup", "ghost"]),
              dict(pos=20, svread=5, refread=0, names=["solo"])], contig="1")
    bb.shard([("dup", 1, 900, 11)], [dict(pos=10, svread=5, refread=0, names=["dup", "q2"])], contig="2")
    batch = bb.build()
    engine.set_thresholds(0, 0)
    res = engine.run(batch)
    assert res.join_row.tolist() == [2, 1, 2, -1, 3, 4, -1]
    assert (res.hap1[0], res.hap2[0], res.totsc1[0], res.totsc2[0]) == (1, 2, 7, 600)
    assert res.feature("hapread_ratio")[0] == 3 / 4
    assert res.ps[2] == 900 and res.hap1[2] == 1


def test_empty_oneps_contig_and_empty_shards(engine):
    bb = BatchBuilder()
    bb.shard([], [], contig="1")                                            # nothing at all
    bb.shard([("a", 1, 100, 9000)], [dict(pos=5, svread=9, refread=0, names=["a", "b"])], contig="2")  # pc>8100 only
    bb.shard([("c", 1, 100, 5)], [], contig="3")                            # reads but no SVs
    bb.shard([], [dict(pos=5, svread=9, refread=0, names=["zz"])], contig="4")   # SVs but no reads
    bb.shard([("d", 2, 300, 5)], [dict(pos=5, svread=9, refread=0, names=["d"])], contig="5")
    batch = bb.build()
    engine.set_thresholds(0, 0)
    res = engine.run(batch)
    assert res.gt.tolist() == [0, 0, 3]
    assert res.order.tolist() == [2]
    assert res.shard_counts[:, 2].tolist() == [0, 0, 0, 0, 1]


def test_tie_order_and_filters(engine):
    """Rows tying on (chrom, pos) come out class 0 first, then class 1, then class 2, VCF order
    inside a class (:206-229); svlen / support / GT filters (:189-190)."""
    reads = [("a", 1, 100, 5), ("b", 2, 100, 5), ("c", 1, 200, 5), ("h", 1, 200, 0)]
    svs = [dict(pos=700, svread=9, refread=0, names=["a", "c"]),            # class 2
           dict(pos=700, svread=9, refread=0, names=["a", "b"]),            # class 1
           dict(pos=700, svread=9, refread=0, names=["nope"]),              # class 0
           dict(pos=700, svread=9, refread=0, names=["a"]),                 # class 1 (later in VCF)
           dict(pos=600, svread=9, refread=0, names=["h"]),                 # pins 200 into the one-PS set
           dict(pos=1, svread=9, refread=0, names=["a"], svlen=10),         # svlen filter
           dict(pos=1, svread=1, refread=0, names=["a"]),                   # support filter
           dict(pos=1, svread=9, refread=0, names=["a"], gt_missing=True)]  # GT ./.
    batch = BatchBuilder().shard(reads, svs).build()
    engine.set_thresholds(50, 2)
    res = engine.run(batch)
    assert res.cls.tolist() == [2, 1, 0, 1, 1, 255, 255, 255]
    assert res.order.tolist() == [4, 2, 1, 3, 0]


@pytest.mark.parametrize("n_ps,n_reads", [(5, 400), (8, 400), (9, 400), (12, 400), (16, 400), (17, 400), (33, 400), (45, 400),
                                          (16, 4000), (17, 4000), (32, 4000), (33, 4000), (45, 4000)])
def test_many_phase_sets_in_one_sv(engine, n_ps, n_reads):
    """k_reduce records up to 16 distinct PS per SV (32 in dense batches, one warp per SV: the 4000-read cases);
    more take the warp-cooperative fallback with its shared-memory list of 32; beyond that the exact quadratic
    path runs.  All must agree with the oracle."""
    rng = np.random.default_rng(n_ps)
    pss = list(range(1000, 1000 + n_ps * 10, 10))
    reads, names = [], []
    for k in range(n_reads):
        ps = int(rng.choice(pss)) if k >= n_ps else pss[k]
        reads.append((f"r{k}", int(rng.integers(1, 3)), ps, int(rng.integers(0, 9000))))
        names.append(f"r{k}")
    helpers, hsv = [], []
    for k, ps in enumerate(pss[3:]):
        helpers.append((f"h{k}", 1, ps, 0))
        hsv.append(dict(pos=1, svread=2, refread=5, names=[f"h{k}"]))
    bb = BatchBuilder().shard(reads + helpers, [dict(pos=1234, svread=40, refread=3, names=names)] + hsv)
    batch = bb.build()
    engine.set_thresholds(0, 0)
    res = engine.run(batch)
    from oracle import ref_port
    joined = [(n, h, p, c) for (n, h, p, c) in reads]
    pred, ps, f = ref_port.predict(joined, 1234, 40, 3, 2, set(pss[3:]))
    assert (res.gt[0], res.ps[0], res.hap1[0], res.hap2[0], res.hap0[0], res.allhap[0]) == \
           (pred, ps, f["hap1"], f["hap2"], f["hap0"], f["allhap"])
    assert (res.totsc1[0], res.totsc2[0]) == (f["hap1_totsc"], f["hap2_totsc"])


def test_hash_collision_is_reported(engine):
    bb = BatchBuilder().shard([("a", 1, 100, 5), ("b", 1, 100, 5)],
                              [dict(pos=5, svread=9, refread=0, names=["a", "b"])])
    batch = bb.build()
    batch.read_key[1] = batch.read_key[0]          # 'b' now collides with 'a' on the 64-bit key only
    engine.set_thresholds(0, 0)
    with pytest.raises(DuetError) as ei:
        engine.run(batch)
    assert ei.value.code == _lib.ERR_HASH_COLLISION


def test_big_shard_uses_global_sort_scratch(engine):
    """One shard with more SVs than the shared-memory sort tiles hold (16384 / 8192)."""
    s = synth.make_sample(12, contigs=["1"], n_reads=60000, n_svs=20000, bp_per_read=700, block_mean=2e5)
    run_and_compare(engine, s)


# ---- the drop-in functions on the golden end-to-end cases recorded from the reference ----------
from conftest import golden_e2e_names  # noqa: E402


@pytest.mark.parametrize("name", golden_e2e_names())
def test_dropin_against_reference_golden(name, golden_workdir):
    """generate_phased_callset / sv_phasing with the reference's signature: rows equal and
    phased_sv.vcf byte-identical to what the unmodified reference produced."""
    from duet_b200 import sv_phasing, sv_phasing_fn
    case, home = golden_workdir(name)
    inc = case.get("include_all_ctgs", False)
    rows = sv_phasing_fn.generate_phased_callset(home + "/sv_calling/variants.vcf", home + "/snp_phasing/",
                                                 case["svlen_thres"], case["suppread_thres"], 1, inc)
    assert rows == case["rows"]
    sv_phasing.sv_phasing(home, case["svlen_thres"], case["suppread_thres"], 1, inc)
    with open(home + "/phased_sv.vcf") as f:
        assert f.read() == case["phased_sv_vcf"]
    # and the join / per-SV features the reference computed on the way
    batch = sv_phasing_fn.generate_callinfo(home + "/sv_calling/variants.vcf",
                                            sv_phasing_fn.read_hap_bam(home + "/snp_phasing/", 1, inc), inc)
    res = sv_phasing_fn.phase_batch(batch, case["svlen_thres"], case["suppread_thres"])
    if name == "all_ctgs":          # '7' and 'chr7' both listed: every contig sees every copy of a 'chr7' record,
        assert batch.n_svs > len(case["joined"])       # so the batch is not the flat callset any more
        return
    assert batch.n_svs == len(case["joined"])
    for i, g in enumerate(case["joined"]):
        rows_i = res.join_row[int(batch.csr_off[i]):int(batch.csr_off[i + 1])].tolist()
        got = [[int(batch.read_hp[r]), int(batch.read_ps[r]), int(batch.read_pc[r])] if r >= 0 else [] for r in rows_i]
        assert got == g["reads"]
    key = {}
    for i in range(batch.n_svs):
        key.setdefault((batch.sv_chrom[i], int(batch.sv_pos[i])), []).append(i)
    used = set()
    for t in case["trace"]:
        cand = [i for i in key[(t["chrom"], t["pos"])] if i not in used and res.cls[i] == t["ps_num"]]
        i = cand[0]
        used.add(i)
        f = t["f"]
        assert (res.hap1[i], res.hap2[i], res.hap0[i], res.allhap[i], res.ps[i]) == \
               (f["hap1"], f["hap2"], f["hap0"], f["allhap"], f["ps"])
        assert (res.totsc1[i], res.totsc2[i]) == (f["hap1_totsc"], f["hap2_totsc"])
        for nm in _lib.FEATURE_NAMES:
            assert float(res.feature(nm)[i]) == float(f[nm])


def test_dropin_reads_real_bam(golden_workdir):
    """Same golden case with the per-contig haplotagged files as real BGZF/BAM: decoded natively
    (no samtools), output byte-identical to the reference's."""
    import os
    from duet_b200 import sv_phasing
    from util_bam import record, write_bam
    case, home = golden_workdir("cutesv_3ctg")
    for fn in os.listdir(home + "/snp_phasing"):
        path = os.path.join(home, "snp_phasing", fn)
        recs = []
        with open(path) as f:
            for line in f:
                s = line.rstrip("\n").split("\t")
                aux = []
                for a in s[11:]:
                    tag, typ, val = a.split(":", 2)
                    aux.append((tag, typ, int(val)) if typ == "i" else (tag, typ, val))
                recs.append(record(s[0], int(s[3]), s[9], s[10], aux)[0])
        write_bam(path, recs, block=30000)
        with open(path, "rb") as f:
            assert f.read(2) == b"\x1f\x8b"
    sv_phasing.sv_phasing(home, case["svlen_thres"], case["suppread_thres"], 4, False)
    with open(home + "/phased_sv.vcf") as f:
        assert f.read() == case["phased_sv_vcf"]


# ---- cohort batches and randomized differential testing against the columnar oracle adapter ---------
def _compare_with_columnar_oracle(engine, batch, svlen, supp, repeats=1):
    """`repeats` > 1 re-runs the whole call (fresh upload each time): the kernels of one call overlap on
    the device, so a synchronisation slip shows up as a result that differs between runs."""
    from oracle.columnar_adapter import phase_batch_oracle
    engine.set_thresholds(svlen, supp)
    res = engine.run(batch)
    want = phase_batch_oracle(batch, svlen, supp)
    assert np.array_equal(res.join_row, want.join_row)
    for _ in range(repeats - 1):
        again = engine.run(batch)
        assert np.array_equal(again.join_row, want.join_row)
        assert np.array_equal(again.gt, res.gt) and np.array_equal(again.order, res.order)
    assert np.array_equal(res.cls, want.cls)
    assert np.array_equal(res.gt, want.gt)
    t = want.traced
    for k in ("ps", "hap1", "hap2", "hap0", "allhap", "totsc1", "totsc2"):
        assert np.array_equal(getattr(res, k)[t], getattr(want, k)[t]), k
    assert np.array_equal(res.features[:, t], want.features[:, t])
    assert np.array_equal(res.order, want.order)
    assert np.array_equal(res.shard_counts, want.shard_counts)
    return res


def test_cohort_batch_many_samples_one_call(engine):
    """BASELINE.json configs[4] in miniature: several samples x contigs = many shards in ONE device call;
    every sample's rows equal the oracle's for that sample alone."""
    samples = [synth.make_sample(100 + i, contigs=["1", "2", "3", "X"], n_reads=5000, n_svs=400, bp_per_read=700,
                                 block_mean=1.2e5, id_base=i << 44) for i in range(6)]
    batch = from_synth(samples)
    assert batch.n_shards == 24
    res = _compare_with_columnar_oracle(engine, batch, 50, 2)
    for i, s in enumerate(samples):
        rows, _ = synth_adapter.phase_sample(s)
        assert res.rows(batch, sample=i) == rows


@pytest.mark.parametrize("seed", range(6))
def test_random_batches_against_columnar_oracle(engine, seed):
    """Arbitrary (not genome-like) batches: negative PS / PC, HP 1/2, PC around the 8100 edge, empty
    shards, reads nobody supports, lists with repeats and misses, unsorted positions, equal positions."""
    rng = np.random.default_rng(1000 + seed)
    bb = BatchBuilder()
    n_shards = int(rng.integers(1, 12))
    for s in range(n_shards):
        n_reads = int(rng.choice([0, 1, 5, 40, 300]))
        names = [f"s{s}r{k}" for k in range(max(n_reads, 1))]
        pss = [int(x) for x in rng.choice([-5, 0, 7, 100, 2**31 - 1, -2**31 + 1, 4242], size=int(rng.integers(1, 6)))]
        reads = []
        for k in range(n_reads):
            nm = names[int(rng.integers(0, len(names)))] if rng.random() < 0.15 else names[k]     # duplicate QNAMEs
            reads.append((nm, int(rng.integers(1, 3)), int(rng.choice(pss)),
                          int(rng.choice([-3, 0, 1, 8099, 8100, 8101, 20000, int(rng.integers(0, 9000))]))))
        svs = []
        for v in range(int(rng.choice([0, 1, 3, 30]))):
            k = int(rng.integers(1, 60))
            lst = [names[int(rng.integers(0, len(names)))] if rng.random() < 0.8 else f"ghost{v}_{i}" for i in range(k)]
            svs.append(dict(pos=int(rng.choice([10, 10, 500, int(rng.integers(-50, 10**6))])),
                            svread=int(rng.integers(1, 40)), refread=int(rng.integers(0, 40)), names=lst,
                            svlen=int(rng.choice([10, 50, 3000])), gt_missing=bool(rng.random() < 0.1)))
        bb.shard(reads, svs, contig=str(s))
    batch = bb.build()
    if batch.n_svs == 0:
        bb.shard([("a", 1, 5, 5)], [dict(pos=1, svread=3, refread=0, names=["a"])], contig="z")
        batch = bb.build()
    _compare_with_columnar_oracle(engine, batch, int(rng.choice([0, 50])), int(rng.choice([0, 2, 10])))


def test_c2_full_size_parity(engine):
    """BASELINE.json configs[1] at FULL size (4.5 M reads / 3.7 M haplotagged, 25 k SVs, 402 k joins, 24
    contigs): every join row, class, genotype, PS, count, score sum, fp64 feature, the emission order and
    the per-contig counters equal the oracle's."""
    batch = from_synth(synth.config_c2(0), with_text=False)
    assert batch.n_reads > 3_500_000 and batch.n_svs == 25_001
    res = _compare_with_columnar_oracle(engine, batch, 50, 2, repeats=4)
    assert res.shard_counts[:, 2].sum() == res.order.shape[0] > 15_000
    # size-independent properties: emitted SVs are exactly those with a genotype; order is a permutation
    # of them, sorted by position inside every contig
    emitted = np.nonzero(res.gt)[0]
    assert np.array_equal(np.sort(res.order), emitted)
    shard = np.searchsorted(batch.sv_off, res.order, side="right") - 1
    assert (np.diff(shard) >= 0).all()
    same = np.diff(shard) == 0
    assert (np.diff(batch.sv_pos[res.order])[same] >= 0).all()


def test_c4_dense_full_size_parity(engine):
    """BASELINE.json configs[3] shape (60x, dense support lists with a tail to 2 000 reads) at full size."""
    batch = from_synth(synth.config_c4(0), with_text=False)
    assert batch.n_joins > 1_500_000 and np.diff(batch.csr_off).max() >= 1500
    _compare_with_columnar_oracle(engine, batch, 50, 2, repeats=6)


def test_c5_share_full_size_parity(engine):
    """BASELINE.json configs[4], one GPU's share at FULL size: 4 samples x WGS 30x = 96 (sample, contig)
    shards, ~14.9 M haplotagged reads and 100 k SVs in ONE device call (the batch bench.py's c5 times), with
    the per-sample id_base << 44 name spaces.  Every join row, class, genotype, PS, count, fp64 feature, the
    emission order and the per-shard counters equal the columnar oracle's; the call is repeated (the
    kernels of a call overlap on the device: a synchronisation slip shows as a run-to-run difference)."""
    samples = [synth.config_c2(4 * 0 + k, id_base=(4 * 0 + k) << 44) for k in range(4)]
    batch = from_synth(samples, with_text=False)
    assert batch.n_shards == 96 and batch.n_svs == 4 * 25_001 and batch.n_reads > 14_000_000
    res = _compare_with_columnar_oracle(engine, batch, 50, 2, repeats=3)
    assert res.shard_counts[:, 2].sum() == res.order.shape[0] > 60_000
    # size-independent: per shard the order is position sorted and covers exactly the SVs with a genotype
    emitted = np.nonzero(res.gt)[0]
    assert np.array_equal(np.sort(res.order), emitted)
    shard = np.searchsorted(batch.sv_off, res.order, side="right") - 1
    assert (np.diff(shard) >= 0).all()
    assert (np.diff(batch.sv_pos[res.order])[np.diff(shard) == 0] >= 0).all()


# ---- property-based differential testing (hypothesis) against the columnar oracle ---------------------------
try:
    from hypothesis import HealthCheck, given, settings, strategies as st
    _HAVE_HYPOTHESIS = True
except Exception:                                            # pragma: no cover
    _HAVE_HYPOTHESIS = False

if _HAVE_HYPOTHESIS:
    _ps = st.sampled_from([-5, 0, 7, 100, 2**31 - 1, -2**31 + 1, 4242, 99_999])
    _pc = st.one_of(st.sampled_from([-3, 0, 1, 8099, 8100, 8101, 20000]), st.integers(0, 9000))

    @st.composite
    def _shards(draw):
        n_shards = draw(st.integers(1, 6))
        out = []
        for s in range(n_shards):
            n_reads = draw(st.sampled_from([0, 1, 3, 17, 90]))
            pool = [f"s{s}r{k}" for k in range(max(n_reads, 1))]
            reads = [(draw(st.sampled_from(pool)) if draw(st.integers(0, 9)) == 0 else pool[k],
                      draw(st.integers(1, 2)), draw(_ps), draw(_pc)) for k in range(n_reads)]
            svs = []
            for v in range(draw(st.sampled_from([0, 1, 2, 9]))):
                names = draw(st.lists(st.one_of(st.sampled_from(pool), st.just(f"ghost{s}_{v}")), min_size=1, max_size=40))
                svs.append(dict(pos=draw(st.one_of(st.sampled_from([10, 500]), st.integers(-50, 10**6))),
                                svread=draw(st.integers(1, 40)), refread=draw(st.integers(0, 40)), names=names,
                                svlen=draw(st.sampled_from([10, 50, 3000])), gt_missing=draw(st.integers(0, 9)) == 0))
            out.append((reads, svs))
        return out

    @settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow,
                                                                     HealthCheck.data_too_large])
    @given(shards=_shards(), svlen=st.sampled_from([0, 50]), supp=st.sampled_from([0, 2, 10]))
    def test_hypothesis_batches_against_columnar_oracle(engine, shards, svlen, supp):
        """Randomised (shrinking) differential test: arbitrary shards -- duplicate QNAMEs, names listed twice or
        absent, PS / PC at the int32 and 8100 edges, empty shards, unsorted and equal positions -- must give the
        oracle's join rows, classes, genotypes, PS, counts, features, order and counters."""
        bb = BatchBuilder()
        for s, (reads, svs) in enumerate(shards):
            bb.shard(reads, svs, contig=str(s))
        batch = bb.build()
        if batch.n_svs == 0:
            bb.shard([("a", 1, 5, 5)], [dict(pos=1, svread=3, refread=0, names=["a"])], contig="z")
            batch = bb.build()
        _compare_with_columnar_oracle(engine, batch, svlen, supp)


def test_dropin_stage_decodes_into_page_locked_columns(golden_workdir, monkeypatch):
    """The stage's own path: both native decoders write into page-locked columns, the device call reads the tag
    records in place, only genotype / phase set / order / counters come back -- same rows as the general readers
    (and as the reference)."""
    from duet_b200 import sv_phasing_fn
    from duet_b200.engine import is_pinned
    case, home = golden_workdir("svim_shuffled")
    args = (home + "/sv_calling/variants.vcf", home + "/snp_phasing/", case["svlen_thres"], case["suppread_thres"], 4, False)
    rows = sv_phasing_fn.generate_phased_callset(*args)
    assert rows == case["rows"]
    assert sv_phasing_fn.last_timings["native_decode"] is True
    b = sv_phasing_fn.last_batch
    for name in _lib.INPUT_COLUMNS:
        arr = getattr(b, name)
        if name in ("read_off", "sv_off") or arr is None:      # host-side descriptors
            continue
        assert is_pinned(arr), name
    assert not is_pinned(np.zeros(16))
    monkeypatch.setenv("DUET_GENERAL_DECODE", "1")
    assert sv_phasing_fn.generate_phased_callset(*args) == rows
    assert sv_phasing_fn.last_timings["native_decode"] is False
