"""Kernel set B.  CPU: the scipy oracle against an independent O(n^2) restatement of the same
spec.  GPU: the CUDA path (through the C-ABI) against the oracle, bit-exact cluster membership.
SVIM parity is unpinned (see oracle/cluster_oracle.py)."""
import numpy as np
import pytest

from duet_b200 import synth
from oracle import cluster_oracle


def small_cases():
    yield "events", synth.make_signatures(1, n=900, contigs=["1", "2", "X"])
    yield "one_contig", synth.make_signatures(2, n=400, contigs=["21"])
    rng = np.random.default_rng(3)
    n = 600                                     # dense: everything inside a few windows, chains of links
    st = rng.integers(1000, 4000, size=n)
    yield "dense", (rng.integers(0, 2, size=n), rng.integers(0, 3, size=n), st, st + rng.integers(0, 700, size=n))
    # exact-threshold pairs: |dc|/900 + |ds|/max = 0.9 exactly representable cases and zero spans
    yield "edges", (np.zeros(8, int), np.zeros(8, int),
                    np.array([100, 100, 505, 505, 5000, 5000, 9000, 9900]),
                    np.array([200, 200, 605, 605, 5000, 5000, 9100, 10000]))


@pytest.mark.parametrize("name,cols", list(small_cases()), ids=lambda x: x if isinstance(x, str) else "")
def test_oracle_matches_bruteforce(name, cols):
    for md in (0.9, 0.3):
        a, na = cluster_oracle.cluster(*cols, max_distance=md)
        b, nb = cluster_oracle.cluster_bruteforce(*[np.asarray(c).tolist() for c in cols], max_distance=md)
        assert na == nb and np.array_equal(a, b)
    assert cluster_oracle.cluster([], [], [], [])[1] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,cols", list(small_cases()), ids=lambda x: x if isinstance(x, str) else "")
def test_gpu_small(name, cols):
    from duet_b200.sv_clustering import cluster_signatures
    for md, win in ((0.9, 1000), (0.3, 1000), (0.9, 50)):
        want, n_want = cluster_oracle.cluster(*cols, max_distance=md, window=win)
        got, n_got, _ = cluster_signatures(*cols, cluster_max_distance=md, partition_window=win)
        assert n_got == n_want
        assert np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_200k_and_properties():
    from duet_b200.sv_clustering import cluster_signatures
    cols = synth.make_signatures(5, n=200_000)
    want, n_want = cluster_oracle.cluster(*cols)
    got, n_got, ms = cluster_signatures(*cols)
    assert n_got == n_want and np.array_equal(got, want)
    # size-independent properties: ids are fixed points, members share (contig, type), permutation
    # of the input permutes the partition
    assert np.array_equal(got[got], got)
    assert np.array_equal(cols[0][got], cols[0]) and np.array_equal(cols[1][got], cols[1])
    p = np.random.default_rng(0).permutation(len(got))
    got_p, n_p, _ = cluster_signatures(*[c[p] for c in cols])
    assert n_p == n_got
    inv = np.empty_like(p); inv[p] = np.arange(len(p))
    # same partition: two signatures share a cluster before iff they share one after
    a = got; b = got_p[inv]
    _, ia = np.unique(a, return_inverse=True); _, ib = np.unique(b, return_inverse=True)
    assert len(set(zip(ia.tolist(), ib.tolist()))) == n_got


@pytest.mark.gpu
def test_gpu_c3_full_size_properties():
    """BASELINE.json configs[2]: 2 M signatures, cluster_max_distance 0.9."""
    from duet_b200.sv_clustering import cluster_signatures
    cols = synth.make_signatures(0, n=2_000_000)
    got, n_got, ms = cluster_signatures(*cols)
    want, n_want = cluster_oracle.cluster(*cols)
    assert n_got == n_want and np.array_equal(got, want)
    assert np.array_equal(got[got], got)


@pytest.mark.gpu
def test_gpu_rejects_bad_input():
    from duet_b200.engine import DuetError
    from duet_b200.sv_clustering import cluster_signatures
    with pytest.raises(DuetError):
        cluster_signatures([0], [0], [10], [5])          # end < start
    assert cluster_signatures([], [], [], [])[1] == 0


def _oracle_cluster_fn(contig, typ, start, end, cluster_max_distance=0.9, *, position_normalizer=900.0, partition_window=1000):
    from oracle import cluster_oracle
    ids, n = cluster_oracle.cluster(contig, typ, start, end, cluster_max_distance, position_normalizer, partition_window)
    return ids, n, 0.0


def _shard_worker(rank, world, port, out_q, on_gpu=False):
    import os
    import sys
    import torch.distributed as dist
    from conftest import ROOT
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if on_gpu:
        import torch
        os.environ["DUET_DEVICE"] = str(rank % torch.cuda.device_count())      # >= 2 GPUs visible: every rank its own
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from duet_b200 import synth
        from duet_b200.sv_clustering import cluster_signatures_sharded
        cols = synth.make_signatures(3, 60_000, contigs=["1", "2", "21", "X"])
        index, ids, total = cluster_signatures_sharded(*cols, cluster_fn=None if on_gpu else _oracle_cluster_fn)
        out_q.put((rank, index, ids, total))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_contig_sharded_clustering_on_gpu():
    """The same with the device path in every rank (each on its own GPU when the box has several)."""
    test_contig_sharded_clustering_equals_single_process(on_gpu=True, world=2)


def test_contig_sharded_clustering_equals_single_process(on_gpu=False, world=3):
    """(contig, type) groups LPT-packed over 3 gloo ranks (the oracle stands in for the device): the union of
    the ranks' slices is the single-process result, every signature is owned exactly once, the all-reduced
    cluster count is the global one."""
    import socket
    import torch.multiprocessing as mp
    from duet_b200 import synth
    from oracle import cluster_oracle
    cols = synth.make_signatures(3, 60_000, contigs=["1", "2", "21", "X"])
    want, n_want = cluster_oracle.cluster(*cols)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q, on_gpu)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = np.full(want.shape[0], -1, np.int64)
    for rank, index, ids, total in got:
        assert total == n_want
        assert (merged[index] == -1).all()
        merged[index] = ids
    assert np.array_equal(merged, want)


@pytest.mark.gpu
def test_gpu_general_path_and_oversize_fallback(monkeypatch):
    """The call normally takes the bucketed path (cluster_fast.cuh).  DUET_CL_GENERAL forces the general path
    (global radix sort + tile kernels); a hot spot that does not fit a bucket's shared memory makes the bucketed
    path stand down and the host run the general one.  All three give the oracle's ids."""
    from duet_b200.sv_clustering import cluster_signatures
    cols = synth.make_signatures(7, n=150_000)
    want, n_want = cluster_oracle.cluster(*cols)
    got, n_got, _ = cluster_signatures(*cols)
    assert n_got == n_want and np.array_equal(got, want)
    monkeypatch.setenv("DUET_CL_GENERAL", "1")
    got, n_got, _ = cluster_signatures(*cols)
    assert n_got == n_want and np.array_equal(got, want)
    monkeypatch.delenv("DUET_CL_GENERAL")
    # 6000 signatures of one type inside 3 kb on top of a genome-wide background: one bucket overflows
    rng = np.random.default_rng(11)
    hot_start = rng.integers(5_000_000, 5_003_000, size=6000)
    hot = (np.zeros(6000, np.int32), np.zeros(6000, np.int32), hot_start.astype(np.int32),
           (hot_start + rng.integers(40, 4000, size=6000)).astype(np.int32))
    mixed = [np.concatenate([a, b]) for a, b in zip(cols, hot)]
    want, n_want = cluster_oracle.cluster(*mixed)
    got, n_got, _ = cluster_signatures(*mixed)
    assert n_got == n_want and np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_bucket_boundaries():
    """Clusters that straddle bucket boundaries (the zone / pending / global-forest machinery): signatures laid
    densely along one contig so that every boundary of the 2^B buckets is crossed by a chain."""
    from duet_b200.sv_clustering import cluster_signatures
    rng = np.random.default_rng(5)
    n = 300_000
    start = np.sort(rng.integers(0, 40_000_000, size=n)).astype(np.int32)       # ~130 bp apart: long chains everywhere
    span = rng.integers(200, 260, size=n).astype(np.int32)
    p = rng.permutation(n)
    cols = (np.zeros(n, np.int32), rng.integers(0, 2, size=n).astype(np.int32)[p], start[p], (start + span)[p])
    for win in (1000, 100):
        want, n_want = cluster_oracle.cluster(*cols, window=win)
        got, n_got, _ = cluster_signatures(*cols, partition_window=win)
        assert n_got == n_want and np.array_equal(got, want)


@pytest.mark.gpu
def test_gpu_random_cases_against_oracle():
    """Property-based differential test: random sizes, coordinate ranges (from everything-in-one-window to a whole
    chromosome: that decides how many key bits and buckets a call gets), windows, thresholds and normalizers, zero
    spans and exact duplicates -- device ids == oracle ids."""
    from hypothesis import HealthCheck, given, settings, strategies as st
    from duet_b200.sv_clustering import cluster_signatures

    @settings(max_examples=120, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(seed=st.integers(0, 2**31 - 1), n=st.integers(1, 3000), n_contig=st.integers(1, 4), n_type=st.integers(1, 3),
           extent=st.sampled_from([50, 3_000, 200_000, 40_000_000, 240_000_000]),
           max_span=st.sampled_from([0, 1, 300, 20_000]),
           window=st.sampled_from([0, 3, 50, 1000, 60_000]),
           md=st.sampled_from([0.0, 0.3, 0.9, 2.5]), norm=st.sampled_from([1.0, 900.0]),
           dup=st.booleans())
    def check(seed, n, n_contig, n_type, extent, max_span, window, md, norm, dup):
        rng = np.random.default_rng(seed)
        contig = rng.integers(0, n_contig, size=n).astype(np.int32) * 7          # sparse ids
        typ = rng.integers(0, n_type, size=n).astype(np.int32)
        start = rng.integers(0, extent + 1, size=n).astype(np.int32)
        span = rng.integers(0, max_span + 1, size=n).astype(np.int32)
        if dup and n > 4:                                                       # exact duplicates and near-ties
            k = n // 3
            src = rng.integers(0, n, size=k)
            dst = rng.integers(0, n, size=k)
            contig[dst], typ[dst], start[dst], span[dst] = contig[src], typ[src], start[src], span[src]
        cols = (contig, typ, start, start + span)
        want, n_want = cluster_oracle.cluster(*cols, max_distance=md, normalizer=norm, window=window)
        got, n_got, _ = cluster_signatures(*cols, cluster_max_distance=md, position_normalizer=norm, partition_window=window)
        assert n_got == n_want
        assert np.array_equal(got, want)

    check()
