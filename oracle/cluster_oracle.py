"""CPU oracle for kernel set B (signature clustering) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference contains no clustering arithmetic; it passes
`--cluster_max_distance` to the external `svim` 1.4.2 CLI
(/root/reference/src/duet/sv_calling.py:14-15), which is neither vendored nor installed here.
This file restates the spec frozen in SURVEY.md §8(c) / csrc/cluster_kernels.cuh -- NOT svim's
own (average-linkage) procedure -- with numpy + scipy.sparse.csgraph:

  c2 = start + end, span = end - start
  edge(i, j)  <=>  same (contig, type), |c2_i - c2_j| <= 2*window and
                   (|c2_i - c2_j| * 0.5) / normalizer + |span_i - span_j| / max(span_i, span_j) <= max_distance
  clusters = connected components; id = smallest original index in the component.
"""
from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


def edges(contig, typ, start, end, max_distance=0.9, normalizer=900.0, window=1000):
    contig, typ = np.asarray(contig, np.int64), np.asarray(typ, np.int64)
    start, end = np.asarray(start, np.int64), np.asarray(end, np.int64)
    n = contig.shape[0]
    c2, span = start + end, end - start
    order = np.lexsort((c2, typ, contig))           # stable; ties keep original order
    sc, st, s2, ss = contig[order], typ[order], c2[order], span[order]
    src, dst = [], []
    d = 1
    while d < n:
        same = (sc[d:] == sc[:-d]) & (st[d:] == st[:-d]) & (s2[d:] - s2[:-d] <= 2 * window)
        if not same.any():
            break                                   # sorted: a wider offset cannot qualify either
        i = np.nonzero(same)[0]
        j = i + d
        dpos = ((s2[j] - s2[i]).astype(np.float64) * 0.5) / np.float64(normalizer)
        mx = np.maximum(ss[i], ss[j])
        with np.errstate(divide="ignore", invalid="ignore"):
            dspan = np.where(mx > 0, np.abs(ss[i] - ss[j]).astype(np.float64) / mx.astype(np.float64), 0.0)
        ok = dpos + dspan <= np.float64(max_distance)
        src.append(order[i[ok]])
        dst.append(order[j[ok]])
        d += 1
    if src:
        return np.concatenate(src), np.concatenate(dst)
    return np.zeros(0, np.int64), np.zeros(0, np.int64)


def cluster(contig, typ, start, end, max_distance=0.9, normalizer=900.0, window=1000):
    """-> (cluster_id[n] = smallest member index, n_clusters)"""
    n = len(contig)
    if n == 0:
        return np.zeros(0, np.int32), 0
    src, dst = edges(contig, typ, start, end, max_distance, normalizer, window)
    g = coo_matrix((np.ones(src.shape[0], np.int8), (src, dst)), shape=(n, n))
    n_comp, lab = connected_components(g, directed=False)
    first = np.full(n_comp, n, np.int64)
    np.minimum.at(first, lab, np.arange(n))
    return first[lab].astype(np.int32), int(n_comp)


def cluster_bruteforce(contig, typ, start, end, max_distance=0.9, normalizer=900.0, window=1000):
    """O(n^2) pure-Python cross-check of `cluster` for small n (no sorting, no windows trick)."""
    n = len(contig)
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for i in range(n):
        for j in range(i + 1, n):
            if contig[i] != contig[j] or typ[i] != typ[j]:
                continue
            d2 = abs((start[i] + end[i]) - (start[j] + end[j]))
            if d2 > 2 * window:
                continue
            si, sj = end[i] - start[i], end[j] - start[j]
            mx = max(si, sj)
            d = (float(d2) * 0.5) / float(normalizer) + (abs(si - sj) / mx if mx > 0 else 0.0)
            if d <= max_distance:
                a, b = find(i), find(j)
                if a != b:
                    parent[max(a, b)] = min(a, b)
    ids = np.array([find(i) for i in range(n)], np.int32)
    return ids, len(set(ids.tolist()))
