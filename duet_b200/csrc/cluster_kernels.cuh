// Kernel set B: span-position-distance clustering of SV signatures (BASELINE.json configs[2]).
//
// The reference does not contain this arithmetic: Duet passes `--cluster_max_distance` to the
// external `svim` CLI (/root/reference/src/duet/sv_calling.py:14-15, default 0.9 utils.py:27-28).
// svim 1.4.2 is not vendored and not installed here, so PARITY WITH SVIM IS UNPINNED -- and this is
// a DIFFERENT algorithm from svim's (svim cuts partitions and runs average-linkage hierarchical
// clustering on them; single linkage chains, so for the same threshold it gives fewer, larger clusters).
// What these kernels implement -- and what oracle/cluster_oracle.py restates on the CPU -- is the spec
// frozen in SURVEY.md §8(c), following the north_star's formulation (sort, windowed pairwise distances
// in shared-memory tiles, union-find):
//
//   signature i = (contig, type, start, end);  c2_i = start+end (twice the centre), span_i = end-start
//   edge(i,j)  <=>  same contig and type,  |c2_i - c2_j| <= 2*window,  and
//                   (|c2_i - c2_j| * 0.5) / normalizer + |span_i - span_j| / max(span_i, span_j)  <=  max_distance
//                   (fp64, this operand order; the span term is 0 when both spans are 0)
//   clusters = connected components;  cluster id = smallest ORIGINAL index in the component.
//
// Device path, no host round trip anywhere:
//   k_cl_keys     64-bit keys contig | type | c2, payload (original index, span), the three maxima
//   k_rs_*        stable LSD radix sort, 8-bit digits taken from the PACKED key (only as many bits as the
//                 maxima need: 36 for a human genome -> five passes; passes beyond that return at once).  A
//                 tile is ranked in one go: every warp ranks its own contiguous 256 keys with match_any, one
//                 block-wide prefix over the warps' digit counts, then the scatter
//   k_cl_runs     maximal runs of sorted neighbours that pass the edge rule: the forest starts with every run
//                 hanging under its first position (block-wide scan, no atomics)
//   k_cl_edges    one block per tile of 256 sorted signatures (+ 256 halo): the windowed pairwise distances
//                 out of shared memory; only pairs of DIFFERENT runs are tested, passing pairs united in a
//                 lock-free union-find (roots = minima)
//   k_cl_label / k_cl_write   roots' smallest original index (one atomic per warp and root), ids in input order
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace duet {

constexpr int kClThreads = 256;
constexpr int kRsItems = 8;                        // radix sort: keys per thread per tile
constexpr int kRsTile = kClThreads * kRsItems;     // 2048
constexpr int kRsBits = 8;
constexpr int kRsBins = 1 << kRsBits;
constexpr int kRsMaxPasses = 7;                    // 16 + 8 + 32 key bits
constexpr int kClHalo = 256;

struct ClMeta {                    // device-side facts about the call
    unsigned max_contig, max_type, max_c2;
    int bad;                       // invalid input seen
    int n_clusters;
    // the bucketed path (cluster_fast.cuh)
    int oversize;                  // a bucket does not fit shared memory: the call takes the general path
    int n_buckets;                 // non-empty buckets
    int ticket;                    // next entry of the bucket list to hand out
    int n_pending;                 // members of components that touch a bucket boundary
    int pad[3];
};

struct ClusterArgs {
    int n;
    const int *contig, *type, *start, *end;
    unsigned long long *key[2];    // ping-pong: contig:16 | type:8 | 0:8 | c2:32
    unsigned long long *pay[2];    // span:32 | original index:32
    int *parent;                   // union-find over sorted positions; roots are component minima
    int *minidx;                   // smallest original index per root
    int *out;                      // [n] cluster id per ORIGINAL index
    ClMeta *meta;
    unsigned *block_hist;          // [kRsBins][n_tiles] bin-major, then [kRsBins] bin totals
    int n_tiles;
    double max_distance, normalizer;
    unsigned window2;              // 2 * partition window
    long long *dbg;                // optional per-block clock stamps of k_cl_edges (DUET_CL_DBG), NULL in production
};

constexpr int kClDbgMarks = 8;
__device__ __forceinline__ void cl_mark(const ClusterArgs &a, int k) {
    if (a.dbg && threadIdx.x == 0) a.dbg[(size_t)blockIdx.x * kClDbgMarks + k] = clock64();
}

__device__ __forceinline__ int bits_for(unsigned v) { return 32 - __clz(v); }          // 0 for v == 0

// how the sort sees a key: only the bits the call's maxima need, contig above type above c2
struct Packing {
    int b2, bt, total;
    __device__ __forceinline__ unsigned long long pack(unsigned long long k) const {
        const unsigned long long c2 = k & 0xFFFFFFFFull, t = (k >> 32) & 0xFFull, c = k >> 40;
        return (c << (bt + b2)) | (t << b2) | c2;
    }
};
__device__ __forceinline__ Packing packing_of(const ClMeta *m) {
    Packing p;
    p.b2 = bits_for(m->max_c2); p.bt = bits_for(m->max_type);
    p.total = p.b2 + p.bt + bits_for(m->max_contig);
    return p;
}
__device__ __forceinline__ int sorted_buffer(const Packing &p) { return ((p.total + kRsBits - 1) / kRsBits) & 1; }

// ---- keys ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kClThreads)
k_cl_keys(ClusterArgs a) {
    const int i = blockIdx.x * kClThreads + threadIdx.x;
    unsigned mc = 0, mt = 0, m2 = 0;
    if (i < a.n) {
        const long long st = a.start[i], en = a.end[i];
        const unsigned c = (unsigned)a.contig[i], t = (unsigned)a.type[i];
        if (st < 0 || en < st || st + en > 0xFFFFFFFFll || c > 0xFFFFu || t > 0xFFu) a.meta->bad = 1;      // invalid input
        mc = c & 0xFFFFu; mt = t & 0xFFu; m2 = (unsigned)(st + en);
        a.key[0][i] = ((unsigned long long)mc << 40) | ((unsigned long long)mt << 32) | (unsigned long long)m2;
        a.pay[0][i] = ((unsigned long long)(unsigned)(en - st) << 32) | (unsigned)i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mc = max(mc, __shfl_xor_sync(0xffffffffu, mc, o));
        mt = max(mt, __shfl_xor_sync(0xffffffffu, mt, o));
        m2 = max(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (mc > a.meta->max_contig) atomicMax(&a.meta->max_contig, mc);
        if (mt > a.meta->max_type) atomicMax(&a.meta->max_type, mt);
        if (m2 > a.meta->max_c2) atomicMax(&a.meta->max_c2, m2);
    }
}

// ---- LSD radix sort, 8-bit digits of the packed key: histogram / scan / ranked scatter -----------------
__global__ void __launch_bounds__(kClThreads)
k_rs_hist(ClusterArgs a, int pass) {
    const Packing pk = packing_of(a.meta);
    const int shift = pass * kRsBits;
    if (shift >= pk.total) return;                       // this call's keys have no such digit
    __shared__ unsigned s_h[kRsBins];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long *key = a.key[pass & 1];
    const int base = blockIdx.x * kRsTile;
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const int i = base + r * kClThreads + threadIdx.x;
        if (i < a.n) atomicAdd(&s_h[(unsigned)(pk.pack(key[i]) >> shift) & (kRsBins - 1)], 1u);
    }
    __syncthreads();
    a.block_hist[threadIdx.x * a.n_tiles + blockIdx.x] = s_h[threadIdx.x];      // bin-major
}

// per-bin scan: block b turns row b of the bin-major histogram (n_tiles counters) into exclusive
// offsets inside the bin and records the bin total; k_rs_scatter adds the bins' bases itself
__global__ void __launch_bounds__(kClThreads)
k_rs_scan(ClusterArgs a, int pass) {
    const Packing pk = packing_of(a.meta);
    if (pass * kRsBits >= pk.total) return;
    __shared__ unsigned s_w[kClThreads / 32];
    __shared__ unsigned s_carry;
    const int n_blocks = a.n_tiles;
    unsigned *row = a.block_hist + (size_t)blockIdx.x * n_blocks;
    unsigned *bin_total = a.block_hist + (size_t)kRsBins * n_blocks;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < n_blocks; base += kClThreads) {
        const int i = base + threadIdx.x;
        const unsigned x = i < n_blocks ? row[i] : 0u;
        unsigned inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_w[w] = inc;
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int k = 0; k < kClThreads / 32; ++k) { const unsigned t = s_w[k]; if (k < w) before += t; total += t; }
        const unsigned carry = s_carry;
        if (i < n_blocks) row[i] = carry + before + inc - x;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) bin_total[blockIdx.x] = s_carry;
}

// One tile of 2048 keys: warp w owns the contiguous keys [256 w, 256 (w+1)) of the tile, in eight rounds of
// 32.  A round ranks its keys inside the warp with match_any (lanes holding the same digit, in lane = key
// order) on top of the warp's running count of that digit; ONE block-wide step then turns the warps' counts
// into offsets (bin base + tile offset + the warps before), and the keys go out.  Stable: a key's place among
// equal digits follows (tile, warp, round, lane) = input order.
__global__ void __launch_bounds__(kClThreads)
k_rs_scatter(ClusterArgs a, int pass) {
    const Packing pk = packing_of(a.meta);
    const int shift = pass * kRsBits;
    if (shift >= pk.total) return;
    __shared__ unsigned s_cnt[kClThreads / 32][kRsBins];
    __shared__ unsigned s_base[kRsBins];
    __shared__ unsigned s_w[kClThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned long long *kin = a.key[pass & 1], *pin = a.pay[pass & 1];
    unsigned long long *kout = a.key[(pass + 1) & 1], *pout = a.pay[(pass + 1) & 1];
    const unsigned *bin_total = a.block_hist + (size_t)kRsBins * a.n_tiles;
    {   // base of bin t = totals of the bins before it (block-wide exclusive scan of 256 totals) + this tile's offset
        const unsigned x = bin_total[threadIdx.x];
        unsigned inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_w[w] = inc;
#pragma unroll
        for (int k = 0; k < kClThreads / 32; ++k) s_cnt[k][threadIdx.x] = 0;
        __syncthreads();
        unsigned before = 0;
        for (int k = 0; k < w; ++k) before += s_w[k];
        s_base[threadIdx.x] = before + inc - x + a.block_hist[threadIdx.x * a.n_tiles + blockIdx.x];
    }
    unsigned long long key[kRsItems], pay[kRsItems];
    unsigned rank[kRsItems];
    const int base = blockIdx.x * kRsTile + w * (32 * kRsItems);
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const int i = base + r * 32 + lane;
        key[r] = i < a.n ? kin[i] : 0ull;
        pay[r] = i < a.n ? pin[i] : 0ull;
    }
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        const bool live = base + r * 32 + lane < a.n;
        const unsigned d = (unsigned)(pk.pack(key[r]) >> shift) & (kRsBins - 1);
        const unsigned peers = __match_any_sync(0xffffffffu, live ? d : kRsBins + lane);      // dead lanes match nobody
        const unsigned before = live ? s_cnt[w][d] : 0u;
        __syncwarp();
        rank[r] = before + __popc(peers & ((1u << lane) - 1u));
        if (live && (peers & ((1u << lane) - 1u)) == 0) s_cnt[w][d] = before + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // digit t: the warps' counts -> where each warp's keys of that digit start
        unsigned run = s_base[threadIdx.x];
#pragma unroll
        for (int k = 0; k < kClThreads / 32; ++k) {
            const unsigned c = s_cnt[k][threadIdx.x];
            s_cnt[k][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRsItems; ++r) {
        if (base + r * 32 + lane >= a.n) continue;
        const unsigned d = (unsigned)(pk.pack(key[r]) >> shift) & (kRsBins - 1);
        const unsigned pos = s_cnt[w][d] + rank[r];
        kout[pos] = key[r];
        pout[pos] = pay[r];
    }
}

// ---- union-find ----------------------------------------------------------------------------------
// Path halving: every write re-points x to its CURRENT grandparent, an ancestor read just now.
// Parents only ever move to smaller indices (roots hang under smaller roots), so concurrent
// finds / unions can never create a cycle.  Works on global and on shared memory alike.
__device__ __forceinline__ int uf_find(volatile int *p, int x) {
    for (;;) {
        const int px = p[x];
        if (px == x) return x;
        const int ppx = p[px];
        if (ppx != px) p[x] = ppx;
        x = px;
    }
}

__device__ __forceinline__ void uf_unite(int *parent, int x, int y) {
    for (;;) {
        x = uf_find(parent, x);
        y = uf_find(parent, y);
        if (x == y) return;
        if (x > y) { const int t = x; x = y; y = t; }
        if (atomicCAS(parent + y, y, x) == y) return;      // larger root hangs under the smaller one
    }
}

// the edge rule for two signatures of one (contig, type) segment, d2 = |c2_i - c2_j| <= window2 already checked.
// A division-free single-precision version of the inequality, multiplied through by normalizer * max(span),
// decides all pairs that are not within 1e-5 (relative) of the threshold; only those pay for the exact fp64
// quotients in the spec's operand order.
__device__ __forceinline__ bool cl_edge(unsigned d2, int si, int sj, float nf, float mdf, double normalizer, double max_distance) {
    const int mx = max(si, sj);
    const float fm = mx > 0 ? (float)mx : 1.0f;
    const float lhs = (float)d2 * 0.5f * fm + (float)abs(si - sj) * nf;       // (dpos + dspan) * nf * mx
    const float rhs = mdf * nf * fm;
    if (lhs > rhs * 1.00001f) return false;
    if (lhs < rhs * 0.99999f) return true;
    const double dpos = ((double)d2 * 0.5) / normalizer;
    const double dspan = mx > 0 ? (double)abs(si - sj) / (double)mx : 0.0;
    return dpos + dspan <= max_distance;
}

// Runs: sorted neighbours i-1, i that pass the edge rule belong together, and in real signature sets that is
// nearly every edge there is (the members of one event sit next to each other).  So the forest STARTS with
// every maximal run of linked neighbours hanging under its first position -- a block-wide max-scan, no
// atomics -- and k_cl_edges only has to unite the runs that some longer-range pair connects.  A run that
// began in the previous tile hangs under that tile's last position (find follows it from there).
// These kernels work on tiles of kClTile = 1024 sorted positions, four per thread: what bounds them is the
// chain of dependent global round trips of a block (meta -> columns -> forest words), not bytes, so a block
// requests four positions' worth at once and the grid is resident in two waves.
constexpr int kClItems = 4;
constexpr int kClTile = kClThreads * kClItems;      // 1024

__device__ __forceinline__ bool cl_linked(unsigned long long kp, unsigned long long ki, int sp, int si, const ClusterArgs &a) {
    const unsigned d2 = (unsigned)ki - (unsigned)kp;            // sorted: non-negative inside a segment
    return (unsigned)(kp >> 32) == (unsigned)(ki >> 32) && d2 <= a.window2 &&
           cl_edge(d2, sp, si, (float)a.normalizer, (float)a.max_distance, a.normalizer, a.max_distance);
}

__global__ void __launch_bounds__(kClThreads)
k_cl_runs(ClusterArgs a) {
    __shared__ unsigned long long s_key[kClTile + 1];
    __shared__ int s_span[kClTile + 1];
    __shared__ int s_w[kClThreads / 32];
    const Packing pk = packing_of(a.meta);
    const int cur = sorted_buffer(pk);
    const unsigned long long *key = a.key[cur], *pay = a.pay[cur];
    const int i0 = blockIdx.x * kClTile;
    for (int t = threadIdx.x; t < kClTile + 1; t += kClThreads) {      // positions i0-1 .. i0+1023
        const int j = i0 - 1 + t;
        const bool ok = j >= 0 && j < a.n;
        s_key[t] = ok ? key[j] : ~0ull;
        s_span[t] = ok ? (int)(pay[j] >> 32) : 0;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int t0 = threadIdx.x * kClItems;                              // my four consecutive positions
    int head[kClItems];
    int run = -1;                                                       // where the current run starts inside the tile
#pragma unroll
    for (int u = 0; u < kClItems; ++u) {
        const int t = t0 + u, i = i0 + t;
        const bool link = i > 0 && i < a.n && cl_linked(s_key[t], s_key[t + 1], s_span[t], s_span[t + 1], a);
        if (!link) run = t;
        head[u] = run;                                                  // -1: continues from before my first position
    }
    int inc = run;                                                      // inclusive max-scan over the threads
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, x);
    }
    if (lane == 31) s_w[w] = inc;
    int before = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) before = -1;
    __syncthreads();
    for (int k = 0; k < w; ++k) before = max(before, s_w[k]);
    int par[kClItems];
#pragma unroll
    for (int u = 0; u < kClItems; ++u) {
        const int hd = head[u] >= 0 ? head[u] : before;
        par[u] = hd >= 0 ? i0 + hd : i0 - 1;
    }
    const int i = i0 + t0;
    if (i + kClItems <= a.n) {                                          // allocations are 16-byte aligned, i is a multiple of 4
        *reinterpret_cast<int4 *>(a.parent + i) = make_int4(par[0], par[1], par[2], par[3]);
        *reinterpret_cast<int4 *>(a.minidx + i) = make_int4(INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX);
    } else {
#pragma unroll
        for (int u = 0; u < kClItems; ++u)
            if (i + u < a.n) { a.parent[i + u] = par[u]; a.minidx[i + u] = INT32_MAX; }
    }
}

// Windowed pairwise distances, out of the tile (+ halo) in shared memory.  The members of a position's own
// run are connected already and are SKIPPED AS A RANGE: a bit per position marks the run starts, and the
// scan of position i begins at the first run start behind it.  From there on, a neighbour inside the segment
// and the partition window that is not yet known to be connected to i is tested; pairs that pass are united
// in a forest of the BLOCK (shared memory: interleaved events make every member a run of its own, and their
// quadratically many unions would each be a chain of global round trips), and the rest of that neighbour's
// run is skipped as well.  At the end every run start that found a smaller root hangs itself under it in
// the global forest: one global union per merged run, not per pair.
__global__ void __launch_bounds__(kClThreads)
k_cl_edges(ClusterArgs a) {
    constexpr int kSpan = kClTile + kClHalo;                            // 1280 positions in shared memory
    __shared__ unsigned long long s_key[kSpan];
    __shared__ int s_span[kSpan];
    __shared__ int s_uf[kSpan];                                         // the block's forest over span positions
    __shared__ unsigned s_brk[kSpan / 32 + 1];
    cl_mark(a, 0);
    const Packing pk = packing_of(a.meta);
    const int cur = sorted_buffer(pk);
    const unsigned long long *key = a.key[cur], *pay = a.pay[cur];
    const int i0 = blockIdx.x * kClTile;
    for (int t = threadIdx.x; t < kSpan; t += kClThreads) {
        const int j = i0 + t;
        s_key[t] = j < a.n ? key[j] : ~0ull;
        s_span[t] = j < a.n ? (int)(pay[j] >> 32) : 0;
    }
    if (threadIdx.x == 0) s_brk[kSpan / 32] = 0xffffffffu;              // sentinel: the scan stops at the end of the span
    __syncthreads();
    cl_mark(a, 1);
    for (int t = threadIdx.x; t < kSpan; t += kClThreads) {             // kSpan is a multiple of 32: whole warps
        const bool brk = t == 0 || !cl_linked(s_key[t - 1], s_key[t], s_span[t - 1], s_span[t], a);
        const unsigned m = __ballot_sync(0xffffffffu, brk);
        if ((threadIdx.x & 31) == 0) s_brk[t >> 5] = m;
        // every position under the start of its run -- or, when that lies before this warp's 32 positions,
        // under the position just before them (a smaller member of the same run: as good a parent)
        const unsigned own = m & (0xffffffffu >> (31 - (t & 31)));
        s_uf[t] = own ? (t & ~31) + 31 - __clz(own) : (t & ~31) - 1;
    }
    __syncthreads();
    cl_mark(a, 2);
    auto next_start = [&](int t) {                                      // first run start behind span position t
        int w = (t + 1) >> 5;
        unsigned m = s_brk[w] & (0xffffffffu << ((t + 1) & 31));
        while (!m) m = s_brk[++w];
        return (w << 5) + __ffs(m) - 1;                                 // >= kSpan when there is none inside the span
    };
    const float nf = (float)a.normalizer, mdf = (float)a.max_distance;
#pragma unroll 1
    for (int u = 0; u < kClItems; ++u) {
        const int ti = u * kClThreads + threadIdx.x, i = i0 + ti;
        if (i >= a.n) break;
        const unsigned long long ki = s_key[ti];
        const unsigned seg = (unsigned)(ki >> 32), c2 = (unsigned)ki;
        const int si = s_span[ti];
        int t = next_start(ti);
        for (;;) {
            const int j = i0 + t;
            if (j >= a.n) break;
            const bool in_smem = t < kSpan;
            const unsigned long long kj = in_smem ? s_key[t] : key[j];
            if ((unsigned)(kj >> 32) != seg) break;
            const unsigned d2 = (unsigned)kj - c2;                      // keys are sorted: non-negative
            if (d2 > a.window2) break;
            if (in_smem) {
                if (uf_find(s_uf, t) != uf_find(s_uf, ti) && cl_edge(d2, si, s_span[t], nf, mdf, a.normalizer, a.max_distance)) {
                    uf_unite(s_uf, ti, t);
                    t = next_start(t);                                  // the rest of j's run came with it
                    continue;
                }
            } else if (cl_edge(d2, si, (int)(pay[j] >> 32), nf, mdf, a.normalizer, a.max_distance)) {
                uf_unite(a.parent, i, j);                               // a window wider than the halo: straight to the global forest
            }
            ++t;
        }
    }
    cl_mark(a, 3);
    __syncthreads();
    cl_mark(a, 4);
    for (int t = threadIdx.x; t < kSpan; t += kClThreads) {
        if (i0 + t >= a.n || !(s_brk[t >> 5] >> (t & 31) & 1u)) continue;
        const int r = uf_find(s_uf, t);
        if (r != t) {
            if (a.dbg) atomicAdd((unsigned long long *)&a.dbg[(size_t)blockIdx.x * kClDbgMarks + 6], 1ull);
            uf_unite(a.parent, i0 + r, i0 + t);
        }
    }
    cl_mark(a, 5);
}

// every position learns its root, and every root the smallest original index below it.  Nearly every
// component lies inside one tile under a root of that tile: those are settled in shared memory (the parent
// words of the tile, one shared-memory atomicMin per position) and cost the global forest one atomicMin per
// root; only positions whose parent lies outside the tile walk the global forest.
__global__ void __launch_bounds__(kClThreads)
k_cl_label(ClusterArgs a) {
    __shared__ int s_par[kClTile], s_min[kClTile];
    const Packing pk = packing_of(a.meta);
    const unsigned long long *pay = a.pay[sorted_buffer(pk)];
    const int i0 = blockIdx.x * kClTile;
    int p0[kClItems], idx[kClItems];
#pragma unroll
    for (int u = 0; u < kClItems; ++u) {
        const int t = u * kClThreads + threadIdx.x, i = i0 + t;
        p0[u] = i < a.n ? a.parent[i] : -1;
        idx[u] = i < a.n ? (int)(unsigned)pay[i] : INT32_MAX;
        s_par[t] = p0[u];
        s_min[t] = INT32_MAX;
    }
    __syncthreads();
    int n_root = 0;
#pragma unroll
    for (int u = 0; u < kClItems; ++u) {
        const int t = u * kClThreads + threadIdx.x, i = i0 + t;
        if (i >= a.n) continue;
        int p = p0[u];
        // follow the parents while they stay inside the tile (shared memory); roots are fixed points
        while (p >= i0 && p != s_par[p - i0]) p = s_par[p - i0];
        if (p >= i0) {
            atomicMin(&s_min[p - i0], idx[u]);
        } else {
            p = uf_find(a.parent, p);
            atomicMin(a.minidx + p, idx[u]);
        }
        a.parent[i] = p;
        n_root += p0[u] == i;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kClItems; ++u) {
        const int t = u * kClThreads + threadIdx.x;
        if (s_min[t] != INT32_MAX) atomicMin(a.minidx + i0 + t, s_min[t]);
    }
    n_root = __reduce_add_sync(0xffffffffu, n_root);
    if ((threadIdx.x & 31) == 0 && n_root) atomicAdd(&a.meta->n_clusters, n_root);
}

__global__ void __launch_bounds__(kClThreads)
k_cl_write(ClusterArgs a) {
    const Packing pk = packing_of(a.meta);
    const unsigned long long *pay = a.pay[sorted_buffer(pk)];
    const int i0 = blockIdx.x * kClTile;
    int r[kClItems], idx[kClItems];
#pragma unroll
    for (int u = 0; u < kClItems; ++u) {
        const int i = i0 + u * kClThreads + threadIdx.x;
        r[u] = i < a.n ? a.parent[i] : -1;
        idx[u] = i < a.n ? (int)(unsigned)pay[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < kClItems; ++u) if (r[u] >= 0) r[u] = a.minidx[r[u]];
#pragma unroll
    for (int u = 0; u < kClItems; ++u)
        if (i0 + u * kClThreads + threadIdx.x < a.n) a.out[idx[u]] = r[u];
}

}  // namespace duet
