// Kernel set B, the path a call normally takes: five short launches chained with programmatic dependent launch,
// the signatures cross the memory system three times (columns in, 8-byte keys, 16-byte bucketed records).
//
// The general path (cluster_kernels.cuh) sorts all signatures with a five-pass radix sort in global memory and
// then runs four tile kernels over the sorted array.  At BASELINE.json's configs[2] (2 M signatures) every one
// of those ~20 launches is bounded by its chain of dependent round trips, not by bytes.  Here instead:
//
//   k_cl_max      maxima of the columns -> how many key bits the call needs; writes the 8-byte key
//                 contig | type | c2 of every signature and makes it a root of the global forest
//   k_cl_hist     histogram of the top B key bits (B <= 14: buckets of ~128+ signatures), one RED per signature
//   k_cl_scatter  every block scans the histogram for the bucket starts, then every signature is written ONCE as a 16-byte record into its bucket's stretch of the bucketed
//                 array (unordered inside the bucket; one returning atomic per signature).  A signature within
//                 the partition window of its bucket's lower boundary is also copied into that bucket's ZONE
//                 list: it is the halo of the bucket before.
//   k_cl_bucket   persistent 512-thread blocks take the non-empty buckets from a ticket counter.  A bucket (+ the zone of the next one)
//                 lives in shared memory from here on: radix-sorted by the remaining key bits, runs of linked
//                 sorted neighbours, the windowed scan over the other runs with a forest of the block, smallest
//                 original index per component -- and the cluster ids of every component that does not touch a
//                 zone are final and written.  Components that do are OPEN: their members go on a pending list
//                 and their zone members are united with the component's local label in a global forest over
//                 ORIGINAL indices (roots = minima), which is how the two blocks that see a zone signature meet.
//   k_cl_fix      pending members take the root of their label: the smallest index of the whole component.
//
// A bucket that does not fit shared memory (more than kBkCap signatures with its halo, or more than kZoneCap in a
// zone) sets `oversize`; the host then runs the general path for the call.  Same spec, same result: cluster ids
// do not depend on the order of equal keys, so neither path needs a stable sort order to agree with the other.
#pragma once

#include "cluster_kernels.cuh"

namespace duet {

constexpr int kSpUnroll = 4;                       // signatures per thread in flight (streaming kernels)
constexpr int kAgThreads = 1024;                   // k_cl_hist / k_cl_scatter: counters aggregated in shared memory
constexpr int kAgItems = 8;                        // signatures per thread there
constexpr int kBkMaxBits = 14;                     // up to 16384 buckets (64 KB of shared-memory counters)
constexpr int kBkTarget = 128;                     // signatures per bucket aimed at (non-empty ones hold several times that)
constexpr int kBkThreads = 512;                    // k_cl_bucket
constexpr int kBkCap = 2048;                       // own + halo signatures of a bucket in shared memory
constexpr int kBkSlotBits = 11;                    // log2(kBkCap)
constexpr int kBkItems = kBkCap / kBkThreads;      // 8
constexpr int kZoneCap = 256;                      // zone copies per bucket
constexpr int kFixBlocks = 148;

struct FastArgs {
    int n;
    const int *contig, *type, *start, *end;
    unsigned long long *key;       // [n] contig:16 | type:8 | 0:8 | c2:32
    ulonglong2 *rec;               // [n] bucketed records: x = packed key, y = span:32 | original index:32
    ulonglong2 *zone;              // [n_buckets][kZoneCap]
    unsigned *hist;                // [n_buckets] signatures per bucket
    unsigned *cursor;              // [n_buckets] reservation cursor of the scatter
    unsigned *zone_n;              // [n_buckets] zone copies (may exceed kZoneCap: oversize)
    int4 *bucket_list;             // [n_buckets] non-empty buckets: {bucket, first record, signatures, 0}
    int *parent;                   // [n] global forest over original indices, -1 = root
    int *pending;                  // [n] members of open components
    int *out;                      // [n] cluster id per original index
    ClMeta *meta;
    double max_distance, normalizer;
    unsigned window2;
    int bucket_bits_override;      // developer aid (DUET_CL_BITS), -1 = automatic
    long long *dbg;                // developer aid (DUET_CL_DBG): cycles per phase of k_cl_bucket, per block
};

// how the call splits its keys: bucket = packed >> shift, the rest orders a bucket.  The c2 field is given
// one bit more than the larger of (max c2, window) needs: two signatures of DIFFERENT (contig, type) segments
// then always lie further apart than the window, and "same segment" never has to be tested separately.
struct Split {
    int b2, bt, total, B, shift;
    __device__ __forceinline__ unsigned long long pack(unsigned c, unsigned t, unsigned c2) const {
        return ((unsigned long long)c << (bt + b2)) | ((unsigned long long)t << b2) | (unsigned long long)c2;
    }
};
__device__ __forceinline__ Split split_of(const ClMeta *m, int n, unsigned window2, int override_bits) {
    Split s;
    s.b2 = max(bits_for(m->max_c2), bits_for(window2)) + 1; s.bt = bits_for(m->max_type);
    s.total = s.b2 + s.bt + bits_for(m->max_contig);
    int want = n > kBkTarget ? bits_for((unsigned)((n - 1) / kBkTarget)) : 0;
    if (override_bits >= 0) want = override_bits;
    // a bucket must be wider than the window, so that a window crosses one boundary at most
    s.B = max(0, min(min(want, kBkMaxBits), s.total - (bits_for(window2) + 1)));
    s.shift = s.total - s.B;
    return s;
}

// ---- the global forest over original indices: -1 marks a root, parents are smaller indices ----
__device__ __forceinline__ int gf_find(volatile int *p, int x) {
    for (;;) {
        const int px = p[x];
        if (px < 0) return x;
        const int ppx = p[px];
        if (ppx >= 0) p[x] = ppx;
        x = px;
    }
}
__device__ __forceinline__ void gf_unite(int *p, int x, int y) {
    for (;;) {
        x = gf_find(p, x);
        y = gf_find(p, y);
        if (x == y) return;
        if (x > y) { const int t = x; x = y; y = t; }
        if (atomicCAS(p + y, -1, x) == -1) return;
    }
}

// ---- the streaming kernels ------------------------------------------------------------------------
__device__ __forceinline__ void fast_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fast_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__global__ void __launch_bounds__(kClThreads)
k_cl_max(FastArgs a) {
    fast_pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31;
    unsigned mc = 0, mt = 0, m2 = 0;
    bool bad = false;
    fast_pdl_wait();                                               // the previous call is done with key / parent / counters
    // a resident grid (8 blocks per SM) strides over the tiles: 7809 one-tile blocks spent a third of the kernel
    // being scheduled
    for (long long base = (long long)blockIdx.x * (kClThreads * kSpUnroll); base < a.n; base += (long long)gridDim.x * (kClThreads * kSpUnroll)) {
        int c[kSpUnroll], t[kSpUnroll], st[kSpUnroll], en[kSpUnroll];
#pragma unroll
        for (int u = 0; u < kSpUnroll; ++u) {
            const long long i = base + u * kClThreads + tid;
            c[u] = t[u] = st[u] = en[u] = 0;
            if (i < a.n) { c[u] = __ldcs(a.contig + i); t[u] = __ldcs(a.type + i); st[u] = a.start[i]; en[u] = __ldcs(a.end + i); }
        }
#pragma unroll
        for (int u = 0; u < kSpUnroll; ++u) {
            const long long i = base + u * kClThreads + tid;
            const long long s2 = (long long)st[u] + en[u];
            bad |= st[u] < 0 || en[u] < st[u] || s2 > 0xFFFFFFFFll || (unsigned)c[u] > 0xFFFFu || (unsigned)t[u] > 0xFFu;
            const unsigned cc = (unsigned)c[u] & 0xFFFFu, tt = (unsigned)t[u] & 0xFFu;
            mc = max(mc, cc); mt = max(mt, tt); m2 = max(m2, (unsigned)s2);
            if (i < a.n) {
                a.key[i] = ((unsigned long long)cc << 40) | ((unsigned long long)tt << 32) | (unsigned long long)(unsigned)s2;
                a.parent[i] = -1;
            }
        }
    }
    mc = __reduce_max_sync(0xffffffffu, mc); mt = __reduce_max_sync(0xffffffffu, mt); m2 = __reduce_max_sync(0xffffffffu, m2);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        if (mc > a.meta->max_contig) atomicMax(&a.meta->max_contig, mc);
        if (mt > a.meta->max_type) atomicMax(&a.meta->max_type, mt);
        if (m2 > a.meta->max_c2) atomicMax(&a.meta->max_c2, m2);
        if (bad) a.meta->bad = 1;
    }
    for (int b = blockIdx.x * kClThreads + tid; b < (1 << kBkMaxBits); b += gridDim.x * kClThreads) {
        a.hist[b] = 0; a.cursor[b] = 0; a.zone_n[b] = 0;
    }
}

__device__ __forceinline__ unsigned long long repack(const Split &sp, unsigned long long key) {
    return sp.pack((unsigned)(key >> 40), (unsigned)(key >> 32) & 0xFFu, (unsigned)key);
}

// Histogram of the bucket ids.  2 M REDs straight onto 16384 global counters cost 28 us (measured,
// tools/ubench_hist.cu: they queue on a few hundred cache lines); counted in shared memory first and flushed
// once per block and non-empty counter it is 12 us.
__global__ void __launch_bounds__(kAgThreads)
k_cl_hist(FastArgs a) {
    extern __shared__ unsigned s_ag[];                             // [1 << B]
    fast_pdl_trigger();
    fast_pdl_wait();                                               // maxima and keys are final
    const Split sp = split_of(a.meta, a.n, a.window2, a.bucket_bits_override);
    if (a.meta->bad || sp.shift + 1 + kBkSlotBits > 64) return;    // invalid input / keys too wide for a bucket's sort items
    const int nb = 1 << sp.B;
    const long long base = (long long)blockIdx.x * (kAgThreads * kAgItems);
    unsigned long long key[kAgItems];
#pragma unroll
    for (int u = 0; u < kAgItems; ++u) {
        const long long i = base + u * kAgThreads + threadIdx.x;
        key[u] = i < a.n ? a.key[i] : 0ull;
    }
    for (int b = threadIdx.x; b < nb; b += kAgThreads) s_ag[b] = 0;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kAgItems; ++u)
        if (base + u * kAgThreads + threadIdx.x < a.n) atomicAdd(&s_ag[(unsigned)(repack(sp, key[u]) >> sp.shift)], 1u);
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += kAgThreads) {
        const unsigned c = s_ag[b];
        if (c) atomicAdd(&a.hist[b], c);
    }
}

// Every signature into its bucket's stretch.  The block counts its own signatures per bucket in shared memory
// (the count a signature sees is its place among the block's), reserves one range per non-empty counter with a
// single returning atomic, and writes the records.
__global__ void __launch_bounds__(kAgThreads, 2)
k_cl_scatter(FastArgs a) {
    extern __shared__ unsigned s_ag[];                             // [1 << B]
    fast_pdl_trigger();
    const long long base = (long long)blockIdx.x * (kAgThreads * kAgItems);
    fast_pdl_wait();                                               // the histogram is final (and, transitively, the keys)
    const Split sp = split_of(a.meta, a.n, a.window2, a.bucket_bits_override);
    if (a.meta->bad || sp.shift + 1 + kBkSlotBits > 64) {          // invalid input / keys too wide for a bucket's sort items
        if (blockIdx.x == 0 && threadIdx.x == 0) a.meta->oversize = 1;
        return;
    }
    const int nb = 1 << sp.B;
    for (int b = threadIdx.x; b < nb; b += kAgThreads) s_ag[b] = 0;
    __syncthreads();
    // Two blocks of 1024 threads per SM (the whole grid resident at once) leave 32 registers per thread: the keys are
    // not kept across the phases -- they are read again for the write (out of L2) -- and a signature's place among
    // the block's (< 8192) is kept in 16 bits.
    unsigned rank2[kAgItems / 2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        constexpr int kH = kAgItems / 2;
        unsigned long long key[kH];
#pragma unroll
        for (int v = 0; v < kH; ++v) {
            const long long i = base + (half * kH + v) * kAgThreads + threadIdx.x;
            key[v] = i < a.n ? a.key[i] : 0ull;
        }
#pragma unroll
        for (int v = 0; v < kH; ++v) {
            const int u = half * kH + v;
            unsigned r = 0;
            if (base + u * kAgThreads + threadIdx.x < a.n) r = atomicAdd(&s_ag[(unsigned)(repack(sp, key[v]) >> sp.shift)], 1u);
            if (u & 1) rank2[u >> 1] |= r << 16; else rank2[u >> 1] = r;
        }
    }
    __syncthreads();
    {   // where the buckets start: every block scans the 2^B counters itself (64 KB out of L2) -- a kernel of one
        // block doing it for all of them took 22 us.  Block 0 also leaves the list of non-empty buckets.
        __shared__ unsigned long long s_scan[kAgThreads / 32];
        const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
        // warp w owns the `per` x 32 consecutive buckets from w * per * 32 on, lane l the buckets j * 32 + l of them:
        // coalesced loads, conflict-free shared memory (a thread owning 16 CONSECUTIVE counters cost 8 us in
        // 64-byte-strided accesses alone)
        const int per = max(nb / kAgThreads, 1);                   // <= 16
        const int wb = w * per * 32;
        // scanned together: signatures : 32 | smaller non-empty buckets : 16 | larger ones : 16.  The list of non-empty
        // buckets puts the LARGER ones first (more than four times the mean of an even spread): k_cl_bucket's blocks
        // draw tickets in list order, and a kernel that ends on its largest buckets ends on a tail (measured: the first
        // block done after 75 us, the last after 105)
        const unsigned large = max(256u, 4u * (unsigned)(a.n >> sp.B));
        auto packed = [&](unsigned c) { return ((unsigned long long)c << 32) | (c > large ? 1ull : c ? 0x10000ull : 0ull); };
        unsigned long long sum = 0;
#pragma unroll 4
        for (int j = 0; j < per; ++j) {
            const int b = wb + j * 32 + lane;
            sum += packed(b < nb ? __ldcg(a.hist + b) : 0u);
        }
        unsigned long long wsum = sum;                             // the warp's total
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
        if (lane == 0) s_scan[w] = wsum;
        __syncthreads();
        unsigned long long carry = 0, total = 0;
        for (int k = 0; k < kAgThreads / 32; ++k) { const unsigned long long v = s_scan[k]; total += v; if (k < w) carry += v; }
        const int n_large = (int)(total & 0xFFFFu), n_small = (int)((total >> 16) & 0xFFFFu);
        bool big = false;
        for (int j = 0; j < per; ++j) {                            // (the counters are read again: registers are what is scarce)
            const int b = wb + j * 32 + lane;
            const unsigned c = b < nb ? __ldcg(a.hist + b) : 0u;
            const unsigned long long x = packed(c);
            unsigned long long inc = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            const unsigned long long before = carry + inc - x;
            carry += __shfl_sync(0xffffffffu, inc, 31);
            if (b < nb) {
                const unsigned at = (unsigned)(before >> 32);
                const unsigned mine = s_ag[b];
                if (mine) s_ag[b] = at + atomicAdd(&a.cursor[b], mine);          // the block's range inside the bucket
                if (blockIdx.x == 0 && c) {
                    const int slot = c > large ? (int)(before & 0xFFFFu) : n_large + (int)((before >> 16) & 0xFFFFu);
                    a.bucket_list[slot] = make_int4(b, (int)at, (int)c, 0);
                }
                big |= c > (unsigned)kBkCap;
            }
        }
        if (blockIdx.x == 0 && tid == 0) a.meta->n_buckets = n_large + n_small;
        if (big) a.meta->oversize = 1;                             // k_cl_bucket and k_cl_fix will stand down
    }
    __syncthreads();
    const unsigned long long low_mask = sp.shift >= 64 ? ~0ull : ((1ull << sp.shift) - 1ull);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        constexpr int kH = kAgItems / 2;
        unsigned long long key[kH];
        int st[kH];
#pragma unroll
        for (int v = 0; v < kH; ++v) {
            const long long i = base + (half * kH + v) * kAgThreads + threadIdx.x;
            key[v] = i < a.n ? __ldcs(a.key + i) : 0ull;
            st[v] = i < a.n ? __ldcs(a.start + i) : 0;
        }
#pragma unroll
        for (int v = 0; v < kH; ++v) {
            const int u = half * kH + v;
            const long long i = base + u * kAgThreads + threadIdx.x;
            if (i >= a.n) continue;
            const unsigned long long pk = repack(sp, key[v]);
            const unsigned b = (unsigned)(pk >> sp.shift);
            const unsigned rank = (rank2[u >> 1] >> ((u & 1) * 16)) & 0xFFFFu;
            const unsigned span = (unsigned)key[v] - 2u * (unsigned)st[v];      // c2 - 2 start = end - start
            const ulonglong2 r = make_ulonglong2(pk, ((unsigned long long)span << 32) | (unsigned)i);
            a.rec[s_ag[b] + rank] = r;
            if (b > 0 && (pk & low_mask) <= a.window2) {           // inside the window of the bucket's lower boundary
                const unsigned z = atomicAdd(&a.zone_n[b], 1u);
                if (z < (unsigned)kZoneCap) a.zone[(size_t)b * kZoneCap + z] = r;
                else a.meta->oversize = 1;
            }
        }
    }
}

// ---- k_cl_bucket -----------------------------------------------------------------------------------
// What bounds this kernel is the latency of its short dependent chains (shared-memory loads, atomics, block
// barriers between the phases of a bucket), not instructions or bytes: 512 threads per bucket and three buckets
// per SM keep 48 warps resident, loops are unrolled for independent loads, and nothing is staged in registers
// across phases (<= 42 registers per thread).
constexpr int kCellBits = 12;                                       // 4096 cells per bucket for the ordering
struct BucketSmem {
    unsigned long long item[2][kBkCap];          // sort items: rel key << kBkSlotBits | slot
    int span[kBkCap], idx[kBkCap];               // by slot (= arrival order)
    unsigned cell[1 << kCellBits];               // ordering: members per cell, then where the cells end
    unsigned brk[kBkCap / 32 + 1];
    unsigned char open[kBkCap];
    unsigned wscan[kBkThreads / 32];
    int4 desc[2];                                // the bucket to work on / the next one: {bucket, first record, signatures, zone copies of the next bucket}
    int n_active;
};

__global__ void __launch_bounds__(kBkThreads, 3)
k_cl_bucket(FastArgs a) {
    extern __shared__ __align__(16) unsigned char s_bk_raw[];
    BucketSmem &S = *reinterpret_cast<BucketSmem *>(s_bk_raw);
    fast_pdl_trigger();
    fast_pdl_wait();                                               // the bucketed records and the zone lists are final
    if (a.meta->oversize) return;
    const Split sp = split_of(a.meta, a.n, a.window2, a.bucket_bits_override);
    const int nb = 1 << sp.B, n_list = a.meta->n_buckets;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float nf = (float)a.normalizer, mdf = (float)a.max_distance;
    const int cshift = kBkSlotBits + max(sp.shift + 1 - kCellBits, 0);      // item >> cshift = cell of the item
    int n_closed = 0;
    long long ph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, t_prev = 0;
    auto mark = [&](int k) { if (a.dbg && tid == 0) { const long long t = clock64(); ph[k] += t - t_prev; t_prev = t; } };
    if (a.dbg && tid == 0) {
        t_prev = clock64();
        long long g0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        a.dbg[(size_t)blockIdx.x * 12 + 10] = g0;                    // when the block started (ns)
    }
    // buckets are handed out by a ticket counter (their sizes differ tenfold between signature types); thread 0
    // draws the NEXT ticket and fetches that bucket's descriptor while the block works on the current one
    if (tid == 0) {
        const int tk = atomicAdd(&a.meta->ticket, 1);
        int4 d = make_int4(-1, 0, 0, 0);
        if (tk < n_list) { d = a.bucket_list[tk]; d.w = d.x + 1 < nb ? (int)a.zone_n[d.x + 1] : 0; }
        S.desc[0] = d;
    }
#pragma unroll
    for (int u = 0; u < (1 << kCellBits) / kBkThreads; ++u) S.cell[u * kBkThreads + tid] = 0;      // (re-zeroed at the end of every bucket)
    for (int it_no = 0;; ++it_no) {
        __syncthreads();                                            // the descriptor is in; the previous bucket is done with shared memory
        const int4 desc = S.desc[it_no & 1];                        // (the other slot takes the next one: no second barrier)
        if (desc.x < 0) break;
        int next_tk = 0;
        if (tid == 0) next_tk = atomicAdd(&a.meta->ticket, 1);      // consumed at the end of this bucket
        const int b = desc.x, m = desc.z;
        const unsigned zn = (unsigned)desc.w;
        const int M = m + (int)min(zn, (unsigned)kZoneCap);
        const bool fits = zn <= (unsigned)kZoneCap && M <= kBkCap;
        mark(0);
        if (!fits) {
            if (tid == 0) { a.meta->oversize = 1; S.desc[(it_no + 1) & 1] = make_int4(-1, 0, 0, 0); }      // the call takes the general path: stop
            continue;
        }
        int4 nd = make_int4(-1, 0, 0, 0);                           // thread 0: the next bucket's descriptor, fetched in steps below
        const unsigned long long base_key = (unsigned long long)b << sp.shift;
        const ulonglong2 *own = a.rec + desc.y, *halo = a.zone + (size_t)(b + 1) * kZoneCap;
        // ---- load: own records, then the zone of the next bucket; the members of every cell are counted on the way.
        //      Ordering = cells first (one counting pass, arrival order inside a cell), then every item's exact
        //      place inside its cell by counting the smaller items there: a cell holds the signatures of about one
        //      window, so that is a few dozen comparisons per item instead of three radix passes ----
#pragma unroll 4
        for (int t = tid; t < M; t += kBkThreads) {
            const ulonglong2 r = t < m ? __ldcs(own + t) : __ldcs(halo + t - m);
            const unsigned long long it = ((r.x - base_key) << kBkSlotBits) | (unsigned long long)t;
            S.item[0][t] = it;
            S.span[t] = (int)(r.y >> 32);
            S.idx[t] = (int)(unsigned)r.y;
            atomicAdd(&S.cell[(unsigned)(it >> cshift)], 1u);
        }
        __syncthreads();
        mark(1);
        mark(2);
        if (tid == 0 && next_tk < n_list) nd = a.bucket_list[next_tk];       // the ticket has long arrived
        {
            constexpr int kPer = (1 << kCellBits) / kBkThreads;     // 8 consecutive cells per thread
            uint4 *cv = reinterpret_cast<uint4 *>(S.cell + tid * kPer);
            uint4 c0 = cv[0], c1 = cv[1];
            const unsigned tot = c0.x + c0.y + c0.z + c0.w + c1.x + c1.y + c1.z + c1.w;
            unsigned inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned x = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += x;
            }
            if (lane == 31) S.wscan[w] = inc;
            __syncthreads();
            unsigned run = inc - tot;
            for (int k = 0; k < w; ++k) run += S.wscan[k];
            uint4 s0, s1;
            s0.x = run; s0.y = s0.x + c0.x; s0.z = s0.y + c0.y; s0.w = s0.z + c0.z;
            s1.x = s0.w + c0.w; s1.y = s1.x + c1.x; s1.z = s1.y + c1.y; s1.w = s1.z + c1.z;
            cv[0] = s0; cv[1] = s1;
        }
        __syncthreads();
        mark(3);
#pragma unroll 4
        for (int t = tid; t < M; t += kBkThreads) {
            const unsigned long long it = S.item[0][t];
            S.item[1][atomicAdd(&S.cell[(unsigned)(it >> cshift)], 1u)] = it;       // cell[c] ends up at the cell's end
        }
        __syncthreads();
        mark(4);
        // (the items of a cell agree above cshift <= 32 bits: comparing the low words is comparing the items)
        {
            const unsigned *lo = reinterpret_cast<const unsigned *>(S.item[1]);
            for (int p = tid; p < M; p += kBkThreads) {
                const unsigned long long it = S.item[1][p];
                const unsigned c = (unsigned)(it >> cshift), me = (unsigned)it;
                const int s0 = c ? (int)S.cell[c - 1] : 0, s1 = (int)S.cell[c];
                int rank = 0, q = s0;
                if (cshift <= 32) {
                    for (; q + 4 <= s1; q += 4) {
                        const unsigned x0 = lo[2 * q], x1 = lo[2 * q + 2], x2 = lo[2 * q + 4], x3 = lo[2 * q + 6];
                        rank += (x0 < me) + (x1 < me) + (x2 < me) + (x3 < me);
                    }
                    for (; q < s1; ++q) rank += lo[2 * q] < me;
                } else {
                    for (; q < s1; ++q) rank += S.item[1][q] < it;
                }
                S.item[0][s0 + rank] = it;
            }
        }
        __syncthreads();
        mark(5);
        if (tid == 0 && nd.x >= 0) nd.w = nd.x + 1 < nb ? (int)a.zone_n[nd.x + 1] : 0;      // ... and so has the descriptor
        const unsigned long long *item = S.item[0];
        int *uf = reinterpret_cast<int *>(S.item[1]);               // the free buffer: forest, then the minima
        int *mn = uf + kBkCap;
        auto rel_of = [&](int k) { return item[k] >> kBkSlotBits; };
        auto slot_of = [&](int k) { return (int)(item[k] & (kBkCap - 1)); };
        // ---- runs of linked neighbours ----
        const int Mw = (M + 31) & ~31;
        if (tid == 0) { S.brk[Mw >> 5] = 0xffffffffu; S.n_active = 0; }      // sentinel: the scan stops behind the last position
        for (int k = tid; k < Mw; k += kBkThreads) {
            bool brk = true;
            if (k > 0 && k < M) {
                const unsigned long long ik = item[k], ip = item[k - 1];
                const unsigned long long d = (ik >> kBkSlotBits) - (ip >> kBkSlotBits);       // other segment => further than the window (Split)
                brk = !(d <= a.window2 &&
                        cl_edge((unsigned)d, S.span[ip & (kBkCap - 1)], S.span[ik & (kBkCap - 1)], nf, mdf, a.normalizer, a.max_distance));
            }
            const unsigned mk = __ballot_sync(0xffffffffu, brk);
            if (lane == 0) S.brk[k >> 5] = mk;
            const unsigned mine = mk & (0xffffffffu >> (31 - (k & 31)));
            uf[k] = mine ? (k & ~31) + 31 - __clz(mine) : (k & ~31) - 1;
            mn[k] = INT32_MAX;
            S.open[k] = 0;
        }
        __syncthreads();
        mark(6);
        // ---- the other runs inside the window: own positions scan forward ----
        auto next_start = [&](int k) {
            int ww = (k + 1) >> 5;
            unsigned mk = S.brk[ww] & (0xffffffffu << ((k + 1) & 31));
            while (!mk) mk = S.brk[++ww];
            return (ww << 5) + __ffs(mk) - 1;                       // >= M when there is none
        };
        // nearly every position has nothing but its own run inside the window; the few that do not (interleaved
        // events) are collected and then scanned by a whole warp each, 32 candidates at a time
        int *active = reinterpret_cast<int *>(S.cell);              // free since the ordering
        for (int k = tid; k < M; k += kBkThreads) {
            if (slot_of(k) >= m) continue;                          // a halo copy: its own bucket scans for it
            const int t = next_start(k);
            if (t < M && rel_of(t) - rel_of(k) <= a.window2) active[atomicAdd(&S.n_active, 1)] = k;
        }
        __syncthreads();
        for (int ai = w; ai < S.n_active; ai += kBkThreads / 32) {
            const int k = active[ai];
            const unsigned long long rk = rel_of(k);
            const int sk = S.span[slot_of(k)];
            for (int t0 = next_start(k); t0 < M; t0 += 32) {
                const int t = t0 + lane;
                const unsigned long long d = t < M ? rel_of(t) - rk : ~0ull;
                const bool in_window = d <= a.window2;
                if (in_window && uf_find(uf, t) != uf_find(uf, k) &&
                    cl_edge((unsigned)d, sk, S.span[slot_of(t)], nf, mdf, a.normalizer, a.max_distance))
                    uf_unite(uf, k, t);
                if (!__shfl_sync(0xffffffffu, (int)in_window, 31)) break;      // sorted: behind the first miss nothing is in the window
            }
        }
        __syncthreads();
        mark(7);
        // ---- smallest original index per component (a warp holds 32 consecutive sorted positions: equal roots
        //      sit next to each other, one shared-memory atomic per stretch); components that touch a zone are open ----
#pragma unroll
        for (int u = 0; u < (1 << kCellBits) / kBkThreads; ++u) S.cell[u * kBkThreads + tid] = 0;      // the list of active positions is dead: cells for the next bucket
        for (int k0 = 0; k0 < M; k0 += kBkThreads) {
            const int k = k0 + tid;
            const bool live = k < M;
            int rt = -1 - lane, v = INT32_MAX;
            if (live) {
                const int sl = slot_of(k);
                rt = uf_find(uf, k);
                v = S.idx[sl];
                if (sl >= m || (b > 0 && rel_of(k) <= a.window2)) S.open[rt] = 1;      // a halo copy, or one of my zone members
            }
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v2 = __shfl_up_sync(0xffffffffu, v, o), r2 = __shfl_up_sync(0xffffffffu, rt, o);
                if (lane >= o && r2 == rt) v = min(v, v2);
            }
            const int r_next = __shfl_down_sync(0xffffffffu, rt, 1);
            if (live) {
                uf[k] = rt;
                if (lane == 31 || r_next != rt) atomicMin(&mn[rt], v);
            }
        }
        __syncthreads();
        mark(8);
        for (int k0 = 0; k0 < M; k0 += kBkThreads) {
            const int k = k0 + tid;
            const bool live = k < M;
            const int rt = live ? uf[k] : 0, sl = live ? slot_of(k) : 0;
            const bool is_own = live && sl < m, is_open = live && S.open[rt];
            const int label = live ? mn[rt] : 0, me = live ? S.idx[sl] : 0;
            if (is_own) a.out[me] = label;
            // open components: own members wait for k_cl_fix, zone members tie the label into the global forest
            const unsigned pend = __ballot_sync(0xffffffffu, is_own && is_open);
            if (pend) {
                int pbase = 0;
                if (lane == 0) pbase = atomicAdd(&a.meta->n_pending, __popc(pend));
                pbase = __shfl_sync(0xffffffffu, pbase, 0);
                if (is_own && is_open) a.pending[pbase + __popc(pend & ((1u << lane) - 1u))] = me;
            }
            if (is_open && (sl >= m || rel_of(k) <= a.window2) && me != label) gf_unite(a.parent, me, label);
            n_closed += live && rt == k && !is_open;
        }
        mark(9);
        if (tid == 0) S.desc[(it_no + 1) & 1] = nd;                 // read behind the barrier at the top
    }
    if (a.dbg && tid == 0) {
        for (int k = 0; k < 10; ++k) a.dbg[(size_t)blockIdx.x * 12 + k] = ph[k];
        long long g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        a.dbg[(size_t)blockIdx.x * 12 + 11] = g1;                    // ... and when it was done
    }
    n_closed = __reduce_add_sync(0xffffffffu, n_closed);
    if (lane == 0 && n_closed) atomicAdd(&a.meta->n_clusters, n_closed);
}

// ---- k_cl_fix --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kClThreads)
k_cl_fix(FastArgs a) {
    if (a.meta->oversize) return;
    const int np = a.meta->n_pending;
    int roots = 0;
    for (int k = blockIdx.x * kClThreads + threadIdx.x; k < np; k += gridDim.x * kClThreads) {
        const int i = a.pending[k];
        const int r = gf_find(a.parent, a.out[i]);
        a.out[i] = r;
        roots += r == i;
    }
    roots = __reduce_add_sync(0xffffffffu, roots);
    if ((threadIdx.x & 31) == 0 && roots) atomicAdd(&a.meta->n_clusters, roots);
}

}  // namespace duet
