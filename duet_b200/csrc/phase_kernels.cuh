// sm_100a kernels of the sv_phasing hot path.
//
// Reference behaviour restated per kernel (citations: /root/reference/src/duet/sv_phasing_fn.py):
//   k_init     start-of-call state in one sequential sweep: EMPTY slot table, join results -1 (the Bloom
//              filter is handed back zeroed by k_reduce) -- which also leaves all three L2 resident for the scattered traffic that follows
//   k_table    the dict-insert side of the JOIN (:26-29 keyed by QNAME) -- here the SMALL side is
//              inserted: every support-read name of the SVs (:46-48) claims a slot (one CAS) and sets its
//              two Bloom-filter bits (one RED)
//   k_probe    the haplotagged reads streamed once through the contig's Bloom filter (shared memory);
//              the block then looks its survivors up in the slot table: a hit records the row index on
//              every support-read entry of that name with atomicMax == "a later row overwrites an
//              earlier one" (:29)
//   k_reduce   per SV: gather the joined reads' tags, class = #distinct PS (:192-194), one-PS candidate
//              (:195-203), class-1 counts and score sums (:74-84), per-PS statistics of class-2 SVs in
//              first-seen order (:85-105)
//   k_tail     one thread-block CLUSTER per contig: its sorted unique one-PS list (:107), then per SV the
//              in-set PS with most reads (:99-105), nearest-PS fallback (:106-111), features (:112-139),
//              the T1-T5 tree (:142-183), then the contig's emission order (:206-229) and counters
//   k_oneps / k_predict / k_order   the same three steps as kernels of their own, for contigs with more
//              SVs than a cluster holds
//
// What bounds these kernels at WGS size is not bandwidth (the whole problem is ~130 MB).  Measured on
// B200 (tools/ubench_atomics.cu, profiles/r2_ubench.txt): a kernel that does ONE scattered access per
// element costs ~8 us however few elements it has; an SM issues ~0.5 scattered atomics or ~0.8 scattered
// loads / stores per clock (shared-memory atomics queue for the same unit); 400 k scattered accesses to lines
// that are not in L2 cost ~8 us more than to lines that are, while WRITING 38 MB front to back costs ~6 us;
// a dependent global access under load is a 1-2 us round trip.  So: (a) the join table is made L2 resident
// by the sequential sweep that initialises it, never by the scattered accesses that use it; (b) as few
// scattered operations per element as possible (one CAS + one RED per name, one slot load per candidate,
// one 16-byte record per joined read); (c) every kernel requests everything independent at once, starts
// from a host-built tile descriptor instead of looking its contig up, and does what does not need its
// predecessor's output before griddepcontrol.wait; (d) the three per-contig steps are one cluster launch.
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/duet_b200.h"

namespace duet {

constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr long long kNoCand = 0x7FFFFFFFFFFFFFFFll;   // "no one-PS candidate" (sorts last)
constexpr long long kNone = (long long)0x8000000000000000ull;
constexpr int kThreads = 256;                         // every kernel
constexpr int kMaxDistinct = 32;                      // distinct in-set PS per SV handled in smem
constexpr int kReduceLanesSparse = 8;                 // lanes per SV in k_reduce, typical support lists (~16 reads)
constexpr int kReduceLanesDense = 32;                 // ... dense lists (mean > 32 reads)
constexpr int kPredictPerBlock = 64;
constexpr int kSortSmemBytes = 16384;                 // slow-path sort tile in shared memory
constexpr int kBloomMaxWords = 16384;                 // 64 KB of shared memory per shard filter: two k_probe blocks per SM

// One slot of the join table: 16 bytes, loaded in one request.
struct __align__(16) Slot {
    unsigned long long key;   // low 64 bits of the name hash, kEmptyKey when free
    int first;                // the support-read entry that claimed the slot
    int head;                 // further entries carrying the same name (chain through `next`), -1 none
};

// include/duet_b200.h :: duet_read_tag as the device reads it (one 16-byte load)
struct __align__(16) ReadTag { int ps, pc; unsigned chk, hp; };       // hp: low byte

// host-built descriptors: what a block needs to know about its tile, in one 64- / 32-byte load
struct PredictTile { int sv0, sv1, shard, b, n, pad[3]; };        // SVs [sv0, sv1) of ONE shard = SVs [b, b + n)
struct BuildTile { int lo, hi, base, mask, bmo, bmw, pad[2]; };   // kThreads * U consecutive support reads: shards lo..hi
struct ProbeTile { long long r0, r1; int shard, base, mask, bmo, bmw, pad; };   // rows [r0, r1) of ONE contig + its table / filter

constexpr int kC2Max = 32;   // distinct PS per class-2 SV a record can hold (k_reduce<G> fills up to c2_cap(G); more -> warp fallback)
__host__ __device__ constexpr int c2_cap(int G) { return G == 32 ? 32 : 16; }     // dense batches (one warp per SV) keep 32, the others 16 (shared memory)
struct C2Ent { int ps, tot, n1, n2; long long s1, s2; int bad, pad; };
struct C2Rec { int n_d, overflow, pad[2]; C2Ent d[kC2Max]; };

struct DevStatus {          // device -> host error report
    int code;               // first DUET_ERR_* seen (atomicCAS from 0)
    int sv;                 // SV index it was seen at
    long long detail;       // offending value (HP, key, ...)
};

// Everything a kernel needs; passed by value.
// developer switches (environment DUET_FLAGS, read once per handle); 0 in production
enum {
    kFlagNoWarm = 2,          // k_init does not warm L2 with the input columns k_reduce / k_tail start from
    kFlagNoStreamHint = 4,    // k_probe's key stream without the L2 evict-first hint
    kFlagFill2 = 16,          // host: load factor <= 1/2 always
    kFlagSplitTail = 32,      // host: k_oneps / k_predict / k_order instead of k_tail
    kFlagTwoBranch = 64,      // k_bloom -> k_stream beside k_init -> k_table, meeting at k_resolve: set by the host for big calls, or forced here
    kFlagSerialChain = 128,   // host: never the two-branch chain
    kFlagNoTagPrefetch = 256, // k_probe does not prefetch the candidates' tag records into L2
};

struct PhaseArgs {
    int n_shards;
    int n_reads, n_svs, n_joins;
    int flags;
    long long n_slots;
    // inputs (device)
    const long long *read_off;   // [n_shards+1]
    const long long *sv_off;     // [n_shards+1]
    const long long *join_off;   // [n_shards+1] csr_off at the shard boundaries (derived at upload)
    const PredictTile *predict_tiles;   // [sum over shards of ceil(n / kPredictPerBlock)] = k_predict grid
    const BuildTile *build_tiles;   // [k_table grid]
    const ProbeTile *probe_tiles;   // [n_probe_tiles] = k_probe grid
    int n_probe_tiles;
    ulonglong2 *cand_list;          // [R] rows that passed their contig's filter, (key, row): tile t appends at [r0(t), ...)
    int *cand_count;                // [n_probe_tiles] candidates of each tile (k_stream -> k_resolve; unused by the fused k_probe)
    const unsigned long long *read_key;
    const ReadTag *read_tag;
    const int *sv_pos, *sv_svlen, *sv_svread, *sv_refread;
    const uint8_t *sv_flags;
    const int *sv_group;
    const long long *csr_off;
    const unsigned long long *csr_key;
    const unsigned *csr_chk;
    // join table: shard s owns slots [tab_off[s], tab_off[s] + tab_mask[s] + 1)
    const int *tab_off;          // [n_shards]
    const int *tab_mask;         // [n_shards]
    Slot *tab;                   // [n_slots] set to all-ones (= free) by k_init
    int *next;                   // [J] next entry carrying the same name, -1 none
    // per-shard Bloom filter over the support-read names: words [bm_off[s], bm_off[s] + bm_wmask[s] + 1)
    const int *bm_off;           // [n_shards]
    const int *bm_wmask;         // [n_shards] (power of two) - 1
    unsigned *bitmap;            // zeroed by k_init
    long long n_bm_words;
    // per-SV intermediates / outputs (device)
    int *join_row;               // [J] row each support read joined to (atomicMax by k_probe), -1 = miss
    int *n_hit;                  // [S] joined reads of the SV
    long long *cand;             // [S] one-PS candidate or kNoCand
    int *oneps;                  // [S] shard s: sorted unique list at [sv_off[s], +oneps_n[s])
    int *oneps_n;                // [n_shards]
    C2Rec *c2rec;                // [S] per-PS statistics of class-2 SVs (k_reduce -> k_tail), c2_stride bytes apart:
    int c2_stride;               //     a record holds the header and c2_cap(reduce lanes) entries, not all kC2Max
    const int *heavy_sv;         // dense batches: the SVs with more than kHeavyReads support reads (k_reduce_heavy), or NULL
    int n_heavy;
    long long *sort_scratch;     // [4*S] global tile for slow-path sorts of big shards
    uint8_t *gt, *cls;
    int *ps, *hap1, *hap2, *hap0, *allhap;
    long long *totsc1, *totsc2;
    double *features;            // [6][S]
    int *order;                  // [S] per-shard regions
    int *n_emit;                 // [n_shards]
    long long *shard_counts;     // [n_shards][8]
    DevStatus *status;
    long long *dbg;              // optional per-block timestamps (duet_debug_timers), NULL in production
};

__device__ __forceinline__ C2Rec *c2_at(const PhaseArgs &a, int sv) {
    return reinterpret_cast<C2Rec *>(reinterpret_cast<char *>(a.c2rec) + (size_t)sv * (size_t)a.c2_stride);
}

// ---- programmatic dependent launch: the kernels of one call are launched back to back with the
// programmatic-serialization attribute, so a kernel's blocks are scheduled as soon as every block of its
// predecessor has STARTED (pdl_trigger) and do their own set-up -- tile descriptors, barrier init, first
// copies of the INPUT columns -- under the predecessor's tail; pdl_wait() returns once the predecessor
// has completed and its writes are visible.  Every kernel calls both, so completion is transitive.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- optional instrumentation: thread 0 of every block stamps (globaltimer, clock64) at mark k ----
constexpr int kDbgBlocks = 2048, kDbgMarks = 8;
__device__ __forceinline__ void dbg_mark(const PhaseArgs &a, int kernel, int k) {
    if (a.dbg && threadIdx.x == 0 && blockIdx.x < kDbgBlocks) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        long long *p = a.dbg + (((size_t)kernel * kDbgBlocks + blockIdx.x) * kDbgMarks + k) * 2;
        p[0] = t;
        p[1] = clock64();
    }
}

__constant__ duet_thresholds c_thr;

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ void report(DevStatus *st, int code, int sv, long long detail) {
    if (atomicCAS(&st->code, 0, code) == 0) {
        st->sv = sv;
        st->detail = detail;
    }
}

// largest s with off[s] <= x  (off is non-decreasing, off[0] == 0, x < off[n])
__device__ __forceinline__ int shard_of(const long long *__restrict__ off, int n, long long x) {
    int lo = 0, hi = n;          // invariant: off[lo] <= x < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ unsigned slot_hash(unsigned long long key) {
    // keys are already well mixed 64-bit hashes; fold so both halves matter
    return (unsigned)(key ^ (key >> 32));
}

template <typename T>
__device__ __forceinline__ T group_sum(T v, unsigned mask, int width) {
    for (int o = width >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
template <typename T>
__device__ __forceinline__ T group_min(T v, unsigned mask, int width) {
    for (int o = width >> 1; o > 0; o >>= 1) { T u = __shfl_xor_sync(mask, v, o); v = u < v ? u : v; }
    return v;
}
template <typename T>
__device__ __forceinline__ T group_max(T v, unsigned mask, int width) {
    for (int o = width >> 1; o > 0; o >>= 1) { T u = __shfl_xor_sync(mask, v, o); v = u > v ? u : v; }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) { return group_sum(v, 0xffffffffu, 32); }

struct OpSum { template <typename T> __device__ T operator()(T a, T b) const { return a + b; } };
struct OpMax { template <typename T> __device__ T operator()(T a, T b) const { return a > b ? a : b; } };
struct OpLast { __device__ long long operator()(long long a, long long b) const { return b != (long long)0x8000000000000000ull ? b : a; } };

// exclusive block scan (blockDim.x == kThreads); *total receives the reduction over the block
template <typename T, typename Op>
__device__ T block_scan_exclusive(T v, T identity, Op op, T *total) {
    __shared__ T warp_tot[kThreads / 32];
    __shared__ T s_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = op(n, inc);
    }
    T exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = identity;
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T run = identity;
        for (int i = 0; i < kThreads / 32; ++i) { const T t = warp_tot[i]; warp_tot[i] = run; run = op(run, t); }
        s_total = run;
    }
    __syncthreads();
    const T res = op(warp_tot[w], exc);
    if (total) *total = s_total;
    __syncthreads();
    return res;
}

// block-wide bitonic sort of n_pad (power of two) elements, in shared or global memory
template <typename T>
__device__ void block_bitonic_sort(T *v, int n_pad) {
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n_pad >> 1); t += blockDim.x) {
                const int i = 2 * t - (t & (j - 1));
                const int p = i + j;
                const bool up = (i & k) == 0;
                const T x = v[i], y = v[p];
                if ((y < x) == up) { v[i] = y; v[p] = x; }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

// ------------------------------------------------------------------------------------------
// Bloom filter bit pattern of a key: one 32-bit word, two bits (a probe is one shared-memory load)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned bloom_word(unsigned long long key, unsigned wmask) {
    return (unsigned)(key >> 40) & wmask;
}
__device__ __forceinline__ unsigned bloom_bits(unsigned long long key) {
    return (1u << ((unsigned)(key >> 5) & 31u)) | (1u << ((unsigned)(key >> 10) & 31u));
}

// ---- mbarrier / bulk-copy (TMA) primitives ------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// one thread: `bytes` (multiple of 16) from 16-byte aligned global memory into shared memory; completion
// is signalled on `bar` (complete_tx)
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// the same, with an L2 eviction-priority hint: the read-key stream is touched exactly once, so its lines are marked
// evict-first and leave the slot table, the join rows and the prefetched tag records alone (C2: 95.1 -> 93.7 us,
// C5 share: 326 -> 323 us per call)
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_load_stream(void *dst, const void *src, unsigned bytes, unsigned long long *bar, unsigned long long policy) {
    if (policy == 0ull) { bulk_load(dst, src, bytes, bar); return; }          // developer switch: no hint
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// ------------------------------------------------------------------------------------------
// k_init: start-of-call state in one sequential sweep: EMPTY slots (all ones), zero filter words, join
// results -1.  Sequential 16-byte stores are the cheap way to get all three L2 resident for the scattered
// traffic of k_table / k_probe (writing 38 MB front to back: ~6 us; 400 k scattered accesses to lines that
// have to come from HBM: ~8 us on top of the same accesses to resident lines -- measured).  The input
// columns the later kernels start from are pulled into L2 along the way (prefetch hints).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void warm_l2(const void *p, size_t bytes, long long tid, long long n_threads) {
    const char *c = reinterpret_cast<const char *>(p);
    const long long lines = (long long)((bytes + 127) >> 7);
    for (long long i = tid; i < lines; i += n_threads) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + (i << 7)));
}

__global__ void __launch_bounds__(kThreads)
k_init(PhaseArgs a) {
    pdl_trigger();
    pdl_wait();                                                  // first kernel of a call: returns at once
    const long long stride = (long long)gridDim.x * kThreads;
    const long long t = (long long)blockIdx.x * kThreads + threadIdx.x;
    const uint4 ones = make_uint4(~0u, ~0u, ~0u, ~0u);
    uint4 *tab = reinterpret_cast<uint4 *>(a.tab);
    for (long long i = t; i < a.n_slots; i += stride) tab[i] = ones;
    uint4 *jr = reinterpret_cast<uint4 *>(a.join_row);           // allocation padded to 16 bytes
    for (long long i = t; i < ((long long)a.n_joins + 3) / 4; i += stride) jr[i] = ones;
    // (the Bloom filter is zero already: duet_phase_upload clears it once and k_reduce hands it back clean)
    if (!(a.flags & kFlagNoWarm)) {
        const size_t S = (size_t)a.n_svs, J = (size_t)a.n_joins;
        warm_l2(a.csr_off, (S + 1) * 8, t, stride);
        if (a.csr_chk) warm_l2(a.csr_chk, J * 4, t, stride);
        warm_l2(a.sv_svlen, S * 4, t, stride); warm_l2(a.sv_svread, S * 4, t, stride); warm_l2(a.sv_flags, S, t, stride);
        warm_l2(a.sv_pos, S * 4, t, stride); warm_l2(a.sv_refread, S * 4, t, stride);
        if (a.sv_group) warm_l2(a.sv_group, S * 4, t, stride);
    }
}

// ------------------------------------------------------------------------------------------
// The build side: k_table claims a slot per support-read name and sets the name's filter bits.  U names
// per thread, all of their atomics in flight together, and a grid that is resident in ONE wave: every block
// has its keys in registers before k_init is over (they are requested before griddepcontrol.wait).
// Dependent chain of a k_table thread: [keys, tile descriptor] -> CAS (-> CAS on a collision) -> stores.
// ------------------------------------------------------------------------------------------
// which shard (table range, filter range) support-read entry j of this tile belongs to
__device__ __forceinline__ void build_where(const PhaseArgs &a, const BuildTile &t, long long j, int &base,
                                            unsigned &mask, int &bmo, unsigned &bmw) {
    base = t.base; mask = (unsigned)t.mask; bmo = t.bmo; bmw = (unsigned)t.bmw;
    if (t.lo != t.hi) {                                          // the tile straddles a contig boundary
        const int s = t.lo + shard_of(a.join_off + t.lo, t.hi - t.lo + 1, j);
        base = __ldg(a.tab_off + s); mask = (unsigned)__ldg(a.tab_mask + s);
        bmo = __ldg(a.bm_off + s); bmw = (unsigned)__ldg(a.bm_wmask + s);
    }
}

template <int U>
__global__ void __launch_bounds__(kThreads)
k_table(PhaseArgs a) {
    dbg_mark(a, 0, 0);
    const long long j0 = (long long)blockIdx.x * (kThreads * U) + threadIdx.x;
    unsigned long long key[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long j = j0 + u * kThreads;
        key[u] = j < a.n_joins ? __ldcs(a.csr_key + j) : 0ull;
    }
    const BuildTile t = a.build_tiles[blockIdx.x];
    unsigned p[U], mask[U], bmw[U];
    int base[U], bmo[U];
    unsigned pend = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long long j = j0 + u * kThreads;
        if (j >= a.n_joins) continue;
        build_where(a, t, j, base[u], mask[u], bmo[u], bmw[u]);
        p[u] = slot_hash(key[u]) & mask[u];
        pend |= 1u << u;
    }
    pdl_trigger();
    pdl_wait();                                                  // the slots and the filter words are initialised
    dbg_mark(a, 0, 1);
    if (!(a.flags & kFlagTwoBranch)) {                           // (two-branch mode: k_bloom has set them)
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (pend >> u & 1u) atomicOr(a.bitmap + bmo[u] + (int)bloom_word(key[u], bmw[u]), bloom_bits(key[u]));    // fire and forget
    }
    while (pend) {
        unsigned long long prev[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (pend >> u & 1u) prev[u] = atomicCAS(&a.tab[base[u] + p[u]].key, kEmptyKey, key[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!(pend >> u & 1u)) continue;
            const int j = (int)(j0 + u * kThreads);
            Slot *sl = a.tab + base[u] + p[u];
            if (prev[u] == kEmptyKey) { sl->first = j; pend &= ~(1u << u); }
            else if (prev[u] == key[u]) { a.next[j] = atomicExch(&sl->head, j); pend &= ~(1u << u); }
            else p[u] = (p[u] + 1) & mask[u];
        }
    }
    dbg_mark(a, 0, 2);
}

// two-branch mode: the filter bits alone, so that the read stream can start while k_table is still claiming slots
__global__ void __launch_bounds__(kThreads)
k_bloom(PhaseArgs a, int per_thread) {
    const BuildTile t = a.build_tiles[blockIdx.x];
    const long long j0 = (long long)blockIdx.x * (kThreads * per_thread) + threadIdx.x;
    pdl_trigger();
    pdl_wait();                                                  // the filter words are zero
    for (int u = 0; u < per_thread; ++u) {
        const long long j = j0 + (long long)u * kThreads;
        if (j >= a.n_joins) break;
        const unsigned long long key = __ldcs(a.csr_key + j);
        int base, bmo; unsigned mask, bmw;
        build_where(a, t, j, base, mask, bmo, bmw);
        atomicOr(a.bitmap + bmo + (int)bloom_word(key, bmw), bloom_bits(key));
    }
}

// ------------------------------------------------------------------------------------------
// k_probe: the haplotagged reads are STREAMED once by one block per tile -- a tile is a row range of ONE
// contig, sized so that the grid is about two blocks per SM.  A producer warp keeps a ring of 16 KB key
// tiles in flight with bulk asynchronous copies (TMA, cp.async.bulk + mbarrier complete_tx); the consumer
// warps take a tile as soon as its barrier flips and hand the stage back through an `empty` barrier --
// no block-wide synchronisation inside the stream.  The ring is full before k_table is over; the contig's
// Bloom filter arrives in shared memory by ONE bulk copy issued the moment k_table is known to be complete,
// so ~90 % of the rows (reads that support no SV) never leave the SM.  The survivors are appended to the
// block's private stretch of the candidate list, one 16-byte record each, and their tag records start
// moving into L2; when the stream is done the whole block resolves its candidates against the slot table,
// kResolveUnroll per thread in flight together.
// (Measured and dropped in round 2: four dedicated resolver warps following the list while twelve consumer warps
// stream -- entries handed out / written counters, batches of 128 claimed below the complete prefix, the consumers
// joining at the end.  Parity green, but C2 93.7 -> 98.4 us and the C5 share 323 -> 335 us: what resolves a
// candidate quickly is the number of slot loads in flight, and the whole block after the stream has 2176 of them
// where the resolver warps had 512 and slowed the stream they ran beside.  One thing learnt on the way: entries
// written by other warps of the block must be read back with PLAIN loads after a block-scope fence -- ld.cg goes
// to L2 past the L1 in which block-scope visibility lives, and lost joins at full size.)
// ------------------------------------------------------------------------------------------
constexpr int kProbeThreads = 512;                               // consumer threads
constexpr int kProbeBlock = kProbeThreads + 32;                  // + one producer warp that only issues copies
constexpr int kProbeBlocksPerSm = 2;
constexpr int kProbeUnroll = 2;                                  // 16-byte pairs per thread per tile
constexpr int kProbeRows = 2 * kProbeUnroll;                     // rows per thread per tile
constexpr int kProbeBatch = kProbeThreads * kProbeUnroll;        // pairs per tile (16 KB)
constexpr int kProbeStages = 3;                                  // tiles in flight per block
constexpr int kProbeRingBytes = kProbeStages * kProbeBatch * 16;
constexpr int kResolveUnroll = 4;                                // candidates per thread in flight in the drain

// MODE 0: the fused kernel (stream, then resolve).  MODE 1 = k_stream: the stream alone, the tile's candidate count
// goes to cand_count.  MODE 2 = k_resolve: the resolve alone.
template <int MODE>
__device__ __forceinline__ void probe_body(const PhaseArgs &a) {
    extern __shared__ __align__(128) unsigned char s_raw[];      // [key ring | filter words]
    __shared__ __align__(8) unsigned long long s_full[kProbeStages], s_empty[kProbeStages], s_bm_full;
    __shared__ int s_count;
    dbg_mark(a, 1, 0);
    const ProbeTile tile = a.probe_tiles[blockIdx.x];            // one contig, one row range, everything needed
    ulonglong2 *ring = reinterpret_cast<ulonglong2 *>(s_raw);
    unsigned *s_bm = reinterpret_cast<unsigned *>(s_raw + kProbeRingBytes);
    const long long R = a.n_reads;
    const long long r0 = tile.r0, r1 = tile.r1;
    const unsigned bmw = (unsigned)tile.bmw;
    const int lane = threadIdx.x & 31;
    const ulonglong2 *pairs = reinterpret_cast<const ulonglong2 *>(a.read_key);
    const long long q0 = r0 >> 1;                                // pairs of rows (2q, 2q+1)
    const long long q1 = (r1 + 1) >> 1;
    const long long q_full = min(q1, R >> 1);                    // pairs that lie completely inside the column
    const int n_tiles = (int)((q1 - q0 + kProbeBatch - 1) / kProbeBatch);
    const unsigned long long stream_policy = (a.flags & kFlagNoStreamHint) ? 0ull : l2_evict_first_policy();
    auto tile_pairs = [&](int t) { return (unsigned)max(0ll, min(q_full, q0 + (long long)(t + 1) * kProbeBatch) - (q0 + (long long)t * kProbeBatch)); };
    if (MODE == 2) {                                             // k_resolve: everything before it has completed
        pdl_trigger();
        pdl_wait();
        if (threadIdx.x == 0) s_count = a.cand_count[blockIdx.x];
    } else {
    if (threadIdx.x == 0) {
        for (int s = 0; s < kProbeStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kProbeThreads / 32); }
        mbar_init(&s_bm_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_count = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                      // fill the ring: the stream starts before the filter is in
        for (int t = 0; t < min(n_tiles, kProbeStages); ++t) {
            const unsigned bytes = tile_pairs(t) * 16u;
            mbar_expect_tx(&s_full[t], bytes);
            if (bytes) bulk_load_stream(ring + (size_t)t * kProbeBatch, pairs + q0 + (long long)t * kProbeBatch, bytes, &s_full[t], stream_policy);
        }
    }
    pdl_trigger();
    pdl_wait();                                                  // k_table is done: filter bits and slots are final
    if (threadIdx.x == 0) {                                      // the contig's filter: one bulk copy
        const unsigned bytes = (bmw + 1u) * 4u;
        mbar_expect_tx(&s_bm_full, bytes);
        bulk_load(s_bm, a.bitmap + tile.bmo, bytes, &s_bm_full);
    }
    dbg_mark(a, 1, 1);
    if (threadIdx.x >= kProbeThreads) {
        // producer warp: refill a stage as soon as every consumer warp has handed it back.  The wait is
        // warp-uniform (all 32 lanes spin together), one lane issues the copy.
        for (int t = kProbeStages; t < n_tiles; ++t) {
            const int stage = t % kProbeStages;
            mbar_wait(&s_empty[stage], (unsigned)(t / kProbeStages - 1) & 1u);
            if (lane == 0) {
                const unsigned bytes = tile_pairs(t) * 16u;
                mbar_expect_tx(&s_full[stage], bytes);
                if (bytes) bulk_load_stream(ring + (size_t)stage * kProbeBatch, pairs + q0 + (long long)t * kProbeBatch, bytes, &s_full[stage], stream_policy);
            }
            __syncwarp();
        }
    } else {
        // consumer warps.  Rows are numbered inside the tile: local row lr <-> row 2*q0 + lr, valid in [lr0, lr1).
        const int row_base = (int)(2 * q0);
        const int lr0 = (int)(r0 - 2 * q0), lr1 = (int)(r1 - 2 * q0);
        ulonglong2 *out = a.cand_list + r0;                      // this block's private stretch of the list
        mbar_wait(&s_bm_full, 0);                                // the filter has landed
        dbg_mark(a, 1, 2);
        for (int t = 0; t < n_tiles; ++t) {
            const int stage = t % kProbeStages;
            const unsigned parity = (unsigned)(t / kProbeStages) & 1u;
            const long long qt = q0 + (long long)t * kProbeBatch;
            mbar_wait(&s_full[stage], parity);                   // the tile's bytes have landed
            unsigned long long key[kProbeRows];
#pragma unroll
            for (int u = 0; u < kProbeUnroll; ++u) {
                const long long qq = qt + (long long)u * kProbeThreads + threadIdx.x;
                ulonglong2 v = make_ulonglong2(0ull, 0ull);
                if (qq < q_full) v = ring[(size_t)stage * kProbeBatch + u * kProbeThreads + threadIdx.x];
                else if (qq < q1) v.x = __ldcs(a.read_key + 2 * qq);                   // the column's odd last row
                key[2 * u] = v.x; key[2 * u + 1] = v.y;
            }
            const int lr_t = 2 * (t * kProbeBatch + (int)threadIdx.x);                // local row of key[0]
            unsigned pass = 0;
#pragma unroll
            for (int u = 0; u < kProbeRows; ++u) {
                const int lr = lr_t + (u >> 1) * (2 * kProbeThreads) + (u & 1);
                const unsigned m = bloom_bits(key[u]);
                if (lr >= lr0 && lr < lr1 && (s_bm[bloom_word(key[u], bmw)] & m) == m) pass |= 1u << u;
            }
            // rows that passed the filter go to the candidate list: one shared-memory atomic per warp
            const int cnt = __popc(pass);
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int x = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += x;
            }
            // Hand the stage back only HERE: the scan has consumed every lane's filter result, so every lane's
            // ring loads have returned their data.  An arrive placed right after the loads is issued while they
            // are still in flight (nothing waits on their scoreboard) and the refill can overtake them
            // (measured in round 1: a warp read the tile three ahead, lost joins).
            if (lane == 0) mbar_arrive(&s_empty[stage]);
            int wbase = 0;
            if (lane == 31 && inc) wbase = atomicAdd(&s_count, inc);
            int pos = __shfl_sync(0xffffffffu, wbase, 31) + inc - cnt;
#pragma unroll
            for (int u = 0; u < kProbeRows; ++u)
                if (pass >> u & 1u) {
                    const int row = row_base + lr_t + (u >> 1) * (2 * kProbeThreads) + (u & 1);
                    out[pos++] = make_ulonglong2(key[u], (unsigned long long)(unsigned)row);
                    // what k_reduce will want from HBM at random -- the row's tag record -- starts moving into
                    // L2 now, under the stream
                    if (!(a.flags & kFlagNoTagPrefetch)) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.read_tag + row));
                }
        }
    }
    }                                                            // MODE != 2
    __syncthreads();
    dbg_mark(a, 1, 3);
    if (MODE == 1) {                                             // k_stream: k_resolve takes it from here
        if (threadIdx.x == 0) a.cand_count[blockIdx.x] = s_count;
        return;
    }
    // Resolve the block's candidates (all 17 warps): one 16-byte slot load decides; a hit pushes the row
    // index to every support-read entry of that name with atomicMax -- a later row overrides an earlier
    // one (sv_phasing_fn.py:29).
    // Dependent chain: candidate (L2, written by this block) -> slot (L2: k_init swept it) -> atomics.
    const int total = s_count;
    const ulonglong2 *q_cand = a.cand_list + r0;
    const Slot *tab = a.tab + tile.base;
    const unsigned mask = (unsigned)tile.mask;
    for (int g0 = threadIdx.x; g0 < total; g0 += kProbeBlock * kResolveUnroll) {
        unsigned long long key[kResolveUnroll];
        int row[kResolveUnroll];
        unsigned p[kResolveUnroll];
        unsigned pend = 0;
#pragma unroll
        for (int u = 0; u < kResolveUnroll; ++u) {
            const int g = g0 + u * kProbeBlock;
            key[u] = 0ull; row[u] = 0;
            if (g < total) { const ulonglong2 c = __ldcg(q_cand + g); key[u] = c.x; row[u] = (int)c.y; pend |= 1u << u; }
        }
#pragma unroll
        for (int u = 0; u < kResolveUnroll; ++u) p[u] = slot_hash(key[u]) & mask;
        while (pend) {                                           // lock step: one slot load per pending candidate
            uint4 sl[kResolveUnroll];
#pragma unroll
            for (int u = 0; u < kResolveUnroll; ++u)
                if (pend >> u & 1u) sl[u] = __ldcg(reinterpret_cast<const uint4 *>(tab + p[u]));     // key, first, head
#pragma unroll
            for (int u = 0; u < kResolveUnroll; ++u) {
                if (!(pend >> u & 1u)) continue;
                const unsigned long long k = ((unsigned long long)sl[u].y << 32) | sl[u].x;
                if (k == key[u]) {
                    atomicMax(a.join_row + (int)sl[u].z, row[u]);
                    for (int h = (int)sl[u].w; h >= 0; h = __ldcg(a.next + h)) atomicMax(a.join_row + h, row[u]);
                    pend &= ~(1u << u);
                } else if (k == kEmptyKey) {
                    pend &= ~(1u << u);                          // filter false positive
                } else {
                    p[u] = (p[u] + 1) & mask;
                }
            }
        }
    }
    dbg_mark(a, 1, 4);
}

__global__ void __launch_bounds__(kProbeBlock, kProbeBlocksPerSm)
k_probe(PhaseArgs a) { probe_body<0>(a); }
// two-branch mode (kFlagTwoBranch): the stream beside k_table, the resolve behind both
__global__ void __launch_bounds__(kProbeBlock, kProbeBlocksPerSm)
k_stream(PhaseArgs a) { probe_body<1>(a); }
__global__ void __launch_bounds__(kProbeBlock, kProbeBlocksPerSm)
k_resolve(PhaseArgs a) { probe_body<2>(a); }

// ------------------------------------------------------------------------------------------
// one-PS list of a shard, as a block of its own (k_oneps: contigs too large for k_tail's cluster).
// Thread t owns the contiguous chunk [t*per, (t+1)*per) of the shard's candidates; shards of up to
// kThreads*kStage SVs keep the chunk in registers so the list is read from L2 exactly once.
// ------------------------------------------------------------------------------------------
// Small shards (<= kSortSmemBytes / 8 SVs): the distinct candidates are collected in a shared-memory
// hash set (a contig has a few hundred phase sets however many SVs it has), and only those are sorted.
constexpr int kOnepsSmall = 4096;                                // SVs per shard handled by the hash-set path

__device__ void oneps_block_big(const PhaseArgs &a, int s, long long *smem_tile);

__device__ void oneps_block_small(const PhaseArgs &a, int s, long long *smem_tile) {
    constexpr int kSlots = kSortSmemBytes / 4;                   // 4096 ints
    constexpr int kPer = kSlots / kThreads;                      // 16 slots per thread
    __shared__ int s_has_min, s_overflow;
    int *tab = reinterpret_cast<int *>(smem_tile);
    const int b = (int)a.sv_off[s], n = (int)a.sv_off[s + 1] - b;
    for (int i = threadIdx.x; i < kSlots; i += kThreads) tab[i] = INT32_MIN;
    if (threadIdx.x == 0) { s_has_min = 0; s_overflow = 0; }
    __syncthreads();
    constexpr int kLoads = kOnepsSmall / kThreads;               // candidates per thread, requested together
    long long cv[kLoads];
#pragma unroll
    for (int u = 0; u < kLoads; ++u) {
        const int i = threadIdx.x + u * kThreads;
        cv[u] = i < n ? __ldcg(a.cand + b + i) : kNoCand;
    }
#pragma unroll
    for (int u = 0; u < kLoads; ++u) {
        if (cv[u] == kNoCand) continue;
        const int x = (int)cv[u];
        if (x == INT32_MIN) { atomicMax(&s_has_min, 1); continue; }   // the empty marker itself: tracked aside
        unsigned h = ((unsigned)x * 2654435761u) >> 20;
        int tries = 0;
        for (; tries < kSlots; ++tries) {
            const int prev = atomicCAS(tab + h, INT32_MIN, x);
            if (prev == INT32_MIN || prev == x) break;
            h = (h + 1) & (kSlots - 1);
        }
        if (tries == kSlots) s_overflow = 1;                     // more distinct phase sets than slots
    }
    __syncthreads();
    if (s_overflow) { __syncthreads(); oneps_block_big(a, s, smem_tile); return; }
    int found[kPer], cnt = 0;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
        found[u] = tab[threadIdx.x * kPer + u];
        cnt += found[u] != INT32_MIN;
    }
    int m;
    int w = block_scan_exclusive(cnt, 0, OpSum(), &m);           // ends with a barrier: the table has been read
#pragma unroll
    for (int u = 0; u < kPer; ++u)
        if (found[u] != INT32_MIN) tab[w++] = found[u];
    const int m_pad = next_pow2(max(m, 1));
    const int off = s_has_min;
    int *dst = a.oneps + b;
    __syncthreads();
    if (m <= 2 * kThreads) {
        // the values are distinct, so a value's rank IS its place in the sorted list: one pass over the
        // list in shared memory (four values per load) instead of a sorting network of barriers
        if (threadIdx.x < 4) tab[m + threadIdx.x] = INT32_MAX;
        __syncthreads();
        const int i0 = threadIdx.x, i1 = threadIdx.x + kThreads;
        const int v0 = i0 < m ? tab[i0] : 0, v1 = i1 < m ? tab[i1] : 0;
        int r0 = 0, r1 = 0;
        for (int j = 0; j < m; j += 4) {
            const int4 x = *reinterpret_cast<const int4 *>(tab + j);
            r0 += (x.x < v0) + (x.y < v0) + (x.z < v0) + (x.w < v0);
            r1 += (x.x < v1) + (x.y < v1) + (x.z < v1) + (x.w < v1);
        }
        if (i0 < m) dst[off + r0] = v0;
        if (i1 < m) dst[off + r1] = v1;
    } else {
        for (int i = m + threadIdx.x; i < m_pad; i += kThreads) tab[i] = INT32_MAX;
        __syncthreads();
        block_bitonic_sort(tab, m_pad);
        for (int i = threadIdx.x; i < m; i += kThreads) dst[off + i] = tab[i];
    }
    if (threadIdx.x == 0) {
        if (off) dst[0] = INT32_MIN;
        a.oneps_n[s] = m + off;
    }
    __syncthreads();
}

// Big shards (more SVs than the hash-set path holds): multi-pass over the candidates in L2 / global
// scratch.  Position-sorted candidates are compacted directly; otherwise runs of equal neighbours
// are squeezed first and what is left is sorted.
__device__ void oneps_block_big(const PhaseArgs &a, int s, long long *smem_tile) {
    const int b = (int)a.sv_off[s], n = (int)a.sv_off[s + 1] - b;
    const int per = (n + kThreads - 1) / kThreads;
    const int c0 = min(n, (int)threadIdx.x * per), c1 = min(n, c0 + per);
    const long long *cand = a.cand + b;
    long long mx = kNone, lastv = kNone;
    for (int i = c0; i < c1; ++i) { const long long v = __ldcg(cand + i); if (v != kNoCand) { mx = max(mx, v); lastv = v; } }
    const long long run = block_scan_exclusive(mx, kNone, OpMax(), (long long *)nullptr);
    long long cur = run;
    int cnt = 0;
    bool ok = true;
    for (int i = c0; i < c1; ++i) {
        const long long v = __ldcg(cand + i);
        if (v != kNoCand) { if (v < cur) ok = false; else if (v > cur) { ++cnt; cur = v; } }
    }
    if (__syncthreads_and(ok)) {                     // already non-decreasing in VCF order
        int total;
        int w = block_scan_exclusive(cnt, 0, OpSum(), &total);
        cur = run;
        for (int i = c0; i < c1; ++i) { const long long v = __ldcg(cand + i); if (v != kNoCand && v > cur) { a.oneps[b + w++] = (int)v; cur = v; } }
        if (threadIdx.x == 0) a.oneps_n[s] = total;
        return;
    }
    const long long prev0 = block_scan_exclusive(lastv, kNone, OpLast(), (long long *)nullptr);
    long long prev = prev0;
    cnt = 0;
    for (int i = c0; i < c1; ++i) { const long long v = __ldcg(cand + i); if (v != kNoCand) { cnt += v != prev; prev = v; } }
    int r;
    int w = block_scan_exclusive(cnt, 0, OpSum(), &r);
    long long *v = a.sort_scratch + 2ll * b;         // run heads: at most n of them
    prev = prev0;
    for (int i = c0; i < c1; ++i) { const long long x = __ldcg(cand + i); if (x != kNoCand) { if (x != prev) v[w++] = x; prev = x; } }
    __syncthreads();
    const int n_pad = next_pow2(max(r, 1));
    if (n_pad * (int)sizeof(long long) <= kSortSmemBytes) {      // few heads: sort them on chip
        for (int i = threadIdx.x; i < n_pad; i += kThreads) smem_tile[i] = i < r ? v[i] : kNoCand;
        v = smem_tile;
    } else {
        for (int i = r + threadIdx.x; i < n_pad; i += kThreads) v[i] = kNoCand;
    }
    __syncthreads();
    block_bitonic_sort(v, n_pad);
    const int per2 = (n_pad + kThreads - 1) / kThreads;
    const int d0 = min(n_pad, (int)threadIdx.x * per2), d1 = min(n_pad, d0 + per2);
    cnt = 0;
    for (int i = d0; i < d1; ++i) cnt += (v[i] != kNoCand && (i == 0 || v[i] != v[i - 1])) ? 1 : 0;
    int total;
    w = block_scan_exclusive(cnt, 0, OpSum(), &total);
    for (int i = d0; i < d1; ++i)
        if (v[i] != kNoCand && (i == 0 || v[i] != v[i - 1])) a.oneps[b + w++] = (int)v[i];
    if (threadIdx.x == 0) a.oneps_n[s] = total;
    __syncthreads();
}

__device__ __forceinline__ void oneps_any(const PhaseArgs &a, int s, long long *smem_tile) {
    if ((int)(a.sv_off[s + 1] - a.sv_off[s]) <= kOnepsSmall) oneps_block_small(a, s, smem_tile);
    else oneps_block_big(a, s, smem_tile);
}

// ------------------------------------------------------------------------------------------
// k_reduce: kReduceLanes lanes per SV.  A lane requests the join rows of up to kReduceUnroll of its
// support reads at once, then their (ps, pc, hp) together, then reduces: class = #distinct PS
// (:192-194), one-PS candidate = PS of the first read with pc <= 8100 (:195-203), class-1 counts and
// score sums (:74-84).  SVs that turn out to span several phase sets (class 2) also get their
// per-PS statistics here, while the tags are still in registers, in first-seen order (:85-105);
// k_predict only has to filter them by the one-PS set.
// Dependent chain: csr_off -> join_row -> tags -> (shuffles) -> stores -> credit.
// ------------------------------------------------------------------------------------------
constexpr int kReduceUnroll = 4;
constexpr int kHeavyReads = 256;                    // dense batches: longer support lists go to k_reduce_heavy

__device__ __forceinline__ ReadTag load_tag(const PhaseArgs &a, int row) {
    const int4 v = __ldg(reinterpret_cast<const int4 *>(a.read_tag + row));
    return ReadTag{v.x, v.y, (unsigned)v.z, (unsigned)v.w};
}

template <int CAP>
struct C2Group {                                   // shared-memory table of one lane group
    int ps[CAP], tot[CAP], n1[CAP], n2[CAP], bad[CAP];
    unsigned long long s1[CAP], s2[CAP];
};

// one step of the per-PS table: every lane of the group offers one read (q = qualifying)
template <int CAP>
__device__ __forceinline__ void c2_update(C2Group<CAP> &g, unsigned gmask, bool q, int ps, int pc, int hp, int &n_d) {
    const int wl = threadIdx.x & 31;
    int id = -1;
    const int known = min(n_d, CAP);
    for (int k = 0; k < known; ++k)
        if (q && g.ps[k] == ps) id = k;
    const bool fresh = q && id < 0;
    // lanes offering the same new PS; the lowest lane of each set speaks for it, in lane (= read) order
    const unsigned peers = __match_any_sync(gmask, fresh ? (long long)ps : (long long)0x4000000000000000ll + wl);
    const int leader = __ffs(peers) - 1;
    const unsigned leaders = __ballot_sync(gmask, fresh && leader == wl) & gmask;
    if (fresh) {
        id = n_d + __popc(leaders & ((1u << leader) - 1u));
        if (leader == wl && id < CAP) {
            g.ps[id] = ps; g.tot[id] = 0; g.n1[id] = 0; g.n2[id] = 0; g.bad[id] = 0; g.s1[id] = 0ull; g.s2[id] = 0ull;
        }
    }
    n_d += __popc(leaders);
    __syncwarp(gmask);
    if (q && id < CAP) {
        atomicAdd(&g.tot[id], 1);
        if (hp == 1) { atomicAdd(&g.n1[id], 1); atomicAdd(&g.s1[id], (unsigned long long)(long long)pc); }
        else if (hp == 2) { atomicAdd(&g.n2[id], 1); atomicAdd(&g.s2[id], (unsigned long long)(long long)pc); }
        else g.bad[id] = hp | 0x100;
    }
    __syncwarp(gmask);
}

template <int G>
__global__ void __launch_bounds__(kThreads, 6)
k_reduce(PhaseArgs a) {
    constexpr int kReducePerBlock = kThreads / G;
    dbg_mark(a, 2, 0);
    constexpr int kCap = c2_cap(G);
    __shared__ C2Group<kCap> s_c2[kReducePerBlock];
    const int lane = threadIdx.x % G, grp = threadIdx.x / G;
    const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) / G * G));
    const int sv0 = blockIdx.x * kReducePerBlock;
    const int sv1 = min(a.n_svs, sv0 + kReducePerBlock);
    const int sv = sv0 + grp;
    const bool live = sv < a.n_svs;
    long long b = 0, e = 0;
    bool kept = false;
    if (live) {
        b = __ldg(a.csr_off + sv); e = __ldg(a.csr_off + sv + 1);
        kept = __ldg(a.sv_svlen + sv) >= c_thr.svlen_thres &&                 // filter of :189-190
               __ldg(a.sv_svread + sv) >= c_thr.suppread_thres &&
               !(__ldg(a.sv_flags + sv) & DUET_SV_GT_MISSING);
    }
    // dense batches: a support list of a thousand reads would keep this one warp busy long after the rest of the grid
    // is done -- such SVs are left to k_reduce_heavy, eight warps each
    const bool skip = G == 32 && a.heavy_sv != nullptr && e - b > kHeavyReads;
    pdl_trigger();
    pdl_wait();                                                  // the join rows are final
    {   // every reader of the Bloom filter is done: hand it back zeroed, so that the next call's filter bits can be
        // set from its first microsecond on (two-branch mode: k_bloom beside k_init)
        uint4 *bm = reinterpret_cast<uint4 *>(a.bitmap);
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.n_bm_words / 4; i += (long long)gridDim.x * kThreads) bm[i] = zero;
    }
    if (skip) return;                                            // (G == 32: the whole warp)
    int hits = 0, ps_lo = INT32_MAX, ps_hi = INT32_MIN;
    int h1 = 0, h2 = 0, nq = 0;
    long long t1 = 0, t2 = 0;
    long long first_q = INT64_MAX;      // CSR index of the first read with pc <= pc_max
    int first_q_ps = 0;
    int row[kReduceUnroll], ps[kReduceUnroll], pc[kReduceUnroll], hp[kReduceUnroll];
    for (long long base = b; base < e; base += G * kReduceUnroll) {
        unsigned want[kReduceUnroll];
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) {
            const long long j = base + u * G + lane;
            row[u] = j < e ? __ldcg(a.join_row + j) : -1;
            want[u] = j < e && a.csr_chk ? __ldg(a.csr_chk + j) : 0u;
        }
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) {
            ps[u] = pc[u] = hp[u] = 0;
            if (row[u] >= 0) {                                   // one 16-byte record = one sector per joined read
                const ReadTag t = load_tag(a, row[u]);
                ps[u] = t.ps; pc[u] = t.pc; hp[u] = (int)(t.hp & 0xffu);
                if (a.csr_chk && t.chk != want[u])               // two different names share the 64-bit key
                    report(a.status, DUET_ERR_HASH_COLLISION, sv, (long long)__ldg(a.csr_key + base + u * G + lane));
            }
        }
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) {
            if (row[u] < 0) continue;
            ++hits;
            ps_lo = min(ps_lo, ps[u]);
            ps_hi = max(ps_hi, ps[u]);
            if (pc[u] <= c_thr.pc_max) {
                ++nq;
                const long long j = base + u * G + lane;
                if (j < first_q) { first_q = j; first_q_ps = ps[u]; }
                if (hp[u] == 1) { ++h1; t1 += pc[u]; }
                else if (hp[u] == 2) { ++h2; t2 += pc[u]; }
            }
        }
    }
    dbg_mark(a, 2, 1);
    hits = group_sum(hits, gmask, G);
    ps_lo = group_min(ps_lo, gmask, G);
    ps_hi = group_max(ps_hi, gmask, G);
    h1 = group_sum(h1, gmask, G); h2 = group_sum(h2, gmask, G); nq = group_sum(nq, gmask, G);
    t1 = group_sum(t1, gmask, G); t2 = group_sum(t2, gmask, G);
    const long long fq = group_min(first_q, gmask, G);
    const unsigned owner = __ballot_sync(gmask, first_q == fq && fq != INT64_MAX) & gmask;
    const int fps = __shfl_sync(gmask, first_q_ps, owner ? __ffs(owner) - 1 : (threadIdx.x & 31));
    const int cls = hits == 0 ? 0 : (ps_lo == ps_hi ? 1 : 2);
    if (live && lane == 0) {
        a.n_hit[sv] = hits;
        a.cls[sv] = kept ? (uint8_t)cls : (uint8_t)DUET_CLS_FILTERED;
        a.gt[sv] = 0;
        a.cand[sv] = (kept && cls == 1 && owner) ? (long long)fps : kNoCand;
        // class-1 view of the statistics (:74-84); k_predict overwrites them for class 2
        a.hap1[sv] = h1; a.hap2[sv] = h2; a.hap0[sv] = 0;
        a.allhap[sv] = cls == 2 ? nq : h1 + h2;
        a.totsc1[sv] = t1; a.totsc2[sv] = t2;
        a.ps[sv] = owner ? ps_lo : 0;   // class 1: every joined read carries the same PS
    }
    if (live && lane < DUET_N_FEATURES) a.features[(size_t)lane * a.n_svs + sv] = 0.0;

    if (live && kept && cls == 2) {                  // group-uniform: per-PS statistics in read order
        C2Group<kCap> &g = s_c2[grp];
        int n_d = 0;
        if (e - b <= G * kReduceUnroll) {            // the tags are still in registers
#pragma unroll
            for (int u = 0; u < kReduceUnroll; ++u)
                c2_update(g, gmask, row[u] >= 0 && pc[u] <= c_thr.pc_max, ps[u], pc[u], hp[u], n_d);
        } else {
            for (long long base = b; base < e; base += G * kReduceUnroll) {      // same batching as the first pass
#pragma unroll
                for (int u = 0; u < kReduceUnroll; ++u) {
                    const long long j = base + u * G + lane;
                    row[u] = j < e ? __ldcg(a.join_row + j) : -1;
                }
#pragma unroll
                for (int u = 0; u < kReduceUnroll; ++u) {
                    ps[u] = pc[u] = hp[u] = 0;
                    if (row[u] >= 0) { const ReadTag t = load_tag(a, row[u]); ps[u] = t.ps; pc[u] = t.pc; hp[u] = (int)(t.hp & 0xffu); }
                }
#pragma unroll
                for (int u = 0; u < kReduceUnroll; ++u)                          // u-major == read order
                    c2_update(g, gmask, row[u] >= 0 && pc[u] <= c_thr.pc_max, ps[u], pc[u], hp[u], n_d);
            }
        }
        C2Rec *rec = c2_at(a, sv);
        if (lane == 0) { rec->n_d = min(n_d, kCap); rec->overflow = n_d > kCap; }
        for (int t = lane; t < min(n_d, kCap); t += G)
            rec->d[t] = C2Ent{g.ps[t], g.tot[t], g.n1[t], g.n2[t], (long long)g.s1[t], (long long)g.s2[t], g.bad[t], 0};
    }

    dbg_mark(a, 2, 2);
}

// k_reduce_heavy (dense batches): ONE BLOCK per SV with more than kHeavyReads support reads.  The list is cut into
// eight contiguous stretches of whole 128-read batches, one per warp; every warp does what a k_reduce<32> warp does on
// its stretch -- the order-free statistics first, then (class 2) its own per-PS table in first-seen order --, and
// warp 0 puts the pieces together IN STRETCH ORDER, which is read order: the sums and extrema combine freely, the first
// qualifying read is the smallest index, and a phase set's place in the merged table is where the first stretch that saw
// it put it.  Same outputs as k_reduce, bit for bit.
struct HeavyPart { int hits, ps_lo, ps_hi, h1, h2, nq, first_q_ps, pad; long long t1, t2, first_q; };

__global__ void __launch_bounds__(kThreads)
k_reduce_heavy(PhaseArgs a) {
    constexpr int G = 32, W = kThreads / 32, kCap = c2_cap(32);
    __shared__ C2Group<kCap> s_c2[W];
    __shared__ C2Group<kCap> s_m;                                // the merged table
    __shared__ HeavyPart s_part[W];
    __shared__ int s_nd[W];                                      // distinct phase sets each warp saw (its own word: the pieces above are still being read)
    const int sv = a.heavy_sv[blockIdx.x];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned gmask = 0xffffffffu;
    const long long b = __ldg(a.csr_off + sv), e = __ldg(a.csr_off + sv + 1);
    const bool kept = __ldg(a.sv_svlen + sv) >= c_thr.svlen_thres && __ldg(a.sv_svread + sv) >= c_thr.suppread_thres &&
                      !(__ldg(a.sv_flags + sv) & DUET_SV_GT_MISSING);
    constexpr int kBatch = G * kReduceUnroll;
    const long long n_batches = (e - b + kBatch - 1) / kBatch, per = (n_batches + W - 1) / W;
    const long long wb = min(e, b + (long long)w * per * kBatch), we = min(e, wb + per * kBatch);
    pdl_trigger();
    pdl_wait();                                                  // the join rows are final
    int hits = 0, ps_lo = INT32_MAX, ps_hi = INT32_MIN, h1 = 0, h2 = 0, nq = 0, first_q_ps = 0;
    long long t1 = 0, t2 = 0, first_q = INT64_MAX;
    int row[kReduceUnroll], ps[kReduceUnroll], pc[kReduceUnroll], hp[kReduceUnroll];
    for (long long base = wb; base < we; base += kBatch) {
        unsigned want[kReduceUnroll];
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) {
            const long long j = base + u * G + lane;
            row[u] = j < we ? __ldcg(a.join_row + j) : -1;
            want[u] = j < we && a.csr_chk ? __ldg(a.csr_chk + j) : 0u;
        }
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) {
            ps[u] = pc[u] = hp[u] = 0;
            if (row[u] >= 0) {
                const ReadTag t = load_tag(a, row[u]);
                ps[u] = t.ps; pc[u] = t.pc; hp[u] = (int)(t.hp & 0xffu);
                if (a.csr_chk && t.chk != want[u])
                    report(a.status, DUET_ERR_HASH_COLLISION, sv, (long long)__ldg(a.csr_key + base + u * G + lane));
            }
        }
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) {
            if (row[u] < 0) continue;
            ++hits;
            ps_lo = min(ps_lo, ps[u]);
            ps_hi = max(ps_hi, ps[u]);
            if (pc[u] <= c_thr.pc_max) {
                ++nq;
                const long long j = base + u * G + lane;
                if (j < first_q) { first_q = j; first_q_ps = ps[u]; }
                if (hp[u] == 1) { ++h1; t1 += pc[u]; }
                else if (hp[u] == 2) { ++h2; t2 += pc[u]; }
            }
        }
    }
    hits = group_sum(hits, gmask, G);
    ps_lo = group_min(ps_lo, gmask, G);
    ps_hi = group_max(ps_hi, gmask, G);
    h1 = group_sum(h1, gmask, G); h2 = group_sum(h2, gmask, G); nq = group_sum(nq, gmask, G);
    t1 = group_sum(t1, gmask, G); t2 = group_sum(t2, gmask, G);
    {
        const long long fq = group_min(first_q, gmask, G);
        const unsigned owner = __ballot_sync(gmask, first_q == fq && fq != INT64_MAX);
        const int fps = __shfl_sync(gmask, first_q_ps, owner ? __ffs(owner) - 1 : 0);
        if (lane == 0) s_part[w] = HeavyPart{hits, ps_lo, ps_hi, h1, h2, nq, fps, 0, t1, t2, fq};
    }
    __syncthreads();
    // every thread puts the eight pieces together (cheap, and the class is needed by all)
    hits = 0; ps_lo = INT32_MAX; ps_hi = INT32_MIN; h1 = h2 = nq = 0; t1 = t2 = 0; first_q = INT64_MAX; first_q_ps = 0;
    for (int k = 0; k < W; ++k) {
        const HeavyPart p = s_part[k];
        hits += p.hits; ps_lo = min(ps_lo, p.ps_lo); ps_hi = max(ps_hi, p.ps_hi);
        h1 += p.h1; h2 += p.h2; nq += p.nq; t1 += p.t1; t2 += p.t2;
        if (p.first_q < first_q) { first_q = p.first_q; first_q_ps = p.first_q_ps; }
    }
    const bool owner = first_q != INT64_MAX;
    const int cls = hits == 0 ? 0 : (ps_lo == ps_hi ? 1 : 2);
    if (threadIdx.x == 0) {
        a.n_hit[sv] = hits;
        a.cls[sv] = kept ? (uint8_t)cls : (uint8_t)DUET_CLS_FILTERED;
        a.gt[sv] = 0;
        a.cand[sv] = (kept && cls == 1 && owner) ? (long long)first_q_ps : kNoCand;
        a.hap1[sv] = h1; a.hap2[sv] = h2; a.hap0[sv] = 0;
        a.allhap[sv] = cls == 2 ? nq : h1 + h2;
        a.totsc1[sv] = t1; a.totsc2[sv] = t2;
        a.ps[sv] = owner ? ps_lo : 0;
    }
    if (threadIdx.x < DUET_N_FEATURES) a.features[(size_t)threadIdx.x * a.n_svs + sv] = 0.0;
    if (!(kept && cls == 2)) return;                             // block-uniform

    // per-PS statistics: every warp its own table over its stretch, in read order
    C2Group<kCap> &g = s_c2[w];
    int n_d = 0;
    for (long long base = wb; base < we; base += kBatch) {
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) {
            const long long j = base + u * G + lane;
            row[u] = j < we ? __ldcg(a.join_row + j) : -1;
        }
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u) {
            ps[u] = pc[u] = hp[u] = 0;
            if (row[u] >= 0) { const ReadTag t = load_tag(a, row[u]); ps[u] = t.ps; pc[u] = t.pc; hp[u] = (int)(t.hp & 0xffu); }
        }
#pragma unroll
        for (int u = 0; u < kReduceUnroll; ++u)                  // u-major == read order
            c2_update(g, gmask, row[u] >= 0 && pc[u] <= c_thr.pc_max, ps[u], pc[u], hp[u], n_d);
    }
    if (lane == 0) s_nd[w] = n_d;
    __syncthreads();
    if (w != 0) return;
    // warp 0: the tables one after the other; lane t looks after merged entry t
    int m = 0;
    bool over = false;
    for (int k = 0; k < W; ++k) {
        const int nk = s_nd[k];
        over |= nk > kCap;
        for (int t = 0; t < min(nk, kCap); ++t) {
            const int psv = s_c2[k].ps[t];
            const unsigned hit = __ballot_sync(gmask, lane < m && s_m.ps[lane] == psv);
            int id = hit ? __ffs(hit) - 1 : m;
            if (!hit) {
                if (m == kCap) { over = true; continue; }
                if (lane == 0) { s_m.ps[m] = psv; s_m.tot[m] = 0; s_m.n1[m] = 0; s_m.n2[m] = 0; s_m.bad[m] = 0; s_m.s1[m] = 0ull; s_m.s2[m] = 0ull; }
                ++m;
            }
            if (lane == 0) {
                s_m.tot[id] += s_c2[k].tot[t]; s_m.n1[id] += s_c2[k].n1[t]; s_m.n2[id] += s_c2[k].n2[t];
                s_m.s1[id] += s_c2[k].s1[t]; s_m.s2[id] += s_c2[k].s2[t];
                if (!s_m.bad[id]) s_m.bad[id] = s_c2[k].bad[t];
            }
            __syncwarp();
        }
    }
    C2Rec *rec = c2_at(a, sv);
    if (lane == 0) { rec->n_d = m; rec->overflow = over; }
    if (lane < m)
        rec->d[lane] = C2Ent{s_m.ps[lane], s_m.tot[lane], s_m.n1[lane], s_m.n2[lane], (long long)s_m.s1[lane], (long long)s_m.s2[lane], s_m.bad[lane], 0};
}

// k_oneps: one block per shard -- the contig's sorted unique one-PS list (:107) from the candidates
// k_reduce left behind.  Its blocks are resident before k_reduce ends (programmatic launch) and start the
// moment it has.
__global__ void __launch_bounds__(kThreads)
k_oneps(PhaseArgs a) {
    __shared__ __align__(16) long long s_tile[kSortSmemBytes / 8];
    pdl_trigger();
    pdl_wait();                                                  // every SV of the shard has been reduced
    const int s = blockIdx.x;
    if (a.sv_off[s + 1] > a.sv_off[s]) oneps_any(a, s, s_tile);
}

// ------------------------------------------------------------------------------------------
// k_predict helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool in_sorted(const int *v, int n, int x) {       // v: global or shared
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int m = v[mid];
        if (m < x) lo = mid + 1; else hi = mid;
    }
    return lo < n && v[lo] == x;
}

// :107-111 -- nearest element of the sorted one-PS list to pos, an exact tie goes up
__device__ __forceinline__ int nearest_ps(const int *v, int n, int pos) {
    int lo = 0, hi = n;               // searchsorted(side='left')
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (v[mid] < pos) lo = mid + 1; else hi = mid;
    }
    const int below = max(lo - 1, 0), above = min(lo, n - 1);
    const int vb = v[below], va = v[above];
    const long long db = llabs((long long)pos - vb);
    const long long da = llabs((long long)pos - va);
    return db < da ? vb : va;
}

struct Class2Stats { int h1, h2, hap0, allhap, ps; long long t1, t2; };

struct Entry { bool q; bool in; int hp, ps, pc; };

__device__ __forceinline__ Entry load_entry(const PhaseArgs &a, long long j, long long e,
                                            const int *__restrict__ oneps, int n_one) {
    Entry r{false, false, 0, 0, 0};
    if (j < e) {
        const int row = __ldcg(a.join_row + j);
        if (row >= 0) {
            const ReadTag t = load_tag(a, row);
            r.pc = t.pc;
            if (r.pc <= c_thr.pc_max) {
                r.q = true;
                r.ps = t.ps;
                r.hp = (int)(t.hp & 0xffu);
                r.in = in_sorted(oneps, n_one, r.ps);
            }
        }
    }
    return r;
}

// exact but quadratic path for SVs whose reads span more than kMaxDistinct in-set phase sets
__device__ void class2_slow(const PhaseArgs &a, long long b, long long e, const int *oneps, int n_one,
                            Class2Stats &st) {
    const int lane = threadIdx.x & 31;
    int best = 0;
    for (long long j = b; j < e; ++j) {
        const Entry cur = load_entry(a, j, e, oneps, n_one);    // warp-uniform
        if (!cur.in) continue;
        bool seen = false;
        for (long long i = b + lane; i < j && !seen; i += 32) {
            const Entry x = load_entry(a, i, e, oneps, n_one);
            seen = x.in && x.ps == cur.ps;
        }
        if (__any_sync(0xffffffffu, seen)) continue;            // not the first occurrence
        int tot = 0, n1 = 0, n2 = 0;
        long long s1 = 0, s2 = 0;
        for (long long i = j + lane; i < e; i += 32) {
            const Entry x = load_entry(a, i, e, oneps, n_one);
            if (x.in && x.ps == cur.ps) {
                ++tot;
                if (x.hp == 1) { ++n1; s1 += x.pc; } else if (x.hp == 2) { ++n2; s2 += x.pc; }
            }
        }
        tot = warp_sum(tot); n1 = warp_sum(n1); n2 = warp_sum(n2);
        s1 = warp_sum(s1); s2 = warp_sum(s2);
        if (tot > best) {
            best = tot;
            st.h1 = n1; st.h2 = n2; st.t1 = s1; st.t2 = s2; st.ps = cur.ps;
            st.hap0 = st.allhap - n1 - n2;
        }
    }
}

struct Class2Smem {
    int ps[kMaxDistinct];
    int cnt[kMaxDistinct][3];                  // tot, n1, n2
    unsigned long long sc[kMaxDistinct][2];
    int bad[kMaxDistinct];                     // an HP outside {1,2} was seen with this PS
};

// all 32 lanes: per-PS statistics of one class-2 SV (:85-105); result is warp-uniform.
// The distinct PS values of the qualifying reads are collected in first-seen order; membership in
// the one-PS set is then tested once per distinct value (not once per read) -- the in-set values
// keep their relative first-seen order, which is what the reference's dict iteration sees.
__device__ void class2_stats(const PhaseArgs &a, int sv, long long b, long long e, const int *oneps, int n_one,
                             Class2Smem &m, Class2Stats &st) {
    const int lane = threadIdx.x & 31;
    st.h1 = st.h2 = st.hap0 = 0; st.t1 = st.t2 = 0; st.ps = 0;
    int n_d = 0;
    bool overflow = false;
    constexpr int kBatch = 4;                                    // reads per lane requested together
    int b_row[kBatch], b_ps[kBatch], b_pc[kBatch], b_hp[kBatch];
    for (long long base = b; base < e && !overflow; base += 32) {
        const int u = (int)((base - b) / 32) % kBatch;
        if (u == 0) {                                            // refill: kBatch x 32 reads, two round trips
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const long long j = base + k * 32 + lane;
                b_row[k] = j < e ? __ldcg(a.join_row + j) : -1;
            }
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                b_ps[k] = b_pc[k] = b_hp[k] = 0;
                if (b_row[k] >= 0) { const ReadTag t = load_tag(a, b_row[k]); b_ps[k] = t.ps; b_pc[k] = t.pc; b_hp[k] = (int)(t.hp & 0xffu); }
            }
        }
        int row = -1, ps = 0, pc = 0, hp = 0;
#pragma unroll
        for (int k = 0; k < kBatch; ++k)
            if (k == u) { row = b_row[k]; ps = b_ps[k]; pc = b_pc[k]; hp = b_hp[k]; }
        const bool q = row >= 0 && pc <= c_thr.pc_max;
        int id = -1;
        for (int t = 0; t < n_d; ++t)
            if (q && m.ps[t] == ps) id = t;
        unsigned fresh = __ballot_sync(0xffffffffu, q && id < 0);
        while (fresh) {                                          // lane order = first-seen order
            const int l0 = __ffs(fresh) - 1;
            const int v = __shfl_sync(0xffffffffu, ps, l0);
            const bool mine = q && id < 0 && ps == v;
            if (n_d == kMaxDistinct) { overflow = true; break; }
            if (lane == 0) {
                m.ps[n_d] = v;
                m.cnt[n_d][0] = m.cnt[n_d][1] = m.cnt[n_d][2] = 0;
                m.sc[n_d][0] = m.sc[n_d][1] = 0ull;
                m.bad[n_d] = 0;
            }
            if (mine) id = n_d;
            ++n_d;
            fresh &= ~__ballot_sync(0xffffffffu, mine);
        }
        __syncwarp();
        if (!overflow && q && id >= 0) {
            atomicAdd(&m.cnt[id][0], 1);
            if (hp == 1 || hp == 2) {
                atomicAdd(&m.cnt[id][hp], 1);
                atomicAdd(&m.sc[id][hp - 1], (unsigned long long)(long long)pc);
            } else {
                m.bad[id] = hp | 0x100;
            }
        }
        __syncwarp();
    }
    if (overflow) {
        class2_slow(a, b, e, oneps, n_one, st);
        for (long long j = b + lane; j < e; j += 32) {           // the KeyError of :96
            const Entry x = load_entry(a, j, e, oneps, n_one);
            if (x.in && x.hp != 1 && x.hp != 2) report(a.status, DUET_ERR_BAD_HP, sv, x.hp);
        }
    } else {
        const bool in = lane < n_d && in_sorted(oneps, n_one, m.ps[lane]);
        const unsigned in_mask = __ballot_sync(0xffffffffu, in);
        if (in && m.bad[lane]) report(a.status, DUET_ERR_BAD_HP, sv, m.bad[lane] & 0xff);
        int best = 0;
        for (int t = 0; t < n_d; ++t) {                          // strict '>' keeps the first seen (:101)
            if ((in_mask >> t & 1u) && m.cnt[t][0] > best) {
                best = m.cnt[t][0];
                st.h1 = m.cnt[t][1]; st.h2 = m.cnt[t][2];
                st.t1 = (long long)m.sc[t][0]; st.t2 = (long long)m.sc[t][1];
                st.ps = m.ps[t];
                st.hap0 = st.allhap - st.h1 - st.h2;
            }
        }
    }
    __syncwarp();
}

// features (:112-132) and the T1-T5 tree (:142-183) of one SV; one thread
struct Decision {
    int pred;                                  // 0 dropped, 1 "1|0", 2 "0|1", 3 "1|1"
    bool valid;                                // false: the reference raises here (reported), nothing is stored
    Class2Stats st;                            // with the phase set the row carries
    double f[DUET_N_FEATURES];
};

__device__ __forceinline__ Decision decide(const PhaseArgs &a, int sv, int cls, Class2Stats st, const int *oneps, int n_one,
                                           int n_list) {
    Decision d;
    d.pred = 0; d.valid = false;
    const int pos = __ldg(a.sv_pos + sv);
    if (cls == 0 || (st.h1 == 0 && st.h2 == 0)) st.ps = nearest_ps(oneps, n_one, pos);     // :106-111
    d.st = st;
#pragma unroll
    for (int k = 0; k < DUET_N_FEATURES; ++k) d.f[k] = 0.0;
    const int svread = __ldg(a.sv_svread + sv), refread = __ldg(a.sv_refread + sv);
    if ((long long)svread + refread == 0 || n_list == 0) {
        report(a.status, DUET_ERR_ZERO_DIVISION, sv, 0);
        return d;
    }
    d.valid = true;
    // Python int/int true division == correctly rounded fp64 division of the exact operands
    const double hapread_ratio = (double)st.allhap / (double)n_list;
    const double a1 = st.h1 > 0 ? (double)st.t1 / (double)st.h1 : 0.0;
    const double a2 = st.h2 > 0 ? (double)st.t2 / (double)st.h2 : 0.0;
    const double sv_ratio = (double)svread / (double)((long long)svread + refread);
    const long long tmin = min(st.t1, st.t2), tmax = max(st.t1, st.t2);
    const double totsc_ratio = tmin > 0 ? (double)tmax / (double)tmin : 0.0;
    const long long onehap_totsc = tmin == 0 ? tmax : 0;
    const double avgsc_diff = fabs(a2 - a1);

    int pred = 0;
    if (cls == 0) {                                                                          // :145-147
        if (sv_ratio == 1.0 && svread >= c_thr.c0_sv_num_min) pred = 3;
    } else if (cls == 2) {                                                                   // :148-155
        if (sv_ratio >= c_thr.c2_sv_ratio_min) {
            if (avgsc_diff <= c_thr.c2_avgsc_diff_max) { if (svread >= c_thr.c2_sv_num_min) pred = 3; }
            else if (st.hap0 >= c_thr.c2_hap0_min) pred = 3;
        }
    } else {                                                                                 // :156-182
        if (onehap_totsc != 0) {
            const bool agree = (hapread_ratio <= c_thr.c1_hapread_ratio && avgsc_diff <= c_thr.c1_avgsc_diff_max) ||
                               hapread_ratio > c_thr.c1_hapread_ratio;
            if (sv_ratio <= c_thr.c1_one_ratio_lo) pred = 0;
            else if (sv_ratio <= c_thr.c1_one_ratio_hi) { if (agree) pred = a1 > 0.0 ? 1 : 2; }
            else if (agree) pred = 3;
        } else {
            const int stronger = st.t1 > st.t2 ? 1 : 2;
            if (sv_ratio <= c_thr.c1_two_ratio_a) pred = 0;
            else if (sv_ratio <= c_thr.c1_two_ratio_b) pred = refread > c_thr.c1_ref_num_max ? 0 : stronger;
            else if (sv_ratio <= c_thr.c1_two_ratio_c) pred = totsc_ratio <= c_thr.c1_totsc_ratio_max ? 3 : stronger;
            else pred = 3;
        }
    }
    d.pred = pred;
    d.f[0] = hapread_ratio; d.f[1] = sv_ratio; d.f[2] = a1; d.f[3] = a2; d.f[4] = totsc_ratio; d.f[5] = avgsc_diff;
    return d;
}

__device__ __forceinline__ void store_decision(const PhaseArgs &a, int sv, const Decision &d) {
    if (!d.valid) return;
    a.gt[sv] = (uint8_t)d.pred;
    a.ps[sv] = d.st.ps;
    a.hap1[sv] = d.st.h1; a.hap2[sv] = d.st.h2; a.hap0[sv] = d.st.hap0; a.allhap[sv] = d.st.allhap;
    a.totsc1[sv] = d.st.t1; a.totsc2[sv] = d.st.t2;
    const size_t S = (size_t)a.n_svs;
#pragma unroll
    for (int k = 0; k < DUET_N_FEATURES; ++k) a.features[k * S + sv] = d.f[k];
}

__device__ int decide_and_store(const PhaseArgs &a, int sv, int cls, Class2Stats st, const int *oneps, int n_one,
                                int n_list) {
    const Decision d = decide(a, sv, cls, st, oneps, n_one, n_list);
    store_decision(a, sv, d);
    return d.pred;
}

// ------------------------------------------------------------------------------------------
// emission order + counters of a shard (k_order, and k_tail's block 0 when the VCF was not sorted)
// ------------------------------------------------------------------------------------------
typedef unsigned __int128 u128;

__device__ __forceinline__ long long order_key(const PhaseArgs &a, int sv, int cls) {
    const unsigned long long grp = a.sv_group ? (unsigned long long)(unsigned)__ldg(a.sv_group + sv) : 0ull;
    const unsigned long long upos = (unsigned long long)((unsigned)__ldg(a.sv_pos + sv) ^ 0x80000000u);
    return (long long)((grp << 34) | (upos << 2) | (unsigned long long)cls);      // < 2^50
}


// shards of up to kThreads*kOrdStage SVs: every thread fetches its chunk's per-SV state with
// independent loads (one round trip), everything after that runs out of registers
template <int kOrdStage>
__device__ void order_block_small(const PhaseArgs &a, int s, long long *smem_tile) {
    __shared__ unsigned long long s_cnt[DUET_N_COUNTERS];
    const int b = (int)a.sv_off[s], n = (int)a.sv_off[s + 1] - b;
    if (threadIdx.x < DUET_N_COUNTERS) s_cnt[threadIdx.x] = 0ull;
    __syncthreads();
    const int per = (n + kThreads - 1) / kThreads;                       // <= kOrdStage
    const int c0 = min(n, (int)threadIdx.x * per), c1 = min(n, c0 + per);
    int g[kOrdStage], cl[kOrdStage], nh[kOrdStage], pos[kOrdStage], grp[kOrdStage];
#pragma unroll
    for (int u = 0; u < kOrdStage; ++u) {
        const bool live = c0 + u < c1;
        const int sv = b + c0 + u;
        g[u] = live ? (int)__ldcg(a.gt + sv) : 0;
        cl[u] = live ? (int)__ldcg(a.cls + sv) : DUET_CLS_FILTERED;
        nh[u] = live ? __ldcg(a.n_hit + sv) : 0;
        pos[u] = live ? __ldg(a.sv_pos + sv) : 0;
        grp[u] = live && a.sv_group ? __ldg(a.sv_group + sv) : 0;
    }
    unsigned long long c_kept = 0, c_emit = 0, c10 = 0, c01 = 0, c11 = 0, c_hits = 0;
    long long key[kOrdStage];
    long long mx = kNone;
#pragma unroll
    for (int u = 0; u < kOrdStage; ++u) {
        c_hits += (unsigned long long)nh[u];
        c_kept += (c0 + u < c1) && cl[u] != DUET_CLS_FILTERED;
        key[u] = kNone;
        if (g[u] != 0) {
            ++c_emit; c10 += g[u] == 1; c01 += g[u] == 2; c11 += g[u] == 3;
            key[u] = (long long)(((unsigned long long)(unsigned)grp[u] << 34) |
                                 ((unsigned long long)((unsigned)pos[u] ^ 0x80000000u) << 2) | (unsigned long long)cl[u]);
            mx = max(mx, key[u]);
        }
    }
    c_kept = warp_sum(c_kept); c_emit = warp_sum(c_emit); c10 = warp_sum(c10);
    c01 = warp_sum(c01); c11 = warp_sum(c11); c_hits = warp_sum(c_hits);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_cnt[1], c_kept); atomicAdd(&s_cnt[2], c_emit); atomicAdd(&s_cnt[3], c10);
        atomicAdd(&s_cnt[4], c01); atomicAdd(&s_cnt[5], c11); atomicAdd(&s_cnt[7], c_hits);
    }
    const long long run = block_scan_exclusive(mx, kNone, OpMax(), (long long *)nullptr);
    long long cur = run;
    int cnt = 0;
    bool ok = true;
#pragma unroll
    for (int u = 0; u < kOrdStage; ++u)
        if (key[u] != kNone) { if (key[u] < cur) ok = false; cur = max(cur, key[u]); ++cnt; }
    const bool sorted = __syncthreads_and(ok);
    const int n_emit = (int)s_cnt[2];
    int w = block_scan_exclusive(cnt, 0, OpSum(), (int *)nullptr);
    if (sorted) {                                    // VCF already in (group, pos, class) order: just compact
#pragma unroll
        for (int u = 0; u < kOrdStage; ++u)
            if (key[u] != kNone) a.order[b + w++] = b + c0 + u;
    } else {                                         // sort (key, index in shard) packed in 62 bits
        const int n_pad = next_pow2(max(n_emit, 1));
        unsigned long long *v = n_pad * 8 <= kSortSmemBytes ? reinterpret_cast<unsigned long long *>(smem_tile)
                                                            : reinterpret_cast<unsigned long long *>(a.sort_scratch) + 4ll * b;
#pragma unroll
        for (int u = 0; u < kOrdStage; ++u)
            if (key[u] != kNone) v[w++] = ((unsigned long long)key[u] << 12) | (unsigned)(c0 + u);
        for (int i = n_emit + threadIdx.x; i < n_pad; i += kThreads) v[i] = ~0ull;
        __syncthreads();
        block_bitonic_sort(v, n_pad);
        for (int i = threadIdx.x; i < n_emit; i += kThreads) a.order[b + i] = b + (int)(v[i] & 0xFFFull);
    }
    if (threadIdx.x == 0) {
        a.n_emit[s] = n_emit;
        long long *c = a.shard_counts + (size_t)s * DUET_N_COUNTERS;
        c[0] = n;
        c[1] = (long long)s_cnt[1]; c[2] = (long long)s_cnt[2]; c[3] = (long long)s_cnt[3];
        c[4] = (long long)s_cnt[4]; c[5] = (long long)s_cnt[5];
        c[6] = a.csr_off[b + n] - a.csr_off[b];
        c[7] = (long long)s_cnt[7];
    }
    __syncthreads();
}

__device__ void order_block_big(const PhaseArgs &a, int s, long long *smem_tile) {
    __shared__ unsigned long long s_cnt[DUET_N_COUNTERS];
    const int b = (int)a.sv_off[s], n = (int)a.sv_off[s + 1] - b;
    if (threadIdx.x < DUET_N_COUNTERS) s_cnt[threadIdx.x] = 0ull;
    __syncthreads();
    const int per = (n + kThreads - 1) / kThreads;
    const int c0 = min(n, (int)threadIdx.x * per), c1 = min(n, c0 + per);
    unsigned long long c_kept = 0, c_emit = 0, c10 = 0, c01 = 0, c11 = 0, c_hits = 0;
    long long mx = kNone;
    for (int i = c0; i < c1; ++i) {
        const int sv = b + i;
        const int g = __ldcg(a.gt + sv), cls = __ldcg(a.cls + sv);
        c_hits += (unsigned long long)__ldcg(a.n_hit + sv);
        c_kept += cls != DUET_CLS_FILTERED;
        if (g != 0) {
            ++c_emit; c10 += g == 1; c01 += g == 2; c11 += g == 3;
            mx = max(mx, order_key(a, sv, cls));
        }
    }
    c_kept = warp_sum(c_kept); c_emit = warp_sum(c_emit); c10 = warp_sum(c10);
    c01 = warp_sum(c01); c11 = warp_sum(c11); c_hits = warp_sum(c_hits);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_cnt[1], c_kept); atomicAdd(&s_cnt[2], c_emit); atomicAdd(&s_cnt[3], c10);
        atomicAdd(&s_cnt[4], c01); atomicAdd(&s_cnt[5], c11); atomicAdd(&s_cnt[7], c_hits);
    }
    // fast path: emitted SVs already sorted by (group, pos, class) in VCF order
    const long long run = block_scan_exclusive(mx, kNone, OpMax(), (long long *)nullptr);
    long long cur = run;
    int cnt = 0;
    bool ok = true;
    for (int i = c0; i < c1; ++i) {
        const int sv = b + i;
        if (__ldcg(a.gt + sv) == 0) continue;
        const long long k = order_key(a, sv, __ldcg(a.cls + sv));
        if (k < cur) ok = false;
        cur = max(cur, k);
        ++cnt;
    }
    const bool sorted = __syncthreads_and(ok);
    const int n_emit = (int)s_cnt[2];
    if (sorted) {
        int w = block_scan_exclusive(cnt, 0, OpSum(), (int *)nullptr);
        for (int i = c0; i < c1; ++i)
            if (__ldcg(a.gt + b + i) != 0) a.order[b + w++] = b + i;
    } else {
        const u128 kPad = ~(u128)0;
        const int n_pad = next_pow2(max(n, 1));
        u128 *v = n_pad * (int)sizeof(u128) <= kSortSmemBytes ? reinterpret_cast<u128 *>(smem_tile)
                                                             : reinterpret_cast<u128 *>(a.sort_scratch) + 2ll * b;
        for (int i = threadIdx.x; i < n_pad; i += kThreads) {
            u128 key = kPad;
            if (i < n && __ldcg(a.gt + b + i) != 0)
                key = ((u128)(unsigned long long)order_key(a, b + i, __ldcg(a.cls + b + i)) << 64) | (u128)(unsigned)i;
            v[i] = key;
        }
        __syncthreads();
        block_bitonic_sort(v, n_pad);
        for (int i = threadIdx.x; i < n_emit; i += kThreads) a.order[b + i] = b + (int)(unsigned)v[i];
    }
    if (threadIdx.x == 0) {
        a.n_emit[s] = n_emit;
        long long *c = a.shard_counts + (size_t)s * DUET_N_COUNTERS;
        c[0] = n;
        c[1] = (long long)s_cnt[1]; c[2] = (long long)s_cnt[2]; c[3] = (long long)s_cnt[3];
        c[4] = (long long)s_cnt[4]; c[5] = (long long)s_cnt[5];
        c[6] = a.csr_off[b + n] - a.csr_off[b];
        c[7] = (long long)s_cnt[7];
    }
    __syncthreads();
}

__device__ __forceinline__ void order_block(const PhaseArgs &a, int s, long long *smem_tile) {
    // register-staged paths (the packed sort key keeps 12 bits for the index in the shard)
    const int n = (int)(a.sv_off[s + 1] - a.sv_off[s]);
    if (n <= kThreads * 8) order_block_small<8>(a, s, smem_tile);
    else if (n <= kThreads * 16) order_block_small<16>(a, s, smem_tile);
    else order_block_big(a, s, smem_tile);
}

// ------------------------------------------------------------------------------------------
// k_predict: one thread per SV (threads 0..63 of a block; all 256 help with the rest).
//   * the one-PS list of the block's first contig is staged in shared memory (binary searches stay
//     on chip); class-2 SVs read the per-PS statistics k_reduce recorded and keep the first-seen
//     in-set phase set with the most reads (:99-105); then features and the T1-T5 tree;
//   * SVs whose reads span more than kC2Max phase sets fall back to a warp-cooperative exact path;
// Dependent chain: [tile, per-SV state] -> [oneps_n, staged list, class-2 record] -> decide -> stores.
// ------------------------------------------------------------------------------------------
constexpr int kOneSmem = 2048;

__global__ void __launch_bounds__(kThreads, 4)
k_predict(PhaseArgs a) {
    __shared__ Class2Smem s_c2[kThreads / 32];
    __shared__ int s_one[kOneSmem];
    __shared__ int s_fb[kPredictPerBlock];
    __shared__ int s_nfb, s_n_one;
    dbg_mark(a, 3, 0);
    const PredictTile tile = a.predict_tiles[blockIdx.x];        // SVs [sv0, sv1) of ONE shard
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int blk0 = tile.sv0, blk1 = tile.sv1, shard = tile.shard;
    const int sv = blk0 + threadIdx.x;
    const bool mine = threadIdx.x < kPredictPerBlock && sv < blk1;
    int cls = DUET_CLS_FILTERED, n_list = 0;
    Class2Stats st{0, 0, 0, 0, 0, 0, 0};
    if (mine) n_list = (int)(__ldg(a.csr_off + sv + 1) - __ldg(a.csr_off + sv));     // input: requested before the wait
    pdl_trigger();
    pdl_wait();                                                  // k_reduce's per-SV results and one-PS lists are final
    if (mine) {                                                  // everything that does not depend on the list
        cls = a.cls[sv];
        st = Class2Stats{a.hap1[sv], a.hap2[sv], 0, a.allhap[sv], a.ps[sv], a.totsc1[sv], a.totsc2[sv]};
    }
    if (threadIdx.x == 0) s_nfb = 0;
    {   // the shard's sorted unique one-PS list (:107), built by the k_reduce block that completed the shard
        const int n_stage = min(tile.n, kOneSmem);
        for (int i = threadIdx.x; i < n_stage; i += kThreads) s_one[i] = a.oneps[tile.b + i];   // only [0, n) is meaningful
        if (threadIdx.x == 0) s_n_one = a.oneps_n[shard];
        __syncthreads();
    }
    const int n_one = s_n_one;                                   // 0: contig skipped (:209-210)
    const int *one = n_one <= kOneSmem ? s_one : a.oneps + tile.b;
    dbg_mark(a, 3, 1);

    if (mine && cls != DUET_CLS_FILTERED) {
        if (n_one > 0) {
            bool ready = true;
            if (cls == 0) st = Class2Stats{0, 0, 0, 0, 0, 0, 0};     // get_phase_info skips both loops
            if (cls == 2) {
                const C2Rec *rec = c2_at(a, sv);
                const int allhap = st.allhap;
                st = Class2Stats{0, 0, 0, allhap, 0, 0, 0};
                if (rec->overflow) {
                    s_fb[atomicAdd(&s_nfb, 1)] = threadIdx.x;
                    ready = false;
                } else {
                    const int n_d = rec->n_d;
                    int best = 0;
                    for (int t = 0; t < n_d; ++t) {                  // first-seen order; strict '>' (:101)
                        const C2Ent ent = rec->d[t];
                        if (!in_sorted(one, n_one, ent.ps)) continue;
                        if (ent.bad) report(a.status, DUET_ERR_BAD_HP, sv, ent.bad & 0xff);
                        if (ent.tot > best) {
                            best = ent.tot;
                            st.h1 = ent.n1; st.h2 = ent.n2; st.t1 = ent.s1; st.t2 = ent.s2; st.ps = ent.ps;
                            st.hap0 = allhap - ent.n1 - ent.n2;
                        }
                    }
                }
            }
            if (ready) decide_and_store(a, sv, cls, st, one, n_one, n_list);
        }
    }
    dbg_mark(a, 3, 2);
    __syncthreads();
    for (int k = w; k < s_nfb; k += kThreads / 32) {                 // rare: > kC2Max phase sets in one SV
        const int sv2 = blk0 + s_fb[k];
        const long long b2 = __ldg(a.csr_off + sv2), e2 = __ldg(a.csr_off + sv2 + 1);
        Class2Stats t{0, 0, 0, a.allhap[sv2], 0, 0, 0};
        class2_stats(a, sv2, b2, e2, one, n_one, s_c2[w], t);
        if (lane == 0) decide_and_store(a, sv2, 2, t, one, n_one, (int)(e2 - b2));
    }

    dbg_mark(a, 3, 3);
}

// k_order: one block per shard -- emission order (:206-229) and counters, once every SV of the shard has
// been decided.
__global__ void __launch_bounds__(kThreads)
k_order(PhaseArgs a) {
    __shared__ __align__(16) long long s_tile[kSortSmemBytes / 8];
    pdl_trigger();
    pdl_wait();
    const int s = blockIdx.x;
    if (a.sv_off[s + 1] > a.sv_off[s]) order_block(a, s, s_tile);
}

// ------------------------------------------------------------------------------------------
// k_tail: ONE thread-block CLUSTER per contig does the three per-contig steps that follow k_reduce --
// the sorted unique one-PS list (:107, :195-203), the per-SV decision (:85-183) and the emission order
// with the counters (:206-229) -- in one launch.
//   1. EVERY block reads the whole contig's candidates (a few thousand values out of L2) into a hash set
//      in its own shared memory and sorts the distinct values (they are distinct: rank = place) -- the
//      contig-wide dependency "every SV's candidate before any decision" costs no exchange at all;
//   2. each block decides its slice's SVs against its copy of the list, keeping the results in registers;
//   3. each block publishes how many of its SVs were emitted, whether they came out in order, their
//      smallest / largest sort key and its counters -- cluster barrier -- and reads the other blocks'
//      summaries through distributed shared memory, one lane per block: now it knows its offset in the
//      contig's emission order and whether the contig was in order all the way (the normal case: a
//      compaction; if not, block 0 sorts);
//   4. only then do the per-SV results go out: no store is in flight when the barrier's release runs.
// Host side: used when every contig of the call has at most kTailMaxSvs SVs (any real callset); larger
// contigs take k_oneps / k_predict / k_order.
// ------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int kTailCluster = 8;
constexpr int kTailPer = 2;                                      // SVs per thread
constexpr int kTailSlice = kThreads * kTailPer;                  // SVs per block
constexpr int kTailMaxSvs = kTailCluster * kTailSlice;           // SVs per contig
constexpr int kTailLoads = kTailMaxSvs / kThreads;               // candidates per thread (whole contig)
constexpr int kTailMinSet = 4096;                                // slots: the set's storage doubles as the 16 KB sort tile

struct TailPub {                                                 // what a block shows its cluster
    int n_emit, sorted;
    long long mn, mx;                                            // smallest / largest sort key among the emitted SVs
    unsigned long long cnt[DUET_N_COUNTERS];
};

__device__ __forceinline__ void set_insert(int *set, unsigned mask, int shift, int x) {
    unsigned h = ((unsigned)x * 2654435761u) >> shift;
    for (;;) {                                                   // never full: at most half the slots are ever taken
        const int prev = atomicCAS(set + h, INT32_MIN, x);
        if (prev == INT32_MIN || prev == x) return;
        h = (h + 1) & mask;
    }
}

// the values held by a hash set -> a dense list (in no particular order); returns how many (block-wide; ends
// with a barrier).  Thread t looks at slots t, t + kThreads, ...: consecutive lanes, consecutive banks.
__device__ int set_compact(const int *set, int n_set, int *dst) {
    int cnt = 0;
    for (int i = threadIdx.x; i < n_set; i += kThreads) cnt += set[i] != INT32_MIN;
    int total;
    int w = block_scan_exclusive(cnt, 0, OpSum(), &total);
    for (int i = threadIdx.x; i < n_set; i += kThreads) { const int v = set[i]; if (v != INT32_MIN) dst[w++] = v; }
    __syncthreads();
    return total;
}

__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }

__global__ void __cluster_dims__(kTailCluster, 1, 1) __launch_bounds__(kThreads)
k_tail(PhaseArgs a, int n_set, int n_vals) {
    extern __shared__ __align__(16) int s_dyn[];                 // [hash set, later the sorted list | distinct values]
    __shared__ TailPub s_pub;
    __shared__ Class2Smem s_c2[kThreads / 32];
    __shared__ int s_fb[kTailSlice];
    __shared__ int s_nfb, s_has_min, s_base, s_all_sorted;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int s = (int)(blockIdx.x / kTailCluster);
    int *s_set = s_dyn, *s_vals = s_dyn + n_set;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    dbg_mark(a, 3, 0);
    const int b = (int)a.sv_off[s], n = (int)a.sv_off[s + 1] - b;
    if (n == 0) return;                                          // the whole cluster leaves, before any barrier
    const int per8 = (n + kTailCluster - 1) / kTailCluster;      // this block's slice of the contig ...
    const int per = (per8 + kThreads - 1) / kThreads;            // ... and this thread's chunk of it: <= kTailPer
    const int e_r = min(n, (rank + 1) * per8);
    const int c0 = min(e_r, rank * per8 + (int)threadIdx.x * per), c1 = min(e_r, c0 + per);
    int pos[kTailPer], grp[kTailPer], n_list[kTailPer];
#pragma unroll
    for (int u = 0; u < kTailPer; ++u) {                         // inputs: requested before the wait
        const bool live = c0 + u < c1;
        const int sv = b + c0 + u;
        pos[u] = live ? __ldg(a.sv_pos + sv) : 0;
        grp[u] = live && a.sv_group ? __ldg(a.sv_group + sv) : 0;
        n_list[u] = live ? (int)(__ldg(a.csr_off + sv + 1) - __ldg(a.csr_off + sv)) : 0;
    }
    for (int i = threadIdx.x; i < n_set; i += kThreads) s_set[i] = INT32_MIN;
    if (threadIdx.x == 0) {
        s_pub.n_emit = 0; s_pub.sorted = 1; s_pub.mn = INT64_MAX; s_pub.mx = kNone;
        for (int k = 0; k < DUET_N_COUNTERS; ++k) s_pub.cnt[k] = 0ull;
        s_nfb = 0; s_has_min = 0;
    }
    pdl_trigger();
    pdl_wait();                                                  // k_reduce's per-SV results are final
    long long cv[kTailLoads];
#pragma unroll
    for (int u = 0; u < kTailLoads; ++u) {                       // the whole contig's candidates (coalesced, L2) ...
        const int i = threadIdx.x + u * kThreads;
        cv[u] = i < n ? __ldcg(a.cand + b + i) : kNoCand;
    }
    int cls[kTailPer], nh[kTailPer];
    Class2Stats st[kTailPer];
#pragma unroll
    for (int u = 0; u < kTailPer; ++u) {                         // ... and what k_reduce left for this thread's SVs
        const bool live = c0 + u < c1;
        const int sv = b + c0 + u;
        cls[u] = live ? (int)__ldcg(a.cls + sv) : DUET_CLS_FILTERED;
        nh[u] = live ? __ldcg(a.n_hit + sv) : 0;
        st[u] = Class2Stats{0, 0, 0, 0, 0, 0, 0};
        if (live) st[u] = Class2Stats{__ldcg(a.hap1 + sv), __ldcg(a.hap2 + sv), 0, __ldcg(a.allhap + sv), __ldcg(a.ps + sv),
                                      __ldcg(a.totsc1 + sv), __ldcg(a.totsc2 + sv)};
    }
    __syncthreads();                                             // the set is initialised
    const unsigned set_mask = (unsigned)n_set - 1u;
    const int set_shift = __clz(n_set) + 1;                      // 32 - log2(n_set)
#pragma unroll
    for (int u = 0; u < kTailLoads; ++u) {
        if (cv[u] == kNoCand) continue;
        const int x = (int)cv[u];
        if (x == INT32_MIN) s_has_min = 1;                       // the set's empty marker itself: tracked aside
        else set_insert(s_set, set_mask, set_shift, x);
    }
    __syncthreads();
    dbg_mark(a, 3, 1);
    const int m = set_compact(s_set, n_set, s_vals);             // the contig's distinct candidates
    int *s_one = s_set;                                          // the set is dead: the sorted list goes there
    const int off = s_has_min;
    if (m <= 2 * kThreads) {
        // the values are distinct, so a value's rank IS its place in the sorted list: one pass over the
        // list in shared memory (four values per load) instead of a sorting network of barriers
        if (threadIdx.x < 4) s_vals[m + threadIdx.x] = INT32_MAX;
        __syncthreads();
        const int i0 = threadIdx.x, i1 = threadIdx.x + kThreads;
        const int v0 = i0 < m ? s_vals[i0] : 0, v1 = i1 < m ? s_vals[i1] : 0;
        int k0 = 0, k1 = 0;
        for (int j = 0; j < m; j += 4) {
            const int4 x = *reinterpret_cast<const int4 *>(s_vals + j);
            k0 += (x.x < v0) + (x.y < v0) + (x.z < v0) + (x.w < v0);
            k1 += (x.x < v1) + (x.y < v1) + (x.z < v1) + (x.w < v1);
        }
        if (i0 < m) s_one[off + k0] = v0;
        if (i1 < m) s_one[off + k1] = v1;
    } else {
        const int m_pad = next_pow2(m);
        for (int i = threadIdx.x; i < m_pad; i += kThreads) s_one[off + i] = i < m ? s_vals[i] : INT32_MAX;
        __syncthreads();
        block_bitonic_sort(s_one + off, m_pad);
    }
    if (off && threadIdx.x == 0) s_one[0] = INT32_MIN;
    const int n_one = m + off;                                   // 0: contig skipped (:209-210)
    __syncthreads();
    dbg_mark(a, 3, 2);

    // ---- the decisions of this block's slice (what k_predict does per tile); results stay in registers ----
    Decision dec[kTailPer];
    int g[kTailPer];
    unsigned deferred = 0;
#pragma unroll
    for (int u = 0; u < kTailPer; ++u) {
        g[u] = 0;
        dec[u].valid = false;
        if (c0 + u >= c1 || cls[u] == DUET_CLS_FILTERED || n_one == 0) continue;
        const int sv = b + c0 + u;
        Class2Stats t = st[u];
        if (cls[u] == 0) t = Class2Stats{0, 0, 0, 0, 0, 0, 0};  // get_phase_info skips both loops
        if (cls[u] == 2) {
            const C2Rec *rec = c2_at(a, sv);
            const int allhap = t.allhap;
            t = Class2Stats{0, 0, 0, allhap, 0, 0, 0};
            if (rec->overflow) {                                 // > kC2Max phase sets: the warps take it together below
                s_fb[atomicAdd(&s_nfb, 1)] = c0 + u;
                deferred |= 1u << u;
                continue;
            }
            const int n_d = rec->n_d;
            int best = 0;
            for (int k = 0; k < n_d; ++k) {                      // first-seen order; strict '>' (:101)
                const C2Ent ent = rec->d[k];
                if (!in_sorted(s_one, n_one, ent.ps)) continue;
                if (ent.bad) report(a.status, DUET_ERR_BAD_HP, sv, ent.bad & 0xff);
                if (ent.tot > best) {
                    best = ent.tot;
                    t.h1 = ent.n1; t.h2 = ent.n2; t.t1 = ent.s1; t.t2 = ent.s2; t.ps = ent.ps;
                    t.hap0 = allhap - ent.n1 - ent.n2;
                }
            }
        }
        dec[u] = decide(a, sv, cls[u], t, s_one, n_one, n_list[u]);
        g[u] = dec[u].pred;
    }
    __syncthreads();
    if (s_nfb) {                                                 // rare: stored right away by the warp that computes them
        for (int k = w; k < s_nfb; k += kThreads / 32) {
            const int sv2 = b + s_fb[k];
            const long long b2 = __ldg(a.csr_off + sv2), e2 = __ldg(a.csr_off + sv2 + 1);
            Class2Stats t{0, 0, 0, __ldcg(a.allhap + sv2), 0, 0, 0};
            class2_stats(a, sv2, b2, e2, s_one, n_one, s_c2[w], t);
            if (lane == 0) decide_and_store(a, sv2, 2, t, s_one, n_one, (int)(e2 - b2));
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kTailPer; ++u)
            if (deferred >> u & 1u) g[u] = (int)__ldcg(a.gt + b + c0 + u);
        __threadfence();                                         // block 0 may have to read them (unsorted contig)
    }
    dbg_mark(a, 3, 3);

    // ---- emission order (:206-229) and counters ----
    unsigned long long c_kept = 0, c_emit = 0, c10 = 0, c01 = 0, c11 = 0, c_hits = 0;
    long long key[kTailPer];
    long long mx = kNone, mn = INT64_MAX;
#pragma unroll
    for (int u = 0; u < kTailPer; ++u) {
        c_hits += (unsigned long long)nh[u];
        c_kept += (c0 + u < c1) && cls[u] != DUET_CLS_FILTERED;
        key[u] = kNone;
        if (g[u] != 0) {
            ++c_emit; c10 += g[u] == 1; c01 += g[u] == 2; c11 += g[u] == 3;
            key[u] = (long long)(((unsigned long long)(unsigned)grp[u] << 34) |
                                 ((unsigned long long)((unsigned)pos[u] ^ 0x80000000u) << 2) | (unsigned long long)cls[u]);
            mx = max(mx, key[u]);
            mn = min(mn, key[u]);
        }
    }
    // six small counters in one 64-bit word (a slice has <= 512 SVs: 10 bits each), hits in another
    unsigned long long packed = c_kept | (c_emit << 10) | (c10 << 20) | (c01 << 30) | (c11 << 40);
    packed = warp_sum(packed);
    c_hits = warp_sum(c_hits);
    mn = group_min(mn, 0xffffffffu, 32);
    if (lane == 0) {
        atomicAdd(&s_pub.cnt[1], packed & 1023ull); atomicAdd(&s_pub.cnt[2], (packed >> 10) & 1023ull);
        atomicAdd(&s_pub.cnt[3], (packed >> 20) & 1023ull); atomicAdd(&s_pub.cnt[4], (packed >> 30) & 1023ull);
        atomicAdd(&s_pub.cnt[5], (packed >> 40) & 1023ull); atomicAdd(&s_pub.cnt[7], c_hits);
        if (mn != INT64_MAX) atomicMin(&s_pub.mn, mn);
    }
    long long blk_mx;
    const long long run = block_scan_exclusive(mx, kNone, OpMax(), &blk_mx);
    long long cur = run;
    int cnt = 0;
    bool ok = true;
#pragma unroll
    for (int u = 0; u < kTailPer; ++u)
        if (key[u] != kNone) { if (key[u] < cur) ok = false; cur = max(cur, key[u]); ++cnt; }
    const int sorted_here = __syncthreads_and(ok);
    int n_emit_blk;
    int w0 = block_scan_exclusive(cnt, 0, OpSum(), &n_emit_blk);
    if (threadIdx.x == 0) { s_pub.n_emit = n_emit_blk; s_pub.sorted = sorted_here; s_pub.mx = blk_mx; }
    cluster.sync();                                              // (1) every block's summary can be read

    if (w == 0) {                                                // lane q reads block q's summary, all eight at once
        const int q = lane & (kTailCluster - 1);
        const TailPub *pp = cluster.map_shared_rank(&s_pub, q);
        const int ne = pp->n_emit, so = pp->sorted;
        const long long pmn = pp->mn, pmx = pp->mx;
        unsigned long long tot[DUET_N_COUNTERS];
#pragma unroll
        for (int k = 0; k < DUET_N_COUNTERS; ++k) tot[k] = lane < kTailCluster ? pp->cnt[k] : 0ull;
#pragma unroll
        for (int k = 0; k < DUET_N_COUNTERS; ++k) tot[k] = warp_sum(tot[k]);
        int base = 0, all = 1;
        long long seen_mx = kNone;
#pragma unroll
        for (int r = 0; r < kTailCluster; ++r) {                 // in block order
            const int ne_r = __shfl_sync(0xffffffffu, ne, r), so_r = __shfl_sync(0xffffffffu, so, r);
            const long long mn_r = __shfl_sync(0xffffffffu, pmn, r), mx_r = __shfl_sync(0xffffffffu, pmx, r);
            if (r < rank) base += ne_r;
            all &= so_r;
            if (ne_r) { if (mn_r < seen_mx) all = 0; seen_mx = max(seen_mx, mx_r); }
        }
        if (lane == 0) {
            s_base = base; s_all_sorted = all;
            if (rank == 0 && all) {
                a.n_emit[s] = (int)tot[2];
                long long *c = a.shard_counts + (size_t)s * DUET_N_COUNTERS;
                c[0] = n;
                c[1] = (long long)tot[1]; c[2] = (long long)tot[2]; c[3] = (long long)tot[3];
                c[4] = (long long)tot[4]; c[5] = (long long)tot[5];
                c[6] = a.csr_off[b + n] - a.csr_off[b];
                c[7] = (long long)tot[7];
            }
        }
    }
    __syncthreads();
    cluster_arrive_relaxed();                                    // (2) this block is done reading the others
    const int all_sorted = s_all_sorted;
#pragma unroll
    for (int u = 0; u < kTailPer; ++u)                           // the per-SV results, at last
        if (c0 + u < c1) store_decision(a, b + c0 + u, dec[u]);
    if (all_sorted) {                                            // VCF already in (group, pos, class) order: a compaction
        int wr = b + s_base + w0;
#pragma unroll
        for (int u = 0; u < kTailPer; ++u)
            if (key[u] != kNone) a.order[wr++] = b + c0 + u;
        dbg_mark(a, 3, 4);
        cluster_wait();                                          // nobody leaves while its summary may still be read
        return;
    }
    // the contig's VCF was not sorted: block 0 re-reads every block's decisions and sorts them
    __threadfence();
    cluster_wait();
    cluster.sync();                                              // (3) all decisions are in global memory
    if (rank == 0) order_block(a, s, reinterpret_cast<long long *>(s_dyn));
    dbg_mark(a, 3, 4);
}

}  // namespace duet
