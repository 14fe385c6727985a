// sm_100a kernels of the sv_phasing hot path.
//
// Reference behaviour restated per kernel (citations: /root/reference/src/duet/sv_phasing_fn.py):
//   k_build   + k_probe   the dict insert / lookup of :26-29 and :46-48 (the JOIN)
//   k_reduce              per-SV class (:192-194), one-PS candidate (:195-203), class-1 counts (:74-84)
//   k_oneps               per-contig set -> sorted unique list (:107)
//   k_predict             class-2 statistics (:85-105), nearest-PS fallback (:106-111),
//                         derived features (:112-139) and the T1-T5 tree (:142-183)
//   k_order               emission order inside a shard (:206-229)
//
// Join direction: the table is built on the SMALL side -- the support-read names of the SVs
// (J entries, L2 resident) -- and the haplotagged reads (R >> J rows) are STREAMED through it
// once, fully coalesced.  A matching row does atomicMax(row index) on its slot, which is the
// reference's "later row overwrites earlier row" rule.  Every shard owns a power-of-two slot
// range, so equal names in different contigs never meet (the reference keeps one dict per contig).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/duet_b200.h"

namespace duet {

constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFull;
constexpr long long kNoCand = 0x7FFFFFFFFFFFFFFFll;   // "no one-PS candidate" (sorts last)
constexpr int kWarpsPerBlock = 8;
constexpr int kMaxDistinct = 32;                      // distinct in-set PS per SV handled in smem

struct DevStatus {          // device -> host error report
    int code;               // first DUET_ERR_* seen (atomicCAS from 0)
    int sv;                 // SV index it was seen at
    long long detail;       // offending value (HP, key, ...)
};

// Everything a kernel needs; passed by value (fits the 4 KB parameter space).
struct PhaseArgs {
    int n_shards;
    int n_reads, n_svs, n_joins;
    // inputs (device)
    const long long *read_off;   // [n_shards+1]
    const long long *sv_off;     // [n_shards+1]
    const unsigned long long *read_key, *read_key_hi;
    const uint8_t *read_hp;
    const int *read_ps, *read_pc;
    const int *sv_pos, *sv_svlen, *sv_svread, *sv_refread;
    const uint8_t *sv_flags;
    const int *sv_group;
    const long long *csr_off;
    const unsigned long long *csr_key, *csr_key_hi;
    // join table: shard s owns slots [tab_off[s], tab_off[s] + tab_mask[s] + 1)
    const int *tab_off;          // [n_shards]
    const int *tab_mask;         // [n_shards]
    unsigned long long *tab_key; // [n_slots]
    unsigned long long *tab_hi;  // [n_slots] hi word of the inserting name (collision check)
    int *tab_row;                // [n_slots] max matching read row, -1 = none
    int *csr_slot;               // [J] slot of each support-read name
    // per-SV intermediates / outputs (device)
    int *join_row;               // [J]
    int *n_hit;                  // [S] joined reads of the SV
    long long *cand;             // [S] one-PS candidate or kNoCand
    int *oneps;                  // [S] shard s: sorted unique list at [sv_off[s], +oneps_n[s])
    int *oneps_n;                // [n_shards]
    long long *sort_scratch;     // [4*S] global fallback for shards too big for shared memory
    int oneps_smem_elems;        // long long elements of dynamic smem given to k_oneps
    int order_smem_elems;        // u128 elements of dynamic smem given to k_order
    uint8_t *gt, *cls;
    int *ps, *hap1, *hap2, *hap0, *allhap;
    long long *totsc1, *totsc2;
    double *features;            // [6][S]
    int *order;                  // [S] per-shard regions
    int *n_emit;                 // [n_shards]
    long long *shard_counts;     // [n_shards][8]
    DevStatus *status;
};

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
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------
// k_build: insert every support-read name into its shard's slot range.  One warp per SV.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_build(PhaseArgs a) {
    const int lane = threadIdx.x & 31;
    const int sv = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (sv >= a.n_svs) return;
    const int s = shard_of(a.sv_off, a.n_shards, sv);
    const int base = __ldg(a.tab_off + s);
    const unsigned mask = (unsigned)__ldg(a.tab_mask + s);
    const long long b = __ldg(a.csr_off + sv), e = __ldg(a.csr_off + sv + 1);
    for (long long j = b + lane; j < e; j += 32) {
        const unsigned long long key = __ldg(a.csr_key + j);
        unsigned p = slot_hash(key) & mask;
        for (;;) {
            const unsigned long long prev = atomicCAS(a.tab_key + base + p, kEmptyKey, key);
            if (prev == kEmptyKey) {
                if (a.csr_key_hi) a.tab_hi[base + p] = __ldg(a.csr_key_hi + j);
                break;
            }
            if (prev == key) break;
            p = (p + 1) & mask;
        }
        a.csr_slot[j] = base + (int)p;
    }
}

// ------------------------------------------------------------------------------------------
// k_probe: stream the haplotagged reads through the table.  kProbePerThread keys per thread,
// all loads issued before the first probe so each thread keeps several L2 requests in flight.
// ------------------------------------------------------------------------------------------
constexpr int kProbeThreads = 256;
constexpr int kProbePerThread = 4;

__device__ __forceinline__ void probe_one(const PhaseArgs &a, unsigned long long key, int row, int base,
                                          unsigned mask) {
    unsigned p = slot_hash(key) & mask;
    for (;;) {
        const unsigned long long k = a.tab_key[base + p];
        if (k == key) {
            if (a.read_key_hi && a.tab_hi[base + p] != __ldg(a.read_key_hi + row)) {
                report(a.status, DUET_ERR_HASH_COLLISION, -1, (long long)key);
                return;
            }
            atomicMax(a.tab_row + base + p, row);
            return;
        }
        if (k == kEmptyKey) return;
        p = (p + 1) & mask;
    }
}

__global__ void __launch_bounds__(kProbeThreads)
k_probe(PhaseArgs a) {
    __shared__ int s_lo, s_hi;
    const long long tile = (long long)blockIdx.x * (kProbeThreads * kProbePerThread);
    if (threadIdx.x == 0) {
        const long long last = min((long long)a.n_reads, tile + kProbeThreads * kProbePerThread) - 1;
        s_lo = shard_of(a.read_off, a.n_shards, tile);
        s_hi = shard_of(a.read_off, a.n_shards, last);
    }
    __syncthreads();
    const int lo = s_lo, hi = s_hi;
    unsigned long long key[kProbePerThread];
    int row[kProbePerThread];
#pragma unroll
    for (int u = 0; u < kProbePerThread; ++u) {
        const long long r = tile + (long long)u * kProbeThreads + threadIdx.x;
        row[u] = r < a.n_reads ? (int)r : -1;
        key[u] = row[u] >= 0 ? __ldcs(a.read_key + r) : 0ull;
    }
    if (lo == hi) {
        const int base = __ldg(a.tab_off + lo);
        const unsigned mask = (unsigned)__ldg(a.tab_mask + lo);
#pragma unroll
        for (int u = 0; u < kProbePerThread; ++u)
            if (row[u] >= 0) probe_one(a, key[u], row[u], base, mask);
    } else {
#pragma unroll
        for (int u = 0; u < kProbePerThread; ++u) {
            if (row[u] < 0) continue;
            const int s = lo + shard_of(a.read_off + lo, hi - lo + 1, row[u]);
            probe_one(a, key[u], row[u], __ldg(a.tab_off + s), (unsigned)__ldg(a.tab_mask + s));
        }
    }
}

// ------------------------------------------------------------------------------------------
// k_reduce: one warp per SV.  Resolves each support read to its read row, then reduces.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_reduce(PhaseArgs a) {
    const int lane = threadIdx.x & 31;
    const int sv = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (sv >= a.n_svs) return;
    const long long b = __ldg(a.csr_off + sv), e = __ldg(a.csr_off + sv + 1);
    // filter of :189-190
    const bool kept = __ldg(a.sv_svlen + sv) >= c_thr.svlen_thres &&
                      __ldg(a.sv_svread + sv) >= c_thr.suppread_thres &&
                      !(__ldg(a.sv_flags + sv) & DUET_SV_GT_MISSING);
    int hits = 0, ps_lo = INT32_MAX, ps_hi = INT32_MIN;
    int h1 = 0, h2 = 0, nq = 0;
    long long t1 = 0, t2 = 0;
    long long first_q = INT64_MAX;      // CSR index of the first read with pc <= pc_max
    int first_q_ps = 0;
    for (long long j = b + lane; j < e; j += 32) {
        const int slot = a.csr_slot[j];
        const int row = a.tab_row[slot];
        if (a.csr_key_hi && a.tab_hi[slot] != __ldg(a.csr_key_hi + j))
            report(a.status, DUET_ERR_HASH_COLLISION, sv, (long long)__ldg(a.csr_key + j));
        a.join_row[j] = row;
        if (row >= 0) {
            const int ps = __ldg(a.read_ps + row);
            const int pc = __ldg(a.read_pc + row);
            const int hp = __ldg(a.read_hp + row);
            ++hits;
            ps_lo = min(ps_lo, ps);
            ps_hi = max(ps_hi, ps);
            if (pc <= c_thr.pc_max) {
                ++nq;
                if (j < first_q) { first_q = j; first_q_ps = ps; }
                if (hp == 1) { ++h1; t1 += pc; }
                else if (hp == 2) { ++h2; t2 += pc; }
            }
        }
    }
    hits = warp_sum(hits);
    ps_lo = warp_min(ps_lo);
    ps_hi = warp_max(ps_hi);
    h1 = warp_sum(h1); h2 = warp_sum(h2); nq = warp_sum(nq);
    t1 = warp_sum(t1); t2 = warp_sum(t2);
    // lane holding the globally first qualifying read
    long long fq = first_q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fq = min(fq, __shfl_xor_sync(0xffffffffu, fq, o));
    const unsigned owner = __ballot_sync(0xffffffffu, first_q == fq && fq != INT64_MAX);
    const int fps = owner ? __shfl_sync(0xffffffffu, first_q_ps, __ffs(owner) - 1) : 0;
    if (lane == 0) {
        const int cls = hits == 0 ? 0 : (ps_lo == ps_hi ? 1 : 2);
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
}

// ------------------------------------------------------------------------------------------
// block-wide bitonic sort of n_pad (power of two) elements, in shared or global memory
// ------------------------------------------------------------------------------------------
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

constexpr int kSortThreads = 1024;
constexpr int kOnepsSmemMaxElems = 16384;    // 128 KB of long long
constexpr int kOrderSmemMaxElems = 8192;     // 128 KB of unsigned __int128

// exclusive block scan of one int per thread (blockDim.x == kSortThreads); returns the prefix,
// *total gets the block sum
__device__ int block_exclusive_scan(int v, int *total) {
    __shared__ int warp_tot[32];
    __shared__ int s_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = warp_tot[lane];
        int ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += n;
        }
        warp_tot[lane] = ti - t;
        if (lane == 31) s_total = ti;
    }
    __syncthreads();
    const int res = warp_tot[w] + inc - v;
    *total = s_total;
    __syncthreads();
    return res;
}

// ------------------------------------------------------------------------------------------
// k_oneps: one block per shard: sort the candidates, keep the distinct ones.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads)
k_oneps(PhaseArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int s = blockIdx.x;
    const int b = (int)a.sv_off[s], n = (int)a.sv_off[s + 1] - b;
    if (n == 0) {
        if (threadIdx.x == 0) a.oneps_n[s] = 0;
        return;
    }
    const int n_pad = next_pow2(n);
    long long *v = n_pad <= a.oneps_smem_elems ? reinterpret_cast<long long *>(smem_raw)
                                            : a.sort_scratch + 2ll * b;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) v[i] = i < n ? a.cand[b + i] : kNoCand;
    __syncthreads();
    block_bitonic_sort(v, n_pad);
    // unique compaction: thread t owns the contiguous chunk [t*per, (t+1)*per)
    const int per = (n_pad + blockDim.x - 1) / blockDim.x;
    const int c0 = threadIdx.x * per, c1 = min(n_pad, c0 + per);
    int cnt = 0;
    for (int i = c0; i < c1; ++i)
        cnt += (v[i] != kNoCand && (i == 0 || v[i] != v[i - 1])) ? 1 : 0;
    int total;
    int w = block_exclusive_scan(cnt, &total);
    for (int i = c0; i < c1; ++i)
        if (v[i] != kNoCand && (i == 0 || v[i] != v[i - 1])) a.oneps[b + w++] = (int)v[i];
    if (threadIdx.x == 0) a.oneps_n[s] = total;
}

// ------------------------------------------------------------------------------------------
// k_predict: one warp per kept SV.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool in_sorted(const int *__restrict__ v, int n, int x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const int m = __ldg(v + mid);
        if (m < x) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(v + lo) == x;
}

// :107-111 -- nearest element of the sorted one-PS list to pos, an exact tie goes up
__device__ __forceinline__ int nearest_ps(const int *__restrict__ v, int n, int pos) {
    int lo = 0, hi = n;               // searchsorted(side='left')
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(v + mid) < pos) lo = mid + 1; else hi = mid;
    }
    const int below = max(lo - 1, 0), above = min(lo, n - 1);
    const long long db = llabs((long long)pos - __ldg(v + below));
    const long long da = llabs((long long)pos - __ldg(v + above));
    return db < da ? __ldg(v + below) : __ldg(v + above);
}

struct Class2Stats { int h1, h2, hap0, allhap, ps; long long t1, t2; };

struct Entry { bool q; bool in; int hp, ps, pc; };

__device__ __forceinline__ Entry load_entry(const PhaseArgs &a, long long j, long long e,
                                            const int *__restrict__ oneps, int n_one) {
    Entry r{false, false, 0, 0, 0};
    if (j < e) {
        const int row = a.join_row[j];
        if (row >= 0) {
            r.pc = __ldg(a.read_pc + row);
            if (r.pc <= c_thr.pc_max) {
                r.q = true;
                r.ps = __ldg(a.read_ps + row);
                r.hp = __ldg(a.read_hp + row);
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

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
k_predict(PhaseArgs a) {
    __shared__ int d_ps[kWarpsPerBlock][kMaxDistinct];
    __shared__ int d_cnt[kWarpsPerBlock][kMaxDistinct][3];           // tot, n1, n2
    __shared__ unsigned long long d_sc[kWarpsPerBlock][kMaxDistinct][2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int sv = blockIdx.x * kWarpsPerBlock + w;
    if (sv >= a.n_svs) return;
    const int cls = a.cls[sv];
    if (cls == DUET_CLS_FILTERED) return;
    const int s = shard_of(a.sv_off, a.n_shards, sv);
    const int n_one = a.oneps_n[s];
    if (n_one == 0) return;                                          // :209-210, gt stays 0
    const int *oneps = a.oneps + a.sv_off[s];
    const long long b = __ldg(a.csr_off + sv), e = __ldg(a.csr_off + sv + 1);
    const int pos = __ldg(a.sv_pos + sv);

    Class2Stats st{a.hap1[sv], a.hap2[sv], 0, a.allhap[sv], a.ps[sv], a.totsc1[sv], a.totsc2[sv]};
    if (cls == 0) { st.h1 = st.h2 = st.allhap = 0; st.t1 = st.t2 = 0; st.ps = 0; }   // get_phase_info skips both loops
    if (cls == 2) {
        st.h1 = st.h2 = st.hap0 = 0; st.t1 = st.t2 = 0; st.ps = 0;
        int n_d = 0;
        bool overflow = false;
        for (long long base = b; base < e && !overflow; base += 32) {
            const Entry x = load_entry(a, base + lane, e, oneps, n_one);
            if (x.in && x.hp != 1 && x.hp != 2) report(a.status, DUET_ERR_BAD_HP, sv, x.hp);
            int id = -1;
            for (int t = 0; t < n_d; ++t)
                if (x.in && d_ps[w][t] == x.ps) id = t;
            unsigned fresh = __ballot_sync(0xffffffffu, x.in && id < 0);
            while (fresh) {                                          // lane order = first-seen order
                const int l0 = __ffs(fresh) - 1;
                const int v = __shfl_sync(0xffffffffu, x.ps, l0);
                const bool mine = x.in && id < 0 && x.ps == v;
                if (n_d == kMaxDistinct) { overflow = true; break; }
                if (lane == 0) {
                    d_ps[w][n_d] = v;
                    d_cnt[w][n_d][0] = d_cnt[w][n_d][1] = d_cnt[w][n_d][2] = 0;
                    d_sc[w][n_d][0] = d_sc[w][n_d][1] = 0ull;
                }
                if (mine) id = n_d;
                ++n_d;
                fresh &= ~__ballot_sync(0xffffffffu, mine);
            }
            __syncwarp();
            if (!overflow && x.in && id >= 0) {
                atomicAdd(&d_cnt[w][id][0], 1);
                if (x.hp == 1 || x.hp == 2) {
                    atomicAdd(&d_cnt[w][id][x.hp], 1);
                    atomicAdd(&d_sc[w][id][x.hp - 1], (unsigned long long)(long long)x.pc);
                }
            }
            __syncwarp();
        }
        if (overflow) {
            class2_slow(a, b, e, oneps, n_one, st);
        } else {
            int best = 0;
            for (int t = 0; t < n_d; ++t) {                          // strict '>' keeps the first seen (:101)
                if (d_cnt[w][t][0] > best) {
                    best = d_cnt[w][t][0];
                    st.h1 = d_cnt[w][t][1]; st.h2 = d_cnt[w][t][2];
                    st.t1 = (long long)d_sc[w][t][0]; st.t2 = (long long)d_sc[w][t][1];
                    st.ps = d_ps[w][t];
                    st.hap0 = st.allhap - st.h1 - st.h2;
                }
            }
        }
    }
    if (lane != 0) return;

    if (cls == 0 || (st.h1 == 0 && st.h2 == 0)) st.ps = nearest_ps(oneps, n_one, pos);     // :106-111
    const int n_list = (int)(e - b);
    const int svread = __ldg(a.sv_svread + sv), refread = __ldg(a.sv_refread + sv);
    if ((long long)svread + refread == 0 || n_list == 0) {
        report(a.status, DUET_ERR_ZERO_DIVISION, sv, 0);
        return;
    }
    // features (:112-132): Python int/int true division == correctly rounded fp64 division
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
    a.gt[sv] = (uint8_t)pred;
    a.ps[sv] = st.ps;
    a.hap1[sv] = st.h1; a.hap2[sv] = st.h2; a.hap0[sv] = st.hap0; a.allhap[sv] = st.allhap;
    a.totsc1[sv] = st.t1; a.totsc2[sv] = st.t2;
    const size_t S = (size_t)a.n_svs;
    a.features[0 * S + sv] = hapread_ratio;
    a.features[1 * S + sv] = sv_ratio;
    a.features[2 * S + sv] = a1;
    a.features[3 * S + sv] = a2;
    a.features[4 * S + sv] = totsc_ratio;
    a.features[5 * S + sv] = avgsc_diff;
}

// ------------------------------------------------------------------------------------------
// k_order: one block per shard: emitted SVs sorted by (group, pos, class, VCF order); counters.
// ------------------------------------------------------------------------------------------
typedef unsigned __int128 u128;

__global__ void __launch_bounds__(kSortThreads)
k_order(PhaseArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long s_cnt[DUET_N_COUNTERS];
    const int s = blockIdx.x;
    const int b = (int)a.sv_off[s], n = (int)a.sv_off[s + 1] - b;
    if (threadIdx.x < DUET_N_COUNTERS) s_cnt[threadIdx.x] = 0ull;
    __syncthreads();
    const u128 kPad = ~(u128)0;
    const int n_pad = next_pow2(max(n, 1));
    u128 *v = n_pad <= a.order_smem_elems ? reinterpret_cast<u128 *>(smem_raw)
                                          : reinterpret_cast<u128 *>(a.sort_scratch) + 2ll * b;
    unsigned long long c_kept = 0, c_emit = 0, c10 = 0, c01 = 0, c11 = 0, c_hits = 0;
    for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        u128 key = kPad;
        if (i < n) {
            const int sv = b + i;
            const int g = a.gt[sv];
            const int cls = a.cls[sv];
            c_hits += (unsigned long long)a.n_hit[sv];
            if (cls != DUET_CLS_FILTERED) ++c_kept;
            if (g != 0) {
                ++c_emit;
                c10 += g == 1; c01 += g == 2; c11 += g == 3;
                const unsigned long long grp = a.sv_group ? (unsigned long long)(unsigned)a.sv_group[sv] : 0ull;
                const unsigned long long upos = (unsigned long long)((unsigned)a.sv_pos[sv] ^ 0x80000000u);
                const unsigned long long hi = (grp << 34) | (upos << 2) | (unsigned long long)cls;
                key = ((u128)hi << 64) | (u128)(unsigned)i;
            }
        }
        v[i] = key;
    }
    c_kept = warp_sum(c_kept); c_emit = warp_sum(c_emit); c10 = warp_sum(c10);
    c01 = warp_sum(c01); c11 = warp_sum(c11); c_hits = warp_sum(c_hits);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_cnt[1], c_kept); atomicAdd(&s_cnt[2], c_emit); atomicAdd(&s_cnt[3], c10);
        atomicAdd(&s_cnt[4], c01); atomicAdd(&s_cnt[5], c11); atomicAdd(&s_cnt[7], c_hits);
    }
    __syncthreads();
    block_bitonic_sort(v, n_pad);
    const int n_emit = (int)s_cnt[2];
    for (int i = threadIdx.x; i < n_emit; i += blockDim.x) a.order[b + i] = b + (int)(unsigned)v[i];
    if (threadIdx.x == 0) {
        a.n_emit[s] = n_emit;
        long long *c = a.shard_counts + (size_t)s * DUET_N_COUNTERS;
        c[0] = n;
        c[1] = (long long)s_cnt[1]; c[2] = (long long)s_cnt[2]; c[3] = (long long)s_cnt[3];
        c[4] = (long long)s_cnt[4]; c[5] = (long long)s_cnt[5];
        c[6] = n ? a.csr_off[b + n] - a.csr_off[b] : 0;
        c[7] = (long long)s_cnt[7];
    }
}

}  // namespace duet
