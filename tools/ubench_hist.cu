// Micro-benchmark: what does a 2 M-element bucket histogram / cursor reservation cost on 16384 buckets?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_hist tools/ubench_hist.cu && tools/ubench_hist
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int N = 2000000, NB = 16384;

__device__ __forceinline__ unsigned bucket_of(unsigned long long k) { return (unsigned)(k >> 40) & (NB - 1); }

template <int STRIDE, int REP, bool RET>
__global__ void k_atomic(const unsigned long long *key, unsigned *ctr, unsigned *out) {
    const int base = blockIdx.x * 1024;
    unsigned long long k[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = base + u * 256 + threadIdx.x; k[u] = i < N ? key[i] : 0; }
    unsigned acc = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = base + u * 256 + threadIdx.x;
        if (i >= N) continue;
        unsigned *p = ctr + ((size_t)(blockIdx.x % REP) * NB + bucket_of(k[u])) * STRIDE;
        if (RET) acc += atomicAdd(p, 1u); else atomicAdd(p, 1u);
    }
    if (RET) out[base + threadIdx.x] = acc;
}

// shared-memory aggregation: 1024 threads x 8 elements, flush non-zero bins
template <bool RET>
__global__ void k_smem(const unsigned long long *key, unsigned *ctr, unsigned *out) {
    extern __shared__ unsigned s_h[];
    for (int b = threadIdx.x; b < NB; b += 1024) s_h[b] = 0;
    __syncthreads();
    const int base = blockIdx.x * 8192;
    unsigned long long k[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = base + u * 1024 + threadIdx.x; k[u] = i < N ? key[i] : 0; }
    unsigned r[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = base + u * 1024 + threadIdx.x; r[u] = i < N ? atomicAdd(&s_h[bucket_of(k[u])], 1u) : 0; }
    __syncthreads();
    for (int b = threadIdx.x; b < NB; b += 1024) {
        const unsigned c = s_h[b];
        if (c) { if (RET) s_h[b] = atomicAdd(ctr + b, c); else atomicAdd(ctr + b, c); }
    }
    if (RET) {
        __syncthreads();
        unsigned acc = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = base + u * 1024 + threadIdx.x; if (i < N) acc += s_h[bucket_of(k[u])] + r[u]; }
        out[base / 8 + threadIdx.x] = acc;
    }
}

// plain streaming read of the keys, for the floor
__global__ void k_read(const unsigned long long *key, unsigned *out) {
    const int base = blockIdx.x * 1024;
    unsigned long long a = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = base + u * 256 + threadIdx.x; a += i < N ? key[i] : 0; }
    if (a == 12345) out[0] = 1;
}

template <typename F> void timeit(const char *name, F f, unsigned *ctr, size_t ctr_bytes) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9, sum = 0;
    for (int it = 0; it < 12; ++it) {
        CK(cudaMemset(ctr, 0, ctr_bytes));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); f(); cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
    }
    printf("%-70s best %7.2f us  mean %7.2f us\n", name, best * 1e3, sum / 10 * 1e3);
}

int main() {
    std::vector<unsigned long long> h(N);
    unsigned long long x = 88172645463325252ull;
    for (int i = 0; i < N; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = x; }
    unsigned long long *key; unsigned *ctr, *out;
    const size_t cb = (size_t)NB * 16 * 8 * 4;
    CK(cudaMalloc(&key, N * 8)); CK(cudaMalloc(&ctr, cb)); CK(cudaMalloc(&out, N * 4 + 4096));
    CK(cudaMemcpy(key, h.data(), N * 8, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, NB * 4));
    CK(cudaFuncSetAttribute(k_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, NB * 4));
    const int g = (N + 1023) / 1024, g8 = (N + 8191) / 8192;
    timeit("read 2M keys (floor)", [&] { k_read<<<g, 256>>>(key, out); }, ctr, cb);
    timeit("RED  2M -> 16384 counters", [&] { k_atomic<1, 1, false><<<g, 256>>>(key, ctr, out); }, ctr, cb);
    timeit("RED  2M -> 16384 counters, one per 32-B sector", [&] { k_atomic<8, 1, false><<<g, 256>>>(key, ctr, out); }, ctr, cb);
    timeit("RED  2M -> 16 replicas x 16384 counters", [&] { k_atomic<1, 16, false><<<g, 256>>>(key, ctr, out); }, ctr, cb);
    timeit("ATOM 2M -> 16384 counters (returning)", [&] { k_atomic<1, 1, true><<<g, 256>>>(key, ctr, out); }, ctr, cb);
    timeit("ATOM 2M -> 16384 counters, one per 32-B sector (returning)", [&] { k_atomic<8, 1, true><<<g, 256>>>(key, ctr, out); }, ctr, cb);
    timeit("ATOM 2M -> 16 replicas x 16384 counters (returning)", [&] { k_atomic<1, 16, true><<<g, 256>>>(key, ctr, out); }, ctr, cb);
    timeit("shared-memory histogram per 8192 + flush (RED)", [&] { k_smem<false><<<g8, 1024, NB * 4>>>(key, ctr, out); }, ctr, cb);
    timeit("shared-memory histogram per 8192 + reservation (returning)", [&] { k_smem<true><<<g8, 1024, NB * 4>>>(key, ctr, out); }, ctr, cb);
    return 0;
}
