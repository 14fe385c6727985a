// Developer micro-benchmark: what do scattered atomics / loads / stores cost on this GPU at the sizes of the
// WGS-30x join (402 k names over a 2.4 M-slot table, 393 k filter words, 316 k tag records out of 3.7 M)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench_atomics.cu && gpurun_out/ubench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long mix(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}

struct Slot { unsigned long long key; int first, head; };

template <int U, int MODE>
__global__ void k_scatter(Slot *tab, unsigned *words, const uint4 *tags, int n, unsigned nslots, unsigned nwords, unsigned ntags, int *sink) {
    const int t0 = (blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long key[U];
    int acc = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) key[u] = mix((unsigned long long)(t0 + u * gridDim.x * blockDim.x) * 2654435761ull + 12345);
    if (MODE == 0) {            // 64-bit CAS, value returned
        unsigned long long prev[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) prev[u] = atomicCAS(&tab[key[u] % nslots].key, ~0ull, key[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) acc += prev[u] == ~0ull;
    } else if (MODE == 1) {     // 32-bit atomicOr, no value needed (RED)
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) atomicOr(&words[(key[u] >> 20) % nwords], 1u << (key[u] & 31));
    } else if (MODE == 2) {     // 16-byte load
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) v[u] = __ldcg(reinterpret_cast<const uint4 *>(tab + key[u] % nslots));
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) acc += v[u].x;
    } else if (MODE == 3) {     // 16-byte store
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) reinterpret_cast<uint4 *>(tab)[key[u] % nslots] = make_uint4(~0u, ~0u, ~0u, ~0u);
    } else if (MODE == 4) {     // 32-bit CAS, value returned
        unsigned prev[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) prev[u] = atomicCAS(reinterpret_cast<unsigned *>(&tab[key[u] % nslots].first), ~0u, (unsigned)key[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) acc += prev[u] == ~0u;
    } else if (MODE == 5) {     // tag gather: 16-byte loads out of a big array
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) v[u] = __ldcg(tags + key[u] % ntags);
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) acc += v[u].x;
    } else if (MODE == 6) {     // atomicMax 32-bit no return (RED)
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) atomicMax(reinterpret_cast<int *>(&words[(key[u] >> 20) % nwords]), (int)key[u]);
    } else if (MODE == 7) {     // 64-bit atomicExch with return
        unsigned long long prev[U];
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) prev[u] = atomicExch(&tab[key[u] % nslots].key, key[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) acc += prev[u] == ~0ull;
    } else if (MODE == 8) {     // prefetch.global.L2 of tag records
#pragma unroll
        for (int u = 0; u < U; ++u) if (t0 + u * gridDim.x * blockDim.x < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(tags + key[u] % ntags));
    }
    if (acc == 0x7fffffff) *sink = acc;
}

__global__ void k_sweep(uint4 *p, size_t n16, uint4 v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_readsweep(const uint4 *p, size_t n16, int *sink) {
    unsigned acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) acc += __ldcg(p + i).x;
    if (acc == 0x7fffffff) *sink = 1;
}
__global__ void k_empty() {}

int main(int argc, char **argv) {
    int n = 401992;
    const bool scaling = argc > 1 && !strcmp(argv[1], "scaling");
    const unsigned nslots = 2359296, nwords = 393216, ntags = 3715404;
    Slot *tab; unsigned *words; uint4 *tags; int *sink; unsigned char *flush;
    CK(cudaMalloc(&tab, (size_t)nslots * 16)); CK(cudaMalloc(&words, (size_t)nwords * 4)); CK(cudaMalloc(&tags, (size_t)ntags * 16));
    CK(cudaMalloc(&sink, 4)); CK(cudaMalloc(&flush, 512u << 20));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto reset = [&](bool cold) {
        k_sweep<<<592, 256>>>(reinterpret_cast<uint4 *>(tab), nslots, make_uint4(~0u, ~0u, ~0u, ~0u));
        CK(cudaMemset(words, 0, (size_t)nwords * 4));
        if (cold) CK(cudaMemset(flush, 1, 512u << 20));
        CK(cudaDeviceSynchronize());
    };
    auto timeit = [&](const char *name, bool cold, auto launch) {
        float best = 1e9f, sum = 0;
        for (int it = 0; it < 7; ++it) {
            reset(cold);
            cudaEventRecord(a); launch(); cudaEventRecord(b); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (it >= 2) { best = ms < best ? ms : best; sum += ms; }
        }
        printf("%-64s %s  best %7.2f us  mean %7.2f us\n", name, cold ? "cold" : "warm", best * 1e3f, sum / 5 * 1e3f);
    };
    timeit("empty kernel", false, [&] { k_empty<<<1, 32>>>(); });
    if (scaling) {
        const int sizes[] = {50000, 100000, 200000, 401992, 803984, 1607968, 3215936};
        for (int sz : sizes) {
            n = sz;
            char label[96];
            auto go = [&](auto kern, const char *what, int U) {
                snprintf(label, sizeof label, "n=%7d %s x%d", sz, what, U);
                timeit(label, false, [&] {
                    const int threads = (n + U - 1) / U;
                    kern<<<(threads + 255) / 256, 256>>>(tab, words, tags, n, nslots, nwords, ntags, sink);
                });
            };
            go(k_scatter<2, 0>, "CAS.64 returning", 2);
            go(k_scatter<2, 1>, "atomicOr RED", 2);
            go(k_scatter<2, 2>, "16-B loads (table)", 2);
            go(k_scatter<4, 2>, "16-B loads (table)", 4);
            go(k_scatter<2, 3>, "16-B stores (table)", 2);
        }
        return 0;
    }
#define RUN(U, MODE, label)                                                                                       \
    for (int cold = 0; cold < 2; ++cold)                                                                          \
        timeit(label " x" #U "/thread", cold, [&] {                                                              \
            const int threads = (n + U - 1) / U;                                                                  \
            k_scatter<U, MODE><<<(threads + 255) / 256, 256>>>(tab, words, tags, n, nslots, nwords, ntags, sink); \
        });
    RUN(1, 0, "402k CAS.64 (returning) over 2.4M 16-B slots")
    RUN(2, 0, "402k CAS.64 (returning) over 2.4M 16-B slots")
    RUN(4, 0, "402k CAS.64 (returning) over 2.4M 16-B slots")
    RUN(2, 4, "402k CAS.32 (returning)")
    RUN(2, 7, "402k EXCH.64 (returning)")
    RUN(1, 1, "402k atomicOr.32 (RED) over 393k words")
    RUN(2, 1, "402k atomicOr.32 (RED) over 393k words")
    RUN(2, 6, "402k atomicMax.32 (RED) over 393k words")
    RUN(2, 2, "402k 16-B loads from the 38 MB table")
    RUN(4, 2, "402k 16-B loads from the 38 MB table")
    RUN(2, 3, "402k 16-B stores into the 38 MB table")
    RUN(2, 5, "402k 16-B loads from 59 MB of tag records")
    RUN(4, 5, "402k 16-B loads from 59 MB of tag records")
    RUN(2, 8, "402k prefetch.L2 of tag records")
    for (int cold = 0; cold < 2; ++cold) {
        timeit("write sweep of the 38 MB table (592 blocks)", cold, [&] { k_sweep<<<592, 256>>>(reinterpret_cast<uint4 *>(tab), nslots, make_uint4(~0u, ~0u, ~0u, ~0u)); });
        timeit("read sweep of the 38 MB table (1184 blocks)", cold, [&] { k_readsweep<<<1184, 256>>>(reinterpret_cast<const uint4 *>(tab), nslots, sink); });
        timeit("read sweep of 59 MB of tags (1184 blocks)", cold, [&] { k_readsweep<<<1184, 256>>>(tags, ntags, sink); });
    }
    // does a prefetch pass make the following gather fast?
    for (int cold = 1; cold < 2; ++cold) {
        timeit("prefetch.L2 of 402k tags, then the gather (both timed)", cold, [&] {
            const int threads = (n + 1) / 2;
            k_scatter<2, 8><<<(threads + 255) / 256, 256>>>(tab, words, tags, n, nslots, nwords, ntags, sink);
            k_scatter<2, 5><<<(threads + 255) / 256, 256>>>(tab, words, tags, n, nslots, nwords, ntags, sink);
        });
    }
    return 0;
}
