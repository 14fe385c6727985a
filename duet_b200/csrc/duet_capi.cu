// C-ABI of the sv_phasing hot path (include/duet_b200.h): handle, staging, launches, results.
// The library owns the stream, the join table and every scratch array; callers pass plain
// column pointers.  No CPU fallback exists: every entry point needs a CUDA device.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "phase_kernels.cuh"
#include "cluster_kernels.cuh"
#include "cluster_fast.cuh"

using namespace duet;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes, bool *grew = nullptr) {
        if (grew) *grew = false;
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) { cap = want; if (grew) *grew = true; }
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// page-locked host staging (descriptors on the way in): grows, never shrinks
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

enum { EV_H2D0, EV_H2D1, EV_X0, EV_K0, EV_K1, EV_K2, EV_K3, EV_K4, EV_K5, EV_K6, EV_K7, EV_D2H0, EV_D2H1, EV_DESC, EV_COUNT };

}  // namespace

struct duet_handle {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[EV_COUNT] = {};
    duet_thresholds thr;
    bool thr_dirty = true;
    std::string err;
    int64_t launches = 0;
    bool staged = false, executed = false, per_kernel = false, have_h2d = false, have_d2h = false;

    PhaseArgs a;                    // device view
    std::vector<long long> h_read_off, h_sv_off;
    long long n_slots = 0, n_bm_words = 0;
    int n_sm = 148;
    int flags = 0;                  // developer switches (DUET_FLAGS), see phase_kernels.cuh

    // staged input copies (HOST mode)
    DevBuf in_read_key, in_read_tag;
    DevBuf in_small;                // sv_* and csr_* columns, laid out by input_layout()
    // descriptors (one page-locked staging buffer -> one device buffer, one copy), table, scratch, outputs
    PinBuf h_desc, h_back;          // h_back: status, per-shard emit counts and the raw order on their way out
    bool desc_in_flight = false;    // EV_DESC marks the end of the last descriptor copy
    DevBuf d_desc, d_c2, d_dbg;
    int build_grid = 0, build_per_thread = 1, build_occ[4] = {8, 8, 8, 8};      // k_table<U>, U = 1, 2, 4, 8
    int probe_grid = 0, predict_grid = 0, tail_set = 0, tail_vals = 0;
    bool tail_fused = true;         // every contig fits a cluster: k_tail; else k_oneps / k_predict / k_order
    int reduce_lanes = kReduceLanesSparse;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    PhaseArgs graph_args;           // what the captured launches were given
    int graph_dims[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool dbg_on = false;
    size_t probe_smem = 0;
    DevBuf d_table;                 // Slot[n_slots], swept to all-ones by k_init at the start of every call
    DevBuf d_bitmap;                // Bloom filter words: zeroed at upload, handed back zeroed by k_reduce
    DevBuf d_cand_list, d_cand_count;
    cudaStream_t side_stream = nullptr;   // two-branch mode: k_table's branch
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    DevBuf d_next, d_n_hit, d_cand, d_oneps, d_oneps_n, d_sort;
    DevBuf d_out;                   // every result column (and the join rows), laid out by output_layout()
    DevBuf d_order, d_n_emit, d_status;
    // kernel set B (signature clustering)
    DevBuf cl_in[4], cl_key[2], cl_idx[2], cl_span, cl_parent, cl_minidx, cl_out, cl_hist, cl_misc, cl_dbg;
    DevBuf cl_rec, cl_zone, cl_ctr; // the bucketed path: records, zone lists, per-bucket counters
    cudaEvent_t cl_ev[8] = {};      // staging, then one per stage boundary
    const char *cl_stage[6] = {};   // names of the stages between cl_ev[1..]
    int cl_stages = 0;
    int cl_passes = 0;
    int cl_bucket_blocks = 0;       // resident blocks of k_cl_bucket (0: not asked yet, -1: it does not fit)
    bool cl_timed = false;
};

namespace {

int fail(duet_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}

#define CU(h, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(h, DUET_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

int stage(duet_handle *h, DevBuf &buf, const void *src, size_t bytes, int mem, const void **dev_view) {
    if (src == nullptr) { *dev_view = nullptr; return DUET_OK; }
    if (mem == DUET_MEM_DEVICE) { *dev_view = src; return DUET_OK; }
    CU(h, buf.reserve(bytes ? bytes : 1));
    if (bytes) CU(h, cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
    *dev_view = buf.p;
    return DUET_OK;
}

long long pow2_at_least(long long n) {
    long long p = 1;
    while (p < n) p <<= 1;
    return p;
}

// Arena layouts.  The small input columns and the results each live in ONE device buffer; a caller whose
// host arrays follow the same layout inside one page-locked allocation (duet_phase_input_layout /
// duet_phase_output_layout; engine.pin_batch, engine.pinned_outputs) gets ONE copy each way instead of a
// dozen ~10 us ones.
constexpr int kInCols = 9, kOutCols = 12;
size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
void input_layout(long long S, long long J, size_t off[kInCols], size_t len[kInCols], size_t *total) {
    const size_t S1 = (size_t)std::max<long long>(S, 0), J1 = (size_t)std::max<long long>(J, 0);
    const size_t bytes[kInCols] = {S1 * 4, S1 * 4, S1 * 4, S1 * 4, S1, S1 * 4, (S1 + 1) * 8, J1 * 8, J1 * 4};   // pos svlen svread refread flags group csr_off csr_key csr_chk
    size_t at = 0;
    for (int k = 0; k < kInCols; ++k) { off[k] = at; len[k] = bytes[k]; at = align256(at + bytes[k]); }
    *total = at;
}
void output_layout(long long S, long long J, int ns, size_t off[kOutCols], size_t len[kOutCols], size_t *total) {
    const size_t S1 = (size_t)std::max<long long>(S, 0), J1 = (size_t)std::max<long long>(J, 0);
    const size_t bytes[kOutCols] = {S1, S1 * 4, S1, S1 * 4, S1 * 4, S1 * 4, S1 * 4, S1 * 8, S1 * 8, S1 * 8 * DUET_N_FEATURES,
                                    (size_t)ns * 8 * DUET_N_COUNTERS, J1 * 4};   // gt ps cls hap1 hap2 hap0 allhap totsc1 totsc2 features shard_counts join_row
    size_t at = 0;
    for (int k = 0; k < kOutCols; ++k) { off[k] = at; len[k] = bytes[k]; at = align256(at + bytes[k] + 16); }
    *total = at;
}

// descriptor arena: sections of one host buffer, 256-byte aligned, copied to the device in one piece
struct Arena {
    std::vector<unsigned char> bytes;
    template <typename T> size_t put(const std::vector<T> &v) {
        const size_t off = (bytes.size() + 255) & ~(size_t)255;
        bytes.resize(off + std::max<size_t>(v.size(), 1) * sizeof(T));
        if (!v.empty()) std::memcpy(bytes.data() + off, v.data(), v.size() * sizeof(T));
        return off;
    }
};

}  // namespace

// One launch of the chain.  `pdl`: with the programmatic-serialization attribute the kernel's blocks may be
// scheduled once every block of the previous kernel has started; the kernel itself waits (pdl_wait in
// phase_kernels.cuh) before it touches anything the previous kernel writes.
// `coop`: cooperative launch -- every block of the grid is resident at once (the runtime refuses the launch
// otherwise); not used by the current chain (a cooperative single-kernel join was measured and lost).
static thread_local int g_launch_priority = 0;
template <typename... Params, typename... Args>
static cudaError_t launch(void (*kernel)(Params...), int grid, int block, size_t smem, cudaStream_t st, bool pdl, bool coop,
                          Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[3];
    int n = 0;
    if (g_launch_priority != 0) {                                // set around a launch that should win the SMs (two-branch mode)
        attr[n].id = cudaLaunchAttributePriority;
        attr[n].val.priority = g_launch_priority;
        ++n;
    }
    if (pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (coop) {
        attr[n].id = cudaLaunchAttributeCooperative;
        attr[n].val.cooperative = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, Params(args)...);
}

static size_t tail_smem_bytes(int n_set, int n_vals) { return ((size_t)n_set + (size_t)n_vals) * sizeof(int); }

extern "C" {

int duet_abi_version(void) { return DUET_ABI_VERSION; }

void duet_default_thresholds(duet_thresholds *t) {
    if (!t) return;
    std::memset(t, 0, sizeof(*t));
    t->svlen_thres = 50;
    t->suppread_thres = 2;
    t->pc_max = 8100;
    t->c0_sv_num_min = 4;
    t->c2_sv_num_min = 3;
    t->c2_hap0_min = 6;
    t->c1_ref_num_max = 10;
    t->c2_sv_ratio_min = 0.72;
    t->c2_avgsc_diff_max = 1369.50;
    t->c1_one_ratio_lo = 0.24;
    t->c1_one_ratio_hi = 0.9;
    t->c1_hapread_ratio = 0.75;
    t->c1_avgsc_diff_max = 2400;
    t->c1_two_ratio_a = 0.3;
    t->c1_two_ratio_b = 0.45;
    t->c1_two_ratio_c = 0.75;
    t->c1_totsc_ratio_max = 9.72;
}

int duet_create(int device_id, duet_handle **out) {
    if (!out) return fail(nullptr, DUET_ERR_INVALID, "duet_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, DUET_ERR_NO_DEVICE,
                    std::string("duet_create: no CUDA device (") + cudaGetErrorString(e) +
                        "); this library has no CPU path");
    if (device_id < 0 || device_id >= n) return fail(nullptr, DUET_ERR_INVALID, "duet_create: bad device id");
    duet_handle *h = new duet_handle();
    h->device = device_id;
    if (const char *f = std::getenv("DUET_FLAGS")) h->flags = std::atoi(f);
    duet_default_thresholds(&h->thr);
    std::memset(&h->a, 0, sizeof(h->a));
    if (cudaSetDevice(device_id) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        std::string msg = std::string("duet_create: ") + cudaGetErrorString(cudaGetLastError());
        delete h;
        return fail(nullptr, DUET_ERR_CUDA, msg);
    }
    h->stream = h->own_stream;
    for (auto &ev : h->ev) cudaEventCreate(&ev);
    for (auto &ev : h->cl_ev) cudaEventCreate(&ev);
    {   // fails here, loudly, if the image was not built for this device (sm_100a only)
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, k_probe);
        cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, kBloomMaxWords * 4 + kProbeRingBytes);
        cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, kBloomMaxWords * 4 + kProbeRingBytes);
        cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)tail_smem_bytes((int)pow2_at_least(2 * kTailMaxSvs), kTailMaxSvs + 8));
        // one shared-memory carveout for all the kernels: switching it between launches drains the SMs
        cudaFuncSetAttribute(k_init, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_table<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_table<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_table<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_table<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_probe, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_reduce<kReduceLanesSparse>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_reduce<kReduceLanesDense>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_tail, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_predict, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_oneps, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_order, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, device_id);
        // k_table runs as ONE resident wave: how many of its blocks an SM holds, per names-per-thread variant
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->build_occ[0], k_table<1>, kThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->build_occ[1], k_table<2>, kThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->build_occ[2], k_table<4>, kThreads, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->build_occ[3], k_table<8>, kThreads, 0);
    }
    if (cudaGetLastError() != cudaSuccess) {
        delete h;
        return fail(nullptr, DUET_ERR_CUDA, "duet_create: kernel image not loadable on this device (built for sm_100a)");
    }
    *out = h;
    return DUET_OK;
}

void duet_destroy(duet_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf *bufs[] = {&h->in_read_key, &h->in_read_tag, &h->in_small, &h->d_desc,
                      &h->d_c2, &h->d_dbg, &h->d_table, &h->d_bitmap, &h->d_cand_list, &h->d_cand_count, &h->d_next,
                      &h->d_n_hit, &h->d_cand, &h->d_oneps, &h->d_oneps_n, &h->d_sort, &h->d_out,
                      &h->d_order, &h->d_n_emit, &h->d_status};
    for (DevBuf *b : bufs) b->release();
    h->h_desc.release();
    h->h_back.release();
    for (DevBuf &b : h->cl_in) b.release();
    for (DevBuf *b : {&h->cl_key[0], &h->cl_key[1], &h->cl_idx[0], &h->cl_idx[1], &h->cl_span, &h->cl_parent,
                      &h->cl_minidx, &h->cl_out, &h->cl_hist, &h->cl_misc, &h->cl_dbg, &h->cl_rec, &h->cl_zone, &h->cl_ctr}) b->release();
    for (auto &ev : h->cl_ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : h->ev) if (ev) cudaEventDestroy(ev);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->graph) cudaGraphDestroy(h->graph);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    delete h;
}

const char *duet_last_error(const duet_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int duet_set_thresholds(duet_handle *h, const duet_thresholds *t) {
    if (!h || !t) return fail(h, DUET_ERR_INVALID, "duet_set_thresholds: NULL argument");
    h->thr = *t;
    h->thr_dirty = true;
    return DUET_OK;
}

int duet_set_stream(duet_handle *h, void *cuda_stream) {
    if (!h) return DUET_ERR_INVALID;
    h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return DUET_OK;
}

int duet_host_alloc(void **ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return DUET_ERR_INVALID;
    *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(ptr, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        g_create_error = std::string("duet_host_alloc: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? DUET_ERR_NO_DEVICE : DUET_ERR_CUDA;
    }
    return DUET_OK;
}

int duet_host_is_pinned(const void *ptr) {
    cudaPointerAttributes pa;
    if (!ptr || cudaPointerGetAttributes(&pa, ptr) != cudaSuccess) { cudaGetLastError(); return 0; }
    return pa.type == cudaMemoryTypeHost ? 1 : 0;
}

int duet_host_free(void *ptr) {
    if (!ptr) return DUET_OK;
    return cudaFreeHost(ptr) == cudaSuccess ? DUET_OK : DUET_ERR_CUDA;
}

int64_t duet_phase_input_layout(int64_t n_svs, int64_t n_joins, int64_t *offsets) {
    size_t off[kInCols], len[kInCols], total;
    input_layout(n_svs, n_joins, off, len, &total);
    if (offsets) for (int k = 0; k < kInCols; ++k) offsets[k] = (int64_t)off[k];
    return (int64_t)total;
}

int64_t duet_phase_output_layout(int64_t n_svs, int64_t n_joins, int32_t n_shards, int64_t *offsets) {
    size_t off[kOutCols], len[kOutCols], total;
    output_layout(n_svs, n_joins, n_shards, off, len, &total);
    if (offsets) for (int k = 0; k < kOutCols; ++k) offsets[k] = (int64_t)off[k];
    return (int64_t)total;
}

int duet_phase_upload(duet_handle *h, const duet_phase_input *in) {
    if (!h || !in) return fail(h, DUET_ERR_INVALID, "duet_phase_upload: NULL argument");
    h->staged = h->executed = false;
    const int ns = in->n_shards;
    const long long R = in->n_reads, S = in->n_svs, J = in->n_joins;
    if (ns < 1 || R < 0 || S < 0 || J < 0 || R >= (1ll << 31) || S >= (1ll << 31) || J >= (1ll << 30))
        return fail(h, DUET_ERR_INVALID, "duet_phase_upload: sizes out of range");
    if (!in->read_off || !in->sv_off || !in->csr_off)
        return fail(h, DUET_ERR_INVALID, "duet_phase_upload: offset arrays are required");
    if ((R && (!in->read_key || !in->read_tag)) ||
        (S && (!in->sv_pos || !in->sv_svlen || !in->sv_svread || !in->sv_refread || !in->sv_flags)) ||
        (J && !in->csr_key))
        return fail(h, DUET_ERR_INVALID, "duet_phase_upload: a required column is NULL");
    if (in->mem != DUET_MEM_HOST && in->mem != DUET_MEM_DEVICE && in->mem != DUET_MEM_HOST_MAPPED)
        return fail(h, DUET_ERR_INVALID, "duet_phase_upload: unknown `mem`");
    if (in->mem == DUET_MEM_DEVICE && ((reinterpret_cast<uintptr_t>(in->read_key) | reinterpret_cast<uintptr_t>(in->csr_key)) & 15u))
        return fail(h, DUET_ERR_INVALID, "duet_phase_upload: read_key and csr_key must be 16-byte aligned");
    if (in->mem != DUET_MEM_HOST && (reinterpret_cast<uintptr_t>(in->read_tag) & 15u))      // read in place, 16 bytes at a time
        return fail(h, DUET_ERR_INVALID, "duet_phase_upload: read_tag must be 16-byte aligned");
    if (in->read_off[0] != 0 || in->sv_off[0] != 0 || in->read_off[ns] != R || in->sv_off[ns] != S)
        return fail(h, DUET_ERR_INVALID, "duet_phase_upload: shard offsets do not span the columns");
    for (int s = 0; s < ns; ++s)
        if (in->read_off[s] > in->read_off[s + 1] || in->sv_off[s] > in->sv_off[s + 1])
            return fail(h, DUET_ERR_INVALID, "duet_phase_upload: shard offsets are not monotone");
    CU(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    CU(h, cudaEventRecord(h->ev[EV_H2D0], st));

    // CSR offsets at shard boundaries size the per-shard slot ranges
    std::vector<long long> join_off(ns + 1);
    std::vector<long long> csr_host;
    const long long *csr_view = nullptr;                          // csr_off as the host can read it
    if (in->mem == DUET_MEM_DEVICE) {
        // the columns are resident: fetch the offsets (one small copy)
        csr_host.resize((size_t)S + 1);
        CU(h, cudaMemcpyAsync(csr_host.data(), in->csr_off, (S + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
        if (csr_host[0] != 0 || csr_host[S] != J) return fail(h, DUET_ERR_INVALID, "duet_phase_upload: csr_off does not span csr_key");
        for (int s = 0; s <= ns; ++s) join_off[s] = csr_host[in->sv_off[s]];
        csr_view = csr_host.data();
    } else {
        const long long *csr = reinterpret_cast<const long long *>(in->csr_off);
        if (csr[0] != 0 || csr[S] != J) return fail(h, DUET_ERR_INVALID, "duet_phase_upload: csr_off does not span csr_key");
        for (int s = 0; s <= ns; ++s) join_off[s] = csr[in->sv_off[s]];
        csr_view = csr;
    }
    std::vector<int> tab_off(ns), tab_mask(ns), bm_off(ns), bm_wmask(ns);
    // slot table: 16-byte slots, load factor <= 1/4 while such a table (<= 8 slots of 16 B per name after
    // rounding up to a power of two) stays within half of the 126 MB L2, else <= 1/2: a sparser table means
    // fewer CAS retry rounds in k_table and fewer probe rounds in k_probe (a warp waits for its unluckiest
    // lane), but one that spills out of L2 costs more than it saves.  k_init sweeps it to all-ones at the start
    // of every call, which is also what makes it L2 resident for the scattered traffic that follows.
    // Bloom filter: 16 bits per name, at most 64 KB per shard (it lives in shared memory, two blocks per SM).
    const long long fill = (J * 8 * (long long)sizeof(Slot) <= (64ll << 20) && !(h->flags & kFlagFill2)) ? 4 : 2;
    long long slots = 0, max_sv = 0, bm_words = 0;
    for (int s = 0; s < ns; ++s) {
        const long long nj = join_off[s + 1] - join_off[s];
        if (nj < 0) return fail(h, DUET_ERR_INVALID, "duet_phase_upload: csr_off is not monotone");
        const long long cap = pow2_at_least(std::max<long long>(fill * nj, 32));
        tab_off[s] = (int)slots;
        tab_mask[s] = (int)(cap - 1);
        slots += cap;
        const long long words = std::min<long long>(pow2_at_least(std::max<long long>(nj / 2, 32)), kBloomMaxWords);
        bm_off[s] = (int)bm_words;
        bm_wmask[s] = (int)(words - 1);
        bm_words += words;
        max_sv = std::max<long long>(max_sv, in->sv_off[s + 1] - in->sv_off[s]);
        if (slots >= (1ll << 31)) return fail(h, DUET_ERR_INVALID, "duet_phase_upload: join table too large");
    }
    h->n_slots = slots;
    h->n_bm_words = bm_words;
    h->h_read_off.assign(in->read_off, in->read_off + ns + 1);
    h->h_sv_off.assign(in->sv_off, in->sv_off + ns + 1);

    PhaseArgs &a = h->a;
    std::memset(&a, 0, sizeof(a));
    a.n_shards = ns; a.n_reads = (int)R; a.n_svs = (int)S; a.n_joins = (int)J;
    a.n_slots = slots;
    a.n_bm_words = bm_words;
    a.flags = h->flags;
    // Big calls (dense callsets, cohort shares: >= 1 M support-read names) take the two-branch chain -- k_bloom ->
    // k_stream beside k_init -> k_table, meeting at k_resolve: the read stream runs in the SM slots the table build
    // leaves and under its tail (measured: C4 512 -> 481 us, C5 share 325 -> 302 us).  Small calls keep the serial
    // chain: two more launches and the fork / join edges cost more than the overlap returns (C2 93 -> 95 us, C1 67 -> 69).
    if (J >= (1ll << 20) && !(h->flags & kFlagSerialChain)) a.flags |= kFlagTwoBranch;
    const int mem = in->mem == DUET_MEM_HOST_MAPPED ? DUET_MEM_HOST : in->mem;      // only read_tag is special
    int rc;
#define STAGE(buf, field, T, count)                                                                  \
    if ((rc = stage(h, h->buf, in->field, sizeof(T) * (size_t)(count), mem,                          \
                    reinterpret_cast<const void **>(&a.field))) != DUET_OK) return rc;
    STAGE(in_read_key, read_key, uint64_t, R)
    if (in->mem == DUET_MEM_HOST_MAPPED && R) {
        // the tag records stay where they are: k_reduce gathers the joined rows' records over the bus
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, in->read_tag) != cudaSuccess || pa.type != cudaMemoryTypeHost || !pa.devicePointer) {
            cudaGetLastError();
            return fail(h, DUET_ERR_INVALID, "duet_phase_upload: DUET_MEM_HOST_MAPPED needs read_tag in page-locked host memory");
        }
        a.read_tag = static_cast<const ReadTag *>(pa.devicePointer);
    } else {
        STAGE(in_read_tag, read_tag, duet_read_tag, R)
    }
    {
        size_t off[kInCols], len[kInCols], total;
        input_layout(S, J, off, len, &total);
        const void *src[kInCols] = {in->sv_pos, in->sv_svlen, in->sv_svread, in->sv_refread, in->sv_flags, in->sv_group,
                                    in->csr_off, in->csr_key, in->csr_chk};
        const void **dst[kInCols] = {reinterpret_cast<const void **>(&a.sv_pos), reinterpret_cast<const void **>(&a.sv_svlen),
                                     reinterpret_cast<const void **>(&a.sv_svread), reinterpret_cast<const void **>(&a.sv_refread),
                                     reinterpret_cast<const void **>(&a.sv_flags), reinterpret_cast<const void **>(&a.sv_group),
                                     reinterpret_cast<const void **>(&a.csr_off), reinterpret_cast<const void **>(&a.csr_key),
                                     reinterpret_cast<const void **>(&a.csr_chk)};
        if (mem == DUET_MEM_DEVICE) {
            for (int k = 0; k < kInCols; ++k) *dst[k] = src[k];
        } else {
            CU(h, h->in_small.reserve(total + 256));
            unsigned char *dev = h->in_small.as<unsigned char>();
            // one copy when the host columns sit in one allocation at the layout's offsets (engine.pin_batch, the stage)
            const unsigned char *base = src[0] ? static_cast<const unsigned char *>(src[0]) - off[0] : nullptr;
            bool arena = base != nullptr;
            for (int k = 0; k < kInCols && arena; ++k)
                if (src[k] != nullptr && static_cast<const unsigned char *>(src[k]) != base + off[k]) arena = false;
            if (arena && src[5] == nullptr) arena = false;              // a missing column: the arena would copy garbage over it
            if (arena && src[8] == nullptr) arena = false;
            if (arena) {
                CU(h, cudaMemcpyAsync(dev, base, off[kInCols - 1] + len[kInCols - 1], cudaMemcpyHostToDevice, st));
            } else {
                for (int k = 0; k < kInCols; ++k)
                    if (src[k] && len[k]) CU(h, cudaMemcpyAsync(dev + off[k], src[k], len[k], cudaMemcpyHostToDevice, st));
            }
            for (int k = 0; k < kInCols; ++k) *dst[k] = src[k] ? dev + off[k] : nullptr;
        }
    }
#undef STAGE

    // ---- descriptors: shard offsets, table / filter ranges and the per-block tiles of every kernel (what a
    // block would otherwise look up with dependent loads), laid out in ONE page-locked buffer -> one copy ----
    auto shard_at = [&](const std::vector<long long> &off, long long x) {
        return (int)(std::upper_bound(off.begin(), off.end(), x) - off.begin()) - 1;
    };
    h->reduce_lanes = (S > 0 && J / std::max<long long>(S, 1) > 32) ? kReduceLanesDense : kReduceLanesSparse;
    // dense callsets have a heavy tail (support lists of a thousand reads beside a median of forty): those SVs get a
    // block of their own in k_reduce_heavy instead of one warp of k_reduce
    std::vector<int> heavy_sv;
    if (h->reduce_lanes == kReduceLanesDense && csr_view)
        for (long long i = 0; i < S; ++i)
            if (csr_view[i + 1] - csr_view[i] > kHeavyReads) heavy_sv.push_back((int)i);
    std::stable_sort(heavy_sv.begin(), heavy_sv.end(),            // longest first: the kernel should not end on its longest lists
                     [&](int x, int y) { return csr_view[x + 1] - csr_view[x] > csr_view[y + 1] - csr_view[y]; });
    // the per-contig steps: one cluster per contig (k_tail) when every contig fits one
    h->tail_fused = max_sv <= kTailMaxSvs && !(h->flags & kFlagSplitTail);
    h->tail_set = (int)pow2_at_least(std::max<long long>(2 * max_sv, kTailMinSet));
    h->tail_vals = (int)((std::max<long long>(max_sv, 2 * kThreads) + 3) / 4 * 4 + 4);
    std::vector<PredictTile> ptiles;                             // k_predict blocks never span shards
    if (!h->tail_fused)
        for (int s = 0; s < ns; ++s) {
            const int b = (int)in->sv_off[s], n = (int)(in->sv_off[s + 1] - in->sv_off[s]);
            for (int o = 0; o < n; o += kPredictPerBlock)
                ptiles.push_back(PredictTile{b + o, std::min(b + n, b + o + kPredictPerBlock), s, b, n, {0, 0, 0}});
        }
    h->predict_grid = (int)ptiles.size();
    // k_table: names per thread so that the whole grid is resident at once (no second wave behind the first);
    // when even eight per thread need a second wave (dense callsets, cohorts), two per thread: more, lighter
    // threads then win (measured: 60x dense and the 4-sample cohort share)
    {
        static const int kU[4] = {1, 2, 4, 8};
        int pick = 1;
        for (int k = 0; k < 4; ++k) {
            const long long blocks = (J + (long long)kThreads * kU[k] - 1) / ((long long)kThreads * kU[k]);
            if (blocks <= (long long)h->n_sm * std::max(1, h->build_occ[k])) { pick = k; break; }
        }
        h->build_per_thread = kU[pick];
    }
    const long long build_tile = (long long)kThreads * h->build_per_thread;
    std::vector<BuildTile> btiles((size_t)((J + build_tile - 1) / build_tile));
    for (size_t t = 0; t < btiles.size(); ++t) {
        const long long first = (long long)t * build_tile, last = std::min<long long>(J, first + build_tile) - 1;
        const int lo = shard_at(join_off, first), hi = shard_at(join_off, last);
        btiles[t] = BuildTile{lo, hi, tab_off[lo], tab_mask[lo], bm_off[lo], bm_wmask[lo], {0, 0}};
    }
    h->build_grid = (int)btiles.size();
    // k_probe tiles: row ranges that never cross a contig, about two per SM in total
    std::vector<ProbeTile> qtiles;
    {
        // tiles of about equal size (a contig's rows are cut into round(rows / per) pieces), as many as fit
        // on the device at once: more would mean a second wave costing a whole block time
        long long max_words = 32;
        for (int s = 0; s < ns; ++s) max_words = std::max<long long>(max_words, (long long)bm_wmask[s] + 1);
        h->probe_smem = (size_t)kProbeRingBytes + (size_t)max_words * 4;
        int occ = kProbeBlocksPerSm;                             // big filters leave room for one block per SM only
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_probe, kProbeBlock, h->probe_smem) != cudaSuccess || occ < 1) occ = 1;
        const long long cap = std::max(1, h->n_sm * std::min(occ, kProbeBlocksPerSm));
        long long per = std::max<long long>((R + cap - 1) / cap, 1);
        for (;;) {
            long long total = 0;
            for (int s = 0; s < ns; ++s) {
                const long long rows = h->h_read_off[s + 1] - h->h_read_off[s];
                if (rows > 0) total += std::max<long long>(1, (rows + per / 2) / per);
            }
            if (total <= cap || per >= R) break;
            per += per / 16 + 1;
        }
        for (int s = 0; s < ns; ++s) {
            const long long b0 = h->h_read_off[s], b1 = h->h_read_off[s + 1];
            if (b1 <= b0) continue;
            const long long pieces = std::max<long long>(1, (b1 - b0 + per / 2) / per);
            for (long long k = 0; k < pieces; ++k) {
                long long q0 = b0 + (b1 - b0) * k / pieces, q1 = b0 + (b1 - b0) * (k + 1) / pieces;
                if (k > 0) q0 += q0 & 1;                         // interior cuts on 16-byte boundaries
                if (k + 1 < pieces) q1 += q1 & 1;
                if (q0 < q1) qtiles.push_back(ProbeTile{q0, q1, s, tab_off[s], tab_mask[s], bm_off[s], bm_wmask[s], 0});
            }
        }
        if (qtiles.empty()) qtiles.push_back(ProbeTile{0, 0, 0, 0, 0, 0, 31, 0});
    }
    h->probe_grid = (int)qtiles.size();
    a.n_probe_tiles = h->probe_grid;
    {
        Arena ar;
        const size_t o_read = ar.put(h->h_read_off), o_sv = ar.put(h->h_sv_off), o_join = ar.put(join_off);
        const size_t o_toff = ar.put(tab_off), o_tmask = ar.put(tab_mask), o_boff = ar.put(bm_off), o_bmask = ar.put(bm_wmask);
        const size_t o_bt = ar.put(btiles), o_pt = ar.put(ptiles), o_qt = ar.put(qtiles), o_hv = ar.put(heavy_sv);
        if (h->desc_in_flight) CU(h, cudaEventSynchronize(h->ev[EV_DESC]));      // the previous copy out of h_desc is over
        CU(h, h->h_desc.reserve(ar.bytes.size()));
        CU(h, h->d_desc.reserve(ar.bytes.size()));
        std::memcpy(h->h_desc.p, ar.bytes.data(), ar.bytes.size());
        CU(h, cudaMemcpyAsync(h->d_desc.p, h->h_desc.p, ar.bytes.size(), cudaMemcpyHostToDevice, st));
        CU(h, cudaEventRecord(h->ev[EV_DESC], st));
        h->desc_in_flight = true;
        const unsigned char *d = h->d_desc.as<unsigned char>();
        a.read_off = reinterpret_cast<const long long *>(d + o_read);
        a.sv_off = reinterpret_cast<const long long *>(d + o_sv);
        a.join_off = reinterpret_cast<const long long *>(d + o_join);
        a.tab_off = reinterpret_cast<const int *>(d + o_toff);
        a.tab_mask = reinterpret_cast<const int *>(d + o_tmask);
        a.bm_off = reinterpret_cast<const int *>(d + o_boff);
        a.bm_wmask = reinterpret_cast<const int *>(d + o_bmask);
        a.predict_tiles = reinterpret_cast<const PredictTile *>(d + o_pt);
        a.build_tiles = reinterpret_cast<const BuildTile *>(d + o_bt);
        a.probe_tiles = reinterpret_cast<const ProbeTile *>(d + o_qt);
        a.heavy_sv = heavy_sv.empty() ? nullptr : reinterpret_cast<const int *>(d + o_hv);
        a.n_heavy = (int)heavy_sv.size();
    }
    CU(h, cudaEventRecord(h->ev[EV_H2D1], st));
    h->have_h2d = true;

    const size_t S1 = (size_t)std::max<long long>(S, 1), J1 = (size_t)std::max<long long>(J, 1);
    CU(h, h->d_table.reserve((size_t)std::max<long long>(slots, 1) * sizeof(Slot)));
    a.tab = h->d_table.as<Slot>();
    CU(h, h->d_bitmap.reserve((size_t)std::max<long long>(bm_words, 32) * 4));
    a.bitmap = h->d_bitmap.as<unsigned>();
    CU(h, cudaMemsetAsync(a.bitmap, 0, (size_t)std::max<long long>(bm_words, 32) * 4, st));     // k_reduce keeps it clean from here on
    CU(h, h->d_cand_list.reserve((size_t)std::max<long long>(R, 1) * 16)); a.cand_list = h->d_cand_list.as<ulonglong2>();
    CU(h, h->d_cand_count.reserve((size_t)std::max(h->probe_grid, 1) * 4)); a.cand_count = h->d_cand_count.as<int>();
    CU(h, h->d_next.reserve(J1 * 4));                    a.next = h->d_next.as<int>();
    CU(h, h->d_n_hit.reserve(S1 * 4));                   a.n_hit = h->d_n_hit.as<int>();
    CU(h, h->d_cand.reserve(S1 * 8));                    a.cand = h->d_cand.as<long long>();
    CU(h, h->d_oneps.reserve(S1 * 4));                   a.oneps = h->d_oneps.as<int>();
    CU(h, h->d_oneps_n.reserve((size_t)ns * 4));         a.oneps_n = h->d_oneps_n.as<int>();
    CU(h, h->d_sort.reserve(S1 * 32));                   a.sort_scratch = h->d_sort.as<long long>();
    a.c2_stride = (int)(offsetof(C2Rec, d) + (size_t)c2_cap(h->reduce_lanes) * sizeof(C2Ent));
    CU(h, h->d_c2.reserve(S1 * (size_t)a.c2_stride + sizeof(C2Rec)));          a.c2rec = h->d_c2.as<C2Rec>();

    {
        size_t off[kOutCols], len[kOutCols], total;
        output_layout(S, J, ns, off, len, &total);
        CU(h, h->d_out.reserve(total + 256));
        unsigned char *o = h->d_out.as<unsigned char>();
        a.gt = o + off[0]; a.ps = reinterpret_cast<int *>(o + off[1]); a.cls = o + off[2];
        a.hap1 = reinterpret_cast<int *>(o + off[3]); a.hap2 = reinterpret_cast<int *>(o + off[4]);
        a.hap0 = reinterpret_cast<int *>(o + off[5]); a.allhap = reinterpret_cast<int *>(o + off[6]);
        a.totsc1 = reinterpret_cast<long long *>(o + off[7]); a.totsc2 = reinterpret_cast<long long *>(o + off[8]);
        a.features = reinterpret_cast<double *>(o + off[9]);
        a.shard_counts = reinterpret_cast<long long *>(o + off[10]);
        a.join_row = reinterpret_cast<int *>(o + off[11]);
    }
    CU(h, h->d_order.reserve(S1 * 4));                   a.order = h->d_order.as<int>();
    CU(h, h->d_n_emit.reserve((size_t)ns * 4));          a.n_emit = h->d_n_emit.as<int>();
    CU(h, h->d_status.reserve(sizeof(DevStatus)));       a.status = h->d_status.as<DevStatus>();
    // state the kernels keep clean between calls: counters of shards without SVs stay zero, status zero
    CU(h, cudaMemsetAsync(h->d_oneps_n.p, 0, (size_t)ns * 4, st));
    CU(h, cudaMemsetAsync(h->d_n_emit.p, 0, (size_t)ns * 4, st));
    CU(h, cudaMemsetAsync(a.shard_counts, 0, (size_t)ns * 8 * DUET_N_COUNTERS, st));
    CU(h, cudaMemsetAsync(h->d_status.p, 0, sizeof(DevStatus), st));
    if (h->dbg_on) {
        CU(h, h->d_dbg.reserve((size_t)4 * kDbgBlocks * kDbgMarks * 2 * 8));
        a.dbg = h->d_dbg.as<long long>();
    }
    h->staged = true;
    return DUET_OK;
}

// The launches of one call, a serial chain on `st`: k_init -> k_table -> k_probe -> k_reduce -> k_tail (or, when a
// contig has more SVs than a cluster holds, k_oneps -> k_predict -> k_order).  With `marks`, an event
// follows each stage.
static int launch_all(duet_handle *h, cudaStream_t st, bool marks) {
    auto mark = [&](int ev) { if (marks) cudaEventRecord(h->ev[ev], st); };
    static const bool no_pdl = std::getenv("DUET_NO_PDL") != nullptr;      // diagnostic switch
    const bool pdl = !marks && !no_pdl;
    PhaseArgs a = h->a;
    const int S = a.n_svs;
    const bool join = a.n_joins > 0;
    const bool probe = a.n_reads && a.n_joins;
    if (marks || !probe) a.flags &= ~kFlagTwoBranch;             // the serial (per-kernel timing) chain is the fused one
    int n = 0;
    if (a.flags & kFlagTwoBranch) {
        if (!h->side_stream) {
            cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
            cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
        }
        cudaEventRecord(h->ev_fork, st);                         // the side branch starts with the call
        cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0);
    }
    if (join) { launch(k_init, h->n_sm * 4, kThreads, 0, st, false, false, a); ++n; }
    mark(EV_K0);
    if (a.flags & kFlagTwoBranch) {
        // developer variant: k_table stays behind k_init on this stream (programmatic launch, as in the serial
        // chain: it is the critical path and gets the higher priority); k_bloom -> k_stream run beside it on a second
        // stream; k_resolve waits for both (captured as a fork / join inside the graph)
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);              // hi = numerically smallest = greatest priority
        g_launch_priority = hi;
        switch (h->build_per_thread) {
            case 1: launch(k_table<1>, h->build_grid, kThreads, 0, st, pdl, false, a); break;
            case 2: launch(k_table<2>, h->build_grid, kThreads, 0, st, pdl, false, a); break;
            case 4: launch(k_table<4>, h->build_grid, kThreads, 0, st, pdl, false, a); break;
            default: launch(k_table<8>, h->build_grid, kThreads, 0, st, pdl, false, a); break;
        }
        g_launch_priority = 0;
        launch(k_bloom, h->build_grid, kThreads, 0, h->side_stream, false, false, a, h->build_per_thread);
        launch(k_stream, h->probe_grid, kProbeBlock, h->probe_smem, h->side_stream, pdl, false, a);
        cudaEventRecord(h->ev_join, h->side_stream);
        cudaStreamWaitEvent(st, h->ev_join, 0);
        launch(k_resolve, h->probe_grid, kProbeBlock, 0, st, false, false, a);
        n += 4;                                                  // k_table, k_bloom, k_stream, k_resolve
    } else {
    if (join) {
        switch (h->build_per_thread) {
            case 1: launch(k_table<1>, h->build_grid, kThreads, 0, st, pdl, false, a); break;
            case 2: launch(k_table<2>, h->build_grid, kThreads, 0, st, pdl, false, a); break;
            case 4: launch(k_table<4>, h->build_grid, kThreads, 0, st, pdl, false, a); break;
            default: launch(k_table<8>, h->build_grid, kThreads, 0, st, pdl, false, a); break;
        }
        ++n;
    }
    mark(EV_K1);
    if (probe) { launch(k_probe, h->probe_grid, kProbeBlock, h->probe_smem, st, pdl, false, a); ++n; }
    }
    mark(EV_K2);
    if (S) {
        const int per = kThreads / h->reduce_lanes;
        if (h->reduce_lanes == kReduceLanesDense) launch(k_reduce<kReduceLanesDense>, (S + per - 1) / per, kThreads, 0, st, pdl && join, false, a);
        else launch(k_reduce<kReduceLanesSparse>, (S + per - 1) / per, kThreads, 0, st, pdl && join, false, a);
        ++n;
        if (a.n_heavy > 0) { launch(k_reduce_heavy, a.n_heavy, kThreads, 0, st, pdl, false, a); ++n; }
    }
    mark(EV_K3);
    if (S && h->tail_fused) {
        launch(k_tail, a.n_shards * kTailCluster, kThreads, tail_smem_bytes(h->tail_set, h->tail_vals), st, pdl, false, a,
               h->tail_set, h->tail_vals);
        ++n;
    }
    mark(EV_K4);
    if (S && !h->tail_fused) { launch(k_oneps, a.n_shards, kThreads, 0, st, pdl, false, a); ++n; }
    mark(EV_K5);
    if (S && !h->tail_fused) { launch(k_predict, h->predict_grid, kThreads, 0, st, pdl, false, a); ++n; }
    mark(EV_K6);
    if (S && !h->tail_fused) { launch(k_order, a.n_shards, kThreads, 0, st, pdl, false, a); ++n; }
    return n;
}

int duet_phase_execute(duet_handle *h, int per_kernel) {
    if (!h) return DUET_ERR_INVALID;
    if (!h->staged) return fail(h, DUET_ERR_STATE, "duet_phase_execute: nothing staged (call duet_phase_upload)");
    CU(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    if (h->thr_dirty) {
        CU(h, cudaMemcpyToSymbolAsync(c_thr, &h->thr, sizeof(h->thr), 0, cudaMemcpyHostToDevice, st));
        h->thr_dirty = false;
    }
    h->per_kernel = per_kernel != 0;
    CU(h, cudaEventRecord(h->ev[EV_X0], st));
    if (h->per_kernel) {
        h->launches += launch_all(h, st, true);
    } else {
        // the launch sequence of a staged batch never changes: replay it as a CUDA graph.  A re-upload
        // of the same shapes lands in the same buffers, so the captured graph stays valid.
        const int dims[8] = {h->probe_grid, h->reduce_lanes, (int)h->probe_smem, h->predict_grid,
                             h->build_grid, h->build_per_thread, h->tail_fused ? h->tail_set : 0, h->tail_vals};
        if (h->graph_exec && (std::memcmp(&h->graph_args, &h->a, sizeof(PhaseArgs)) != 0 ||
                              std::memcmp(h->graph_dims, dims, sizeof(dims)) != 0)) {
            cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr;
            cudaGraphDestroy(h->graph); h->graph = nullptr;
        }
        int n_graph = 0;
        if (!h->graph_exec) {
            h->graph_args = h->a;
            std::memcpy(h->graph_dims, dims, sizeof(dims));
            if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                launch_all(h, st, false);
                if (cudaStreamEndCapture(st, &h->graph) != cudaSuccess ||
                    cudaGraphInstantiate(&h->graph_exec, h->graph, 0) != cudaSuccess) {
                    h->graph_exec = nullptr;
                }
            }
            cudaGetLastError();
        }
        if (h->graph_exec) {
            CU(h, cudaGraphLaunch(h->graph_exec, st));
            const PhaseArgs &a = h->a;
            const bool probe = a.n_reads && a.n_joins;
            n_graph = (a.n_joins ? 2 : 0) + (probe ? ((a.flags & kFlagTwoBranch) ? 3 : 1) : 0) + (a.n_svs ? (h->tail_fused ? 2 : 4) : 0) +
                      (a.n_svs && a.n_heavy > 0 ? 1 : 0);
            h->launches += n_graph;
        } else {
            h->launches += launch_all(h, st, false);
        }
    }
    CU(h, cudaEventRecord(h->ev[EV_K7], st));
    CU(h, cudaGetLastError());
    h->executed = true;
    return DUET_OK;
}

int duet_sync(duet_handle *h) {
    if (!h) return DUET_ERR_INVALID;
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    return DUET_OK;
}

int duet_phase_download(duet_handle *h, duet_phase_output *out) {
    if (!h || !out) return fail(h, DUET_ERR_INVALID, "duet_phase_download: NULL argument");
    if (!h->executed) return fail(h, DUET_ERR_STATE, "duet_phase_download: nothing executed");
    CU(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const PhaseArgs &a = h->a;
    const size_t S = (size_t)a.n_svs, J = (size_t)a.n_joins;
    // what the host post-processes (status, emit counts, raw order) lands in page-locked staging: plain DMA
    const size_t ns = (size_t)a.n_shards;
    const size_t off_emit = 64, off_order = (off_emit + ns * 4 + 63) & ~(size_t)63;
    CU(h, h->h_back.reserve(off_order + (out->order ? S * 4 : 0) + 64));
    unsigned char *back = static_cast<unsigned char *>(h->h_back.p);
    DevStatus &status = *reinterpret_cast<DevStatus *>(back);
    int *n_emit = reinterpret_cast<int *>(back + off_emit);
    int *order_raw = reinterpret_cast<int *>(back + off_order);
    CU(h, cudaEventRecord(h->ev[EV_D2H0], st));
    CU(h, cudaMemcpyAsync(&status, a.status, sizeof(status), cudaMemcpyDeviceToHost, st));
    {
        size_t off[kOutCols], len[kOutCols], total;
        output_layout((long long)S, (long long)J, a.n_shards, off, len, &total);
        void *dst[kOutCols] = {out->gt, out->ps, out->cls, out->hap1, out->hap2, out->hap0, out->allhap, out->totsc1, out->totsc2,
                               out->features, out->shard_counts, out->join_row};
        const unsigned char *dev = h->d_out.as<unsigned char>();
        // one copy when the host arrays sit in one allocation at the layout's offsets (engine.pinned_outputs)
        unsigned char *base = dst[0] ? static_cast<unsigned char *>(dst[0]) - off[0] : nullptr;
        int last = -1;
        bool arena = base != nullptr;
        for (int k = 0; k < kOutCols && arena; ++k) {
            if (dst[k] == nullptr) { for (int m = k + 1; m < kOutCols; ++m) if (dst[m]) arena = false; break; }   // only a tail may be missing
            if (static_cast<unsigned char *>(dst[k]) != base + off[k]) arena = false;
            last = k;
        }
        if (arena && last >= 0) {
            CU(h, cudaMemcpyAsync(base, dev, off[last] + len[last], cudaMemcpyDeviceToHost, st));
        } else {
            for (int k = 0; k < kOutCols; ++k)
                if (dst[k] && len[k]) CU(h, cudaMemcpyAsync(dst[k], dev + off[k], len[k], cudaMemcpyDeviceToHost, st));
        }
    }
    CU(h, cudaMemcpyAsync(n_emit, a.n_emit, sizeof(int) * ns, cudaMemcpyDeviceToHost, st));
    if (out->order && S) CU(h, cudaMemcpyAsync(order_raw, a.order, S * 4, cudaMemcpyDeviceToHost, st));
    CU(h, cudaEventRecord(h->ev[EV_D2H1], st));
    CU(h, cudaStreamSynchronize(st));
    h->have_d2h = true;
    if (status.code != 0) {
        char buf[160];
        const char *what = status.code == DUET_ERR_HASH_COLLISION ? "64-bit read-name key collision"
                         : status.code == DUET_ERR_BAD_HP ? "HP outside {1,2} in a multi-phase-set SV"
                         : status.code == DUET_ERR_ZERO_DIVISION ? "svread + refread == 0 (or empty read list)"
                         : "device error";
        std::snprintf(buf, sizeof(buf), "%s (sv=%d detail=%lld)", what, status.sv, status.detail);
        cudaMemsetAsync(h->d_status.p, 0, sizeof(DevStatus), st);
        cudaStreamSynchronize(st);
        return fail(h, status.code, buf);
    }
    // shard regions of `order` -> one compact list
    long long total = 0;
    for (int s = 0; s < a.n_shards; ++s) {
        if (out->order && S)
            std::memcpy(out->order + total, order_raw + h->h_sv_off[s], sizeof(int) * (size_t)n_emit[s]);
        total += n_emit[s];
    }
    out->n_emitted = total;
    return DUET_OK;
}

int duet_phase_run(duet_handle *h, const duet_phase_input *in, duet_phase_output *out) {
    int rc = duet_phase_upload(h, in);
    if (rc == DUET_OK) rc = duet_phase_execute(h, 0);
    if (rc == DUET_OK) rc = duet_phase_download(h, out);
    return rc;
}

int duet_get_timings(duet_handle *h, duet_timings *t) {
    if (!h || !t) return DUET_ERR_INVALID;
    std::memset(t, 0, sizeof(*t));
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->have_h2d) cudaEventElapsedTime(&t->h2d_ms, h->ev[EV_H2D0], h->ev[EV_H2D1]);
    if (h->executed) {
        cudaEventElapsedTime(&t->device_ms, h->ev[EV_X0], h->ev[EV_K7]);
        if (h->per_kernel) {
            const int seq[] = {EV_X0, EV_K0, EV_K1, EV_K2, EV_K3, EV_K4, EV_K5, EV_K6, EV_K7};
            for (int i = 0; i < 8; ++i)
                if (cudaEventElapsedTime(&t->kernel_ms[i], h->ev[seq[i]], h->ev[seq[i + 1]]) != cudaSuccess) t->kernel_ms[i] = 0.f;
        }
    }
    if (h->have_d2h) cudaEventElapsedTime(&t->d2h_ms, h->ev[EV_D2H0], h->ev[EV_D2H1]);
    cudaGetLastError();
    return DUET_OK;
}

int64_t duet_launch_count(const duet_handle *h) { return h ? h->launches : 0; }

int duet_debug_timers(duet_handle *h, int enable, int64_t *out) {
    if (!h) return DUET_ERR_INVALID;
    const size_t bytes = (size_t)4 * kDbgBlocks * kDbgMarks * 2 * 8;
    if (out && h->dbg_on && h->d_dbg.p) {
        CU(h, cudaSetDevice(h->device));
        CU(h, cudaStreamSynchronize(h->stream));
        CU(h, cudaMemcpy(out, h->d_dbg.p, bytes, cudaMemcpyDeviceToHost));
        CU(h, cudaMemset(h->d_dbg.p, 0, bytes));
    }
    h->dbg_on = enable != 0;          // takes effect at the next duet_phase_upload
    return DUET_OK;
}

// ---- kernel set B: signature clustering -----------------------------------------------------------

void duet_default_cluster_params(duet_cluster_params *p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->max_distance = 0.9;            // --cluster_max_distance default, utils.py:27-28
    p->position_normalizer = 900.0;
    p->partition_window = 1000;
}

// the bucketed path (cluster_fast.cuh): five launches chained with programmatic dependent launch.  Returns DUET_OK
// with *done = false when the call has to take the general path (a bucket that does not fit shared memory).
static int cluster_fast(duet_handle *h, const ClusterArgs &g, ClMeta *meta, bool *done) {
    *done = false;
    cudaStream_t st = h->stream;
    const size_t N = (size_t)g.n, NB = (size_t)1 << kBkMaxBits;
    const size_t bucket_smem = sizeof(BucketSmem);
    if (h->cl_bucket_blocks == 0) {
        int per_sm = 0;
        if (cudaFuncSetAttribute(k_cl_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NB * sizeof(unsigned))) == cudaSuccess &&
            cudaFuncSetAttribute(k_cl_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NB * sizeof(unsigned))) == cudaSuccess &&
            cudaFuncSetAttribute(k_cl_bucket, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bucket_smem) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cl_bucket, kBkThreads, bucket_smem) == cudaSuccess && per_sm > 0)
            h->cl_bucket_blocks = per_sm * h->n_sm;
        else
            h->cl_bucket_blocks = -1;
        cudaGetLastError();
    }
    if (h->cl_bucket_blocks < 0) return DUET_OK;
    FastArgs f;
    std::memset(&f, 0, sizeof(f));
    f.n = g.n; f.contig = g.contig; f.type = g.type; f.start = g.start; f.end = g.end;
    CU(h, h->cl_key[0].reserve(N * 8));                  f.key = h->cl_key[0].as<unsigned long long>();
    CU(h, h->cl_rec.reserve(N * 16));                    f.rec = h->cl_rec.as<ulonglong2>();
    CU(h, h->cl_zone.reserve(NB * kZoneCap * 16));       f.zone = h->cl_zone.as<ulonglong2>();
    CU(h, h->cl_ctr.reserve(NB * (4 * 4 + 16)));
    f.bucket_list = h->cl_ctr.as<int4>();
    f.hist = reinterpret_cast<unsigned *>(f.bucket_list + NB); f.cursor = f.hist + NB; f.zone_n = f.cursor + NB;
    f.parent = g.parent; f.pending = g.minidx; f.out = g.out; f.meta = g.meta;
    f.max_distance = g.max_distance; f.normalizer = g.normalizer; f.window2 = g.window2;
    const char *ob = std::getenv("DUET_CL_BITS");       // developer aid: bucket bits
    f.bucket_bits_override = ob ? std::atoi(ob) : -1;
    const int stream_grid = (int)((g.n + kClThreads * kSpUnroll - 1) / (kClThreads * kSpUnroll));
    const bool dbg = std::getenv("DUET_CL_DBG") != nullptr;          // developer aid: cycles per phase of k_cl_bucket
    if (dbg) {
        CU(h, h->cl_dbg.reserve((size_t)4096 * 12 * 8));
        CU(h, cudaMemsetAsync(h->cl_dbg.p, 0, (size_t)4096 * 12 * 8, st));
        f.dbg = h->cl_dbg.as<long long>();
    }
    const int bucket_grid = (int)std::max<long long>(1, std::min<long long>(h->cl_bucket_blocks, (g.n + 255) / 256));
    cudaEvent_t de[5] = {};
    if (dbg) for (auto &e : de) CU(h, cudaEventCreate(&e));
    if (dbg) CU(h, cudaEventRecord(de[0], st));
    CU(h, launch(k_cl_max, std::min(stream_grid, h->n_sm * 8), kClThreads, 0, st, false, false, f));
    if (dbg) CU(h, cudaEventRecord(de[1], st));
    const int ag_grid = (int)((g.n + kAgThreads * kAgItems - 1) / (kAgThreads * kAgItems));
    const size_t ag_smem = NB * sizeof(unsigned);
    CU(h, launch(k_cl_hist, ag_grid, kAgThreads, ag_smem, st, !dbg, false, f));
    if (dbg) CU(h, cudaEventRecord(de[2], st));
    if (dbg) CU(h, cudaEventRecord(de[3], st));
    CU(h, launch(k_cl_scatter, ag_grid, kAgThreads, ag_smem, st, !dbg, false, f));
    if (dbg) CU(h, cudaEventRecord(de[4], st));
    CU(h, cudaEventRecord(h->cl_ev[2], st));
    CU(h, launch(k_cl_bucket, bucket_grid, kBkThreads, bucket_smem, st, false, false, f));
    CU(h, cudaEventRecord(h->cl_ev[3], st));
    CU(h, launch(k_cl_fix, kFixBlocks, kClThreads, 0, st, false, false, f));
    CU(h, cudaEventRecord(h->cl_ev[4], st));
    CU(h, cudaMemcpyAsync(meta, g.meta, sizeof(*meta), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    CU(h, cudaGetLastError());
    if (dbg) {
        static const char *const kK[4] = {"k_cl_max", "k_cl_hist", "-", "k_cl_scatter"};
        for (int k = 0; k < 4; ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, de[k], de[k + 1]);
            std::fprintf(stderr, "%-14s %7.1f us (serial, no programmatic overlap)\n", kK[k], ms * 1e3);
        }
        for (auto &e : de) cudaEventDestroy(e);
        std::vector<long long> d((size_t)bucket_grid * 12);
        CU(h, cudaMemcpy(d.data(), f.dbg, d.size() * 8, cudaMemcpyDeviceToHost));
        // (two stamps with nothing between them: what a stamp itself costs, to be subtracted from every phase)
        static const char *const kPh[10] = {"wait for the block", "records + cell counts", "(a stamp's own cost)", "cell starts", "grouped by cell",
                                            "ranked inside cells", "runs", "window scan", "minima", "output"};
        {
            long long s0 = INT64_MAX, s1 = 0, e0 = INT64_MAX, e1 = 0;
            for (int b = 0; b < bucket_grid; ++b) {
                const long long st0 = d[(size_t)b * 12 + 10], en0 = d[(size_t)b * 12 + 11];
                if (!st0 || !en0) continue;
                s0 = std::min(s0, st0); s1 = std::max(s1, st0); e0 = std::min(e0, en0); e1 = std::max(e1, en0);
            }
            std::fprintf(stderr, "k_cl_bucket blocks start within %.1f us, the first is done after %.1f us, the last after %.1f us\n",
                         (s1 - s0) * 1e-3, (e0 - s0) * 1e-3, (e1 - s0) * 1e-3);
        }
        for (int k = 0; k < 10; ++k) {
            double sum = 0; long long mx = 0;
            for (int b = 0; b < bucket_grid; ++b) { sum += (double)d[(size_t)b * 12 + k]; mx = std::max(mx, d[(size_t)b * 12 + k]); }
            std::fprintf(stderr, "k_cl_bucket %-22s mean %9.0f cycles per block, max %9lld  (%d buckets, %d blocks)\n", kPh[k], sum / bucket_grid, mx,
                         meta->n_buckets, bucket_grid);
        }
    }
    if (meta->oversize && !meta->bad) return DUET_OK;
    static const char *const kNames[3] = {"k_cl_max + k_cl_hist + k_cl_scatter", "k_cl_bucket", "k_cl_fix"};
    for (int k = 0; k < 3; ++k) h->cl_stage[k] = kNames[k];
    h->cl_stages = 3;
    h->launches += 5;
    *done = true;
    return DUET_OK;
}

int duet_cluster_run(duet_handle *h, const duet_cluster_input *in, const duet_cluster_params *params,
                     int32_t *cluster_id, int64_t *n_clusters, float *device_ms) {
    if (!h || !in || !params || !cluster_id) return fail(h, DUET_ERR_INVALID, "duet_cluster_run: NULL argument");
    const long long n = in->n;
    if (n < 0 || n >= (1ll << 31)) return fail(h, DUET_ERR_INVALID, "duet_cluster_run: n out of range");
    if (n_clusters) *n_clusters = 0;
    if (device_ms) *device_ms = 0.f;
    h->cl_timed = false;
    if (n == 0) return DUET_OK;
    if (!in->contig || !in->type || !in->start || !in->end)
        return fail(h, DUET_ERR_INVALID, "duet_cluster_run: a column is NULL");
    if (!(params->max_distance >= 0) || !(params->position_normalizer > 0) || params->partition_window < 0)
        return fail(h, DUET_ERR_INVALID, "duet_cluster_run: bad parameters");
    CU(h, cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    ClusterArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n = (int)n;
    const void *dv;
    int rc;
    const int32_t *cols[4] = {in->contig, in->type, in->start, in->end};
    const int **dst[4] = {&a.contig, &a.type, &a.start, &a.end};
    CU(h, cudaEventRecord(h->cl_ev[0], st));
    for (int c = 0; c < 4; ++c) {
        if ((rc = stage(h, h->cl_in[c], cols[c], sizeof(int32_t) * (size_t)n, in->mem, &dv))) return rc;
        *dst[c] = static_cast<const int *>(dv);
    }
    const size_t N = (size_t)n;
    CU(h, h->cl_parent.reserve(N * 4));  a.parent = h->cl_parent.as<int>();
    CU(h, h->cl_minidx.reserve(N * 4));  a.minidx = h->cl_minidx.as<int>();
    CU(h, h->cl_misc.reserve(sizeof(ClMeta)));
    a.meta = h->cl_misc.as<ClMeta>();
    if (in->mem == DUET_MEM_DEVICE) a.out = cluster_id;
    else { CU(h, h->cl_out.reserve(N * 4)); a.out = h->cl_out.as<int>(); }
    a.max_distance = params->max_distance;
    a.normalizer = params->position_normalizer;
    a.window2 = 2u * (unsigned)params->partition_window;
    static const char kBadInput[] = "duet_cluster_run: need 0 <= start <= end, start+end < 2^32, contig < 65536, type < 256";

    ClMeta meta;
    CU(h, cudaMemsetAsync(h->cl_misc.p, 0, sizeof(ClMeta), st));
    CU(h, cudaEventRecord(h->cl_ev[1], st));
    bool done = false;
    int last_ev = 4;
    if (!std::getenv("DUET_CL_GENERAL")) {                             // developer / test switch: general path only
        if ((rc = cluster_fast(h, a, &meta, &done))) return rc;
        if (done && meta.bad) return fail(h, DUET_ERR_INVALID, kBadInput);
    }
    if (!done) {
        // the general path: global radix sort by (contig, type, c2), then tile kernels over the sorted array.
        // One stream, no host round trip: how many key bits (sort passes) the call needs is decided on the device
        a.n_tiles = (int)((n + kRsTile - 1) / kRsTile);
        for (int k = 0; k < 2; ++k) {
            CU(h, h->cl_key[k].reserve(N * 8));  a.key[k] = h->cl_key[k].as<unsigned long long>();
            CU(h, h->cl_idx[k].reserve(N * 8));  a.pay[k] = h->cl_idx[k].as<unsigned long long>();
        }
        CU(h, h->cl_hist.reserve(((size_t)a.n_tiles + 1) * kRsBins * 4));
        a.block_hist = h->cl_hist.as<unsigned>();
        const int blocks = (int)((n + kClThreads - 1) / kClThreads);
        CU(h, cudaMemsetAsync(h->cl_misc.p, 0, sizeof(ClMeta), st));
        CU(h, cudaEventRecord(h->cl_ev[1], st));
        k_cl_keys<<<blocks, kClThreads, 0, st>>>(a);
        CU(h, cudaEventRecord(h->cl_ev[2], st));
        for (int pass = 0; pass < kRsMaxPasses; ++pass) {
            k_rs_hist<<<a.n_tiles, kClThreads, 0, st>>>(a, pass);
            k_rs_scan<<<kRsBins, kClThreads, 0, st>>>(a, pass);
            k_rs_scatter<<<a.n_tiles, kClThreads, 0, st>>>(a, pass);
        }
        CU(h, cudaEventRecord(h->cl_ev[3], st));
        const int tiles = (int)((n + kClTile - 1) / kClTile);
        const bool cl_dbg = std::getenv("DUET_CL_DBG") != nullptr;         // developer aid: per-block clock stamps of k_cl_edges
        if (cl_dbg) {
            CU(h, h->cl_dbg.reserve((size_t)tiles * kClDbgMarks * 8));
            CU(h, cudaMemsetAsync(h->cl_dbg.p, 0, (size_t)tiles * kClDbgMarks * 8, st));
            a.dbg = h->cl_dbg.as<long long>();
        }
        k_cl_runs<<<tiles, kClThreads, 0, st>>>(a);
        k_cl_edges<<<tiles, kClThreads, 0, st>>>(a);
        CU(h, cudaEventRecord(h->cl_ev[4], st));
        k_cl_label<<<tiles, kClThreads, 0, st>>>(a);
        k_cl_write<<<tiles, kClThreads, 0, st>>>(a);
        CU(h, cudaEventRecord(h->cl_ev[5], st));
        last_ev = 5;
        CU(h, cudaMemcpyAsync(&meta, a.meta, sizeof(meta), cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
        CU(h, cudaGetLastError());
        if (cl_dbg) {
            std::vector<long long> d((size_t)tiles * kClDbgMarks);
            CU(h, cudaMemcpy(d.data(), a.dbg, d.size() * 8, cudaMemcpyDeviceToHost));
            for (int k = 1; k <= 5; ++k) {
                std::vector<long long> v(tiles);
                for (int b = 0; b < tiles; ++b) v[b] = d[(size_t)b * kClDbgMarks + k] - d[(size_t)b * kClDbgMarks + k - 1];
                const int arg = (int)(std::max_element(v.begin(), v.end()) - v.begin());
                std::sort(v.begin(), v.end());
                std::fprintf(stderr, "k_cl_edges phase %d cycles: p50 %lld  p90 %lld  p99 %lld  max %lld (block %d)\n", k, v[tiles / 2],
                             v[(size_t)tiles * 9 / 10], v[(size_t)tiles * 99 / 100], v[tiles - 1], arg);
            }
        }
        if (meta.bad) return fail(h, DUET_ERR_INVALID, kBadInput);
        const int key_bits = (meta.max_c2 ? 32 - __builtin_clz(meta.max_c2) : 0) + (meta.max_type ? 32 - __builtin_clz(meta.max_type) : 0) +
                             (meta.max_contig ? 32 - __builtin_clz(meta.max_contig) : 0);
        h->cl_passes = (key_bits + kRsBits - 1) / kRsBits;
        h->launches += 1 + 3 * h->cl_passes + 4;            // the passes beyond the key's bits return at once: not counted
        static const char *const kNames[4] = {"k_cl_keys", "k_rs_hist + k_rs_scan + k_rs_scatter", "k_cl_runs + k_cl_edges", "k_cl_label + k_cl_write"};
        for (int k = 0; k < 4; ++k) h->cl_stage[k] = kNames[k];
        h->cl_stages = 4;
    }
    if (in->mem != DUET_MEM_DEVICE) {
        CU(h, cudaMemcpyAsync(cluster_id, a.out, N * 4, cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
    }
    h->cl_timed = true;
    if (n_clusters) *n_clusters = meta.n_clusters;
    if (device_ms) cudaEventElapsedTime(device_ms, h->cl_ev[1], h->cl_ev[last_ev]);
    return DUET_OK;
}

int duet_cluster_timings(duet_handle *h, const char **names, float *ms, int cap) {
    if (!h || !names || !ms || !h->cl_timed || cap < h->cl_stages) return 0;
    for (int k = 0; k < h->cl_stages; ++k) {
        names[k] = h->cl_stage[k];
        if (cudaEventElapsedTime(&ms[k], h->cl_ev[1 + k], h->cl_ev[2 + k]) != cudaSuccess) ms[k] = 0.f;
    }
    cudaGetLastError();
    return h->cl_stages;
}

}  // extern "C"
