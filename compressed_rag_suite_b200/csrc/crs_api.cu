// C ABI of libcrs.so (include/crs.h): index object, memory, dispatch.
//
// The index keeps the stored rows of ONE shard contiguous in HBM
// ([count][row_bytes], row_bytes a multiple of 128) plus per-search scratch.
// All kernels are enqueued on the index's stream; nothing here computes on the CPU
// except parameter plumbing (thresholds, plans).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <sys/types.h>
#include <unistd.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "crs_internal.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return e == cudaErrorMemoryAllocation ? CRS_ENOMEM : CRS_ECUDA;
}
#define CRS_CUDA(call)                                   \
    do {                                                 \
        cudaError_t _e = (call);                         \
        if (_e != cudaSuccess) return cuda_fail(_e, #call); \
    } while (0)

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

const int kFloatNch[] = {1, 2, 3, 4, 6, 8, 12, 16};
const int kIntNch[] = {1, 2, 3, 4, 6, 8};

// rows are zero-padded to one of the row widths the scan kernels are instantiated for
int padded_dim_for(int dim, crs_dtype store) {
    const int unit = (store == CRS_F16 || store == CRS_BF16) ? 64 : (store == CRS_I8 ? 128 : 1024);
    const int* tab = (store == CRS_F16 || store == CRS_BF16) ? kFloatNch : kIntNch;
    const int n = (store == CRS_F16 || store == CRS_BF16) ? 8 : 6;
    for (int i = 0; i < n; ++i)
        if (tab[i] * unit >= dim) return tab[i] * unit;
    return -1;
}
size_t row_bytes_for(int dim_padded, crs_dtype store) {
    switch (store) {
        case CRS_F16: case CRS_BF16: return (size_t)dim_padded * 2;
        case CRS_I8: return (size_t)dim_padded;
        default: return (size_t)dim_padded / 8;
    }
}

// temporary device buffer of one call: freed on every return path (cudaFree waits for pending work)
struct DevTmp {
    void* p = nullptr;
    DevTmp() = default;
    DevTmp(const DevTmp&) = delete;
    DevTmp& operator=(const DevTmp&) = delete;
    ~DevTmp() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
};

template <typename T>
struct DevScratch {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct crs_index {
    int dim = 0, dim_padded = 0;
    crs_dtype store = CRS_F16;
    crs_metric metric = CRS_COSINE;
    int device = 0;
    uint32_t row_base = 0;
    size_t row_bytes = 0;
    int64_t count = 0, capacity = 0, reserve_hint = 0;
    uint8_t* codes = nullptr;
    uint32_t* id_map = nullptr;        // optional: global id of every local row (crs_index_map_ids); else row_base + row
    int64_t id_map_cap = 0;
    float i8_scale = 1.0f;
    float row_norm_bound = 1.00390625f;
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    // options
    int force_path = -1;
    int force_exact = 0;
    int gemm_cluster = 0;
    int gemm_prefetch = 0;      // corpus tiles prefetched into L2 ahead of the TMA ring (0 = off)
    int64_t sample_rows = 0;    // rows of an optional sample pass that seeds the contraction's per-query floors (0 = off)
    int share_floor = 1;        // contraction: the slices of a query share their k-th best score while the launch runs
    int gemm_warm = 8;          // contraction: first tiles of every slice that only seed the floor and are redone last
    int gemm_lockstep = 6;      // contraction: tiles a cluster may run ahead of the slowest cluster streaming the same corpus
                                // slice (they share the slice through L2 only while they stay close); 0 = off
    int fuse_encode = 1;        // single-query scans encode the query in their own prologue (no separate encode launch)
    int short_lists = 1;        // integer scans with k > 32 keep 32 keys per CTA + certification (0 = full 128-key lists)
    int multi_scan = 8;         // largest group of short-row integer queries that shares one corpus pass (<= 1: off)
    int gemm_min_nq = 2;        // batches of at least this many queries take the tensor-core path: one corpus read for
                                // the whole batch (a 4-query batch over 10M x 384 fp16: 4.4 ms as scans, 1.1 ms here)
    double eps_scale = 1.0;
    // scratch
    DevScratch<float> qsrc, qnorms, norms_tmp;
    DevScratch<uint8_t> qcodes, stage_rows, vec_codes;
    DevScratch<uint64_t> cand;
    DevScratch<int32_t> flags, counts_dev;
    DevScratch<uint32_t> ids_dev;
    DevScratch<uint8_t> scores_dev;
    DevScratch<uint8_t> out_pack;               // host-output searches: [ids | scores | counts], copied back in one piece
    uint8_t* h_pack = nullptr;                  // pinned staging of that copy
    size_t h_pack_cap = 0;
    DevScratch<uint32_t> allow_dev, floors;     // floors: [tau_q (nq, padded to 128) | pub lists] of the contraction
    int32_t* n_flagged = nullptr;      // device counters: [0] uncertified this search, [1] since create,
                                       // [2] float bits of the largest |fast - exact| score seen in finalize
    crs_search_stats stats{};
    int profiling = 0;
    cudaEvent_t ev_switch = nullptr;           // orders a newly set stream after the previous one
    cudaEvent_t evs[32][2] = {};               // ring of event pairs bracketing the dominant kernel(s) of each search
    int64_t ev_count = 0;                      // searches timed so far
    std::mutex mu;
};

// Peer-memory exchange of one rank (exchange.cu): its receive buffer, the mapped buffers of its peers, the
// step stamp and the staging area the local search writes its [nq, k] result into.
struct crs_exchange {
    int device = 0, rank = 0, world = 1, max_nq = 0, max_k = 0, flag_ctas = 0;
    size_t slots_words = 0, flags_words = 0;
    uint32_t* buf = nullptr;                       // receive buffer: [slots | flags], cudaMalloc'ed and zeroed
    uint32_t* ctl = nullptr;                       // [0] step stamp, [1] error word
    uint32_t* local = nullptr;                     // [ids (max_nq*max_k) | scores (max_nq*max_k) | counts (max_nq)]
    void* peers[crs::kMaxWorld] = {};              // every rank's receive buffer as seen from this device (own = buf)
    bool ipc_opened[crs::kMaxWorld] = {};
    bool wired = false;
    crs::PeerShards shards{};                      // every rank's stored rows (crs_exchange_open_shards), world = 0: none
    bool shard_ipc[crs::kMaxWorld] = {};
    DevScratch<uint8_t> gather;                    // candidate rows fetched for crs_exchange_score_rows
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int grow(crs_index* ix, int64_t need) {
    if (need <= ix->capacity) return CRS_OK;
    int64_t cap = std::max<int64_t>(need, ix->capacity + ix->capacity / 2);
    cap = std::max<int64_t>(cap, ix->reserve_hint);
    cap = std::max<int64_t>(cap, 1024);
    uint8_t* p = nullptr;
    CRS_CUDA(cudaMalloc(&p, (size_t)cap * ix->row_bytes));
    cudaError_t e = cudaSuccess;
    if (ix->count > 0)
        e = cudaMemcpyAsync(p, ix->codes, (size_t)ix->count * ix->row_bytes, cudaMemcpyDeviceToDevice, ix->stream);
    if (e == cudaSuccess && ix->codes) e = cudaStreamSynchronize(ix->stream);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "grow"); }
    if (ix->codes) cudaFree(ix->codes);
    ix->codes = p;
    ix->capacity = cap;
    return CRS_OK;
}

// smallest int raw score whose float similarity passes fl32(sim) >= fl32(min_similarity)
int32_t min_raw_for(const crs_index* ix, float min_similarity) {
    if (!(min_similarity > -INFINITY)) return INT32_MIN;
    if (isnan(min_similarity)) return INT32_MAX;
    auto sim = [&](int64_t raw) -> float {
        if (ix->store == CRS_I8) {
            const float s2 = (float)(((double)ix->i8_scale / 127.0) * ((double)ix->i8_scale / 127.0));
            return (float)(int32_t)raw * s2;
        }
        return (float)((double)raw / (double)ix->dim);
    };
    const int64_t lo = (ix->store == CRS_I8) ? -(int64_t)ix->dim_padded * 127 * 127 : -(int64_t)ix->dim;
    const int64_t hi = -lo;
    if (sim(hi) < min_similarity) return INT32_MAX;
    int64_t a = lo, b = hi;                 // sim is monotone non-decreasing in raw: binary search
    while (a < b) {
        const int64_t m = a + (b - a) / 2;
        if (sim(m) >= min_similarity) b = m; else a = m + 1;
    }
    return (int32_t)a;
}

__global__ void fill_pad_kernel(uint32_t* ids, void* scores, int32_t* counts, int nq, int k, int is_int) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq * k) {
        ids[i] = CRS_PAD_ID;
        if (is_int) reinterpret_cast<int32_t*>(scores)[i] = INT32_MIN;
        else reinterpret_cast<float*>(scores)[i] = -INFINITY;
    }
    if (i < nq) counts[i] = 0;
}
__global__ void iota_ids_kernel(uint32_t* dst, int64_t n, uint32_t first) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = first + (uint32_t)i;
}
__global__ void set_flags_kernel(int32_t* flags, int nq, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) flags[i] = v;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

const char* crs_last_error(void) { return g_last_error.c_str(); }
int crs_version(void) { return 100; }

int crs_index_create(crs_index** out, int dim, crs_dtype store, crs_metric metric,
                     int device, uint32_t row_base, int64_t reserve_rows) {
    if (!out) return fail(CRS_EINVAL, "out is NULL");
    *out = nullptr;
    if (dim <= 0) return fail(CRS_EINVAL, "dim must be positive");
    if (store != CRS_F16 && store != CRS_BF16 && store != CRS_I8 && store != CRS_B1)
        return fail(CRS_EINVAL, "store dtype must be F16, BF16, I8 or B1");
    if (metric != CRS_COSINE && metric != CRS_IP) return fail(CRS_EINVAL, "metric must be COSINE or IP");
    // int8 / 1-bit codes are defined on unit-norm rows (one global scale, sign bits): un-normalised
    // inner-product rows would saturate at +-127 and rank wrongly
    if (metric == CRS_IP && (store == CRS_I8 || store == CRS_B1))
        return fail(CRS_EINVAL, "inner-product space needs a float store (F16 or BF16): I8 / B1 codes are defined on unit rows");
    const int dp = padded_dim_for(dim, store);
    if (dp < 0) return fail(CRS_EINVAL, "dim too large for this store dtype");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(CRS_ECUDA, "no CUDA device: libcrs has no CPU implementation");
    }
    if (device < 0 || device >= ndev) return fail(CRS_EINVAL, "device ordinal out of range");
    cudaDeviceProp prop;
    CRS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(CRS_ECUDA, "libcrs kernels are built for sm_100a only");
    DeviceGuard g(device);
    crs_index* ix = new crs_index();
    ix->dim = dim; ix->dim_padded = dp; ix->store = store; ix->metric = metric;
    ix->device = device; ix->row_base = row_base;
    ix->row_bytes = row_bytes_for(dp, store);
    ix->reserve_hint = reserve_rows;
    ix->num_sms = prop.multiProcessorCount;
    e = cudaMalloc(&ix->n_flagged, 4 * sizeof(int32_t));
    if (e != cudaSuccess) { delete ix; return cuda_fail(e, "cudaMalloc"); }
    cudaMemset(ix->n_flagged, 0, 4 * sizeof(int32_t));
    if (reserve_rows > 0) {
        int rc = grow(ix, reserve_rows);
        if (rc != CRS_OK) { cudaFree(ix->n_flagged); delete ix; return rc; }
    }
    *out = ix;
    return CRS_OK;
}

int crs_index_destroy(crs_index* ix) {
    if (!ix) return CRS_OK;
    {
        DeviceGuard g(ix->device);
        cudaStreamSynchronize(ix->stream);
        if (ix->codes) cudaFree(ix->codes);
        if (ix->id_map) cudaFree(ix->id_map);
        if (ix->n_flagged) cudaFree(ix->n_flagged);
        for (auto& p : ix->evs) { if (p[0]) cudaEventDestroy(p[0]); if (p[1]) cudaEventDestroy(p[1]); }
        if (ix->ev_switch) cudaEventDestroy(ix->ev_switch);
        ix->qsrc.release(); ix->qnorms.release(); ix->norms_tmp.release(); ix->qcodes.release();
        ix->stage_rows.release(); ix->vec_codes.release(); ix->cand.release(); ix->flags.release(); ix->counts_dev.release();
        ix->ids_dev.release(); ix->scores_dev.release(); ix->allow_dev.release(); ix->floors.release(); ix->out_pack.release();
        if (ix->h_pack) cudaFreeHost(ix->h_pack);
    }
    delete ix;
    return CRS_OK;
}

int crs_index_set_stream(crs_index* ix, void* cuda_stream) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    std::lock_guard<std::mutex> lk(ix->mu);
    cudaStream_t ns = reinterpret_cast<cudaStream_t>(cuda_stream);
    if (ns != ix->stream) {
        // the per-index scratch (candidate lists, encoded queries, flags) is shared by consecutive calls:
        // work enqueued on the new stream must not start before what the old stream still has to do
        DeviceGuard g(ix->device);
        cudaStreamCaptureStatus cs_old = cudaStreamCaptureStatusNone, cs_new = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(ix->stream, &cs_old);
        cudaStreamIsCapturing(ns, &cs_new);
        cudaGetLastError();
        if (cs_old == cudaStreamCaptureStatusNone && cs_new == cudaStreamCaptureStatusNone) {
            if (!ix->ev_switch) CRS_CUDA(cudaEventCreateWithFlags(&ix->ev_switch, cudaEventDisableTiming));
            CRS_CUDA(cudaEventRecord(ix->ev_switch, ix->stream));
            CRS_CUDA(cudaStreamWaitEvent(ns, ix->ev_switch, 0));
        }
        ix->stream = ns;
    }
    return CRS_OK;
}

int crs_index_set_option(crs_index* ix, const char* name, int64_t value) {
    if (!ix || !name) return fail(CRS_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (!strcmp(name, "force_path")) ix->force_path = (int)value;
    else if (!strcmp(name, "force_exact")) ix->force_exact = (int)value;
    else if (!strcmp(name, "gemm_cluster")) ix->gemm_cluster = (int)value;
    else if (!strcmp(name, "gemm_min_nq")) ix->gemm_min_nq = (int)value;
    else if (!strcmp(name, "gemm_prefetch")) ix->gemm_prefetch = (int)value;
    else if (!strcmp(name, "multi_scan")) ix->multi_scan = (int)value;
    else if (!strcmp(name, "short_lists")) ix->short_lists = (int)value;
    else if (!strcmp(name, "sample_rows")) ix->sample_rows = value;
    else if (!strcmp(name, "share_floor")) ix->share_floor = (int)value;
    else if (!strcmp(name, "fuse_encode")) ix->fuse_encode = (int)value;
    else if (!strcmp(name, "gemm_lockstep")) ix->gemm_lockstep = (int)value;
    else if (!strcmp(name, "gemm_warm")) ix->gemm_warm = (int)std::max<int64_t>(0, std::min<int64_t>(value, 64));
    else if (!strcmp(name, "eps_scale")) ix->eps_scale = (double)value / 1000.0;
    else if (!strcmp(name, "profiling")) {
        DeviceGuard g(ix->device);
        ix->profiling = (int)value;
        if (ix->profiling && !ix->evs[0][0]) {
            for (auto& p : ix->evs) { CRS_CUDA(cudaEventCreate(&p[0])); CRS_CUDA(cudaEventCreate(&p[1])); }
        }
    }
    else return fail(CRS_EINVAL, std::string("unknown option ") + name);
    return CRS_OK;
}

int crs_index_add(crs_index* ix, const void* rows, int64_t n, crs_dtype src_dtype) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    if (n < 0) return fail(CRS_EINVAL, "n must be >= 0");
    if (n == 0) return CRS_OK;
    if (!rows) return fail(CRS_EINVAL, "rows is NULL");
    if (src_dtype != CRS_F32) return fail(CRS_EINVAL, "rows must be float32");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if ((uint64_t)ix->row_base + (uint64_t)ix->count + (uint64_t)n >= 0xFFFFFFFFull)
        return fail(CRS_EINVAL, "row ids would overflow uint32");
    int rc = grow(ix, ix->count + n);
    if (rc != CRS_OK) return rc;
    const bool dev = is_device_ptr(rows);
    const int64_t chunk = dev ? n : std::min<int64_t>(n, 1 << 16);
    if (!dev) CRS_CUDA(ix->stage_rows.ensure((size_t)chunk * ix->dim * sizeof(float)));
    if (ix->metric == CRS_IP && (ix->store == CRS_F16 || ix->store == CRS_BF16))
        CRS_CUDA(ix->norms_tmp.ensure((size_t)chunk));
    for (int64_t off = 0; off < n; off += chunk) {
        const int64_t m = std::min(chunk, n - off);
        const float* src = reinterpret_cast<const float*>(rows) + off * ix->dim;
        if (!dev) {
            CRS_CUDA(cudaMemcpyAsync(ix->stage_rows.p, src, (size_t)m * ix->dim * sizeof(float),
                                     cudaMemcpyHostToDevice, ix->stream));
            src = reinterpret_cast<const float*>(ix->stage_rows.p);
        }
        float* norms = (ix->metric == CRS_IP && ix->norms_tmp.p) ? ix->norms_tmp.p : nullptr;
        CRS_CUDA(crs::launch_encode(ix->stream, src, m, ix->dim, ix->dim_padded, ix->store, ix->metric,
                                    ix->i8_scale, ix->codes + (size_t)(ix->count + off) * ix->row_bytes, norms));
        if (norms) {   // inner-product space: track the largest stored-row norm for the error bound
            std::vector<float> h((size_t)m);
            CRS_CUDA(cudaMemcpyAsync(h.data(), norms, (size_t)m * sizeof(float), cudaMemcpyDeviceToHost, ix->stream));
            CRS_CUDA(cudaStreamSynchronize(ix->stream));
            for (float v : h) ix->row_norm_bound = std::max(ix->row_norm_bound, v);
        } else if (!dev) {
            CRS_CUDA(cudaStreamSynchronize(ix->stream));   // staging buffer is reused by the next chunk
        }
    }
    ix->count += n;
    return CRS_OK;
}

int crs_index_count(const crs_index* ix, int64_t* out_count) {
    if (!ix || !out_count) return fail(CRS_EINVAL, "bad argument");
    *out_count = ix->count;
    return CRS_OK;
}

int crs_index_map_ids(crs_index* ix, int64_t first_row, int64_t n, uint32_t first_global_id) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    if (first_row < 0 || n < 0 || first_row + n > ix->count) return fail(CRS_EINVAL, "row range outside the index");
    if ((uint64_t)first_global_id + (uint64_t)n >= 0xFFFFFFFFull) return fail(CRS_EINVAL, "row ids would overflow uint32");
    if (n == 0) return CRS_OK;
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t st = ix->stream;
    if (ix->id_map_cap < ix->capacity) {                    // (re)allocate beside the codes; unmapped rows keep row_base + row
        uint32_t* p = nullptr;
        CRS_CUDA(cudaMalloc(&p, (size_t)ix->capacity * sizeof(uint32_t)));
        cudaError_t e = cudaSuccess;
        if (ix->id_map) e = cudaMemcpyAsync(p, ix->id_map, (size_t)ix->id_map_cap * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st);
        const int64_t have = ix->id_map ? ix->id_map_cap : 0;
        if (e == cudaSuccess && ix->capacity > have) {
            iota_ids_kernel<<<(unsigned)((ix->capacity - have + 255) / 256), 256, 0, st>>>(p + have, ix->capacity - have,
                                                                                         ix->row_base + (uint32_t)have);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "crs_index_map_ids"); }
        if (ix->id_map) cudaFree(ix->id_map);
        ix->id_map = p;
        ix->id_map_cap = ix->capacity;
    }
    iota_ids_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ix->id_map + first_row, n, first_global_id);
    CRS_CUDA(cudaGetLastError());
    return CRS_OK;
}

int crs_index_info(const crs_index* ix, int32_t* dim, int32_t* dim_padded, int64_t* row_bytes,
                   int32_t* store, int32_t* metric) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    if (dim) *dim = ix->dim;
    if (dim_padded) *dim_padded = ix->dim_padded;
    if (row_bytes) *row_bytes = (int64_t)ix->row_bytes;
    if (store) *store = (int32_t)ix->store;
    if (metric) *metric = (int32_t)ix->metric;
    return CRS_OK;
}

int crs_index_similarity_scale(const crs_index* ix, double* out_scale) {
    if (!ix || !out_scale) return fail(CRS_EINVAL, "bad argument");
    switch (ix->store) {
        case CRS_I8: *out_scale = (double)(float)(((double)ix->i8_scale / 127.0) * ((double)ix->i8_scale / 127.0)); break;
        case CRS_B1: *out_scale = 1.0 / (double)ix->dim; break;
        default: *out_scale = 1.0;
    }
    return CRS_OK;
}

int crs_index_last_stats(const crs_index* ix, crs_search_stats* out) {
    if (!ix || !out) return fail(CRS_EINVAL, "bad argument");
    *out = ix->stats;
    int32_t tot = 0;
    DeviceGuard g(ix->device);
    int32_t cnt[3] = {0, 0, 0};
    // on the index's own stream (non-blocking streams are not ordered with the legacy stream)
    CRS_CUDA(cudaMemcpyAsync(cnt, ix->n_flagged, sizeof(cnt), cudaMemcpyDeviceToHost, ix->stream));
    CRS_CUDA(cudaStreamSynchronize(ix->stream));
    tot = cnt[1];
    out->uncertified_total = tot;
    memcpy(&out->max_fast_error, &cnt[2], sizeof(float));
    return CRS_OK;
}

int crs_index_last_kernel_ms(crs_index* ix, float* out_ms) {
    if (!ix || !out_ms) return fail(CRS_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (!ix->profiling || ix->ev_count == 0) return fail(CRS_ESTATE, "profiling is off or no search was timed");
    DeviceGuard g(ix->device);
    cudaEvent_t* p = ix->evs[(ix->ev_count - 1) % 32];
    CRS_CUDA(cudaEventSynchronize(p[1]));
    CRS_CUDA(cudaEventElapsedTime(out_ms, p[0], p[1]));
    return CRS_OK;
}

int crs_index_kernel_ms_history(crs_index* ix, float* out_ms, int max_n, int* n_out) {
    if (!ix || !out_ms || !n_out || max_n < 0) return fail(CRS_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    *n_out = 0;
    if (!ix->profiling) return fail(CRS_ESTATE, "profiling is off");
    DeviceGuard g(ix->device);
    const int64_t have = std::min<int64_t>(ix->ev_count, 32);
    const int n = (int)std::min<int64_t>(have, max_n);
    for (int i = 0; i < n; ++i) {                          // oldest of the last n first
        cudaEvent_t* p = ix->evs[(ix->ev_count - n + i) % 32];
        CRS_CUDA(cudaEventSynchronize(p[1]));
        CRS_CUDA(cudaEventElapsedTime(out_ms + i, p[0], p[1]));
    }
    *n_out = n;
    return CRS_OK;
}

namespace {
__global__ void bump_word_kernel(uint32_t* w) { *w += 1u; }
}

// ex / xphase: optional cross-shard exchange (crs_index_search_sharded and its split forms);
// xphase 0 = push + wait + merge, 1 = push only (no outputs)
static int search_impl(crs_index* ix, const void* queries, int nq, int k, float min_similarity,
                       const uint32_t* allow_bits, uint32_t* out_ids, void* out_scores, int32_t* out_counts,
                       crs_exchange* ex = nullptr, int xphase = 0) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    if (nq < 0 || k <= 0) return fail(CRS_EINVAL, "nq must be >= 0 and k > 0");
    if (nq == 0) return CRS_OK;
    const bool push_only = ex != nullptr && xphase == 1;
    if (!queries || (!push_only && (!out_ids || !out_scores || !out_counts))) return fail(CRS_EINVAL, "NULL buffer");
    if (ex != nullptr) {
        if (!ex->wired) return fail(CRS_ESTATE, "exchange is not wired to its peers yet");
        if (ex->device != ix->device) return fail(CRS_EINVAL, "exchange and index live on different devices");
        if (nq > ex->max_nq || k > ex->max_k) return fail(CRS_EINVAL, "nq / k exceed what the exchange was created for");
    }
    const bool is_float = ix->store == CRS_F16 || ix->store == CRS_BF16;
    const bool is_int = !is_float;
    // float stores need head-room above k for certification (finalize.cu)
    int lpl;
    if (is_float) { if (k <= 16) lpl = 1; else if (k <= 112) lpl = 4; else return fail(CRS_EINVAL, "k > 112 not supported for float stores"); }
    else { if (k <= 32) lpl = 1; else if (k <= 128) lpl = 4; else return fail(CRS_EINVAL, "k > 128 not supported"); }
    // batches go to the tcgen05 contraction (K4/K5); single queries / unsupported shapes stream-scan (K1-K3)
    const bool use_gemm = (is_float || ix->store == CRS_I8) && ix->force_path != 0 &&
                          crs::gemm_supported((int)ix->row_bytes, k) && (ix->force_path == 1 || nq >= ix->gemm_min_nq);
    if (use_gemm && is_float) lpl = (k <= 24) ? 1 : 4;     // slice lists hold 16 / 32 keys: M = 32 covers k <= 24
    const int M = 32 * lpl;

    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t st = ix->stream;
    int launches = 0;

    const bool q_dev = is_device_ptr(queries);
    const bool ids_dev = is_device_ptr(out_ids), sc_dev = is_device_ptr(out_scores), cn_dev = is_device_ptr(out_counts);
    if (!push_only && !(ids_dev == sc_dev && sc_dev == cn_dev))
        return fail(CRS_EINVAL, "out_ids/out_scores/out_counts must all be host or all be device buffers");
    const bool out_dev = ids_dev || push_only;
    const size_t nk = (size_t)nq * k;

    uint32_t* d_ids = out_ids; void* d_scores = out_scores; int32_t* d_counts = out_counts;
    const size_t pack_bytes = nk * 8 + (size_t)nq * 4;       // host results: [ids | scores | counts] in ONE device buffer,
    if (!out_dev) {                                          // so that they come back in one copy instead of three
        CRS_CUDA(ix->out_pack.ensure(pack_bytes));
        if (ix->h_pack_cap < pack_bytes) {
            if (ix->h_pack) cudaFreeHost(ix->h_pack);
            ix->h_pack = nullptr; ix->h_pack_cap = 0;
            CRS_CUDA(cudaHostAlloc(&ix->h_pack, pack_bytes + 4096, cudaHostAllocDefault));
            ix->h_pack_cap = pack_bytes + 4096;
        }
        d_ids = reinterpret_cast<uint32_t*>(ix->out_pack.p);
        d_scores = ix->out_pack.p + nk * 4;
        d_counts = reinterpret_cast<int32_t*>(ix->out_pack.p + nk * 8);
    }
    // with an exchange the local search writes into the exchange's staging area and the exchange kernel
    // produces the final (global) result
    uint32_t* f_ids = d_ids; void* f_scores = d_scores; int32_t* f_counts = d_counts;
    if (ex != nullptr) {
        const size_t cap = (size_t)ex->max_nq * ex->max_k;
        d_ids = ex->local; d_scores = ex->local + cap; d_counts = reinterpret_cast<int32_t*>(ex->local + 2 * cap);
    }

    bool copied = false;
    // device pack -> pinned staging (one copy, enqueued); unpack_out() moves it into the caller's arrays after the sync
    auto copy_out = [&]() -> cudaError_t {
        return cudaMemcpyAsync(ix->h_pack, f_ids, pack_bytes, cudaMemcpyDeviceToHost, st);
    };
    auto unpack_out = [&]() {
        memcpy(out_ids, ix->h_pack, nk * 4);
        memcpy(out_scores, ix->h_pack + nk * 4, nk * 4);
        memcpy(out_counts, ix->h_pack + nk * 8, (size_t)nq * 4);
    };

    if (ix->count == 0) {
        fill_pad_kernel<<<(unsigned)((std::max(nk, (size_t)nq) + 255) / 256), 256, 0, st>>>(d_ids, d_scores, d_counts, nq, k, is_int);
        CRS_CUDA(cudaGetLastError());
        ++launches;
        if (ex != nullptr) {                      // an empty shard still takes part in the step
            bump_word_kernel<<<1, 1, 0, st>>>(ex->ctl);
            CRS_CUDA(cudaGetLastError());
            ++launches;
        }
    } else {
        // ---- queries -> canonical stored codes
        const float* qd = reinterpret_cast<const float*>(queries);
        if (!q_dev) {
            CRS_CUDA(ix->qsrc.ensure((size_t)nq * ix->dim));
            CRS_CUDA(cudaMemcpyAsync(ix->qsrc.p, queries, (size_t)nq * ix->dim * sizeof(float), cudaMemcpyHostToDevice, st));
            qd = ix->qsrc.p;
        }
        CRS_CUDA(ix->qcodes.ensure((size_t)nq * ix->row_bytes));
        CRS_CUDA(ix->qnorms.ensure((size_t)nq));
        // ---- plan: persistent grid, one candidate list per CTA
        crs::ScanPlan plan;
        plan.lpl = lpl;
        plan.grid = ix->num_sms;
        // a single-query scan encodes the query in its own prologue; everything else gets an encode launch
        const bool fused_q = !use_gemm && nq == 1 && ix->fuse_encode != 0 && ix->dim <= 2048 &&
                             !(is_float && ix->force_exact != 0);
        if (fused_q) {
            plan.fq.src = qd; plan.fq.dim = ix->dim; plan.fq.cosine = ix->metric == CRS_COSINE;
            plan.fq.i8_mult = 127.0 / (double)ix->i8_scale;
            plan.fq.qcodes_out = ix->qcodes.p; plan.fq.qnorm_out = ix->qnorms.p;
            plan.fq.zero_word = ix->n_flagged; plan.fq.inc_word = ex ? ex->ctl : nullptr;
        } else {
            CRS_CUDA(crs::launch_encode(st, qd, nq, ix->dim, ix->dim_padded, ix->store, ix->metric, ix->i8_scale,
                                        ix->qcodes.p, ix->qnorms.p, ix->n_flagged /*reset "uncertified this search"*/,
                                        ex ? ex->ctl : nullptr /*advance the exchange's step stamp*/));
            ++launches;
        }
        if (allow_bits) {                       // row bitmap of a where / where_document filter
            const size_t words = (size_t)((ix->count + 31) / 32);
            if (is_device_ptr(allow_bits)) {
                plan.allow = allow_bits;
            } else {
                CRS_CUDA(ix->allow_dev.ensure(words));
                CRS_CUDA(cudaMemcpyAsync(ix->allow_dev.p, allow_bits, words * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
                plan.allow = ix->allow_dev.p;
            }
        }
        const int n_lists = plan.grid;
        CRS_CUDA(ix->cand.ensure((size_t)nq * n_lists * M));
        CRS_CUDA(ix->flags.ensure((size_t)nq));

        crs::FinalizeArgs fa{};
        fa.cand = ix->cand.p; fa.n_lists = n_lists; fa.list_len = M; fa.list_stride = M; fa.lpl = lpl; fa.nq = nq; fa.k = k;
        fa.codes = ix->codes; fa.qcodes = ix->qcodes.p; fa.qnorms = ix->qnorms.p;
        fa.dim_padded = ix->dim_padded; fa.bf16 = ix->store == CRS_BF16;
        fa.row_norm_bound = ix->row_norm_bound;
        fa.min_similarity = min_similarity; fa.row_base = ix->row_base; fa.id_map = ix->id_map;
        fa.out_ids = d_ids; fa.out_scores = d_scores; fa.out_counts = d_counts;
        fa.flags = ix->flags.p; fa.n_flagged = ix->n_flagged; fa.is_int = is_int;

        // integer scans with k > 32: every CTA keeps only its best 32 keys (the insert cost of a 128-key
        // list is what bounds shared-pass batches) and finalize certifies that no CTA held more than 32
        // of the global top-k; a query that fails is recomputed with full lists by the exact pass
        const bool short_lists = is_int && !use_gemm && lpl == 4 && ix->short_lists != 0;
        const int scan_lpl = short_lists ? 1 : lpl;
        const int SM = 32 * scan_lpl;                       // keys per CTA list of the fast pass
        plan.lpl = scan_lpl;
        if (!use_gemm) { fa.list_len = SM; fa.list_stride = SM; }
        ix->stats.path = 0; ix->stats.grid = plan.grid; ix->stats.list_len = SM;
        const int32_t min_raw = is_int ? min_raw_for(ix, min_similarity) : 0;
        bool need_exact = is_float && ix->force_exact != 0;      // test hook: skip the fast pass
        bool certifying = false;                                  // the fast pass can flag queries for the exact pass

        if (!need_exact) {
            // ---- fast pass: candidate lists
            float tau_pre = -INFINITY;
            if (is_float) {
                // error bound of the fp32 scan score (any summation order): Dp * 2^-24 * |q| * |c|;
                // tensor-core accumulation may truncate instead of round: one ulp per term
                fa.eps_rel = (float)((double)ix->dim_padded * ldexp(1.0, use_gemm ? -23 : -24) * (use_gemm ? 1.05 : 1.01) *
                                     ix->eps_scale);
                // cosine queries are unit rows (|q| <= 1.0039); for ip the norm is only known on the
                // device, so the pre-filter is left off and the threshold applied on the exact score.
                if (ix->metric == CRS_COSINE && min_similarity > -INFINITY)
                    tau_pre = min_similarity - fa.eps_rel * 1.00390625f * ix->row_norm_bound;
            }
            if (ix->profiling) CRS_CUDA(cudaEventRecord(ix->evs[ix->ev_count % 32][0], st));
            if (use_gemm) {
                int n_slices = 0;
                uint32_t tau_bits;
                if (is_float) memcpy(&tau_bits, &tau_pre, sizeof(tau_bits)); else tau_bits = (uint32_t)min_raw;
                const int kind = ix->store == CRS_F16 ? 0 : (ix->store == CRS_BF16 ? 1 : 2);
                // Per-query floors.  The slices of a query run on different SMs with a list each; left alone every
                // list warms up by itself.  (a) share_floor: the slices publish their lists' scores and a helper
                // warp keeps the k-th best of the union in tau_q, re-read by every slice once per tile;
                // (b) optional sample pass over the first `sample_rows` rows that seeds tau_q before the launch
                // (valid while k fits in a slice list: the floor is the L-th best sampled key).
                crs::GemmFloorArgs fl{};
                const int L = crs::gemm_list_len(k);
                const int64_t sample = ix->sample_rows;
                const bool sampling = sample > 0 && nq > 128 && k <= L && ix->count >= 8 * sample;
                const bool sharing = ix->share_floor != 0;
                if (sampling || sharing) {
                    const int slices_full = crs::gemm_n_slices(ix->count, nq, ix->num_sms, ix->gemm_cluster);
                    const size_t nq_pad = ((size_t)nq + 127) / 128 * 128;
                    const size_t pub_words = sharing ? nq_pad * (size_t)slices_full * L : 0;
                    const size_t prog_words = ix->gemm_lockstep
                        ? (size_t)crs::gemm_progress_words(ix->count, nq, ix->num_sms, ix->gemm_cluster) : 0;
                    const size_t words = nq_pad + pub_words + prog_words;
                    CRS_CUDA(ix->floors.ensure(words));
                    CRS_CUDA(cudaMemsetAsync(ix->floors.p, 0, words * sizeof(uint32_t), st));
                    fl.tau_q = ix->floors.p;
                    fl.pub = sharing ? ix->floors.p + nq_pad : nullptr;
                    fl.progress = prog_words ? ix->floors.p + nq_pad + pub_words : nullptr;
                    fl.lead_tiles = ix->gemm_lockstep;
                    fl.qnorms = ix->qnorms.p;
                    fl.margin_rel = is_float ? 3.0f * fa.eps_rel * ix->row_norm_bound : 0.f;
                    if (is_float) fa.tau_q = fl.tau_q;
                }
                if (sampling) {
                    int s_slices = 0;
                    CRS_CUDA(crs::launch_gemm_topk(st, ix->codes, sample, (int)ix->row_bytes, kind, ix->qcodes.p, nq, k,
                                                   tau_bits, ix->cand.p, ix->num_sms, ix->gemm_cluster, &s_slices, plan.allow));
                    CRS_CUDA(crs::launch_sample_tau(st, ix->cand.p, s_slices, L, nq, is_int ? 1 : 0,
                                                    ix->qnorms.p, fl.margin_rel, fl.tau_q));
                    launches += 2;
                }
                CRS_CUDA(crs::launch_gemm_topk(st, ix->codes, ix->count, (int)ix->row_bytes, kind, ix->qcodes.p, nq, k,
                                               tau_bits, ix->cand.p, ix->num_sms, ix->gemm_cluster, &n_slices, plan.allow,
                                               fl.tau_q ? &fl : nullptr, ix->gemm_prefetch, ix->gemm_warm));
                ++launches;
                fa.n_lists = n_slices;
                fa.list_len = crs::gemm_list_len(k);
                fa.list_stride = 32;
                ix->stats.path = 1; ix->stats.grid = n_slices * ((nq + 127) / 128); ix->stats.list_len = fa.list_len;
            } else {
                for (int q = 0; q < nq; ++q) {
                    const uint8_t* qc = ix->qcodes.p + (size_t)q * ix->row_bytes;
                    uint64_t* cd = ix->cand.p + (size_t)q * n_lists * SM;
                    const int grp = crs::scan_multi_group(ix->store, (int)ix->row_bytes, scan_lpl, nq - q, ix->multi_scan);
                    if (grp > 1) {              // a group of short-row integer queries shares one corpus pass
                        CRS_CUDA(crs::launch_scan_int_multi(st, ix->store, ix->codes, ix->count, (int)ix->row_bytes, ix->dim,
                                                            qc, grp, min_raw, cd, (size_t)n_lists * SM, plan));
                        q += grp - 1;
                        ++launches;
                        continue;
                    }
                    if (is_float)
                        CRS_CUDA(crs::launch_scan_f16(st, ix->codes, ix->count, ix->dim_padded, fa.bf16, qc, tau_pre, cd, plan));
                    else if (ix->store == CRS_I8)
                        CRS_CUDA(crs::launch_scan_i8(st, ix->codes, ix->count, ix->dim_padded, qc, min_raw, cd, plan));
                    else
                        CRS_CUDA(crs::launch_scan_b1(st, ix->codes, ix->count, ix->dim_padded, ix->dim, qc, min_raw, cd, plan));
                    ++launches;
                }
            }
            if (ix->profiling) { CRS_CUDA(cudaEventRecord(ix->evs[ix->ev_count % 32][1], st)); ++ix->ev_count; }

            // ---- finalize: merge lists; float stores rescore + certify; integer keys are exact already and
            // need certification only when a slice list is shorter than k
            fa.mode = is_float ? 0 : 1;
            fa.only_flagged = 0;
            fa.certify_exact = (is_int && fa.list_len < k) ? 1 : 0;
            certifying = is_float || fa.certify_exact;
            CRS_CUDA(crs::launch_finalize(st, fa));
            ++launches;
            if (certifying) {
                if (out_dev || ex != nullptr) {
                    need_exact = true;          // cannot look at the flags without a sync: enqueue the conditional pass
                } else {
                    int32_t nf = 0;             // results and the flag count come back in one sync
                    CRS_CUDA(copy_out());
                    CRS_CUDA(cudaMemcpyAsync(&nf, ix->n_flagged, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
                    CRS_CUDA(cudaStreamSynchronize(st));
                    need_exact = nf > 0;
                    copied = !need_exact;
                    if (copied) unpack_out();
                }
            }
        } else {
            set_flags_kernel<<<(nq + 255) / 256, 256, 0, st>>>(ix->flags.p, nq, 1);
            CRS_CUDA(cudaGetLastError());
            ++launches;
        }
        if (need_exact) {
            // exact pass over the whole shard for the flagged queries (exits at once when none is)
            plan.lpl = lpl;
            CRS_CUDA(crs::launch_exact_scan(st, ix->codes, ix->count, (int)ix->row_bytes, ix->dim, ix->store, ix->qcodes.p, nq,
                                            ix->flags.p, min_similarity, min_raw, ix->cand.p, plan,
                                            certifying ? ix->n_flagged : nullptr));
            fa.mode = 1; fa.only_flagged = 1; fa.certify_exact = 0;
            fa.n_lists = n_lists; fa.list_len = M; fa.list_stride = M;     // the exact scan writes one full list per CTA
            CRS_CUDA(crs::launch_finalize(st, fa));
            launches += 2;
        }
    }

    if (ex != nullptr) {
        // ---- exchange: push the local lists to every peer, wait for theirs, merge (one kernel)
        crs::XchgArgs xa{};
        for (int r = 0; r < ex->world; ++r) {
            xa.peer_slots[r] = reinterpret_cast<uint32_t*>(ex->peers[r]);
            xa.peer_flags[r] = reinterpret_cast<uint32_t*>(ex->peers[r]) + ex->slots_words;
        }
        xa.my_slots = ex->buf; xa.my_flags = ex->buf + ex->slots_words;
        xa.step_word = ex->ctl; xa.err_word = ex->ctl + 1;
        xa.local_ids = d_ids; xa.local_scores = reinterpret_cast<const uint32_t*>(d_scores);
        xa.out_ids = f_ids; xa.out_scores = f_scores; xa.out_counts = f_counts;
        xa.rank = ex->rank; xa.world = ex->world; xa.nq = nq; xa.k = k; xa.max_nq = ex->max_nq; xa.max_k = ex->max_k;
        xa.flag_ctas = ex->flag_ctas; xa.is_int = is_int ? 1 : 0; xa.phase = push_only ? 1 : 0;
        CRS_CUDA(crs::launch_xmerge(st, xa));
        ++launches;
    }
    if (!out_dev && !copied) {
        CRS_CUDA(copy_out());
        CRS_CUDA(cudaStreamSynchronize(st));
        unpack_out();
    }
    ix->stats.kernel_launches = launches;
    ix->stats.searches_total += nq;
    return CRS_OK;
}

int crs_index_search(crs_index* ix, const void* queries, int nq, int k, float min_similarity,
                     uint32_t* out_ids, void* out_scores, int32_t* out_counts) {
    return search_impl(ix, queries, nq, k, min_similarity, nullptr, out_ids, out_scores, out_counts);
}

int crs_index_search_filtered(crs_index* ix, const void* queries, int nq, int k, float min_similarity,
                              const uint32_t* allow_bits, uint32_t* out_ids, void* out_scores,
                              int32_t* out_counts) {
    return search_impl(ix, queries, nq, k, min_similarity, allow_bits, out_ids, out_scores, out_counts);
}

int crs_index_fetch_rows(crs_index* ix, const uint32_t* ids, int n, void* out_codes) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    if (n < 0) return fail(CRS_EINVAL, "n must be >= 0");
    if (n == 0) return CRS_OK;
    if (!ids || !out_codes) return fail(CRS_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t st = ix->stream;
    const bool ids_dev = is_device_ptr(ids), out_dev = is_device_ptr(out_codes);
    DevTmp d_ids, d_out;
    const uint32_t* ids_p = ids; void* out_p = out_codes;
    const size_t bytes = (size_t)n * ix->row_bytes;
    if (!ids_dev) {
        CRS_CUDA(d_ids.alloc((size_t)n * sizeof(uint32_t)));
        CRS_CUDA(cudaMemcpyAsync(d_ids.p, ids, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        ids_p = reinterpret_cast<const uint32_t*>(d_ids.p);
    }
    if (!out_dev) {
        CRS_CUDA(d_out.alloc(bytes));
        CRS_CUDA(cudaMemcpyAsync(d_out.p, out_codes, bytes, cudaMemcpyHostToDevice, st));   // keep rows of other shards
        out_p = d_out.p;
    }
    CRS_CUDA(crs::launch_gather_rows(st, ix->codes, ix->row_bytes, ix->count, ix->row_base, ids_p, n, out_p));
    if (!out_dev) CRS_CUDA(cudaMemcpyAsync(out_codes, d_out.p, bytes, cudaMemcpyDeviceToHost, st));
    if (!ids_dev || !out_dev) CRS_CUDA(cudaStreamSynchronize(st));
    return CRS_OK;
}

int crs_index_score_rows(crs_index* ix, const void* queries, int nq, const uint32_t* ids, int m, void* out_scores) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    if (nq < 0 || m < 0) return fail(CRS_EINVAL, "nq and m must be >= 0");
    if (nq == 0 || m == 0) return CRS_OK;
    if (!queries || !ids || !out_scores) return fail(CRS_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t st = ix->stream;
    const size_t total = (size_t)nq * m;
    const bool q_dev = is_device_ptr(queries), ids_dev = is_device_ptr(ids), out_dev = is_device_ptr(out_scores);
    const float* qd = reinterpret_cast<const float*>(queries);
    if (!q_dev) {
        CRS_CUDA(ix->qsrc.ensure((size_t)nq * ix->dim));
        CRS_CUDA(cudaMemcpyAsync(ix->qsrc.p, queries, (size_t)nq * ix->dim * sizeof(float), cudaMemcpyHostToDevice, st));
        qd = ix->qsrc.p;
    }
    CRS_CUDA(ix->qcodes.ensure((size_t)nq * ix->row_bytes));
    CRS_CUDA(ix->qnorms.ensure((size_t)nq));
    CRS_CUDA(crs::launch_encode(st, qd, nq, ix->dim, ix->dim_padded, ix->store, ix->metric, ix->i8_scale,
                                ix->qcodes.p, ix->qnorms.p));
    const uint32_t* d_ids = ids;
    if (!ids_dev) {
        CRS_CUDA(ix->ids_dev.ensure(total));
        CRS_CUDA(cudaMemcpyAsync(ix->ids_dev.p, ids, total * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        d_ids = ix->ids_dev.p;
    }
    void* d_out = out_scores;
    if (!out_dev) {
        CRS_CUDA(ix->scores_dev.ensure(total * 4));
        d_out = ix->scores_dev.p;
    }
    CRS_CUDA(crs::launch_score_rows(st, ix->codes, ix->count, ix->row_base, (int)ix->row_bytes, ix->dim, ix->store,
                                    ix->qcodes.p, d_ids, nq, m, d_out));
    if (!out_dev) CRS_CUDA(cudaMemcpyAsync(out_scores, d_out, total * 4, cudaMemcpyDeviceToHost, st));
    if (!q_dev || !ids_dev || !out_dev) CRS_CUDA(cudaStreamSynchronize(st));
    return CRS_OK;
}

int crs_index_score_vectors(crs_index* ix, const void* queries, int nq, const void* rows, int m, void* out_scores) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    if (nq < 0 || m < 0) return fail(CRS_EINVAL, "nq and m must be >= 0");
    if (nq == 0 || m == 0) return CRS_OK;
    if (!queries || !rows || !out_scores) return fail(CRS_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t st = ix->stream;
    const size_t total = (size_t)nq * m;
    const bool q_dev = is_device_ptr(queries), r_dev = is_device_ptr(rows), out_dev = is_device_ptr(out_scores);
    const float* qd = reinterpret_cast<const float*>(queries);
    if (!q_dev) {
        CRS_CUDA(ix->qsrc.ensure((size_t)nq * ix->dim));
        CRS_CUDA(cudaMemcpyAsync(ix->qsrc.p, queries, (size_t)nq * ix->dim * sizeof(float), cudaMemcpyHostToDevice, st));
        qd = ix->qsrc.p;
    }
    CRS_CUDA(ix->qcodes.ensure((size_t)nq * ix->row_bytes));
    CRS_CUDA(ix->qnorms.ensure((size_t)nq));
    CRS_CUDA(crs::launch_encode(st, qd, nq, ix->dim, ix->dim_padded, ix->store, ix->metric, ix->i8_scale,
                                ix->qcodes.p, ix->qnorms.p));
    const float* rd = reinterpret_cast<const float*>(rows);
    if (!r_dev) {
        CRS_CUDA(ix->stage_rows.ensure(total * ix->dim * sizeof(float)));
        CRS_CUDA(cudaMemcpyAsync(ix->stage_rows.p, rows, total * ix->dim * sizeof(float), cudaMemcpyHostToDevice, st));
        rd = reinterpret_cast<const float*>(ix->stage_rows.p);
    }
    // the candidate rows go through the same encoder as stored rows, into scratch
    CRS_CUDA(ix->vec_codes.ensure(total * ix->row_bytes));
    CRS_CUDA(crs::launch_encode(st, rd, (int64_t)total, ix->dim, ix->dim_padded, ix->store, ix->metric, ix->i8_scale,
                                ix->vec_codes.p, nullptr));
    void* d_out = out_scores;
    if (!out_dev) {
        CRS_CUDA(ix->scores_dev.ensure(total * 4));
        d_out = ix->scores_dev.p;
    }
    CRS_CUDA(crs::launch_score_rows(st, ix->vec_codes.p, (int64_t)total, 0, (int)ix->row_bytes, ix->dim, ix->store,
                                    ix->qcodes.p, nullptr, nq, m, d_out));
    if (!out_dev) CRS_CUDA(cudaMemcpyAsync(out_scores, d_out, total * 4, cudaMemcpyDeviceToHost, st));
    if (!q_dev || !r_dev || !out_dev) CRS_CUDA(cudaStreamSynchronize(st));
    return CRS_OK;
}

int crs_select_topk(void* cuda_stream, const uint32_t* ids, const void* scores, int is_int, int nq, int m, int k_out,
                    uint32_t* out_ids, void* out_scores, int32_t* out_counts) {
    if (nq < 0 || m <= 0 || k_out <= 0 || k_out > m) return fail(CRS_EINVAL, "bad sizes");
    if (m > crs::kMaxListLen) return fail(CRS_EINVAL, "m > 128 not supported");
    if (nq == 0) return CRS_OK;
    if (!is_device_ptr(ids) || !is_device_ptr(scores) || !is_device_ptr(out_ids) || !is_device_ptr(out_scores) ||
        !is_device_ptr(out_counts))
        return fail(CRS_EINVAL, "crs_select_topk takes device buffers");
    CRS_CUDA(crs::launch_merge_topk(reinterpret_cast<cudaStream_t>(cuda_stream), ids, scores, is_int, 1, nq, m, k_out,
                                    out_ids, out_scores, out_counts, /*sorted_input=*/false));
    return CRS_OK;
}

int crs_mmr(crs_index* ix, const void* vecs, const double* relevance, int nq, int m, int k_out,
            double lambda, int32_t* out_order) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    if (nq < 0 || m < 0 || k_out <= 0) return fail(CRS_EINVAL, "bad sizes");
    if (nq == 0 || m == 0) return CRS_OK;
    if (m > 128) return fail(CRS_EINVAL, "m > 128 candidates not supported");
    if ((size_t)m * (ix->row_bytes + 16) > 200 * 1024)
        return fail(CRS_EINVAL, "candidate set too large for the MMR kernel's shared memory (m * (row_bytes + 16) must stay below 200 KB: "
                                "e.g. m <= 99 for 1024-d fp16 rows)");
    if (!vecs || !relevance || !out_order) return fail(CRS_EINVAL, "NULL buffer");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t st = ix->stream;
    const size_t vbytes = (size_t)nq * m * ix->row_bytes, rbytes = (size_t)nq * m * sizeof(double);
    const size_t obytes = (size_t)nq * k_out * sizeof(int32_t);
    DevTmp d_v, d_r, d_o;
    const void* v = vecs; const double* r = relevance; int32_t* o = out_order;
    if (!is_device_ptr(vecs)) { CRS_CUDA(d_v.alloc(vbytes)); CRS_CUDA(cudaMemcpyAsync(d_v.p, vecs, vbytes, cudaMemcpyHostToDevice, st)); v = d_v.p; }
    if (!is_device_ptr(relevance)) { CRS_CUDA(d_r.alloc(rbytes)); CRS_CUDA(cudaMemcpyAsync(d_r.p, relevance, rbytes, cudaMemcpyHostToDevice, st)); r = reinterpret_cast<const double*>(d_r.p); }
    const bool o_dev = is_device_ptr(out_order);
    if (!o_dev) { CRS_CUDA(d_o.alloc(obytes)); o = reinterpret_cast<int32_t*>(d_o.p); }
    CRS_CUDA(crs::launch_mmr(st, v, ix->store, ix->dim_padded, ix->dim, r, nq, m, k_out, lambda, o));
    if (!o_dev) CRS_CUDA(cudaMemcpyAsync(out_order, d_o.p, obytes, cudaMemcpyDeviceToHost, st));
    if (d_v.p || d_r.p || d_o.p) CRS_CUDA(cudaStreamSynchronize(st));
    return CRS_OK;
}

int crs_mmr_select(crs_index* ix, const void* vecs, const uint32_t* ids, const void* raw_scores, const int32_t* counts,
                   int nq, int m, int k_out, double lambda, uint32_t* out_ids, float* out_sims, double* out_scores,
                   int32_t* out_counts) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    if (nq < 0 || m <= 0 || k_out <= 0) return fail(CRS_EINVAL, "bad sizes");
    if (nq == 0) return CRS_OK;
    if (m > 128) return fail(CRS_EINVAL, "m > 128 candidates not supported");
    if ((size_t)m * (ix->row_bytes + 16) > 200 * 1024)
        return fail(CRS_EINVAL, "candidate set too large for the MMR kernel's shared memory (m * (row_bytes + 16) must stay below 200 KB)");
    const void* ptrs[] = {vecs, ids, raw_scores, counts, out_ids, out_sims, out_scores, out_counts};
    for (const void* p : ptrs)
        if (!is_device_ptr(p)) return fail(CRS_EINVAL, "crs_mmr_select takes device buffers");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    crs::MmrFusedArgs fa{};
    fa.ids = ids; fa.raw = raw_scores; fa.counts = counts;
    fa.sim_scale = (float)(((double)ix->i8_scale / 127.0) * ((double)ix->i8_scale / 127.0));
    fa.out_ids = out_ids; fa.out_sims = out_sims; fa.out_rel = out_scores; fa.out_counts = out_counts;
    CRS_CUDA(crs::launch_mmr(ix->stream, vecs, ix->store, ix->dim_padded, ix->dim, nullptr, nq, m, k_out, lambda, nullptr, &fa));
    return CRS_OK;
}

int crs_merge_topk(void* cuda_stream, const uint32_t* ids, const void* scores, int is_int,
                   int n_lists, int nq, int k_in, int k_out,
                   uint32_t* out_ids, void* out_scores, int32_t* out_counts) {
    if (nq < 0 || n_lists <= 0 || k_in <= 0 || k_out <= 0 || k_out > k_in)
        return fail(CRS_EINVAL, "bad sizes");
    if (k_in > crs::kMaxListLen) return fail(CRS_EINVAL, "k_in > 128 not supported");
    if (nq == 0) return CRS_OK;
    if (!is_device_ptr(ids) || !is_device_ptr(scores) || !is_device_ptr(out_ids) || !is_device_ptr(out_scores) ||
        !is_device_ptr(out_counts))
        return fail(CRS_EINVAL, "crs_merge_topk takes device buffers");
    CRS_CUDA(crs::launch_merge_topk(reinterpret_cast<cudaStream_t>(cuda_stream), ids, scores, is_int, n_lists, nq,
                                    k_in, k_out, out_ids, out_scores, out_counts));
    return CRS_OK;
}

int crs_merge_topk_strided(void* cuda_stream, const uint32_t* ids, const void* scores, int is_int,
                           int n_lists, int nq, int k_in, int k_out, int64_t list_stride,
                           uint32_t* out_ids, void* out_scores, int32_t* out_counts) {
    if (nq < 0 || n_lists <= 0 || k_in <= 0 || k_out <= 0 || k_out > k_in || list_stride < (int64_t)nq * k_in)
        return fail(CRS_EINVAL, "bad sizes");
    if (k_in > crs::kMaxListLen) return fail(CRS_EINVAL, "k_in > 128 not supported");
    if (nq == 0) return CRS_OK;
    if (!is_device_ptr(ids) || !is_device_ptr(scores) || !is_device_ptr(out_ids) || !is_device_ptr(out_scores) ||
        !is_device_ptr(out_counts))
        return fail(CRS_EINVAL, "crs_merge_topk_strided takes device buffers");
    CRS_CUDA(crs::launch_merge_topk(reinterpret_cast<cudaStream_t>(cuda_stream), ids, scores, is_int, n_lists, nq,
                                    k_in, k_out, out_ids, out_scores, out_counts, true, (size_t)list_stride));
    return CRS_OK;
}

// ---- cross-shard exchange over peer memory ------------------------------------------------
int crs_exchange_create(crs_exchange** out, int device, int rank, int world, int max_nq, int max_k) {
    if (!out) return fail(CRS_EINVAL, "out is NULL");
    *out = nullptr;
    if (world < 1 || world > crs::kMaxWorld || rank < 0 || rank >= world) return fail(CRS_EINVAL, "bad rank / world");
    if (max_nq <= 0 || max_k <= 0 || max_k > crs::kMaxListLen) return fail(CRS_EINVAL, "bad max_nq / max_k");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(CRS_ECUDA, "no CUDA device"); }
    if (device < 0 || device >= ndev) return fail(CRS_EINVAL, "device ordinal out of range");
    DeviceGuard g(device);
    crs_exchange* ex = new crs_exchange();
    ex->device = device; ex->rank = rank; ex->world = world; ex->max_nq = max_nq; ex->max_k = max_k;
    ex->flag_ctas = crs::xmerge_ctas(max_nq);
    const size_t cap = (size_t)max_nq * max_k;
    ex->slots_words = 2 * (size_t)world * 2 * cap;
    ex->flags_words = 2 * (size_t)world * ex->flag_ctas;
    const size_t buf_bytes = (ex->slots_words + ex->flags_words) * sizeof(uint32_t);
    cudaError_t e = cudaMalloc(&ex->buf, buf_bytes);
    if (e == cudaSuccess) e = cudaMemset(ex->buf, 0, buf_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&ex->ctl, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(ex->ctl, 0, 2 * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&ex->local, (2 * cap + max_nq) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (ex->buf) cudaFree(ex->buf);
        if (ex->ctl) cudaFree(ex->ctl);
        if (ex->local) cudaFree(ex->local);
        delete ex;
        return cuda_fail(e, "crs_exchange_create");
    }
    ex->peers[rank] = ex->buf;
    ex->wired = world == 1;
    *out = ex;
    return CRS_OK;
}

int crs_exchange_destroy(crs_exchange* ex) {
    if (!ex) return CRS_OK;
    {
        DeviceGuard g(ex->device);
        cudaDeviceSynchronize();
        for (int r = 0; r < ex->world; ++r) {
            if (ex->ipc_opened[r] && ex->peers[r]) cudaIpcCloseMemHandle(ex->peers[r]);
            if (ex->shard_ipc[r] && ex->shards.codes[r]) cudaIpcCloseMemHandle(const_cast<uint8_t*>(ex->shards.codes[r]));
        }
        ex->gather.release();
        if (ex->buf) cudaFree(ex->buf);
        if (ex->ctl) cudaFree(ex->ctl);
        if (ex->local) cudaFree(ex->local);
        cudaGetLastError();
    }
    delete ex;
    return CRS_OK;
}

int crs_exchange_ipc_handle(crs_exchange* ex, void* out_handle64) {
    if (!ex || !out_handle64) return fail(CRS_EINVAL, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    DeviceGuard g(ex->device);
    cudaIpcMemHandle_t h;
    CRS_CUDA(cudaIpcGetMemHandle(&h, ex->buf));
    memcpy(out_handle64, &h, sizeof(h));
    return CRS_OK;
}

int crs_exchange_open_peers(crs_exchange* ex, const void* handles) {
    if (!ex || !handles) return fail(CRS_EINVAL, "bad argument");
    DeviceGuard g(ex->device);
    for (int r = 0; r < ex->world; ++r) {
        if (r == ex->rank || ex->peers[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, reinterpret_cast<const uint8_t*>(handles) + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        CRS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ex->peers[r] = p;
        ex->ipc_opened[r] = true;
    }
    ex->wired = true;
    return CRS_OK;
}

int crs_exchange_buffer(crs_exchange* ex, void** out_ptr) {
    if (!ex || !out_ptr) return fail(CRS_EINVAL, "bad argument");
    *out_ptr = ex->buf;
    return CRS_OK;
}

int crs_exchange_set_peer_buffers(crs_exchange* ex, void* const* bufs) {
    if (!ex || !bufs) return fail(CRS_EINVAL, "bad argument");
    DeviceGuard g(ex->device);
    for (int r = 0; r < ex->world; ++r) {
        if (r == ex->rank) continue;
        if (!bufs[r]) return fail(CRS_EINVAL, "NULL peer buffer");
        cudaPointerAttributes a;
        CRS_CUDA(cudaPointerGetAttributes(&a, bufs[r]));
        if (a.type != cudaMemoryTypeDevice) return fail(CRS_EINVAL, "peer buffer is not device memory");
        if (a.device != ex->device) {
            int can = 0;
            CRS_CUDA(cudaDeviceCanAccessPeer(&can, ex->device, a.device));
            if (!can) return fail(CRS_ECUDA, "devices cannot access each other's memory (no P2P)");
            cudaError_t e = cudaDeviceEnablePeerAccess(a.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
        }
        ex->peers[r] = bufs[r];
    }
    ex->wired = true;
    return CRS_OK;
}

int crs_exchange_status(crs_exchange* ex, void* cuda_stream, int* timed_out, uint32_t* step) {
    if (!ex) return fail(CRS_EINVAL, "exchange is NULL");
    DeviceGuard g(ex->device);
    uint32_t w[2] = {0, 0};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    CRS_CUDA(cudaMemcpyAsync(w, ex->ctl, sizeof(w), cudaMemcpyDeviceToHost, st));
    CRS_CUDA(cudaStreamSynchronize(st));
    if (step) *step = w[0];
    if (timed_out) *timed_out = (int)w[1];
    return CRS_OK;
}

// ---- peer access to the shards' stored rows (candidate vectors for MMR / rescoring without a second collective)
int crs_index_codes_handle(crs_index* ix, void* out_handle64, void** out_ptr, uint32_t* row_base, int64_t* count) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    if (out_handle64) {
        if (!ix->codes) return fail(CRS_ESTATE, "the index holds no rows yet");
        cudaIpcMemHandle_t h;
        CRS_CUDA(cudaIpcGetMemHandle(&h, ix->codes));
        memcpy(out_handle64, &h, sizeof(h));
    }
    if (out_ptr) *out_ptr = ix->codes;
    if (row_base) *row_base = ix->row_base;
    if (count) *count = ix->count;
    return CRS_OK;
}

static int set_shards(crs_exchange* ex, crs_index* own, const void* handles, void* const* ptrs,
                      const uint32_t* row_bases, const int64_t* counts) {
    if (!ex || !own || !row_bases || !counts || (!handles && !ptrs)) return fail(CRS_EINVAL, "bad argument");
    if (own->device != ex->device) return fail(CRS_EINVAL, "exchange and index live on different devices");
    DeviceGuard g(ex->device);
    for (int r = 0; r < ex->world; ++r) {
        if (ex->shard_ipc[r] && ex->shards.codes[r]) cudaIpcCloseMemHandle(const_cast<uint8_t*>(ex->shards.codes[r]));
        ex->shard_ipc[r] = false;
        ex->shards.codes[r] = nullptr;
    }
    for (int r = 0; r < ex->world; ++r) {
        ex->shards.row_base[r] = row_bases[r];
        ex->shards.count[r] = counts[r];
        if (r == ex->rank) { ex->shards.codes[r] = own->codes; continue; }
        if (counts[r] == 0) continue;
        if (handles) {
            cudaIpcMemHandle_t h;
            memcpy(&h, reinterpret_cast<const uint8_t*>(handles) + (size_t)r * sizeof(h), sizeof(h));
            void* p = nullptr;
            CRS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            ex->shards.codes[r] = reinterpret_cast<const uint8_t*>(p);
            ex->shard_ipc[r] = true;
        } else {
            cudaPointerAttributes a;
            CRS_CUDA(cudaPointerGetAttributes(&a, ptrs[r]));
            if (a.type != cudaMemoryTypeDevice) return fail(CRS_EINVAL, "shard pointer is not device memory");
            if (a.device != ex->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(a.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
                cudaGetLastError();
            }
            ex->shards.codes[r] = reinterpret_cast<const uint8_t*>(ptrs[r]);
        }
    }
    ex->shards.world = ex->world;
    ex->shards.row_bytes = (int)own->row_bytes;
    return CRS_OK;
}

int crs_exchange_open_shards(crs_exchange* ex, crs_index* own, const void* handles, const uint32_t* row_bases,
                             const int64_t* counts) {
    return set_shards(ex, own, handles, nullptr, row_bases, counts);
}

int crs_exchange_set_shards(crs_exchange* ex, crs_index* own, void* const* codes_ptrs, const uint32_t* row_bases,
                            const int64_t* counts) {
    return set_shards(ex, own, nullptr, codes_ptrs, row_bases, counts);
}

int crs_exchange_fetch_rows(crs_exchange* ex, void* cuda_stream, const uint32_t* ids, int n, void* out_codes) {
    if (!ex) return fail(CRS_EINVAL, "exchange is NULL");
    if (ex->shards.world == 0) return fail(CRS_ESTATE, "no shards registered with this exchange");
    if (n < 0) return fail(CRS_EINVAL, "n must be >= 0");
    if (n == 0) return CRS_OK;
    if (!is_device_ptr(ids) || !is_device_ptr(out_codes)) return fail(CRS_EINVAL, "crs_exchange_fetch_rows takes device buffers");
    DeviceGuard g(ex->device);
    CRS_CUDA(crs::launch_peer_gather(reinterpret_cast<cudaStream_t>(cuda_stream), ex->shards, ids, n, out_codes));
    return CRS_OK;
}

int crs_exchange_score_rows(crs_index* ix, crs_exchange* ex, const void* queries, int nq, const uint32_t* ids, int m,
                            void* out_scores) {
    if (!ix || !ex) return fail(CRS_EINVAL, "bad argument");
    if (ex->shards.world == 0) return fail(CRS_ESTATE, "no shards registered with this exchange");
    if (nq < 0 || m < 0) return fail(CRS_EINVAL, "nq and m must be >= 0");
    if (nq == 0 || m == 0) return CRS_OK;
    if (!is_device_ptr(queries) || !is_device_ptr(ids) || !is_device_ptr(out_scores))
        return fail(CRS_EINVAL, "crs_exchange_score_rows takes device buffers");
    if ((int)ix->row_bytes != ex->shards.row_bytes) return fail(CRS_EINVAL, "index and registered shards differ in row width");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t st = ix->stream;
    const size_t total = (size_t)nq * m;
    CRS_CUDA(ix->qcodes.ensure((size_t)nq * ix->row_bytes));
    CRS_CUDA(ix->qnorms.ensure((size_t)nq));
    CRS_CUDA(crs::launch_encode(st, reinterpret_cast<const float*>(queries), nq, ix->dim, ix->dim_padded, ix->store, ix->metric,
                                ix->i8_scale, ix->qcodes.p, ix->qnorms.p));
    CRS_CUDA(ex->gather.ensure(total * ix->row_bytes));
    CRS_CUDA(crs::launch_peer_gather(st, ex->shards, ids, (int)total, ex->gather.p));
    // pair i scores gathered row i (ids = NULL form of K8); pad ids were gathered as zero rows and are masked by the caller
    CRS_CUDA(crs::launch_score_rows(st, ex->gather.p, (int64_t)total, 0, (int)ix->row_bytes, ix->dim, ix->store,
                                    ix->qcodes.p, nullptr, nq, m, out_scores));
    return CRS_OK;
}

int crs_index_search_sharded(crs_index* ix, crs_exchange* ex, const void* queries, int nq, int k, float min_similarity,
                             uint32_t* out_ids, void* out_scores, int32_t* out_counts) {
    if (!ex) return fail(CRS_EINVAL, "exchange is NULL");
    return search_impl(ix, queries, nq, k, min_similarity, nullptr, out_ids, out_scores, out_counts, ex, 0);
}

int crs_index_search_push(crs_index* ix, crs_exchange* ex, const void* queries, int nq, int k, float min_similarity,
                          const uint32_t* allow_bits) {
    if (!ex) return fail(CRS_EINVAL, "exchange is NULL");
    return search_impl(ix, queries, nq, k, min_similarity, allow_bits, nullptr, nullptr, nullptr, ex, 1);
}

int crs_exchange_merge(crs_exchange* ex, void* cuda_stream, int nq, int k, int is_int,
                       uint32_t* out_ids, void* out_scores, int32_t* out_counts) {
    if (!ex) return fail(CRS_EINVAL, "exchange is NULL");
    if (!ex->wired) return fail(CRS_ESTATE, "exchange is not wired to its peers yet");
    if (nq < 0 || k <= 0 || nq > ex->max_nq || k > ex->max_k) return fail(CRS_EINVAL, "bad sizes");
    if (nq == 0) return CRS_OK;
    if (!is_device_ptr(out_ids) || !is_device_ptr(out_scores) || !is_device_ptr(out_counts))
        return fail(CRS_EINVAL, "crs_exchange_merge takes device buffers");
    DeviceGuard g(ex->device);
    crs::XchgArgs xa{};
    for (int r = 0; r < ex->world; ++r) {
        xa.peer_slots[r] = reinterpret_cast<uint32_t*>(ex->peers[r]);
        xa.peer_flags[r] = reinterpret_cast<uint32_t*>(ex->peers[r]) + ex->slots_words;
    }
    xa.my_slots = ex->buf; xa.my_flags = ex->buf + ex->slots_words;
    xa.step_word = ex->ctl; xa.err_word = ex->ctl + 1;
    xa.out_ids = out_ids; xa.out_scores = out_scores; xa.out_counts = out_counts;
    xa.rank = ex->rank; xa.world = ex->world; xa.nq = nq; xa.k = k; xa.max_nq = ex->max_nq; xa.max_k = ex->max_k;
    xa.flag_ctas = ex->flag_ctas; xa.is_int = is_int ? 1 : 0; xa.phase = 2;
    CRS_CUDA(crs::launch_xmerge(reinterpret_cast<cudaStream_t>(cuda_stream), xa));
    return CRS_OK;
}

// ---- persistence: 64-byte header + raw stored rows -------------------------------------
struct CrsFileHeader {
    char magic[8];          // "CRSIDX1\0"
    int32_t dim, dim_padded, store, metric;
    int64_t count;
    float i8_scale, row_norm_bound;
    uint8_t pad[24];
};
static_assert(sizeof(CrsFileHeader) == 64, "header is 64 bytes");

// rows [from_row, count) of the index -> f at its current position
static int write_rows(crs_index* ix, FILE* f, int64_t from_row) {
    const size_t chunk_rows = std::max<size_t>(1, (64u << 20) / ix->row_bytes);
    std::vector<uint8_t> buf(chunk_rows * ix->row_bytes);
    for (int64_t off = from_row; off < ix->count; off += (int64_t)chunk_rows) {
        const size_t m = (size_t)std::min<int64_t>((int64_t)chunk_rows, ix->count - off);
        cudaError_t e = cudaMemcpyAsync(buf.data(), ix->codes + (size_t)off * ix->row_bytes, m * ix->row_bytes,
                                        cudaMemcpyDeviceToHost, ix->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
        if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy");
        if (fwrite(buf.data(), ix->row_bytes, m, f) != m) return fail(CRS_EIO, "write failed");
    }
    return CRS_OK;
}

static bool flush_to_disk(FILE* f) { return fflush(f) == 0 && fsync(fileno(f)) == 0; }

static void fill_header(const crs_index* ix, CrsFileHeader* h) {
    memset(h, 0, sizeof(*h));
    memcpy(h->magic, "CRSIDX1", 8);
    h->dim = ix->dim; h->dim_padded = ix->dim_padded; h->store = ix->store; h->metric = ix->metric;
    h->count = ix->count; h->i8_scale = ix->i8_scale; h->row_norm_bound = ix->row_norm_bound;
}

// Whole index -> path, atomically: written to "<path>.tmp", flushed to disk, renamed over the old file.
// A crash leaves either the old file or the new one, never a truncated mix.
int crs_index_save(crs_index* ix, const char* path) {
    if (!ix || !path) return fail(CRS_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(CRS_EIO, std::string("cannot open ") + tmp);
    CrsFileHeader h;
    fill_header(ix, &h);
    int rc = fwrite(&h, sizeof(h), 1, f) == 1 ? CRS_OK : fail(CRS_EIO, "write failed");
    if (rc == CRS_OK) rc = write_rows(ix, f, 0);
    if (rc == CRS_OK && !flush_to_disk(f)) rc = fail(CRS_EIO, std::string("flush failed: ") + tmp);
    if (fclose(f) != 0 && rc == CRS_OK) rc = fail(CRS_EIO, std::string("close failed: ") + tmp);
    if (rc == CRS_OK && rename(tmp.c_str(), path) != 0) rc = fail(CRS_EIO, std::string("rename failed: ") + path);
    if (rc != CRS_OK) remove(tmp.c_str());
    return rc;
}

// Appends the rows the file does not hold yet (incremental indexing: O(new rows), not O(all rows) per add).
// Order: new rows are written behind the old ones and flushed, THEN the header's row count is advanced and
// flushed — a crash in between leaves a file whose header still describes a consistent prefix.
int crs_index_append(crs_index* ix, const char* path) {
    if (!ix || !path) return fail(CRS_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(ix->mu);
    DeviceGuard g(ix->device);
    FILE* f = fopen(path, "r+b");
    if (!f) return fail(CRS_EIO, std::string("cannot open ") + path);
    CrsFileHeader h{};
    int rc = CRS_OK;
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "CRSIDX1", 8) != 0) rc = fail(CRS_EIO, "not a CRS index file");
    if (rc == CRS_OK && (h.dim != ix->dim || h.dim_padded != ix->dim_padded || h.store != (int32_t)ix->store ||
                         h.metric != (int32_t)ix->metric))
        rc = fail(CRS_EIO, "file holds an index of another shape");
    if (rc == CRS_OK && (h.count < 0 || h.count > ix->count)) rc = fail(CRS_EIO, "file holds more rows than the index");
    if (rc == CRS_OK && fseeko(f, (off_t)sizeof(h) + (off_t)h.count * (off_t)ix->row_bytes, SEEK_SET) != 0)
        rc = fail(CRS_EIO, "seek failed");
    if (rc == CRS_OK) rc = write_rows(ix, f, h.count);
    if (rc == CRS_OK && !flush_to_disk(f)) rc = fail(CRS_EIO, "flush failed");
    if (rc == CRS_OK) {
        fill_header(ix, &h);
        if (fseeko(f, 0, SEEK_SET) != 0 || fwrite(&h, sizeof(h), 1, f) != 1 || !flush_to_disk(f))
            rc = fail(CRS_EIO, "header update failed");
    }
    fclose(f);
    return rc;
}

// Drops the rows from new_count on (a host whose sidecar holds fewer rows than the blob after a crash).
int crs_index_truncate(crs_index* ix, int64_t new_count) {
    if (!ix) return fail(CRS_EINVAL, "index is NULL");
    std::lock_guard<std::mutex> lk(ix->mu);
    if (new_count < 0 || new_count > ix->count) return fail(CRS_EINVAL, "new_count out of range");
    ix->count = new_count;
    return CRS_OK;
}

int crs_index_load(crs_index** out, const char* path, int device, uint32_t row_base) {
    if (!out || !path) return fail(CRS_EINVAL, "bad argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(CRS_EIO, std::string("cannot open ") + path);
    CrsFileHeader h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "CRSIDX1", 8) != 0) {
        fclose(f);
        return fail(CRS_EIO, "not a CRS index file");
    }
    // the header is untrusted input: every field is checked against what this library can hold and against
    // the size of the file (a row count beyond the rows actually present is clamped: torn append)
    if (h.dim <= 0 || h.count < 0 || (h.store != CRS_F16 && h.store != CRS_BF16 && h.store != CRS_I8 && h.store != CRS_B1) ||
        (h.metric != CRS_COSINE && h.metric != CRS_IP) || !(h.i8_scale > 0.f) || !(h.row_norm_bound > 0.f)) {
        fclose(f);
        return fail(CRS_EIO, "corrupt index header");
    }
    const int dp = padded_dim_for(h.dim, (crs_dtype)h.store);
    if (dp < 0 || dp != h.dim_padded) { fclose(f); return fail(CRS_EIO, "layout mismatch"); }
    const size_t rb = row_bytes_for(dp, (crs_dtype)h.store);
    int64_t rows_in_file = 0;
    if (fseeko(f, 0, SEEK_END) == 0) rows_in_file = ((int64_t)ftello(f) - (int64_t)sizeof(h)) / (int64_t)rb;
    if (fseeko(f, (off_t)sizeof(h), SEEK_SET) != 0) { fclose(f); return fail(CRS_EIO, "seek failed"); }
    const int64_t count = std::max<int64_t>(0, std::min<int64_t>(h.count, rows_in_file));
    if ((uint64_t)row_base + (uint64_t)count >= 0xFFFFFFFFull) { fclose(f); return fail(CRS_EIO, "row ids would overflow uint32"); }
    crs_index* ix = nullptr;
    int rc = crs_index_create(&ix, h.dim, (crs_dtype)h.store, (crs_metric)h.metric, device, row_base, count);
    if (rc != CRS_OK) { fclose(f); return rc; }
    ix->i8_scale = h.i8_scale; ix->row_norm_bound = h.row_norm_bound;
    DeviceGuard g(device);
    const size_t chunk_rows = std::max<size_t>(1, (64u << 20) / ix->row_bytes);
    std::vector<uint8_t> buf(chunk_rows * ix->row_bytes);
    for (int64_t off = 0; off < count; off += (int64_t)chunk_rows) {
        const size_t m = (size_t)std::min<int64_t>((int64_t)chunk_rows, count - off);
        if (fread(buf.data(), ix->row_bytes, m, f) != m) { fclose(f); crs_index_destroy(ix); return fail(CRS_EIO, "short read"); }
        cudaError_t e = cudaMemcpy(ix->codes + (size_t)off * ix->row_bytes, buf.data(), m * ix->row_bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { fclose(f); crs_index_destroy(ix); return cuda_fail(e, "cudaMemcpy"); }
    }
    fclose(f);
    ix->count = count;
    *out = ix;
    return CRS_OK;
}

#pragma GCC visibility pop
}  // extern "C"
