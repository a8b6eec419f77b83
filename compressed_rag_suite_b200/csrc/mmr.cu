// K6 — greedy MMR (diversity) rerank over the m candidate vectors of each query.
//
// Replaces ContextRetriever._apply_diversity (reference rag/retrieval.py:219-277).
// The reference re-embeds the surviving chunk texts (:238-239); here the stored
// vectors of the candidates are used instead.  Semantics kept,
// line by line (restated in oracle/postprocess.py: pairwise_sims_f32 / mmr_order):
//   * first pick is position 0 (:242); every later pick is the first position, in
//     ascending order, whose mmr score is strictly greater than the running best
//     (:246-268); the loop runs until k_out positions are picked (the reference
//     runs to k_out = m and never drops an item).
//   * sim(a,b) = dot / (|a| * |b|) evaluated in fp32 (:258-260): dot and the two
//     squared norms are the exactly rounded fl32 of an fp64 sequential sum; sqrt,
//     product and quotient are IEEE fp32 operations.
//   * max_sim starts as the Python float 0.0 and becomes an np.float32 once some
//     sim > it (:255,261); numpy-2 typing then makes
//     lambda*rel - (1-lambda)*max_sim an fp32 expression, otherwise it stays fp64
//     (:264).  Mixed comparisons at :266 round the fp64 side to fp32.  Both are
//     reproduced: every value carries an "is fp32" flag.
// One CTA per query, one thread per candidate, vectors staged in padded shared
// memory (row stride + 16 B keeps the per-thread 16-byte walks conflict-free).
#include "common.cuh"
#include "crs_internal.h"

namespace crs {

constexpr int kMmrThreads = 128;      // m <= 128

template <int STORE>
__device__ __forceinline__ double mmr_dot(const uint8_t* a, const uint8_t* b, int row_bytes, int dim) {
    // sequential over stored elements j = 0..; exact products, fp64 accumulate
    double acc = 0.0;
    if constexpr (STORE == CRS_F16 || STORE == CRS_BF16) {
        for (int c = 0; c < row_bytes; c += 16) {
            const uint4 va = *reinterpret_cast<const uint4*>(a + c);
            const uint4 vb = *reinterpret_cast<const uint4*>(b + c);
            const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float2 fa, fb;
                if constexpr (STORE == CRS_BF16) {
                    fa = make_float2(__uint_as_float(wa[i] << 16), __uint_as_float(wa[i] & 0xFFFF0000u));
                    fb = make_float2(__uint_as_float(wb[i] << 16), __uint_as_float(wb[i] & 0xFFFF0000u));
                } else {
                    fa = __half22float2(*reinterpret_cast<const __half2*>(&wa[i]));
                    fb = __half22float2(*reinterpret_cast<const __half2*>(&wb[i]));
                }
                acc = fma((double)fa.x, (double)fb.x, acc);
                acc = fma((double)fa.y, (double)fb.y, acc);
            }
        }
    } else if constexpr (STORE == CRS_I8) {
        int s = 0;
        for (int c = 0; c < row_bytes; c += 16) {
            const uint4 va = *reinterpret_cast<const uint4*>(a + c);
            const uint4 vb = *reinterpret_cast<const uint4*>(b + c);
            s = __dp4a((int)va.x, (int)vb.x, s); s = __dp4a((int)va.y, (int)vb.y, s);
            s = __dp4a((int)va.z, (int)vb.z, s); s = __dp4a((int)va.w, (int)vb.w, s);
        }
        acc = (double)s;
    } else {   // B1: +-1 vectors over the real dims; pad bits are equal (0) in both rows
        int h = 0;
        for (int c = 0; c < row_bytes; c += 16) {
            const uint4 va = *reinterpret_cast<const uint4*>(a + c);
            const uint4 vb = *reinterpret_cast<const uint4*>(b + c);
            h += __popc(va.x ^ vb.x) + __popc(va.y ^ vb.y) + __popc(va.z ^ vb.z) + __popc(va.w ^ vb.w);
        }
        acc = (double)(dim - 2 * h);
    }
    return acc;
}

// numpy-2 comparison a > b for scalars that are np.float32 (f32 flag) or Python floats
__device__ __forceinline__ bool nep50_gt(double a, bool a32, double b, bool b32) {
    if (!a32 && !b32) return a > b;
    return (float)a > (float)b;
}

// The reference's `score` of a hit from the raw search score, in the arithmetic Python uses
// (rag/indexing.py:171-176 Chroma distance 1 - sim; rag/retrieval.py:75-77): every operation is an
// explicitly rounded IEEE double op, so the compiler cannot contract them into FMAs.
template <int STORE>
__device__ __forceinline__ float raw_to_similarity(const void* raw, size_t i, float sim_scale, int dim) {
    if constexpr (STORE == CRS_F16 || STORE == CRS_BF16) return reinterpret_cast<const float*>(raw)[i];
    else if constexpr (STORE == CRS_I8) return __fmul_rn((float)reinterpret_cast<const int32_t*>(raw)[i], sim_scale);
    else return (float)__ddiv_rn((double)reinterpret_cast<const int32_t*>(raw)[i], (double)dim);
}
__device__ __forceinline__ double reference_relevance(float sim) {
    double d = (double)__fsub_rn(1.0f, sim);        // Chroma's distance is a float32 (1.0f - cos), widened by Python
    d = fmax(0.0, fmin(2.0, d));
    const double s = __dsub_rn(1.0, __ddiv_rn(__dmul_rn(d, d), 2.0));
    return fmax(0.0, fmin(1.0, s));
}

// Optional fused form (search -> MMR in one launch, BASELINE config 4): instead of a relevance
// array the kernel gets the search output (ids, raw scores, counts), derives similarity and the
// reference's relevance itself, and writes the selected ids / similarities / relevances.
struct MmrFused {
    const uint32_t* ids;       // [nq, m] global row ids of the candidates (nullptr = plain form)
    const void* raw;           // [nq, m] raw scores (f32 | i32)
    const int32_t* counts;     // [nq] valid candidates per query
    float sim_scale;           // I8: (a/127)^2 as fp32
    uint32_t* out_ids;         // [nq, k_out], pad 0xFFFFFFFF
    float* out_sims;           // [nq, k_out], pad -inf
    double* out_rel;           // [nq, k_out], pad -inf
    int32_t* out_counts;       // [nq]
};

template <int STORE>
__global__ void __launch_bounds__(kMmrThreads)
mmr_kernel(const uint8_t* __restrict__ vecs, int row_bytes, int dim, const double* __restrict__ relevance,
           int m, int k_out, double lambda, int32_t* __restrict__ out_order, MmrFused fz) {
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ double s_val[kMmrThreads];
    __shared__ uint8_t s_is32[kMmrThreads];
    __shared__ uint8_t s_taken[kMmrThreads];
    __shared__ float s_norm[kMmrThreads];
    __shared__ int s_last;
    const int q = blockIdx.x, t = threadIdx.x;
    const int stride = row_bytes + 16;
    // stage the candidate vectors
    const uint4* src = reinterpret_cast<const uint4*>(vecs + (size_t)q * m * row_bytes);
    const int chunks = row_bytes / 16;
    for (int i = t; i < m * chunks; i += kMmrThreads) {
        const int r = i / chunks, c = i % chunks;
        *reinterpret_cast<uint4*>(sm + (size_t)r * stride + c * 16) = src[i];
    }
    if (t < kMmrThreads) s_taken[t] = 0;
    __syncthreads();

    const uint8_t* mine = sm + (size_t)t * stride;
    float norm = 0.f;
    double rel = 0.0;
    float my_sim = -INFINITY;
    if (t < m) {
        norm = __fsqrt_rn((float)mmr_dot<STORE>(mine, mine, row_bytes, dim));
        if (fz.ids != nullptr) {
            rel = -INFINITY;                                   // padding candidates can never be picked
            if (t < fz.counts[q]) {
                my_sim = raw_to_similarity<STORE>(fz.raw, (size_t)q * m + t, fz.sim_scale, dim);
                rel = reference_relevance(my_sim);
            }
        } else {
            rel = relevance[(size_t)q * m + t];
        }
        s_norm[t] = norm;
    }
    float max_sim = 0.f;
    bool max32 = false;              // False: still the Python float 0.0
    const double t1 = lambda * rel;                 // Python float product (:264)
    const float c32 = (float)(1.0 - lambda);        // (1-lambda) rounded when it meets an np.float32

    __shared__ int s_order[kMmrThreads];
    __shared__ int s_picked;
    if (t == 0) { s_last = 0; s_taken[0] = 1; s_order[0] = 0; s_picked = 1; }
    __syncthreads();
    const int picks = min(k_out, m);
    for (int step = 1; step < picks; ++step) {
        const int last = s_last;
        if (t < m && !s_taken[t]) {
            const uint8_t* other = sm + (size_t)last * stride;
            const float d32 = (float)mmr_dot<STORE>(mine, other, row_bytes, dim);
            const float n_last = s_norm[last];
            const float sim = __fdiv_rn(d32, __fmul_rn(norm, n_last));
            // max(max_sim, sim): returns sim iff sim > max_sim (numpy fp32 compare either way)
            if (sim > max_sim) { max_sim = sim; max32 = true; }
            if (max32) {
                s_val[t] = (double)__fsub_rn((float)t1, __fmul_rn(c32, max_sim));
                s_is32[t] = 1;
            } else {
                s_val[t] = t1 - (1.0 - lambda) * 0.0;
                s_is32[t] = 0;
            }
        }
        __syncthreads();
        if (t == 0) {
            int best = -1; double bv = -INFINITY; bool b32 = false;
            for (int i = 1; i < m; ++i) {
                if (s_taken[i]) continue;
                if (nep50_gt(s_val[i], s_is32[i] != 0, bv, b32)) { bv = s_val[i]; b32 = s_is32[i] != 0; best = i; }
            }
            s_last = best;
            if (best >= 0) { s_taken[best] = 1; s_order[step] = best; s_picked = step + 1; }
        }
        __syncthreads();
        if (s_last < 0) break;       // nothing comparable left (NaNs / padding): stop like the reference's `break`
    }
    __syncthreads();
    // output: the greedy order (plain form) or the selected hits themselves (fused form)
    __shared__ float s_sim[kMmrThreads];
    __shared__ double s_rel[kMmrThreads];
    if (fz.ids != nullptr && t < m) { s_sim[t] = my_sim; s_rel[t] = rel; }
    __syncthreads();
    int picked = s_picked;
    if (fz.ids != nullptr && fz.counts[q] == 0) picked = 0;    // position 0 is a pad too: nothing to return
    for (int r = t; r < k_out; r += kMmrThreads) {
        const int p = (r < picked) ? s_order[r] : -1;
        if (out_order != nullptr) out_order[(size_t)q * k_out + r] = p;
        if (fz.ids != nullptr) {
            fz.out_ids[(size_t)q * k_out + r] = p >= 0 ? fz.ids[(size_t)q * m + p] : CRS_PAD_ID;
            fz.out_sims[(size_t)q * k_out + r] = p >= 0 ? s_sim[p] : -INFINITY;
            fz.out_rel[(size_t)q * k_out + r] = p >= 0 ? s_rel[p] : -INFINITY;
        }
    }
    if (fz.ids != nullptr && t == 0) fz.out_counts[q] = picked;
}

cudaError_t launch_mmr(cudaStream_t st, const void* vecs, crs_dtype store, int dim_padded, int dim,
                       const double* relevance, int nq, int m, int k_out, double lambda, int32_t* out_order,
                       const MmrFusedArgs* fused) {
    MmrFused fz{};
    if (fused) {
        fz.ids = fused->ids; fz.raw = fused->raw; fz.counts = fused->counts; fz.sim_scale = fused->sim_scale;
        fz.out_ids = fused->out_ids; fz.out_sims = fused->out_sims; fz.out_rel = fused->out_rel; fz.out_counts = fused->out_counts;
    }
    if (nq <= 0 || m <= 0) return cudaSuccess;
    if (m > kMmrThreads || k_out <= 0) return cudaErrorInvalidValue;
    int row_bytes;
    switch (store) {
        case CRS_F16: case CRS_BF16: row_bytes = dim_padded * 2; break;
        case CRS_I8: row_bytes = dim_padded; break;
        case CRS_B1: row_bytes = dim_padded / 8; break;
        default: return cudaErrorInvalidValue;
    }
    const size_t smem = (size_t)m * (row_bytes + 16);
    const uint8_t* v = reinterpret_cast<const uint8_t*>(vecs);
#define CRS_MMR_LAUNCH(S)                                                                              \
    do {                                                                                               \
        cudaError_t e = cudaFuncSetAttribute(mmr_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                \
        mmr_kernel<S><<<nq, kMmrThreads, smem, st>>>(v, row_bytes, dim, relevance, m, k_out, lambda, out_order, fz);    \
    } while (0)
    switch (store) {
        case CRS_F16:  CRS_MMR_LAUNCH(CRS_F16); break;
        case CRS_BF16: CRS_MMR_LAUNCH(CRS_BF16); break;
        case CRS_I8:   CRS_MMR_LAUNCH(CRS_I8); break;
        default:       CRS_MMR_LAUNCH(CRS_B1); break;
    }
#undef CRS_MMR_LAUNCH
    return cudaGetLastError();
}

}  // namespace crs
