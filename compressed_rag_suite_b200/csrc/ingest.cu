// K0 — ingest: fp32 rows -> stored codes.
//
// Replaces what Chroma does with `collection.add(embeddings=...)`
// (reference rag/indexing.py:114-119; cosine space normalises, :83) and
// additionally casts / quantises / sign-packs the rows into the HBM layout the
// scan and GEMM kernels stream.  Arithmetic is the canonical definition in
// oracle/encode.py: n2 accumulated sequentially in fp64, y = x / sqrt(n2) in fp64,
// one rounding to the store dtype.  Bit-exact with the oracle by construction.
//
// Mapping: one warp owns 32 consecutive rows.  Row segments of 32 columns are
// loaded coalesced (lane = column) into a padded smem tile and walked by
// lane = row, so the fp64 accumulation order is j = 0..D-1 for every row.
#include "common.cuh"
#include "crs_internal.h"

namespace crs {

constexpr int kIngestWarps = 4;

template <int STORE>   // crs_dtype value
__global__ void __launch_bounds__(kIngestWarps * 32)
encode_kernel(const float* __restrict__ src, int64_t n, int dim, int dim_padded, int cosine,
              double i8_mult, void* __restrict__ dst, float* __restrict__ norms_out, int32_t* __restrict__ zero_word,
              uint32_t* __restrict__ inc_word) {
    __shared__ float tile[kIngestWarps][32][33];
    if (zero_word != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *zero_word = 0;   // per-search counter reset rides along
    if (inc_word != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *inc_word += 1u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = ((int64_t)blockIdx.x * kIngestWarps + warp) * 32;
    if (row0 >= n) return;
    const int rows_here = (int)min((int64_t)32, n - row0);
    float (*t)[33] = tile[warp];

    // pass 1: sequential fp64 sum of squares (also needed for ip: stored-row norm bound)
    double n2 = 0.0;
    for (int c0 = 0; c0 < dim; c0 += 32) {
        const int c = c0 + lane;
#pragma unroll
        for (int r = 0; r < 32; ++r)         // 32 independent loads in flight per lane
            t[r][lane] = (r < rows_here && c < dim) ? src[(row0 + r) * dim + c] : 0.f;
        __syncwarp();
        // columns past `dim` were staged as 0 and fma(0, 0, n2) == n2 exactly, so the tile is always
        // walked in full: the 32 loads + converts are issued up front and only the fp64 FMA chain
        // (the canonical sequential order j = 0..D-1) is serial
        double xv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) xv[j] = (double)t[lane][j];
#pragma unroll
        for (int j = 0; j < 32; ++j) n2 = fma(xv[j], xv[j], n2);        // x*x is exact in fp64, so fma == mul + add
        __syncwarp();
    }
    const double norm = sqrt(n2);
    const double div = (cosine && norm > 0.0) ? norm : 1.0;
    const bool zero_row = cosine && !(norm > 0.0);
    if (norms_out && lane < rows_here) {
        // upper bound of the stored row's norm: unit rows round to <= 1 + 2^-8 (bf16 worst case)
        const float nb = cosine ? 1.00390625f : __double2float_ru(norm) * 1.00390625f;
        norms_out[row0 + lane] = nb;
    }

    // pass 2: normalise + convert, lane = column so stores are coalesced
    for (int c0 = 0; c0 < dim_padded; c0 += 32) {
        const int c = c0 + lane;
        float v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r)         // issue all loads of the tile before the fp64 divides
            v[r] = (r < rows_here && c < dim) ? src[(row0 + r) * dim + c] : 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            if (r >= rows_here) break;       // warp-uniform
            const double dv = __shfl_sync(CRS_FULL_MASK, div, r);
            const bool zr = __shfl_sync(CRS_FULL_MASK, (int)zero_row, r) != 0;
            double y = 0.0;
            if (c < dim && !zr) y = (double)v[r] / dv;
            const int64_t o = (row0 + r) * (int64_t)dim_padded + c;
            if constexpr (STORE == CRS_F16) {
                reinterpret_cast<__half*>(dst)[o] = __double2half(y);
            } else if constexpr (STORE == CRS_BF16) {
                reinterpret_cast<__nv_bfloat16*>(dst)[o] = __double2bfloat16(y);
            } else if constexpr (STORE == CRS_I8) {
                double q = rint(y * i8_mult);                    // round half to even
                q = fmin(127.0, fmax(-127.0, q));
                reinterpret_cast<int8_t*>(dst)[o] = (int8_t)(int)q;
            } else {   // CRS_B1: bit j%32 of word j/32
                const unsigned w = __ballot_sync(CRS_FULL_MASK, y > 0.0);
                if (lane == 0) reinterpret_cast<uint32_t*>(dst)[o >> 5] = w;
            }
        }
    }
}

// K0 for a handful of rows (queries): the general kernel walks a row in 32-column steps and pays one
// global-memory round trip per step (12 for 384 dims, twice) — 19 us for ONE query, which is what a
// single-query search then waits for.  Here a CTA first copies its (up to) 32 rows into shared memory
// with independent coalesced loads, and both passes run from there.  Same arithmetic, same order.
constexpr int kSmallThreads = 128;
constexpr int kSmallRows = 8;            // rows per CTA: 1024 queries -> 128 CTAs (32 rows per CTA left 116 SMs idle: 42 us)

template <int STORE>
__global__ void __launch_bounds__(kSmallThreads)
encode_small_kernel(const float* __restrict__ src, int64_t n, int dim, int dim_padded, int cosine,
                    double i8_mult, void* __restrict__ dst, float* __restrict__ norms_out, int32_t* __restrict__ zero_word,
                    uint32_t* __restrict__ inc_word) {
    extern __shared__ float rows_sm[];                       // [kSmallRows][dim + 1]
    __shared__ double s_div[32];
    __shared__ int s_zero[32];
    if (zero_word != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *zero_word = 0;
    if (inc_word != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *inc_word += 1u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row0 = (int64_t)blockIdx.x * kSmallRows;
    const int rows_here = (int)min((int64_t)kSmallRows, n - row0);
    const int ld = dim + 1;
    for (int i = threadIdx.x; i < rows_here * dim; i += kSmallThreads) {
        const int r = i / dim, c = i - r * dim;
        rows_sm[r * ld + c] = src[(row0 + r) * dim + c];
    }
    __syncthreads();
    if (warp == 0) {
        double n2 = 0.0;
        if (lane < rows_here) {
            const float* t = rows_sm + lane * ld;
            for (int j0 = 0; j0 < dim; j0 += 16) {           // loads and converts run ahead of the serial FMA chain
                double xv[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) xv[j] = (j0 + j < dim) ? (double)t[j0 + j] : 0.0;
#pragma unroll
                for (int j = 0; j < 16; ++j) n2 = fma(xv[j], xv[j], n2);
            }
        }
        const double norm = sqrt(n2);
        s_div[lane] = (cosine && norm > 0.0) ? norm : 1.0;
        s_zero[lane] = (cosine && !(norm > 0.0)) ? 1 : 0;
        if (norms_out && lane < rows_here)
            norms_out[row0 + lane] = cosine ? 1.00390625f : __double2float_ru(norm) * 1.00390625f;
    }
    __syncthreads();
    // normalise + convert: a warp takes (row, 32 consecutive columns) so sign bits pack with one ballot
    const int col_blocks = dim_padded / 32;
    for (int w = warp; w < rows_here * col_blocks; w += kSmallThreads / 32) {
        const int r = w / col_blocks, c = (w - r * col_blocks) * 32 + lane;
        double y = 0.0;
        if (c < dim && !s_zero[r]) y = (double)rows_sm[r * ld + c] / s_div[r];
        const int64_t o = (row0 + r) * (int64_t)dim_padded + c;
        if constexpr (STORE == CRS_F16) {
            reinterpret_cast<__half*>(dst)[o] = __double2half(y);
        } else if constexpr (STORE == CRS_BF16) {
            reinterpret_cast<__nv_bfloat16*>(dst)[o] = __double2bfloat16(y);
        } else if constexpr (STORE == CRS_I8) {
            double q = rint(y * i8_mult);
            q = fmin(127.0, fmax(-127.0, q));
            reinterpret_cast<int8_t*>(dst)[o] = (int8_t)(int)q;
        } else {
            const unsigned bits = __ballot_sync(CRS_FULL_MASK, y > 0.0);
            if (lane == 0) reinterpret_cast<uint32_t*>(dst)[o >> 5] = bits;
        }
    }
}

cudaError_t launch_encode(cudaStream_t st, const float* src, int64_t n, int dim, int dim_padded,
                          crs_dtype store, crs_metric metric, float i8_scale, void* dst, float* norms_out,
                          int32_t* zero_word, uint32_t* inc_word) {
    if (n <= 0) return cudaSuccess;
    const int64_t rows_per_block = kIngestWarps * 32;
    const unsigned grid = (unsigned)((n + rows_per_block - 1) / rows_per_block);
    const int cosine = metric == CRS_COSINE;
    const double mult = 127.0 / (double)i8_scale;
    const size_t small_smem = (size_t)kSmallRows * (dim + 1) * sizeof(float);
    if (n <= 4096 && small_smem <= 160 * 1024) {           // query-sized inputs
        const unsigned g = (unsigned)((n + kSmallRows - 1) / kSmallRows);
#define CRS_ENC_SMALL(S)                                                                                         \
        do {                                                                                                     \
            cudaError_t e = cudaFuncSetAttribute(encode_small_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)small_smem); \
            if (e != cudaSuccess) return e;                                                                      \
            encode_small_kernel<S><<<g, kSmallThreads, small_smem, st>>>(src, n, dim, dim_padded, cosine, mult, dst, norms_out, zero_word, inc_word); \
        } while (0)
        switch (store) {
            case CRS_F16:  CRS_ENC_SMALL(CRS_F16); break;
            case CRS_BF16: CRS_ENC_SMALL(CRS_BF16); break;
            case CRS_I8:   CRS_ENC_SMALL(CRS_I8); break;
            case CRS_B1:   CRS_ENC_SMALL(CRS_B1); break;
            default: return cudaErrorInvalidValue;
        }
#undef CRS_ENC_SMALL
        return cudaGetLastError();
    }
    switch (store) {
        case CRS_F16:  encode_kernel<CRS_F16><<<grid, kIngestWarps * 32, 0, st>>>(src, n, dim, dim_padded, cosine, mult, dst, norms_out, zero_word, inc_word); break;
        case CRS_BF16: encode_kernel<CRS_BF16><<<grid, kIngestWarps * 32, 0, st>>>(src, n, dim, dim_padded, cosine, mult, dst, norms_out, zero_word, inc_word); break;
        case CRS_I8:   encode_kernel<CRS_I8><<<grid, kIngestWarps * 32, 0, st>>>(src, n, dim, dim_padded, cosine, mult, dst, norms_out, zero_word, inc_word); break;
        case CRS_B1:   encode_kernel<CRS_B1><<<grid, kIngestWarps * 32, 0, st>>>(src, n, dim, dim_padded, cosine, mult, dst, norms_out, zero_word, inc_word); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

__global__ void gather_rows_kernel(const uint4* __restrict__ codes, int chunks_per_row, int64_t n_rows,
                                   uint32_t row_base, const uint32_t* __restrict__ ids, int n,
                                   uint4* __restrict__ out) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const uint32_t gid = ids[i];
    if (gid < row_base || (int64_t)(gid - row_base) >= n_rows) return;
    const int64_t r = gid - row_base;
    for (int c = threadIdx.x; c < chunks_per_row; c += blockDim.x)
        out[(int64_t)i * chunks_per_row + c] = codes[r * chunks_per_row + c];
}

cudaError_t launch_gather_rows(cudaStream_t st, const void* codes, size_t row_bytes, int64_t n_rows,
                               uint32_t row_base, const uint32_t* ids, int n, void* out) {
    if (n <= 0) return cudaSuccess;
    gather_rows_kernel<<<n, 64, 0, st>>>(reinterpret_cast<const uint4*>(codes), (int)(row_bytes / 16), n_rows,
                                         row_base, ids, n, reinterpret_cast<uint4*>(out));
    return cudaGetLastError();
}

}  // namespace crs
