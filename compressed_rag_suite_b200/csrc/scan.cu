// K1 / K2 / K3 — single-query stream scans over the stored rows with fused
// threshold and in-register top-M.
//
// Replaces the O(N*D) distance pass that `collection.query` performs inside
// ChromaDB for the reference (rag/indexing.py:171-176) with an exhaustive,
// HBM-bandwidth-bound pass.  Algorithmic bytes per launch: n_rows * row_bytes.
//
//   K1  fp16 / bf16 rows : fp32 FMA dot (candidate ranking only, see finalize.cu)
//   K2  int8 rows        : dp4a, exact int32 dot
//   K3  1-bit rows       : xor + popc, exact Hamming -> raw = dim - 2h
//
// Structure (persistent, one CTA per SM):
//   * warp NW (producer): one elected lane streams tiles of NW*4*ITERS consecutive
//     rows with cp.async.bulk (UBLKCP) into a STAGES-deep shared-memory ring,
//     completion on "full" mbarriers; it re-arms a slot when the NW consumer
//     warps have arrived on its "empty" mbarrier.
//   * warps 0..NW-1 (consumers): each takes 4*ITERS rows of the tile, 8 lanes per
//     row; lane `sub` reads 16-byte chunks sub, sub+8, ... of its row (a
//     quarter-warp covers 128 contiguous bytes -> conflict-free LDS.128), does the
//     math against the query held in registers, 3 xor-shuffles finish the dot.
//     Every row goes through the same instruction sequence, so equal rows get
//     equal scores.
//   * candidates: score >= threshold and key > the warp's floor go into a
//     warp-distributed top-M ("replace the minimum", WarpTopM).  At the end the
//     warp lists are bitonic-sorted and tree-merged through smem; the CTA writes
//     one sorted list of M keys.  finalize.cu merges the CTA lists (and, for the
//     float stores, rescores exactly in fp64 and certifies the result).
//
// For fp16/bf16 the score computed here only picks candidates; its error against
// the canonical fp64 score is bounded by Dp * 2^-24 * |q| * |c| (any summation order).
#include "common.cuh"
#include "crs_internal.h"

namespace crs {

constexpr int kScanMaxStages = 16;
enum ScanKind { kF16 = 0, kBF16 = 1, kI8 = 2, kB1 = 3 };

template <int KIND>
__device__ __forceinline__ float2 cvt_pair(uint32_t v) {
    if constexpr (KIND == kBF16) {
        return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xFFFF0000u));
    } else {
        return __half22float2(*reinterpret_cast<const __half2*>(&v));
    }
}

// Fused query encode (see FusedQuery): the NT consumer threads of a CTA turn the fp32 query into its stored
// form in shared memory.  Arithmetic and order are those of ingest.cu: n2 accumulated sequentially in fp64 by
// ONE thread, y = x / sqrt(n2) in fp64, one rounding to the store dtype.
template <int KIND, int NT>
__device__ __forceinline__ void encode_query_cta(const FusedQuery& fq, int dim_padded, float* qstage, uint8_t* qenc,
                                                 int tid, bool first_cta) {
    __shared__ double s_div;
    __shared__ int s_zero;
    for (int j = tid; j < fq.dim; j += NT) qstage[j] = fq.src[j];
    named_barrier_1<NT>();
    if (tid == 0) {
        double n2 = 0.0;
        for (int j0 = 0; j0 < fq.dim; j0 += 16) {                 // loads and converts run ahead of the serial FMA chain
            double xv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) xv[j] = (j0 + j < fq.dim) ? (double)qstage[j0 + j] : 0.0;
#pragma unroll
            for (int j = 0; j < 16; ++j) n2 = fma(xv[j], xv[j], n2);
        }
        const double norm = sqrt(n2);
        s_div = (fq.cosine && norm > 0.0) ? norm : 1.0;
        s_zero = (fq.cosine && !(norm > 0.0)) ? 1 : 0;
        if (first_cta) {
            if (fq.qnorm_out) *fq.qnorm_out = fq.cosine ? 1.00390625f : __double2float_ru(norm) * 1.00390625f;
            if (fq.zero_word) *fq.zero_word = 0;
            if (fq.inc_word) *fq.inc_word += 1u;
        }
    }
    named_barrier_1<NT>();
    const double dv = s_div;
    const bool zr = s_zero != 0;
    for (int j = tid; j < dim_padded; j += NT) {                  // NT is a multiple of 32: a warp covers 32 consecutive j
        double y = 0.0;
        if (j < fq.dim && !zr) y = (double)qstage[j] / dv;
        if constexpr (KIND == kF16) {
            reinterpret_cast<__half*>(qenc)[j] = __double2half(y);
        } else if constexpr (KIND == kBF16) {
            reinterpret_cast<__nv_bfloat16*>(qenc)[j] = __double2bfloat16(y);
        } else if constexpr (KIND == kI8) {
            double q = rint(y * fq.i8_mult);
            q = fmin(127.0, fmax(-127.0, q));
            reinterpret_cast<int8_t*>(qenc)[j] = (int8_t)(int)q;
        } else {
            const unsigned bits = __ballot_sync(CRS_FULL_MASK, y > 0.0);
            if ((tid & 31) == 0) reinterpret_cast<uint32_t*>(qenc)[j >> 5] = bits;
        }
    }
    named_barrier_1<NT>();
    if (first_cta && fq.qcodes_out) {
        constexpr int kEsz8 = (KIND == kF16 || KIND == kBF16) ? 16 : (KIND == kI8 ? 8 : 1);     // bits per element
        const int bytes = dim_padded * kEsz8 / 8;
        for (int i = tid; i < bytes / 16; i += NT)
            reinterpret_cast<uint4*>(fq.qcodes_out)[i] = reinterpret_cast<const uint4*>(qenc)[i];
    }
}

// NCH: 128-byte chunk groups per row (row_bytes = 128*NCH).  QREG: query in registers.
template <int KIND, int NCH, int NW, int ITERS, int LPL, bool QREG>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
scan_kernel(const uint8_t* __restrict__ codes, int64_t n_rows, const uint8_t* __restrict__ qcodes,
            uint32_t ord_min, int32_t b1_dim, uint64_t* __restrict__ cand, int stages,
            const uint32_t* __restrict__ allow, FusedQuery fq) {
    constexpr bool kFloat = (KIND == kF16 || KIND == kBF16);
    constexpr int ROWB = NCH * 128;
    constexpr int ROWS_PER_WARP = 4 * ITERS;
    constexpr int TILE_ROWS = NW * ROWS_PER_WARP;
    constexpr int TILE_BYTES = TILE_ROWS * ROWB;
    constexpr int M = 32 * LPL;
    constexpr int QN = kFloat ? NCH * 8 : NCH * 4;          // query registers per lane

    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full_bar[kScanMaxStages];
    __shared__ uint64_t empty_bar[kScanMaxStages];
    // largest "M-th best" any warp of this CTA has reached: a key below it cannot be in the CTA's
    // top-M, so every warp filters against it (cuts the insert rate by ~NW)
    __shared__ unsigned long long cta_floor;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;

    if (threadIdx.x == 0) {
        cta_floor = 0ull;
        for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NW); }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NW) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const int64_t r0 = t * TILE_ROWS;
                const uint32_t bytes = (uint32_t)(min((int64_t)TILE_ROWS, n_rows - r0) * ROWB);
                mbar_arrive_expect_tx(&full_bar[stage], bytes);
                bulk_g2s(smem + (size_t)stage * TILE_BYTES, codes + r0 * ROWB, bytes, &full_bar[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumers
    if (fq.src != nullptr) {                  // fused query encode: behind the ring, while the producer fills it
        uint8_t* qenc = smem + (size_t)stages * TILE_BYTES;
        encode_query_cta<KIND, NW * 32>(fq, kFloat ? NCH * 64 : (KIND == kI8 ? NCH * 128 : NCH * 1024),
                                        reinterpret_cast<float*>(qenc + ROWB), qenc, threadIdx.x, blockIdx.x == 0);
        qcodes = qenc;
    }
    const int sub = lane & 7, grp = lane >> 3;
    float qf[(kFloat && QREG) ? QN : 1];
    uint32_t qi[kFloat ? 1 : QN];
    if constexpr (kFloat && QREG) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(qcodes + (c * 8 + sub) * 16);
            const float2 a = cvt_pair<KIND>(v.x), b = cvt_pair<KIND>(v.y), d = cvt_pair<KIND>(v.z), e = cvt_pair<KIND>(v.w);
            qf[c * 8 + 0] = a.x; qf[c * 8 + 1] = a.y; qf[c * 8 + 2] = b.x; qf[c * 8 + 3] = b.y;
            qf[c * 8 + 4] = d.x; qf[c * 8 + 5] = d.y; qf[c * 8 + 6] = e.x; qf[c * 8 + 7] = e.y;
        }
    }
    if constexpr (!kFloat) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(qcodes + (c * 8 + sub) * 16);
            qi[c * 4 + 0] = v.x; qi[c * 4 + 1] = v.y; qi[c * 4 + 2] = v.z; qi[c * 4 + 3] = v.w;
        }
    }

    WarpTopM<LPL> top; top.init();
    uint64_t published = 0ull;

    int stage = 0; uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait(&full_bar[stage], phase);
        uint64_t floor_eff = u64max(top.floor_key, *reinterpret_cast<volatile unsigned long long*>(&cta_floor));
        const uint8_t* tile = smem + (size_t)stage * TILE_BYTES;
        uint32_t ord[ITERS];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int row_in_tile = warp * ROWS_PER_WARP + it * 4 + grp;
            const int64_t row = t * TILE_ROWS + row_in_tile;
            const uint8_t* rp = tile + (size_t)row_in_tile * ROWB + sub * 16;
            if constexpr (kFloat) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                if (row < n_rows) {          // rows past the end of a partial tile were not copied
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const uint4 v = *reinterpret_cast<const uint4*>(rp + c * 128);
                        const float2 f0 = cvt_pair<KIND>(v.x), f1 = cvt_pair<KIND>(v.y), f2 = cvt_pair<KIND>(v.z), f3 = cvt_pair<KIND>(v.w);
                        if constexpr (QREG) {
                            a0 = fmaf(f0.x, qf[c * 8 + 0], a0); a1 = fmaf(f0.y, qf[c * 8 + 1], a1);
                            a2 = fmaf(f1.x, qf[c * 8 + 2], a2); a3 = fmaf(f1.y, qf[c * 8 + 3], a3);
                            a0 = fmaf(f2.x, qf[c * 8 + 4], a0); a1 = fmaf(f2.y, qf[c * 8 + 5], a1);
                            a2 = fmaf(f3.x, qf[c * 8 + 6], a2); a3 = fmaf(f3.y, qf[c * 8 + 7], a3);
                        } else {
                            const uint4 w = *reinterpret_cast<const uint4*>(qcodes + (c * 8 + sub) * 16);
                            const float2 g0 = cvt_pair<KIND>(w.x), g1 = cvt_pair<KIND>(w.y), g2 = cvt_pair<KIND>(w.z), g3 = cvt_pair<KIND>(w.w);
                            a0 = fmaf(f0.x, g0.x, a0); a1 = fmaf(f0.y, g0.y, a1);
                            a2 = fmaf(f1.x, g1.x, a2); a3 = fmaf(f1.y, g1.y, a3);
                            a0 = fmaf(f2.x, g2.x, a0); a1 = fmaf(f2.y, g2.y, a1);
                            a2 = fmaf(f3.x, g3.x, a2); a3 = fmaf(f3.y, g3.y, a3);
                        }
                    }
                }
                float s = (a0 + a1) + (a2 + a3);
                s += __shfl_xor_sync(CRS_FULL_MASK, s, 1);
                s += __shfl_xor_sync(CRS_FULL_MASK, s, 2);
                s += __shfl_xor_sync(CRS_FULL_MASK, s, 4);
                ord[it] = orderable_f32(s);
            } else {
                int acc = 0;
                if (row < n_rows) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const uint4 v = *reinterpret_cast<const uint4*>(rp + c * 128);
                        if constexpr (KIND == kI8) {
                            acc = __dp4a((int)v.x, (int)qi[c * 4 + 0], acc);
                            acc = __dp4a((int)v.y, (int)qi[c * 4 + 1], acc);
                            acc = __dp4a((int)v.z, (int)qi[c * 4 + 2], acc);
                            acc = __dp4a((int)v.w, (int)qi[c * 4 + 3], acc);
                        } else {
                            acc += __popc(v.x ^ qi[c * 4 + 0]) + __popc(v.y ^ qi[c * 4 + 1]) +
                                   __popc(v.z ^ qi[c * 4 + 2]) + __popc(v.w ^ qi[c * 4 + 3]);
                        }
                    }
                }
                acc += __shfl_xor_sync(CRS_FULL_MASK, acc, 1);
                acc += __shfl_xor_sync(CRS_FULL_MASK, acc, 2);
                acc += __shfl_xor_sync(CRS_FULL_MASK, acc, 4);
                if constexpr (KIND == kB1) acc = b1_dim - 2 * acc;
                ord[it] = orderable_i32(acc);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);      // data is in registers: free the slot
        if (++stage == stages) { stage = 0; phase ^= 1; }

#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int64_t row = t * TILE_ROWS + warp * ROWS_PER_WARP + it * 4 + grp;
            const uint64_t key = make_key(ord[it], (uint32_t)row);
            bool pass = (sub == 0) && (row < n_rows) && (ord[it] >= ord_min) && (key > floor_eff);
            if (allow != nullptr && pass) pass = (allow[row >> 5] >> (row & 31)) & 1u;   // where / where_document filter
            unsigned bal = __ballot_sync(CRS_FULL_MASK, pass);
            while (bal) {
                const int src = __ffs(bal) - 1;
                bal &= bal - 1;
                const uint64_t kb = shfl_u64(key, src);
                if (kb > floor_eff) {
                    top.insert(kb, lane);
                    floor_eff = u64max(floor_eff, top.floor_key);
                }
            }
        }
        if (top.floor_key > published) {                 // warp-uniform; non-zero only once the list is full
            published = top.floor_key;
            if (lane == 0) atomicMax(&cta_floor, (unsigned long long)published);
        }
    }

    // ------------------------------------------------- CTA reduction of the warp lists
    warp_sort_desc<LPL>(top.e, lane);
    // every consumer has passed its last full-barrier wait, so all copies have landed
    // and the ring can be reused as merge scratch once all consumers are here.
    named_barrier_1<NW * 32>();
    uint64_t* stage_keys = reinterpret_cast<uint64_t*>(smem);
    block_merge_lists<LPL, NW>(top.e, stage_keys, warp, lane);
    if (warp == 0) {
#pragma unroll
        for (int sl = 0; sl < LPL; ++sl)
            cand[(size_t)blockIdx.x * M + lane * LPL + sl] = top.e[sl];
    }
}

// K2 / K3 for short rows (128 or 256 bytes, e.g. 1024-bit codes): one THREAD per row.
// With 8 lanes per row a 128-byte row costs three shuffles, a key build and a ballot per 16
// bytes of payload and the kernel is issue-bound (ncu: 19.6 warp-instructions per row, 3.9 TB/s).
// Here lane l walks row l of its warp's 32 rows in the rotated chunk order (c + l) mod CH, so
// a quarter-warp still touches 8 distinct 16-byte bank groups (conflict-free LDS.128) and each
// lane keeps the query pre-rotated the same way in registers.  Integer scores are
// order-independent, so the result is unchanged.  Same ring, same top-M, same output.
template <int KIND, int NCH, int NW, int LPL>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
scan_rows_kernel(const uint8_t* __restrict__ codes, int64_t n_rows, const uint8_t* __restrict__ qcodes,
                 uint32_t ord_min, int32_t b1_dim, uint64_t* __restrict__ cand, int stages,
                 const uint32_t* __restrict__ allow, FusedQuery fq) {
    static_assert(KIND == kI8 || KIND == kB1, "integer stores only");
    static_assert(NCH == 1 || NCH == 2, "rows of 128 or 256 bytes");
    constexpr int ROWB = NCH * 128;
    constexpr int CH = NCH * 8;                       // 16-byte chunks per row
    constexpr int TILE_ROWS = NW * 32;
    constexpr int TILE_BYTES = TILE_ROWS * ROWB;
    constexpr int M = 32 * LPL;

    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full_bar[kScanMaxStages];
    __shared__ uint64_t empty_bar[kScanMaxStages];
    __shared__ unsigned long long cta_floor;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;

    if (threadIdx.x == 0) {
        cta_floor = 0ull;
        for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NW); }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NW) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const int64_t r0 = t * TILE_ROWS;
                const uint32_t bytes = (uint32_t)(min((int64_t)TILE_ROWS, n_rows - r0) * ROWB);
                mbar_arrive_expect_tx(&full_bar[stage], bytes);
                bulk_g2s(smem + (size_t)stage * TILE_BYTES, codes + r0 * ROWB, bytes, &full_bar[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    if (fq.src != nullptr) {                  // fused query encode: behind the ring, while the producer fills it
        uint8_t* qenc = smem + (size_t)stages * TILE_BYTES;
        encode_query_cta<KIND, NW * 32>(fq, KIND == kI8 ? NCH * 128 : NCH * 1024,
                                        reinterpret_cast<float*>(qenc + ROWB), qenc, threadIdx.x, blockIdx.x == 0);
        qcodes = qenc;
    }
    uint32_t qi[CH * 4];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(qcodes + ((c + lane) & (CH - 1)) * 16);
        qi[c * 4 + 0] = v.x; qi[c * 4 + 1] = v.y; qi[c * 4 + 2] = v.z; qi[c * 4 + 3] = v.w;
    }

    WarpTopM<LPL> top; top.init();
    uint64_t published = 0ull;
    int stage = 0; uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait(&full_bar[stage], phase);
        uint64_t floor_eff = u64max(top.floor_key, *reinterpret_cast<volatile unsigned long long*>(&cta_floor));
        const int row_in_tile = warp * 32 + lane;
        const int64_t row = t * TILE_ROWS + row_in_tile;
        const uint8_t* rp = smem + (size_t)stage * TILE_BYTES + (size_t)row_in_tile * ROWB;
        int acc = 0;
        if (row < n_rows) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const uint4 v = *reinterpret_cast<const uint4*>(rp + ((c + lane) & (CH - 1)) * 16);
                if constexpr (KIND == kI8) {
                    acc = __dp4a((int)v.x, (int)qi[c * 4 + 0], acc);
                    acc = __dp4a((int)v.y, (int)qi[c * 4 + 1], acc);
                    acc = __dp4a((int)v.z, (int)qi[c * 4 + 2], acc);
                    acc = __dp4a((int)v.w, (int)qi[c * 4 + 3], acc);
                } else {
                    acc += __popc(v.x ^ qi[c * 4 + 0]) + __popc(v.y ^ qi[c * 4 + 1]) +
                           __popc(v.z ^ qi[c * 4 + 2]) + __popc(v.w ^ qi[c * 4 + 3]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }

        if constexpr (KIND == kB1) acc = b1_dim - 2 * acc;
        const uint32_t ord = orderable_i32(acc);
        // cheap 32-bit pre-test on the score; the 64-bit key is only built for survivors
        bool pass = (row < n_rows) && (ord >= ord_min) && (ord >= key_ord(floor_eff));
        uint64_t key = 0ull;
        if (pass) {
            key = make_key(ord, (uint32_t)row);
            pass = key > floor_eff;
            if (allow != nullptr && pass) pass = (allow[row >> 5] >> (row & 31)) & 1u;
        }
        unsigned bal = __ballot_sync(CRS_FULL_MASK, pass);
        while (bal) {
            const int src = __ffs(bal) - 1;
            bal &= bal - 1;
            const uint64_t kb = shfl_u64(key, src);
            if (kb > floor_eff) {
                top.insert(kb, lane);
                floor_eff = u64max(floor_eff, top.floor_key);
            }
        }
        if (top.floor_key > published) {
            published = top.floor_key;
            if (lane == 0) atomicMax(&cta_floor, (unsigned long long)published);
        }
    }

    warp_sort_desc<LPL>(top.e, lane);
    named_barrier_1<NW * 32>();
    uint64_t* stage_keys = reinterpret_cast<uint64_t*>(smem);
    block_merge_lists<LPL, NW>(top.e, stage_keys, warp, lane);
    if (warp == 0) {
#pragma unroll
        for (int sl = 0; sl < LPL; ++sl)
            cand[(size_t)blockIdx.x * M + lane * LPL + sl] = top.e[sl];
    }
}

// K2 / K3, small batches: NQ queries share ONE pass over short rows (thread per row as above).
// A batch of b Hamming queries costs ceil(b / NQ) corpus reads instead of b (BASELINE config 5,
// b = 8).  The queries live in shared memory in stored form; lane l reads chunk (c + l) mod CH
// of the row AND of each query, so both loads are conflict-free LDS.128; NQ accumulators and NQ
// warp-distributed top-M lists stay in registers.  Per row and query the result is the same
// integer the single-query kernel computes, so the lists are identical.
template <int KIND, int NCH, int NW, int LPL, int NQ>
__global__ void __launch_bounds__((NW + 1) * 32, 1)
scan_rows_multi_kernel(const uint8_t* __restrict__ codes, int64_t n_rows, const uint8_t* __restrict__ qcodes,
                       uint32_t ord_min, int32_t b1_dim, uint64_t* __restrict__ cand, size_t cand_q_stride, int stages,
                       const uint32_t* __restrict__ allow) {
    static_assert(KIND == kI8 || KIND == kB1, "integer stores only");
    constexpr int ROWB = NCH * 128;
    constexpr int CH = NCH * 8;
    constexpr int TILE_ROWS = NW * 32;
    constexpr int TILE_BYTES = TILE_ROWS * ROWB;
    constexpr int M = 32 * LPL;

    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(16) uint8_t qs[NQ * ROWB];
    __shared__ uint64_t full_bar[kScanMaxStages];
    __shared__ uint64_t empty_bar[kScanMaxStages];
    __shared__ unsigned long long cta_floor[NQ];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (n_rows + TILE_ROWS - 1) / TILE_ROWS;

    for (int i = threadIdx.x; i < NQ * ROWB / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(qs)[i] = reinterpret_cast<const uint4*>(qcodes)[i];
    if (threadIdx.x < NQ) cta_floor[threadIdx.x] = 0ull;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NW); }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NW) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                const int64_t r0 = t * TILE_ROWS;
                const uint32_t bytes = (uint32_t)(min((int64_t)TILE_ROWS, n_rows - r0) * ROWB);
                mbar_arrive_expect_tx(&full_bar[stage], bytes);
                bulk_g2s(smem + (size_t)stage * TILE_BYTES, codes + r0 * ROWB, bytes, &full_bar[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    WarpTopM<LPL> top[NQ];
    uint64_t published[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) { top[qi].init(); published[qi] = 0ull; }
    uint32_t thr[NQ];
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) thr[qi] = ord_min;
    int iter = 0;

    int stage = 0; uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        mbar_wait(&full_bar[stage], phase);
        const int row_in_tile = warp * 32 + lane;
        const int64_t row = t * TILE_ROWS + row_in_tile;
        const uint8_t* rp = smem + (size_t)stage * TILE_BYTES + (size_t)row_in_tile * ROWB;
        int acc[NQ];
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) acc[qi] = 0;
        if (row < n_rows) {
            if constexpr (KIND == kI8) {
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int off = ((c + lane) & (CH - 1)) * 16;
                    const uint4 v = *reinterpret_cast<const uint4*>(rp + off);
#pragma unroll
                    for (int qi = 0; qi < NQ; ++qi) {
                        const uint4 w = *reinterpret_cast<const uint4*>(qs + qi * ROWB + off);
                        acc[qi] = __dp4a((int)v.x, (int)w.x, acc[qi]); acc[qi] = __dp4a((int)v.y, (int)w.y, acc[qi]);
                        acc[qi] = __dp4a((int)v.z, (int)w.z, acc[qi]); acc[qi] = __dp4a((int)v.w, (int)w.w, acc[qi]);
                    }
                }
            } else {
                // Hamming distance with a carry-save adder (first Harley-Seal level).  POPC issues at a
                // quarter of the LOP3 rate, and the shared-pass kernel is bound by whichever pipe is busier:
                // 4 POPC per chunk and query saturate the POPC pipe, a full two-level adder (1 POPC, 14 LOP3)
                // saturates the ALU pipe.  Balanced point: fold the four xor words into a running `ones`
                // bit-plane (4 LOP3) and count the two weight-2 carries (2 POPC).
                //   h = popc(ones) + 2 * sum(popc(carries))
                uint32_t ones[NQ];
#pragma unroll
                for (int qi = 0; qi < NQ; ++qi) ones[qi] = 0u;
                const int one = (stages > 0) ? 1 : 0;              // always 1; opaque to the compiler
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int off = ((c + lane) & (CH - 1)) * 16;
                    const uint4 v = *reinterpret_cast<const uint4*>(rp + off);
#pragma unroll
                    for (int qi = 0; qi < NQ; ++qi) {
                        const uint4 w = *reinterpret_cast<const uint4*>(qs + qi * ROWB + off);
                        const uint32_t x0 = v.x ^ w.x, x1 = v.y ^ w.y, x2 = v.z ^ w.z, x3 = v.w ^ w.w;
                        const uint32_t o1 = ones[qi] ^ x0 ^ x1;
                        const uint32_t ca = (ones[qi] & x0) | (ones[qi] & x1) | (x0 & x1);
                        const uint32_t o2 = o1 ^ x2 ^ x3;
                        const uint32_t cb = (o1 & x2) | (o1 & x3) | (x2 & x3);
                        ones[qi] = o2;
                        // the ALU pipe (LOP3) is the busiest one here (ncu: 75 %): the two additions are issued as
                        // IMADs (multiplier `one` is a run-time 1) so they go to the idle FMA pipe instead
                        acc[qi] = __popc(ca) * one + acc[qi];
                        acc[qi] = __popc(cb) * one + acc[qi];
                    }
                }
#pragma unroll
                for (int qi = 0; qi < NQ; ++qi) acc[qi] = 2 * acc[qi] + __popc(ones[qi]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }

        // Common case per (row, query): ONE 32-bit compare against a cached threshold and one vote.  The
        // threshold (largest of the caller's minimum, this warp's floor and the CTA floor, in the orderable
        // score domain) is refreshed after every insert and every 8 tiles; a stale value is only ever too
        // low, so nothing that belongs in the list is skipped.  Everything 64-bit happens behind the vote.
        const bool in_range = row < n_rows;
        const bool refresh = ((iter++) & 7) == 0;
#pragma unroll
        for (int qi = 0; qi < NQ; ++qi) {
            int a = acc[qi];
            if constexpr (KIND == kB1) a = b1_dim - 2 * a;
            const uint32_t ord = orderable_i32(a);
            if (refresh) {
                const uint64_t f = u64max(top[qi].floor_key, *reinterpret_cast<volatile unsigned long long*>(&cta_floor[qi]));
                thr[qi] = max(ord_min, key_ord(f));
            }
            if (!__any_sync(CRS_FULL_MASK, in_range && ord >= thr[qi])) continue;
            uint64_t floor_eff = u64max(top[qi].floor_key, *reinterpret_cast<volatile unsigned long long*>(&cta_floor[qi]));
            bool pass = in_range && (ord >= ord_min) && (ord >= key_ord(floor_eff));
            uint64_t key = 0ull;
            if (pass) {
                key = make_key(ord, (uint32_t)row);
                pass = key > floor_eff;
                if (allow != nullptr && pass) pass = (allow[row >> 5] >> (row & 31)) & 1u;   // filter: hits only
            }
            unsigned bal = __ballot_sync(CRS_FULL_MASK, pass);
            while (bal) {
                const int src = __ffs(bal) - 1;
                bal &= bal - 1;
                const uint64_t kb = shfl_u64(key, src);
                if (kb > floor_eff) {
                    top[qi].insert(kb, lane);
                    floor_eff = u64max(floor_eff, top[qi].floor_key);
                }
            }
            thr[qi] = max(ord_min, key_ord(floor_eff));
            if (top[qi].floor_key > published[qi]) {
                published[qi] = top[qi].floor_key;
                if (lane == 0) atomicMax(&cta_floor[qi], (unsigned long long)published[qi]);
            }
        }
    }

    named_barrier_1<NW * 32>();                       // all copies landed, ring reusable as merge scratch
    uint64_t* stage_keys = reinterpret_cast<uint64_t*>(smem);
#pragma unroll
    for (int qi = 0; qi < NQ; ++qi) {
        warp_sort_desc<LPL>(top[qi].e, lane);
        block_merge_lists<LPL, NW>(top[qi].e, stage_keys, warp, lane);
        if (warp == 0) {
#pragma unroll
            for (int sl = 0; sl < LPL; ++sl)
                cand[(size_t)qi * cand_q_stride + (size_t)blockIdx.x * M + lane * LPL + sl] = top[qi].e[sl];
        }
    }
}

template <int KIND, int NCH, int LPL, int NQ>
static cudaError_t launch_rows_multi(cudaStream_t st, const void* codes, int64_t n, const void* qcodes, uint32_t ord_min,
                                     int32_t b1_dim, uint64_t* cand, size_t cand_q_stride, int grid, const uint32_t* allow) {
    // 8 top-128 lists per warp need ~200 registers per thread: half the warps for that combination
    constexpr int NW = (NCH == 1 && !(LPL == 4 && NQ == 8)) ? 16 : 8;
    constexpr int TILE_BYTES = NW * 32 * NCH * 128;
    constexpr int M = 32 * LPL;
    int stages = (190 * 1024) / TILE_BYTES;
    if (stages > kScanMaxStages) stages = kScanMaxStages;
    size_t smem = (size_t)stages * TILE_BYTES;
    const size_t merge_bytes = (size_t)NW * M * sizeof(uint64_t);
    if (smem < merge_bytes) smem = merge_bytes;
    auto kern = scan_rows_multi_kernel<KIND, NCH, NW, LPL, NQ>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, (NW + 1) * 32, smem, st>>>(reinterpret_cast<const uint8_t*>(codes), n,
                                            reinterpret_cast<const uint8_t*>(qcodes), ord_min, b1_dim, cand, cand_q_stride,
                                            stages, allow);
    return cudaGetLastError();
}

template <int KIND, int NCH, int LPL>
static cudaError_t rows_multi_by_nq(int nq, cudaStream_t st, const void* codes, int64_t n, const void* qcodes,
                                    uint32_t ord_min, int32_t b1_dim, uint64_t* cand, size_t cand_q_stride, int grid,
                                    const uint32_t* allow) {
    switch (nq) {
        case 8: return launch_rows_multi<KIND, NCH, LPL, 8>(st, codes, n, qcodes, ord_min, b1_dim, cand, cand_q_stride, grid, allow);
        case 4: return launch_rows_multi<KIND, NCH, LPL, 4>(st, codes, n, qcodes, ord_min, b1_dim, cand, cand_q_stride, grid, allow);
        case 2: return launch_rows_multi<KIND, NCH, LPL, 2>(st, codes, n, qcodes, ord_min, b1_dim, cand, cand_q_stride, grid, allow);
        default: return cudaErrorInvalidValue;
    }
}

// Largest group of queries (8, 4, 2) the shared-pass kernel takes for this row width and list
// size, 1 = not available (the caller scans that query alone).
int scan_multi_group(crs_dtype store, int row_bytes, int lpl, int nq_left, int max_group) {
    if ((store != CRS_B1 && store != CRS_I8) || (row_bytes != 128 && row_bytes != 256)) return 1;
    int cap = (lpl == 4) ? ((row_bytes == 128) ? 8 : 4) : 8;
    if (max_group < cap) cap = max_group;
    for (int g = cap; g >= 2; g >>= 1) if (nq_left >= g) return g;
    return 1;
}

cudaError_t launch_scan_int_multi(cudaStream_t st, crs_dtype store, const void* codes, int64_t n, int row_bytes, int dim,
                                  const void* qcodes, int nq_group, int32_t min_raw, uint64_t* cand, size_t cand_q_stride,
                                  const ScanPlan& p) {
    const uint32_t om = orderable_i32(min_raw);
    const int nch = row_bytes / 128;
#define CRS_MULTI(KIND_, NCH_)                                                                                          \
    (p.lpl == 1 ? rows_multi_by_nq<KIND_, NCH_, 1>(nq_group, st, codes, n, qcodes, om, dim, cand, cand_q_stride, p.grid, p.allow) \
                : rows_multi_by_nq<KIND_, NCH_, 4>(nq_group, st, codes, n, qcodes, om, dim, cand, cand_q_stride, p.grid, p.allow))
    if (store == CRS_B1) return nch == 1 ? CRS_MULTI(kB1, 1) : CRS_MULTI(kB1, 2);
    if (store == CRS_I8) return nch == 1 ? CRS_MULTI(kI8, 1) : CRS_MULTI(kI8, 2);
#undef CRS_MULTI
    return cudaErrorInvalidValue;
}

// dynamic shared memory of a scan: the ring, and behind it (fused query encode) the encoded query + its fp32 source
static size_t scan_smem_bytes(int stages, int tile_bytes, size_t merge_bytes, int row_bytes, const FusedQuery& fq) {
    size_t smem = (size_t)stages * tile_bytes;
    if (fq.src != nullptr) smem += (size_t)row_bytes + (size_t)fq.dim * sizeof(float) + 16;
    return smem < merge_bytes ? merge_bytes : smem;
}

template <int KIND, int NCH, int NW, int LPL>
static cudaError_t launch_rows(cudaStream_t st, const void* codes, int64_t n, const void* qcodes, uint32_t ord_min,
                               int32_t b1_dim, uint64_t* cand, int grid, const uint32_t* allow, const FusedQuery& fq) {
    constexpr int TILE_BYTES = NW * 32 * NCH * 128;
    constexpr int M = 32 * LPL;
    int stages = (200 * 1024) / TILE_BYTES;
    if (stages > kScanMaxStages) stages = kScanMaxStages;
    const size_t smem = scan_smem_bytes(stages, TILE_BYTES, (size_t)NW * M * sizeof(uint64_t), NCH * 128, fq);
    auto kern = scan_rows_kernel<KIND, NCH, NW, LPL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, (NW + 1) * 32, smem, st>>>(reinterpret_cast<const uint8_t*>(codes), n,
                                            reinterpret_cast<const uint8_t*>(qcodes), ord_min, b1_dim, cand, stages, allow, fq);
    return cudaGetLastError();
}

template <int KIND, int NCH>
static cudaError_t rows_by_lpl(cudaStream_t st, const void* codes, int64_t n, const void* qcodes, uint32_t ord_min,
                               int32_t b1_dim, uint64_t* cand, const ScanPlan& p) {
    constexpr int NW = (NCH == 1) ? 16 : 8;          // 64 query registers per lane at NCH = 2
    if (p.lpl == 1) return launch_rows<KIND, NCH, NW, 1>(st, codes, n, qcodes, ord_min, b1_dim, cand, p.grid, p.allow, p.fq);
    if (p.lpl == 4) return launch_rows<KIND, NCH, NW, 4>(st, codes, n, qcodes, ord_min, b1_dim, cand, p.grid, p.allow, p.fq);
    return cudaErrorInvalidValue;
}

template <int KIND, int NCH, int NW, int ITERS, int LPL, bool QREG>
static cudaError_t launch_one(cudaStream_t st, const void* codes, int64_t n, const void* qcodes, uint32_t ord_min,
                              int32_t b1_dim, uint64_t* cand, int grid, const uint32_t* allow, const FusedQuery& fq) {
    constexpr int TILE_BYTES = NW * 4 * ITERS * NCH * 128;
    constexpr int M = 32 * LPL;
    int stages = (200 * 1024) / TILE_BYTES;
    if (stages > kScanMaxStages) stages = kScanMaxStages;
    if (stages < 2) return cudaErrorInvalidConfiguration;
    const size_t smem = scan_smem_bytes(stages, TILE_BYTES, (size_t)NW * M * sizeof(uint64_t), NCH * 128, fq);
    auto kern = scan_kernel<KIND, NCH, NW, ITERS, LPL, QREG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, (NW + 1) * 32, smem, st>>>(reinterpret_cast<const uint8_t*>(codes), n,
                                            reinterpret_cast<const uint8_t*>(qcodes), ord_min, b1_dim, cand, stages, allow, fq);
    return cudaGetLastError();
}

template <int KIND, int NCH, int NW, int ITERS, bool QREG>
static cudaError_t by_lpl(cudaStream_t st, const void* codes, int64_t n, const void* qcodes, uint32_t ord_min,
                          int32_t b1_dim, uint64_t* cand, const ScanPlan& p) {
    if (p.lpl == 1) return launch_one<KIND, NCH, NW, ITERS, 1, QREG>(st, codes, n, qcodes, ord_min, b1_dim, cand, p.grid, p.allow, p.fq);
    if (p.lpl == 4) return launch_one<KIND, NCH, NW, ITERS, 4, QREG>(st, codes, n, qcodes, ord_min, b1_dim, cand, p.grid, p.allow, p.fq);
    return cudaErrorInvalidValue;
}

template <int KIND>
static cudaError_t by_rowbytes_float(cudaStream_t st, const void* codes, int64_t n, int nch, const void* qcodes,
                                     uint32_t ord_min, uint64_t* cand, const ScanPlan& p) {
    switch (nch) {   // row_bytes = 128 * nch, Dp = 64 * nch
        case 1:  return by_lpl<KIND, 1, 16, 4, true>(st, codes, n, qcodes, ord_min, 0, cand, p);
        case 2:  return by_lpl<KIND, 2, 16, 2, true>(st, codes, n, qcodes, ord_min, 0, cand, p);
        case 3:  return by_lpl<KIND, 3, 16, 1, true>(st, codes, n, qcodes, ord_min, 0, cand, p);
        case 4:  return by_lpl<KIND, 4, 16, 1, true>(st, codes, n, qcodes, ord_min, 0, cand, p);
        case 6:  return by_lpl<KIND, 6, 16, 1, true>(st, codes, n, qcodes, ord_min, 0, cand, p);
        case 8:  return by_lpl<KIND, 8, 8, 1, true>(st, codes, n, qcodes, ord_min, 0, cand, p);
        case 12: return by_lpl<KIND, 12, 8, 1, false>(st, codes, n, qcodes, ord_min, 0, cand, p);
        case 16: return by_lpl<KIND, 16, 8, 1, false>(st, codes, n, qcodes, ord_min, 0, cand, p);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_scan_f16(cudaStream_t st, const void* codes, int64_t n, int dim_padded, bool bf16,
                            const void* qcodes, float tau_pre, uint64_t* cand, const ScanPlan& plan) {
    const uint32_t ord_min = orderable_f32(tau_pre);
    return bf16 ? by_rowbytes_float<kBF16>(st, codes, n, dim_padded / 64, qcodes, ord_min, cand, plan)
                : by_rowbytes_float<kF16>(st, codes, n, dim_padded / 64, qcodes, ord_min, cand, plan);
}

template <int KIND>
static cudaError_t by_rowbytes_int(cudaStream_t st, const void* codes, int64_t n, int nch, const void* qcodes,
                                   uint32_t ord_min, int32_t b1_dim, uint64_t* cand, const ScanPlan& p) {
    switch (nch) {   // row_bytes = 128 * nch
        case 1:  return rows_by_lpl<KIND, 1>(st, codes, n, qcodes, ord_min, b1_dim, cand, p);
        case 2:  return rows_by_lpl<KIND, 2>(st, codes, n, qcodes, ord_min, b1_dim, cand, p);
        case 3:  return by_lpl<KIND, 3, 16, 2, true>(st, codes, n, qcodes, ord_min, b1_dim, cand, p);
        case 4:  return by_lpl<KIND, 4, 16, 2, true>(st, codes, n, qcodes, ord_min, b1_dim, cand, p);
        case 6:  return by_lpl<KIND, 6, 16, 1, true>(st, codes, n, qcodes, ord_min, b1_dim, cand, p);
        case 8:  return by_lpl<KIND, 8, 16, 1, true>(st, codes, n, qcodes, ord_min, b1_dim, cand, p);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_scan_i8(cudaStream_t st, const void* codes, int64_t n, int dim_padded,
                           const void* qcodes, int32_t min_raw, uint64_t* cand, const ScanPlan& plan) {
    return by_rowbytes_int<kI8>(st, codes, n, dim_padded / 128, qcodes, orderable_i32(min_raw), 0, cand, plan);
}

cudaError_t launch_scan_b1(cudaStream_t st, const void* codes, int64_t n, int dim_padded, int dim,
                           const void* qcodes, int32_t min_raw, uint64_t* cand, const ScanPlan& plan) {
    return by_rowbytes_int<kB1>(st, codes, n, dim_padded / 1024, qcodes, orderable_i32(min_raw), dim, cand, plan);
}

}  // namespace crs
