// K7 — cross-shard merge: G sorted lists of k_in (id, raw score) per query ->
// global top k_out by (score desc, id asc).
//
// No reference counterpart (the reference is single-node); it follows the one
// collective of the sharded search, an allgather of each rank's local top-k
// (SURVEY.md §8e).  Scores are the canonical raw scores, identical on every rank,
// so the merge is exact.  One warp per query, lists folded with the same
// register-resident bitonic merge the scans use.
#include "common.cuh"
#include "crs_internal.h"

namespace crs {

constexpr int kMergeWarps = 4;

template <int LPL, bool SORT>
__global__ void __launch_bounds__(kMergeWarps * 32)
merge_topk_kernel(const uint32_t* __restrict__ ids, const void* __restrict__ scores, int is_int,
                  int n_lists, int nq, int k_in, int k_out, size_t list_stride,
                  uint32_t* __restrict__ out_ids, void* __restrict__ out_scores, int32_t* __restrict__ out_counts) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kMergeWarps + warp;
    if (q >= nq) return;
    uint64_t e[LPL];
#pragma unroll
    for (int s = 0; s < LPL; ++s) e[s] = 0ull;
    for (int l = 0; l < n_lists; ++l) {
        uint64_t b[LPL];
        const size_t base = (size_t)l * list_stride + (size_t)q * k_in;
#pragma unroll
        for (int s = 0; s < LPL; ++s) {
            const int i = lane * LPL + s;
            uint64_t key = 0ull;
            if (i < k_in) {
                const uint32_t id = ids[base + i];
                if (id != CRS_PAD_ID) {
                    const uint32_t ord = is_int ? orderable_i32(reinterpret_cast<const int32_t*>(scores)[base + i])
                                                : orderable_f32(reinterpret_cast<const float*>(scores)[base + i]);
                    key = make_key(ord, id);
                }
            }
            b[s] = key;
        }
        if constexpr (SORT) warp_sort_desc<LPL>(b, lane);     // unsorted candidates (rescored lists)
        warp_merge_desc<LPL>(e, b, lane);
    }
    int nvalid = 0;
#pragma unroll
    for (int s = 0; s < LPL; ++s) nvalid += (e[s] != 0ull);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nvalid += __shfl_xor_sync(CRS_FULL_MASK, nvalid, off);
    const int count = min(nvalid, k_out);
#pragma unroll
    for (int s = 0; s < LPL; ++s) {
        const int i = lane * LPL + s;
        if (i < k_out) {
            const bool ok = i < count;
            out_ids[(size_t)q * k_out + i] = ok ? key_id(e[s]) : CRS_PAD_ID;
            if (is_int) reinterpret_cast<int32_t*>(out_scores)[(size_t)q * k_out + i] = ok ? unorderable_i32(key_ord(e[s])) : INT32_MIN;
            else        reinterpret_cast<float*>(out_scores)[(size_t)q * k_out + i] = ok ? unorderable_f32(key_ord(e[s])) : -INFINITY;
        }
    }
    if (lane == 0) out_counts[q] = count;
}

cudaError_t launch_merge_topk(cudaStream_t st, const uint32_t* ids, const void* scores, int is_int,
                              int n_lists, int nq, int k_in, int k_out,
                              uint32_t* out_ids, void* out_scores, int32_t* out_counts, bool sorted_input,
                              size_t list_stride) {
    if (nq <= 0) return cudaSuccess;
    if (list_stride == 0) list_stride = (size_t)nq * k_in;
    if (k_in > kMaxListLen || k_out > k_in || k_out <= 0 || n_lists <= 0) return cudaErrorInvalidValue;
    const int grid = (nq + kMergeWarps - 1) / kMergeWarps;
    const int threads = kMergeWarps * 32;
    if (sorted_input) {
        if (k_in <= 32) merge_topk_kernel<1, false><<<grid, threads, 0, st>>>(ids, scores, is_int, n_lists, nq, k_in, k_out, list_stride, out_ids, out_scores, out_counts);
        else            merge_topk_kernel<4, false><<<grid, threads, 0, st>>>(ids, scores, is_int, n_lists, nq, k_in, k_out, list_stride, out_ids, out_scores, out_counts);
    } else {
        if (k_in <= 32) merge_topk_kernel<1, true><<<grid, threads, 0, st>>>(ids, scores, is_int, n_lists, nq, k_in, k_out, list_stride, out_ids, out_scores, out_counts);
        else            merge_topk_kernel<4, true><<<grid, threads, 0, st>>>(ids, scores, is_int, n_lists, nq, k_in, k_out, list_stride, out_ids, out_scores, out_counts);
    }
    return cudaGetLastError();
}

}  // namespace crs
