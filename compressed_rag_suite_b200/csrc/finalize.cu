// Finalize — per query: merge the per-CTA candidate lists, rescore exactly,
// certify, apply the threshold and emit the top-k.  Plus the exact fp64 fallback
// scan used when a float-store query cannot be certified.
//
// Why: the fast passes (K1 scan, K4 tcgen05 GEMM) rank rows on fp32-accumulated
// scores whose summation order differs from any CPU brute force.  The canonical
// result (oracle/search.py) is defined on fl32(fp64-sequential dot), so each query's
// M best rows by fast score are rescored here in fp64, sequentially over
// j = 0..Dp-1 (products of fp16/bf16 values are exact in fp64, so the value equals
// the CPU's bit for bit), then re-ranked.  The result is *certified* when no row
// outside the M candidates can reach the k-th exact score:
//      list not full                          -> nothing was cut
//      fast(M-th candidate) + eps < exact(k-th)   (or < threshold when count < k)
// with eps = Dp * 2^-24 * |q| * max|c| * scale bounding the fast pass' error.
// Otherwise flags[q] = 1 and the exact fallback (exact_scan + finalize mode 1)
// recomputes that query exhaustively in fp64.  Integer stores (I8/B1) are exact
// already and use mode 1 directly.
#include "common.cuh"
#include "crs_internal.h"

namespace crs {

// 4 warps per query: the CTA is latency-bound (list loads, then M threads rescoring serially in fp64), so what matters
// is how many queries are resident per SM — 128-thread CTAs fit 8+ per SM, 512-thread ones fit 2 (1024 queries: 55 -> ~15 us)
constexpr int kFinalizeWarpsMany = 4;    // many queries: occupancy
constexpr int kFinalizeWarpsFew = 16;    // a handful of queries: each CTA's own latency is what the caller waits for

template <bool BF16>
__device__ __forceinline__ void unpack8(const uint4& v, double (&o)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if constexpr (BF16) {
            o[2 * i] = (double)__uint_as_float(w[i] << 16);
            o[2 * i + 1] = (double)__uint_as_float(w[i] & 0xFFFF0000u);
        } else {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            o[2 * i] = (double)f.x;
            o[2 * i + 1] = (double)f.y;
        }
    }
}

// canonical score: sequential fp64 accumulation over j = 0..Dp-1, one rounding to fp32
template <bool BF16>
__device__ __forceinline__ float exact_dot(const uint4* __restrict__ row, const uint4* __restrict__ q, int chunks) {
    // rows are multiples of 128 bytes: 8 chunks (16 loads) are issued before their 64 serial FMAs, so a
    // row costs chunks/8 memory round trips instead of one per chunk
    double acc = 0.0;
    for (int c0 = 0; c0 < chunks; c0 += 8) {
        uint4 rv[8], qv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { rv[i] = row[c0 + i]; qv[i] = q[c0 + i]; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double a[8], b[8];
            unpack8<BF16>(rv[i], a);
            unpack8<BF16>(qv[i], b);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fma(a[j], b[j], acc);
        }
    }
    return (float)acc;
}

template <int LPL, int kFinalizeWarps>
__global__ void __launch_bounds__(kFinalizeWarps * 32)
finalize_kernel(FinalizeArgs a) {
    constexpr int M = 32 * LPL;
    __shared__ uint64_t stage[kFinalizeWarps * M];
    __shared__ uint64_t fast_keys[M];
    __shared__ uint64_t exact_keys[M];
    __shared__ uint64_t cut_stage[kFinalizeWarps];
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (a.only_flagged && a.flags[q] == 0) return;

    // 1. merge the n_lists sorted lists of this query (warp w takes lists w, w+16, ...)
    uint64_t e[LPL];
#pragma unroll
    for (int s = 0; s < LPL; ++s) e[s] = 0ull;
    const uint64_t* base = a.cand + (size_t)q * a.n_lists * a.list_stride;
    uint64_t cut = 0ull;        // largest key at which any input list was cut (0 = no list was full)
    // lists are fetched in batches of kBatch independent loads (one L2 round trip per batch, not per list)
    constexpr int kBatch = (LPL == 1) ? 8 : 2;
    for (int l0 = warp; l0 < a.n_lists; l0 += kFinalizeWarps * kBatch) {
        uint64_t b[kBatch][LPL];
        uint64_t last[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const int l = l0 + j * kFinalizeWarps;
            const bool have = l < a.n_lists;
#pragma unroll
            for (int s = 0; s < LPL; ++s) {
                const int i = lane * LPL + s;
                b[j][s] = (have && i < a.list_len) ? base[(size_t)l * a.list_stride + i] : 0ull;
            }
            last[j] = have ? base[(size_t)l * a.list_stride + a.list_len - 1] : 0ull;
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            cut = u64max(cut, last[j]);
            warp_merge_desc<LPL>(e, b[j], lane);
        }
    }
    if (lane == 0) cut_stage[warp] = cut;
    block_merge_lists<LPL, kFinalizeWarps>(e, stage, warp, lane);
    if (warp == 0) {
#pragma unroll
        for (int s = 0; s < LPL; ++s) fast_keys[lane * LPL + s] = e[s];
    }
    __syncthreads();

    // 2. exact rescoring, one thread per candidate (float stores only).  The candidate rows are scattered
    // over HBM: when they fit, the whole CTA first copies them (and the query) into shared memory with
    // independent coalesced loads — one memory round trip — and the serial fp64 chains then run from there
    // (row stride + 16 bytes: the per-thread 16-byte walks of consecutive rows hit different banks).
    if (a.mode == 0) {
        extern __shared__ __align__(16) uint8_t rows_sm[];
        const int chunks = a.dim_padded / 8;
        const int stride16 = chunks + 1;                       // in 16-byte units
        const uint4* qglob = reinterpret_cast<const uint4*>(a.qcodes) + (size_t)q * chunks;
        // Candidates more than 2 eps below the k-th best FAST score cannot reach the exact top-k: their exact
        // score is below fast_k - eps, and each of the k best-by-fast rows scores at least that exactly.  (If one
        // of those k fails the threshold, so does everything below it.)  Their rows are neither fetched nor
        // rescored: about k + a few random row reads per query instead of M.
        float skip_below = -INFINITY;
        if (a.k <= M && fast_keys[a.k - 1] != 0ull)
            skip_below = unorderable_f32(key_ord(fast_keys[a.k - 1])) - 2.0f * a.eps_rel * a.qnorms[q] * a.row_norm_bound;
        if (a.stage_rows) {
            uint4* sm = reinterpret_cast<uint4*>(rows_sm);
            for (int i = threadIdx.x; i < (M + 1) * chunks; i += blockDim.x) {
                const int r = i / chunks, c = i - r * chunks;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (r == M) v = qglob[c];                      // last slot: the query
                else if (fast_keys[r] != 0ull && unorderable_f32(key_ord(fast_keys[r])) >= skip_below)
                    v = (reinterpret_cast<const uint4*>(a.codes) + (size_t)key_id(fast_keys[r]) * chunks)[c];
                sm[(size_t)r * stride16 + c] = v;
            }
            __syncthreads();
        }
        if (threadIdx.x < M) {
            const uint64_t fk = fast_keys[threadIdx.x];
            uint64_t ek = 0ull;
            float err = 0.f;
            if (fk != 0ull && unorderable_f32(key_ord(fk)) >= skip_below) {
                const uint32_t id = key_id(fk);
                const uint4* row = a.stage_rows ? reinterpret_cast<const uint4*>(rows_sm) + (size_t)threadIdx.x * stride16
                                                : reinterpret_cast<const uint4*>(a.codes) + (size_t)id * chunks;
                const uint4* qv = a.stage_rows ? reinterpret_cast<const uint4*>(rows_sm) + (size_t)M * stride16 : qglob;
                const float s = a.bf16 ? exact_dot<true>(row, qv, chunks) : exact_dot<false>(row, qv, chunks);
                if (s >= a.min_similarity) ek = make_key(orderable_f32(s), id);
                err = fabsf(unorderable_f32(key_ord(fk)) - s);
            }
            exact_keys[threadIdx.x] = ek;
            // statistic: largest fast-vs-exact score gap seen (evidence for the eps bound)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) err = fmaxf(err, __shfl_xor_sync(CRS_FULL_MASK, err, off));
            unsigned* slot = reinterpret_cast<unsigned*>(a.n_flagged + 2);
            if (lane == 0 && __float_as_uint(err) > *reinterpret_cast<volatile unsigned*>(slot))
                atomicMax(slot, __float_as_uint(err));
        }
        __syncthreads();
    }

    // 3. warp 0: final order, certification, output
    if (warp != 0) return;
    uint64_t x[LPL];
    const uint64_t* srck = (a.mode == 0) ? exact_keys : fast_keys;
#pragma unroll
    for (int s = 0; s < LPL; ++s) x[s] = srck[lane * LPL + s];
    if (a.mode == 0) warp_sort_desc<LPL>(x, lane);

    // number of valid keys (sorted desc, empties (0) at the end)
    int nvalid = 0;
#pragma unroll
    for (int s = 0; s < LPL; ++s) nvalid += (x[s] != 0ull);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nvalid += __shfl_xor_sync(CRS_FULL_MASK, nvalid, off);
    const int count = min(nvalid, a.k);

    if (a.mode == 0) {
        // everything dropped before this point scored (fast) no more than the largest cut:
        // the M-th merged candidate when the merged list is full, or the last entry of any
        // input list that was full
        uint64_t last_fast = fast_keys[M - 1];
#pragma unroll
        for (int w = 0; w < kFinalizeWarps; ++w) last_fast = u64max(last_fast, cut_stage[w]);
        // the fast pass may have started from a sampled per-query floor: everything below it was dropped too
        if (a.tau_q != nullptr) {
            const uint32_t o = a.tau_q[q];
            if (o != 0u) last_fast = u64max(last_fast, make_key(o, 0u));
        }
        bool certified = true;
        if (last_fast != 0ull) {
            const float a_min = unorderable_f32(key_ord(last_fast));
            const float eps = a.eps_rel * a.qnorms[q] * a.row_norm_bound;
            float bound;
            if (nvalid >= a.k) {
                // k-th exact score: element k-1 of x
                const int kk = a.k - 1;
                uint64_t kth = 0ull;
#pragma unroll
                for (int s = 0; s < LPL; ++s) {
                    const uint64_t v = shfl_u64(x[s], kk / LPL);
                    if (s == kk % LPL) kth = v;
                }
                bound = unorderable_f32(key_ord(kth));
            } else {
                bound = a.min_similarity;                     // fewer than k pass: nothing cut may pass either
            }
            certified = (a_min + eps < bound);
        }
        if (lane == 0) {
            a.flags[q] = certified ? 0 : 1;
            if (!certified) { atomicAdd(a.n_flagged, 1); atomicAdd(a.n_flagged + 1, 1); }
        }
    } else if (a.certify_exact) {
        // exact keys (integer stores through the tensor-core path, where each slice list keeps
        // fewer than k keys): everything a full list dropped sorts below that list's last key,
        // so the top-k is final iff the k-th key beats the largest such cut
        uint64_t cutmax = 0ull;
#pragma unroll
        for (int w = 0; w < kFinalizeWarps; ++w) cutmax = u64max(cutmax, cut_stage[w]);
        bool certified = true;
        if (cutmax != 0ull) {
            certified = false;
            if (nvalid >= a.k) {
                const int kk = a.k - 1;
                uint64_t kth = 0ull;
#pragma unroll
                for (int s = 0; s < LPL; ++s) {
                    const uint64_t v = shfl_u64(x[s], kk / LPL);
                    if (s == kk % LPL) kth = v;
                }
                certified = kth > cutmax;
            }
        }
        if (lane == 0) {
            a.flags[q] = certified ? 0 : 1;
            if (!certified) { atomicAdd(a.n_flagged, 1); atomicAdd(a.n_flagged + 1, 1); }
        }
    } else if (a.only_flagged && lane == 0) {
        a.flags[q] = 0;
    }

#pragma unroll
    for (int s = 0; s < LPL; ++s) {
        const int i = lane * LPL + s;
        if (i < a.k) {
            const bool ok = i < count;
            const uint64_t key = x[s];
            uint32_t gid = CRS_PAD_ID;
            if (ok) gid = a.id_map ? a.id_map[key_id(key)] : key_id(key) + a.row_base;
            a.out_ids[(size_t)q * a.k + i] = gid;
            if (a.is_int) {
                reinterpret_cast<int32_t*>(a.out_scores)[(size_t)q * a.k + i] =
                    ok ? unorderable_i32(key_ord(key)) : INT32_MIN;
            } else {
                reinterpret_cast<float*>(a.out_scores)[(size_t)q * a.k + i] =
                    ok ? unorderable_f32(key_ord(key)) : -INFINITY;
            }
        }
    }
    if (lane == 0) a.out_counts[q] = count;
}

cudaError_t launch_finalize(cudaStream_t st, const FinalizeArgs& a_in) {
    FinalizeArgs a = a_in;
    if (a.nq <= 0) return cudaSuccess;
    if (a.k > 32 * a.lpl) return cudaErrorInvalidValue;
    // shared-memory staging of the candidate rows (mode 0): M rows + the query, padded stride
    const int M = 32 * a.lpl;
    const size_t row_bytes = (size_t)a.dim_padded * 2;
    size_t smem = (a.mode == 0) ? (size_t)(M + 1) * (row_bytes + 16) : 0;
    if (smem > 110 * 1024) smem = 0;                       // wide rows x 128 candidates: read straight from HBM
    a.stage_rows = smem > 0 ? 1 : 0;
    const bool few = a.nq <= 32;
#define CRS_FINALIZE(LPL_, NW_)                                                                                  \
    do {                                                                                                         \
        if (smem > 48 * 1024) {                                                                                  \
            cudaError_t e = cudaFuncSetAttribute(finalize_kernel<LPL_, NW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                      \
        }                                                                                                        \
        finalize_kernel<LPL_, NW_><<<a.nq, NW_ * 32, smem, st>>>(a);                                              \
    } while (0)
    if (a.lpl == 1) { if (few) CRS_FINALIZE(1, kFinalizeWarpsFew); else CRS_FINALIZE(1, kFinalizeWarpsMany); }
    else if (a.lpl == 4) { if (few) CRS_FINALIZE(4, kFinalizeWarpsFew); else CRS_FINALIZE(4, kFinalizeWarpsMany); }
    else return cudaErrorInvalidValue;
#undef CRS_FINALIZE
    return cudaGetLastError();
}

// ------------------------------------------------------------------ exact fallback scan
// Thread per row, fp64 sequential — slow by design (rows are walked uncoalesced), only
// run for the rare queries finalize could not certify.  One sorted list per CTA and
// flagged query, keys carry the exact fl32 score, so finalize mode 1 just merges.
constexpr int kExactWarps = 8;

// canonical 32-bit orderable score of one stored row against one stored query (all stores)
template <int STORE>
__device__ __forceinline__ uint32_t exact_ord(const uint4* __restrict__ row, const uint4* __restrict__ q, int chunks,
                                              int dim, float min_similarity, uint32_t ord_min, bool& pass) {
    if constexpr (STORE == CRS_F16 || STORE == CRS_BF16) {
        const float s = exact_dot<STORE == CRS_BF16>(row, q, chunks);
        pass = s >= min_similarity;
        return orderable_f32(s);
    } else {
        int acc = 0;
        for (int c = 0; c < chunks; ++c) {
            const uint4 a = row[c], b = q[c];
            if constexpr (STORE == CRS_I8) {
                acc = __dp4a((int)a.x, (int)b.x, acc); acc = __dp4a((int)a.y, (int)b.y, acc);
                acc = __dp4a((int)a.z, (int)b.z, acc); acc = __dp4a((int)a.w, (int)b.w, acc);
            } else {
                acc += __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z) + __popc(a.w ^ b.w);
            }
        }
        if constexpr (STORE == CRS_B1) acc = dim - 2 * acc;
        const uint32_t o = orderable_i32(acc);
        pass = o >= ord_min;
        return o;
    }
}

template <int LPL, int STORE>
__global__ void __launch_bounds__(kExactWarps * 32)
exact_scan_kernel(const uint4* __restrict__ codes, int64_t n_rows, int chunks, int dim, const uint4* __restrict__ qcodes,
                  int nq, const int32_t* __restrict__ flags, float min_similarity, uint32_t ord_min,
                  uint64_t* __restrict__ cand, const uint32_t* __restrict__ allow, const int32_t* __restrict__ n_flagged) {
    constexpr int M = 32 * LPL;
    // the usual case: finalize certified every query of this search -> nothing to do
    if (n_flagged != nullptr && *n_flagged == 0) return;
    __shared__ uint64_t stage[kExactWarps * M];
    __shared__ int s_list[1024];
    __shared__ int s_cnt;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // find the flagged queries with one strided pass instead of walking nq flags serially
    for (int base = 0; base < nq; base += 1024) {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < min(1024, nq - base); i += blockDim.x)
        if (flags[base + i] != 0) s_list[atomicAdd(&s_cnt, 1)] = base + i;
    __syncthreads();
    const int n_fl = s_cnt;
    for (int j = 0; j < n_fl; ++j) {
        const int q = s_list[j];
        WarpTopM<LPL> top; top.init();
        const uint4* qv = qcodes + (size_t)q * chunks;
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        for (int64_t r0 = (int64_t)blockIdx.x * blockDim.x + warp * 32; r0 < n_rows; r0 += stride) {
            const int64_t row = r0 + lane;
            uint64_t key = 0ull;
            if (row < n_rows && (allow == nullptr || ((allow[row >> 5] >> (row & 31)) & 1u))) {
                bool pass;
                const uint32_t o = exact_ord<STORE>(codes + row * chunks, qv, chunks, dim, min_similarity, ord_min, pass);
                if (pass) key = make_key(o, (uint32_t)row);
            }
            unsigned bal = __ballot_sync(CRS_FULL_MASK, key > top.floor_key);
            while (bal) {
                const int src = __ffs(bal) - 1;
                bal &= bal - 1;
                const uint64_t kb = shfl_u64(key, src);
                if (kb > top.floor_key) top.insert(kb, lane);
            }
        }
        warp_sort_desc<LPL>(top.e, lane);
        block_merge_lists<LPL, kExactWarps>(top.e, stage, warp, lane);
        if (warp == 0) {
#pragma unroll
            for (int s = 0; s < LPL; ++s)
                cand[((size_t)q * gridDim.x + blockIdx.x) * M + lane * LPL + s] = top.e[s];
        }
        __syncthreads();
    }
    __syncthreads();
    }
}

// Serial order of the flagged list (s_list is filled with atomics) does not matter: every
// flagged query writes its own lists.
cudaError_t launch_exact_scan(cudaStream_t st, const void* codes, int64_t n, int row_bytes, int dim, crs_dtype store,
                              const void* qcodes, int nq, const int32_t* flags, float min_similarity, int32_t min_raw,
                              uint64_t* cand, const ScanPlan& plan, const int32_t* n_flagged) {
    if (nq <= 0) return cudaSuccess;
    const int chunks = row_bytes / 16;
    const uint4* c = reinterpret_cast<const uint4*>(codes);
    const uint4* qv = reinterpret_cast<const uint4*>(qcodes);
    const int threads = kExactWarps * 32;
    const uint32_t ord_min = orderable_i32(min_raw);
#define CRS_EXACT(LPL_, S_) exact_scan_kernel<LPL_, S_><<<plan.grid, threads, 0, st>>>(c, n, chunks, dim, qv, nq, flags, \
                                min_similarity, ord_min, cand, plan.allow, n_flagged)
    if (plan.lpl == 1) {
        switch (store) {
            case CRS_F16: CRS_EXACT(1, CRS_F16); break;
            case CRS_BF16: CRS_EXACT(1, CRS_BF16); break;
            case CRS_I8: CRS_EXACT(1, CRS_I8); break;
            default: CRS_EXACT(1, CRS_B1); break;
        }
    } else {
        switch (store) {
            case CRS_F16: CRS_EXACT(4, CRS_F16); break;
            case CRS_BF16: CRS_EXACT(4, CRS_BF16); break;
            case CRS_I8: CRS_EXACT(4, CRS_I8); break;
            default: CRS_EXACT(4, CRS_B1); break;
        }
    }
#undef CRS_EXACT
    return cudaGetLastError();
}

// ------------------------------------------------------------------ K8: candidate rescoring
// Canonical score of given (query, row) pairs: fl32 of the fp64-sequential dot for the float
// stores, exact int32 dot for I8, dim - 2*hamming for B1 — the same values a search on this
// index reports.  One thread per pair (rows are walked uncoalesced: this is a latency-bound
// pass over <= 128 candidates per query).  Rows this shard does not own, and pad ids, get the
// "absent" score (-inf / INT32_MIN) so that a MAX reduction over the shards assembles the row.
// ids == NULL: `codes` holds the nq*m candidate rows themselves ([nq][m], freshly encoded).
template <int STORE>
__global__ void __launch_bounds__(128)
score_rows_kernel(const uint8_t* __restrict__ codes, int64_t n_rows, uint32_t row_base, int row_bytes, int dim,
                  const uint8_t* __restrict__ qcodes, const uint32_t* __restrict__ ids, int nq, int m,
                  void* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)nq * m) return;
    const int q = (int)(i / m);
    const uint32_t id = ids ? ids[i] : (uint32_t)i + row_base;      // NULL ids: pair i scores row i of `codes`
    const int64_t local = (int64_t)id - (int64_t)row_base;
    const bool have = (id != CRS_PAD_ID) && local >= 0 && local < n_rows;
    const uint4* row = reinterpret_cast<const uint4*>(codes + (size_t)(have ? local : 0) * row_bytes);
    const uint4* qv = reinterpret_cast<const uint4*>(qcodes + (size_t)q * row_bytes);
    const int chunks = row_bytes / 16;
    if constexpr (STORE == CRS_F16 || STORE == CRS_BF16) {
        float s = -INFINITY;
        if (have) s = exact_dot<STORE == CRS_BF16>(row, qv, chunks);
        reinterpret_cast<float*>(out)[i] = s;
    } else {
        int32_t s = INT32_MIN;
        if (have) {
            int acc = 0;
            for (int c = 0; c < chunks; ++c) {
                const uint4 a = row[c], b = qv[c];
                if constexpr (STORE == CRS_I8) {
                    acc = __dp4a((int)a.x, (int)b.x, acc); acc = __dp4a((int)a.y, (int)b.y, acc);
                    acc = __dp4a((int)a.z, (int)b.z, acc); acc = __dp4a((int)a.w, (int)b.w, acc);
                } else {
                    acc += __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z) + __popc(a.w ^ b.w);
                }
            }
            s = (STORE == CRS_B1) ? dim - 2 * acc : acc;
        }
        reinterpret_cast<int32_t*>(out)[i] = s;
    }
}

cudaError_t launch_score_rows(cudaStream_t st, const void* codes, int64_t n_rows, uint32_t row_base, int row_bytes,
                              int dim, crs_dtype store, const void* qcodes, const uint32_t* ids, int nq, int m,
                              void* out) {
    const int64_t total = (int64_t)nq * m;
    if (total <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((total + 127) / 128);
    const uint8_t* c = reinterpret_cast<const uint8_t*>(codes);
    const uint8_t* qc = reinterpret_cast<const uint8_t*>(qcodes);
    switch (store) {
        case CRS_F16:  score_rows_kernel<CRS_F16><<<grid, 128, 0, st>>>(c, n_rows, row_base, row_bytes, dim, qc, ids, nq, m, out); break;
        case CRS_BF16: score_rows_kernel<CRS_BF16><<<grid, 128, 0, st>>>(c, n_rows, row_base, row_bytes, dim, qc, ids, nq, m, out); break;
        case CRS_I8:   score_rows_kernel<CRS_I8><<<grid, 128, 0, st>>>(c, n_rows, row_base, row_bytes, dim, qc, ids, nq, m, out); break;
        case CRS_B1:   score_rows_kernel<CRS_B1><<<grid, 128, 0, st>>>(c, n_rows, row_base, row_bytes, dim, qc, ids, nq, m, out); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace crs
