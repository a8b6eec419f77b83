// K7x — cross-shard exchange + merge in ONE kernel over NVLink peer memory.
//
// No reference counterpart (the reference is single-node, rag/indexing.py); this is the multi-GPU step that
// follows each rank's local exact top-k.  The NCCL form of it (allgather of k*(id, score),
// then merge_topk) costs a collective launch plus a merge launch per search and dominates a sub-millisecond
// step; here every rank's kernel
//   (1) stores its own [nq, k] ids / raw scores straight into EVERY peer's receive buffer (plain global stores
//       on peer-mapped pointers: CUDA IPC between processes, cudaDeviceEnablePeerAccess inside one process),
//   (2) publishes a per-CTA flag (step stamp) to every peer after a system-scope fence,
//   (3) waits until the same CTA of every peer has published its flag for this step, and
//   (4) merges the G lists of its queries out of its own receive buffer (one warp per query).
// A CTA only ever waits for pushes of the SAME CTA index on other GPUs, and every CTA pushes before it
// waits, so there is no ordering between CTAs of one GPU to get wrong.  Receive slots are double-buffered by
// step parity: a rank can only be one step ahead of the slowest peer (it needs that peer's push to finish its
// own step), so step s+2 never overwrites what a peer is still merging for step s.
//
// The step stamp lives in device memory and is advanced by the search's first kernel (the query encode), so a
// CUDA-graph replay of the whole step works unchanged.
//
// A wait that is not satisfied within ~2 s (a peer died) sets the exchange's error word and lets the kernel
// finish; the host checks it (crs_exchange_status).
#include "common.cuh"
#include "crs_internal.h"

namespace crs {

constexpr int kXWarps = 4;              // queries per CTA (one warp each), as in merge.cu

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// phase: 0 = push + wait + merge, 1 = push only, 2 = wait + merge only
template <int LPL>
__global__ void __launch_bounds__(kXWarps * 32)
xmerge_kernel(XchgArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * kXWarps;
    const uint32_t step = *reinterpret_cast<volatile const uint32_t*>(a.step_word);
    const int par = (int)(step & 1u);
    const size_t nk = (size_t)a.max_nq * a.max_k;                 // words per block of a slot
    const size_t slot_words = 2 * nk;                             // ids block + scores block
    const size_t flags_per_par = (size_t)a.world * a.flag_ctas;

    if (a.phase != 2) {
        // ---- (1) push this CTA's queries' local lists into slot [par][rank] of every rank's receive buffer
        const int nq_here = min(kXWarps, a.nq - q0);
        const int words = nq_here * a.k;                          // per block (ids / scores)
        for (int p = 0; p < a.world; ++p) {
            uint32_t* dst = a.peer_slots[p] + ((size_t)par * a.world + a.rank) * slot_words;
            for (int i = threadIdx.x; i < words; i += blockDim.x) {
                const size_t o = (size_t)q0 * a.k + i;
                dst[o] = a.local_ids[o];
                dst[nk + o] = a.local_scores[o];
            }
        }
        __threadfence_system();
        __syncthreads();
        // ---- (2) flag: "CTA blockIdx.x of rank `rank` has pushed step `step`"
        if ((int)threadIdx.x < a.world) {
            uint32_t* f = a.peer_flags[threadIdx.x] + (size_t)par * flags_per_par + (size_t)a.rank * a.flag_ctas + blockIdx.x;
            st_release_sys_u32(f, step);
        }
        if (a.phase == 1) return;
    }

    // ---- (3) wait for the same CTA of every rank
    if ((int)threadIdx.x < a.world) {
        const uint32_t* f = a.my_flags + (size_t)par * flags_per_par + (size_t)threadIdx.x * a.flag_ctas + blockIdx.x;
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_sys_u32(f) != step) {
            if (globaltimer_ns() - t0 > 2000000000ull) { atomicExch(a.err_word, 1u); break; }
            __nanosleep(64);
        }
    }
    __syncthreads();

    // ---- (4) merge the world's lists of this warp's query
    const int q = q0 + warp;
    if (q >= a.nq) return;
    uint64_t e[LPL];
#pragma unroll
    for (int s = 0; s < LPL; ++s) e[s] = 0ull;
    for (int r = 0; r < a.world; ++r) {
        const uint32_t* slot = a.my_slots + ((size_t)par * a.world + r) * slot_words;
        uint64_t b[LPL];
#pragma unroll
        for (int s = 0; s < LPL; ++s) {
            const int i = lane * LPL + s;
            uint64_t key = 0ull;
            if (i < a.k) {
                const size_t o = (size_t)q * a.k + i;
                const uint32_t id = ld_relaxed_sys_u32(slot + o);             // written by a peer: not through L1
                if (id != CRS_PAD_ID) {
                    const uint32_t bits = ld_relaxed_sys_u32(slot + nk + o);
                    const uint32_t ord = a.is_int ? orderable_i32((int32_t)bits) : orderable_f32(__uint_as_float(bits));
                    key = make_key(ord, id);
                }
            }
            b[s] = key;
        }
        warp_merge_desc<LPL>(e, b, lane);
    }
    int nvalid = 0;
#pragma unroll
    for (int s = 0; s < LPL; ++s) nvalid += (e[s] != 0ull);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nvalid += __shfl_xor_sync(CRS_FULL_MASK, nvalid, off);
    const int count = min(nvalid, a.k);
#pragma unroll
    for (int s = 0; s < LPL; ++s) {
        const int i = lane * LPL + s;
        if (i < a.k) {
            const bool ok = i < count;
            a.out_ids[(size_t)q * a.k + i] = ok ? key_id(e[s]) : CRS_PAD_ID;
            if (a.is_int) reinterpret_cast<int32_t*>(a.out_scores)[(size_t)q * a.k + i] = ok ? unorderable_i32(key_ord(e[s])) : INT32_MIN;
            else          reinterpret_cast<float*>(a.out_scores)[(size_t)q * a.k + i] = ok ? unorderable_f32(key_ord(e[s])) : -INFINITY;
        }
    }
    if (lane == 0) a.out_counts[q] = count;
}

// ---- candidate rows by global id, read straight out of the owning shard's HBM over NVLink peer memory
// (replaces the MAX all-reduce that used to assemble [nq, m, row_bytes] candidate codes on every rank).
// One CTA per id; rows no shard owns (pad ids) are zero-filled.
__global__ void __launch_bounds__(64)
peer_gather_kernel(PeerShards sh, const uint32_t* __restrict__ ids, int n, uint4* __restrict__ out) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const uint32_t gid = ids[i];
    const int chunks = sh.row_bytes / 16;
    const uint4* src = nullptr;
    if (gid != CRS_PAD_ID) {
        for (int r = 0; r < sh.world; ++r) {
            const int64_t local = (int64_t)gid - (int64_t)sh.row_base[r];
            if (local >= 0 && local < sh.count[r]) {
                src = reinterpret_cast<const uint4*>(sh.codes[r]) + local * chunks;
                break;
            }
        }
    }
    for (int c = threadIdx.x; c < chunks; c += blockDim.x)
        out[(int64_t)i * chunks + c] = src ? src[c] : make_uint4(0u, 0u, 0u, 0u);
}

cudaError_t launch_peer_gather(cudaStream_t st, const PeerShards& sh, const uint32_t* ids, int n, void* out) {
    if (n <= 0) return cudaSuccess;
    peer_gather_kernel<<<n, 64, 0, st>>>(sh, ids, n, reinterpret_cast<uint4*>(out));
    return cudaGetLastError();
}

int xmerge_ctas(int nq) { return (nq + kXWarps - 1) / kXWarps; }

cudaError_t launch_xmerge(cudaStream_t st, const XchgArgs& a) {
    if (a.nq <= 0) return cudaSuccess;
    if (a.k > kMaxListLen || a.k <= 0 || a.world < 1 || a.world > kMaxWorld) return cudaErrorInvalidValue;
    const int grid = xmerge_ctas(a.nq);
    if (grid > a.flag_ctas) return cudaErrorInvalidValue;
    if (a.k <= 32) xmerge_kernel<1><<<grid, kXWarps * 32, 0, st>>>(a);
    else           xmerge_kernel<4><<<grid, kXWarps * 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace crs
