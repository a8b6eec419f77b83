// Shared device helpers for the retrieval kernels (sm_100a only).
//
// Ordering convention used everywhere: a candidate is one 64-bit key
//     key = (orderable32(raw score) << 32) | (0xFFFFFFFF - row id)
// so that a plain unsigned "larger key wins" implements the reference-facing
// order "score descending, ties -> lowest id" (SURVEY.md §7 hard-part 3).
// key == 0 is the empty slot (row id 0xFFFFFFFF is the pad id and never stored).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define CRS_FULL_MASK 0xFFFFFFFFu

namespace crs {

// ---------------------------------------------------------------- keys
__host__ __device__ __forceinline__ uint32_t orderable_f32(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float unorderable_f32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint32_t orderable_i32(int32_t v) { return (uint32_t)v ^ 0x80000000u; }
__host__ __device__ __forceinline__ int32_t unorderable_i32(uint32_t o) { return (int32_t)(o ^ 0x80000000u); }

__host__ __device__ __forceinline__ uint64_t make_key(uint32_t ord, uint32_t id) {
    return ((uint64_t)ord << 32) | (uint64_t)(0xFFFFFFFFu - id);
}
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)k; }
__host__ __device__ __forceinline__ uint32_t key_ord(uint64_t k) { return (uint32_t)(k >> 32); }

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    return __shfl_sync(CRS_FULL_MASK, v, src);
}
__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
    return __shfl_xor_sync(CRS_FULL_MASK, v, m);
}
__device__ __forceinline__ uint64_t u64max(uint64_t a, uint64_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint64_t u64min(uint64_t a, uint64_t b) { return a < b ? a : b; }

// ------------------------------------------------- warp-distributed top-M lists
// A list of M = 32*LPL keys lives in registers: element i sits in lane i / LPL,
// slot i % LPL ("blocked" layout), so compare-exchange distances < LPL stay in a
// thread and the rest are one shuffle.

// One bitonic compare-exchange stage at distance J inside blocks of size K
// (final order: descending).
template <int LPL, int K, int J>
__device__ __forceinline__ void bitonic_stage(uint64_t (&e)[LPL], int lane) {
    if constexpr (J >= LPL) {
        constexpr int LJ = J / LPL;
#pragma unroll
        for (int s = 0; s < LPL; ++s) {
            const int i = lane * LPL + s;
            const uint64_t o = shfl_xor_u64(e[s], LJ);
            const bool desc = ((i & K) == 0);
            const bool lower = ((i & J) == 0);
            e[s] = (lower == desc) ? u64max(e[s], o) : u64min(e[s], o);
        }
    } else {
#pragma unroll
        for (int s = 0; s < LPL; ++s) {
            if ((s & J) == 0) {
                const int i = lane * LPL + s;
                const bool desc = ((i & K) == 0);
                const uint64_t a = e[s], b = e[s ^ J];
                const uint64_t hi = u64max(a, b), lo = u64min(a, b);
                e[s] = desc ? hi : lo;
                e[s ^ J] = desc ? lo : hi;
            }
        }
    }
}

template <int LPL, int K, int J>
struct BitonicInner {
    static __device__ __forceinline__ void run(uint64_t (&e)[LPL], int lane) {
        bitonic_stage<LPL, K, J>(e, lane);
        if constexpr (J > 1) BitonicInner<LPL, K, J / 2>::run(e, lane);
    }
};
template <int LPL, int K>
struct BitonicOuter {
    static __device__ __forceinline__ void run(uint64_t (&e)[LPL], int lane) {
        if constexpr (K > 2) BitonicOuter<LPL, K / 2>::run(e, lane);
        BitonicInner<LPL, K, K / 2>::run(e, lane);
    }
};

// Full sort, descending, of the 32*LPL keys held by the warp.
template <int LPL>
__device__ __forceinline__ void warp_sort_desc(uint64_t (&e)[LPL], int lane) {
    BitonicOuter<LPL, 32 * LPL>::run(e, lane);
}

// a, b sorted descending -> a := the 32*LPL largest of a U b, sorted descending.
template <int LPL>
__device__ __forceinline__ void warp_merge_desc(uint64_t (&a)[LPL], const uint64_t (&b)[LPL], int lane) {
#pragma unroll
    for (int s = 0; s < LPL; ++s) {
        const uint64_t o = shfl_u64(b[LPL - 1 - s], 31 - lane);     // b reversed
        a[s] = u64max(a[s], o);                                     // bitonic sequence of the top half
    }
    // K = 64*LPL is larger than every index, so every block sorts descending.
    BitonicInner<LPL, 64 * LPL, 16 * LPL>::run(a, lane);
}

// Unsorted running top-M with a warp-uniform floor ("replace the minimum").
template <int LPL>
struct WarpTopM {
    uint64_t e[LPL];
    uint64_t floor_key;      // min over all 32*LPL entries, warp-uniform

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < LPL; ++s) e[s] = 0ull;
        floor_key = 0ull;
    }
    // key is warp-uniform and > floor_key
    __device__ __forceinline__ void insert(uint64_t key, int lane) {
        unsigned hit = 0;
#pragma unroll
        for (int s = 0; s < LPL; ++s) hit |= (e[s] == floor_key) ? (1u << s) : 0u;
        const unsigned owners = __ballot_sync(CRS_FULL_MASK, hit != 0);
        const int owner = __ffs(owners) - 1;
        if (lane == owner) {
            const int slot = __ffs(hit) - 1;
#pragma unroll
            for (int s = 0; s < LPL; ++s) if (s == slot) e[s] = key;
        }
        uint64_t m = e[0];
#pragma unroll
        for (int s = 1; s < LPL; ++s) m = u64min(m, e[s]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = u64min(m, shfl_xor_u64(m, off));
        floor_key = m;
    }
};

// Block-level reduction of per-warp sorted lists through shared memory.
// stage: [NWARPS][32*LPL] keys. On return warp 0 holds the block's top-M (sorted).
// All NWARPS warps must call; NWARPS is a power of two.
// Synchronises on named barrier 1 over exactly NWARPS*32 threads, so other warps
// of the block (e.g. a copy-producer warp) need not take part.
template <int NTHREADS>
__device__ __forceinline__ void named_barrier_1() {
    asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
}

template <int LPL, int NWARPS>
__device__ __forceinline__ void block_merge_lists(uint64_t (&e)[LPL], uint64_t* stage, int warp, int lane) {
    constexpr int M = 32 * LPL;
#pragma unroll
    for (int half = NWARPS / 2; half >= 1; half >>= 1) {
        if (warp >= half && warp < 2 * half) {
#pragma unroll
            for (int s = 0; s < LPL; ++s) stage[warp * M + lane * LPL + s] = e[s];
        }
        named_barrier_1<NWARPS * 32>();
        if (warp < half) {
            uint64_t b[LPL];
#pragma unroll
            for (int s = 0; s < LPL; ++s) b[s] = stage[(warp + half) * M + lane * LPL + s];
            warp_merge_desc<LPL>(e, b, lane);
        }
        named_barrier_1<NWARPS * 32>();
    }
}

// ---------------------------------------------------------------- PTX: mbarrier + bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}
// 1-D bulk async copy global -> shared (UBLKCP), completion counted on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

}  // namespace crs
