// K4 — batched search as a dense contraction S = Q * C^T on the tcgen05 tensor cores
// with a fused threshold + top-L epilogue: the score matrix never reaches HBM.
//
// Replaces, for a whole batch of queries, the distance pass `collection.query` performs
// inside ChromaDB (reference rag/indexing.py:171-176; the reference itself can only ask
// one query at a time).  FLOPs per launch: 2 * nq_padded * n_rows * Dp.
//
// Work split: CTA = (query tile of 128 queries, contiguous slice of corpus tiles).
// blockIdx = slice * n_qtiles + qtile, so the CTAs that stream the same corpus slice are
// launched next to each other and share it through L2.
//
// Roles inside a CTA (192 threads, 1 CTA / SM):
//   warp 0  TMA producer : loads the CTA's query tile once (A operand, stays resident in
//           smem: KCH chunks of [128 x 64] fp16, SWIZZLE_128B), then streams corpus tiles
//           as [256 rows x 64 k] chunks (B operand) through a STAGES-deep mbarrier ring.
//   warp 1  MMA issuer   : one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//           (M=128, N=256, K=16) 4 per chunk, KCH chunks per corpus tile, accumulating
//           in TMEM; tcgen05.commit frees the smem slot / publishes the accumulator.
//           Accumulators are double-buffered (2 x 256 TMEM columns).
//   warps 2-5 epilogue   : TMEM lane = query, column = corpus row.  Each thread reads its
//           query's 256 scores in 32-column slabs (tcgen05.ld.32x32b.x32), tests the slab
//           maximum against its running threshold (the L-th best so far), and only on a
//           hit walks the slab and inserts into a sorted register list of L keys.
// At the end every thread writes its sorted list: cand[q][slice][0..L) (stride M = 32).
// finalize.cu merges the slices' lists, rescores the candidates exactly in fp64,
// certifies the top-k and falls back to the exhaustive fp64 pass when it cannot.
#include <cuda.h>

#include "common.cuh"
#include "crs_internal.h"

namespace crs {

#ifdef CRS_GEMM_PROFILE
// dev instrumentation: cycles the MMA issuer spends waiting for operands / for a drained accumulator
__device__ unsigned long long g_gemm_prof[8];
#endif

constexpr int kGemmThreads = 192;
constexpr int kTileQ = 128;        // UMMA M
constexpr int kTileC = 256;        // UMMA N (corpus rows per tile)
constexpr int kChunkK = 64;        // fp16 elements per 128-byte swizzle row
constexpr int kStages = 4;
constexpr int kAChunkBytes = kTileQ * kChunkK * 2;     // 16 KB
constexpr int kBStageBytes = kTileC * kChunkK * 2;     // 32 KB

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// multicast variant: the box lands at the same smem offset in every CTA of `mask`, and each of
// those CTAs' mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
// L2 prefetch of one TMA box (no shared memory, no barrier): issued a couple of corpus tiles ahead so the
// ring's own loads find their data in L2 — half of them used to go to DRAM (ncu: 51 % L2 hit rate) with
// a latency the 4-stage ring cannot cover.
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_512(uint32_t* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// kind::i8: signed 8-bit operands, int32 accumulators, K = 32 per instruction (K5)
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at this offset in every CTA of `mask` once the MMAs issued so far retire
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: one MMA spans the two SMs of a cluster; each CTA stages its own
// 128 queries (A) and HALF of the corpus tile (B), so every SM reads half the B bytes from shared memory
// per FLOP.  All barriers the issuing (rank-0) CTA waits on live in rank 0.
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t smem_addr) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(smem_addr));
    return r;
}
// TMA load into THIS CTA's smem whose completion bytes are counted on a barrier of the pair's rank-0 CTA
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_rank0) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_rank0), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_512_pair(uint32_t* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this offset in both CTAs of the pair once the pair-MMAs issued so far retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row groups
// 1024 bytes apart (SBO), descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

// sorted (descending) register list of L keys; insert keeps the L largest
template <int L>
__device__ __forceinline__ void list_insert(uint64_t (&a)[L], uint64_t x) {
#pragma unroll
    for (int i = 0; i < L; ++i) {
        const uint64_t hi = u64max(a[i], x);
        x = u64min(a[i], x);
        a[i] = hi;
    }
}

// One thread's 32 scores (consecutive corpus rows row0..row0+31 of its query).  Fast path:
// tree max against the running threshold.  Slow path (rare once the list is warm): a hit mask
// and ONE copy of the insert code, walked only over this thread's own hits, so a warp runs
// max-over-lanes(hits) insert bodies instead of one per column any lane hit.
template <bool INT> struct ScoreT { using type = float; };
template <> struct ScoreT<true> { using type = int32_t; };

template <bool INT>
__device__ __forceinline__ typename ScoreT<INT>::type score_of(uint32_t bits) {
    if constexpr (INT) return (int32_t)bits; else return __uint_as_float(bits);
}
template <bool INT>
__device__ __forceinline__ uint32_t ord_of(typename ScoreT<INT>::type v) {
    if constexpr (INT) return orderable_i32(v); else return orderable_f32(v);
}
template <bool INT>
__device__ __forceinline__ typename ScoreT<INT>::type from_ord(uint32_t o) {
    if constexpr (INT) return unorderable_i32(o); else return unorderable_f32(o);
}
template <typename T> __device__ __forceinline__ T smax(T a, T b) { return a > b ? a : b; }
template <> __device__ __forceinline__ float smax<float>(float a, float b) { return fmaxf(a, b); }

template <int L, bool INT>
__device__ __forceinline__ void slab_scan(const uint32_t (&r)[32], int64_t row0, int64_t n_rows,
                                          typename ScoreT<INT>::type tau_pre, typename ScoreT<INT>::type& tau,
                                          uint64_t (&best)[L], const uint32_t* __restrict__ allow) {
    using T = typename ScoreT<INT>::type;
    // maximum of the 32 scores as a tree of 3-input max (FMNMX3 / VIMNMX3 on sm_100): 16 instructions
    T m[11];
#pragma unroll
    for (int i = 0; i < 10; ++i)
        m[i] = smax<T>(smax<T>(score_of<INT>(r[3 * i]), score_of<INT>(r[3 * i + 1])), score_of<INT>(r[3 * i + 2]));
    m[10] = smax<T>(score_of<INT>(r[30]), score_of<INT>(r[31]));
    m[0] = smax<T>(smax<T>(m[0], m[1]), m[2]);
    m[3] = smax<T>(smax<T>(m[3], m[4]), m[5]);
    m[6] = smax<T>(smax<T>(m[6], m[7]), m[8]);
    m[9] = smax<T>(m[9], m[10]);
    m[0] = smax<T>(smax<T>(m[0], m[3]), m[6]);
    m[0] = smax<T>(m[0], m[9]);
    if (m[0] >= tau) {
        unsigned mask = 0;
        T tmp[32];                           // dynamically indexed -> local memory, touched on this path only
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const T v = score_of<INT>(r[i]);
            tmp[i] = v;
            mask |= (v >= tau) ? (1u << i) : 0u;
        }
        while (mask) {
            const int i = __ffs(mask) - 1;
            mask &= mask - 1;
            const T v = tmp[i];
            const int64_t row = row0 + i;
            // row bitmap of a where / where_document filter: looked up for hits only
            if (v >= tau && row < n_rows && (allow == nullptr || ((allow[row >> 5] >> (row & 31)) & 1u))) {
                const uint64_t key = make_key(ord_of<INT>(v), (uint32_t)row);
                if (key > best[L - 1]) {
                    list_insert<L>(best, key);
                    if (best[L - 1] != 0ull) tau = smax<T>(tau_pre, from_ord<INT>(key_ord(best[L - 1])));
                }
            }
        }
    }
}

template <int KCH, int L, int CS, bool INT, bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c,
                 int64_t n_rows, int n_qtiles, int n_slices, uint32_t idesc, uint32_t tau_pre_bits,
                 uint64_t* __restrict__ cand, int nq, int list_stride, const uint32_t* __restrict__ allow,
                 const uint32_t* __restrict__ tau_q, int prefetch_tiles) {
    static_assert(!PAIR || CS == 2, "the CTA-pair MMA needs clusters of exactly two CTAs");
    // pair mode: a stage holds this CTA's half (128 rows) of a corpus chunk -> twice the stages in the same smem
    constexpr int STAGES = PAIR ? 2 * kStages : kStages;
    constexpr int BSTAGE = PAIR ? kBStageBytes / 2 : kBStageBytes;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[2 * kStages], empty_bar[2 * kStages], a_bar, tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                   // KCH x 16 KB
    uint8_t* smem_b = smem + KCH * kAChunkBytes;              // STAGES x BSTAGE (128 KB)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // a cluster = CS query tiles working on the same corpus slice; its CTAs split every corpus
    // chunk CS ways and multicast the pieces to each other (one L2 read feeds CS SMs)
    const int crank = (CS > 1) ? (int)cluster_ctarank() : 0;
    const int cluster_id = blockIdx.x / CS;
    const int qgroups = n_qtiles / CS;                      // n_qtiles is a multiple of CS
    const int qtile = (cluster_id % qgroups) * CS + crank;
    const int slice = cluster_id / qgroups;
    constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1);
    const int64_t tiles_total = (n_rows + kTileC - 1) / kTileC;
    const int64_t tile_lo = tiles_total * slice / n_slices;
    const int64_t tile_hi = tiles_total * (slice + 1) / n_slices;
    const int n_tiles = (int)(tile_hi - tile_lo);

    if (threadIdx.x == 0) {
        // pair mode: one commit (multicast to both CTAs) frees a stage; the accumulator-drained barrier of
        // rank 0 collects the 4 epilogue warps of BOTH CTAs
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], PAIR ? 1 : CS); }
        mbar_init(&a_bar, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], PAIR ? 8 : 4); }
        fence_mbar_init();
    }
    if (warp == 1) { if constexpr (PAIR) tmem_alloc_512_pair(&tmem_base_slot); else tmem_alloc_512(&tmem_base_slot); }
    tc_fence_before();
    __syncthreads();
    if constexpr (CS > 1) cluster_sync_all();               // peers' barriers exist before anything is multicast
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            if constexpr (PAIR) {
                // both CTAs' query tiles and corpus halves are counted on rank 0's barriers
                if (crank == 0) mbar_arrive_expect_tx(&a_bar, 2 * KCH * kAChunkBytes);
                const uint32_t a_bar0 = mapa_rank0(smem_u32(&a_bar));
                for (int kc = 0; kc < KCH; ++kc)
                    tma_load_2d_pair(smem_a + kc * kAChunkBytes, &map_q, kc * (INT ? 128 : kChunkK), qtile * kTileQ, a_bar0);
                int stage = 0; uint32_t phase = 0;
                for (int t = 0; t < n_tiles; ++t) {
                    const int row0 = (int)((tile_lo + t) * kTileC) + crank * (kTileC / 2);
                    if (prefetch_tiles > 0 && t + prefetch_tiles < n_tiles)
                        for (int kc = 0; kc < KCH; ++kc)
                            tma_prefetch_2d(&map_c, kc * (INT ? 128 : kChunkK), row0 + prefetch_tiles * kTileC);
                    for (int kc = 0; kc < KCH; ++kc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * BSTAGE);
                        tma_load_2d_pair(smem_b + stage * BSTAGE, &map_c, kc * (INT ? 128 : kChunkK), row0,
                                         mapa_rank0(smem_u32(&full_bar[stage])));
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            } else {
            mbar_arrive_expect_tx(&a_bar, KCH * kAChunkBytes);
            for (int kc = 0; kc < KCH; ++kc)
                tma_load_2d(smem_a + kc * kAChunkBytes, &map_q, kc * (INT ? 128 : kChunkK), qtile * kTileQ, &a_bar);
            int stage = 0; uint32_t phase = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const int row0 = (int)((tile_lo + t) * kTileC);
                if (prefetch_tiles > 0 && t + prefetch_tiles < n_tiles) {   // this CTA's piece of a tile ahead -> L2
                    const int prow = (int)((tile_lo + t + prefetch_tiles) * kTileC) + crank * (kTileC / CS);
                    for (int kc = 0; kc < KCH; ++kc) tma_prefetch_2d(&map_c, kc * (INT ? 128 : kChunkK), prow);
                }
                for (int kc = 0; kc < KCH; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], kBStageBytes);      // all CS pieces land here
                    if constexpr (CS == 1) {
                        tma_load_2d(smem_b + stage * kBStageBytes, &map_c, kc * (INT ? 128 : kChunkK), row0, &full_bar[stage]);
                    } else {
                        constexpr int kPieceRows = kTileC / CS;
                        tma_load_2d_mc(smem_b + stage * kBStageBytes + crank * kPieceRows * 128, &map_c, kc * (INT ? 128 : kChunkK),
                                       row0 + crank * kPieceRows, &full_bar[stage], kMask);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (pair mode: rank 0 only)
        if (lane == 0 && (!PAIR || crank == 0)) {
            mbar_wait(&a_bar, 0);
            tc_fence_after();
#ifdef CRS_GEMM_PROFILE
            long long w_full = 0, w_empty = 0, t_begin = clock64();
#endif
            int stage = 0; uint32_t phase = 0;
            for (int t = 0; t < n_tiles; ++t) {
                const int buf = t & 1;
                const uint32_t tphase = (t >> 1) & 1;
#ifdef CRS_GEMM_PROFILE
                long long c0 = clock64();
#endif
                mbar_wait(&tempty_bar[buf], tphase ^ 1);            // epilogue has drained this accumulator
#ifdef CRS_GEMM_PROFILE
                w_empty += clock64() - c0;
#endif
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * kTileC;
                for (int kc = 0; kc < KCH; ++kc) {
#ifdef CRS_GEMM_PROFILE
                    long long c1 = clock64();
#endif
                    mbar_wait(&full_bar[stage], phase);
#ifdef CRS_GEMM_PROFILE
                    w_full += clock64() - c1;
#endif
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc(smem_u32(smem_a + kc * kAChunkBytes));
                    const uint64_t b_desc = make_smem_desc(smem_u32(smem_b + stage * BSTAGE));
#pragma unroll
                    for (int k = 0; k < kChunkK / 16; ++k) {        // +32 bytes per K=16 step inside the swizzle row
                        if constexpr (PAIR) {
                            if constexpr (INT) umma_i8_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0);
                            else umma_f16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0);
                        } else {
                            if constexpr (INT) umma_i8(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0);
                            else umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) != 0);
                        }
                    }
                    if constexpr (PAIR) umma_commit_pair(&empty_bar[stage]);  // frees the stage in both CTAs
                    else if constexpr (CS == 1) umma_commit(&empty_bar[stage]);   // slot reusable once these MMAs retire
                    else umma_commit_mc(&empty_bar[stage], kMask);          // ... in every CTA that writes into it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if constexpr (PAIR) umma_commit_pair(&tfull_bar[buf]);      // accumulators complete in both CTAs
                else umma_commit(&tfull_bar[buf]);                          // accumulator complete
            }
#ifdef CRS_GEMM_PROFILE
            atomicAdd(&g_gemm_prof[0], (unsigned long long)w_full);
            atomicAdd(&g_gemm_prof[1], (unsigned long long)w_empty);
            atomicAdd(&g_gemm_prof[2], (unsigned long long)(clock64() - t_begin));
            atomicAdd(&g_gemm_prof[3], 1ull);
#endif
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int lane_grp = warp & 3;                              // TMEM lanes this warp may touch
        const int qrow = lane_grp * 32 + lane;
        const int q = qtile * kTileQ + qrow;
        uint64_t best[L];
#pragma unroll
        for (int i = 0; i < L; ++i) best[i] = 0ull;
        // running threshold: L-th best so far.  Padding rows of the last query tile (all-zero
        // queries, every score 0) must never enter the slow path: their threshold is +inf.
        using T = typename ScoreT<INT>::type;
        const T tau_pre = score_of<INT>(tau_pre_bits);
        T tau;
        if constexpr (INT) tau = (q < nq) ? tau_pre : INT32_MAX; else tau = (q < nq) ? tau_pre : INFINITY;
        // optional per-query starting threshold from a sample of the shard (see sample_tau_kernel): every
        // slice starts where a scan of the sample would have ended instead of warming its list up from nothing
        T tau_start = tau_pre;
        if (tau_q != nullptr && q < nq) { tau_start = smax<T>(tau_pre, score_of<INT>(tau_q[q])); tau = tau_start; }
        for (int t = 0; t < n_tiles; ++t) {
            const int buf = t & 1;
            const uint32_t tphase = (t >> 1) & 1;
            mbar_wait(&tfull_bar[buf], tphase);
            tc_fence_after();
            const int64_t row0 = (tile_lo + t) * kTileC;
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + buf * kTileC;
            // two register slabs in flight: the TMEM load of slab s+1 overlaps the scan of slab s
            uint32_t ra[32], rb[32];
            tmem_ld32(taddr, ra);
            tmem_ld_wait();
#pragma unroll 1
            for (int slab = 0; slab < kTileC / 32; slab += 2) {
                tmem_ld32(taddr + (slab + 1) * 32, rb);
                slab_scan<L, INT>(ra, row0 + slab * 32, n_rows, tau_start, tau, best, allow);
                __syncwarp();                                       // tcgen05.ld / wait are .aligned: reconverge first
                tmem_ld_wait();
                if (slab + 2 < kTileC / 32) tmem_ld32(taddr + (slab + 2) * 32, ra);
                slab_scan<L, INT>(rb, row0 + (slab + 1) * 32, n_rows, tau_start, tau, best, allow);
                __syncwarp();
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(mapa_rank0(smem_u32(&tempty_bar[buf])));
                else mbar_arrive(&tempty_bar[buf]);
            }
        }
        if (q < nq) {
            uint64_t* dst = cand + ((size_t)q * n_slices + slice) * list_stride;
#pragma unroll
            for (int i = 0; i < L; ++i) dst[i] = best[i];
            for (int i = L; i < list_stride; ++i) dst[i] = 0ull;
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CS > 1) cluster_sync_all();               // no CTA leaves while peers may still signal it
    if (warp == 1) { if constexpr (PAIR) tmem_dealloc_512_pair(tmem_base); else tmem_dealloc_512(tmem_base); }
}

#ifdef CRS_GEMM_PROFILE
}  // namespace crs
extern "C" __attribute__((visibility("default"))) int crs_debug_gemm_profile(unsigned long long* out8, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out8, crs::g_gemm_prof, sizeof(unsigned long long) * 8);
    if (e == cudaSuccess && reset) {
        unsigned long long z[8] = {0};
        e = cudaMemcpyToSymbol(crs::g_gemm_prof, z, sizeof(z));
    }
    return e == cudaSuccess ? 0 : 2;
}
namespace crs {
#endif

// Starting thresholds from a sample pass: the contraction above is first run over the first m rows of the
// shard; this kernel merges each query's slice lists of that pass (one warp per query) and takes the L-th
// best key.  There are L distinct rows scoring at least that, so with k <= L no row below it can be in the
// top-k: it is a valid floor for EVERY slice of the full pass.  Float stores subtract `margin` (3 x the fast
// pass' error bound) so that certification against this cut always succeeds; integer scores are exact.
__global__ void __launch_bounds__(128)
sample_tau_kernel(const uint64_t* __restrict__ cand, int n_lists, int list_len, int list_stride, int nq, int is_int,
                  const float* __restrict__ qnorms, float margin_rel, uint32_t* __restrict__ tau_q) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * 4 + warp;
    if (q >= nq) return;
    uint64_t e[1] = {0ull};
    const uint64_t* base = cand + (size_t)q * n_lists * list_stride;
    for (int l = 0; l < n_lists; ++l) {
        uint64_t b[1];
        b[0] = (lane < list_len) ? base[(size_t)l * list_stride + lane] : 0ull;
        warp_merge_desc<1>(e, b, lane);
    }
    const uint64_t lth = shfl_u64(e[0], list_len - 1);
    if (lane == 0) {
        uint32_t bits;
        if (is_int) {
            bits = (lth != 0ull) ? (uint32_t)unorderable_i32(key_ord(lth)) : (uint32_t)INT32_MIN;
        } else {
            const float t = (lth != 0ull) ? unorderable_f32(key_ord(lth)) - margin_rel * qnorms[q] : -INFINITY;
            bits = __float_as_uint(t);
        }
        tau_q[q] = bits;
    }
}

cudaError_t launch_sample_tau(cudaStream_t st, const uint64_t* cand, int n_lists, int list_len, int nq, int is_int,
                              const float* qnorms, float margin_rel, uint32_t* tau_q) {
    if (nq <= 0) return cudaSuccess;
    sample_tau_kernel<<<(nq + 3) / 4, 128, 0, st>>>(cand, n_lists, list_len, 32, nq, is_int, qnorms, margin_rel, tau_q);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// kind: 0 = fp16, 1 = bf16, 2 = int8.  One 128-byte swizzle row holds 64 fp16 / 128 int8 elements.
static bool make_map(CUtensorMap* map, const void* base, int64_t rows, int row_bytes, int box_rows, int kind) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const int esz = kind == 2 ? 1 : 2;
    cuuint64_t gdim[2] = {(cuuint64_t)(row_bytes / esz), (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = kind == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                 : (kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    return fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int KCH, int L, int CS, bool INT, bool PAIR>
static cudaError_t launch_kch(cudaStream_t st, const CUtensorMap& mq, const CUtensorMap& mc, int64_t n, int n_qtiles,
                              int n_slices, uint32_t idesc, uint32_t tau_pre, uint64_t* cand, int nq, int list_stride,
                              const uint32_t* allow, const uint32_t* tau_q, int prefetch_tiles) {
    const size_t smem = (size_t)KCH * kAChunkBytes + (size_t)kStages * kBStageBytes + 1024;
    auto kern = gemm_topk_kernel<KCH, L, CS, INT, PAIR>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(n_qtiles * n_slices));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre, cand, nq, list_stride, allow, tau_q, prefetch_tiles);
}

template <int KCH, int L, bool INT>
static cudaError_t launch_cs(int cs, cudaStream_t st, const CUtensorMap& mq, const CUtensorMap& mc, int64_t n,
                             int n_qtiles, int n_slices, uint32_t idesc, uint32_t tau_pre, uint64_t* cand, int nq,
                             const uint32_t* allow, const uint32_t* tau_q, int prefetch_tiles) {
    if (cs == 22) {     // CTA pair: M = 256 across the two SMs of a cluster
        const uint32_t idesc_pair = (idesc & ~(0x1Fu << 24)) | ((uint32_t)(2 * kTileQ >> 4) << 24);
        return launch_kch<KCH, L, 2, INT, true>(st, mq, mc, n, n_qtiles, n_slices, idesc_pair, tau_pre, cand, nq, 32, allow, tau_q, prefetch_tiles);
    }
    if (cs == 4) return launch_kch<KCH, L, 4, INT, false>(st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre, cand, nq, 32, allow, tau_q, prefetch_tiles);
    if (cs == 2) return launch_kch<KCH, L, 2, INT, false>(st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre, cand, nq, 32, allow, tau_q, prefetch_tiles);
    return launch_kch<KCH, L, 1, INT, false>(st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre, cand, nq, 32, allow, tau_q, prefetch_tiles);
}

// kind: 0 fp16, 1 bf16, 2 int8
bool gemm_supported(int row_bytes, int k) {
    const int kch = row_bytes / 128;
    return (kch == 1 || kch == 2 || kch == 3 || kch == 4 || kch == 6) && k <= 128;
}

int gemm_list_len(int k) { return k <= 10 ? 16 : 32; }

cudaError_t launch_gemm_topk(cudaStream_t st, const void* codes, int64_t n, int row_bytes, int kind,
                             const void* qcodes, int nq, int k, uint32_t tau_pre_bits, uint64_t* cand, int num_sms,
                             int cluster, int* n_slices_out, const uint32_t* allow, const uint32_t* tau_q,
                             int prefetch_tiles) {
    const int kch = row_bytes / 128;
    int n_qtiles = (nq + kTileQ - 1) / kTileQ;
    // cluster size: query tiles that share one corpus stream through TMA multicast
    const bool pair = (cluster == 22) && n_qtiles >= 2;      // option value 22: CTA-pair MMA (cta_group::2)
    int cs = pair ? 2 : (cluster > 0 && cluster != 22 ? cluster : (n_qtiles >= 2 ? 2 : 1));
    if (cs != 1 && cs != 2 && cs != 4) cs = 1;
    while (cs > 1 && n_qtiles < cs) cs >>= 1;
    n_qtiles = (n_qtiles + cs - 1) / cs * cs;              // padded query tiles are all-zero (TMA OOB fill)
    int n_slices = num_sms / n_qtiles;
    if (n_slices < 1) n_slices = 1;
    const int64_t tiles_total = (n + kTileC - 1) / kTileC;
    if (n_slices > tiles_total) n_slices = (int)tiles_total;
    *n_slices_out = n_slices;
    CUtensorMap mq, mc;
    if (!make_map(&mq, qcodes, nq, row_bytes, kTileQ, kind) || !make_map(&mc, codes, n, row_bytes, kTileC / cs, kind))
        return cudaErrorInvalidValue;
    // instruction descriptor: both operands K-major, N=256, M=128;
    // f16/bf16: D=f32 (c_format 1), a/b format 0|1;  int8: D=s32 (c_format 2), a/b format 1 (signed)
    const uint32_t cfmt = kind == 2 ? 2u : 1u;
    const uint32_t fmt = kind == 0 ? 0u : 1u;
    const uint32_t idesc = (cfmt << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kTileC >> 3) << 17) | ((uint32_t)(kTileQ >> 4) << 24);
    const int L = gemm_list_len(k);
    if (pair) cs = 22;
#define CRS_GEMM_CASE(KCH_)                                                                                              \
    case KCH_:                                                                                                           \
        if (kind == 2)                                                                                                   \
            return L == 16 ? launch_cs<KCH_, 16, true>(cs, st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre_bits, cand, nq, allow, tau_q, prefetch_tiles)   \
                           : launch_cs<KCH_, 32, true>(cs, st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre_bits, cand, nq, allow, tau_q, prefetch_tiles);  \
        return L == 16 ? launch_cs<KCH_, 16, false>(cs, st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre_bits, cand, nq, allow, tau_q, prefetch_tiles)      \
                       : launch_cs<KCH_, 32, false>(cs, st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre_bits, cand, nq, allow, tau_q, prefetch_tiles);
    switch (kch) {
        CRS_GEMM_CASE(1)
        CRS_GEMM_CASE(2)
        CRS_GEMM_CASE(3)
        CRS_GEMM_CASE(4)
        CRS_GEMM_CASE(6)
        default: return cudaErrorInvalidValue;
    }
#undef CRS_GEMM_CASE
}

}  // namespace crs
