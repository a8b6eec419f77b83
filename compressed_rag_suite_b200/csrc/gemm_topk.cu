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
// Roles inside a CTA (9 warps = 288 threads, 1 CTA / SM):
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
//           hit walks the slab and inserts into a sorted register list of L keys.  The first
//           `warm_tiles` tiles of a slice are only looked at (group maxima seed the floor) and are
//           computed a second time at the end of the slice (deferred warm-up, see n_warm).
//   warps 6-7  floor sharing: the slices of a query tile run on different SMs and each keeps its own
//           list, so without help every (query, slice) list warms up on its own (16 ln(n/16) inserts
//           each, ~13x what one list over the whole shard would need).  Epilogue threads publish
//           the scores of their list after a tile that changed it (pub[q][slice][0..L), 32-bit
//           orderable scores, each slot monotone over time); this warp keeps merging the published
//           lists of the queries assigned to its CTA and stores the k-th best score of the union as
//           the query's floor (tau_q), which every slice re-reads once per tile.  A floor is valid
//           whenever k published scores reach it — k distinct rows score at least that — whatever
//           mix of old and new slot values a racing reader sees: the published slots are scores of
//           distinct rows and only ever grow (atomicMax), so for any threshold t the number of slots
//           a reader sees at or above t never exceeds the number of rows that really score >= t.
//   warp 8  slice lock-step monitor: keeps the clusters that stream the same corpus slice within a few
//           tiles of each other so that they share the slice through L2 (see lock_allowed).
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
// progress curve: [j] = sum over issuers of the cycles since CTA start at which tile 2^(j-1) (j = 0: tile 0) was begun;
// [15] = at the end of the last tile
__device__ unsigned long long g_gemm_tl[16];
// floor-sharing helper: [0] rounds, [1] cycles spent in rounds, [2] helpers, [3] cycles (since CTA start) at which the
// first floor was stored, [4] helpers that stored one; epilogue (warp 2, lane 0): [5] candidates inserted in the real
// list, [6] slabs that took the insert path, [7] inserts during the warm-up tiles
__device__ unsigned long long g_gemm_hp[8];
#endif

constexpr int kGemmThreads = 288;      // TMA warp, MMA warp, 4 epilogue warps, 2 floor-sharing warps, lock-step monitor warp
constexpr int kTileQ = 128;        // UMMA M
constexpr int kTileC = 256;        // UMMA N (corpus rows per tile)
constexpr int kChunkK = 64;        // fp16 elements per 128-byte swizzle row
constexpr int kStages = 4;
constexpr int kProgStride = 128;   // progress words per corpus slice (lock-step): every slice on its own L2 lines, up to 128 query groups
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

// One thread's candidate list: L entries sorted by (score descending, row ascending), kept as two 32-bit
// register arrays (orderable score, 0xFFFFFFFF - row); ord == 0 marks an empty slot.
//
// Insert of an entry that beats the last one: its position is the number of entries that sort before it.
// All L position tests are independent and every slot then takes one of {itself, its left neighbour, x}:
// two dependent steps instead of an L-step compare-exchange chain (the epilogue has one warp per scheduler
// and nothing to hide a dependent chain behind).
template <int L>
__device__ __forceinline__ void list_insert(uint32_t (&ord)[L], uint32_t (&nid)[L], uint32_t x_ord, uint32_t x_nid) {
    bool keep[L];
#pragma unroll
    for (int i = 0; i < L; ++i) keep[i] = ord[i] > x_ord || (ord[i] == x_ord && nid[i] > x_nid);
#pragma unroll
    for (int i = L - 1; i >= 1; --i) {
        ord[i] = keep[i] ? ord[i] : (keep[i - 1] ? x_ord : ord[i - 1]);
        nid[i] = keep[i] ? nid[i] : (keep[i - 1] ? x_nid : nid[i - 1]);
    }
    ord[0] = keep[0] ? ord[0] : x_ord;
    nid[0] = keep[0] ? nid[0] : x_nid;
}
// the same for a list of scores only (warm-up: group maxima)
template <int L>
__device__ __forceinline__ void scores_insert(uint32_t (&ord)[L], uint32_t x_ord) {
    bool keep[L];
#pragma unroll
    for (int i = 0; i < L; ++i) keep[i] = ord[i] >= x_ord;
#pragma unroll
    for (int i = L - 1; i >= 1; --i) ord[i] = keep[i] ? ord[i] : (keep[i - 1] ? x_ord : ord[i - 1]);
    ord[0] = keep[0] ? ord[0] : x_ord;
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

template <bool INT>
__device__ __forceinline__ typename ScoreT<INT>::type slab_max(const uint32_t (&r)[32]) {
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
    return smax<T>(m[0], m[9]);
}

// Warm-up tiles (see the kernel): the maxima of the slab's G-element groups go into a sorted list of L scores.
// Each entry is the score of a different row, so the list's last entry — once the list is full — is a
// floor that at least L rows of this slice reach.
template <int L, bool INT, int G>
__device__ __forceinline__ void slab_warm(const uint32_t (&r)[32], uint32_t (&top)[L], bool live, bool& dirty) {
    using T = typename ScoreT<INT>::type;
#pragma unroll
    for (int g = 0; g < 32 / G; ++g) {
        T m = score_of<INT>(r[g * G]);
#pragma unroll
        for (int i = 1; i < G; ++i) m = smax<T>(m, score_of<INT>(r[g * G + i]));
        const uint32_t o = ord_of<INT>(m);
        if (live && o > top[L - 1]) {
            scores_insert<L>(top, o);
            dirty = true;
        }
    }
}

template <int L, bool INT>
__device__ __forceinline__ void slab_scan(const uint32_t (&r)[32], int64_t row0, int64_t n_rows,
                                          typename ScoreT<INT>::type& tau, uint32_t (&best_ord)[L], uint32_t (&best_nid)[L],
                                          const uint32_t* __restrict__ allow, bool& dirty
#ifdef CRS_GEMM_PROFILE
                                          , unsigned& n_slow, unsigned& n_ins
#endif
                                          ) {
    using T = typename ScoreT<INT>::type;
    if (slab_max<INT>(r) >= tau) {
#ifdef CRS_GEMM_PROFILE
        ++n_slow;
#endif
        unsigned mask = 0;
        T tmp[32];                           // dynamically indexed -> local memory, touched on this path only
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const T v = score_of<INT>(r[i]);
            tmp[i] = v;
            mask |= (v >= tau) ? (1u << i) : 0u;
        }
        // walk this thread's own hits; the next hit's score is fetched (local memory) while the current one is inserted
        int i = __ffs(mask) - 1;
        mask &= mask - 1;
        T v = tmp[i & 31];
        while (i >= 0) {
            int i_next = -1;
            T v_next = v;
            if (mask) {
                i_next = __ffs(mask) - 1;
                mask &= mask - 1;
                v_next = tmp[i_next];
            }
            const int64_t row = row0 + i;
            // row bitmap of a where / where_document filter: looked up for hits only
            if (v >= tau && row < n_rows && (allow == nullptr || ((allow[row >> 5] >> (row & 31)) & 1u))) {
                const uint32_t o = ord_of<INT>(v);
                const uint32_t nr = 0xFFFFFFFFu - (uint32_t)row;
                if (o > best_ord[L - 1] || (o == best_ord[L - 1] && nr > best_nid[L - 1])) {
                    list_insert<L>(best_ord, best_nid, o, nr);
#ifdef CRS_GEMM_PROFILE
                    ++n_ins;
#endif
                    dirty = true;
                    if (best_ord[L - 1] != 0u) tau = smax<T>(tau, from_ord<INT>(best_ord[L - 1]));
                }
            }
            i = i_next;
            v = v_next;
        }
    }
}

// Per-query floors (see the header comment).  tau_q holds ORDERABLE 32-bit scores (0 = none yet).
struct GemmFloor {
    uint32_t* tau_q;        // [nq] in/out; NULL = no floors at all
    uint32_t* pub;          // [nq][n_slices][L] published list scores; NULL = tau_q is only read once (sample pass result)
    const float* qnorms;    // [nq] (float stores: the floor is lowered by margin_rel * |q|)
    float margin_rel;
    int k;                  // rank of the union that makes the floor
    uint32_t* progress;     // [n_slices][query groups] tile each cluster has reached (slice lock-step); NULL = off
    int lead_tiles;         // how far a cluster may run ahead of the slowest cluster of its slice
};

__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// One query's floor: k-th best of the union of its slices' published lists (a warp; LPL = 1 for k <= 32).
// A racing reader may see a list half-updated (unsorted): the merge network then still passes on distinct
// slots of the union — every step is a compare-exchange or a pick-one-of-two — just not necessarily the
// largest ones, and the final sort makes position k-1 the k-th largest of those: a valid, if weaker, floor.
// Returns 0 (no floor) when told to stop early.
template <int LPL>
__device__ __noinline__ uint32_t union_kth_ord(const uint32_t* __restrict__ pubq, int n_slices, int list_len, int k, int lane,
                                                  volatile int* done) {
    uint64_t e[LPL];
#pragma unroll
    for (int s = 0; s < LPL; ++s) e[s] = 0ull;
    for (int sl = 0; sl < n_slices; ++sl) {
        if (*done >= 4) return 0u;                          // the epilogue has finished: nobody needs the floor any more
        uint64_t b[LPL];
#pragma unroll
        for (int s = 0; s < LPL; ++s) {
            const int i = lane * LPL + s;
            const uint32_t o = (i < list_len) ? ld_cg_u32(pubq + (size_t)sl * list_len + i) : 0u;
            b[s] = (uint64_t)o << 32;
        }
        if (__shfl_sync(CRS_FULL_MASK, b[0], 0) == 0ull) continue;     // nothing published by this slice yet
        warp_merge_desc<LPL>(e, b, lane);
    }
    warp_sort_desc<LPL>(e, lane);
    const int kk = k - 1;
    uint64_t kth = 0ull;
#pragma unroll
    for (int s = 0; s < LPL; ++s) {
        const uint64_t v = shfl_u64(e[s], kk / LPL);
        if (s == kk % LPL) kth = v;
    }
    return (uint32_t)(kth >> 32);
}

// The common case (k <= 32): the query's published scores are one contiguous array of n_slices * L words,
// fetched in 32-word chunks (two lists of 16 or one of 32).  A refresh is latency-bound — one warp, and every
// step of a bitonic network waits for a shuffle — so the work is arranged for instruction-level parallelism:
// 8 chunks are fetched with independent loads, sorted side by side and tree-merged (3 levels) before they
// meet the running top-32.  For L = 16 the second list of a chunk is read back to front, which makes the chunk
// a bitonic sequence that one 5-step merge sorts.  Values below the previous k-th best cannot move the k-th
// best any more and are dropped on load.  A list caught between two slot updates may be unsorted: the
// networks then still pass on distinct slots, and the final sort makes position k-1 a valid floor.
template <int L>
__device__ __noinline__ uint32_t union_kth_small(const uint32_t* __restrict__ pubq, int n_words, int k, int lane,
                                                 volatile int* done, uint32_t prev_kth) {
    constexpr int G = 8;
    const int n_chunks = (n_words + 31) / 32;
    const int off = (L == 16 && lane >= 16) ? 47 - lane : lane;        // second list of a chunk: reversed
    uint64_t e[1] = {0ull};
    for (int c0 = 0; c0 < n_chunks; c0 += G) {
        if (*done >= 4) return 0u;
        uint64_t b[G][1];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int w = (c0 + j) * 32 + off;
            const uint32_t v = (w < n_words) ? ld_cg_u32(pubq + w) : 0u;
            b[j][0] = (uint64_t)(v >= prev_kth ? v : 0u) << 32;
        }
        if constexpr (L == 16) {
#pragma unroll
            for (int j = 0; j < G; ++j) BitonicInner<1, 64, 16>::run(b[j], lane);
        }
#pragma unroll
        for (int stride = 1; stride < G; stride *= 2) {
#pragma unroll
            for (int i = 0; i + stride < G; i += 2 * stride) warp_merge_desc<1>(b[i], b[i + stride], lane);
        }
        warp_merge_desc<1>(e, b[0], lane);
    }
    warp_sort_desc<1>(e, lane);
    return (uint32_t)(shfl_u64(e[0], k - 1) >> 32);
}

template <int KCH, int L, int CS, bool INT, bool PAIR>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_c,
                 int64_t n_rows, int n_qtiles, int n_slices, uint32_t idesc, uint32_t tau_pre_bits,
                 uint64_t* __restrict__ cand, int nq, int list_stride, const uint32_t* __restrict__ allow,
                 GemmFloor fl, int prefetch_tiles, int warm_tiles) {
    static_assert(!PAIR || CS == 2, "the CTA-pair MMA needs clusters of exactly two CTAs");
    // pair mode: a stage holds this CTA's half (128 rows) of a corpus chunk -> twice the stages in the same smem
    constexpr int STAGES = PAIR ? 2 * kStages : kStages;
    constexpr int BSTAGE = PAIR ? kBStageBytes / 2 : kBStageBytes;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[2 * kStages], empty_bar[2 * kStages], a_bar, tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_slot;
    __shared__ int epi_done;                                  // epilogue warps that have finished their slice
    // Slice lock-step.  The clusters that stream the same corpus slice (one per pair of query tiles) share it through
    // L2 only while they stay within a few MB of each other; left alone they drift apart (their epilogues stall at
    // different times) and every one of them ends up reading the slice from DRAM (ncu: 18.9 GB read for a 7.68 GB
    // corpus; 7.8 GB with the lock-step).  Each cluster publishes the tile it has reached and a leader waits while it
    // is more than `lead_tiles` ahead of the slowest cluster of its slice.  A slice is as slow as its slowest cluster,
    // so holding the leaders back costs nothing, and the DRAM power it saves is clock headroom under the power cap.
    __shared__ uint32_t lock_allowed;
    __shared__ int lock_done;

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                   // KCH x 16 KB
    uint8_t* smem_b = smem + KCH * kAChunkBytes;              // STAGES x BSTAGE (128 KB)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef CRS_GEMM_PROFILE
    const long long t_cta = clock64();
#endif
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
    // Deferred warm-up.  A slice's list starts empty, and until its floor is tight nearly every score is a
    // candidate: inserting them kept the epilogue far behind the tensor core for the first ~30 tiles of every
    // slice (measured: ~200 K cycles per launch).  Instead the first n_warm tiles are only LOOKED at — the
    // maxima of small groups of scores go into a score-only list, published to the floor sharing like a real
    // list — and are computed a second time at the END of the slice, when the floor is tight and almost
    // nothing in them is a candidate.  Costs n_warm extra tiles of MMA per slice.
    const int n_warm = (allow == nullptr && warm_tiles > 0 && n_tiles >= 4 * warm_tiles) ? warm_tiles : 0;
    const int n_iter = n_tiles + n_warm;                      // iteration it works on tile (it < n_tiles ? it : it - n_tiles)

    if (threadIdx.x == 0) {
        epi_done = 0;
        lock_allowed = (uint32_t)(fl.lead_tiles > 0 ? fl.lead_tiles : 1);
        lock_done = 0;
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
                for (int it = 0; it < n_iter; ++it) {
                    const int t = it < n_tiles ? it : it - n_tiles;
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
            // slice lock-step (see lock_allowed above)
            uint32_t* prog = (fl.progress != nullptr && crank == 0 && qgroups > 1) ? fl.progress + (size_t)slice * kProgStride : nullptr;
            const int my_qg = cluster_id % qgroups;
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < n_iter; ++it) {
                const int t = it < n_tiles ? it : it - n_tiles;
                const int row0 = (int)((tile_lo + t) * kTileC);
                if (prog != nullptr) {
                    // publish where this cluster is; wait (bounded) while it is too far ahead of the slowest cluster of
                    // its slice.  `lock_allowed` is kept up to date by lane 1 of this warp, so the check is one shared-
                    // memory read and never puts an L2 round trip in front of the TMA loads.
                    *reinterpret_cast<volatile uint32_t*>(prog + my_qg) = (uint32_t)it;
                    for (int spin = 0; spin < 200000 && (uint32_t)it > *reinterpret_cast<volatile uint32_t*>(&lock_allowed); ++spin)
                        __nanosleep(40);
                }
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
            if (prog != nullptr) *reinterpret_cast<volatile uint32_t*>(prog + my_qg) = 0xFFFFFFFFu;   // done: never the slowest
            *reinterpret_cast<volatile int*>(&lock_done) = 1;
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
            for (int t = 0; t < n_iter; ++t) {
                const int buf = t & 1;
                const uint32_t tphase = (t >> 1) & 1;
#ifdef CRS_GEMM_PROFILE
                long long c0 = clock64();
                if ((t & (t - 1)) == 0) {                   // t = 0, 1, 2, 4, 8, ...
                    const int j = t == 0 ? 0 : 32 - __clz(t);
                    if (j < 15) atomicAdd(&g_gemm_tl[j], (unsigned long long)(c0 - t_cta));
                }
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
            atomicAdd(&g_gemm_prof[7], (unsigned long long)n_iter);
            atomicAdd(&g_gemm_tl[15], (unsigned long long)(clock64() - t_cta));
#endif
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------ slice lock-step monitor (its own warp: a spinning,
        // sleeping lane inside the TMA or MMA warp would stall that warp's one working lane)
        if (lane == 0 && !PAIR && fl.progress != nullptr && crank == 0 && qgroups > 1) {
            // tile index up to which this cluster may load = slowest cluster of the slice + lead
            const uint32_t* prog = fl.progress + (size_t)slice * kProgStride;
            while (*reinterpret_cast<volatile int*>(&lock_done) == 0) {
                uint32_t slowest = 0xFFFFFFFFu;
                for (int j = 0; j < qgroups; ++j) slowest = min(slowest, ld_cg_u32(prog + j));
                const uint32_t allowed = slowest > 0xFFFFFFFFu - (uint32_t)fl.lead_tiles ? 0xFFFFFFFFu : slowest + (uint32_t)fl.lead_tiles;
                *reinterpret_cast<volatile uint32_t*>(&lock_allowed) = allowed;
                __nanosleep(1500);                                  // ~half a tile: polling faster only fights over the progress lines in L2
            }
        }
    } else if (warp >= 6) {
        // ------------------------------------------------------------ floor sharing (see the header comment)
        if (fl.pub != nullptr) {
            volatile int* done = &epi_done;
            __shared__ uint32_t prev[8];                            // k-th best of the last refresh, per assigned query
            if (lane < 8 && (lane & 1) == warp - 6) prev[lane] = 0u;     // this warp's queries
            __syncwarp();
#ifdef CRS_GEMM_PROFILE
            long long h_rounds = 0, h_cyc = 0, h_first = 0;
#endif
            while (*done < 4) {
#ifdef CRS_GEMM_PROFILE
                const long long h0 = clock64();
#endif
#pragma unroll 1
                for (int j = warp - 6; j < 8; j += 2) {             // each query of the tile belongs to one slice's CTA (and one of its two helper warps)
                    const int i = slice + j * n_slices;
                    const int q = qtile * kTileQ + i;
                    if (i >= kTileQ || q >= nq || *done >= 4) break;
                    const uint32_t* pubq = fl.pub + (size_t)q * n_slices * L;
                    const uint32_t o = (fl.k <= 32) ? union_kth_small<L>(pubq, n_slices * L, fl.k, lane, done, prev[j])
                                                    : union_kth_ord<4>(pubq, n_slices, L, fl.k, lane, done);
                    if (o != 0u) {
                        if (lane == 0) {
                            prev[j] = o;
                            uint32_t f = o;
                            if constexpr (!INT) f = orderable_f32(unorderable_f32(o) - fl.margin_rel * fl.qnorms[q]);
                            atomicMax(fl.tau_q + q, f);
#ifdef CRS_GEMM_PROFILE
                            if (h_first == 0) h_first = clock64() - t_cta;
#endif
                        }
                        __syncwarp();
                    }
                }
#ifdef CRS_GEMM_PROFILE
                h_cyc += clock64() - h0; ++h_rounds;
#endif
                __nanosleep(200);
            }
#ifdef CRS_GEMM_PROFILE
            if (lane == 0) {
                atomicAdd(&g_gemm_hp[0], (unsigned long long)h_rounds);
                atomicAdd(&g_gemm_hp[1], (unsigned long long)h_cyc);
                atomicAdd(&g_gemm_hp[2], 1ull);
                if (h_first) { atomicAdd(&g_gemm_hp[3], (unsigned long long)h_first); atomicAdd(&g_gemm_hp[4], 1ull); }
            }
#endif
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..5)
        const int lane_grp = warp & 3;                              // TMEM lanes this warp may touch
        const int qrow = lane_grp * 32 + lane;
        const int q = qtile * kTileQ + qrow;
        uint32_t best_ord[L], best_nid[L];
#pragma unroll
        for (int i = 0; i < L; ++i) { best_ord[i] = 0u; best_nid[i] = 0u; }
        // running threshold: L-th best so far.  Padding rows of the last query tile (all-zero
        // queries, every score 0) must never enter the slow path: their threshold is +inf.
        using T = typename ScoreT<INT>::type;
        const T tau_pre = score_of<INT>(tau_pre_bits);
        T tau;
        if constexpr (INT) tau = (q < nq) ? tau_pre : INT32_MAX; else tau = (q < nq) ? tau_pre : INFINITY;
        // per-query floor: from a sample pass (read once) and / or shared between the slices while they run
        // (re-read once per tile; the load is issued here and consumed after the tile's slabs)
        const bool floors = fl.tau_q != nullptr && q < nq;
        const bool sharing = floors && fl.pub != nullptr;
        if (floors) {
            const uint32_t o = ld_cg_u32(fl.tau_q + q);
            if (o != 0u) tau = smax<T>(tau, from_ord<INT>(o));
        }
        bool dirty = false;
#ifdef CRS_GEMM_PROFILE
        long long e_wait = 0, e_begin = clock64();
        unsigned p_slow = 0, p_ins = 0, p_slow_early = 0, p_ins_early = 0;
#endif
        for (int it = 0; it < n_iter; ++it) {
            const int buf = it & 1;
            const uint32_t tphase = (it >> 1) & 1;
            const int t = it < n_tiles ? it : it - n_tiles;
#ifdef CRS_GEMM_PROFILE
            const long long ew0 = clock64();
#endif
            mbar_wait(&tfull_bar[buf], tphase);
#ifdef CRS_GEMM_PROFILE
            e_wait += clock64() - ew0;
#endif
            tc_fence_after();
            uint32_t floor_ord = 0u;
            if (sharing) floor_ord = ld_cg_u32(fl.tau_q + q);
            const int64_t row0 = (tile_lo + t) * kTileC;
            const uint32_t taddr = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + buf * kTileC;
            // two register slabs in flight: the TMEM load of slab s+1 overlaps the scan of slab s
            uint32_t ra[32], rb[32];
            tmem_ld32(taddr, ra);
            tmem_ld_wait();
            if (it < n_warm) {
                // warm-up tile: group maxima only (best_ord serves as the score-only list, best_nid is unused)
#pragma unroll 1
                for (int slab = 0; slab < kTileC / 32; slab += 2) {
                    tmem_ld32(taddr + (slab + 1) * 32, rb);
                    slab_warm<L, INT, kTileC / L>(ra, best_ord, q < nq, dirty);
                    __syncwarp();
                    tmem_ld_wait();
                    if (slab + 2 < kTileC / 32) tmem_ld32(taddr + (slab + 2) * 32, ra);
                    slab_warm<L, INT, kTileC / L>(rb, best_ord, q < nq, dirty);
                    __syncwarp();
                    tmem_ld_wait();
                }
            } else {
#pragma unroll 1
                for (int slab = 0; slab < kTileC / 32; slab += 2) {
                    tmem_ld32(taddr + (slab + 1) * 32, rb);
                    slab_scan<L, INT>(ra, row0 + slab * 32, n_rows, tau, best_ord, best_nid, allow, dirty
#ifdef CRS_GEMM_PROFILE
                                      , p_slow, p_ins
#endif
                                      );
                    __syncwarp();                                   // tcgen05.ld / wait are .aligned: reconverge first
                    tmem_ld_wait();
                    if (slab + 2 < kTileC / 32) tmem_ld32(taddr + (slab + 2) * 32, ra);
                    slab_scan<L, INT>(rb, row0 + (slab + 1) * 32, n_rows, tau, best_ord, best_nid, allow, dirty
#ifdef CRS_GEMM_PROFILE
                                      , p_slow, p_ins
#endif
                                      );
                    __syncwarp();
                    tmem_ld_wait();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(mapa_rank0(smem_u32(&tempty_bar[buf])));
                else mbar_arrive(&tempty_bar[buf]);
            }
            if (sharing) {
                if (dirty) {
                    // publish this list's scores: every slot only ever grows (atomic max), whichever of the two
                    // lists — warm-up scores, then the real list — it is fed from
                    uint32_t* dstp = fl.pub + ((size_t)q * n_slices + slice) * L;
#pragma unroll
                    for (int i = 0; i < L; ++i)
                        if (best_ord[i] != 0u) atomicMax(dstp + i, best_ord[i]);
                    dirty = false;
                }
                if (floor_ord != 0u) tau = smax<T>(tau, from_ord<INT>(floor_ord));
            }
#ifdef CRS_GEMM_PROFILE
            if (it + 1 == n_warm + 8) { p_slow_early = p_slow; p_ins_early = p_ins; }
#endif
            if (it + 1 == n_warm) {
                // end of the warm-up: its L-th best group maximum is this slice's first floor; the real list starts empty
                if (q < nq && best_ord[L - 1] != 0u) tau = smax<T>(tau, from_ord<INT>(best_ord[L - 1]));
#pragma unroll
                for (int i = 0; i < L; ++i) { best_ord[i] = 0u; best_nid[i] = 0u; }
            }
        }
        __syncwarp();
        if (lane == 0) atomicAdd(&epi_done, 1);
#ifdef CRS_GEMM_PROFILE
        if (lane == 0 && warp == 2) {            // one epilogue warp per CTA: cycles waiting for an accumulator / in total
            atomicAdd(&g_gemm_prof[4], (unsigned long long)e_wait);
            atomicAdd(&g_gemm_prof[5], (unsigned long long)(clock64() - e_begin));
            atomicAdd(&g_gemm_prof[6], 1ull);
        }
        {   // insert-path statistics of every epilogue thread: [5] inserts, [6] slabs on the insert path, [7] the same two
            // packed for the first 8 tiles after the warm-up (inserts << 32 | slabs)
            atomicAdd(&g_gemm_hp[5], (unsigned long long)p_ins);
            atomicAdd(&g_gemm_hp[6], (unsigned long long)p_slow);
            atomicAdd(&g_gemm_hp[7], ((unsigned long long)p_ins_early << 32) | (unsigned long long)p_slow_early);
        }
#endif
        if (q < nq) {
            uint64_t* dst = cand + ((size_t)q * n_slices + slice) * list_stride;
#pragma unroll
            for (int i = 0; i < L; ++i) dst[i] = ((uint64_t)best_ord[i] << 32) | (uint64_t)best_nid[i];
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
extern "C" __attribute__((visibility("default"))) int crs_debug_gemm_profile(unsigned long long* out24 /*[32]*/, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out24, crs::g_gemm_prof, sizeof(unsigned long long) * 8);
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out24 + 8, crs::g_gemm_tl, sizeof(unsigned long long) * 16);
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out24 + 24, crs::g_gemm_hp, sizeof(unsigned long long) * 8);
    if (e == cudaSuccess && reset) {
        unsigned long long z[16] = {0};
        e = cudaMemcpyToSymbol(crs::g_gemm_prof, z, sizeof(unsigned long long) * 8);
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(crs::g_gemm_tl, z, sizeof(z));
        if (e == cudaSuccess) e = cudaMemcpyToSymbol(crs::g_gemm_hp, z, sizeof(unsigned long long) * 8);
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
    if (lane == 0) {                                   // orderable score, 0 = no floor
        uint32_t o = 0u;
        if (lth != 0ull) o = is_int ? key_ord(lth) : orderable_f32(unorderable_f32(key_ord(lth)) - margin_rel * qnorms[q]);
        tau_q[q] = o;
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
                              const uint32_t* allow, const GemmFloor& fl, int prefetch_tiles, int warm_tiles) {
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
    return cudaLaunchKernelEx(&cfg, kern, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre, cand, nq, list_stride, allow, fl, prefetch_tiles, warm_tiles);
}

template <int KCH, int L, bool INT>
static cudaError_t launch_cs(int cs, cudaStream_t st, const CUtensorMap& mq, const CUtensorMap& mc, int64_t n,
                             int n_qtiles, int n_slices, uint32_t idesc, uint32_t tau_pre, uint64_t* cand, int nq,
                             const uint32_t* allow, const GemmFloor& fl, int prefetch_tiles, int warm_tiles) {
    if (cs == 22) {     // CTA pair: M = 256 across the two SMs of a cluster
        const uint32_t idesc_pair = (idesc & ~(0x1Fu << 24)) | ((uint32_t)(2 * kTileQ >> 4) << 24);
        return launch_kch<KCH, L, 2, INT, true>(st, mq, mc, n, n_qtiles, n_slices, idesc_pair, tau_pre, cand, nq, 32, allow, fl, prefetch_tiles, warm_tiles);
    }
    if (cs == 4) return launch_kch<KCH, L, 4, INT, false>(st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre, cand, nq, 32, allow, fl, prefetch_tiles, warm_tiles);
    if (cs == 2) return launch_kch<KCH, L, 2, INT, false>(st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre, cand, nq, 32, allow, fl, prefetch_tiles, warm_tiles);
    return launch_kch<KCH, L, 1, INT, false>(st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre, cand, nq, 32, allow, fl, prefetch_tiles, warm_tiles);
}

// kind: 0 fp16, 1 bf16, 2 int8
bool gemm_supported(int row_bytes, int k) {
    const int kch = row_bytes / 128;
    return (kch == 1 || kch == 2 || kch == 3 || kch == 4 || kch == 6) && k <= 128;
}

int gemm_list_len(int k) { return k <= 10 ? 16 : 32; }

// Clusters of `cs` CTAs (1 CTA / SM: every instantiation asks for ~225 KB of shared memory) that can be resident
// at once on this GPU — clusters must sit inside one GPC, so this can be less than num_sms / cs.  A grid with
// more clusters than that would run in two waves.  Asked once per cluster size (for the widest instantiation).
static int max_resident_clusters(int cs) {
    static int cache[5] = {0, 0, 0, 0, 0};
    if (cs < 1 || cs > 4) return 0;
    if (cache[cs] != 0) return cache[cs];
    const void* kern = cs == 4 ? (const void*)gemm_topk_kernel<6, 16, 4, false, false>
                     : cs == 2 ? (const void*)gemm_topk_kernel<6, 16, 2, false, false>
                               : (const void*)gemm_topk_kernel<6, 16, 1, false, false>;
    const size_t smem = (size_t)6 * kAChunkBytes + (size_t)kStages * kBStageBytes + 1024;
    int n = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(cs * 64);
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
    cache[cs] = n > 0 ? n : -1;
    return cache[cs];
}

// work split of one launch: query tiles (padded to the cluster size), corpus slices, cluster size
static void gemm_plan(int64_t n, int nq, int num_sms, int cluster, int* n_qtiles_out, int* n_slices_out, int* cs_out,
                      bool* pair_out) {
    int n_qtiles = (nq + kTileQ - 1) / kTileQ;
    // cluster size: query tiles that share one corpus stream through TMA multicast
    const bool pair = (cluster == 22) && n_qtiles >= 2;      // option value 22: CTA-pair MMA (cta_group::2)
    int cs = pair ? 2 : (cluster > 0 && cluster != 22 ? cluster : (n_qtiles >= 2 ? 2 : 1));
    if (cs != 1 && cs != 2 && cs != 4) cs = 1;
    while (cs > 1 && n_qtiles < cs) cs >>= 1;
    n_qtiles = (n_qtiles + cs - 1) / cs * cs;              // padded query tiles are all-zero (TMA OOB fill)
    int n_slices = num_sms / n_qtiles;
    const int fit = max_resident_clusters(cs);             // one wave: no more clusters than can be resident together
    if (fit > 0 && n_slices * (n_qtiles / cs) > fit) n_slices = fit / (n_qtiles / cs);
    if (n_slices < 1) n_slices = 1;
    const int64_t tiles_total = (n + kTileC - 1) / kTileC;
    if (n_slices > tiles_total) n_slices = (int)tiles_total;
    *n_qtiles_out = n_qtiles; *n_slices_out = n_slices; *cs_out = cs; *pair_out = pair;
}

int gemm_n_slices(int64_t n, int nq, int num_sms, int cluster) {
    int n_qtiles, n_slices, cs; bool pair;
    gemm_plan(n, nq, num_sms, cluster, &n_qtiles, &n_slices, &cs, &pair);
    return n_slices;
}
int gemm_progress_words(int64_t n, int nq, int num_sms, int cluster) {
    int n_qtiles, n_slices, cs; bool pair;
    gemm_plan(n, nq, num_sms, cluster, &n_qtiles, &n_slices, &cs, &pair);
    return (n_qtiles / cs) <= kProgStride ? n_slices * kProgStride : 0;       // more query groups than a slice has words: no lock-step
}

cudaError_t launch_gemm_topk(cudaStream_t st, const void* codes, int64_t n, int row_bytes, int kind,
                             const void* qcodes, int nq, int k, uint32_t tau_pre_bits, uint64_t* cand, int num_sms,
                             int cluster, int* n_slices_out, const uint32_t* allow, const GemmFloorArgs* floors,
                             int prefetch_tiles, int warm_tiles) {
    const int kch = row_bytes / 128;
    int n_qtiles, n_slices, cs; bool pair;
    gemm_plan(n, nq, num_sms, cluster, &n_qtiles, &n_slices, &cs, &pair);
    *n_slices_out = n_slices;
    GemmFloor fl{};
    if (floors != nullptr) {
        fl.tau_q = floors->tau_q; fl.pub = floors->pub; fl.qnorms = floors->qnorms; fl.margin_rel = floors->margin_rel;
        fl.progress = floors->progress;
        fl.lead_tiles = floors->lead_tiles > 0 ? floors->lead_tiles : 6;
    }
    fl.k = k;
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
            return L == 16 ? launch_cs<KCH_, 16, true>(cs, st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre_bits, cand, nq, allow, fl, prefetch_tiles, warm_tiles)   \
                           : launch_cs<KCH_, 32, true>(cs, st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre_bits, cand, nq, allow, fl, prefetch_tiles, warm_tiles);  \
        return L == 16 ? launch_cs<KCH_, 16, false>(cs, st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre_bits, cand, nq, allow, fl, prefetch_tiles, warm_tiles)      \
                       : launch_cs<KCH_, 32, false>(cs, st, mq, mc, n, n_qtiles, n_slices, idesc, tau_pre_bits, cand, nq, allow, fl, prefetch_tiles, warm_tiles);
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
