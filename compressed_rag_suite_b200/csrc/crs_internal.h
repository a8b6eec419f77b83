// Internal declarations shared by the translation units of libcrs.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/crs.h"

namespace crs {

// Widest candidate list a search keeps per query: M = 32 * LPL keys, LPL in {1, 4}.
constexpr int kMaxListLen = 128;

// Optional: a single-query scan encodes the fp32 query ITSELF in its prologue (every CTA redundantly, in shared
// memory, while its producer warp already streams the first tiles), instead of a separate encode launch in
// front of it — the launch and its gap were ~10 us of a 150 us single-query search.  Same arithmetic as
// ingest.cu (oracle/encode.py).  CTA 0 also stores the codes / norm bound for finalize.
struct FusedQuery {
    const float* src = nullptr;        // [dim] fp32 on the device; NULL = the query arrives encoded (qcodes)
    int dim = 0;
    int cosine = 0;
    double i8_mult = 127.0;
    void* qcodes_out = nullptr;        // [row_bytes]
    float* qnorm_out = nullptr;        // [1]
    int32_t* zero_word = nullptr;      // the counters the encode kernel resets / advances
    uint32_t* inc_word = nullptr;
};

struct ScanPlan {
    int grid;          // persistent CTAs
    int lpl;           // list entries per lane (1 -> M = 32, 4 -> M = 128)
    const uint32_t* allow = nullptr;   // optional row bitmap (bit r%32 of word r/32): metadata / document filter
    FusedQuery fq;                     // optional fused query encode (single-query scans)
};

// K0: fp32 rows -> stored codes (normalise for cosine, cast / quantise / sign-pack).
//   norms_out (optional, [n]): fp32 upper bound of the Euclidean norm of each STORED row.
cudaError_t launch_encode(cudaStream_t st, const float* src, int64_t n, int dim, int dim_padded,
                          crs_dtype store, crs_metric metric, float i8_scale, void* dst, float* norms_out,
                          int32_t* zero_word = nullptr /*optional device word the kernel sets to 0 (saves a memset node)*/,
                          uint32_t* inc_word = nullptr /*optional device word the kernel increments (exchange step stamp)*/);

// K1: fp16 / bf16 stream scan of `n` stored rows against ONE stored query; writes one
// sorted candidate list of M keys per CTA: cand[cta * M + i].
cudaError_t launch_scan_f16(cudaStream_t st, const void* codes, int64_t n, int dim_padded, bool bf16,
                            const void* qcodes, float tau_pre, uint64_t* cand, const ScanPlan& plan);
// K2 / K3: exact integer scans (dp4a dot, xor+popcount); same candidate-list output.
cudaError_t launch_scan_i8(cudaStream_t st, const void* codes, int64_t n, int dim_padded,
                           const void* qcodes, int32_t min_raw, uint64_t* cand, const ScanPlan& plan);
cudaError_t launch_scan_b1(cudaStream_t st, const void* codes, int64_t n, int dim_padded, int dim,
                           const void* qcodes, int32_t min_raw, uint64_t* cand, const ScanPlan& plan);

// K2 / K3 for a small batch: `nq_group` (8, 4 or 2, from scan_multi_group) consecutive stored queries share
// one pass over rows of 128 / 256 bytes; query j's lists go to cand + j * cand_q_stride.
int scan_multi_group(crs_dtype store, int row_bytes, int lpl, int nq_left, int max_group);
cudaError_t launch_scan_int_multi(cudaStream_t st, crs_dtype store, const void* codes, int64_t n, int row_bytes, int dim,
                                  const void* qcodes, int nq_group, int32_t min_raw, uint64_t* cand, size_t cand_q_stride,
                                  const ScanPlan& plan);

struct FinalizeArgs {
    const uint64_t* cand;      // [nq][n_lists][M] sorted lists
    int n_lists;
    int list_len;              // valid entries per list (<= M = 32*lpl); a list whose last valid slot is
                               // occupied was cut there, and its cut score bounds everything it dropped
    int list_stride;           // keys between consecutive lists of a query (>= list_len)
    int lpl;
    int nq;
    int k;
    int mode;                  // 0: float store, rescore + certify; 1: keys are already exact (ints or fallback)
    int only_flagged;          // 1: process only queries with flags[q] != 0 (fallback pass), then clear the flag
    int certify_exact;         // mode 1 only: lists may be shorter than k -> flag queries whose k-th key does not
                               // beat the largest list cut (integer stores on the tensor-core path)
    const void* codes;         // stored corpus rows (mode 0)
    const void* qcodes;        // [nq][dim_padded]
    const float* qnorms;       // [nq]
    int dim_padded;
    int bf16;
    float eps_rel;             // approx-score error bound = eps_rel * qnorm * row_norm_bound
    float row_norm_bound;
    float min_similarity;      // float-domain threshold on the exact score (mode 0 only)
    uint32_t row_base;
    const uint32_t* id_map;    // optional: global id of every local row (else row_base + row)
    uint32_t* out_ids;         // [nq][k]
    void* out_scores;          // [nq][k] f32 or i32 raw
    int32_t* out_counts;       // [nq]
    int32_t* flags;            // [nq] 1 = not certified (mode 0 writes, only_flagged reads)
    int32_t* n_flagged;        // device counter of uncertified queries (statistics)
    const uint32_t* tau_q;     // mode 0, optional: per-query fast-score floor (ORDERABLE f32, 0 = none) the fast pass
                               // applied: everything below it was dropped, so it counts as a cut
    int is_int;                // raw scores are int32
    int stage_rows;            // set by launch_finalize: candidate rows are staged in shared memory before rescoring
};
cudaError_t launch_finalize(cudaStream_t st, const FinalizeArgs& a);

// Exact fallback: canonical score (fp64-sequential for float stores, integer dot / Hamming
// otherwise) of every row for each flagged query; one sorted list per CTA, keys carry the exact score.
cudaError_t launch_exact_scan(cudaStream_t st, const void* codes, int64_t n, int row_bytes, int dim, crs_dtype store,
                              const void* qcodes, int nq, const int32_t* flags, float min_similarity /*float stores*/,
                              int32_t min_raw /*integer stores*/, uint64_t* cand, const ScanPlan& plan,
                              const int32_t* n_flagged /*device counter; kernel exits at once when it is 0; NULL = always run*/);

// K8: canonical scores of given (query, global row id) pairs; rows of other shards -> -inf / INT32_MIN.
cudaError_t launch_score_rows(cudaStream_t st, const void* codes, int64_t n_rows, uint32_t row_base, int row_bytes,
                              int dim, crs_dtype store, const void* qcodes, const uint32_t* ids, int nq, int m,
                              void* out);

// K4 / K5: tcgen05 Q*C^T with fused top-L epilogue.  kind 0 = fp16, 1 = bf16 (kind::f16, f32
// accumulate, candidates for finalize mode 0), 2 = int8 (kind::i8, exact int32 scores, finalize
// mode 1).  Rows of 128..768 bytes.  Writes one sorted list of gemm_list_len(k) (16 or 32) keys per
// (query, corpus slice), list stride 32; *n_slices_out = number of lists per query.  For k above the
// list length a slice may hold more than a list's worth of the global top-k: finalize detects that
// (certification) and the query is recomputed by the exact fallback.
// tau_pre_bits: threshold as float bits (kind 0/1) or int32 bits (kind 2).
bool gemm_supported(int row_bytes, int k);
int gemm_list_len(int k);
int gemm_n_slices(int64_t n, int nq, int num_sms, int cluster);     // lists per query the launch will write
int gemm_progress_words(int64_t n, int nq, int num_sms, int cluster);   // words of GemmFloorArgs::progress (zeroed by the caller)
// Per-query floors of the contraction.  tau_q: [nq] ORDERABLE 32-bit scores (orderable_f32 / orderable_i32,
// 0 = none), read when a slice starts and — when `pub` is given — raised while the launch runs: the slices
// of a query publish their lists' scores into pub ([nq][n_slices][gemm_list_len(k)] uint32, zeroed by the
// caller) and a helper warp stores the k-th best score of their union (float stores: minus
// margin_rel * |q|) back into tau_q.  After the launch tau_q[q] bounds every floor query q used.
struct GemmFloorArgs {
    uint32_t* tau_q;
    uint32_t* pub;
    const float* qnorms;
    float margin_rel;
    uint32_t* progress = nullptr;   // optional [gemm_progress_words] zeroed words: the clusters of a corpus slice stay in
                                    // lock-step (within a few MB), so that they share the slice through L2
    int lead_tiles = 6;             // tiles (256 rows) a cluster may run ahead of the slowest cluster of its slice
};
cudaError_t launch_gemm_topk(cudaStream_t st, const void* codes, int64_t n, int row_bytes, int kind,
                             const void* qcodes, int nq, int k, uint32_t tau_pre_bits, uint64_t* cand, int num_sms,
                             int cluster /*0 = auto, else 1|2|4 query tiles per multicast cluster*/, int* n_slices_out,
                             const uint32_t* allow /*optional row bitmap, applied to epilogue hits*/,
                             const GemmFloorArgs* floors = nullptr,
                             int prefetch_tiles = 0 /*corpus tiles the producer prefetches into L2 ahead of its ring*/,
                             int warm_tiles = 0 /*first tiles of every slice that only seed the floor and are redone last*/);
// per-query floors (orderable) from the lists of a sample pass (L-th best key minus margin_rel * |q|)
cudaError_t launch_sample_tau(cudaStream_t st, const uint64_t* cand, int n_lists, int list_len, int nq, int is_int,
                              const float* qnorms, float margin_rel, uint32_t* tau_q);

// K7: merge G lists of k_in (id, raw score) per query into the global top k_out.
cudaError_t launch_merge_topk(cudaStream_t st, const uint32_t* ids, const void* scores, int is_int,
                              int n_lists, int nq, int k_in, int k_out,
                              uint32_t* out_ids, void* out_scores, int32_t* out_counts, bool sorted_input = true,
                              size_t list_stride = 0 /*elements between consecutive lists' [nq,k_in] blocks; 0 = nq*k_in*/);

// K7x: exchange + merge over NVLink peer memory (exchange.cu).  Receive buffer of a rank:
//   slots [2 parities][world][2 blocks: ids, raw-score bits][max_nq * max_k] uint32
//   flags [2 parities][world][flag_ctas] uint32 step stamps
constexpr int kMaxWorld = 16;
struct XchgArgs {
    uint32_t* peer_slots[kMaxWorld];   // every rank's slots base (own rank included), peer-mapped
    uint32_t* peer_flags[kMaxWorld];
    const uint32_t* my_slots;
    const uint32_t* my_flags;
    const uint32_t* step_word;         // device word holding the current step stamp (advanced by the search's first kernel)
    uint32_t* err_word;                // set to 1 when a wait times out
    const uint32_t* local_ids;         // this rank's [nq, k] result of the local search
    const uint32_t* local_scores;
    uint32_t* out_ids; void* out_scores; int32_t* out_counts;
    int rank, world, nq, k, max_nq, max_k, flag_ctas, is_int;
    int phase;                         // 0 push + wait + merge, 1 push only, 2 wait + merge only
};
int xmerge_ctas(int nq);
cudaError_t launch_xmerge(cudaStream_t st, const XchgArgs& a);
// every rank's stored rows as seen from this device (peer-mapped): rank r holds global ids [row_base[r], row_base[r] + count[r])
struct PeerShards {
    const uint8_t* codes[kMaxWorld];
    uint32_t row_base[kMaxWorld];
    int64_t count[kMaxWorld];
    int world;
    int row_bytes;
};
cudaError_t launch_peer_gather(cudaStream_t st, const PeerShards& sh, const uint32_t* ids, int n, void* out);

// K6: greedy MMR over m candidate vectors per query.  `fused` (optional): take the search output instead of
// a relevance array and emit the selected hits (ids / similarity / reference score) directly.
struct MmrFusedArgs {
    const uint32_t* ids; const void* raw; const int32_t* counts; float sim_scale;
    uint32_t* out_ids; float* out_sims; double* out_rel; int32_t* out_counts;
};
cudaError_t launch_mmr(cudaStream_t st, const void* vecs, crs_dtype store, int dim_padded, int dim,
                       const double* relevance, int nq, int m, int k_out, double lambda,
                       int32_t* out_order, const MmrFusedArgs* fused = nullptr);

// gather stored rows by local row index
cudaError_t launch_gather_rows(cudaStream_t st, const void* codes, size_t row_bytes, int64_t n_rows,
                               uint32_t row_base, const uint32_t* ids, int n, void* out);

}  // namespace crs
