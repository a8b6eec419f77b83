/* crs.h — C ABI of libcrs.so: exact similarity search on B200 (sm_100a).
 *
 * This is the drop-in boundary for the one hot path of
 * zahraamselim/compressed-rag-suite that this repository replaces: the calls
 * `rag/indexing.py` makes into ChromaDB and the numeric half of
 * `rag/retrieval.py`.  Every entry point names the reference interface it
 * replaces (paths relative to the reference tree).  The reference is Python and
 * binds nothing native today; a maintainer would bind these symbols with the
 * ctypes stub shown in INTEGRATION.md (that stub is what
 * compressed_rag_suite_b200/_native.py ships).
 *
 * Conventions
 *  - plain C, no exceptions across the boundary: every function returns 0
 *    (CRS_OK) or a CRS_E* code; crs_last_error() gives the thread-local message.
 *    The Python host turns a non-zero status into the ValueError/RuntimeError
 *    the reference raises after logging (rag/indexing.py:121-123,178-180).
 *  - one index = one GPU shard (one process per GPU; cross-GPU exchange is done
 *    by the host with torch.distributed/NCCL and crs_merge_topk).  Row ids are
 *    uint32 insertion indices, global id = row_base + local row.
 *  - in/out buffers belong to the caller and may be host or device pointers
 *    (detected with cudaPointerGetAttributes).  With device buffers a call only
 *    enqueues work on the index's stream (crs_index_set_stream); with host
 *    buffers it returns after the results are in the host buffer.
 *  - calls on one index are serialised by an internal mutex.
 *  - there is no CPU implementation behind these symbols: without a usable
 *    sm_100 device crs_index_create fails with CRS_ECUDA.
 */
#ifndef CRS_H_
#define CRS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crs_index crs_index;

typedef enum { CRS_F32 = 0, CRS_F16 = 1, CRS_BF16 = 2, CRS_I8 = 3, CRS_B1 = 4 } crs_dtype;
/* Chroma "hnsw:space" values the reference can meet (rag/indexing.py:83 always
 * creates "cosine"; rag/retrieval.py:84-87 also knows "ip"). */
typedef enum { CRS_COSINE = 0, CRS_IP = 1 } crs_metric;

enum { CRS_OK = 0, CRS_EINVAL = 1, CRS_ECUDA = 2, CRS_ENOMEM = 3, CRS_ESTATE = 4, CRS_EIO = 5 };

#define CRS_PAD_ID 0xFFFFFFFFu

/* Counters of the last crs_index_search on an index (bench.py's gpu_launches). */
typedef struct crs_search_stats {
    int32_t kernel_launches;   /* kernels of libcrs launched by the call */
    int32_t path;              /* 0 = stream scan (K1/K2/K3), 1 = tcgen05 GEMM (K4/K5) */
    int32_t grid;              /* CTAs of the dominant kernel */
    int32_t list_len;          /* candidates kept per query per CTA (M) */
    int64_t uncertified_total; /* float stores: queries that needed the exact fp64 pass, since create */
    int64_t searches_total;
    float max_fast_error;      /* float stores: largest |fast score - exact score| of any rescored candidate
                                  since create (the certification bound eps must stay above it) */
    int32_t reserved;
} crs_search_stats;

const char* crs_last_error(void);
int crs_version(void);

/* replaces chromadb.Client()/PersistentClient() + client.create_collection(name,
 * metadata={"hnsw:space": ...})  — rag/indexing.py:31-37,81-84.
 *   store        : F16 | BF16 | I8 | B1 (how rows are kept in HBM)
 *   device       : CUDA ordinal;  row_base: global id of local row 0
 *   reserve_rows : capacity hint (0 = grow on demand) */
int crs_index_create(crs_index** out, int dim, crs_dtype store, crs_metric metric,
                     int device, uint32_t row_base, int64_t reserve_rows);
/* replaces client.delete_collection(name) — rag/indexing.py:186 */
int crs_index_destroy(crs_index* idx);
/* cudaStream_t the index enqueues on (NULL = default stream). */
int crs_index_set_stream(crs_index* idx, void* cuda_stream);

/* replaces collection.add(embeddings=...) — rag/indexing.py:114-119 (the
 * ids/documents/metadatas of that call stay with the Python host).
 *   rows: [n, dim] row-major, src_dtype must be CRS_F32; host or device.
 * Inner-product space (CRS_IP) needs a float store: I8 / B1 codes are defined on unit rows. */
int crs_index_add(crs_index* idx, const void* rows, int64_t n, crs_dtype src_dtype);
/* gives the local rows [first_row, first_row + n) the global ids first_global_id, first_global_id + 1, ...
 * instead of row_base + row: for a host that deals the rows of one collection out to several indexes in
 * turns (one GPU each) and still wants the insertion index as the global id.  Ids must grow with the local
 * row (ties inside a shard go to the lowest LOCAL row).  Searches report the mapped ids;
 * crs_index_fetch_rows / crs_index_score_rows keep addressing rows as row_base + local row. */
int crs_index_map_ids(crs_index* idx, int64_t first_row, int64_t n, uint32_t first_global_id);
/* replaces collection.count() — rag/indexing.py:52,120,147,152,206 */
int crs_index_count(const crs_index* idx, int64_t* out_count);
/* dim, padded dim, bytes per stored row, store dtype, metric */
int crs_index_info(const crs_index* idx, int32_t* dim, int32_t* dim_padded, int64_t* row_bytes,
                   int32_t* store, int32_t* metric);

/* replaces collection.query(query_embeddings, n_results) — rag/indexing.py:171-176 —
 * plus the similarity_threshold filter of rag/retrieval.py:143 pushed down as
 * min_similarity (cosine/dot domain, -INFINITY = off).
 *   queries    : [nq, dim] fp32, host or device
 *   out_ids    : [nq, k] global row ids, descending score, ties -> lowest id, CRS_PAD_ID padded
 *   out_scores : [nq, k] raw inner product of the stored codes — float32 for F16/BF16,
 *                int32 for I8 (dot) and B1 (dim - 2*hamming); pad = -inf / INT32_MIN
 *   out_counts : [nq] number of valid entries per query */
int crs_index_search(crs_index* idx, const void* queries, int nq, int k, float min_similarity,
                     uint32_t* out_ids, void* out_scores, int32_t* out_counts);
/* the same with the `where` / `where_document` arguments of collection.query
 * (rag/indexing.py:129-130,174-175; rag/retrieval.py:97,120): the host evaluates the
 * metadata / document predicates (they are string and dict work) and passes the result as a
 * row bitmap — bit (r % 32) of word (r / 32) set = local row r may be returned;
 * ceil(count / 32) words, host or device.  NULL = no filter. */
int crs_index_search_filtered(crs_index* idx, const void* queries, int nq, int k, float min_similarity,
                              const uint32_t* allow_bits, uint32_t* out_ids, void* out_scores,
                              int32_t* out_counts);
int crs_index_last_stats(const crs_index* idx, crs_search_stats* out);
/* Tuning / test hooks (defaults in brackets):
 *   "force_path"    [-1] -1 auto, 0 stream scans (K1-K3), 1 tensor-core contraction (K4/K5)
 *   "gemm_min_nq"   [2]  smallest batch that takes the contraction
 *   "gemm_cluster"  [0]  0 auto (2 query tiles per TMA-multicast cluster), 1 | 2 | 4, 22 = CTA-pair MMA (cta_group::2)
 *   "gemm_prefetch" [0]  corpus tiles prefetched into L2 ahead of the TMA ring
 *   "share_floor"   [1]  contraction: the corpus slices of a query share their k-th best score while the launch runs
 *   "gemm_warm"     [8]  contraction: first tiles of every slice that only seed the floor and are computed again last
 *   "gemm_lockstep" [6]  contraction: tiles a cluster may run ahead of the slowest cluster streaming the same corpus slice
 *                        (they share the slice through L2 only while they stay close); 0 = off
 *   "fuse_encode"   [1]  single-query scans encode the query in their own prologue instead of a separate launch
 *   "sample_rows"   [0]  rows of an optional sample pass that seeds the contraction's per-query floors (0 = off)
 *   "multi_scan"    [8]  largest group of short-row integer queries that shares one corpus pass (<= 1 = off)
 *   "short_lists"   [1]  integer scans with k > 32 keep 32 keys per CTA + certification (0 = full 128-key lists)
 *   "force_exact"   [0]  1 = skip the fast pass of float stores, always run the exhaustive fp64 pass
 *   "eps_scale"     [1000] certification error bound x value/1000
 *   "profiling"     [0]  1 = bracket the dominant kernel(s) of each search with CUDA events on the index's stream
 * None of them changes a result: every combination returns the canonical top-k (tests compare them). */
int crs_index_set_option(crs_index* idx, const char* name, int64_t value);
/* device time of the dominant kernel(s) (scan passes or GEMM) of the last search, from the
 * CUDA events recorded when "profiling" is on; waits for that search to finish. */
int crs_index_last_kernel_ms(crs_index* idx, float* out_ms);
/* the same for up to the last 32 searches (oldest first), so a benchmark can read the
 * per-step kernel times after its timed loop without synchronising inside it. */
int crs_index_kernel_ms_history(crs_index* idx, float* out_ms, int max_n, int* n_out);

/* raw -> float similarity scale of this index: sim = raw * scale (1 for F16/BF16,
 * (a/127)^2 for I8, 1/dim for B1). */
int crs_index_similarity_scale(const crs_index* idx, double* out_scale);

/* copies stored rows (codes, row_bytes each) of global ids into out; ids outside this
 * shard are skipped (their output rows are left untouched).  Feeds crs_mmr. */
int crs_index_fetch_rows(crs_index* idx, const uint32_t* ids, int n, void* out_codes);

/* K8, candidate rescoring (no reference counterpart: BASELINE config 5, "Hamming top-100 with
 * fp16 rescoring of candidates"; also how a coarse index of one store dtype is refined by a
 * finer one).  Canonical score — the value crs_index_search reports for that row — of every
 * (query q, row ids[q][j]) pair:
 *   queries    : [nq, dim] fp32, host or device
 *   ids        : [nq, m] global row ids (CRS_PAD_ID allowed), host or device
 *   out_scores : [nq, m] float32 (F16/BF16) or int32 (I8/B1); rows this shard does not hold
 *                and pad ids get -inf / INT32_MIN, so a MAX over shards assembles the result */
int crs_index_score_rows(crs_index* idx, const void* queries, int nq, const uint32_t* ids, int m,
                         void* out_scores);
/* the same for candidate vectors the CALLER supplies instead of stored rows (BASELINE config 5 at
 * full size: the fp16 originals of 1 B x 1024-d rows do not fit in HBM, so the candidates'
 * fp32 vectors are re-materialised or read from host memory): the rows go through this index's
 * encoder (normalise + round to the store dtype) and are scored canonically.
 *   rows : [nq, m, dim] fp32, host or device;  out_scores : [nq, m] */
int crs_index_score_vectors(crs_index* idx, const void* queries, int nq, const void* rows, int m,
                            void* out_scores);
/* orders UNSORTED candidates (e.g. rescored ones) by (score desc, id asc) and keeps k_out:
 *   ids/scores : [nq, m] device pointers, m <= 128; pad ids are skipped */
int crs_select_topk(void* cuda_stream, const uint32_t* ids, const void* scores, int is_int,
                    int nq, int m, int k_out, uint32_t* out_ids, void* out_scores, int32_t* out_counts);

/* replaces ContextRetriever._apply_diversity — rag/retrieval.py:219-277 — on stored
 * vectors instead of re-embedded texts.
 *   vecs      : [nq, m, row_bytes] stored codes of the m candidates (position order)
 *   relevance : [nq, m] the chunks' `score` values (Python floats -> double)
 *   lambda    : 1 - diversity_penalty
 *   out_order : [nq, k_out] positions 0..m-1 in greedy MMR order (first = position 0) */
int crs_mmr(crs_index* idx, const void* vecs, const double* relevance, int nq, int m, int k_out,
            double lambda, int32_t* out_order);

/* the same selection fed straight from a search (BASELINE config 4: top-100 -> MMR -> 10, one
 * launch): takes the candidates' stored codes plus the search output and derives, per candidate,
 * the cosine-domain similarity and the reference's `score` (rag/retrieval.py:75-77 applied to the
 * Chroma distance 1 - similarity, in IEEE double arithmetic as Python evaluates it); candidates
 * at positions >= counts[q] are padding.  Device buffers only.
 *   out_ids [nq,k_out] (pad CRS_PAD_ID), out_sims [nq,k_out] f32, out_scores [nq,k_out] f64
 *   (the reference's `score`), out_counts [nq] */
int crs_mmr_select(crs_index* idx, const void* vecs, const uint32_t* ids, const void* raw_scores,
                   const int32_t* counts, int nq, int m, int k_out, double lambda,
                   uint32_t* out_ids, float* out_sims, double* out_scores, int32_t* out_counts);

/* cross-shard merge after the allgather of local top-k lists (no reference
 * counterpart: the reference is single-node).
 *   ids/scores : [n_lists, nq, k_in] device pointers;  is_int: scores are int32 */
int crs_merge_topk(void* cuda_stream, const uint32_t* ids, const void* scores, int is_int,
                   int n_lists, int nq, int k_in, int k_out,
                   uint32_t* out_ids, void* out_scores, int32_t* out_counts);

/* the same when rank l's [nq, k_in] block starts `list_stride` elements after rank l-1's (ids and
 * scores alike): lets each rank allgather ONE buffer holding its ids block followed by its scores
 * block and merge straight out of the gathered buffer, with no pack / unpack copies. */
int crs_merge_topk_strided(void* cuda_stream, const uint32_t* ids, const void* scores, int is_int,
                           int n_lists, int nq, int k_in, int k_out, int64_t list_stride,
                           uint32_t* out_ids, void* out_scores, int32_t* out_counts);

/* ---- cross-shard exchange over NVLink peer memory (no reference counterpart: the reference is single-node).
 * The sharded search of the north star — local top-k per GPU, exchange of the k*(id, score) candidates, final
 * merge — as ONE kernel after the local search: every rank stores its [nq, k] lists straight into every
 * peer's receive buffer (peer-mapped global memory), flags them, waits for the peers' flags and merges.
 * The NCCL form (allgather + crs_merge_topk[_strided]) stays available; both give identical results.
 *
 *   one crs_exchange per rank (= per GPU), sized for the largest nq and k it will carry (k <= 128);
 *   ranks in different processes:  crs_exchange_ipc_handle -> allgather the 64-byte handles with any host
 *                                  mechanism (torch.distributed, MPI, a file) -> crs_exchange_open_peers;
 *   ranks in one process:          crs_exchange_buffer of every rank -> crs_exchange_set_peer_buffers
 *                                  (enables peer access between the devices).
 * Every rank must run the same sequence of sharded searches (same nq and k per step). */
typedef struct crs_exchange crs_exchange;
int crs_exchange_create(crs_exchange** out, int device, int rank, int world, int max_nq, int max_k);
int crs_exchange_destroy(crs_exchange* ex);
int crs_exchange_ipc_handle(crs_exchange* ex, void* out_handle64);
int crs_exchange_open_peers(crs_exchange* ex, const void* handles /* world x 64 bytes, rank order */);
int crs_exchange_buffer(crs_exchange* ex, void** out_ptr);
int crs_exchange_set_peer_buffers(crs_exchange* ex, void* const* bufs /* world device pointers, rank order */);
/* waits for `cuda_stream`, then reports the step stamp and whether any wait for a peer timed out (~2 s) */
int crs_exchange_status(crs_exchange* ex, void* cuda_stream, int* timed_out, uint32_t* step);
/* crs_index_search on this rank's shard + exchange + merge: every rank ends with the GLOBAL top-k
 * (same buffers and conventions as crs_index_search). */
int crs_index_search_sharded(crs_index* idx, crs_exchange* ex, const void* queries, int nq, int k,
                             float min_similarity, uint32_t* out_ids, void* out_scores, int32_t* out_counts);
/* the same in two calls, for a host that drives several GPUs from one thread: push = local search + stores
 * to the peers (waits for nobody); merge = wait for the peers' pushes of this step + merge (device buffers). */
int crs_index_search_push(crs_index* idx, crs_exchange* ex, const void* queries, int nq, int k, float min_similarity,
                          const uint32_t* allow_bits /* optional row bitmap as in crs_index_search_filtered, NULL = none */);
int crs_exchange_merge(crs_exchange* ex, void* cuda_stream, int nq, int k, int is_int,
                       uint32_t* out_ids, void* out_scores, int32_t* out_counts);

/* Candidate vectors without a second collective (BASELINE configs 4 and 5: MMR over the top-100, fp16 rescoring of
 * Hamming candidates): every rank maps the other ranks' stored rows (CUDA IPC handle of the code buffer, or the
 * pointer itself inside one process) and reads the candidates' rows straight out of the owning GPU's HBM over
 * NVLink.  Register the shards after the corpus is built (growing an index re-allocates its rows).
 *   crs_index_codes_handle : IPC handle (64 bytes, may be NULL) / device pointer / id range of this index's rows
 *   crs_exchange_open_shards / _set_shards : rank-ordered tables; `own` is this rank's index
 *   crs_exchange_fetch_rows  : ids [n] (device) -> stored codes [n, row_bytes] (device); pad ids -> zero rows
 *   crs_exchange_score_rows  : K8 on global ids wherever the rows live: queries [nq, dim] fp32, ids [nq, m]
 *                              (device) -> canonical scores [nq, m] of idx's store dtype (pad ids: caller masks) */
int crs_index_codes_handle(crs_index* idx, void* out_handle64, void** out_ptr, uint32_t* row_base, int64_t* count);
int crs_exchange_open_shards(crs_exchange* ex, crs_index* own, const void* handles /* world x 64 bytes */,
                             const uint32_t* row_bases, const int64_t* counts);
int crs_exchange_set_shards(crs_exchange* ex, crs_index* own, void* const* codes_ptrs, const uint32_t* row_bases,
                            const int64_t* counts);
int crs_exchange_fetch_rows(crs_exchange* ex, void* cuda_stream, const uint32_t* ids, int n, void* out_codes);
int crs_exchange_score_rows(crs_index* idx, crs_exchange* ex, const void* queries, int nq, const uint32_t* ids, int m,
                            void* out_scores);

/* replaces chromadb.PersistentClient(path) persistence + get_collection reload —
 * rag/indexing.py:32-34,46-55.  Raw code blob + small header; the host keeps
 * ids/documents/metadatas in a sidecar. */
/* whole index -> path, atomically (written to "<path>.tmp", flushed, renamed) */
int crs_index_save(crs_index* idx, const char* path);
/* appends the rows `path` does not hold yet and then advances its header (incremental indexing; a crash in
 * between leaves a consistent prefix) */
int crs_index_append(crs_index* idx, const char* path);
/* header fields are validated; a row count beyond the rows present in the file is clamped (torn append) */
int crs_index_load(crs_index** out, const char* path, int device, uint32_t row_base);
/* drops the rows from new_count on (a host sidecar that holds fewer rows than the blob after a crash) */
int crs_index_truncate(crs_index* idx, int64_t new_count);

#ifdef __cplusplus
}
#endif
#endif /* CRS_H_ */
