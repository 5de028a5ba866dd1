/*
 * drs_b200.h -- C ABI of the B200-native hybrid-retrieval hot path.
 *
 * Drop-in boundary for Luyu-H/Domain-RAG-System's retrieval path.  The reference has
 * no FFI of its own (pure Python); these entry points are what a binding for that path
 * binds, one per arithmetic stage the reference delegates to NumPy / rank_bm25 /
 * qdrant-client.  Reference citations are relative to the reference repository root.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - Every pointer named d_* is DEVICE memory owned by the caller; the library never
 *     frees or retains it and allocates no persistent device memory.  Scratch space is
 *     the caller-provided workspace (size from drs_*_workspace_bytes).
 *   - All work is enqueued on the caller's stream (cudaStream_t passed as void*); no
 *     hidden synchronisation, graph-capturable.
 *   - Return value: 0 on success, negative DRS_ERR_* otherwise; drs_last_error() gives
 *     the thread-local message.  There is no CPU fallback anywhere.
 *   - Candidate lists are [Q, k] row-major, sorted by (score desc, row desc) -- the
 *     determinised form of the reference's `np.argsort(scores)[::-1]` (larger row first
 *     among exactly equal scores) -- and padded with id -1.
 */
#ifndef DRS_B200_H_
#define DRS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRS_OK 0
#define DRS_ERR_ARG (-1)         /* null pointer, negative size, misaligned pointer */
#define DRS_ERR_UNSUPPORTED (-2) /* shape outside the compiled limits */
#define DRS_ERR_WORKSPACE (-3)   /* workspace too small */
#define DRS_ERR_CUDA (-4)        /* a CUDA call failed; message has the cudaError string */

#define DRS_MAX_KPRIME 128  /* largest per-retriever list length (k' = 2*top_k) */
#define DRS_MAX_DIM 1024    /* largest embedding dimension; must be a multiple of 8 */
#define DRS_MAX_POOL 1024   /* largest rerank pool */

/* kernel family selector for drs_dense_topk */
#define DRS_DENSE_AUTO 0
#define DRS_DENSE_SIMT 1 /* fp32 CUDA-core streaming kernel (HBM-bound; small Q, any D%8==0) */
#define DRS_DENSE_UMMA 2 /* tcgen05 bf16 GEMM + fused top-k epilogue; picks one of the two below  */
#define DRS_DENSE_UMMA_TS 3   /* single-CTA form, queries parked in tensor memory (D%128==0, D<=768);
                                 serves a single block of <= 128 queries                              */
#define DRS_DENSE_UMMA_PAIR 4 /* CTA-pair form (cta_group::2, M256 x N256, both operands streamed by
                                 TMA; D%64==0): the main kernel for >= 2 blocks of 128 queries        */

int drs_version(void);
const char* drs_last_error(void);

/* Number of SMs / device properties the library sees on the current device (diagnostic). */
int drs_device_sm_count(void);

/* ---------------------------------------------------------------------------------
 * Index build helper: inv_norm[r] = 1/||row r||_2 (0 for an all-zero row).
 * Replaces the upsert-time normalisation qdrant-client's local mode performs for
 * Distance.COSINE (called from src/indexing/vector_indexer.py:149-154).
 * ------------------------------------------------------------------------------- */
int drs_row_inv_norms(const void* d_corpus_bf16, int64_t n_rows, int dim, float* d_inv_norm,
                      void* stream);

/* ---------------------------------------------------------------------------------
 * Exact cosine top-k' of Q queries against all N rows.
 * Replaces QdrantClient.query_points in local mode as called by VectorIndexer.search
 * (src/indexing/vector_indexer.py:202-209): normalise query, scores = M @ q,
 * argsort descending, first `limit` hits.
 *
 *   d_queries_f32  [Q, dim] fp32 (what MedicalEmbedder.encode returns, embedder.py:50-73)
 *   d_corpus_bf16  [N, dim] bf16 row-major, 16-byte aligned
 *   d_inv_norm     [N] from drs_row_inv_norms
 *   d_row_mask     optional [N] uint8, 0 = row excluded BEFORE top-k (payload filter,
 *                  vector_indexer.py:185-196); NULL = no filter
 *   d_out_scores   [Q, kprime] fp32: cosine, exact dot in fp64 rounded once to fp32
 *   d_out_ids      [Q, kprime] int64 row ids, -1 padded
 *   d_out_flags    optional [Q] int32: 0 = the result is PROVEN exact w.r.t. the fp64
 *                  scores; 1 = not proven by the candidate-margin bound (UMMA path only;
 *                  such queries are re-run through the fp32 kernel when `escalate` != 0)
 * ------------------------------------------------------------------------------- */
size_t drs_dense_workspace_bytes(int n_queries, int64_t n_rows, int dim, int kprime, int kernel);
int drs_dense_topk(const float* d_queries_f32, int n_queries, const void* d_corpus_bf16,
                   const float* d_inv_norm, const uint8_t* d_row_mask, int64_t n_rows, int dim,
                   int kprime, int kernel, int escalate, float* d_out_scores, int64_t* d_out_ids,
                   int32_t* d_out_flags, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------
 * BM25 index on the device: term-major CSR inverted index, document-partitioned into
 * tiles of `tile_docs` consecutive documents.
 * Replaces rank_bm25.BM25Okapi (built at src/indexing/bm25_indexer.py:102-106).
 * ------------------------------------------------------------------------------- */
typedef struct drs_bm25_index {
  int64_t n_docs;               /* N                                                     */
  int32_t n_terms;              /* V                                                     */
  int32_t tile_docs;            /* documents per tile, power of two, 256..8192          */
  int32_t n_tiles;              /* ceil(N / tile_docs)                                   */
  int32_t impact_scale_log2;    /* s: stored impacts are round(impact * 2^s), |value| <= 2^26 */
  const int64_t* d_term_offsets;/* [V+1] postings range of each term                     */
  const uint32_t* d_post_doc;   /* [nnz] document id, ascending inside a term            */
  const int32_t* d_post_impact; /* [nnz] fixed point: round(idf*tf*(k1+1)/(tf+kd[doc]) * 2^s).
                                   d_post_doc / d_post_impact are read in aligned groups of four:
                                   allocate (and zero) 4 entries beyond nnz                */
  const uint16_t* d_post_tf;    /* [nnz] term frequency (saturated at 65535)             */
  const uint32_t* d_tile_ptr;   /* [V*(n_tiles+1)] first posting of term t with doc >=
                                   j*tile_docs, relative to d_term_offsets[t]            */
  const double* d_idf;          /* [V] idf incl. the epsilon floor (bm25 `idf` dict)     */
  const double* d_doc_kd;       /* [N] k1*(1-b+b*doc_len/avgdl), fp64, reference op order */
  double k1_plus_1;             /* k1 + 1                                                */
  /* Head terms (present in a large fraction of the documents) additionally carry a DENSE row of
     per-document impacts: adding a row is a vector add (2 bytes/document, no scatter).  Row values are
     14-bit unsigned: round(stored_impact / 2^dense_shift), 0 where the term is absent -- four rows can be
     summed as packed 16-bit lanes without a carry.  Pre-selection scores are INTEGER sums of impacts:
     integer addition is associative, so a document's pre-selection score does not depend on the order
     its terms are visited in and documents with the same (term, impact) multiset tie exactly, as they
     do in the reference.  A head term is ALWAYS added from its row, so the coarser rounding of the rows
     keeps that property; it only widens the error bound of the exactness proof.
     Layout of a row: per tile, 16-byte group i (i < tile_docs/8) holds the four documents 4i..4i+3 of the
     tile's first half followed by the four documents 4i..4i+3 of its second half.                   */
  const int32_t* d_dense_slot;  /* [V] row of the term in d_dense_impact, -1 = sparse only; may be NULL */
  const uint16_t* d_dense_impact;/* [n_dense, dense_stride]                              */
  int64_t dense_stride;         /* n_tiles * tile_docs (rows are zero padded)            */
  int32_t dense_shift;          /* row value * 2^dense_shift ~ stored impact             */
  int32_t impacts_nonneg;       /* 1 = no stored impact is negative (no negative idf floor): scores only
                                   grow while a tile accumulates, which lets the kernel catch threshold
                                   crossings instead of scanning the tile; 0 is always safe         */
} drs_bm25_index;

/* ---------------------------------------------------------------------------------
 * BM25 top-k' of Q tokenised queries.
 * Replaces BM25Okapi.get_scores + np.argsort(scores)[::-1][:k'] as called by
 * BM25Indexer.search (src/indexing/bm25_indexer.py:127-128).
 *
 *   d_q_terms    flattened term ids of every query token OCCURRENCE, in order;
 *                -1 = out-of-vocabulary token (contributes nothing)
 *   d_q_offsets  [Q+1]
 *   d_out_scores [Q, kprime] fp64, bit-identical to rank_bm25's fp64 arithmetic
 *   d_out_ids    [Q, kprime] int64 document ids, -1 padded; zero-score documents ARE
 *                returned when fewer than k' documents match (reference behaviour)
 * ------------------------------------------------------------------------------- */
size_t drs_bm25_workspace_bytes(int n_queries, const drs_bm25_index* index, int kprime);
int drs_bm25_topk(const drs_bm25_index* index, const int32_t* d_q_terms, const int32_t* d_q_offsets,
                  int n_queries, int kprime, double* d_out_scores, int64_t* d_out_ids,
                  void* d_workspace, size_t workspace_bytes, void* stream);
/* Same call with the exactness machinery exposed (same reference lines).  Documents are pre-selected on
 * integer (fixed-point) scores, the k'+margin survivors are re-scored in fp64, and an error bound then
 * PROVES that no other document can belong to the top-k'; queries it cannot prove (large groups of
 * near-ties around the k'-th score) are escalated on the device: every document that can still qualify
 * is re-scored exactly.
 *   mode          0 = prove, escalate the unproven (what drs_bm25_topk does);
 *                 1 = prove only (diagnostics: flag 1 marks unproven queries);
 *                 2 = escalate every query (test hook for the escalation pass)
 *   d_out_flags   [Q] int32 or NULL: 0 proven, 1 unproven (mode 1 only), 2 resolved by escalation   */
int drs_bm25_topk_ex(const drs_bm25_index* index, const int32_t* d_q_terms, const int32_t* d_q_offsets,
                     int n_queries, int kprime, int mode, double* d_out_scores, int64_t* d_out_ids,
                     int32_t* d_out_flags, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------
 * Score fusion + final top-K.
 * Replaces HybridIndexer._reciprocal_rank_fusion (src/indexing/hybrid_indexer.py:112-156)
 * and HybridIndexer._weighted_fusion (:158-218), arithmetic in fp64 in the reference's
 * operation order; stable order (ties keep insertion order, vector list first).
 *
 *   mode            0 = RRF (uses rrf_k, default 60), 1 = weighted (uses w_vector, w_bm25)
 *   d_v_* / d_b_*   the two candidate lists [Q, kv] / [Q, kb]; id -1 = absent.  Vector scores
 *                   are fp32 (v_scores_f64 = 0, what drs_dense_topk emits; widened exactly) or
 *                   fp64 (v_scores_f64 = 1, caller-built lists); BM25 scores are fp64
 *   d_fusion_key    optional [n_rows] int64: fusion identity of a row (the reference fuses
 *                   on chunk_id STRINGS; rows sharing a chunk_id share a key).  NULL = row id
 *   d_out_ids       [Q, K] representative row of each fused entry (-1 padded)
 *   d_out_scores    [Q, K] fp64 fused score
 *   d_out_norm      optional [Q, K] fp64: the `normalized_score` the weighted mode leaks
 *                   into the representative dict (hybrid_indexer.py:184), NaN if not set
 * ------------------------------------------------------------------------------- */
int drs_fuse_topk(int mode, int n_queries, int kv, int kb, const void* d_v_scores, int v_scores_f64,
                  const int64_t* d_v_ids, const double* d_b_scores, const int64_t* d_b_ids,
                  const int64_t* d_fusion_key, double rrf_k, double w_vector, double w_bm25, int top_k,
                  int64_t* d_out_ids, double* d_out_scores, double* d_out_norm, void* stream);

/* ---------------------------------------------------------------------------------
 * "simple" cosine rerank of a pool of rows.
 * Replaces SimpleSimilarityReranker.score_pairs/_cosine + the stable sort of
 * BaseReranker.rerank (src/reranker/rerank.py:53-68, :29-43).  Bit-identical to the
 * reference under NumPy >= 2: three sequential fp32 sums, fp64 sqrt, fp32 divide.
 *
 *   d_rq_f32       [Q, dim] fp32 embedding of the RAW query (scripts/query.py:160)
 *   d_pool_ids     [Q, top_n] rows in fused order, -1 padded
 *   d_out_scores   [Q, top_k] fp32 rerank_score;  d_out_ids [Q, top_k] rows;
 *   d_out_pos      optional [Q, top_k] int32 position of each output in the pool
 * ------------------------------------------------------------------------------- */
int drs_rerank_cosine(const void* d_corpus_bf16, int dim, const float* d_rq_f32, const int64_t* d_pool_ids,
                      int n_queries, int top_n, int top_k, float* d_out_scores, int64_t* d_out_ids,
                      int32_t* d_out_pos, void* stream);

/* ---------------------------------------------------------------------------------
 * Merge of P candidate lists per query into one (the step after the NCCL all-gather of
 * per-GPU top-k' lists; ids must already be GLOBAL rows).  Input [P, Q, k], output [Q, k_out],
 * order (score desc, id desc).  No reference counterpart (the reference is single-process);
 * equals argsort over the concatenation.
 * ------------------------------------------------------------------------------- */
int drs_merge_candidates_f32(const float* d_scores, const int64_t* d_ids, int n_parts, int n_queries,
                             int k, int k_out, float* d_out_scores, int64_t* d_out_ids, void* stream);
int drs_merge_candidates_f64(const double* d_scores, const int64_t* d_ids, int n_parts, int n_queries,
                             int k, int k_out, double* d_out_scores, int64_t* d_out_ids, void* stream);

/* ---------------------------------------------------------------------------------
 * Row-sharded corpora (one process per GPU; no reference counterpart: the reference is single-process).
 * Each rank fills ONE send buffer with its local lists, an all-gather concatenates the ranks' buffers
 * ("parts", part_stride_bytes apart), and these entry points work on the gathered buffer in place.
 *
 * drs_candidate_cosines: the "simple" reranker's cosine (src/reranker/rerank.py:53-58, the exact fp32
 *   arithmetic of drs_rerank_cosine) of the rerank query against every LOCAL candidate row of a
 *   [Q, n] list (-1 -> 0), so that the fused pool can be reranked after the exchange without a second
 *   one: the cosine travels with the candidate.
 * drs_merge_gathered: merge of the P parts' [Q, k] lists into [Q, k_out], order (score desc, id desc).
 *   Scores at off_scores (fp32, or fp64 when scores_f64), int64 LOCAL ids at off_ids, optional fp32
 *   cosines at off_cos inside each part; d_id_offset[p] = global row of part p's row 0 (NULL: ids are
 *   global already).  Equals drs_merge_candidates_* on the unpacked lists.
 * drs_rerank_from_cosines: BaseReranker.rerank's stable sort (src/reranker/rerank.py:29-43) of the fused
 *   pool [Q, top_n] by cosines looked up in the merged vector / BM25 lists; equals drs_rerank_cosine
 *   on the unsharded corpus bit for bit.
 * ------------------------------------------------------------------------------- */
int drs_candidate_cosines(const void* d_corpus_bf16, int dim, const float* d_rq_f32, const int64_t* d_ids,
                          int n_queries, int n, float* d_out_cos, void* stream);
int drs_merge_gathered(const void* d_gathered, int n_parts, size_t part_stride_bytes, size_t off_scores,
                       int scores_f64, size_t off_ids, size_t off_cos, int has_cos, const int64_t* d_id_offset,
                       int n_queries, int k, int k_out, void* d_out_scores, int64_t* d_out_ids, float* d_out_cos,
                       void* stream);
int drs_rerank_from_cosines(const int64_t* d_pool_ids, int n_queries, int top_n, int top_k, const int64_t* d_v_ids,
                            const float* d_v_cos, int kv, const int64_t* d_b_ids, const float* d_b_cos, int kb,
                            float* d_out_scores, int64_t* d_out_ids, int32_t* d_out_pos, void* stream);

/* ---------------------------------------------------------------------------------
 * Query tokenisation + vocabulary lookup for a batch of ASCII query strings.
 * Replaces BM25Indexer.tokenize (src/indexing/bm25_indexer.py:50-78: text.lower(), re.findall(r'\b\w+\b'),
 * one-character tokens kept only when they are abbreviation keys) and the per-token dictionary lookups of
 * rank_bm25.BM25Okapi.get_scores as called from BM25Indexer.search (src/indexing/bm25_indexer.py:125-127).
 * For ASCII bytes those semantics are byte-level and reproduced exactly; strings with any byte >= 0x80 must be
 * tokenised by the caller (Unicode-defined lower() / \w) -- pass them as empty strings here.
 *
 * d_text: the strings' UTF-8 bytes back to back; d_text_offsets int64 [n_texts + 1].
 * drs_single_chars: bit c set = the one-character token c (lower case) is kept.
 * drs_vocab_table: open addressing keyed by 64-bit FNV-1a (offset basis 0xcbf29ce484222325, prime 0x100000001b3) of
 *   the lower-cased token; home slot = hash & (capacity - 1), capacity a power of two; entries sit at or after their
 *   home slot in home-slot order WITHOUT wrap-around (n_slots >= capacity + number of entries), empty slots hold
 *   term -1.  d_str_offsets int64 [V + 1] / d_pool: the vocabulary strings by term id, used to confirm a hash match
 *   byte by byte (the lookup is exact).  capacity 0 = no vocabulary (every token maps to -1).
 * drs_tokenize_count: d_counts[q] = tokens of text q.  drs_tokenize_lookup: term ids (or -1: out of vocabulary,
 *   which drs_bm25_topk skips like rank_bm25 does) written at d_out_terms[d_out_offsets[q] ..), d_out_offsets being
 *   the exclusive prefix sum of the counts -- i.e. the (q_terms, q_offsets) pair drs_bm25_topk reads.
 * ------------------------------------------------------------------------------- */
typedef struct drs_single_chars {
  uint32_t bits[8];
} drs_single_chars;
typedef struct drs_vocab_table {
  const uint64_t* d_slot_hash;   /* [n_slots] */
  const int32_t* d_slot_term;    /* [n_slots] term id or -1 */
  const int64_t* d_str_offsets;  /* [V + 1] */
  const uint8_t* d_pool;         /* vocabulary strings (lower case), term id order */
  int64_t capacity;              /* power of two, or 0 */
  int64_t n_slots;
} drs_vocab_table;
int drs_tokenize_count(const uint8_t* d_text, const int64_t* d_text_offsets, int n_texts,
                       const drs_single_chars* keep1, int32_t* d_counts, void* stream);
int drs_tokenize_lookup(const uint8_t* d_text, const int64_t* d_text_offsets, int n_texts,
                        const drs_single_chars* keep1, const drs_vocab_table* table, const int32_t* d_out_offsets,
                        int32_t* d_out_terms, void* stream);

/* Which kernel the last drs_dense_topk call on this thread used: DRS_DENSE_SIMT (1), DRS_DENSE_UMMA (2, the
 * single-CTA tensor kernel) or 3 (the CTA-pair tensor kernel). */
int drs_dense_last_kernel(void);

/* ---------------------------------------------------------------------------------
 * Measurement hooks (bench.py): when enabled, the library brackets its kernels with CUDA
 * events on the caller's stream, per kernel family, so the device time of the dominant
 * kernel can be read without a profiler.  drs_profile_read synchronises on the recorded
 * events and returns the summed milliseconds and the number of launches since the last
 * drs_profile_enable(1).
 * ------------------------------------------------------------------------------- */
#define DRS_PROF_DENSE_SELECT 0 /* tcgen05 GEMM+top-k kernel or the fp32 streaming kernel */
#define DRS_PROF_DENSE_TAIL 1   /* merge + exact re-score (+ escalation)                   */
#define DRS_PROF_BM25_SELECT 2  /* postings scatter/accumulate + tile top-k                */
#define DRS_PROF_BM25_TAIL 3    /* merge + fp64 re-score                                   */
#define DRS_PROF_FUSE 4
#define DRS_PROF_RERANK 5
#define DRS_PROF_KINDS 6
/* kernels launched by this library in this process so far */
unsigned long long drs_kernel_launches(void);
int drs_profile_enable(int on);
int drs_profile_read(int kind, double* total_ms, int* launches);

#ifdef __cplusplus
}
#endif
#endif /* DRS_B200_H_ */
