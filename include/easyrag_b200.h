/*
 * easyrag_b200 -- C ABI of the B200-native coarse-ranking path.
 *
 * The reference (BUAADreamer/EasyRAG) has no native code and therefore no FFI
 * for this path: its boundary is the Python class surface of
 * src/easyrag/custom/retrievers.py.  Each entry point below replaces the
 * arithmetic behind one of those methods; the Python classes in
 * easyrag_b200/retrievers.py keep the reference's signatures and call these
 * through ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success, a negative ezr_status otherwise;
 *     ezr_last_error() returns a thread-local message for the last failure.
 *   - all pointers are DEVICE pointers unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - launch functions never synchronise and never allocate: outputs and the
 *     workspace are caller-owned (query the size with the *_workspace call).
 *   - document ids are int32, local to the shard the index was built over;
 *     `id_base` is added on output so a row-sharded corpus yields global ids.
 *   - canonical rank order everywhere: score descending, then id descending
 *     (== numpy argsort(kind="stable")[::-1]).
 *   - the library targets sm_100a only; ezr_device_check() fails elsewhere.
 */
#ifndef EASYRAG_B200_H
#define EASYRAG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ezr_status {
    EZR_OK = 0,
    EZR_ERR_INVALID = -1,
    EZR_ERR_CUDA = -2,
    EZR_ERR_WORKSPACE = -3,
    EZR_ERR_UNSUPPORTED = -4,
    EZR_ERR_ARCH = -5
} ezr_status;

typedef enum ezr_score_type {
    EZR_F64 = 0, /* rank_bm25.BM25Okapi, bm25_type 0 (retrievers.py:113-118) */
    EZR_F32 = 1  /* bm25s, bm25_type 1 (retrievers.py:107-111); dense cosine */
} ezr_score_type;

int ezr_version(void);
const char* ezr_last_error(void);
/* fails with EZR_ERR_ARCH unless the current device is compute capability 10.x */
int ezr_device_check(void);

/* ------------------------------------------------------------------ BM25 --
 * Term-major (CSC) posting lists with per-posting precomputed contribution
 *   w = idf[t] * tf*(k1+1) / (tf + k1*((1-b) + b*dl/avgdl))
 * evaluated with IEEE round-to-nearest, no FMA contraction, in exactly the
 * operation order numpy applies to rank_bm25's expression (oracle/bm25.py).
 * Documents are cut into ranges of `range_size` ids; range_off[t*(n_ranges+1)+r]
 * is the offset (relative to indptr[t]) of the first posting of term t whose
 * document id is >= r*range_size.
 */
typedef struct ezr_bm25_index {
    int64_t n_docs;
    int64_t n_postings;
    int32_t vocab;
    int32_t score_type;        /* ezr_score_type of post_w */
    int32_t range_size;        /* ezr_bm25_range_size(): 8192 in this build */
    int32_t n_ranges;          /* ceil(n_docs / range_size) */
    const int64_t* indptr;     /* [vocab+1] */
    const int32_t* post_doc;   /* [n_postings] ascending inside a term */
    const void* post_w;        /* [n_postings] double or float */
    const uint32_t* range_off; /* [vocab*(n_ranges+1)] */
    const int32_t* doc_group;  /* [n_docs] metadata class of each document, or NULL */
    int32_t monotone;          /* 1 if every post_w >= 0 (no negative idf): enables crossing-based selection */
    int32_t pk_scale_log2;     /* e of ezr_bm25_pack (informational) */
    const uint32_t* post_pk;   /* [n_postings] packed postings from ezr_bm25_pack, or NULL: enables the two-phase
                                  top-k (integer candidate pass + exact float64 rescoring) for F64 / monotone / k<=32 */
    const uint32_t* term_max;  /* [vocab] from ezr_bm25_term_max, or NULL: lets the candidate pass skip the posting
                                  lists of a query's lowest-weight terms (MaxScore); results are unchanged */
} ezr_bm25_index;

/* documents per range the library was built for (the `range_size` an index must use) */
int ezr_bm25_range_size(void);

/* K_d[i] = k1 * (one_minus_b + (b*doc_len[i]) / avgdl)   -- rank_bm25 get_scores denominator term */
int ezr_bm25_doc_norm(const int32_t* doc_len, int64_t n_docs, double k1, double b, double one_minus_b,
                      double avgdl, double* out_kd, void* stream);

/* out_w[p] = (S)( idf[term(p)] * ((tf*num_scale) / (tf + K_d[doc])) ); num_scale = k1+1 (Okapi) or 1 (bm25s) */
int ezr_bm25_weights(const int64_t* indptr, const int32_t* post_doc, const int32_t* post_tf, int32_t vocab,
                     int64_t n_postings, const double* idf, const double* kd, double num_scale,
                     int32_t score_type, void* out_w, void* stream);

int ezr_bm25_range_index(const int64_t* indptr, const int32_t* post_doc, int32_t vocab, int32_t range_size,
                         int32_t n_ranges, uint32_t* out_range_off, void* stream);

/* ---- index construction (replaces the dict building of BM25Retriever.__init__, retrievers.py:98-118) ----
 * Corpus = int32 term ids tokens[n_tokens] + int64 doc_ptr[n_docs+1] (device).  Two phases because the number of
 * postings is only known after counting:
 *   count: df[vocab], indptr[vocab+1] (int64), first_pos[vocab] (uint64 corpus position of each term's first
 *          occurrence, ~0 = absent: rank_bm25 sums idf in first-seen term order) + the workspace for phase two.
 *          Documents longer than 8192 tokens sort in the global scratch long_keys (long_docs[n_long] document indices,
 *          long_off[n_long] key offsets, sum(next_pow2(len)) uint64 keys); NULL / 0 when there are none.
 *          *status_host = 0, or 1 + the index of a document holding a token id outside [0, vocab).  Synchronises.
 *   fill : post_doc / post_tf [indptr[vocab]], term-major, ascending document id inside a term (no sort: documents
 *          are placed block by block in order).
 * ezr_bm25_shard_*: postings of documents [doc_lo, doc_hi) of a built index (ids rebased to doc_lo). */
int ezr_bm25_build_block(void);
size_t ezr_bm25_build_workspace(int64_t n_docs, int64_t n_tokens, int32_t vocab);
int ezr_bm25_build_count(const int32_t* tokens, const int64_t* doc_ptr, int64_t n_docs, int64_t n_tokens, int32_t vocab,
                         int32_t max_doc_len, int64_t* out_df, int64_t* out_indptr, uint64_t* out_first_pos,
                         const int32_t* long_docs, const int64_t* long_off, uint64_t* long_keys, int32_t n_long,
                         void* workspace, size_t workspace_bytes, int32_t* status_host, void* stream);
int ezr_bm25_build_fill(const int64_t* doc_ptr, int64_t n_docs, int64_t n_tokens, int32_t vocab, const int64_t* indptr,
                        int32_t* out_post_doc, int32_t* out_post_tf, void* workspace, size_t workspace_bytes,
                        void* stream);
int ezr_bm25_shard_count(const int64_t* indptr, const int32_t* post_doc, int32_t vocab, int32_t doc_lo, int32_t doc_hi,
                         int64_t* out_first, int64_t* out_df_local, int64_t* out_indptr_local, void* stream);
int ezr_bm25_shard_copy(const int64_t* first, const int64_t* indptr_local, const int32_t* post_doc,
                        const int32_t* post_tf, int32_t vocab, int32_t doc_lo, int32_t* out_post_doc,
                        int32_t* out_post_tf, void* stream);

/* Packed postings for the candidate pass of ezr_bm25_topk: out_pk[p] = (post_doc[p] mod range_size) << W |
 * ceil(post_w[p] * 2^e), W = 32 - log2(range_size), e chosen from the largest weight so that every field fits
 * (returned in *out_scale_log2).  Rounding up makes the integer sums upper bounds of the float64 scores; the
 * exact scores of the surviving candidates are recomputed from post_w in token order, so results stay
 * bit-identical to rank_bm25 (retrievers.py:128-151).  Needs non-negative finite weights (else EZR_ERR_INVALID).
 * scratch16: 16 bytes of device memory.  Synchronises the stream (index-build time). */
int ezr_bm25_pack(const int32_t* post_doc, const double* post_w, int64_t n_postings, int32_t range_size,
                  uint32_t* out_pk, int32_t* out_scale_log2, void* scratch16, void* stream);

/* out_term_max[t] = largest packed weight among term t's postings (0 for an empty list). */
int ezr_bm25_term_max(const int64_t* indptr, const uint32_t* post_pk, int32_t vocab, uint32_t* out_term_max,
                      void* stream);

/* 1: the candidate pass of ezr_bm25_topk may skip non-essential terms when the index carries term_max;
 * 0 (default): it reads every posting.  Results are identical either way; on the measured workload the extra
 * candidates cost more rescoring time than the skipped postings save (profiles/README.md), hence the default. */
int ezr_bm25_set_skipping(int32_t on);

/* 1 (default): every candidate launch is preceded by a plan kernel that resolves token -> posting segment for all its
 * (query, range) pairs in parallel; 0: the candidate CTAs walk that chain of dependent loads themselves (A/B switch;
 * results are identical). */
int ezr_bm25_set_plan(int32_t on);

/* Document ranges of the FIRST candidate launch (default 4; then the same number again, then doubling up to 32).  A
 * tuning switch: fewer, larger launches on short shards; results are identical for every value. */
int ezr_bm25_set_span(int32_t first_ranges);

/* candidates per query the two-phase path can hold before it hands a query to the ordered kernel (0: not built) */
int ezr_bm25_cand_capacity(void);

/* BM25Retriever.get_scores + .filter for a batch of queries (retrievers.py:128-151,191-210):
 * query i has terms q_terms[q_ptr[i] .. q_ptr[i+1]) in token order (duplicates repeat, <0 or >=vocab = unknown).
 * Only documents with score > 0 qualify (retrievers.py:195-196); q_group[i] >= 0 additionally requires
 * doc_group[d] == q_group[i] (filter_dict, retrievers.py:198-202); -1 = no filter.
 * Outputs: out_scores[Q*k] (double/float per index->score_type), out_ids[Q*k] (-1 padded), out_counts[Q].
 * k <= 32 runs fused (accumulators never leave shared memory); larger k (<=1024) goes through a score row. */
size_t ezr_bm25_topk_workspace(const ezr_bm25_index* index, int32_t n_queries, int32_t k);
int ezr_bm25_topk(const ezr_bm25_index* index, const int32_t* q_ptr, const int32_t* q_terms, int32_t n_queries,
                  int32_t k, const int32_t* q_group, int32_t id_base, void* out_scores, int32_t* out_ids,
                  int32_t* out_counts, void* workspace, size_t workspace_bytes, void* stream);

/* BM25Retriever.get_scores (retrievers.py:128-151): full score rows, out_scores[Q * n_docs] */
int ezr_bm25_scores(const ezr_bm25_index* index, const int32_t* q_ptr, const int32_t* q_terms,
                    int32_t n_queries, void* out_scores, void* stream);

/* ------------------------------------------------------- generic top-k --
 * Row-wise top-k of a score matrix (k <= 1024): scores[q*row_stride + j], j < n_cols.
 * positive_only != 0 keeps only scores > 0 (BM25Retriever.filter). */
size_t ezr_select_rows_workspace(int32_t n_rows, int64_t n_cols, int32_t k, int32_t score_type);
int ezr_select_rows(const void* scores, int32_t score_type, int32_t n_rows, int64_t n_cols, int64_t row_stride,
                    int32_t k, int32_t positive_only, const int32_t* doc_group, const int32_t* q_group,
                    int32_t id_base, void* out_scores, int32_t* out_ids, int32_t* out_counts, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Merge per-shard / per-partition candidate lists: row q has n_cand (score,id) pairs at q*cand_stride,
 * id < 0 = empty slot.  Used after the all-gather of per-shard top-k. k <= 1024. */
size_t ezr_merge_topk_workspace(int32_t n_rows, int32_t n_cand, int32_t k, int32_t score_type);
int ezr_merge_topk(const void* cand_scores, const int32_t* cand_ids, int32_t score_type, int32_t n_rows,
                   int32_t n_cand, int64_t cand_stride, int32_t k, void* out_scores, int32_t* out_ids,
                   int32_t* out_counts, void* workspace, size_t workspace_bytes, void* stream);

/* Same merge over candidates that sit in n_parts separate segments: segment p of row q starts at
 * (char*)cand_x + p*part_stride_bytes + q*cand_stride*sizeof(elem) and holds n_cand entries.  This is the layout of
 * the all-gathered per-shard records (easyrag_b200/dist.py: one byte record per rank, gathered back to back), so
 * the shard merge reads the NCCL output in place -- no unpack / transpose kernels.  k <= 32. */
int ezr_merge_topk_parts(const void* cand_scores, const int32_t* cand_ids, int32_t score_type, int32_t n_rows,
                         int32_t n_cand, int64_t cand_stride, int32_t n_parts, int64_t part_stride_bytes, int32_t k,
                         void* out_scores, int32_t* out_ids, int32_t* out_counts, void* stream);

/* ------------------------------------------------------------- dense ----
 * QdrantRetriever (retrievers.py:37-52) over a COSINE collection (ingestion.py:180-182):
 * corpus rows and queries are L2-normalised bf16; score = fp32-accumulated dot product.
 * ld_* are row strides in elements.  q_group / doc_group implement the `dir` payload filter
 * (ingestion.py:207-216).  Rows short of k are padded with id -1, score -inf. */
size_t ezr_dense_topk_workspace(int64_t n_rows, int32_t dim, int32_t n_queries, int32_t k);
int ezr_dense_topk(const void* corpus_bf16, int64_t n_rows, int32_t dim, int64_t ld_corpus,
                   const void* queries_bf16, int32_t n_queries, int64_t ld_queries, int32_t k,
                   const int32_t* doc_group, const int32_t* q_group, int32_t id_base, float* out_scores,
                   int32_t* out_ids, int32_t* out_counts, void* workspace, size_t workspace_bytes,
                   void* stream);
/* Insert path of the in-HBM vector store: out[r] = bf16(x[r] / max(||x[r]||, 1e-12)) in fp32 math (what a
 * Distance.COSINE collection does at insert, ingestion.py:180-182).  x is float32 (x_is_f32 != 0) or bf16; strides
 * in elements; out may be a slice of a larger preallocated corpus matrix (append without rebuilding). */
int ezr_normalize_rows(const void* x, int32_t x_is_f32, int64_t ldx, int64_t n_rows, int32_t dim, void* out_bf16,
                       int64_t ldo, void* stream);
/* 0 = pick automatically, 1 = force the generic SIMT kernel, 2 = force tcgen05 with the query block in shared
 * memory (SS), 3 = force tcgen05 with the query block in tensor memory (TS) and 64-row corpus tiles, 4 = TS with
 * 128-row corpus tiles (the automatic choice), 5 = 4 run in cluster pairs (two neighbouring query blocks on the same
 * corpus split; each CTA loads half of every corpus tile and TMA-multicasts it to both); 2-5 error if the shape is
 * unsupported */
int ezr_dense_set_kernel(int32_t which);
/* name of the kernel the last ezr_dense_topk call on this thread launched
 * ("tcgen05" / "tcgen05-ts" / "tcgen05-ts128" / "tcgen05-ts128-mc2" / "simt") */
const char* ezr_dense_last_kernel(void);

/* Cap the TMA ring of the tcgen05 kernels at `stages` stages (0 = use all shared memory, the default).  A capped
 * ring leaves shared memory on every SM for kernels of another stream: CoarseRanker(overlap=True) uses it to let
 * BM25 CTAs co-reside with the persistent dense CTA (tensor pipe vs. integer/issue bound work). */
int ezr_dense_set_stage_cap(int32_t stages);

/* Measurement probes for the tcgen05 TS kernel (results become garbage; never use outside bench experiments):
 * bit mask: 1 = pipeline without TMA loads, 2 = one k-chunk of MMAs per tile, 4 = epilogue without the
 * insertion path. */
int ezr_dense_set_probe(int32_t probe);

/* ------------------------------------------------------------- fusion ---
 * HybridRetriever.reciprocal_rank_fusion (retrievers.py:256-274): list a first, then list b
 * (the reference passes [sparse, dense], retrievers.py:290); score += 1/(rank+K), rank from 1, fp64;
 * key = canon[id] (documents with identical text share a key, retrievers.py:263-265; NULL = identity);
 * the returned id is the LAST occurrence of the key (text_to_node overwrite, :264); ties keep
 * first-insertion order (stable sort, :266).  ids_x is [Q][stride_in], cnt_x[Q] valid entries each. */
int ezr_rrf_fuse(const int32_t* ids_a, const int32_t* cnt_a, const int32_t* ids_b, const int32_t* cnt_b,
                 int32_t n_queries, int32_t stride_in, const int32_t* canon, int32_t canon_base, int32_t K,
                 int32_t k_out, int32_t* out_ids, double* out_scores, int32_t* out_counts, void* stream);

/* HybridRetriever.fusion (retrievers.py:239-253): concatenate, drop later items whose text was seen,
 * stable sort by raw score descending, keep k_out. */
int ezr_fusion_simple(const int32_t* ids_a, const double* scores_a, const int32_t* cnt_a, const int32_t* ids_b,
                      const double* scores_b, const int32_t* cnt_b, int32_t n_queries, int32_t stride_in,
                      const int32_t* canon, int32_t canon_base, int32_t k_out, int32_t* out_ids,
                      double* out_scores, int32_t* out_counts, void* stream);

/* Both fusions over ANY number of rank lists (the reference's loops take a list of lists, retrievers.py:243,261;
 * the pipeline passes two).  ids_host / scores_host / cnt_host are HOST arrays of n_lists DEVICE pointers (each list
 * laid out like ids_a / scores_a / cnt_a above; scores_host may be NULL for RRF); list order = insertion order.
 * rrf != 0: reciprocal_rank_fusion, else fusion.  n_lists <= 8, n_lists * stride_in <= 2048. */
int ezr_fuse_lists(int32_t rrf, int32_t n_lists, const int32_t* const* ids_host, const double* const* scores_host,
                   const int32_t* const* cnt_host, int32_t n_queries, int32_t stride_in, const int32_t* canon,
                   int32_t canon_base, int32_t K, int32_t k_out, int32_t* out_ids, double* out_scores,
                   int32_t* out_counts, void* stream);

/* ------------------------------------------------- reranker hand-off ---
 * The coarse ranker's fused top-k -> the token sequences LLMRerank scores (rerankers.py:196-293 get_inputs /
 * get_inputs_v2_5, sliced 32 at a time by _postprocess_nodes :309-322), built on the device.  Pair p = q*k + r is
 * candidate r of query q: [bos] + query[: 3/4 max_length] + sep + passage (the pair truncated to max_length, the
 * passage gives way) + sep + prompt.  Queries ("A: ..." ids, q_ptr/q_tok) and passages ("B: ..." ids of every chunk,
 * tokenised once at index time, p_ptr/p_tok) are device CSR arrays.  Output is packed: ids[T], cu[P+1]; pairs past a
 * query's count are empty.  plan: lengths, their scan (int64) and get_inputs_v2_5's query_lengths; *total_host = T
 * (synchronises).  fill: the tokens + an int32 copy of cu (the cu_seqlens format of the encoder kernels). */
int ezr_rerank_pack_plan(const int32_t* cand_ids, const int32_t* cand_cnt, int32_t n_queries, int32_t k, int32_t k_stride,
                         int32_t id_base, const int32_t* q_ptr, const int64_t* p_ptr, int32_t n_sep, int32_t n_prompt,
                         int32_t max_length, int64_t* out_len, int64_t* out_cu, int32_t* out_query_len,
                         int64_t* total_host, void* stream);
int ezr_rerank_pack_fill(const int32_t* cand_ids, const int32_t* cand_cnt, int32_t n_queries, int32_t k, int32_t k_stride,
                         int32_t id_base, const int32_t* q_ptr, const int32_t* q_tok, const int64_t* p_ptr,
                         const int32_t* p_tok, const int32_t* sep, int32_t n_sep, const int32_t* prompt, int32_t n_prompt,
                         int32_t bos, int32_t max_length, const int64_t* cu, int32_t* out_ids, int32_t* out_cu32,
                         void* stream);

/* ------------------------------------------------------------ encoder ---
 * Building blocks of the chunk/query embedding forward pass (GTEEmbedding._embed, gte_embeddings.py:59-72 ->
 * Qwen2Model.forward, modeling_qwen.py:956-1116; HuggingFaceEmbedding._embed, hf_embeddings.py:112-123 ->
 * a BERT-shaped encoder).  Activations are bf16, row-major, PACKED: sequence b owns rows
 * [cu_seqlens[b], cu_seqlens[b+1]) -- no padding tokens.  The Python classes in easyrag_b200/encoder.py chain
 * these per layer on one stream. */

/* out[M,N'] = epi(A[M,K] . W[N,K]^T + bias) (+ residual); tcgen05/TMEM/TMA.  epilogue: 0 none, 1 GELU(erf),
 * 2 SwiGLU (W rows interleaved per 256: 128 gate rows then the matching 128 up rows; N' = N/2).  K % 64 == 0. */
int ezr_gemm_bf16(const void* a, int32_t m, int32_t k, int64_t lda, const void* w, int32_t n, int64_t ldw,
                  const void* bias, const void* residual, int64_t ldr, void* out, int64_t ldo, int32_t epilogue,
                  void* stream);
/* non-causal attention over packed q|k|v rows ([n_tokens, (H + 2*KV) * hd], row stride ld); head_dim 64 or 128; GQA
 * via n_kv_heads.  Default kernel: tcgen05 (S = QK^T and O += PV on the 5th-gen tensor cores, S/P/O in tensor memory,
 * Q/K/V tiles by TMA).  n_tokens bounds the TMA tensor map (tiles that run past the last token are zero-filled). */
int ezr_attn_bidir(const void* qkv, int64_t n_tokens, int64_t ld, const int32_t* cu_seqlens, int32_t n_seq,
                   int32_t max_len, int32_t n_heads, int32_t n_kv_heads, int32_t head_dim, float softmax_scale, void* out,
                   int64_t ldo, void* stream);
/* 0 = tcgen05 kernel (default), 1 = the warp-level mma.sync kernel it replaced (kept as an independent cross-check) */
int ezr_attn_set_kernel(int32_t which);
/* "tcgen05" / "mma.sync": what the last ezr_attn_bidir call on this thread launched */
const char* ezr_attn_last_kernel(void);
int ezr_embed_gather(const int32_t* ids, int32_t n_tokens, const void* table, int64_t ldt, int32_t vocab, int32_t dim,
                     void* out, int64_t ldo, void* stream);
/* BERT embeddings: LayerNorm(word[id] + type[0] + position[pos]) */
int ezr_bert_embed(const int32_t* ids, const int32_t* positions, int32_t n_tokens, const void* word, const void* pos,
                   const void* type0, const void* gamma, const void* beta, float eps, int32_t vocab, int32_t max_pos,
                   int32_t dim, void* out, void* stream);
/* Qwen2RMSNorm (modeling_qwen.py:91-96) */
int ezr_rmsnorm(const void* x, int64_t ldx, const void* gamma, float eps, int32_t n_rows, int32_t dim, void* out,
                int64_t ldo, void* stream);
int ezr_layernorm(const void* x, int64_t ldx, const void* gamma, const void* beta, float eps, int32_t n_rows,
                  int32_t dim, void* out, int64_t ldo, void* stream);
/* rotary embedding in place on the q and k heads of packed qkv rows (modeling_qwen.py:137-169); bf16 cos/sin tables
 * [max_pos, head_dim/2] */
int ezr_rope(void* qkv, int64_t ld, const int32_t* positions, const void* cos_table, const void* sin_table,
             int32_t max_pos, int32_t n_heads_qk, int32_t head_dim, int32_t n_tokens, void* stream);
/* pooling (0 last token, 1 first/CLS, 2 mean) + optional final RMSNorm of the pooled row + L2 normalisation
 * (l2_mode 0 none, 1 bf16 semantics of gte_embeddings.py:70, 2 fp32 semantics); writes bf16 [n_seq, dim] and,
 * if out_f32 != NULL, the float copy the embedding API returns */
int ezr_pool_normalize(const void* hidden, int64_t ldh, const int32_t* cu_seqlens, int32_t n_seq, int32_t pool,
                       int32_t final_norm, const void* gamma, float eps, int32_t l2_mode, int32_t dim, void* out_bf16,
                       float* out_f32, void* stream);

/* ----------------------------------------------------------- profiling ---
 * Per-kernel CUDA-event timing on the launching stream, for bench.py's roofline figures.
 * ezr_profile_read synchronises on the recorded events and returns the summed duration of the
 * kernel launches recorded in `slot` since the last reset. */
typedef enum ezr_prof_slot {
    EZR_PROF_BM25_SCORE = 0, /* bm25_score_kernel (fused top-k or score rows) */
    EZR_PROF_DENSE_TC = 1,   /* dense_tc_kernel (tcgen05) */
    EZR_PROF_DENSE_SIMT = 2, /* dense_scores_simt_kernel */
    EZR_PROF_MERGE = 3,      /* merge / select kernels */
    EZR_PROF_FUSE = 4,       /* rrf / simple fusion */
    EZR_PROF_ENC_GEMM = 5,   /* encoder GEMMs */
    EZR_PROF_ENC_ATTN = 6,   /* encoder attention */
    EZR_PROF_ENC_OTHER = 7,  /* encoder norms / elementwise */
    EZR_PROF_BM25_CAND = 8,  /* bm25_cand_kernel (integer candidate pass over packed postings) */
    EZR_PROF_BM25_RESCORE = 9, /* bm25_rescore_kernel (exact float64 rescoring + top-k of the candidates) */
    EZR_PROF_COUNT = 10
} ezr_prof_slot;
/* kernels launched by this library since it was loaded (every launch site counts itself) */
long long ezr_launch_count(void);
int ezr_profile_enable(int32_t on);
int ezr_profile_reset(void);
int ezr_profile_read(int32_t slot, double* total_ms, int32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* EASYRAG_B200_H */
