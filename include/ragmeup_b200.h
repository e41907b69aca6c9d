/*
 * ragmeup_b200 — C ABI of the B200-native dense-retrieval hot path.
 *
 * The reference (AI-Commandos/RAGMeUp) is pure Python and reaches this path through
 * three LangChain objects owned by RAGHelper:
 *   self.embeddings  HuggingFaceEmbeddings      server/RAGHelper_local.py:114-117, RAGHelper_cloud.py:101-103
 *   self.db          Milvus / PGVector          server/RAGHelper.py:388-394, 399-404, 431, 497-499, 525
 *   self.compressor  ScoredCrossEncoderReranker(HuggingFaceCrossEncoder)
 *                                               server/RAGHelper.py:483-486, server/ScoredCrossEncoderReranker.py:25-45
 * There is no FFI in the reference; the binding a maintainer adds is the ctypes stub in
 * INTEGRATION.md (ragmeup_b200/_lib.py is that stub).  Each entry point below names the
 * reference interface it stands behind.
 *
 * Conventions: plain C, opaque handles, int return codes (0 = ok, <0 = error, text from
 * rmu_last_error()), every pointer is a DEVICE pointer unless its name ends in _h, every
 * asynchronous call takes the cudaStream_t to run on (as void*), no hidden device syncs unless
 * stated.  Handles own the corpus / weights and their scratch memory; callers own inputs and
 * outputs.  All calls on one handle are serialised internally (Flask worker threads share one
 * RAGHelper: server/server.py:141-146,394).
 */
#ifndef RAGMEUP_B200_H
#define RAGMEUP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMU_OK 0
#define RMU_ERR_ARG -1
#define RMU_ERR_CUDA -2
#define RMU_ERR_UNSUPPORTED -3

/* metric ids — Milvus FLAT/L2 is the reference default store metric (langchain-milvus 0.1.3
 * via server/RAGHelper.py:388-394); PGVector uses cosine distance (server/RAGHelper.py:399-404). */
#define RMU_METRIC_IP 0      /* score = <q,x>            larger is better  */
#define RMU_METRIC_COSINE 1  /* score = cos(q,x)         larger is better  */
#define RMU_METRIC_L2 2      /* score = ||q-x||^2        smaller is better */

/* search modes */
#define RMU_SEARCH_AUTO 0    /* tcgen05 coarse scan + exact fp32 re-score + certified fallback */
#define RMU_SEARCH_EXACT 1   /* fp32 CUDA-core scan only (what AUTO falls back to)            */
#define RMU_SEARCH_TENSOR_NOFALLBACK 2 /* diagnostics: coarse+re-score, flags reported, no fallback */

const char* rmu_last_error(void);
int rmu_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t rmu_launch_count(void);

/* per-kernel-class device timing (CUDA events on the launching stream; used by bench.py for the
 * roofline objects).  classes: 0 scan, 1 finalize, 2 exact scan, 3 merge, 4 gemm, 5 attention,
 * 6 layernorm, 7 embedding, 8 pool/head, 9 misc, 10 second-pass scan launches (queries whose certificate failed; exit at once when there are none).  rmu_profile_read synchronises the recorded events. */
void rmu_profile_enable(int on);
void rmu_profile_reset(void);
int rmu_profile_read(int cls, double* total_ms, int64_t* launches);

/* ------------------------------------------------------------------ flat vector index
 * Stands behind the vector store: Milvus.from_documents / PGVector ctor (RAGHelper.py:385-404),
 * db.add_documents (:431,525) and the retriever's col.search (:497-499). */
typedef struct rmu_index rmu_index;

int rmu_index_create(int dim, int metric, rmu_index** out);
void rmu_index_destroy(rmu_index* idx);
int rmu_index_reserve(rmu_index* idx, int64_t rows);
/* append n rows of `dim` fp32; src_is_host selects a host (pinned or pageable) source */
int rmu_index_add(rmu_index* idx, const float* vecs, int64_t n, int src_is_host, void* stream);
/* overwrite existing rows in place (PGVector's upsert-on-id, langchain-postgres add_embeddings): rows[n] int64 local row
 * numbers and vecs [n, dim] fp32, both DEVICE pointers; rows outside [0, size) are ignored */
int rmu_index_set_rows(rmu_index* idx, const int64_t* rows, const float* vecs, int64_t n, void* stream);
int64_t rmu_index_size(const rmu_index* idx);
int rmu_index_dim(const rmu_index* idx);
int rmu_index_metric(const rmu_index* idx);
int rmu_index_clear(rmu_index* idx);
/* raw device pointer to the [size, dim] fp32 corpus (persistence / tests) */
const float* rmu_index_data(const rmu_index* idx);

/* top-k of every query against the whole index.
 *   queries [nq, dim] fp32, out_scores [nq, k] fp32, out_ids [nq, k] int64 = id_offset + row,
 *   missing results (k > size) are id -1 with score -inf (+inf for L2).
 *   stats_h (nullable, host int32[4]): {queries whose certificate failed in the first tensor pass, first-pass
 *   tensor-scan launches, queries that still needed the exact fp32 CUDA-core scan after the second pass, 0};
 *   requesting it makes the call synchronise the stream. */
int rmu_index_search(rmu_index* idx, const float* queries, int nq, int k, int64_t id_offset, int mode,
                     float* out_scores, int64_t* out_ids, int32_t* stats_h, void* stream);
/* same, HOST buffers in and out (H2D + D2H inside the call, synchronises `stream`) */
int rmu_index_search_host(rmu_index* idx, const float* queries_h, int nq, int k, int64_t id_offset, int mode,
                          float* out_scores_h, int64_t* out_ids_h, void* stream);
/* diagnostics: raw TF32 accumulators of the first 256 rows (one CTA-pair tile) for nq <= 64 queries, out [64, 256] */
int rmu_debug_scan_tile(rmu_index* idx, const float* queries, int nq, float* out, void* stream);
/* rows[n] (int64, local row numbers) -> out [n, dim]; feeds MMR (langchain-milvus fetches the
 * fetch_k stored vectors the same way after col.search) */
int rmu_index_gather(rmu_index* idx, const int64_t* rows, int n, float* out, void* stream);

/* merge R per-shard result lists (after the NCCL all-gather): [R, nq, k] -> [nq, k] */
int rmu_topk_merge(const float* scores, const int64_t* ids, int R, int nq, int k, int metric,
                   float* out_scores, int64_t* out_ids, void* stream);

/* the same on the buffer ONE all-gather fills: rank r's scores start at scores + r * rank_stride_scores (elements),
 * its ids at ids + r * rank_stride_ids, so a per-rank record {scores[nq*k], ids[nq*k]} needs no repacking */
int rmu_topk_merge_strided(const float* scores, const int64_t* ids, int64_t rank_stride_scores, int64_t rank_stride_ids,
                           int R, int nq, int k, int metric, float* out_scores, int64_t* out_ids, void* stream);

/* greedy maximal-marginal-relevance re-selection (langchain `maximal_marginal_relevance`,
 * used by as_retriever(search_type="mmr"), server/RAGHelper.py:497-499,533-535):
 * q [nq, dim], cand [nq, fetch_k, dim], n_cand[nq] valid candidates, out_sel [nq, k] int32
 * positions into the candidate list (-1 padded). */
int rmu_mmr_select(const float* q, const float* cand, const int32_t* n_cand, int nq, int fetch_k, int dim,
                   int k, float lambda_mult, int32_t* out_sel, void* stream);

/* cosine DISTANCE (1 - cosine similarity) between consecutive rows of x [n, dim] -> out [n - 1] (device, float64):
 * the sentence-to-sentence distances of langchain_experimental's SemanticChunker, which the reference builds at
 * server/RAGHelper.py:329-341 and runs through text_splitter.split_documents (:368). */
int rmu_adjacent_cosine_distance(const float* x, int64_t n, int dim, double* out, void* stream);

/* ------------------------------------------------------------------ BM25 sparse leg
 * Stands behind langchain_community.retrievers.BM25Retriever (rank_bm25.BM25Okapi.get_scores /
 * get_top_n), built at server/RAGHelper.py:436-443 and queried through the EnsembleRetriever
 * (:501-503).  The inverted index is CSR by term: postings of term t are
 * post_doc/post_tf[post_ptr[t] .. post_ptr[t+1]) with documents ascending; den[d] =
 * k1 * (1 - b + b * doc_len[d] / avgdl) and idf[t] are float64, computed by the host exactly as
 * rank_bm25 does.  Scores are float64 and bit-identical to rank_bm25's; results are ordered score
 * descending, equal scores by larger document index (numpy stable argsort reversed). */
typedef struct rmu_bm25 rmu_bm25;
int rmu_bm25_create(int64_t n_docs, int64_t n_terms, const int64_t* post_ptr_h, const int32_t* post_doc_h,
                    const int32_t* post_tf_h, const double* den_h, const double* idf_h, double k1_plus_1,
                    rmu_bm25** out);                    /* host arrays, copied to the device */
int rmu_bm25_destroy(rmu_bm25* h);
int64_t rmu_bm25_size(const rmu_bm25* h);
int64_t rmu_bm25_terms(const rmu_bm25* h);
/* Q queries: terms of query i are q_terms[q_ptr[i] .. q_ptr[i+1]) (term ids in query order, repeats
 * kept); out_scores [Q, k] float64, out_ids [Q, k] int64 document numbers, -1 padded when the corpus
 * has fewer than k documents.  1 <= k <= 256.  Device pointers: */
int rmu_bm25_search(rmu_bm25* h, const int32_t* q_ptr, const int32_t* q_terms, int Q, int k, double* out_scores,
                    int64_t* out_ids, void* stream);
/* the same with HOST buffers (copies + synchronise inside the call) */
int rmu_bm25_search_host(rmu_bm25* h, const int32_t* q_ptr_h, const int32_t* q_terms_h, int Q, int k,
                         double* out_scores_h, int64_t* out_ids_h, void* stream);

/* Host-only helper (no GPU work): the inverted index rank_bm25 would build from
 * `[text.split() for text in texts]` (BM25Retriever.from_texts, server/RAGHelper.py:436-443) — Python
 * str.split() whitespace semantics on UTF-8, vocabulary ids in first-seen order, postings CSR by term
 * with documents ascending.  texts_utf8[i] has text_bytes[i] bytes (no terminator needed). */
typedef struct rmu_bm25_csr rmu_bm25_csr;
int rmu_bm25_csr_build(const char* const* texts_utf8, const int64_t* text_bytes, int64_t n_docs, rmu_bm25_csr** out);
int rmu_bm25_csr_sizes(const rmu_bm25_csr* c, int64_t* n_terms, int64_t* nnz, int64_t* vocab_bytes);
/* doc_len [n_docs], post_ptr [n_terms+1], post_doc/post_tf [nnz], vocab_off [n_terms+1], vocab_bytes */
int rmu_bm25_csr_export(const rmu_bm25_csr* c, int64_t* doc_len, int64_t* post_ptr, int32_t* post_doc,
                        int32_t* post_tf, int64_t* vocab_off, char* vocab_bytes);
int rmu_bm25_csr_free(rmu_bm25_csr* c);

/* ------------------------------------------------------------------ BERT encoder
 * Stands behind HuggingFaceEmbeddings.embed_documents/embed_query (sentence-transformers
 * encode -> BertModel -> Pooling -> Normalize) and HuggingFaceCrossEncoder.score
 * (CrossEncoder.predict -> BertForSequenceClassification). */
typedef struct rmu_bert_config {
    int32_t vocab_size, hidden, layers, heads, ffn, max_pos, type_vocab, num_labels;
    float ln_eps;
} rmu_bert_config;

typedef struct rmu_encoder rmu_encoder;

/* weights_h: host fp32 arrays in the canonical order documented in ragmeup_b200/encoder.py
 * (HuggingFace BertModel tensor order; pooler + classifier last when has_head). */
int rmu_encoder_create(const rmu_bert_config* cfg, const float* const* weights_h, int n_weights, int has_head,
                       rmu_encoder** out);
void rmu_encoder_destroy(rmu_encoder* enc);

#define RMU_POOL_MEAN 0
#define RMU_POOL_CLS 1

/* ragged batch: ids/type_ids [total_tokens] int32, cu_seqlens [B+1] int32 (all device).
 * out [B, hidden] fp32 sentence embeddings. */
int rmu_encoder_embed(rmu_encoder* enc, const int32_t* ids, const int32_t* type_ids, const int32_t* cu_seqlens,
                      int B, int total_tokens, int max_seqlen, int pool_mode, int normalize, float* out,
                      void* stream);
/* cross-encoder logits: out [B, num_labels] fp32 (raw logits; activation is the caller's) */
int rmu_encoder_classify(rmu_encoder* enc, const int32_t* ids, const int32_t* type_ids, const int32_t* cu_seqlens,
                         int B, int total_tokens, int max_seqlen, float* out, void* stream);
/* last hidden state [total_tokens, hidden] fp32 (tests / provenance) */
int rmu_encoder_hidden(rmu_encoder* enc, const int32_t* ids, const int32_t* type_ids, const int32_t* cu_seqlens,
                       int B, int total_tokens, int max_seqlen, float* out, void* stream);
/* host-buffer variants (H2D of the token batch + D2H of the result inside; synchronise) */
int rmu_encoder_embed_host(rmu_encoder* enc, const int32_t* ids_h, const int32_t* type_ids_h,
                           const int32_t* cu_seqlens_h, int B, int total_tokens, int max_seqlen, int pool_mode,
                           int normalize, float* out_h, void* stream);
int rmu_encoder_classify_host(rmu_encoder* enc, const int32_t* ids_h, const int32_t* type_ids_h,
                              const int32_t* cu_seqlens_h, int B, int total_tokens, int max_seqlen, float* out_h,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RAGMEUP_B200_H */
