/*
 * sdag_b200 — C ABI of the B200 (sm_100a) library behind the two hot paths of
 * sagie-dekel/Sparse-Document-Attention-RAG-.
 *
 * The reference is pure Python and has no FFI of its own; each entry point below names the
 * reference function (file:line under /root/reference) whose arithmetic it replaces.  The Python
 * package binds these with ctypes (see INTEGRATION.md for the stubs a maintainer of the reference
 * would add).
 *
 * Conventions: raw device pointers, explicit sizes / strides (in ELEMENTS unless a name says
 * bytes), a cudaStream_t passed as void*; return 0 on success, a cudaError_t value (< 1000) or one
 * of the SDAG_ERR_* codes otherwise.  No call allocates device memory, throws, or synchronises the
 * stream except where documented; workspaces are caller-owned and sized by the *_bytes queries.
 * All functions are re-entrant per stream.
 */
#ifndef SDAG_B200_H_
#define SDAG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDAG_OK 0
#define SDAG_ERR_INVALID 1001     /* bad argument (null pointer, negative size, misaligned stride) */
#define SDAG_ERR_UNSUPPORTED 1002 /* shape outside what the kernels implement (head_dim, nbr_k, ...) */
#define SDAG_ERR_WORKSPACE 1003   /* caller workspace too small */
#define SDAG_ERR_DRIVER 1004      /* tensor-map encoding unavailable / failed */

#define SDAG_TILE 128             /* query rows and key columns per attention tile */
#define SDAG_PREFILL_NO_UNIFORM_ROWS 1 /* sdag_prefill_fwd flag: no query row is without an allowed key */
#define SDAG_PREFILL_FORCE_SINGLE 2    /* sdag_prefill_fwd flag: single-CTA kernel even where the CTA-pair kernel applies */
#define SDAG_PREFILL_FORCE_PAIR 4      /* sdag_prefill_fwd flag: CTA-pair kernel (head_dim 128, H / Hkv even) or SDAG_ERR_UNSUPPORTED */
#define SDAG_BLOCKMAP_HEADER 16   /* int32 words */
#define SDAG_LAYOUT_HEADER 8      /* int32 words */
#define SDAG_MAX_NBR_K 16
#define SDAG_MAX_DOCS 512

int sdag_b200_version(void);
const char* sdag_b200_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * Path A — SDAG mask / block map.
 *
 * A prompt layout is an int32 record  [L, S, qa_start, n_docs, use_nbrs, 0, 0, 0,
 *   spans[max_docs][2], nbrs[max_docs][nbr_k] (-1 padded)]   (stride = sdag_layout_ints()).
 * It carries exactly the arguments of build_blocked_causal_mask_full (SDAG.py:68-76):
 * base_seq_len, sys_user_len, doc_token_spans, qa_start, doc_neighbors (use_nbrs = the
 * `len(doc_neighbors) == num_docs` test of SDAG.py:111).
 * ---------------------------------------------------------------------------------------------- */
int sdag_layout_ints(int max_docs, int nbr_k);

/* Dense bool mask [L, L] (1 byte per entry) — the 1:1 replacement of
 * build_blocked_causal_mask_full, SDAG.py:68-127, including its slice-clipping behaviour on odd
 * inputs.  Only used at that seam and by tests; the attention path never materialises it. */
int sdag_build_mask(const int32_t* layout_dev, int max_docs, int nbr_k, uint8_t* mask_dev, int L,
                    void* stream);

/* Block map computed from the document token offsets (replaces SDAG.py:68-127 + the bool->additive
 * conversion of SDAG.py:164-173 on the attention path).  One map per prompt:
 *   header[16] = {n_qtiles, n_kvtiles, status, n_gap_rows, L, cap_qtiles, cap_kvtiles,
 *                 bit_words_per_row, pairs_lo, pairs_hi, bits_offset_ints, ...}
 *   qtiles[cap_qtiles]   = {q0, nrows, kv_off, kv_cnt}   sorted by kv_cnt descending
 *   kvtiles[cap_kvtiles] = {k0, ncols, doff, flags}      allowed(r,c) = c < ncols && c <= r + doff
 * status: 0 ok; 1 layout not in canonical order (S <= s_0, spans sorted and disjoint,
 * e_last <= qa_start) -> caller must use sdag_blockmap_from_mask; 2 capacity exceeded. */
size_t sdag_blockmap_bytes(int L, int n_docs, int nbr_k);
int sdag_build_blockmap(const int32_t* layouts_dev, int n_prompts, int max_docs, int nbr_k, int max_L,
                        int32_t* blockmaps_dev, size_t blockmap_stride_bytes, void* stream);

/* General path: block map of an arbitrary dense bool mask [Lq = L, L] (what
 * generate_with_custom_mask accepts, SDAG.py:130-136).  Tiles are globally aligned; non-empty tiles
 * carry a bit-packed copy of the mask stored inside the map buffer. */
size_t sdag_maskmap_bytes(int L);
int sdag_blockmap_from_mask(const uint8_t* mask_dev, int64_t row_stride, int L, int32_t* blockmap_dev,
                            size_t blockmap_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Path A — attention.  Replaces the HF attention call the reference's mask reaches:
 * LlamaAttention.forward -> sdpa/eager  softmax(Q K^T * scaling + bias) V
 * (transformers models/llama/modeling_llama.py:199-221, integrations/sdpa_attention.py:40-104),
 * reached from SDAG.py:176 (prefill) and SDAG.py:193-197 (decode).
 * bf16 in / bf16 out, fp32 softmax and accumulation; GQA: q head h reads kv head h / (H / Hkv).
 * Rows with no allowed key produce mean_j V[j] over all Skv keys (uniform softmax over finfo.min
 * logits — SURVEY.md Appendix A2).
 * ---------------------------------------------------------------------------------------------- */
size_t sdag_prefill_workspace_bytes(int B, int Hkv, int D);

/* q [B,H,Lq,D] / k,v [B,Hkv,Skv,D] / out [B,Lq,H,D] addressed through element strides (innermost
 * stride 1, others multiples of 8).  blockmaps: one per batch entry (stride in bytes, 0 = shared).
 * D in {64, 128}.  flags: SDAG_PREFILL_NO_UNIFORM_ROWS skips the mean-of-V pass when the caller knows
 * that no query row is without an allowed key (plain causal and bidirectional maps). */
int sdag_prefill_fwd(const void* q, int64_t sq_b, int64_t sq_h, int64_t sq_l,
                     const void* k, int64_t sk_b, int64_t sk_h, int64_t sk_l,
                     const void* v, int64_t sv_b, int64_t sv_h, int64_t sv_l,
                     void* out, int64_t so_b, int64_t so_l, int64_t so_h,
                     int B, int H, int Hkv, int Lq, int Skv, int D, float scaling,
                     const int32_t* blockmaps_dev, size_t blockmap_stride_bytes, int cap_qtiles,
                     void* workspace, size_t workspace_bytes, int flags, void* stream);

/* One decode step, dense over the cache (SDAG.py:193-197 passes no mask): q [B,H,D] (stride sq_b,
 * sq_h), k/v cache [B,Hkv,S_cap,D], kv_len valid positions, out [B,H,D] contiguous.  When kv_len_dev is
 * not NULL the kernel reads the valid length from device memory (kv_len is then the upper bound used to
 * size the grid), so a captured CUDA graph can be replayed for every generated token; kv_len_dev_stride
 * = 1 reads one length per prompt (ragged batches, e.g. generate_batch_seq2seq,
 * attack/content_generation.py:44-94), 0 one value for the whole batch. */
size_t sdag_decode_workspace_bytes(int B, int H, int D, int kv_len);
int sdag_decode_fwd(const void* q, int64_t sq_b, int64_t sq_h,
                    const void* k, int64_t sk_b, int64_t sk_h, int64_t sk_l,
                    const void* v, int64_t sv_b, int64_t sv_h, int64_t sv_l,
                    void* out, int B, int H, int Hkv, int kv_len, int D, float scaling,
                    void* workspace, size_t workspace_bytes, const int32_t* kv_len_dev, int kv_len_dev_stride,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Path B — exact max-inner-product top-k.  Replaces faiss IndexFlatIP.search as called by
 * search_index, src/pipeline/retrieval/dense.py:68-84.
 *
 * Stage 1 (ip_topk_fwd): bf16 tcgen05 GEMM over the HBM-resident bf16 shadow of the corpus, fused
 * with a per-query threshold filter; survivors (bf16 score, row id) are appended to per-query
 * candidate lists.  Thresholds are provable lower bounds of the final k-th exact score minus the
 * bf16 error bound, so the true top-k is always among the candidates.
 * Stage 2 (ip_rescore_topk): exact fp32 re-scoring of the candidates against the fp32 corpus and
 * final ordering (score descending, id ascending) -> (D float32[Q,k], I int64[Q,k]), -1 / -FLT_MAX
 * padding when fewer than k rows exist (dense.py:118-125 relies on that).
 * ---------------------------------------------------------------------------------------------- */
/* fp32 rows -> bf16 shadow rows; also max row L2 norm (atomicMax on *max_norm_dev, caller zeroes). */
int ip_cast_rows_bf16(const float* x, void* x_bf16, float* max_norm_dev, int64_t n, int d, void* stream);

typedef struct {
  int cand_cap;     /* candidate slots per query */
  int sample_rows;  /* rows of the first (dump) epoch */
} ip_search_config_t;

size_t ip_search_workspace_bytes(int Q, int d, int k, int cand_cap);
/* Full single-shard search.  xq fp32 [Q,d] (device), corpus fp32 [N,d] + bf16 shadow [N,d],
 * id_offset added to row numbers (corpus sharding), max_norm = max corpus row norm (device scalar).
 * Outputs on device: scores float32 [Q,k], ids int64 [Q,k]; stats int32[4] =
 * {overflowed queries (re-run exactly inside), total candidates, epochs, reserved}.
 * With SDAG_IP_EARLY_RESCORE=1 (off by default: measured slower) the exact re-scoring of the candidates found before the
 * last epoch runs on an internal second stream under the last epoch's GEMM (forked from and joined back into `stream` with
 * events, so the call stays stream-ordered and capturable); the library keeps one such stream per device. */
int ip_search(const float* xq, int Q, const float* corpus_f32, const void* corpus_bf16, int64_t N, int d,
              int64_t id_offset, const float* max_norm_dev, int k, float* out_scores, int64_t* out_ids,
              int32_t* stats_dev, void* workspace, size_t workspace_bytes, int cand_cap, void* stream);

/* Merge G per-shard results [G,Q,k] (e.g. after ncclAllGather) into the global top-k, same order. */
int topk_merge(const float* scores, const int64_t* ids, int G, int Q, int k, float* out_scores,
               int64_t* out_ids, void* stream);
/* The exchange step of the sharded search with ONE buffer per rank: topk_pack writes the local result as
 * n = Q*k records of three int32 {score bits, id low word, id high word}; after one ncclAllGather of those records
 * (rank-major [G,Q,k,3]) topk_merge_packed merges them into the global top-k with topk_merge's order rule (score
 * descending, then id ascending, padding ids -1 last).  Replaces faiss' single-index heap over the whole corpus,
 * dense.py:83, when the corpus rows are partitioned over GPUs. */
/* The same exchange as ONE kernel over peer memory (NVLink / NVSwitch), no collective library call: every rank stores
 * its packed records straight into the receive area of every rank, raises a per-(source rank, block) flag there (release,
 * system scope), waits for the G flags of its own blocks and merges.  Every rank owns one buffer of
 * topk_exchange_buffer_bytes(G, Q_cap, k) bytes, ZERO-FILLED ONCE and mapped into all ranks (CUDA peer access / IPC /
 * symmetric memory); peer_buffers_dev is a device array of the G base addresses as THIS process sees them (entry [rank] = its
 * own buffer).  epoch >= 1 must increase by one per call, identically on every rank; all ranks must make the call (a rank
 * that never does lets the others trap after ~2 s instead of hanging).  Q <= Q_cap. */
size_t topk_exchange_buffer_bytes(int G, int Q_cap, int k);
int topk_exchange_merge(const float* scores, const int64_t* ids, int Q, int Q_cap, int k, int rank, int G,
                        const uint64_t* peer_buffers_dev, uint32_t epoch, float* out_scores, int64_t* out_ids, void* stream);
int topk_pack(const float* scores, const int64_t* ids, int64_t n, void* records, void* stream);
int topk_merge_packed(const void* records, int G, int Q, int k, float* out_scores, int64_t* out_ids, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Path B — selector.  Replaces the numeric core of _select_malicious_docs_for_query,
 * src/pipeline/attack/malicious_selection.py:81-128: centroid = mean of retrieved embeddings (not
 * re-normalised), sims = cand . centroid, order by sims (descending for near, ascending for far;
 * ties -> lower index first, as np.argmax / np.argmin).  Batched over queries:
 * retrieved [nq, R_max, d] with r_count[nq] valid rows, cand [nq, C_max, d] with c_count[nq].
 * Outputs: sims float32 [nq, C_max], order int32 [nq, C_max] (-1 past c_count).
 * ---------------------------------------------------------------------------------------------- */
int centroid_select(const float* retrieved, const int32_t* r_count, int R_max, const float* cand,
                    const int32_t* c_count, int C_max, int nq, int d, int furthest, float* sims_out,
                    int32_t* order_out, void* stream);

/* Directed kNN among the k prompt documents (compute_doc_knn_for_docs, SDAG.py:50-63): sims = E E^T
 * fp32, per row the k_neighbors most similar other rows, most similar first. emb [n,d] -> nbrs
 * int32 [n, k_neighbors] (-1 when fewer exist). */
int doc_knn(const float* emb, int n, int d, int k_neighbors, int32_t* nbrs_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Path A4 — element-wise glue of the decoder layer around the attention call (the prefill forward
 * generate_with_custom_mask drives, SDAG.py:175-176 -> HF LlamaDecoderLayer.forward,
 * models/llama/modeling_llama.py).  bf16 tensors, HF module rounding order.
 * ---------------------------------------------------------------------------------------------- */
/* LlamaRMSNorm.forward; with `residual` != NULL first s = x + residual (written to sum_out), then
 * out = weight * bf16(s * rsqrt(mean(s^2) + eps)).  rows x dim, dim % 8 == 0. */
int llm_rmsnorm(const void* x, const void* residual, const void* weight, void* sum_out, void* out,
                int64_t rows, int dim, float eps, void* stream);
/* apply_rotary_pos_emb + KV-cache write: qkv [B*L, (H + 2*Hkv)*D] -> q_out [B,L,H,D], k_out / v_out
 * [B,Hkv,S_cap,D] at positions pos0..pos0+L-1 (element strides kv_stride_b / kv_stride_h);
 * cos / sin: bf16 tables [>= pos0+L, D] indexed by absolute position.  pos0_dev != NULL: read pos0
 * from device memory (CUDA-graph replay), one value per prompt when pos0_dev_stride = 1 (ragged
 * batches), one for the batch when 0. */
int llm_rope_qkv(const void* qkv, const void* cos, const void* sin, void* q_out, void* k_out, void* v_out,
                 int B, int L, int H, int Hkv, int D, int64_t kv_stride_b, int64_t kv_stride_h, int pos0,
                 const int32_t* pos0_dev, int pos0_dev_stride, void* stream);
/* LlamaMLP activation: gate_up [rows, 2*inter] (gate first) -> out [rows, inter] = silu(gate) * up. */
int llm_swiglu(const void* gate_up, void* out, int64_t rows, int inter, void* stream);
/* Prefill linear layer (the q/k/v, o, gate/up and down projections HF runs inside the prefill forward,
 * SDAG.py:175-176 -> LlamaAttention / LlamaMLP): out [M, N] = x [M, K] . weight [N, K]^T (+ bias), bf16 with fp32
 * accumulation on CTA pairs (tcgen05 cta_group::2).  x rows ldx elements apart, out rows ldo apart; K, ldx,
 * ldo, N multiples of 8, 16-byte aligned pointers.  epilogue LLM_GEMM_SWIGLU: weight is the fused [gate; up]
 * matrix (N = 2*I rows, no bias) and out [M, I] = silu(gate) * up with the LlamaMLP bf16 roundings — the
 * [M, 2*I] intermediate is never written. */
#define LLM_GEMM_PLAIN 0
#define LLM_GEMM_SWIGLU 1
#define LLM_GEMM_GELU 2 /* out = gelu(x W^T + bias), erf form, applied to the bf16-rounded linear output (BertIntermediate) */
int llm_gemm(const void* x, int64_t ldx, const void* weight, const void* bias, void* out, int64_t ldo, int64_t M,
             int N, int K, int epilogue, void* stream);
/* The same GEMM with the residual add of the decoder / encoder layer in its epilogue:
 * out = bf16(bf16(x W^T + bias) + residual) — `hidden_states = residual + self.o_proj(...)` and
 * `residual + self.mlp(...)` of LlamaDecoderLayer.forward (HF modeling_llama.py, reached from SDAG.py:175-176), and the
 * `dense(x) + input_tensor` of BertSelfOutput / BertOutput in the encoder.  residual [M, N] rows ldr elements apart
 * (ldr % 8 == 0, 16-byte aligned); it may alias `out` (every element is read and written by the same thread). */
int llm_gemm_residual(const void* x, int64_t ldx, const void* weight, const void* bias, const void* residual,
                      int64_t ldr, void* out, int64_t ldo, int64_t M, int N, int K, void* stream);
/* The prefill's q/k/v projection with apply_rotary_pos_emb and the KV-cache write in the GEMM epilogue (llm_gemm on the
 * fused [(H + 2*Hkv)*D, K] q|k|v weight followed by llm_rope_qkv, in one launch; the [B*L, (H+2Hkv)*D] intermediate is
 * never written): x [B*L, K] rows ldx apart, token row b*L + l sits at position pos0 + l; q_out [B, L, H, D]; k_out /
 * v_out, cos / sin as in llm_rope_qkv.  D = 64 or 128, H and Hkv multiples of 256 / D (SDAG_ERR_UNSUPPORTED otherwise). */
int llm_gemm_rope(const void* x, int64_t ldx, const void* weight, const void* bias, const void* cos, const void* sin,
                  void* q_out, void* k_out, void* v_out, int B, int L, int H, int Hkv, int D, int K,
                  int64_t kv_stride_b, int64_t kv_stride_h, int pos0, void* stream);
/* Decode-step linear layer (SDAG.py:191-211, one token per prompt, B <= 4): out [B,N] = xin [B,K] . weight[N,K]^T
 * (+ bias) (+ residual [B,N]), weight rows streamed once.  prologue 0: xin = x [B,K]; 1: xin = RMSNorm(x) *
 * norm_weight (eps); 2: x is a gate|up row [B,2K] and xin = silu(gate) * up.  pair_swiglu != 0: weight is a
 * fused gate|up matrix [N = 2*I, K] and out [B, I] = silu(gate row j) * (up row j + I) — no bias / residual.
 * K % 8 == 0, N % 4 == 0. */
int llm_gemv(const void* x, const void* weight, const void* bias, const void* norm_weight, const void* residual,
             void* out, int B, int K, int N, int prologue, float eps, int pair_swiglu, void* stream);

/* The q/k/v projection of a decode step with apply_rotary_pos_emb and the KV-cache write in its epilogue (llm_gemv with
 * the RMSNorm prologue when norm_weight != NULL, followed by llm_rope_qkv for L = 1, in one launch): weight is the fused
 * [(H + 2*Hkv)*D, K] q|k|v matrix; q_out [B, 1, H, D]; k_out / v_out, cos / sin and the positions as in llm_rope_qkv. */
int llm_gemv_rope(const void* x, const void* weight, const void* bias, const void* norm_weight, float eps,
                  const void* cos, const void* sin, void* q_out, void* k_out, void* v_out, int B, int K, int H,
                  int Hkv, int D, int64_t kv_stride_b, int64_t kv_stride_h, int pos0, const int32_t* pos0_dev,
                  int pos0_dev_stride, void* stream);

/* The same decode-step linear layer (SDAG.py:191-211) for 1..32 prompts as a TMA weight stream on the tensor cores
 * (tcgen05, prompts on the padded M side; csrc/llm_skinny.cu): out [B, N] rows ldo apart = xin [B, K] (rows ldx apart) .
 * weight[N, K]^T (+ bias) (+ residual [B, N], rows ldr apart; out may alias it).  norm_weight != NULL: xin = RMSNorm(x) *
 * norm_weight (eps), computed in the kernel (B * K small enough to stage: SDAG_ERR_UNSUPPORTED otherwise — normalise with
 * llm_rmsnorm first).  pair_swiglu != 0: weight is the fused gate|up matrix [N = 2*I, K] and out [B, I] = silu(gate row j) *
 * (up row j + I), no bias / residual.  K % 64 == 0, N % 8 == 0 (16 with pair_swiglu), ldx / ldr / ldo % 8 == 0, B <= 32. */
int llm_skinny_linear(const void* x, int64_t ldx, const void* weight, const void* bias, const void* norm_weight, float eps,
                      const void* residual, int64_t ldr, void* out, int64_t ldo, int B, int N, int K, int pair_swiglu,
                      void* stream);

/* Hint: pull `bytes` at `ptr` (16-byte aligned) into L2.  The generation engine (SDAG.py:191-211 decode loop) launches it
 * on a side stream for the next layer's o_proj weights while the decode attention kernels run. */
int llm_prefetch_l2(const void* ptr, size_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Paths B1 / B6 — the sentence encoder behind `ranker.encode(...)` (embed_queries, dense.py:44-65;
 * encode_texts_with_ranker, malicious_selection.py:12-34; compute_doc_knn_for_docs, SDAG.py:46-48):
 * BERT-large geometry, bidirectional attention = sdag_prefill_fwd under a block map whose tiles are all
 * visible, cuBLAS linear layers, and the row-wise glue below.  bf16 tensors.
 * ---------------------------------------------------------------------------------------------- */
/* Block maps of B bidirectional sequences of lengths_dev[b] <= max_L tokens (right-padded batch): every
 * query tile sees every key tile of its own sequence, nothing of the padding. */
size_t sdag_full_blockmap_bytes(int max_L);
int sdag_build_blockmap_full(const int32_t* lengths_dev, int B, int max_L, int32_t* out, size_t stride_bytes,
                             void* stream);
/* BertSelfOutput / BertOutput: out = LayerNorm(x + residual) * weight + bias (residual may be NULL),
 * rows x dim, dim % 8 == 0. */
int enc_layernorm(const void* x, const void* residual, const void* weight, const void* bias, void* out,
                  int64_t rows, int dim, float eps, void* stream);
/* BertIntermediate activation: exact (erf) GELU over n_elems contiguous values, n_elems % 8 == 0. */
int enc_gelu(const void* x, void* out, int64_t n_elems, void* stream);
/* sentence-transformers Pooling(mean) + Normalize: hidden [B, L, dim] (element strides stride_b, stride_l),
 * lengths[b] valid tokens -> fp32 row (dst_rows ? dst_rows[b] : row0 + b) of out_f32 [*, dim], and its
 * bf16 copy in out_bf16 when not NULL — i.e. straight into the index shard ip_search reads. */
int enc_mean_pool_normalize(const void* hidden, int64_t stride_b, int64_t stride_l, const int32_t* lengths,
                            int B, int L, int dim, const int64_t* dst_rows, int64_t row0, float* out_f32,
                            void* out_bf16, int normalize, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SDAG_B200_H_ */
