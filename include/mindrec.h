/*
 * mindrec.h -- C ABI of libmindrec.so, the B200 (sm_100a) implementation of the TwoTower
 * news-recommendation hot path of tyh666/News-Recommendation-MIND.
 *
 * Every entry point below replaces one piece of PyTorch library work the reference does on the
 * path  models/TwoTower.py -> models/Encoders/<encoder>.py -> models/Modules/Attention.py  (driven by
 * utils/Manager.py::_train / _eval_fast).  The reference interface each one stands in for is cited
 * as file:line (paths relative to the reference root).
 *
 * Conventions
 *   - plain C: raw device pointers, explicit int64 sizes, no torch types;
 *   - the CALLER owns every buffer (inputs, outputs, saved-for-backward, workspace); the library
 *     allocates nothing persistent and never synchronises the device or touches the default stream;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - return value: 0 = MR_OK, negative = MR_ERR_*; a human-readable reason is kept per thread
 *     and returned by mr_last_error(); nothing throws or aborts across this boundary;
 *   - token ids / masks / labels may be int32 or int64 (`*_i64` flag) because the reference batch
 *     dict carries int64 (utils/MIND.py:352-363);
 *   - `precision`: MR_F32  = fp32 SIMT kernels (verification mode, 1e-5 parity with the reference)
 *                  MR_BF16 = bf16 operands on tcgen05 tensor cores with fp32 accumulation; every
 *                            non-GEMM step (bias, ReLU, tanh, softmax, pooling, cell state) in fp32.
 *   - there is no CPU fallback: on a device that is not compute capability 10.x every compute entry
 *     returns MR_ERR_NOT_SM100.
 */
#ifndef MINDREC_H_
#define MINDREC_H_

#include <stdint.h>

#if defined(__GNUC__)
#define MR_API __attribute__((visibility("default")))
#else
#define MR_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

enum {
  MR_OK = 0,
  MR_ERR_BAD_SHAPE = -1,
  MR_ERR_UNSUPPORTED = -2,
  MR_ERR_NOT_SM100 = -3,
  MR_ERR_LAUNCH = -4,
  MR_ERR_NULL = -5,
  MR_ERR_WORKSPACE = -6
};

enum { MR_F32 = 0, MR_BF16 = 1 };
enum { MR_RNN_LSTM = 0, MR_RNN_GRU = 1 };

MR_API int mr_version(void);
MR_API const char* mr_last_error(void);
/* MR_OK iff `device` is an sm_100-family GPU (B200). */
MR_API int mr_device_check(int device);
/* number of kernels this library has launched in this process (bench.py's gpu_launches). */
MR_API int64_t mr_launch_count(void);

/* ----------------------------------------------------------------------------------------------
 * Token-embedding gather.   models/Embeddings/BERT.py:24-41  (word_embeds = W[ids])
 *   out[t, :] = table[ids[t], :]           ids [T], table [V,E] fp32, out [T,E] fp32
 * Only needed when a caller wants the materialised [B,*,L,E] tensor (stand-alone module
 * interface); the fused news encoder below never materialises it.
 * -------------------------------------------------------------------------------------------- */
MR_API int mr_embed_gather_f32(const void* ids, int ids_i64, const float* table, float* out,
                        int64_t T, int64_t E, int64_t V, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Title rows by news id.   utils/MIND.py:347-355 (train) / :389-398 (dev): the dataset looks up
 * ``encoded_news[cdd_ids][:, :L]`` / ``attn_mask[...]`` per sample on the host and the batch carries the int64
 * token tensors over PCIe (7.2 MB per 256 impressions).  With the token table resident in HBM (int32
 * tok_ids / tok_mask [n_rows, L], row 0 = the empty article of MIND.py:125-127) the batch carries news ids only:
 *   out_ids[r, :]  = tok_ids[nid[r], :],  out_mask[r, :] = tok_mask[nid[r], :]
 * rows 0..n_a-1 from nid_a (candidates), rows n_a..n_a+n_b-1 from nid_b (clicked history; may be NULL with n_b = 0).
 * An id outside [0, n_rows) reads row 0.
 * -------------------------------------------------------------------------------------------- */
MR_API int mr_gather_titles(const int32_t* tok_ids, const int32_t* tok_mask, int64_t n_rows, int64_t L,
                     const void* nid_a, int64_t n_a, const void* nid_b, int64_t n_b, int nid_i64,
                     int32_t* out_ids, int32_t* out_mask, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Embedding-table gradient.   autograd of BERT.py:39 = embedding_dense_backward, padding_idx=0
 *   d_table[v, :] = sum over t with ids[t]==v of d_emb[t, :];  row `padding_idx` (and every id
 *   that does not occur) is written as zeros.  Atomic-free: a stable key sort of the token ids
 *   followed by a two-level segmented row reduction, so the result is bit-reproducible.
 *   d_emb is fp32 (dtype MR_F32) or bf16 (MR_BF16).
 * -------------------------------------------------------------------------------------------- */
MR_API int64_t mr_embed_grad_workspace_bytes(int64_t T, int64_t E, int64_t V);
MR_API int mr_embed_grad_segreduce(const void* ids, int ids_i64, const void* d_emb, int d_emb_dtype,
                            int64_t d_emb_ld /* row pitch of d_emb in elements, 0 = E; pad columns must be 0 */,
                            float* d_table, int64_t T, int64_t E, int64_t V, int64_t padding_idx,
                            void* workspace, int64_t workspace_bytes, void* stream);

/* ----------------------------------------------------------------------------------------------
 * CNN news encoder.   models/Encoders/CNN.py:30-51 (+ BERT.py:39 fused in front of it,
 *                      Attention.py:5-30,56-80 fused behind it)
 *   c[n,l,:]  = relu(conv_b + sum_tap conv_w[:, :, tap] . x[n, l+tap-1, :])     zero padded
 *   key       = tanh(c proj_w^T + proj_b)
 *   prob[n,:] = masked_softmax_l( query . key[n,l,:] / sqrt(H) , mask[n,:] )
 *   news[n,:] = sum_l prob[n,l] c[n,l,:]
 * Input is EITHER token ids (ids != NULL: rows of `table` are gathered on the fly, table is
 * [V,E] fp32 in MR_F32 or the bf16 shadow [V, align_up(E,64)] (zero padded) in MR_BF16) OR a dense
 * embedding tensor (ids == NULL, emb [N*L, E] fp32).
 * Saved for backward (caller allocated): c_save, key_save ([N*L,H] fp32, or bf16 [N*L, align_up(H,16)]
 * in MR_BF16; there c_save needs 32 more bytes per token BEHIND its N*L rows: the forward leaves the sign mask of c,
 * one bit per column, for the backward), prob [N,L] fp32.  MR_BF16 limits: H <= 256, L <= 128.  c_out (optional, may be NULL) receives the fp32 token-level output
 * the module interface returns as its first value.
 * -------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t N, L, E, H, V;
  int precision;
} mr_cnn_shape;

/* Token ids whose table rows the MR_BF16 gather keeps in shared memory (at most 4; e.g. PAD, [CLS], [SEP] --
 * in MIND-shaped batches they are >25 % of all positions and every SM would otherwise hit the same few L2
 * lines).  Purely a performance hint: results are identical.  Process-wide, set before launching. */
MR_API int mr_news_cnn_set_hot_tokens(const int64_t* ids, int n);
/* reps > 0: the bf16 table shadow passed to mr_news_cnn_fwd/bwd has n_hot*reps extra rows after row V-1; row
 * V + h*reps + k (k < reps) is a copy of hot row h.  The gather then runs on TMA (tile::gather4) and every CTA
 * reads its own replica of the hot rows.  0 (default): cp.async gather + shared-memory hot-row cache. */
MR_API int mr_news_cnn_set_hot_replicas(int reps);
MR_API int64_t mr_news_cnn_workspace_bytes(const mr_cnn_shape* s, int backward);
MR_API int mr_news_cnn_fwd(const mr_cnn_shape* s,
                    const void* ids, int ids_i64, const float* emb,
                    const void* mask, int mask_i64,           /* NULL = no mask */
                    const void* table,
                    const float* conv_w, const float* conv_b, /* [H,E,3], [H] */
                    const float* proj_w, const float* proj_b, /* [H,H], [H]   */
                    const float* query,                       /* [H]          */
                    void* c_save, void* key_save, float* prob, float* news,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* Backward of the above.  d_news [N,H] fp32 (+ optional d_c [N*L,H] fp32 for the token-level
 * output).  Produces d_conv_w [H,E,3], d_conv_b [H], d_proj_w [H,H], d_proj_b [H], d_query [H]
 * (all overwritten, fp32) and d_emb (fp32 [N*L,E] in MR_F32, bf16 [N*L, align_up(E,16)] with zero
 * padding columns in MR_BF16; may be NULL when the input needs no gradient).  `table`/`ids` or `emb` are the forward inputs. */
MR_API int mr_news_cnn_bwd(const mr_cnn_shape* s,
                    const void* ids, int ids_i64, const float* emb, const void* table,
                    const float* conv_w, const float* proj_w, const float* query,
                    const void* c_save, const void* key_save, const float* prob,
                    const float* d_news, const float* d_c,
                    float* d_conv_w, float* d_conv_b, float* d_proj_w, float* d_proj_b,
                    float* d_query, void* d_emb,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* MR_BF16, ids path: the same backward, but the gradient of the token table is produced directly (dense [V,E] fp32,
 * padding row zero; replaces mr_news_cnn_bwd(d_emb) + mr_embed_grad_segreduce = autograd of BERT.py:39 + CNN.py:41).
 * The conv-output gradient is first summed per vocabulary row (sorted token positions, fixed-order partial sums, no
 * atomics), so the table- and filter-gradient GEMMs run over V rows instead of N*L tokens.  `table_bf16` is the padded
 * bf16 table the forward gathered from, with table_rows >= align_up(V, 32) rows (rows >= V zero).  E % 4 == 0. */
MR_API int64_t mr_news_cnn_bwd_table_workspace_bytes(const mr_cnn_shape* s);
/* The grouping plan (token positions sorted by id, segment bounds, chunk table) depends on the ids only.  It can be
 * built ahead of the backward -- e.g. on a second stream while the forward runs -- and passed as `group_plan`
 * (NULL: mr_news_cnn_bwd_table builds it itself).  n_tokens = N*L. */
MR_API int64_t mr_token_group_plan_bytes(int64_t n_tokens, int64_t V);
MR_API int mr_token_group_plan(const void* ids, int ids_i64, int64_t n_tokens, int64_t V,
                    void* plan, int64_t plan_bytes, void* stream);
MR_API int mr_news_cnn_bwd_table(const mr_cnn_shape* s,
                    const void* ids, int ids_i64, const void* table_bf16, int64_t table_rows,
                    const float* conv_w, const float* proj_w, const float* query,
                    const void* c_save, const void* key_save, const float* prob, const float* d_news,
                    float* d_conv_w, float* d_conv_b, float* d_proj_w, float* d_proj_b, float* d_query,
                    float* d_table, int64_t padding_idx,
                    const void* group_plan, int64_t group_plan_bytes,
                    void* table_ready_event,   /* optional cudaEvent_t recorded on `stream` as soon as d_table is complete */
                    void* workspace, int64_t workspace_bytes, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Recurrent user encoders.   models/Encoders/RNN.py:36-73 (RNN_User_Encoder, LSTM / GRU) and
 *                             RNN.py:76-104 (LSTUR: h0 given, no length masking)
 *   x [B,S,H] fp32 history vectors, lens [B] int32 (= his_mask.sum, >=1), optional h0 [B,H].
 *   user[b,:] = hidden state after step lens[b]-1 (pack_padded_sequence + h_n semantics).
 *   Gate order i,f,g,o (LSTM) / r,z,n (GRU); both bias vectors are added.
 *   `reverse` != 0 runs over the flipped sequence (descend_history / LSTUR).
 *   Saved for backward: gates [B,S,G*H] fp32 (activated gates), hs [B,S,H], cs [B,S,H] (LSTM only); entries of steps
 *   >= lens[b] are unspecified in MR_BF16 (never read by mr_rnn_user_bwd), zero in MR_F32.
 * -------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t B, S, H;
  int kind;       /* MR_RNN_LSTM | MR_RNN_GRU */
  int reverse;
  int precision;
} mr_rnn_shape;

MR_API int64_t mr_rnn_workspace_bytes(const mr_rnn_shape* s, int backward);
MR_API int mr_rnn_user_fwd(const mr_rnn_shape* s, const float* x, const int32_t* lens, const float* h0,
                    const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                    float* gates, float* hs, float* cs, float* user,
                    void* workspace, int64_t workspace_bytes, void* stream);
MR_API int mr_rnn_user_bwd(const mr_rnn_shape* s, const float* x, const int32_t* lens, const float* h0,
                    const float* w_ih, const float* w_hh,
                    const float* gates, const float* hs, const float* cs, const float* d_user,
                    float* d_x, float* d_h0, float* d_w_ih, float* d_w_hh, float* d_b_ih, float* d_b_hh,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Query attention pooling over a set of vectors.
 *   models/Encoders/Pooling.py:12-25 (Attention_Pooling) and the pooling step of
 *   MHA.py:38,71 -- scaled_dp_attention(query[1,H], r, r, mask) with key = value = r.
 *   r [B,S,H] fp32, mask [B,S] fp32 (NULL = none), query [H] -> out [B,H], prob [B,S] (saved).
 * Average_Pooling (Pooling.py:32-43): mean over S, mask ignored.
 * -------------------------------------------------------------------------------------------- */
MR_API int mr_attnpool_fwd(const float* r, const float* mask, const float* query, float* prob, float* out,
                    int64_t B, int64_t S, int64_t H, void* stream);
MR_API int mr_attnpool_bwd(const float* r, const float* query, const float* prob, const float* d_out,
                    float* d_r, float* d_query_partial /* [B,H] */,
                    int64_t B, int64_t S, int64_t H, void* stream);
MR_API int mr_avgpool_fwd(const float* r, float* out, int64_t B, int64_t S, int64_t H, void* stream);
MR_API int mr_avgpool_bwd(const float* d_out, float* d_r, int64_t B, int64_t S, int64_t H, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Multi-head self-attention block.   models/Modules/Attention.py:83-147 (shared q/k projection,
 *   no output projection) as used by MHA_Encoder (MHA.py:21-39) and MHA_User_Encoder (:58-75).
 *   qk [n,len,hn*dk], v [n,len,hn*dv] are the already projected tensors (projection GEMMs are
 *   mr_linear_*).  mask [n,len] fp32 0/1 -> pair mask m_i*m_j (Attention.py:33-53).
 *   ctx [n,len,hn*dv];  prob [n,hn,len,len] saved for backward.
 * -------------------------------------------------------------------------------------------- */
MR_API int mr_mha_core_fwd(const float* qk, const float* v, const float* mask, float* prob, float* ctx,
                    int64_t n, int64_t len, int64_t hn, int64_t dk, int64_t dv, void* stream);
MR_API int mr_mha_core_bwd(const float* qk, const float* v, const float* prob, const float* d_ctx,
                    float* d_qk, float* d_v,
                    int64_t n, int64_t len, int64_t hn, int64_t dk, int64_t dv, void* stream);

/* The same attention core with explicit row pitches, register-tiled (csrc/mha_attn.cu); limits len <= 64, dk, dv <= 32 (what
 * the news / user encoders use).  q|k rows at qk + row * ldq (+ h * dk), v rows at v + row * ldv (+ h * dv): both may be column
 * slices of one projection output.  ctx [n, len, hn*dv] and d_ctx are dense; the backward writes d_qk with pitch ldo_q and d_v
 * with pitch ldo_v (slices of the projection's gradient buffer).  mr_mha_core_* use these kernels whenever the shape fits. */
MR_API int mr_mha_attn_fwd(const float* qk, int64_t ldq, const float* v, int64_t ldv, const float* mask, float* prob, float* ctx,
                    int64_t n, int64_t len, int64_t hn, int64_t dk, int64_t dv, void* stream);
MR_API int mr_mha_attn_bwd(const float* qk, int64_t ldq, const float* v, int64_t ldv, const float* prob, const float* d_ctx,
                    float* d_qk, int64_t ldo_q, float* d_v, int64_t ldo_v,
                    int64_t n, int64_t len, int64_t hn, int64_t dk, int64_t dv, void* stream);

/* Dense layer y = act(x W^T + b) and its gradients (nn.Linear, Attention.py:101-102; CNN.py:22).
 *   x [M,K], w [N,K], b [N] (may be NULL), y [M,N];  act: 0 none, 1 relu, 2 tanh.
 *   bwd: d_x = d_y W  (NULL to skip), d_w = d_y^T x, d_b = colsum(d_y); d_y is the gradient wrt
 *   the pre-activation (callers fold the activation derivative in). */
MR_API int64_t mr_linear_workspace_bytes(int64_t M, int64_t N, int64_t K);
MR_API int mr_linear_fwd(const float* x, const float* w, const float* b, float* y,
                  int64_t M, int64_t N, int64_t K, int act, int precision, void* stream);
MR_API int mr_linear_bwd(const float* x, const float* w, const float* d_y, float* d_x, float* d_w, float* d_b,
                  int64_t M, int64_t N, int64_t K, int precision,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* The same dense layer on the tcgen05 tensor cores (MR_BF16 path of the attention projections keyProject / valueProject,
 * Attention.py:101-102,125-127): bf16 operands, fp32 accumulation and fp32 output y [M, ldy] (ldy % 4 == 0, ldy >= N rounded up
 * to 4; columns N..ldy-1 are padding).  The input is EITHER a dense fp32 matrix x [M, K] OR (x == NULL) rows of the padded bf16
 * token table gathered by ids [M] inside the GEMM (BERT.py:39 fused in, the [M, K] embedding tensor is never materialised).
 * bwd: dense -> d_x [M, K] (NULL to skip), d_w [N, K], d_b [N] (NULL to skip); gather -> d_table [V, K] instead of d_x (row
 * `padding_idx` zero; the gradient rows are first summed per token id, so both GEMMs run over V rows), K % 4 == 0, and the bf16
 * table must have table_rows >= V rounded up to 32 (zero padded).  d_y [M, ldy] fp32. */
MR_API int64_t mr_linear_tc_workspace_bytes(int64_t M, int64_t N, int64_t K, int64_t V /* 0 = dense */, int backward);
MR_API int mr_linear_tc_fwd(const float* x, const void* ids, int ids_i64, const void* table_bf16, int64_t table_ld, int64_t V,
                     const float* w, const float* b, float* y, int64_t ldy, int64_t M, int64_t N, int64_t K,
                     void* workspace, int64_t workspace_bytes, void* stream);
MR_API int mr_linear_tc_bwd(const float* x, const void* ids, int ids_i64, const void* table_bf16, int64_t table_ld,
                     int64_t table_rows, int64_t V, int64_t padding_idx, const float* w, const float* d_y, int64_t ldy,
                     float* d_x, float* d_table, float* d_w, float* d_b, int64_t M, int64_t N, int64_t K,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* LayerNorm over the last axis, eps 1e-5 (MHA.py:18,37), with optional fused inverted-dropout keep
 * mask (uint8, NULL = none).  mean/rstd [M] saved. */
MR_API int mr_layernorm_fwd(const float* x, const float* gamma, const float* beta, const uint8_t* keep, float keep_scale,
                     float* y, float* mean, float* rstd, int64_t M, int64_t H, void* stream);
MR_API int mr_layernorm_bwd(const float* x, const float* gamma, const uint8_t* keep, float keep_scale,
                     const float* mean, const float* rstd, const float* d_y,
                     float* d_x, float* d_gamma_partial, float* d_beta_partial, int64_t n_partial,
                     int64_t M, int64_t H, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Candidate scoring.   models/TwoTowerBaseModel.py:51-75 and utils/Manager.py:381-382,641
 *   score[b,c] = <cdd[b,c,:], user[b,:]> / sqrt(H)
 *   training: logp = log_softmax_c(score); optional fused NLL: loss_sum += -logp[b,label[b]]
 *   eval:     prob = sigmoid(score)
 * bwd takes d_logp [B,C] (gradient wrt the log-probabilities) and returns d_cdd, d_user.
 * -------------------------------------------------------------------------------------------- */
MR_API int mr_score_logsoftmax_fwd(const float* cdd, const float* user, float* logp,
                            const void* label, int label_i64, float* loss_mean /* 1 float or NULL */,
                            int64_t B, int64_t C, int64_t H, void* stream);
MR_API int mr_score_logsoftmax_bwd(const float* cdd, const float* user, const float* logp, const float* d_logp,
                            float* d_cdd, float* d_user, int64_t B, int64_t C, int64_t H, void* stream);
/* apply_sigmoid = 0 returns the raw scores of compute_score (TwoTowerBaseModel.py:61). */
MR_API int mr_score_sigmoid_fwd(const float* cdd, const float* user, float* prob,
                         int64_t B, int64_t C, int64_t H, int apply_sigmoid, void* stream);
/* Fast evaluation (TwoTowerBaseModel.py:78-84 + Manager.py:514-517), batched over impressions in
 * CSR form: candidates of impression i are cdd_id[offsets[i] .. offsets[i+1]);
 *   prob[j] = sigmoid(<news_table[cdd_id[j]], user[i]> / sqrt(H)). */
MR_API int mr_score_sigmoid_gather_fwd(const float* news_table, const void* cdd_id, int id_i64,
                                const int64_t* offsets, const float* user, float* prob,
                                int64_t n_impr, int64_t n_cand, int64_t n_news_rows, int64_t H, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Ranking metrics per impression (CSR).   utils/Manager.py:1205-1344 (cal_metric: auc, mean_mrr,
 * ndcg@5, ndcg@10) and Manager.py:842-850 (ordinal rank for prediction.txt).
 *   order: descending score, ties by ascending candidate position.
 *   metrics [n_impr,4] fp64 = (auc, mrr, ndcg@5, ndcg@10); rank [n_cand] int32 (1 = best; may be NULL).
 * -------------------------------------------------------------------------------------------- */
MR_API int mr_rank_metrics(const float* prob, const float* label, const int64_t* offsets,
                    double* metrics, int32_t* rank, int64_t n_impr, int64_t n_cand, void* stream);

/* ----------------------------------------------------------------------------------------------
 * Adam.   utils/Manager.py:404-413 (torch.optim.Adam, betas .9/.999, eps 1e-8, no weight decay)
 *   one launch per parameter tensor; `step` is 1-based; grad_scale multiplies g first (1/world
 *   for a summed all-reduce).  Optionally refreshes a bf16 shadow copy (row pitch `shadow_ld`
 *   elements for rows of `row_len`; NULL = none).
 * -------------------------------------------------------------------------------------------- */
MR_API int mr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int64_t step,
                 double lr, double beta1, double beta2, double eps, double grad_scale,
                 void* shadow_bf16, int64_t row_len, int64_t shadow_ld, void* stream);

/* The same update for up to 24 parameter tensors in one launch (p/g/m/v/numel/lr are HOST arrays of n_tensors entries;
 * lr per tensor = the two learning-rate groups of Manager._get_optim).  `shadow_tensor` = index of the tensor whose bf16
 * shadow is refreshed (-1 or shadow_bf16 == NULL: none). */
MR_API int mr_adam_step_multi(int n_tensors, float* const* p, const float* const* g, float* const* m, float* const* v,
                 const int64_t* numel, const double* lr, int64_t step, double beta1, double beta2, double eps,
                 double grad_scale, int shadow_tensor, void* shadow_bf16, int64_t row_len, int64_t shadow_ld, void* stream);

/* Same, with every step-dependent scalar read from DEVICE memory: `dyn_device` = 4 + n_tensors floats
 *   {1 / (1 - beta1^t), 1 / sqrt(1 - beta2^t), grad_scale, unused, lr[0], ..., lr[n_tensors-1]}
 * (`step`, the host `lr` array and `grad_scale` are then ignored) -- the form a captured CUDA graph can replay for every
 * step while a learning-rate schedule (Manager.py:414-420) keeps moving the rates.  dyn_device == NULL: mr_adam_step_multi. */
MR_API int mr_adam_step_multi_dyn(int n_tensors, float* const* p, const float* const* g, float* const* m, float* const* v,
                 const int64_t* numel, const double* lr, int64_t step, double beta1, double beta2, double eps,
                 double grad_scale, int shadow_tensor, void* shadow_bf16, int64_t row_len, int64_t shadow_ld,
                 const float* dyn_device, void* stream);

/* Mean negative log-likelihood (nn.NLLLoss(), utils/Manager.py:381-382,641) over logp [B,C]:
 *   fwd: loss[0] = -(1/B) sum_b logp[b, label[b]];   bwd: d_logp[b,c] = -(d_loss/B) [c == label[b]] */
MR_API int mr_nll_loss_fwd(const float* logp, const void* label, int label_i64, float* loss,
                    int64_t B, int64_t C, void* stream);
MR_API int mr_nll_loss_bwd(const void* label, int label_i64, const float* d_loss, float* d_logp,
                    int64_t B, int64_t C, void* stream);
/* Diagnostic: one tcgen05 tile D[128,N] = A x B^T (bf16 operands, fp32 accumulate) built with the
 * same shared-memory descriptors the production kernels use (csrc/tc05.cuh); a_mn/b_mn select
 * K-major (rows = M/N index) or MN-major (rows = K index) operands, a_shift reads A `a_shift` rows
 * later (zero halo).  Used by tests/test_gpu_tc.py to pin the descriptor conventions on hardware. */
MR_API int mr_tc_selftest(const float* a, int64_t ra, int64_t ca, const float* b, int64_t rb, int64_t cb, float* d,
                   int a_mn, int b_mn, int64_t N, int64_t K, int64_t a_shift, int64_t halo, int swap,
                   int a_layout, int b_layout, int base_off_mode, void* stream);

/* fp32 [rows, cols] -> bf16 [rows, ld] (zero padded columns). */
MR_API int mr_cast_pad_bf16(const float* src, void* dst, int64_t rows, int64_t cols, int64_t ld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MINDREC_H_ */
