// Fused "tail" of the CNN news encoder (models/Encoders/CNN.py:44-48: wordQueryProject + tanh, additive-attention
// pooling over the title) -- one persistent tcgen05 kernel per direction instead of a GEMM, a pooling kernel and a
// weight-gradient GEMM that each stream c / key / dkp through HBM (cnn_tail.cu).
#pragma once
#include "common.cuh"

namespace mr {

// titles of 16..32 tokens, Hp = align16(H) <= 160 (3 * Hp TMEM columns), row pitch of c / key = Hp
bool cnn_tail_supported(int64_t L, int64_t Hp);
int64_t cnn_tail_bwd_workspace_bytes(int64_t n_titles, int64_t L, int64_t Hp);
// backward of  key = tanh(c Wq^T + bq),  p = masked softmax(<q, key> / sqrt(H)),  news = sum_l p c   given d_news:
//   dconv      = relu'(c) * (p d_news + dkp Wq)            bf16 [T, Hp]   (gradient wrt the conv pre-activation)
//   d_proj_w   = dkp^T c,  d_proj_b = sum_t dkp,  d_query = sum_t ds key,  d_conv_b = sum_t dconv
// with dkp = ds q (1 - key^2), ds = softmax backward of <d_news, c> (Attention.py:77-80).
// wq_img: bf16 panel image of Wq as the K-major B operand [k][n] (tapgemm_pack(proj_w, ., 1, Hp, Hp, H, H, 1, H, 0), replica 0).
// cmask: [T][32 bytes], bit j of a row = (c[row, j] > 0), written by the forward (conv epilogue / cmask_from_c_kernel).
int cnn_tail_bwd(int64_t n_titles, int64_t L, int64_t H, const __nv_bfloat16* c, const __nv_bfloat16* key, const uint8_t* cmask, const float* prob,
                 const float* d_news, const float* query, const uint8_t* wq_img, __nv_bfloat16* dconv, float* d_proj_w,
                 float* d_proj_b, float* d_query, float* d_conv_b, void* ws, int64_t wsb, cudaStream_t st);

// forward: key (bf16 [T, Hp], saved for the backward), prob (fp32 [T]) and news (fp32 [N, H]) from c (bf16 [T, Hp]).
// wq_img: bf16 panel image of Wq as the K-major B operand [n][k] (tapgemm_pack(proj_w, ., 1, Hp, Hp, H, H, H, 1, 0), replica 0).
int cnn_tail_fwd(int64_t n_titles, int64_t L, int64_t H, const __nv_bfloat16* c, const void* mask, int mask_i64, const float* query,
                 const float* proj_b, const uint8_t* wq_img, __nv_bfloat16* key, float* prob, float* news, cudaStream_t st);

}  // namespace mr
