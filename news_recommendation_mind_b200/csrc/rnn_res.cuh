// Persistent recurrence kernels with the recurrent weights resident in shared memory (MR_BF16 path of
// the LSTM / GRU user encoders, models/Encoders/RNN.py:36-104).  See rnn_res.cu.
#pragma once
#include "common.cuh"

namespace mr {

// true when W_hh (bf16) plus the per-step scratch fits one SM's shared memory for this shape
bool rnn_res_supported(int kind, int H);
int rnn_res_bpc(int kind, int B, int H);

// xp: [B*S, ldx] fp32 input projection (+ biases), ldx >= G*H, indexed by step; w_hh [G*H, H] fp32; the rest as rnn.cu
int rnn_res_fwd(int kind, const float* xp, int ldx, const float* w_hh, const float* b_hh, const float* h0, const int32_t* lens,
                float* gates, float* hs, float* cs, float* user, int B, int S, int H, void* scratch, cudaStream_t st);
// gib / ghb (optional, pre-zeroed bf16 [>= B*S rows, GHp16]): gate gradients written directly in the layout of the
// weight-gradient GEMMs instead of fp32 dgi / dgh; bias_part (optional, [grid = ceil(B / rnn_res_bpc)][2][G*H]): per-CTA
// column sums of dgi ([.][0]) and dgh ([.][1])
int rnn_res_bwd(int kind, const float* w_hh, const float* h0, const int32_t* lens, const float* gates, const float* hs,
                const float* cs, const float* d_user, float* dgi, float* dgh, float* d_h0, int B, int S, int H,
                void* scratch, cudaStream_t st, __nv_bfloat16* gib = nullptr, __nv_bfloat16* ghb = nullptr, int GHp16 = 0,
                float* bias_part = nullptr);
// scratch for the bf16 image of W_hh the kernels bulk-copy into shared memory
int64_t rnn_res_scratch_bytes(int kind, int H);


// ---- rnn_mma.cu: the same recurrence on mma.sync tensor-core MMAs, W_hh in registers + shared memory, h as bf16 hi + lo ----
bool rnn_mma_supported(int kind, int H);               // H <= 160, G*H <= 768
int64_t rnn_mma_scratch_bytes(int kind, int H);
int rnn_mma_fwd(int kind, const float* xp, int ldx, const float* w_hh, const float* b_hh, const float* h0, const int32_t* lens,
                float* gates, float* hs, float* cs, float* user, int B, int S, int H, void* scratch, cudaStream_t st);

// ---- rnn_tc.cu: the LSTM recurrence as tcgen05.mma batches, W_hh resident in shared memory (A operand), h as bf16 hi + lo (B operand) ----
bool rnn_tc_supported(int kind, int H);                // LSTM, H <= 160 (MINDREC_RNN_TC=0 disables)
int64_t rnn_tc_scratch_bytes(int kind, int H);
int rnn_tc_fwd(int kind, const float* xp, int ldx, const float* w_hh, const float* h0, const int32_t* lens, float* gates, float* hs,
               float* cs, float* user, int B, int S, int H, void* scratch, cudaStream_t st);

// backward recurrence (LSTM): W_hh^T resident in tensor memory; writes the bf16 gate-gradient rows (pitch GHp16, pre-zeroed),
// d_h0 (optional) and bias_part [rnn_tc_bwd_rows(B)][2][4H]
bool rnn_tc_bwd_supported(int kind, int H, int B);
int rnn_tc_bwd_rows(int B);
int64_t rnn_tc_bwd_scratch_bytes(int kind, int H);
int rnn_tc_bwd(int kind, const float* w_hh, const int32_t* lens, const float* gates, const float* cs, const float* d_user,
               __nv_bfloat16* gib, int GHp16, float* d_h0, float* bias_part, int B, int S, int H, void* scratch, cudaStream_t st);

}  // namespace mr
