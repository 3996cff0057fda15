// "Token-reduction GEMM": weight gradients of the bf16 news encoder on tcgen05.
//
//     out[tap][i][j] = sum over tokens t=(title,l) of  P[(title, l + tap - c), i] * Q[t, j]        c = (taps-1)/2
//
// (zero outside the title).  P is the conv input (token rows gathered from the bf16 table by id) or a dense
// bf16 activation matrix; Q is a dense bf16 gradient matrix.  With taps = 3, P = x, Q = dconv this is the
// Conv1d filter gradient (autograd of models/Encoders/CNN.py:41); with taps = 1, P = c, Q = dkey it is the
// gradient of wordQueryProject.weight (CNN.py:46).
//
// The token index is the K dimension of the MMA, so both operands are MN-major views of the same panel
// layout the forward uses (tc05.cuh): a tile of 128 token rows is staged once (rows ordered l*G + g, zero
// halo), and tap shifts are descriptor start-address shifts of G rows.  CTA (m, s) owns the 128-wide slice
// m of the i index and every S-th token tile; it accumulates taps x [128, NQ] fp32 in TMEM over all its
// tiles, writes one partial, and tokred_reduce sums the S partials in a fixed order (deterministic, no atomics).
#pragma once
#include "common.cuh"
#include "tc05.cuh"
#include "tma.cuh"

namespace mr {

constexpr int TR_THREADS = 288;   // warps 0-3 epilogue, 4 MMA, 5-8 producers

struct TokRedArgs {
  int64_t n_titles, n_tiles;
  int L, G, taps;
  const void* ids; int ids_i64;
  int64_t hot_ids[4]; int n_hot;                    // gather mode: table rows kept in shared memory (see tapgemm.cuh)
  const __nv_bfloat16* p; int64_t ldp, V; int KP;   // KP: columns of P to reduce (multiple of 8)
  const __nv_bfloat16* q; int64_t ldq; int NQ;      // NQ: columns of Q (multiple of 16, taps*NQ <= 512)
  float* partial;                                   // [n_mtiles][S][taps][128][NQ]
  int n_mtiles, S, halo, n_stages;
  uint32_t p_ps, q_ps, p_bytes, stage_bytes, q_rb;   // block strides (bytes), Q row bytes
  int p_tma, q_tma;                                    // operand staged by TMA tile loads instead of cp.async
  long long* dbg;                                      // optional wait counters (debug)
  int q_layout;                                        // UMMA layout type of Q (2 / 4 / 6)
};

struct TokRedPlan {
  alignas(64) CUtensorMap pmap;       // dense P: 3-D (column, title, position) view, box {64, G, L}, SWIZZLE_128B
  alignas(64) CUtensorMap qmap;       // Q: same view, box {q_rb/2, G, L}, swizzle = q_rb
  TokRedArgs args;
  size_t smem_bytes;
  int grid;
};

int64_t tokred_partial_bytes(int64_t n_titles, int L, int taps, int KP, int NQ);
int tokred_plan(TokRedArgs& a, TokRedPlan* plan);
int tokred_launch(const TokRedPlan& plan, cudaStream_t stream);
// dst[j*dj + i*di + tap*dt] = sum_s partial[i/128][s][tap][i%128][j]   for i < i_valid, j < j_valid
int tokred_reduce(const TokRedPlan& plan, float* dst, int i_valid, int j_valid, int64_t dj, int64_t di, int64_t dt,
                  cudaStream_t stream);

}  // namespace mr
