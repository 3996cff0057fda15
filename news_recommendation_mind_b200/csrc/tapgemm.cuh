// "Tap GEMM": the tensor-core workhorse of the bf16 news-encoder path (forward conv, projection,
// projection backward, conv data gradient).  For every token row t = (title, l) it computes
//
//     out[t, n] = epi( sum_{tap} sum_{k<K}  A[(title, l + dir*(tap-c)), k] * W[tap][n, k] )
//
// with zeros outside the title (the conv's zero padding, models/Encoders/CNN.py:12-17), c = (taps-1)/2.
// A is either gathered on the fly from the bf16 token table by token id (BERT.py:39 fused in) or a dense
// bf16 activation matrix.  No im2col copy is made: a tile of G titles is staged ONCE in shared memory in
// the panel layout of tc05.cuh with rows ordered  r = l*G + g  (position-major), so "one position to the
// left/right" is "G rows up/down" for every title at once, and the rows above/below the tile are a
// permanent zero halo.  Each tap then feeds tcgen05.mma the same staged bytes through a descriptor whose
// start address is moved by G rows.
//
// One persistent CTA per SM, 10 warps:
//   warps 0-3  epilogue   : TMEM -> registers (tcgen05.ld) -> bias/activation -> bf16 -> global
//   warp  4    MMA issuer : one thread issues tcgen05.mma into a double-buffered TMEM accumulator
//   warp  5    W producer : cp.async.bulk of pre-packed weight blocks (one block per k-chunk x tap)
//   warps 6-9  A producers: 16-byte cp.async gather of token rows into the panel layout (ring of k-chunks)
// mbarrier rings: a_full/a_empty, b_full/b_empty (smem), t_full/t_empty (TMEM accumulators).
#pragma once
#include "common.cuh"
#include "tc05.cuh"
#include "tma.cuh"

namespace mr {

enum { TG_EPI_BIAS_RELU = 0, TG_EPI_BIAS_TANH = 1, TG_EPI_RELUGRAD = 2, TG_EPI_STORE = 3, TG_EPI_BIAS_F32 = 4,
       // out = relu'(e1) * (acc + prob[t] * dnp[title, :] (+ e0)): the pooling gradient p * d_news is formed here instead of
       // being read back from memory, and relu'(e1) comes as a 1-bit-per-column mask (n_total <= 160)
       TG_EPI_RELUGRAD_POOL = 5 };
constexpr int TG_KC = 64;          // k-chunk (elements) = 8 panels
constexpr int TG_THREADS = 320;
constexpr int TG_THREADS2 = 384;   // two epilogue groups (dense A by TMA): warps 0-3 / 8-11 epilogue, 4 MMA, 5 weights, 6 TMA
constexpr int TG_MAX_SLOTS = 8;
constexpr int TG_W_REPS = 4;       // replicas of the packed weights in global memory (spreads L2 slice load)

struct TapGemmArgs {
  int64_t n_titles, n_tiles;
  int L, G, taps, dir, K;
  int n_sub, nsz[2];                       // N sub-tiles, each a multiple of 16 and <= 256
  const void* ids; int ids_i64;            // gather mode when ids != nullptr
  int hot_reps;                            // > 0: table rows V + h*hot_reps + k (k < hot_reps) are copies of hot row h
  int64_t hot_ids[4]; int n_hot;           // gather mode: table rows kept in shared memory (PAD/[CLS]/[SEP]: every
                                           // CTA would otherwise hammer the same few L2 lines)
  const __nv_bfloat16* a; int64_t lda, V;  // gather: table [V, lda];  dense: activations [n_titles*L, lda]
  const uint8_t* wpack;                    // TG_W_REPS replicas of [chunk][tap] blocks of b_slot_bytes
  int w_reps; int64_t w_rep_stride;
  int rotate;                              // per-CTA rotated (k-chunk, tap) order (MINDREC_ROTATE=1); off = batch-invariant sums
  int epi; const float* bias; const float* bias2; int n_valid;   // bias2: optional second bias (LSTM b_ih + b_hh)
  int64_t n_rows;                          // rows of the problem (<= n_titles * L); 0 = n_titles * L
  float* out_f32; int n_store;             // TG_EPI_BIAS_F32: fp32 output [rows, ldo], columns >= n_store are not written
  const __nv_bfloat16* e0; const __nv_bfloat16* e1; int64_t lde;
  const float* prob; const float* dnp; int64_t ldn;   // TG_EPI_RELUGRAD_POOL: prob [rows], dnp [n_titles, ldn] fp32 (ldn % 4 == 0)
  const uint8_t* cmask;                    // ... and the sign mask of e1: [rows][32 bytes], bit j of the row = (e1[row, j] > 0)
  float* colsum_out;                       // RELUGRAD / RELUGRAD_POOL: optional [grid * 4][n_total] partial column sums of `out`
  uint8_t* cmask_out;                      // two-CTA BIAS_RELU epilogue: optional [rows][32 bytes] sign mask of the stored rows (bit j = out[row, j] > 0)
  __nv_bfloat16* out; int64_t ldo;
  int ns_a, ns_b, halo;
  int use_tma;                             // A tiles staged by TMA (3-D tile load / gather4) instead of cp.async
  long long* dbg;                          // optional [grid][4 roles][5] wait-cycle counters (see tapgemm.cu)
  uint32_t a_slot_bytes, b_slot_bytes, a_ps;
};

// ---- host side -------------------------------------------------------------------------------
struct TapGemmPlan {
  alignas(64) CUtensorMap tmap;            // A operand: 2-D table (gather4) or 3-D [cols, title, position] view
  TapGemmArgs args;
  size_t smem_bytes;
  int grid;
  bool epi2;                               // two-epilogue-group instantiation (384 threads)
  int colsum_rows;                         // rows of args.colsum_out this launch writes (grid x epilogue warps)
};

inline int64_t tapgemm_pack_bytes(int taps, int n_total, int K) {      // all replicas
  return (int64_t)TG_W_REPS * ((K + TG_KC - 1) / TG_KC) * taps * (TG_KC / 8) * n_total * 16;
}

int tapgemm_pack(const float* src, uint8_t* dst, int taps, int n_total, int K, int n_valid, int k_valid, int64_t sn,
                 int64_t sk, int64_t st, cudaStream_t stream);
// K index = concatenated blocks of k_mod columns: W[n, b*k_mod + kk] = src[n*sn + kk*sk + b*sb] for kk < k_valid (tapgemm.cu)
int tapgemm_pack_blocks(const float* src, uint8_t* dst, int taps, int n_total, int K, int n_valid, int k_valid, int64_t sn,
                        int64_t sk, int64_t st, int k_mod, int64_t sb, cudaStream_t stream);
// fills the derived fields of `a` (G, halo, slots, tiles); returns MR_OK or an error
int tapgemm_plan(TapGemmArgs& a, TapGemmPlan* plan);
int tapgemm_launch(const TapGemmPlan& plan, cudaStream_t stream);
// two-CTA (cta_group::2) variant with the weights resident in shared memory -- tapgemm2.cu
bool tapgemm2_supported(const TapGemmArgs& a);
int64_t tapgemm2_pack_bytes(int taps, int N, int K);
int tapgemm2_pack(const float* src, uint8_t* dst, int taps, int N, int K, int n_valid, int k_valid, int64_t sn, int64_t sk,
                  int64_t st, cudaStream_t stream);
int tapgemm2_run(TapGemmArgs a, const uint8_t* wpack2, cudaStream_t stream);
int sm_count();
bool use_tma_gather();
bool use_tma_default();                  // MINDREC_TMA=0 switches the producers back to cp.async
constexpr int TG_MAX_HOT = 4;
// process-wide list of "hot" token ids (set through mr_news_cnn_set_hot_tokens); n <= TG_MAX_HOT
int hot_tokens(int64_t* out);
int hot_replicas();                      // replicas per hot row appended to the bf16 table (0 = none)
// debug: when set, every tap-GEMM launch writes its per-role wait counters here ([148][4][5] int64)
extern long long* g_tapgemm_dbg;

}  // namespace mr
