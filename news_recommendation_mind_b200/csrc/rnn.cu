// Recurrent user encoders (LSTM / GRU over the clicked-history news vectors).
// Reference: models/Encoders/RNN.py:36-73 (RNN_User_Encoder: pack_padded_sequence + h_n) and
// RNN.py:76-104 (LSTUR_User_Encoder: h0 from the user table, flipped history, no packing).
//
// Structure (both precisions share it; the state, gates and cell arithmetic are always fp32):
//   1. input projection  xp[b,s,:] = x[b,pos(s),:] W_ih^T + b_ih (+ b_hh for the LSTM) -- one GEMM;
//   2. persistent recurrence kernel: one CTA owns RNN_BPC sequences for all S steps, the hidden state
//      lives in shared memory, W_hh^T is streamed through L1/L2 (fp32) -- no per-step launch;
//   3. backward: persistent reverse-time kernel producing the pre-activation gate gradients, then
//      three GEMMs (d_x, d_W_ih, d_W_hh) and column sums for the biases.
// pos(s) = s, or S-1-s when `reverse` (the reference flips the padded history *before* packing).
#include "gemm_simt.cuh"
#include "rnn_res.cuh"
#include <stdlib.h>
#include "tapgemm.cuh"
#include "tokred.cuh"

namespace mr {

constexpr int RNN_BPC = 4;        // sequences per CTA
constexpr int RNN_THREADS = 256;

__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
  // out[c, r] = in[r, c]
  __shared__ float tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int r = r0 + j;
    if (r < R && c < C) tile[j][threadIdx.x] = in[(int64_t)r * C + c];
  }
  __syncthreads();
  int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int cc = c0 + j;
    if (r < R && cc < C) out[(int64_t)cc * R + r] = tile[threadIdx.x][j];
  }
}

template <int KIND>   // 0 LSTM, 1 GRU
__global__ void __launch_bounds__(RNN_THREADS)
rnn_fwd_kernel(const float* __restrict__ xp,      // [B,S,G*H] input projection (step-major: index s)
               const float* __restrict__ whhT,    // [H, G*H]
               const float* __restrict__ b_hh,    // [G*H] (GRU only; LSTM folds it into xp)
               const float* __restrict__ h0,      // [B,H] or null
               const int32_t* __restrict__ lens,  // [B] or null (= S)
               float* __restrict__ gates,         // [B,S,G*H] activated gates
               float* __restrict__ hs,            // [B,S,H]
               float* __restrict__ cs,            // [B,S,H]  LSTM: cell state; GRU: W_hn h + b_hn
               float* __restrict__ user,          // [B,H]
               int B, int S, int H) {
  constexpr int G = KIND == 0 ? 4 : 3;
  extern __shared__ float smem[];
  const int GH = G * H;
  float* h_s = smem;                      // [BPC][H]
  float* c_s = h_s + RNN_BPC * H;         // [BPC][H]
  float* pre = c_s + RNN_BPC * H;         // [BPC][GH]
  __shared__ int len_s[RNN_BPC];
  const int b0 = blockIdx.x * RNN_BPC;
  const int tid = threadIdx.x;
  for (int i = tid; i < RNN_BPC * H; i += RNN_THREADS) {
    int b = b0 + i / H;
    h_s[i] = (h0 != nullptr && b < B) ? h0[(int64_t)b * H + (i % H)] : 0.f;
    c_s[i] = 0.f;
  }
  if (tid < RNN_BPC) {
    int b = b0 + tid;
    int l = (b < B) ? (lens ? lens[b] : S) : 0;
    len_s[tid] = l < 0 ? 0 : (l > S ? S : l);
  }
  __syncthreads();
  int max_len = 0;
#pragma unroll
  for (int i = 0; i < RNN_BPC; ++i) max_len = max(max_len, len_s[i]);

  for (int s = 0; s < max_len; ++s) {
    // recurrent matvec for all BPC sequences; thread owns gate columns n = tid, tid+256, ...
    for (int n = tid; n < GH; n += RNN_THREADS) {
      float acc[RNN_BPC];
#pragma unroll
      for (int b = 0; b < RNN_BPC; ++b) acc[b] = 0.f;
      for (int k = 0; k < H; ++k) {
        float w = __ldg(whhT + (int64_t)k * GH + n);
#pragma unroll
        for (int b = 0; b < RNN_BPC; ++b) acc[b] = fmaf(w, h_s[b * H + k], acc[b]);
      }
      float bh = (KIND == 1) ? __ldg(b_hh + n) : 0.f;
#pragma unroll
      for (int b = 0; b < RNN_BPC; ++b) pre[b * GH + n] = acc[b] + bh;
    }
    __syncthreads();
    for (int i = tid; i < RNN_BPC * H; i += RNN_THREADS) {
      int bl = i / H, j = i - bl * H, b = b0 + bl;
      if (b >= B || s >= len_s[bl]) continue;
      const float* xps = xp + ((int64_t)b * S + s) * GH;
      float* gs = gates + ((int64_t)b * S + s) * GH;
      const float* pr = pre + bl * GH;
      int64_t o = ((int64_t)b * S + s) * H + j;
      if (KIND == 0) {
        float gi = sigmoidf_(xps[j] + pr[j]);
        float gf = sigmoidf_(xps[H + j] + pr[H + j]);
        float gg = tanhf(xps[2 * H + j] + pr[2 * H + j]);
        float go = sigmoidf_(xps[3 * H + j] + pr[3 * H + j]);
        float c = gf * c_s[i] + gi * gg;
        float h = go * tanhf(c);
        gs[j] = gi; gs[H + j] = gf; gs[2 * H + j] = gg; gs[3 * H + j] = go;
        c_s[i] = c; h_s[i] = h;
        cs[o] = c; hs[o] = h;
      } else {
        float r = sigmoidf_(xps[j] + pr[j]);
        float z = sigmoidf_(xps[H + j] + pr[H + j]);
        float hn = pr[2 * H + j];
        float nn = tanhf(xps[2 * H + j] + r * hn);
        float h = (1.f - z) * nn + z * h_s[i];
        gs[j] = r; gs[H + j] = z; gs[2 * H + j] = nn;
        h_s[i] = h;
        cs[o] = hn; hs[o] = h;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < RNN_BPC * H; i += RNN_THREADS) {
    int b = b0 + i / H;
    if (b < B) user[(int64_t)b * H + (i % H)] = h_s[i];
  }
}

template <int KIND>
__global__ void __launch_bounds__(RNN_THREADS)
rnn_bwd_kernel(const float* __restrict__ whh,     // [G*H, H]
               const float* __restrict__ h0, const int32_t* __restrict__ lens, const float* __restrict__ gates,
               const float* __restrict__ hs, const float* __restrict__ cs, const float* __restrict__ d_user,
               float* __restrict__ dgi,           // [B,S,G*H] grad wrt input-side pre-activations
               float* __restrict__ dgh,           // [B,S,G*H] grad wrt hidden-side pre-activations (GRU; == dgi for LSTM)
               float* __restrict__ d_h0,          // [B,H] or null
               int B, int S, int H) {
  constexpr int G = KIND == 0 ? 4 : 3;
  extern __shared__ float smem[];
  const int GH = G * H;
  float* dh_s = smem;                     // [BPC][H] carried dL/dh
  float* dc_s = dh_s + RNN_BPC * H;       // [BPC][H] carried dL/dc (LSTM) / direct dh*z term (GRU)
  float* dp = dc_s + RNN_BPC * H;         // [BPC][GH] hidden-side pre-activation grads of this step
  __shared__ int len_s[RNN_BPC];
  const int b0 = blockIdx.x * RNN_BPC;
  const int tid = threadIdx.x;
  for (int i = tid; i < RNN_BPC * H; i += RNN_THREADS) { dh_s[i] = 0.f; dc_s[i] = 0.f; }
  if (tid < RNN_BPC) {
    int b = b0 + tid;
    int l = (b < B) ? (lens ? lens[b] : S) : 0;
    len_s[tid] = l < 0 ? 0 : (l > S ? S : l);
  }
  __syncthreads();
  int max_len = 0;
#pragma unroll
  for (int i = 0; i < RNN_BPC; ++i) max_len = max(max_len, len_s[i]);
  // steps beyond a sequence's length contribute nothing: zero their gate gradients
  for (int bl = 0; bl < RNN_BPC; ++bl) {
    int b = b0 + bl;
    if (b >= B) continue;
    for (int64_t i = (int64_t)len_s[bl] * GH + tid; i < (int64_t)S * GH; i += RNN_THREADS) {
      dgi[(int64_t)b * S * GH + i] = 0.f;
      if (KIND == 1) dgh[(int64_t)b * S * GH + i] = 0.f;
    }
  }

  for (int s = max_len - 1; s >= 0; --s) {
    for (int i = tid; i < RNN_BPC * H; i += RNN_THREADS) {
      int bl = i / H, j = i - bl * H, b = b0 + bl;
      float* dpl = dp + bl * GH;
      if (b >= B || s >= len_s[bl]) {
        for (int g = 0; g < G; ++g) dpl[g * H + j] = 0.f;
        continue;
      }
      float dh = dh_s[i];
      if (s == len_s[bl] - 1) dh += d_user[(int64_t)b * H + j];
      const float* gs = gates + ((int64_t)b * S + s) * GH;
      int64_t o = ((int64_t)b * S + s) * H + j;
      float* gi_out = dgi + ((int64_t)b * S + s) * GH;
      if (KIND == 0) {
        float gi = gs[j], gf = gs[H + j], gg = gs[2 * H + j], go = gs[3 * H + j];
        float c = cs[o];
        float cprev = s > 0 ? cs[o - H] : 0.f;
        float tc = tanhf(c);
        float dc = dc_s[i] + dh * go * (1.f - tc * tc);
        float d_i = dc * gg * gi * (1.f - gi);
        float d_f = dc * cprev * gf * (1.f - gf);
        float d_g = dc * gi * (1.f - gg * gg);
        float d_o = dh * tc * go * (1.f - go);
        dc_s[i] = dc * gf;
        dpl[j] = d_i; dpl[H + j] = d_f; dpl[2 * H + j] = d_g; dpl[3 * H + j] = d_o;
        gi_out[j] = d_i; gi_out[H + j] = d_f; gi_out[2 * H + j] = d_g; gi_out[3 * H + j] = d_o;
      } else {
        float r = gs[j], z = gs[H + j], nn = gs[2 * H + j];
        float hn = cs[o];
        float hprev = s > 0 ? hs[o - H] : (h0 ? h0[(int64_t)b * H + j] : 0.f);
        float dn = dh * (1.f - z) * (1.f - nn * nn);
        float dz = dh * (hprev - nn) * z * (1.f - z);
        float dr = dn * hn * r * (1.f - r);
        dc_s[i] = dh * z;
        gi_out[j] = dr; gi_out[H + j] = dz; gi_out[2 * H + j] = dn;
        float* gh_out = dgh + ((int64_t)b * S + s) * GH;
        gh_out[j] = dr; gh_out[H + j] = dz; gh_out[2 * H + j] = dn * r;
        dpl[j] = dr; dpl[H + j] = dz; dpl[2 * H + j] = dn * r;
      }
    }
    __syncthreads();
    // dh_{s-1} = dp W_hh (+ direct z path for the GRU)
    for (int i = tid; i < RNN_BPC * H; i += RNN_THREADS) {
      int bl = i / H, k = i - bl * H;
      const float* dpl = dp + bl * GH;
      float acc = (KIND == 1) ? dc_s[i] : 0.f;
      if (b0 + bl < B && s < len_s[bl]) {
        for (int n = 0; n < GH; ++n) acc = fmaf(dpl[n], __ldg(whh + (int64_t)n * H + k), acc);
        dh_s[i] = acc;
      }
    }
    __syncthreads();
  }
  if (d_h0)
    for (int i = tid; i < RNN_BPC * H; i += RNN_THREADS) {
      int b = b0 + i / H;
      if (b < B) d_h0[(int64_t)b * H + (i % H)] = dh_s[i];
    }
}

// x[b, pos(s), :] viewed as a [B*S, H] matrix indexed by (b*S+s)
struct SeqView {
  const float* x; int S, H, reverse;
  __device__ __forceinline__ int64_t row(int64_t m) const {
    if (!reverse) return m;
    int64_t b = m / S; int s = (int)(m - b * S);
    return b * S + (S - 1 - s);
  }
  __device__ __forceinline__ float operator()(int64_t m, int64_t k) const { return __ldg(x + row(m) * H + k); }
};
struct SeqViewKM {           // as a B operand: (k=m, n)
  SeqView v;
  __device__ __forceinline__ float operator()(int64_t k, int64_t n) const { return v(k, n); }
};
struct PrevHiddenKM {        // (k=(b,s), n) -> h_{s-1}[b, n]  (h0 or 0 at s = 0)
  const float* hs; const float* h0; int S, H;
  __device__ __forceinline__ float operator()(int64_t k, int64_t n) const {
    int64_t b = k / S; int s = (int)(k - b * S);
    if (s == 0) return h0 ? __ldg(h0 + b * H + n) : 0.f;
    return __ldg(hs + (k - 1) * H + n);
  }
};
struct SeqStoreEpi {         // d_x[b, pos(s), n] = v
  float* out; int S, H, reverse;
  __device__ __forceinline__ void operator()(int64_t m, int64_t n, float v) const {
    int64_t r = m;
    if (reverse) { int64_t b = m / S; int s = (int)(m - b * S); r = b * S + (S - 1 - s); }
    out[r * H + n] = v;
  }
};
struct TwoBiasEpi {
  float* out; int64_t ld; const float* b1; const float* b2;
  __device__ __forceinline__ void operator()(int64_t m, int64_t n, float v) const {
    v += __ldg(b1 + n);
    if (b2) v += __ldg(b2 + n);
    out[m * ld + n] = v;
  }
};

// ---- MR_BF16: the three big GEMMs of the recurrent encoders on the tensor-core kernels -------------------------
// rows m = b*S + s (step order: s-th step reads x[b, S-1-s] when `reverse`), padded to a multiple of 128 rows
__global__ void seq_cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t M, int64_t Mp, int S,
                                     int H, int Hp, int reverse) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp * Hp) return;
  const int64_t m = i / Hp;
  const int k = (int)(i - m * Hp);
  float v = 0.f;
  if (m < M && k < H) {
    const int64_t b = m / S;
    const int s = (int)(m - b * S);
    v = x[(b * S + (reverse ? S - 1 - s : s)) * H + k];
  }
  out[i] = __float2bfloat16(v);
}
// h_{s-1} of every step (h0 or 0 at s = 0)
// (steps at or beyond a sequence's length give zero rows: the recurrence kernels never write hs there)
__global__ void hprev_bf16_kernel(const float* __restrict__ hs, const float* __restrict__ h0, const int32_t* __restrict__ lens,
                                  __nv_bfloat16* __restrict__ out, int64_t M, int64_t Mp, int S, int H, int Hp) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp * Hp) return;
  const int64_t m = i / Hp;
  const int k = (int)(i - m * Hp);
  float v = 0.f;
  if (m < M && k < H) {
    const int64_t b = m / S;
    const int s = (int)(m - b * S);
    const int len = lens ? lens[b] : S;
    if (s < len) v = s > 0 ? hs[(m - 1) * H + k] : (h0 ? h0[b * H + k] : 0.f);
  }
  out[i] = __float2bfloat16(v);
}
__global__ void cast_rows_pad_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t M, int64_t Mp,
                                          int C, int Cp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mp * Cp) return;
  const int64_t m = i / Cp;
  const int c = (int)(i - m * Cp);
  dst[i] = __float2bfloat16((m < M && c < C) ? src[m * C + c] : 0.f);
}
// d_x[b, pos(s), :] = tmp[(b,s), :H]
__global__ void unseq_copy_kernel(const float* __restrict__ tmp, float* __restrict__ d_x, int64_t M, int S, int H, int Hp, int reverse) {
  pdl_trigger();
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * H) return;
  const int64_t m = i / H;
  const int k = (int)(i - m * H);
  const int64_t b = m / S;
  const int s = (int)(m - b * S);
  d_x[(b * S + (reverse ? S - 1 - s : s)) * H + k] = tmp[m * Hp + k];
}

struct RnnTcGeom {
  int64_t M, Mp, Hp, GH, GHp, nblk, nbsz;
};
static RnnTcGeom rnn_tc_geom(const mr_rnn_shape* s) {
  RnnTcGeom g;
  g.M = s->B * s->S;
  g.Mp = align_up(g.M, 128);
  g.Hp = align_up(s->H, 16);
  g.GH = (s->kind == MR_RNN_LSTM ? 4 : 3) * s->H;
  g.GHp = align_up(g.GH, 16);
  g.nblk = ceil_div(g.GHp, 256);
  g.nbsz = align_up(ceil_div(g.GHp, g.nblk), 16);
  return g;
}
static bool rnn_tc_ok(const mr_rnn_shape* s) { return s->precision == MR_BF16 && align_up(s->H, 16) <= 256; }
static int64_t rnn_tc_ws(const mr_rnn_shape* s, int backward) {
  const RnnTcGeom g = rnn_tc_geom(s);
  int64_t b = arena_bytes(g.Mp * g.Hp, 2);                                      // x as bf16, step order
  if (!backward) return b + arena_bytes(tapgemm_pack_bytes(1, (int)g.nbsz, (int)g.Hp), 1);
  b += arena_bytes(g.Mp * g.Hp, 2);                                             // h_{s-1}
  b += 2 * arena_bytes(g.Mp * g.GHp, 2);                                        // dgi, dgh as bf16
  b += arena_bytes(tapgemm_pack_bytes(1, (int)g.Hp, (int)g.GHp), 1);            // W_ih for d_x
  b += arena_bytes(g.M * g.Hp, 4);                                              // d_x in step order
  b += arena_bytes(tokred_partial_bytes(g.Mp / 128, 128, 1, (int)g.GHp, (int)g.Hp), 1);
  return b;
}

// xp = x_seq W_ih^T + b_ih (+ b_hh)
static int rnn_tc_input_proj(const mr_rnn_shape* s, const float* x, const float* w_ih, const float* b_ih, const float* b_hh2,
                             float* xp, Arena& ar, cudaStream_t st) {
  const RnnTcGeom g = rnn_tc_geom(s);
  const int H = (int)s->H;
  __nv_bfloat16* xb = ar.take<__nv_bfloat16>(g.Mp * g.Hp);
  uint8_t* wp = ar.take<uint8_t>(tapgemm_pack_bytes(1, (int)g.nbsz, (int)g.Hp));
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_rnn_user_fwd: workspace too small");
  launch_pdl(seq_cast_bf16_kernel, dim3((unsigned)ceil_div(g.Mp * g.Hp, 256)), dim3(256), 0, st, x, xb, g.M, g.Mp, (int)s->S, H, (int)g.Hp, s->reverse);
  MR_CHECK_LAUNCH("seq_cast_bf16_kernel");
  for (int64_t blk = 0; blk < g.nblk; ++blk) {
    const int64_t n0 = blk * g.nbsz;
    const int64_t nb = (g.GHp - n0) < g.nbsz ? (g.GHp - n0) : g.nbsz;
    const int64_t nv = (g.GH - n0) < nb ? (g.GH - n0) : nb;
    if (nv <= 0) break;
    if (int rc = tapgemm_pack(w_ih + n0 * H, wp, 1, (int)nb, (int)g.Hp, (int)nv, H, H, 1, 0, st)) return rc;
    TapGemmArgs a{};
    TapGemmPlan plan;
    a.n_titles = g.Mp / 128; a.L = 128; a.taps = 1; a.dir = 1; a.K = (int)g.Hp;
    a.n_sub = 1; a.nsz[0] = (int)nb;
    a.ids = nullptr; a.a = xb; a.lda = g.Hp;
    a.wpack = wp; a.epi = TG_EPI_BIAS_F32; a.bias = b_ih + n0; a.bias2 = b_hh2 ? b_hh2 + n0 : nullptr; a.n_valid = (int)nv;
    a.n_rows = g.M; a.out_f32 = xp + n0; a.ldo = align_up(g.GH, 4); a.n_store = (int)align_up(nv, 4);
    if (int rc = tapgemm_plan(a, &plan)) return rc;
    if (int rc = tapgemm_launch(plan, st)) return rc;
  }
  return MR_OK;
}

// d_x, d_w_ih, d_w_hh from the gate gradients
static int rnn_tc_grad_gemms(const mr_rnn_shape* s, const float* x, const float* h0, const int32_t* lens, const float* w_ih, const float* hs,
                             const float* dgi, const float* dgh, float* d_x, float* d_w_ih, float* d_w_hh, Arena& ar,
                             cudaStream_t st, __nv_bfloat16* gib, __nv_bfloat16* ghb, bool precast) {
  // gib / ghb: bf16 [Mp, GHp] gate gradients; `precast`: already written by the recurrence kernel (else cast from dgi / dgh here)
  const RnnTcGeom g = rnn_tc_geom(s);
  const int H = (int)s->H, S = (int)s->S;
  __nv_bfloat16* xb = ar.take<__nv_bfloat16>(g.Mp * g.Hp);
  __nv_bfloat16* hb = ar.take<__nv_bfloat16>(g.Mp * g.Hp);
  uint8_t* wp = ar.take<uint8_t>(tapgemm_pack_bytes(1, (int)g.Hp, (int)g.GHp));
  float* tmp = ar.take<float>(g.M * g.Hp);
  float* partial = ar.take<float>(tokred_partial_bytes(g.Mp / 128, 128, 1, (int)g.GHp, (int)g.Hp) / 4);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_rnn_user_bwd: workspace too small");
  launch_pdl(seq_cast_bf16_kernel, dim3((unsigned)ceil_div(g.Mp * g.Hp, 256)), dim3(256), 0, st, x, xb, g.M, g.Mp, S, H, (int)g.Hp, s->reverse);
  MR_CHECK_LAUNCH("seq_cast_bf16_kernel");
  launch_pdl(hprev_bf16_kernel, dim3((unsigned)ceil_div(g.Mp * g.Hp, 256)), dim3(256), 0, st, hs, h0, lens, hb, g.M, g.Mp, S, H, (int)g.Hp);
  MR_CHECK_LAUNCH("hprev_bf16_kernel");
  if (!precast) {
    cast_rows_pad_bf16_kernel<<<(unsigned)ceil_div(g.Mp * g.GHp, 256), 256, 0, st>>>(dgi, gib, g.M, g.Mp, (int)g.GH, (int)g.GHp);
    MR_CHECK_LAUNCH("cast_rows_pad_bf16_kernel");
    if (dgh != dgi) {
      cast_rows_pad_bf16_kernel<<<(unsigned)ceil_div(g.Mp * g.GHp, 256), 256, 0, st>>>(dgh, ghb, g.M, g.Mp, (int)g.GH, (int)g.GHp);
      MR_CHECK_LAUNCH("cast_rows_pad_bf16_kernel");
    }
  }
  if (s->kind == MR_RNN_LSTM) ghb = gib;
  if (d_x) {        // d_x_seq[m, k] = sum_n dgi[m, n] W_ih[n, k]
    if (int rc = tapgemm_pack(w_ih, wp, 1, (int)g.Hp, (int)g.GHp, H, (int)g.GH, 1, H, 0, st)) return rc;
    TapGemmArgs a{};
    TapGemmPlan plan;
    a.n_titles = g.Mp / 128; a.L = 128; a.taps = 1; a.dir = 1; a.K = (int)g.GHp;
    a.n_sub = 1; a.nsz[0] = (int)g.Hp;
    a.ids = nullptr; a.a = gib; a.lda = g.GHp;
    a.wpack = wp; a.epi = TG_EPI_BIAS_F32; a.bias = nullptr; a.bias2 = nullptr; a.n_valid = H;
    a.n_rows = g.M; a.out_f32 = tmp; a.ldo = g.Hp; a.n_store = (int)g.Hp;
    if (int rc = tapgemm_plan(a, &plan)) return rc;
    if (int rc = tapgemm_launch(plan, st)) return rc;
    launch_pdl(unseq_copy_kernel, dim3((unsigned)ceil_div(g.M * H, 256)), dim3(256), 0, st, tmp, d_x, g.M, S, H, (int)g.Hp, s->reverse);
    MR_CHECK_LAUNCH("unseq_copy_kernel");
  }
  for (int which = 0; which < 2; ++which) {   // d_w_ih[n,k] = sum_m dgi[m,n] x[m,k];  d_w_hh[n,k] = sum_m dgh[m,n] h_{s-1}[m,k]
    TokRedArgs a{};
    TokRedPlan plan;
    a.n_titles = g.Mp / 128; a.L = 128; a.taps = 1;
    a.ids = nullptr; a.p = which == 0 ? gib : ghb; a.ldp = g.GHp; a.KP = (int)g.GHp;
    a.q = which == 0 ? xb : hb; a.ldq = g.Hp; a.NQ = (int)g.Hp; a.partial = partial;
    if (int rc = tokred_plan(a, &plan)) return rc;
    if (int rc = tokred_launch(plan, st)) return rc;
    if (int rc = tokred_reduce(plan, which == 0 ? d_w_ih : d_w_hh, (int)g.GH, H, 1, H, 0, st)) return rc;
  }
  return MR_OK;
}

static int64_t rnn_ws(const mr_rnn_shape* s, int backward) {
  const int64_t G = s->kind == MR_RNN_LSTM ? 4 : 3, GH = G * s->H, BS = s->B * s->S;
  int64_t b = 0;
  if (!backward) {
    b += arena_bytes(BS * align_up(GH, 4), 4);   // xp
    b += arena_bytes(s->H * GH + rnn_res_scratch_bytes(s->kind, (int)s->H) / 4 + rnn_mma_scratch_bytes(s->kind, (int)s->H) / 4 +
                     rnn_tc_scratch_bytes(s->kind, (int)s->H) / 4, 4);   // whhT / bf16 images of W_hh
    if (rnn_tc_ok(s)) b += rnn_tc_ws(s, 0);
    return b + 256;
  }
  b += 2 * arena_bytes(BS * GH, 4);         // dgi, dgh
  b += arena_bytes((int64_t)s->B * 2 * GH, 4);   // per-CTA bias-gradient partials of the resident recurrence kernel
  b += arena_bytes(64 * GH * s->H, 4);      // split-K partial
  b += arena_bytes(colsum_chunks(BS) * GH, 4);
  if (rnn_tc_ok(s)) b += rnn_tc_ws(s, 1);
  return b + 256;
}

}  // namespace mr

// MINDREC_RNN_MMA=0 keeps the SIMT resident-weights recurrence (A/B)
static bool rnn_use_mma() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MINDREC_RNN_MMA");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

extern "C" {

int64_t mr_rnn_workspace_bytes(const mr_rnn_shape* s, int backward) {
  if (!s) return -1;
  return mr::rnn_ws(s, backward);
}

static int rnn_check(const mr_rnn_shape* s, const char* who) {
  using namespace mr;
  MR_REQUIRE(s != nullptr, MR_ERR_NULL, "%s: null shape", who);
  MR_REQUIRE(s->B >= 0 && s->S >= 1 && s->H >= 1, MR_ERR_BAD_SHAPE, "%s: B=%lld S=%lld H=%lld", who, (long long)s->B,
             (long long)s->S, (long long)s->H);
  MR_REQUIRE(s->kind == MR_RNN_LSTM || s->kind == MR_RNN_GRU, MR_ERR_UNSUPPORTED, "%s: kind %d", who, s->kind);
  MR_REQUIRE(s->H <= 1024, MR_ERR_UNSUPPORTED, "%s: hidden_dim %lld > 1024", who, (long long)s->H);
  return MR_OK;
}

int mr_rnn_user_fwd(const mr_rnn_shape* s, const float* x, const int32_t* lens, const float* h0, const float* w_ih,
                    const float* w_hh, const float* b_ih, const float* b_hh, float* gates, float* hs, float* cs,
                    float* user, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  if (int rc = rnn_check(s, "mr_rnn_user_fwd")) return rc;
  MR_REQUIRE(x && w_ih && w_hh && b_ih && b_hh && gates && hs && cs && user, MR_ERR_NULL, "mr_rnn_user_fwd: null pointer");
  if (s->B == 0) return MR_OK;
  cudaStream_t st = as_stream(stream);
  const int B = (int)s->B, S = (int)s->S, H = (int)s->H;
  const int G = s->kind == MR_RNN_LSTM ? 4 : 3, GH = G * H;
  Arena ar(workspace, workspace_bytes);
  const int64_t ldx = rnn_tc_ok(s) ? align_up(GH, 4) : GH;       // pitch of the input projection (16-byte rows for the tensor-core epilogue)
  float* xp = ar.take<float>((int64_t)B * S * ldx);
  float* whhT = ar.take<float>((int64_t)H * GH + rnn_res_scratch_bytes(s->kind, H) / 4 + rnn_mma_scratch_bytes(s->kind, H) / 4 +
                               rnn_tc_scratch_bytes(s->kind, H) / 4);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_rnn_user_fwd: workspace too small (%lld given)", (long long)workspace_bytes);
  if (rnn_tc_ok(s)) {
    if (int rc = rnn_tc_input_proj(s, x, w_ih, b_ih, s->kind == MR_RNN_LSTM ? b_hh : nullptr, xp, ar, st)) return rc;
  } else {
    SeqView xv{x, S, H, s->reverse};
    cudaError_t e = gemm_simt<true, false>((int64_t)B * S, GH, H, xv, Transposed{w_ih, H},
                                           TwoBiasEpi{xp, GH, b_ih, s->kind == MR_RNN_LSTM ? b_hh : nullptr}, 1, nullptr, st);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "rnn input projection: %s", cudaGetErrorString(e));
  }
  // steps past a sequence's length are never written by the kernels.  The bf16 path's backward never reads them either
  // (resident recurrence: s < len only; h_{s-1} operand: zero rows beyond len), so only the other paths clear them.
  const bool resident_path = s->precision == MR_BF16 && rnn_tc_ok(s) && rnn_res_supported(s->kind, H);
  if (!resident_path) {
    cudaMemsetAsync(gates, 0, sizeof(float) * (int64_t)B * S * GH, st);
    cudaMemsetAsync(hs, 0, sizeof(float) * (int64_t)B * S * H, st);
    cudaMemsetAsync(cs, 0, sizeof(float) * (int64_t)B * S * H, st);
  }
  if (s->precision == MR_BF16 && resident_path && rnn_tc_supported(s->kind, H))
    return rnn_tc_fwd(s->kind, xp, (int)ldx, w_hh, h0, lens, gates, hs, cs, user, B, S, H, whhT, st);
  if (s->precision == MR_BF16 && rnn_mma_supported(s->kind, H) && rnn_use_mma())
    return rnn_mma_fwd(s->kind, xp, (int)ldx, w_hh, b_hh, h0, lens, gates, hs, cs, user, B, S, H, whhT, st);
  if (s->precision == MR_BF16 && rnn_res_supported(s->kind, H))
    return rnn_res_fwd(s->kind, xp, (int)ldx, w_hh, b_hh, h0, lens, gates, hs, cs, user, B, S, H, whhT, st);
  dim3 tg((unsigned)ceil_div(H, 32), (unsigned)ceil_div(GH, 32)), tb(32, 8);
  transpose_kernel<<<tg, tb, 0, st>>>(w_hh, whhT, GH, H);
  MR_CHECK_LAUNCH("transpose_kernel");
  size_t smem = sizeof(float) * (2 * RNN_BPC * H + RNN_BPC * GH);
  unsigned grid = (unsigned)ceil_div(B, RNN_BPC);
  if (s->kind == MR_RNN_LSTM) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(rnn_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rnn_fwd_kernel<0><<<grid, RNN_THREADS, smem, st>>>(xp, whhT, b_hh, h0, lens, gates, hs, cs, user, B, S, H);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(rnn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rnn_fwd_kernel<1><<<grid, RNN_THREADS, smem, st>>>(xp, whhT, b_hh, h0, lens, gates, hs, cs, user, B, S, H);
  }
  MR_CHECK_LAUNCH("rnn_fwd_kernel");
  return MR_OK;
}

int mr_rnn_user_bwd(const mr_rnn_shape* s, const float* x, const int32_t* lens, const float* h0, const float* w_ih,
                    const float* w_hh, const float* gates, const float* hs, const float* cs, const float* d_user,
                    float* d_x, float* d_h0, float* d_w_ih, float* d_w_hh, float* d_b_ih, float* d_b_hh, void* workspace,
                    int64_t workspace_bytes, void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  if (int rc = rnn_check(s, "mr_rnn_user_bwd")) return rc;
  MR_REQUIRE(x && w_ih && w_hh && gates && hs && cs && d_user && d_w_ih && d_w_hh && d_b_ih && d_b_hh, MR_ERR_NULL,
             "mr_rnn_user_bwd: null pointer");
  cudaStream_t st = as_stream(stream);
  const int B = (int)s->B, S = (int)s->S, H = (int)s->H;
  const int G = s->kind == MR_RNN_LSTM ? 4 : 3, GH = G * H;
  if (B == 0) {
    cudaMemsetAsync(d_w_ih, 0, sizeof(float) * GH * H, st);
    cudaMemsetAsync(d_w_hh, 0, sizeof(float) * GH * H, st);
    cudaMemsetAsync(d_b_ih, 0, sizeof(float) * GH, st);
    cudaMemsetAsync(d_b_hh, 0, sizeof(float) * GH, st);
    return MR_OK;
  }
  const int64_t BS = (int64_t)B * S;
  Arena ar(workspace, workspace_bytes);
  float* dgi = ar.take<float>(BS * GH);
  float* dgh = ar.take<float>(BS * GH);
  float* sp = ar.take<float>((int64_t)64 * GH * H);
  float* cp = ar.take<float>(colsum_chunks(BS) * GH);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_rnn_user_bwd: workspace too small (%lld given)", (long long)workspace_bytes);
  if (s->kind == MR_RNN_LSTM) dgh = dgi;
  size_t smem = sizeof(float) * (2 * RNN_BPC * H + RNN_BPC * GH);
  unsigned grid = (unsigned)ceil_div(B, RNN_BPC);
  const bool resident = s->precision == MR_BF16 && rnn_res_supported(s->kind, H);
  // resident recurrence + tensor-core GEMMs: the kernel writes the gate gradients as bf16 GEMM operands and the bias
  // partial sums itself (no fp32 dgi / dgh round trip, no cast and column-sum passes)
  const bool fused_out = resident && rnn_tc_ok(s);
  __nv_bfloat16* gib = nullptr;
  __nv_bfloat16* ghb = nullptr;
  float* bias_part = nullptr;
  RnnTcGeom tg{};
  if (rnn_tc_ok(s)) {
    tg = rnn_tc_geom(s);
    gib = ar.take<__nv_bfloat16>(tg.Mp * tg.GHp);
    ghb = ar.take<__nv_bfloat16>(tg.Mp * tg.GHp);
    bias_part = ar.take<float>((int64_t)B * 2 * GH);
    MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_rnn_user_bwd: workspace too small (%lld given)", (long long)workspace_bytes);
  }
  // LSTM: the reverse recurrence as tcgen05 batches with W_hh^T resident in tensor memory (rnn_tc.cu)
  const bool tc_bwd = fused_out && rnn_tc_bwd_supported(s->kind, H, B);
  if (resident) {
    if (fused_out) {
      cudaMemsetAsync(gib, 0, (size_t)tg.Mp * tg.GHp * 2, st);
      if (s->kind != MR_RNN_LSTM) cudaMemsetAsync(ghb, 0, (size_t)tg.Mp * tg.GHp * 2, st);
    }
    if (tc_bwd) {
      if (int rc = rnn_tc_bwd(s->kind, w_hh, lens, gates, cs, d_user, gib, (int)tg.GHp, d_h0, bias_part, B, S, H, sp, st)) return rc;
    } else if (int rc = rnn_res_bwd(s->kind, w_hh, h0, lens, gates, hs, cs, d_user, dgi, dgh, d_h0, B, S, H, sp, st,
                                    fused_out ? gib : nullptr, fused_out ? ghb : nullptr, fused_out ? (int)tg.GHp : 0,
                                    fused_out ? bias_part : nullptr))
      return rc;
  } else if (s->kind == MR_RNN_LSTM) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(rnn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rnn_bwd_kernel<0><<<grid, RNN_THREADS, smem, st>>>(w_hh, h0, lens, gates, hs, cs, d_user, dgi, dgh, d_h0, B, S, H);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(rnn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rnn_bwd_kernel<1><<<grid, RNN_THREADS, smem, st>>>(w_hh, h0, lens, gates, hs, cs, d_user, dgi, dgh, d_h0, B, S, H);
  }
  if (!resident) MR_CHECK_LAUNCH("rnn_bwd_kernel");
  cudaError_t e;
  if (rnn_tc_ok(s)) {
    if (int rc = rnn_tc_grad_gemms(s, x, h0, lens, w_ih, hs, dgi, dgh, d_x, d_w_ih, d_w_hh, ar, st, gib, ghb, fused_out)) return rc;
  } else {
    if (d_x) {
      e = gemm_simt<true, true>(BS, H, GH, RowMajor{dgi, GH}, RowMajor{w_ih, H}, SeqStoreEpi{d_x, S, H, s->reverse}, 1, nullptr, st);
      MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "rnn d_x: %s", cudaGetErrorString(e));
    }
    e = gemm_simt<false, true>(GH, H, BS, Transposed{dgi, GH}, SeqViewKM{SeqView{x, S, H, s->reverse}}, StoreEpi{d_w_ih, H},
                               pick_splits(GH, H, BS), sp, st);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "rnn d_w_ih: %s", cudaGetErrorString(e));
    e = gemm_simt<false, true>(GH, H, BS, Transposed{dgh, GH}, PrevHiddenKM{hs, h0, S, H}, StoreEpi{d_w_hh, H},
                               pick_splits(GH, H, BS), sp, st);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "rnn d_w_hh: %s", cudaGetErrorString(e));
  }
  if (fused_out) {
    // bias gradients: fold the per-CTA partials (pitch 2*GH: [0] input side, [1] hidden side) in a fixed order
    const int64_t rows = tc_bwd ? rnn_tc_bwd_rows(B) : ceil_div(B, rnn_res_bpc(s->kind, B, H));
    e = colsum_small(bias_part, 2 * GH, d_b_ih, rows, GH, st);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "rnn d_b_ih: %s", cudaGetErrorString(e));
    e = colsum_small(bias_part + (s->kind == MR_RNN_LSTM ? 0 : GH), 2 * GH, d_b_hh, rows, GH, st);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "rnn d_b_hh: %s", cudaGetErrorString(e));
    return MR_OK;
  }
  e = colsum(dgi, d_b_ih, BS, GH, cp, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "rnn d_b_ih: %s", cudaGetErrorString(e));
  e = colsum(dgh, d_b_hh, BS, GH, cp, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "rnn d_b_hh: %s", cudaGetErrorString(e));
  return MR_OK;
}

}  // extern "C"
