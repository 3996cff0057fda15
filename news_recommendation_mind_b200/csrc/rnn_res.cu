// Persistent LSTM / GRU recurrence with resident recurrent weights (MR_BF16 path).
// Reference: models/Encoders/RNN.py:36-73 (pack_padded_sequence + nn.LSTM/nn.GRU, h_n) and :76-104 (LSTUR).
//
// One CTA owns BPC sequences for all S steps; there is no inter-CTA communication and no per-step launch.
//   * W_hh is converted to bf16 ONCE per launch and stays in shared memory (LSTM H=150: 600x150x2 B = 180 KB);
//     the hidden / cell state, the gate pre-activations and all activations are fp32.
//   * step = (1) matvec: 600 threads, each owns 8 consecutive gate columns (one 16-byte shared load per k)
//              and one of 8 k-slices, fp32 FMAs against the broadcast hidden state, partial sums to smem;
//            (2) gates: one thread per (sequence, hidden unit) folds the 8 partials, applies the
//              nonlinearities, updates c/h (only while s < len: the packed-sequence semantics), stores the
//              saved tensors.  The input projection of the step is prefetched into registers before (1).
//   * backward: same structure in reverse time with W_hh [n][k]-major in shared memory:
//            (A) gate gradients from the carried dh/dc, (B) dh_prev = dgates . W_hh (16 n-slices), (C) fold.
// BPC = ceil(B / #SM) rounded up to 1/2/4, so B=256 runs as 128 CTAs of 2 sequences.
#include "rnn_res.cuh"
#include "tapgemm.cuh"   // sm_count()
#include "tc05.cuh"

namespace mr {

constexpr int RR_THREADS = 640;
constexpr int RR_KS = 8;      // k-slices of the forward matvec
constexpr int RR_NS = 16;     // n-slices of the backward matvec

__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + __expf(-x)); }

struct RRGeom {
  int G, GH, NT, GHp, KR, Hp, NTk, NR;
  size_t fwd_bytes, bwd_bytes;
};
static inline RRGeom rr_geom(int kind, int H, int bpc) {
  RRGeom g;
  g.G = kind == MR_RNN_LSTM ? 4 : 3;
  g.GH = g.G * H;
  g.NT = (g.GH + 7) / 8;
  g.GHp = g.NT * 8;
  g.KR = (H + RR_KS - 1) / RR_KS;
  g.Hp = (H + 7) / 8 * 8;
  g.NTk = g.Hp / 8;
  g.NR = (g.GH + RR_NS - 1) / RR_NS;
  g.fwd_bytes = (size_t)H * g.GHp * 2 + (size_t)RR_KS * bpc * g.GHp * 4 + (size_t)2 * bpc * H * 4 + 64;
  g.bwd_bytes = (size_t)g.GH * g.Hp * 2 + (size_t)RR_NS * bpc * g.Hp * 4 + (size_t)bpc * g.GH * 4 + 64;
  return g;
}

static bool rr_fits(int kind, int H, int bpc) {
  RRGeom g = rr_geom(kind, H, bpc);
  const size_t lim = 227 * 1024;
  return g.fwd_bytes <= lim && g.bwd_bytes <= lim && bpc * H <= RR_THREADS && g.NT * RR_KS <= RR_THREADS &&
         g.NTk * RR_NS <= RR_THREADS;
}

// sequences per CTA: enough to cover the batch with one wave of CTAs when the scratch for it still fits
int rnn_res_bpc(int kind, int B, int H) {
  const int per = (B + sm_count() - 1) / sm_count();
  int bpc = per <= 1 ? 1 : (per <= 2 ? 2 : 4);
  while (bpc > 1 && !rr_fits(kind, H, bpc)) bpc >>= 1;
  return bpc;
}

bool rnn_res_supported(int kind, int H) { return rr_fits(kind, H, 1); }

// W_hh [GH, H] fp32 -> bf16 images of the two shared-memory layouts (done once per call; every CTA then pulls its
// copy with a few cp.async.bulk transfers instead of 90k strided loads):
//   wt [H][GHp]  (forward:  Wt[k][n] = W_hh[n][k], zero padded columns)      w [GH][Hp]  (backward, zero padded)
__global__ void rnn_res_prep_kernel(const float* __restrict__ w_hh, __nv_bfloat16* __restrict__ wt, __nv_bfloat16* __restrict__ w,
                                    int GH, int H, int GHp, int Hp) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (wt != nullptr && i < H * GHp) {
    const int k = i / GHp, n = i - k * GHp;
    wt[i] = __float2bfloat16(n < GH ? w_hh[(int64_t)n * H + k] : 0.f);
  }
  if (w != nullptr && i < GH * Hp) {
    const int n = i / Hp, k = i - n * Hp;
    w[i] = __float2bfloat16(k < H ? w_hh[(int64_t)n * H + k] : 0.f);
  }
}

// pulls `bytes` (multiple of 16) from global into shared memory with bulk copies; all threads wait on the barrier
__device__ __forceinline__ void bulk_fill(uint8_t* dst, const void* src, uint32_t bytes, uint64_t* bar, int tid) {
  if (tid == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (tid == 0) {
    tc::mbar_arrive_expect_tx(bar, bytes);
    for (uint32_t off = 0; off < bytes; off += 32768) {
      const uint32_t n = bytes - off < 32768 ? bytes - off : 32768;
      tc::bulk_g2s(tc::smem_u32(dst + off), static_cast<const uint8_t*>(src) + off, n, bar);
    }
  }
  tc::mbar_wait(bar, 0);
}

template <int KIND, int BPC>
__global__ void __launch_bounds__(RR_THREADS, 1)
rnn_res_fwd_kernel(const float* __restrict__ xp, int ldx, const __nv_bfloat16* __restrict__ wt_g, const float* __restrict__ b_hh,
                   const float* __restrict__ h0, const int32_t* __restrict__ lens, float* __restrict__ gates,
                   float* __restrict__ hs, float* __restrict__ cs, float* __restrict__ user, int B, int S, int H) {
  pdl_trigger();
  pdl_wait();
  constexpr int G = KIND == 0 ? 4 : 3;
  const int GH = G * H, NT = (GH + 7) / 8, GHp = NT * 8, KR = (H + RR_KS - 1) / RR_KS;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Wt = reinterpret_cast<__nv_bfloat16*>(smem_raw);                 // [H][GHp]   Wt[k][n] = W_hh[n][k]
  float* part = reinterpret_cast<float*>(smem_raw + (size_t)H * GHp * 2);         // [KS][BPC][GHp]
  float* h_s = part + (size_t)RR_KS * BPC * GHp;                                  // [H][BPC]
  float* c_s = h_s + H * BPC;                                                     // [BPC][H]
  int* len_s = reinterpret_cast<int*>(c_s + BPC * H);
  const int tid = threadIdx.x, b0 = blockIdx.x * BPC;

  __shared__ uint64_t w_bar;
  bulk_fill(reinterpret_cast<uint8_t*>(Wt), wt_g, (uint32_t)(H * GHp * 2), &w_bar, tid);
  for (int i = tid; i < BPC * H; i += RR_THREADS) {
    const int bl = i / H, j = i - bl * H, b = b0 + bl;
    h_s[j * BPC + bl] = (h0 != nullptr && b < B) ? h0[(int64_t)b * H + j] : 0.f;
    c_s[i] = 0.f;
  }
  if (tid < BPC) {
    const int b = b0 + tid;
    const int l = b < B ? (lens ? lens[b] : S) : 0;
    len_s[tid] = l < 0 ? 0 : (l > S ? S : l);
  }
  __syncthreads();
  int max_len = 0;
#pragma unroll
  for (int i = 0; i < BPC; ++i) max_len = max(max_len, len_s[i]);

  // matvec role
  const bool mv = tid < NT * RR_KS;
  const int ng = tid % NT, ks = tid / NT;
  const int k0 = ks * KR, k1 = min(H, k0 + KR);
  // gate role: one (sequence, unit) per thread
  const bool gt = tid < BPC * H;
  const int bl = gt ? tid / H : 0, j = gt ? tid - bl * H : 0, b = b0 + bl;
  const int my_len = gt && b < B ? len_s[bl] : 0;
  float bh[G];
#pragma unroll
  for (int g = 0; g < G; ++g) bh[g] = (KIND == 1 && gt) ? __ldg(b_hh + g * H + j) : 0.f;

  for (int s = 0; s < max_len; ++s) {
    float xv[G];
    const bool act = gt && s < my_len;
    if (act) {
      const float* xps = xp + ((int64_t)b * S + s) * ldx + j;
#pragma unroll
      for (int g = 0; g < G; ++g) xv[g] = __ldg(xps + g * H);          // consumed after the matvec: latency hidden
    }
    if (mv) {
      float acc[BPC][8];
#pragma unroll
      for (int q = 0; q < BPC; ++q)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;
      const uint4* wp = reinterpret_cast<const uint4*>(Wt + (size_t)k0 * GHp + ng * 8);
      const int wstride = GHp / 8;
#pragma unroll 2
      for (int k = k0; k < k1; ++k) {
        const uint4 w = wp[(size_t)(k - k0) * wstride];
        const float wf[8] = {bf_lo(w.x), bf_hi(w.x), bf_lo(w.y), bf_hi(w.y), bf_lo(w.z), bf_hi(w.z), bf_lo(w.w), bf_hi(w.w)};
#pragma unroll
        for (int q = 0; q < BPC; ++q) {
          const float hv = h_s[k * BPC + q];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[q][i] = fmaf(wf[i], hv, acc[q][i]);
        }
      }
#pragma unroll
      for (int q = 0; q < BPC; ++q) {
        float4* d = reinterpret_cast<float4*>(part + ((size_t)ks * BPC + q) * GHp + ng * 8);
        d[0] = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
        d[1] = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
      }
    }
    __syncthreads();
    if (act) {
      float pre[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float a = bh[g];
#pragma unroll
        for (int q = 0; q < RR_KS; ++q) a += part[((size_t)q * BPC + bl) * GHp + g * H + j];
        pre[g] = a;
      }
      float* gs = gates + ((int64_t)b * S + s) * GH + j;
      const int64_t o = ((int64_t)b * S + s) * H + j;
      if (KIND == 0) {
        const float gi = sigm(xv[0] + pre[0]), gf = sigm(xv[1] + pre[1]);
        const float gg = tanhf(xv[2] + pre[2]), go = sigm(xv[3] + pre[3]);
        const float c = gf * c_s[tid] + gi * gg;
        const float h = go * tanhf(c);
        gs[0] = gi; gs[H] = gf; gs[2 * H] = gg; gs[3 * H] = go;
        c_s[tid] = c; h_s[j * BPC + bl] = h;
        cs[o] = c; hs[o] = h;
      } else {
        const float r = sigm(xv[0] + pre[0]), z = sigm(xv[1] + pre[1]);
        const float hn = pre[2];
        const float nn = tanhf(xv[2] + r * hn);
        const float h = (1.f - z) * nn + z * h_s[j * BPC + bl];
        gs[0] = r; gs[H] = z; gs[2 * H] = nn;
        h_s[j * BPC + bl] = h;
        cs[o] = hn; hs[o] = h;
      }
    }
    __syncthreads();
  }
  if (gt && b < B) user[(int64_t)b * H + j] = h_s[j * BPC + bl];
}

template <int KIND, int BPC>
__global__ void __launch_bounds__(RR_THREADS, 1)
rnn_res_bwd_kernel(const __nv_bfloat16* __restrict__ w_g, const float* __restrict__ h0, const int32_t* __restrict__ lens,
                   const float* __restrict__ gates, const float* __restrict__ hs, const float* __restrict__ cs,
                   const float* __restrict__ d_user, float* __restrict__ dgi, float* __restrict__ dgh,
                   float* __restrict__ d_h0, int B, int S, int H, __nv_bfloat16* __restrict__ gib, __nv_bfloat16* __restrict__ ghb,
                   int GHp16, float* __restrict__ bias_part) {
  pdl_trigger();
  pdl_wait();
  // gib / ghb != nullptr: the gate gradients are written as bf16 rows of pitch GHp16 (the layout the weight-gradient
  // GEMMs read; pre-zeroed by the host) instead of fp32 dgi / dgh, and the bias gradients (column sums of dgi / dgh over
  // this CTA's sequences and steps) go to bias_part[blockIdx][2][GH]
  constexpr int G = KIND == 0 ? 4 : 3;
  const int GH = G * H, Hp = (H + 7) / 8 * 8, NTk = Hp / 8, NR = (GH + RR_NS - 1) / RR_NS;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* W = reinterpret_cast<__nv_bfloat16*>(smem_raw);                  // [GH][Hp]
  float* part = reinterpret_cast<float*>(smem_raw + (size_t)GH * Hp * 2);         // [NS][BPC][Hp]
  float* dp = part + (size_t)RR_NS * BPC * Hp;                                    // [GH][BPC]
  int* len_s = reinterpret_cast<int*>(dp + (size_t)BPC * GH);
  const int tid = threadIdx.x, b0 = blockIdx.x * BPC;

  __shared__ uint64_t w_bar;
  bulk_fill(reinterpret_cast<uint8_t*>(W), w_g, (uint32_t)(GH * Hp * 2), &w_bar, tid);
  if (tid < BPC) {
    const int b = b0 + tid;
    const int l = b < B ? (lens ? lens[b] : S) : 0;
    len_s[tid] = l < 0 ? 0 : (l > S ? S : l);
  }
  __syncthreads();
  int max_len = 0;
#pragma unroll
  for (int i = 0; i < BPC; ++i) max_len = max(max_len, len_s[i]);
  // steps beyond a sequence's length contribute nothing: zero their gate gradients
  if (gib == nullptr)
  for (int q = 0; q < BPC; ++q) {
    const int b = b0 + q;
    if (b >= B) continue;
    for (int64_t i = (int64_t)len_s[q] * GH + tid; i < (int64_t)S * GH; i += RR_THREADS) {
      dgi[(int64_t)b * S * GH + i] = 0.f;
      if (KIND == 1) dgh[(int64_t)b * S * GH + i] = 0.f;
    }
  }

  const bool mv = tid < NTk * RR_NS;
  const int kg = tid % NTk, ns = tid / NTk;
  const int n0 = ns * NR, n1 = min(GH, n0 + NR);
  const bool gt = tid < BPC * H;
  const int bl = gt ? tid / H : 0, j = gt ? tid - bl * H : 0, b = b0 + bl;
  const int my_len = gt && b < B ? len_s[bl] : 0;
  float dh_c = 0.f, dc_c = 0.f;      // carried dL/dh and dL/dc (LSTM) / direct z-path term (GRU), private to (b, j)
  float bsum[G], bsum_hn = 0.f;      // bias-gradient partial sums of this (b, j) over its steps (bsum_hn: GRU hidden-side n gate)
#pragma unroll
  for (int g = 0; g < G; ++g) bsum[g] = 0.f;

  for (int s = max_len - 1; s >= 0; --s) {
    const bool act = gt && s < my_len;
    float d[G];
#pragma unroll
    for (int g = 0; g < G; ++g) d[g] = 0.f;
    if (act) {
      float dh = dh_c;
      if (s == my_len - 1) dh += d_user[(int64_t)b * H + j];
      const float* gs = gates + ((int64_t)b * S + s) * GH + j;
      const int64_t o = ((int64_t)b * S + s) * H + j;
      float* gi_out = dgi + ((int64_t)b * S + s) * GH + j;
      if (KIND == 0) {
        const float gi = gs[0], gf = gs[H], gg = gs[2 * H], go = gs[3 * H];
        const float c = cs[o];
        const float cprev = s > 0 ? cs[o - H] : 0.f;
        const float tc = tanhf(c);
        const float dc = dc_c + dh * go * (1.f - tc * tc);
        d[0] = dc * gg * gi * (1.f - gi);
        d[1] = dc * cprev * gf * (1.f - gf);
        d[2] = dc * gi * (1.f - gg * gg);
        d[3] = dh * tc * go * (1.f - go);
        dc_c = dc * gf;
        if (gib != nullptr) {
          __nv_bfloat16* go16 = gib + ((int64_t)b * S + s) * GHp16 + j;
#pragma unroll
          for (int g = 0; g < G; ++g) { go16[g * H] = __float2bfloat16(d[g]); bsum[g] += d[g]; }
        } else {
          gi_out[0] = d[0]; gi_out[H] = d[1]; gi_out[2 * H] = d[2]; gi_out[3 * H] = d[3];
        }
      } else {
        const float r = gs[0], z = gs[H], nn = gs[2 * H];
        const float hn = cs[o];
        const float hprev = s > 0 ? hs[o - H] : (h0 ? h0[(int64_t)b * H + j] : 0.f);
        const float dn = dh * (1.f - z) * (1.f - nn * nn);
        const float dz = dh * (hprev - nn) * z * (1.f - z);
        const float dr = dn * hn * r * (1.f - r);
        dc_c = dh * z;
        if (gib != nullptr) {
          __nv_bfloat16* gi16 = gib + ((int64_t)b * S + s) * GHp16 + j;
          __nv_bfloat16* gh16 = ghb + ((int64_t)b * S + s) * GHp16 + j;
          gi16[0] = __float2bfloat16(dr); gi16[H] = __float2bfloat16(dz); gi16[2 * H] = __float2bfloat16(dn);
          gh16[0] = __float2bfloat16(dr); gh16[H] = __float2bfloat16(dz); gh16[2 * H] = __float2bfloat16(dn * r);
          bsum[0] += dr; bsum[1] += dz; bsum[2] += dn; bsum_hn += dn * r;
        } else {
          gi_out[0] = dr; gi_out[H] = dz; gi_out[2 * H] = dn;
          float* gh_out = dgh + ((int64_t)b * S + s) * GH + j;
          gh_out[0] = dr; gh_out[H] = dz; gh_out[2 * H] = dn * r;
        }
        d[0] = dr; d[1] = dz; d[2] = dn * r;
      }
    }
    if (gt) {
#pragma unroll
      for (int g = 0; g < G; ++g) dp[(size_t)(g * H + j) * BPC + bl] = d[g];
    }
    __syncthreads();
    if (mv) {
      float acc[BPC][8];
#pragma unroll
      for (int q = 0; q < BPC; ++q)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;
      const uint4* wp = reinterpret_cast<const uint4*>(W + (size_t)n0 * Hp + kg * 8);
#pragma unroll 2
      for (int n = n0; n < n1; ++n) {
        const uint4 w = wp[(size_t)(n - n0) * NTk];
        const float wf[8] = {bf_lo(w.x), bf_hi(w.x), bf_lo(w.y), bf_hi(w.y), bf_lo(w.z), bf_hi(w.z), bf_lo(w.w), bf_hi(w.w)};
#pragma unroll
        for (int q = 0; q < BPC; ++q) {
          const float dv = dp[(size_t)n * BPC + q];
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[q][i] = fmaf(wf[i], dv, acc[q][i]);
        }
      }
#pragma unroll
      for (int q = 0; q < BPC; ++q) {
        float4* o4 = reinterpret_cast<float4*>(part + ((size_t)ns * BPC + q) * Hp + kg * 8);
        o4[0] = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
        o4[1] = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
      }
    }
    __syncthreads();
    if (act) {
      float a = (KIND == 1) ? dc_c : 0.f;
#pragma unroll
      for (int q = 0; q < RR_NS; ++q) a += part[((size_t)q * BPC + bl) * Hp + j];
      dh_c = a;
    }
  }
  if (d_h0 != nullptr && gt && b < B) d_h0[(int64_t)b * H + j] = dh_c;
  if (bias_part != nullptr) {
    // fold the BPC sequences of this CTA (fixed order) through the dp scratch: dp[(g*H + j) * BPC + bl]
    __syncthreads();
    float* dp2 = dp;                                   // [2][GH][BPC] would not fit for the hidden-side copy: two passes
    for (int pass = 0; pass < 2; ++pass) {
      if (gt) {
#pragma unroll
        for (int g = 0; g < G; ++g)
          dp2[(size_t)(g * H + j) * BPC + bl] = (pass == 1 && KIND == 1 && g == 2) ? bsum_hn : bsum[g];
      }
      __syncthreads();
      for (int n = tid; n < GH; n += RR_THREADS) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < BPC; ++q) a += dp2[(size_t)n * BPC + q];
        bias_part[((size_t)blockIdx.x * 2 + pass) * GH + n] = a;
      }
      __syncthreads();
    }
  }
}

template <int KIND>
static int launch_fwd(int bpc, const float* xp, int ldx, const __nv_bfloat16* w_hh, const float* b_hh, const float* h0, const int32_t* lens,
                      float* gates, float* hs, float* cs, float* user, int B, int S, int H, cudaStream_t st) {
  const RRGeom g = rr_geom(KIND == 0 ? MR_RNN_LSTM : MR_RNN_GRU, H, bpc);
  const unsigned grid = (unsigned)ceil_div(B, bpc);
#define RR_LAUNCH_F(BPC)                                                                                               \
  {                                                                                                                    \
    cudaFuncSetAttribute(rnn_res_fwd_kernel<KIND, BPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.fwd_bytes); \
    launch_pdl(rnn_res_fwd_kernel<KIND, BPC>, dim3(grid), dim3(RR_THREADS), g.fwd_bytes, st, xp, ldx, w_hh, b_hh, h0, lens, gates, hs, cs, user, B, S, H); \
  }
  if (bpc == 1) RR_LAUNCH_F(1) else if (bpc == 2) RR_LAUNCH_F(2) else RR_LAUNCH_F(4)
#undef RR_LAUNCH_F
  MR_CHECK_LAUNCH("rnn_res_fwd_kernel");
  return MR_OK;
}

template <int KIND>
static int launch_bwd(int bpc, const __nv_bfloat16* w_hh, const float* h0, const int32_t* lens, const float* gates, const float* hs,
                      const float* cs, const float* d_user, float* dgi, float* dgh, float* d_h0, int B, int S, int H,
                      __nv_bfloat16* gib, __nv_bfloat16* ghb, int GHp16, float* bias_part, cudaStream_t st) {
  const RRGeom g = rr_geom(KIND == 0 ? MR_RNN_LSTM : MR_RNN_GRU, H, bpc);
  const unsigned grid = (unsigned)ceil_div(B, bpc);
#define RR_LAUNCH_B(BPC)                                                                                               \
  {                                                                                                                    \
    cudaFuncSetAttribute(rnn_res_bwd_kernel<KIND, BPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.bwd_bytes); \
    launch_pdl(rnn_res_bwd_kernel<KIND, BPC>, dim3(grid), dim3(RR_THREADS), g.bwd_bytes, st, w_hh, h0, lens, gates, hs, cs, d_user, dgi, dgh, d_h0, B, S, H, gib, ghb, GHp16, bias_part); \
  }
  if (bpc == 1) RR_LAUNCH_B(1) else if (bpc == 2) RR_LAUNCH_B(2) else RR_LAUNCH_B(4)
#undef RR_LAUNCH_B
  MR_CHECK_LAUNCH("rnn_res_bwd_kernel");
  return MR_OK;
}

int64_t rnn_res_scratch_bytes(int kind, int H) {
  const RRGeom g = rr_geom(kind, H, 1);
  const int64_t a = (int64_t)H * g.GHp * 2, b = (int64_t)g.GH * g.Hp * 2;
  return (a > b ? a : b) + 256;
}

int rnn_res_fwd(int kind, const float* xp, int ldx, const float* w_hh_f32, const float* b_hh, const float* h0, const int32_t* lens,
                float* gates, float* hs, float* cs, float* user, int B, int S, int H, void* scratch, cudaStream_t st) {
  const int bpc = rnn_res_bpc(kind, B, H);
  const RRGeom g = rr_geom(kind, H, bpc);
  __nv_bfloat16* w_hh = static_cast<__nv_bfloat16*>(scratch);
  launch_pdl(rnn_res_prep_kernel, dim3((unsigned)ceil_div((int64_t)H * g.GHp, 256)), dim3(256), 0, st, w_hh_f32, w_hh, nullptr, g.GH, H, g.GHp, g.Hp);
  MR_CHECK_LAUNCH("rnn_res_prep_kernel");
  return kind == MR_RNN_LSTM ? launch_fwd<0>(bpc, xp, ldx, w_hh, b_hh, h0, lens, gates, hs, cs, user, B, S, H, st)
                             : launch_fwd<1>(bpc, xp, ldx, w_hh, b_hh, h0, lens, gates, hs, cs, user, B, S, H, st);
}

int rnn_res_bwd(int kind, const float* w_hh_f32, const float* h0, const int32_t* lens, const float* gates, const float* hs,
                const float* cs, const float* d_user, float* dgi, float* dgh, float* d_h0, int B, int S, int H,
                void* scratch, cudaStream_t st, __nv_bfloat16* gib, __nv_bfloat16* ghb, int GHp16, float* bias_part) {
  const int bpc = rnn_res_bpc(kind, B, H);
  const RRGeom g = rr_geom(kind, H, bpc);
  __nv_bfloat16* w_hh = static_cast<__nv_bfloat16*>(scratch);
  launch_pdl(rnn_res_prep_kernel, dim3((unsigned)ceil_div((int64_t)g.GH * g.Hp, 256)), dim3(256), 0, st, w_hh_f32, nullptr, w_hh, g.GH, H, g.GHp, g.Hp);
  MR_CHECK_LAUNCH("rnn_res_prep_kernel");
  return kind == MR_RNN_LSTM ? launch_bwd<0>(bpc, w_hh, h0, lens, gates, hs, cs, d_user, dgi, dgh, d_h0, B, S, H, gib, ghb, GHp16, bias_part, st)
                             : launch_bwd<1>(bpc, w_hh, h0, lens, gates, hs, cs, d_user, dgi, dgh, d_h0, B, S, H, gib, ghb, GHp16, bias_part, st);
}

}  // namespace mr
