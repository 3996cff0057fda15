// Multi-head self-attention core, register-tiled (models/Modules/Attention.py:115-147: shared q/k projection, pair mask
// m_i * m_j, XSoftmax, P V; no output projection) for the shapes the news / user encoders use: len <= 64 positions,
// per-head widths dk, dv <= 32 (title 48 x 30 / 15, history 50 x 15 / 15).  fp32 throughout -- it serves both precisions.
//
// One CTA of 128 threads per (sequence, head).  The q(=k), v (and, backward, d_ctx) rows of the head sit in shared memory
// with a row pitch of 36 floats (= 4 mod 32: eight consecutive rows read as float4 hit eight disjoint bank groups).
// Thread (ty, tx) of the 16 x 8 layout owns the RA x CB register tile of rows ty + 16 a, columns tx + 8 b of every
// len x len product (S = q q^T, dP = dO v^T) and RA x 4 tiles of the len x width products (P v, P^T dO, (dS + dS^T) q),
// reading its operands as float4 along the contraction index.  The round-1 kernel had one thread per query row with scalar
// shared-memory operands (2 loads per FMA): 45 ms forward + 35 ms backward at the config-5 shape (28,160 titles x 10 heads).
// Inputs and outputs carry explicit row pitches: q|k and v are column slices of ONE projection output [rows, ldq]
// (linear_tc.cu), and the backward writes d_qk / d_v straight into the slices of that layer's gradient.
#include "common.cuh"

namespace mr {

constexpr int AT_THREADS = 128;
constexpr int AT_P = 36;          // row pitch (floats) of the q / v / d_ctx tiles

template <int RA, int CB>
__global__ void __launch_bounds__(AT_THREADS)
mha_attn_fwd_kernel(const float* __restrict__ qk, int64_t ldq, const float* __restrict__ v, int64_t ldv,
                    const float* __restrict__ mask, float* __restrict__ prob, float* __restrict__ ctx, int64_t ldc, int len,
                    int hn, int dk, int dv) {
  constexpr int LT = 16 * RA, PP = LT + 4;
  static_assert(LT == 8 * CB, "tile geometry");
  extern __shared__ float4 at_smem4[];
  float* Qs = reinterpret_cast<float*>(at_smem4);      // [LT][36]
  float* Vs = Qs + LT * AT_P;                          // [LT][36]
  float* Ps = Vs + LT * AT_P;                          // [LT][PP]
  float* ms = Ps + LT * PP;                            // [LT]
  const int64_t n = blockIdx.x / hn;
  const int h = blockIdx.x % hn;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  for (int i = tid; i < LT * AT_P; i += AT_THREADS) {
    const int r = i / AT_P, c = i - r * AT_P;
    Qs[i] = (r < len && c < dk) ? qk[(n * len + r) * ldq + h * dk + c] : 0.f;
    Vs[i] = (r < len && c < dv) ? v[(n * len + r) * ldv + h * dv + c] : 0.f;
  }
  for (int i = tid; i < LT; i += AT_THREADS) ms[i] = i < len ? (mask ? mask[n * len + i] : 1.f) : 0.f;
  __syncthreads();
  // ---- S = q q^T -------------------------------------------------------------------------------------------------
  float acc[RA][CB];
#pragma unroll
  for (int a = 0; a < RA; ++a)
#pragma unroll
    for (int b = 0; b < CB; ++b) acc[a][b] = 0.f;
  for (int k4 = 0; k4 < dk; k4 += 4) {
    float4 qa[RA], qb[CB];
#pragma unroll
    for (int a = 0; a < RA; ++a) qa[a] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * a) * AT_P + k4);
#pragma unroll
    for (int b = 0; b < CB; ++b) qb[b] = *reinterpret_cast<const float4*>(Qs + (tx + 8 * b) * AT_P + k4);
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
      for (int b = 0; b < CB; ++b)
        acc[a][b] = fmaf(qa[a].w, qb[b].w, fmaf(qa[a].z, qb[b].z, fmaf(qa[a].y, qb[b].y, fmaf(qa[a].x, qb[b].x, acc[a][b]))));
  }
  // ---- masked softmax over the row (XSoftmax, Attention.py:56-76): the 8 threads of a row are 8 consecutive lanes ----------
  const float inv = rsqrtf((float)dk);
#pragma unroll
  for (int a = 0; a < RA; ++a) {
    const int i = ty + 16 * a;
    const bool row_on = ms[i] != 0.f;
    float mx = -INFINITY;
#pragma unroll
    for (int b = 0; b < CB; ++b) {
      acc[a][b] *= inv;
      if (row_on && ms[tx + 8 * b] != 0.f) mx = fmaxf(mx, acc[a][b]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
    float sum = 0.f;
#pragma unroll
    for (int b = 0; b < CB; ++b) {
      const float e = (row_on && ms[tx + 8 * b] != 0.f) ? expf(acc[a][b] - mx) : 0.f;
      acc[a][b] = e;
      sum += e;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    const float rs = sum > 0.f ? 1.f / sum : 0.f;        // an all-masked row gives zeros, not NaN
#pragma unroll
    for (int b = 0; b < CB; ++b) Ps[i * PP + tx + 8 * b] = acc[a][b] * rs;
  }
  __syncthreads();
  // the probabilities are saved for the backward: coalesced copy of the len x len block
  {
    float* pg = prob + ((n * hn + h) * (int64_t)len) * len;
    for (int i = tid; i < len * len; i += AT_THREADS) {
      const int r = i / len, c = i - r * len;
      pg[i] = Ps[r * PP + c];
    }
  }
  // ---- ctx = P v : rows ty + 16 a, columns tx + 8 z ----------------------------------------------------------------
  float o[RA][4];
#pragma unroll
  for (int a = 0; a < RA; ++a)
#pragma unroll
    for (int z = 0; z < 4; ++z) o[a][z] = 0.f;
  for (int j4 = 0; j4 < len; j4 += 4) {
    float4 p[RA];
#pragma unroll
    for (int a = 0; a < RA; ++a) p[a] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * a) * PP + j4);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float vv[4];
#pragma unroll
      for (int z = 0; z < 4; ++z) vv[z] = Vs[(j4 + jj) * AT_P + tx + 8 * z];
#pragma unroll
      for (int a = 0; a < RA; ++a) {
        const float pj = jj == 0 ? p[a].x : (jj == 1 ? p[a].y : (jj == 2 ? p[a].z : p[a].w));
#pragma unroll
        for (int z = 0; z < 4; ++z) o[a][z] = fmaf(pj, vv[z], o[a][z]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < RA; ++a) {
    const int i = ty + 16 * a;
    if (i < len) {
#pragma unroll
      for (int z = 0; z < 4; ++z)
        if (tx + 8 * z < dv) ctx[(n * len + i) * ldc + h * dv + tx + 8 * z] = o[a][z];
    }
  }
}

// d_ctx [rows, ldg] -> d_qk [rows, ldo_q] (+ h * dk), d_v [rows, ldo_v] (+ h * dv); prob from the forward
template <int RA, int CB>
__global__ void __launch_bounds__(AT_THREADS)
mha_attn_bwd_kernel(const float* __restrict__ qk, int64_t ldq, const float* __restrict__ v, int64_t ldv,
                    const float* __restrict__ prob, const float* __restrict__ d_ctx, int64_t ldg, float* __restrict__ d_qk,
                    int64_t ldo_q, float* __restrict__ d_v, int64_t ldo_v, int len, int hn, int dk, int dv) {
  constexpr int LT = 16 * RA, PP = LT + 4;
  extern __shared__ float4 at_smem4[];
  float* Qs = reinterpret_cast<float*>(at_smem4);      // [LT][36]
  float* Vs = Qs + LT * AT_P;
  float* Gs = Vs + LT * AT_P;                          // d_ctx rows
  float* Ps = Gs + LT * AT_P;                          // [LT][PP]
  float* Ds = Ps + LT * PP;                            // [LT][PP]  dS
  const int64_t n = blockIdx.x / hn;
  const int h = blockIdx.x % hn;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  for (int i = tid; i < LT * AT_P; i += AT_THREADS) {
    const int r = i / AT_P, c = i - r * AT_P;
    Qs[i] = (r < len && c < dk) ? qk[(n * len + r) * ldq + h * dk + c] : 0.f;
    Vs[i] = (r < len && c < dv) ? v[(n * len + r) * ldv + h * dv + c] : 0.f;
    Gs[i] = (r < len && c < dv) ? d_ctx[(n * len + r) * ldg + h * dv + c] : 0.f;
  }
  for (int i = tid; i < LT * PP; i += AT_THREADS) Ps[i] = 0.f;
  __syncthreads();
  {
    const float* pg = prob + ((n * hn + h) * (int64_t)len) * len;
    for (int i = tid; i < len * len; i += AT_THREADS) {
      const int r = i / len, c = i - r * len;
      Ps[r * PP + c] = pg[i];
    }
  }
  __syncthreads();
  // ---- dP = d_ctx v^T, dS = P (dP - <P, dP>_row) / sqrt(dk)   (XSoftmax backward, Attention.py:77-80) ----------------------
  {
    float acc[RA][CB];
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
      for (int b = 0; b < CB; ++b) acc[a][b] = 0.f;
    for (int c4 = 0; c4 < dv; c4 += 4) {
      float4 ga[RA], vb[CB];
#pragma unroll
      for (int a = 0; a < RA; ++a) ga[a] = *reinterpret_cast<const float4*>(Gs + (ty + 16 * a) * AT_P + c4);
#pragma unroll
      for (int b = 0; b < CB; ++b) vb[b] = *reinterpret_cast<const float4*>(Vs + (tx + 8 * b) * AT_P + c4);
#pragma unroll
      for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int b = 0; b < CB; ++b)
          acc[a][b] = fmaf(ga[a].w, vb[b].w, fmaf(ga[a].z, vb[b].z, fmaf(ga[a].y, vb[b].y, fmaf(ga[a].x, vb[b].x, acc[a][b]))));
    }
    const float inv = rsqrtf((float)dk);
#pragma unroll
    for (int a = 0; a < RA; ++a) {
      const int i = ty + 16 * a;
      float pr[CB];
      float dot = 0.f;
#pragma unroll
      for (int b = 0; b < CB; ++b) {
        pr[b] = Ps[i * PP + tx + 8 * b];
        dot = fmaf(pr[b], acc[a][b], dot);
      }
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
#pragma unroll
      for (int b = 0; b < CB; ++b) Ds[i * PP + tx + 8 * b] = pr[b] * (acc[a][b] - dot) * inv;
    }
  }
  __syncthreads();
  // ---- d_v[j, c] = sum_i P[i, j] d_ctx[i, c] : rows j = ty + 16 a, columns c = tx + 8 z --------------------------------------
  {
    float o[RA][4];
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
      for (int z = 0; z < 4; ++z) o[a][z] = 0.f;
    for (int i = 0; i < len; ++i) {
      float gg[4], pc[RA];
#pragma unroll
      for (int z = 0; z < 4; ++z) gg[z] = Gs[i * AT_P + tx + 8 * z];
#pragma unroll
      for (int a = 0; a < RA; ++a) pc[a] = Ps[i * PP + ty + 16 * a];
#pragma unroll
      for (int a = 0; a < RA; ++a)
#pragma unroll
        for (int z = 0; z < 4; ++z) o[a][z] = fmaf(pc[a], gg[z], o[a][z]);
    }
#pragma unroll
    for (int a = 0; a < RA; ++a) {
      const int j = ty + 16 * a;
      if (j < len) {
#pragma unroll
        for (int z = 0; z < 4; ++z)
          if (tx + 8 * z < dv) d_v[(n * len + j) * ldo_v + h * dv + tx + 8 * z] = o[a][z];
      }
    }
  }
  // ---- d_qk[i, c] = sum_j (dS[i, j] + dS[j, i]) q[j, c]  (the projection is shared by queries and keys, Attention.py:125-126) ---
  {
    float o[RA][4];
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
      for (int z = 0; z < 4; ++z) o[a][z] = 0.f;
    for (int j4 = 0; j4 < len; j4 += 4) {
      float4 dr[RA];
#pragma unroll
      for (int a = 0; a < RA; ++a) dr[a] = *reinterpret_cast<const float4*>(Ds + (ty + 16 * a) * PP + j4);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        float qq[4];
#pragma unroll
        for (int z = 0; z < 4; ++z) qq[z] = Qs[(j4 + jj) * AT_P + tx + 8 * z];
#pragma unroll
        for (int a = 0; a < RA; ++a) {
          const float rowv = jj == 0 ? dr[a].x : (jj == 1 ? dr[a].y : (jj == 2 ? dr[a].z : dr[a].w));
          const float t = rowv + Ds[(j4 + jj) * PP + ty + 16 * a];
#pragma unroll
          for (int z = 0; z < 4; ++z) o[a][z] = fmaf(t, qq[z], o[a][z]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < RA; ++a) {
      const int i = ty + 16 * a;
      if (i < len) {
#pragma unroll
        for (int z = 0; z < 4; ++z)
          if (tx + 8 * z < dk) d_qk[(n * len + i) * ldo_q + h * dk + tx + 8 * z] = o[a][z];
      }
    }
  }
}

static size_t at_smem_fwd(int LT) { return sizeof(float) * (size_t)(2 * LT * AT_P + LT * (LT + 4) + LT); }
static size_t at_smem_bwd(int LT) { return sizeof(float) * (size_t)(3 * LT * AT_P + 2 * LT * (LT + 4)); }

bool mha_attn_supported(int64_t len, int64_t dk, int64_t dv) { return len >= 1 && len <= 64 && dk >= 1 && dk <= 32 && dv >= 1 && dv <= 32; }

int mha_attn_fwd(const float* qk, int64_t ldq, const float* v, int64_t ldv, const float* mask, float* prob, float* ctx, int64_t ldc,
                 int64_t n, int64_t len, int64_t hn, int64_t dk, int64_t dv, cudaStream_t st) {
  const unsigned grid = (unsigned)(n * hn);
#define AT_LAUNCH_F(RA, CB)                                                                                                         \
  {                                                                                                                                 \
    const size_t smem = at_smem_fwd(16 * RA);                                                                                       \
    if (smem > 48 * 1024) cudaFuncSetAttribute(mha_attn_fwd_kernel<RA, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    mha_attn_fwd_kernel<RA, CB><<<grid, AT_THREADS, smem, st>>>(qk, ldq, v, ldv, mask, prob, ctx, ldc, (int)len, (int)hn, (int)dk, (int)dv); \
  }
  if (len <= 32) AT_LAUNCH_F(2, 4) else if (len <= 48) AT_LAUNCH_F(3, 6) else AT_LAUNCH_F(4, 8)
#undef AT_LAUNCH_F
  MR_CHECK_LAUNCH("mha_attn_fwd_kernel");
  return MR_OK;
}

int mha_attn_bwd(const float* qk, int64_t ldq, const float* v, int64_t ldv, const float* prob, const float* d_ctx, int64_t ldg,
                 float* d_qk, int64_t ldo_q, float* d_v, int64_t ldo_v, int64_t n, int64_t len, int64_t hn, int64_t dk, int64_t dv,
                 cudaStream_t st) {
  const unsigned grid = (unsigned)(n * hn);
#define AT_LAUNCH_B(RA, CB)                                                                                                         \
  {                                                                                                                                 \
    const size_t smem = at_smem_bwd(16 * RA);                                                                                       \
    if (smem > 48 * 1024) cudaFuncSetAttribute(mha_attn_bwd_kernel<RA, CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    mha_attn_bwd_kernel<RA, CB><<<grid, AT_THREADS, smem, st>>>(qk, ldq, v, ldv, prob, d_ctx, ldg, d_qk, ldo_q, d_v, ldo_v, (int)len, \
                                                               (int)hn, (int)dk, (int)dv);                                          \
  }
  if (len <= 32) AT_LAUNCH_B(2, 4) else if (len <= 48) AT_LAUNCH_B(3, 6) else AT_LAUNCH_B(4, 8)
#undef AT_LAUNCH_B
  MR_CHECK_LAUNCH("mha_attn_bwd_kernel");
  return MR_OK;
}

}  // namespace mr

extern "C" {
using namespace mr;

/* pitched form of mr_mha_core_*: q|k rows at qk + row * ldq (+ h * dk), v rows at v + row * ldv (+ h * dv) */
int mr_mha_attn_fwd(const float* qk, int64_t ldq, const float* v, int64_t ldv, const float* mask, float* prob, float* ctx,
                    int64_t n, int64_t len, int64_t hn, int64_t dk, int64_t dv, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(qk && v && prob && ctx, MR_ERR_NULL, "mr_mha_attn_fwd: null pointer");
  MR_REQUIRE(n >= 0 && hn >= 1 && ldq >= hn * dk && ldv >= hn * dv && n * hn < (1ll << 31), MR_ERR_BAD_SHAPE, "mr_mha_attn_fwd: bad shape");
  MR_REQUIRE(mha_attn_supported(len, dk, dv), MR_ERR_UNSUPPORTED, "mr_mha_attn_fwd: len=%lld dk=%lld dv=%lld (limits 64 / 32 / 32)",
             (long long)len, (long long)dk, (long long)dv);
  if (n == 0) return MR_OK;
  return mha_attn_fwd(qk, ldq, v, ldv, mask, prob, ctx, hn * dv, n, len, hn, dk, dv, as_stream(stream));
}

int mr_mha_attn_bwd(const float* qk, int64_t ldq, const float* v, int64_t ldv, const float* prob, const float* d_ctx,
                    float* d_qk, int64_t ldo_q, float* d_v, int64_t ldo_v, int64_t n, int64_t len, int64_t hn, int64_t dk,
                    int64_t dv, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(qk && v && prob && d_ctx && d_qk && d_v, MR_ERR_NULL, "mr_mha_attn_bwd: null pointer");
  MR_REQUIRE(n >= 0 && hn >= 1 && ldq >= hn * dk && ldv >= hn * dv && ldo_q >= hn * dk && ldo_v >= hn * dv && n * hn < (1ll << 31),
             MR_ERR_BAD_SHAPE, "mr_mha_attn_bwd: bad shape");
  MR_REQUIRE(mha_attn_supported(len, dk, dv), MR_ERR_UNSUPPORTED, "mr_mha_attn_bwd: len=%lld dk=%lld dv=%lld (limits 64 / 32 / 32)",
             (long long)len, (long long)dk, (long long)dv);
  if (n == 0) return MR_OK;
  return mha_attn_bwd(qk, ldq, v, ldv, prob, d_ctx, hn * dv, d_qk, ldo_q, d_v, ldo_v, n, len, hn, dk, dv, as_stream(stream));
}

}  // extern "C"
