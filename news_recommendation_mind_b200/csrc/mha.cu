// Dense layer, multi-head self-attention core and LayerNorm for the MHA news / user encoders.
// Reference: models/Modules/Attention.py:83-147 (MultiheadAttention: shared q/k projection, pair
// mask, XSoftmax, no output projection), models/Encoders/MHA.py:21-39,58-75.
#include "gemm_simt.cuh"

namespace mr {
// register-tiled kernels for len <= 64, dk, dv <= 32 (mha_attn.cu); the kernels below remain for larger shapes
bool mha_attn_supported(int64_t len, int64_t dk, int64_t dv);
int mha_attn_fwd(const float* qk, int64_t ldq, const float* v, int64_t ldv, const float* mask, float* prob, float* ctx, int64_t ldc,
                 int64_t n, int64_t len, int64_t hn, int64_t dk, int64_t dv, cudaStream_t st);
int mha_attn_bwd(const float* qk, int64_t ldq, const float* v, int64_t ldv, const float* prob, const float* d_ctx, int64_t ldg,
                 float* d_qk, int64_t ldo_q, float* d_v, int64_t ldo_v, int64_t n, int64_t len, int64_t hn, int64_t dk, int64_t dv,
                 cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// attention core: one CTA per (sequence, head).  qk/v rows are staged in shared memory, each
// thread owns query rows i = tid, tid+128, ...
// ---------------------------------------------------------------------------------------------
constexpr int MHA_THREADS = 128;

__global__ void __launch_bounds__(MHA_THREADS)
mha_core_fwd_kernel(const float* __restrict__ qk, const float* __restrict__ v, const float* __restrict__ mask,
                    float* __restrict__ prob, float* __restrict__ ctx, int len, int hn, int dk, int dv) {
  extern __shared__ float sm[];
  float* q_s = sm;                        // [len][dk+1]
  float* v_s = q_s + len * (dk + 1);      // [len][dv+1]
  float* m_s = v_s + len * (dv + 1);      // [len]
  const int64_t n = blockIdx.x / hn;
  const int h = blockIdx.x % hn;
  const int tid = threadIdx.x;
  for (int i = tid; i < len * dk; i += MHA_THREADS) {
    int r = i / dk, c = i - r * dk;
    q_s[r * (dk + 1) + c] = qk[(n * len + r) * (int64_t)(hn * dk) + h * dk + c];
  }
  for (int i = tid; i < len * dv; i += MHA_THREADS) {
    int r = i / dv, c = i - r * dv;
    v_s[r * (dv + 1) + c] = v[(n * len + r) * (int64_t)(hn * dv) + h * dv + c];
  }
  for (int i = tid; i < len; i += MHA_THREADS) m_s[i] = mask ? mask[n * len + i] : 1.f;
  __syncthreads();
  const float inv = rsqrtf((float)dk);
  float* prow_base = prob + ((n * hn + h) * (int64_t)len) * len;
  for (int i = tid; i < len; i += MHA_THREADS) {
    float* prow = prow_base + (int64_t)i * len;
    const bool row_on = m_s[i] != 0.f;
    float mx = -INFINITY;
    for (int j = 0; j < len; ++j) {
      float s = 0.f;
      for (int c = 0; c < dk; ++c) s = fmaf(q_s[i * (dk + 1) + c], q_s[j * (dk + 1) + c], s);
      s *= inv;
      prow[j] = s;
      if (row_on && m_s[j] != 0.f) mx = fmaxf(mx, s);
    }
    float sum = 0.f;
    for (int j = 0; j < len; ++j) {
      float e = (row_on && m_s[j] != 0.f) ? expf(prow[j] - mx) : 0.f;
      prow[j] = e;
      sum += e;
    }
    const float rs = sum > 0.f ? 1.f / sum : 0.f;
    float* out = ctx + (n * len + i) * (int64_t)(hn * dv) + h * dv;
    for (int c = 0; c < dv; ++c) out[c] = 0.f;
    for (int j = 0; j < len; ++j) {
      float p = prow[j] * rs;
      prow[j] = p;
      if (p != 0.f)
        for (int c = 0; c < dv; ++c) out[c] = fmaf(p, v_s[j * (dv + 1) + c], out[c]);
    }
  }
}

__global__ void __launch_bounds__(MHA_THREADS)
mha_core_bwd_kernel(const float* __restrict__ qk, const float* __restrict__ v, const float* __restrict__ prob,
                    const float* __restrict__ d_ctx, float* __restrict__ d_qk, float* __restrict__ d_v, int len, int hn,
                    int dk, int dv) {
  extern __shared__ float sm[];
  float* q_s = sm;                        // [len][dk+1]
  float* v_s = q_s + len * (dk + 1);      // [len][dv+1]
  float* g_s = v_s + len * (dv + 1);      // [len][dv+1]  d_ctx rows
  float* ds_s = g_s + len * (dv + 1);     // [len][len+1] dS
  const int64_t n = blockIdx.x / hn;
  const int h = blockIdx.x % hn;
  const int tid = threadIdx.x;
  for (int i = tid; i < len * dk; i += MHA_THREADS) {
    int r = i / dk, c = i - r * dk;
    q_s[r * (dk + 1) + c] = qk[(n * len + r) * (int64_t)(hn * dk) + h * dk + c];
  }
  for (int i = tid; i < len * dv; i += MHA_THREADS) {
    int r = i / dv, c = i - r * dv;
    int64_t o = (n * len + r) * (int64_t)(hn * dv) + h * dv + c;
    v_s[r * (dv + 1) + c] = v[o];
    g_s[r * (dv + 1) + c] = d_ctx[o];
  }
  __syncthreads();
  const float inv = rsqrtf((float)dk);
  const float* pbase = prob + ((n * hn + h) * (int64_t)len) * len;
  for (int i = tid; i < len; i += MHA_THREADS) {
    const float* prow = pbase + (int64_t)i * len;
    float dot = 0.f;
    for (int j = 0; j < len; ++j) {
      float dp = 0.f;
      for (int c = 0; c < dv; ++c) dp = fmaf(g_s[i * (dv + 1) + c], v_s[j * (dv + 1) + c], dp);
      ds_s[i * (len + 1) + j] = dp;
      dot = fmaf(prow[j], dp, dot);
    }
    for (int j = 0; j < len; ++j) ds_s[i * (len + 1) + j] = prow[j] * (ds_s[i * (len + 1) + j] - dot) * inv;
  }
  __syncthreads();
  // d_v[j,:] = sum_i P[i,j] d_ctx[i,:] ;  d_qk[i,:] = sum_j (dS[i,j] + dS[j,i]) qk[j,:]
  for (int j = tid; j < len; j += MHA_THREADS) {
    float* dvo = d_v + (n * len + j) * (int64_t)(hn * dv) + h * dv;
    for (int c = 0; c < dv; ++c) {
      float s = 0.f;
      for (int i = 0; i < len; ++i) s = fmaf(pbase[(int64_t)i * len + j], g_s[i * (dv + 1) + c], s);
      dvo[c] = s;
    }
    float* dqo = d_qk + (n * len + j) * (int64_t)(hn * dk) + h * dk;
    for (int c = 0; c < dk; ++c) {
      float s = 0.f;
      for (int i = 0; i < len; ++i)
        s = fmaf(ds_s[j * (len + 1) + i] + ds_s[i * (len + 1) + j], q_s[i * (dk + 1) + c], s);
      dqo[c] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5) with optional inverted-dropout keep mask fused behind it; warp per row
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const uint8_t* __restrict__ keep, float keep_scale, float* __restrict__ y, float* __restrict__ mean,
                     float* __restrict__ rstd, int64_t M, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= M) return;
  const float* xr = x + r * H;
  float s = 0.f;
  for (int h = lane; h < H; h += 32) s += xr[h];
  const float mu = warp_sum(s) / (float)H;
  float q = 0.f;
  for (int h = lane; h < H; h += 32) { float d = xr[h] - mu; q = fmaf(d, d, q); }
  const float rs = rsqrtf(warp_sum(q) / (float)H + 1e-5f);
  for (int h = lane; h < H; h += 32) {
    float o = (xr[h] - mu) * rs * __ldg(gamma + h) + __ldg(beta + h);
    if (keep) o = keep[r * H + h] ? o * keep_scale : 0.f;
    y[r * H + h] = o;
  }
  if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
}

// each block handles rows r = blockIdx.x, +gridDim.x, ... (8 warps); writes per-block partials
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const uint8_t* __restrict__ keep,
                     float keep_scale, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ d_y, float* __restrict__ d_x, float* __restrict__ dg_part,
                     float* __restrict__ db_part, int64_t M, int H) {
  extern __shared__ float sm[];           // [8][H] dgamma, [8][H] dbeta
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* dg_s = sm + w * H;
  float* db_s = sm + 8 * H + w * H;
  for (int h = lane; h < H; h += 32) { dg_s[h] = 0.f; db_s[h] = 0.f; }
  for (int64_t r = (int64_t)blockIdx.x * 8 + w; r < M; r += (int64_t)gridDim.x * 8) {
    const float mu = mean[r], rs = rstd[r];
    float a = 0.f, b = 0.f;
    for (int h = lane; h < H; h += 32) {
      float g = d_y[r * H + h];
      if (keep) g = keep[r * H + h] ? g * keep_scale : 0.f;
      float xh = (x[r * H + h] - mu) * rs;
      float gg = g * __ldg(gamma + h);
      a += gg; b = fmaf(gg, xh, b);
      dg_s[h] = fmaf(g, xh, dg_s[h]);
      db_s[h] += g;
    }
    a = warp_sum(a) / (float)H; b = warp_sum(b) / (float)H;
    for (int h = lane; h < H; h += 32) {
      float g = d_y[r * H + h];
      if (keep) g = keep[r * H + h] ? g * keep_scale : 0.f;
      float xh = (x[r * H + h] - mu) * rs;
      d_x[r * H + h] = rs * (g * __ldg(gamma + h) - a - xh * b);
    }
  }
  __syncthreads();
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sg += sm[k * H + h]; sb += sm[8 * H + k * H + h]; }
    dg_part[(int64_t)blockIdx.x * H + h] = sg;
    db_part[(int64_t)blockIdx.x * H + h] = sb;
  }
}

}  // namespace mr

extern "C" {
using namespace mr;

int64_t mr_linear_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (M < 0 || N < 1 || K < 1) return -1;
  return arena_bytes(64 * N * K, 4) + arena_bytes(colsum_chunks(M) * N, 4) + 256;
}

int mr_linear_fwd(const float* x, const float* w, const float* b, float* y, int64_t M, int64_t N, int64_t K, int act,
                  int precision, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(x && w && y, MR_ERR_NULL, "mr_linear_fwd: null pointer");
  MR_REQUIRE(M >= 0 && N >= 1 && K >= 1 && act >= 0 && act <= 2, MR_ERR_BAD_SHAPE, "mr_linear_fwd: bad shape");
  (void)precision;
  if (M == 0) return MR_OK;
  cudaError_t e = gemm_simt<true, false>(M, N, K, RowMajor{x, K}, Transposed{w, K}, BiasActEpi{y, N, b, act}, 1, nullptr,
                                         as_stream(stream));
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "mr_linear_fwd: %s", cudaGetErrorString(e));
  return MR_OK;
}

int mr_linear_bwd(const float* x, const float* w, const float* d_y, float* d_x, float* d_w, float* d_b, int64_t M,
                  int64_t N, int64_t K, int precision, void* workspace, int64_t workspace_bytes, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(d_y, MR_ERR_NULL, "mr_linear_bwd: null d_y");
  MR_REQUIRE(M >= 0 && N >= 1 && K >= 1, MR_ERR_BAD_SHAPE, "mr_linear_bwd: bad shape");
  (void)precision;
  cudaStream_t st = as_stream(stream);
  if (M == 0) {
    if (d_w) cudaMemsetAsync(d_w, 0, sizeof(float) * N * K, st);
    if (d_b) cudaMemsetAsync(d_b, 0, sizeof(float) * N, st);
    return MR_OK;
  }
  Arena ar(workspace, workspace_bytes);
  float* sp = ar.take<float>(64 * N * K);
  float* cp = ar.take<float>(colsum_chunks(M) * N);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_linear_bwd: workspace too small");
  cudaError_t e;
  if (d_x) {
    MR_REQUIRE(w != nullptr, MR_ERR_NULL, "mr_linear_bwd: d_x needs w");
    e = gemm_simt<true, true>(M, K, N, RowMajor{d_y, N}, RowMajor{w, K}, StoreEpi{d_x, K}, 1, nullptr, st);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "mr_linear_bwd d_x: %s", cudaGetErrorString(e));
  }
  if (d_w) {
    MR_REQUIRE(x != nullptr, MR_ERR_NULL, "mr_linear_bwd: d_w needs x");
    e = gemm_simt<false, true>(N, K, M, Transposed{d_y, N}, RowMajor{x, K}, StoreEpi{d_w, K}, pick_splits(N, K, M), sp, st);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "mr_linear_bwd d_w: %s", cudaGetErrorString(e));
  }
  if (d_b) {
    e = colsum(d_y, d_b, M, N, cp, st);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "mr_linear_bwd d_b: %s", cudaGetErrorString(e));
  }
  return MR_OK;
}

static size_t mha_smem_fwd(int64_t len, int64_t dk, int64_t dv) { return sizeof(float) * (len * (dk + 1) + len * (dv + 1) + len); }
static size_t mha_smem_bwd(int64_t len, int64_t dk, int64_t dv) {
  return sizeof(float) * (len * (dk + 1) + 2 * len * (dv + 1) + len * (len + 1));
}

int mr_mha_core_fwd(const float* qk, const float* v, const float* mask, float* prob, float* ctx, int64_t n, int64_t len,
                    int64_t hn, int64_t dk, int64_t dv, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(qk && v && prob && ctx, MR_ERR_NULL, "mr_mha_core_fwd: null pointer");
  MR_REQUIRE(n >= 0 && len >= 1 && hn >= 1 && dk >= 1 && dv >= 1, MR_ERR_BAD_SHAPE, "mr_mha_core_fwd: bad shape");
  if (n == 0) return MR_OK;
  if (mha_attn_supported(len, dk, dv) && n * hn < (1ll << 31))
    return mha_attn_fwd(qk, hn * dk, v, hn * dv, mask, prob, ctx, hn * dv, n, len, hn, dk, dv, as_stream(stream));
  size_t smem = mha_smem_fwd(len, dk, dv);
  MR_REQUIRE(smem <= 200 * 1024, MR_ERR_UNSUPPORTED, "mr_mha_core_fwd: len=%lld too long", (long long)len);
  if (smem > 48 * 1024) cudaFuncSetAttribute(mha_core_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mha_core_fwd_kernel<<<(unsigned)(n * hn), MHA_THREADS, smem, as_stream(stream)>>>(qk, v, mask, prob, ctx, (int)len, (int)hn,
                                                                                    (int)dk, (int)dv);
  MR_CHECK_LAUNCH("mha_core_fwd_kernel");
  return MR_OK;
}

int mr_mha_core_bwd(const float* qk, const float* v, const float* prob, const float* d_ctx, float* d_qk, float* d_v,
                    int64_t n, int64_t len, int64_t hn, int64_t dk, int64_t dv, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(qk && v && prob && d_ctx && d_qk && d_v, MR_ERR_NULL, "mr_mha_core_bwd: null pointer");
  MR_REQUIRE(n >= 0 && len >= 1 && hn >= 1 && dk >= 1 && dv >= 1, MR_ERR_BAD_SHAPE, "mr_mha_core_bwd: bad shape");
  if (n == 0) return MR_OK;
  if (mha_attn_supported(len, dk, dv) && n * hn < (1ll << 31))
    return mha_attn_bwd(qk, hn * dk, v, hn * dv, prob, d_ctx, hn * dv, d_qk, hn * dk, d_v, hn * dv, n, len, hn, dk, dv, as_stream(stream));
  size_t smem = mha_smem_bwd(len, dk, dv);
  MR_REQUIRE(smem <= 200 * 1024, MR_ERR_UNSUPPORTED, "mr_mha_core_bwd: len=%lld too long", (long long)len);
  if (smem > 48 * 1024) cudaFuncSetAttribute(mha_core_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mha_core_bwd_kernel<<<(unsigned)(n * hn), MHA_THREADS, smem, as_stream(stream)>>>(qk, v, prob, d_ctx, d_qk, d_v, (int)len,
                                                                                    (int)hn, (int)dk, (int)dv);
  MR_CHECK_LAUNCH("mha_core_bwd_kernel");
  return MR_OK;
}

int mr_layernorm_fwd(const float* x, const float* gamma, const float* beta, const uint8_t* keep, float keep_scale, float* y,
                     float* mean, float* rstd, int64_t M, int64_t H, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(x && gamma && beta && y && mean && rstd, MR_ERR_NULL, "mr_layernorm_fwd: null pointer");
  MR_REQUIRE(M >= 0 && H >= 1, MR_ERR_BAD_SHAPE, "mr_layernorm_fwd: bad shape");
  if (M == 0) return MR_OK;
  layernorm_fwd_kernel<<<(unsigned)ceil_div(M, 8), 256, 0, as_stream(stream)>>>(x, gamma, beta, keep, keep_scale, y, mean,
                                                                                rstd, M, (int)H);
  MR_CHECK_LAUNCH("layernorm_fwd_kernel");
  return MR_OK;
}

int mr_layernorm_bwd(const float* x, const float* gamma, const uint8_t* keep, float keep_scale, const float* mean,
                     const float* rstd, const float* d_y, float* d_x, float* d_gamma_partial, float* d_beta_partial,
                     int64_t n_partial, int64_t M, int64_t H, void* stream) {
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(x && gamma && mean && rstd && d_y && d_x && d_gamma_partial && d_beta_partial, MR_ERR_NULL,
             "mr_layernorm_bwd: null pointer");
  MR_REQUIRE(M >= 0 && H >= 1 && n_partial >= 1 && H <= 2048, MR_ERR_BAD_SHAPE, "mr_layernorm_bwd: bad shape");
  size_t smem = sizeof(float) * 16 * H;
  if (smem > 48 * 1024) cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  layernorm_bwd_kernel<<<(unsigned)n_partial, 256, smem, as_stream(stream)>>>(x, gamma, keep, keep_scale, mean, rstd, d_y, d_x,
                                                                             d_gamma_partial, d_beta_partial, M, (int)H);
  MR_CHECK_LAUNCH("layernorm_bwd_kernel");
  return MR_OK;
}

}  // extern "C"
