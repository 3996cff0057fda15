// Small shared kernels: column sums, casts, Adam.
#include "gemm_simt.cuh"

namespace mr {

// row chunks of the first level: enough CTAs to fill the GPU (column groups x chunks), few enough that the
// fixed-order second level (one thread per column walking the chunks) stays a few microseconds
int64_t colsum_chunks(int64_t R) {
  int64_t c = ceil_div(R, 512);
  return c < 1 ? 1 : (c > 96 ? 96 : c);
}

template <class T> __device__ __forceinline__ float cs_load(const T* p);
template <> __device__ __forceinline__ float cs_load<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float cs_load<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <class T>
__global__ void colsum_partial_kernel(const T* __restrict__ x, int64_t ld, float* __restrict__ partial, int64_t R, int64_t C,
                                      int64_t rows_per_chunk) {
  __shared__ float sm[8][33];
  const int cx = threadIdx.x, ry = threadIdx.y;
  const int64_t c = (int64_t)blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = min(R, r0 + rows_per_chunk);
  float s = 0.f;
  if (c < C)
    for (int64_t r = r0 + ry; r < r1; r += 8) s += cs_load<T>(x + r * ld + c);
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][cx];
    partial[(int64_t)blockIdx.y * C + c] = t;
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ partial, float* __restrict__ out, int64_t chunks, int64_t C) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int64_t i = 0; i < chunks; ++i) s += partial[i * C + c];
  out[c] = s;
}

template <class T>
static cudaError_t colsum_t(const T* x, int64_t ld, float* out, int64_t R, int64_t C, float* partial, cudaStream_t st) {
  int64_t chunks = colsum_chunks(R);
  int64_t rpc = ceil_div(R, chunks);
  dim3 grid((unsigned)ceil_div(C, 32), (unsigned)chunks), block(32, 8);
  colsum_partial_kernel<T><<<grid, block, 0, st>>>(x, ld, partial, R, C, rpc);
  colsum_final_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, st>>>(partial, out, chunks, C);
  count_launch(2);
  return cudaGetLastError();
}
// single launch, for a few thousand rows at most (per-CTA partials of a producer kernel): one CTA per 32 columns,
// 8 row groups, fixed summation order
__global__ void __launch_bounds__(1024) colsum_small_kernel(const float* __restrict__ x, int64_t ld, float* __restrict__ out, int64_t R,
                                                            int64_t C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sm[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < C)
    for (int64_t r = ry; r < R; r += 32) s += x[r * ld + c];
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sm[i][cx];
    out[c] = t;
  }
}
cudaError_t colsum_small(const float* x, int64_t ld, float* out, int64_t R, int64_t C, cudaStream_t st) {
  launch_pdl(colsum_small_kernel, dim3((unsigned)ceil_div(C, 32)), dim3(1024), 0, st, x, ld, out, R, C);
  count_launch(1);
  return cudaGetLastError();
}
cudaError_t colsum(const float* x, float* out, int64_t R, int64_t C, float* partial, cudaStream_t st) {
  return colsum_t<float>(x, C, out, R, C, partial, st);
}
cudaError_t colsum_bf16(const __nv_bfloat16* x, int64_t ld, float* out, int64_t R, int64_t C, float* partial, cudaStream_t st) {
  return colsum_t<__nv_bfloat16>(x, ld, out, R, C, partial, st);
}

// ---- Adam (utils/Manager.py:404-413 -> torch.optim.Adam defaults) -----------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr_over_bc1, float inv_sqrt_bc2, float beta1,
                            float omb1, float beta2, float omb2, float eps, float grad_scale,
                            __nv_bfloat16* __restrict__ shadow, int64_t row_len, int64_t shadow_ld) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float gi = g[i] * grad_scale;
    float mi = beta1 * m[i] + omb1 * gi;
    float vi = beta2 * v[i] + omb2 * gi * gi;
    float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    float pi = p[i] - lr_over_bc1 * (mi / denom);
    m[i] = mi; v[i] = vi; p[i] = pi;
    if (shadow) {
      int64_t r = i / row_len, c = i - r * row_len;
      shadow[r * shadow_ld + c] = __float2bfloat16(pi);
    }
  }
}

// ---- multi-tensor Adam: every parameter of the model in ONE launch -------------------------------------------
constexpr int AD_MAX_TENSORS = 24;
constexpr int AD_CHUNK = 4096;          // elements per CTA (256 threads x 4 float4)
struct AdamMultiArgs {
  float* p[AD_MAX_TENSORS];
  const float* g[AD_MAX_TENSORS];
  float* m[AD_MAX_TENSORS];
  float* v[AD_MAX_TENSORS];
  int64_t n[AD_MAX_TENSORS];
  float lr_over_bc1[AD_MAX_TENSORS];
  int32_t blk_off[AD_MAX_TENSORS + 1];  // first CTA of tensor i
  int n_tensors;
  float inv_sqrt_bc2, beta1, omb1, beta2, omb2, eps, grad_scale;
  const float* dyn;                     // optional device block {1 / (1 - beta1^t), 1 / sqrt(1 - beta2^t), grad_scale, unused, lr[0..n_tensors)}:
                                        // when set, every step-dependent scalar (bias corrections, 1/world, the per-tensor learning rates a
                                        // schedule moves) is read here, so that one captured launch stays valid for every step (CUDA-graph replay)
  int shadow_tensor;                    // tensor whose bf16 shadow is refreshed (-1 = none)
  __nv_bfloat16* shadow; int64_t row_len, shadow_ld;
};

__device__ __forceinline__ float adam_one(float pi, float gi, float& mi, float& vi, const AdamMultiArgs& a, float lr, float inv_sqrt_bc2,
                                          float grad_scale) {
  gi *= grad_scale;
  mi = a.beta1 * mi + a.omb1 * gi;
  vi = a.beta2 * vi + a.omb2 * gi * gi;
  const float denom = sqrtf(vi) * inv_sqrt_bc2 + a.eps;
  return pi - lr * (mi / denom);
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const __grid_constant__ AdamMultiArgs a) {
  pdl_trigger();
  pdl_wait();
  int t = 0;
#pragma unroll 1
  while (t + 1 < a.n_tensors && (int)blockIdx.x >= a.blk_off[t + 1]) ++t;
  const int64_t n = a.n[t];
  float* __restrict__ p = a.p[t];
  const float* __restrict__ g = a.g[t];
  float* __restrict__ m = a.m[t];
  float* __restrict__ v = a.v[t];
  const float lr = a.dyn != nullptr ? __ldg(a.dyn + 4 + t) * __ldg(a.dyn) : a.lr_over_bc1[t];
  const float isb2 = a.dyn != nullptr ? __ldg(a.dyn + 1) : a.inv_sqrt_bc2;
  const float gs = a.dyn != nullptr ? __ldg(a.dyn + 2) : a.grad_scale;
  const bool sh = t == a.shadow_tensor;
  const int64_t base = (int64_t)(blockIdx.x - a.blk_off[t]) * AD_CHUNK;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                         reinterpret_cast<uintptr_t>(v)) & 15) == 0;          // gradients may be views into a flat all-reduce buffer
  const bool vec = aligned && (n & 3) == 0 && (!sh || (a.row_len & 3) == 0);
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int64_t i = base + ((int64_t)it * 256 + threadIdx.x) * 4;
    if (i >= n) break;
    if (vec) {
      const float4 g4 = *reinterpret_cast<const float4*>(g + i);
      float4 p4 = *reinterpret_cast<float4*>(p + i), m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i);
      p4.x = adam_one(p4.x, g4.x, m4.x, v4.x, a, lr, isb2, gs);
      p4.y = adam_one(p4.y, g4.y, m4.y, v4.y, a, lr, isb2, gs);
      p4.z = adam_one(p4.z, g4.z, m4.z, v4.z, a, lr, isb2, gs);
      p4.w = adam_one(p4.w, g4.w, m4.w, v4.w, a, lr, isb2, gs);
      *reinterpret_cast<float4*>(p + i) = p4;
      *reinterpret_cast<float4*>(m + i) = m4;
      *reinterpret_cast<float4*>(v + i) = v4;
      if (sh) {
        const int64_t r = i / a.row_len, c = i - r * a.row_len;
        __nv_bfloat162 lo = __floats2bfloat162_rn(p4.x, p4.y), hi = __floats2bfloat162_rn(p4.z, p4.w);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(a.shadow + r * a.shadow_ld + c) = o;
      }
    } else {
      for (int e = 0; e < 4 && i + e < n; ++e) {
        float mi = m[i + e], vi = v[i + e];
        const float pi = adam_one(p[i + e], g[i + e], mi, vi, a, lr, isb2, gs);
        p[i + e] = pi; m[i + e] = mi; v[i + e] = vi;
        if (sh) {
          const int64_t r = (i + e) / a.row_len, c = (i + e) - r * a.row_len;
          a.shadow[r * a.shadow_ld + c] = __float2bfloat16(pi);
        }
      }
    }
  }
}

__global__ void cast_pad_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows,
                                     int64_t cols, int64_t ld) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = rows * ld, stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int64_t r = i / ld, c = i - r * ld;
    dst[i] = __float2bfloat16(c < cols ? src[r * cols + c] : 0.f);
  }
}

// ---- title rows by news id (device-resident replacement of the per-sample token assembly of utils/MIND.py:347-355) ----
// one warp per output row: out_ids[r, :] = tok_ids[nid[r], :], out_mask[r, :] = tok_mask[nid[r], :]; rows 0..n_a-1 come from
// nid_a (candidates), rows n_a..n_a+n_b-1 from nid_b (clicked history); an id outside [0, n_rows) reads row 0 (the empty article)
__global__ void __launch_bounds__(256) gather_titles_kernel(const int32_t* __restrict__ tok_ids, const int32_t* __restrict__ tok_mask,
                                                            int64_t n_rows, int L, const void* __restrict__ nid_a, int64_t n_a,
                                                            const void* __restrict__ nid_b, int64_t n_b, int nid_i64,
                                                            int32_t* __restrict__ out_ids, int32_t* __restrict__ out_mask) {
  pdl_trigger();
  pdl_wait();
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= n_a + n_b) return;
  int64_t nid = r < n_a ? load_index(nid_a, nid_i64, r) : load_index(nid_b, nid_i64, r - n_a);
  if (nid < 0 || nid >= n_rows) nid = 0;
  const int32_t* si = tok_ids + nid * L;
  const int32_t* sm_ = tok_mask + nid * L;
  for (int l = threadIdx.x & 31; l < L; l += 32) {
    out_ids[r * L + l] = __ldg(si + l);
    out_mask[r * L + l] = __ldg(sm_ + l);
  }
}

}  // namespace mr

extern "C" {

int mr_gather_titles(const int32_t* tok_ids, const int32_t* tok_mask, int64_t n_rows, int64_t L, const void* nid_a, int64_t n_a,
                     const void* nid_b, int64_t n_b, int nid_i64, int32_t* out_ids, int32_t* out_mask, void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(tok_ids && tok_mask && out_ids && out_mask, MR_ERR_NULL, "mr_gather_titles: null pointer");
  MR_REQUIRE(n_rows >= 1 && L >= 1 && L <= 4096 && n_a >= 0 && n_b >= 0, MR_ERR_BAD_SHAPE, "mr_gather_titles: rows=%lld L=%lld n_a=%lld n_b=%lld",
             (long long)n_rows, (long long)L, (long long)n_a, (long long)n_b);
  MR_REQUIRE((n_a == 0 || nid_a) && (n_b == 0 || nid_b), MR_ERR_NULL, "mr_gather_titles: null id list");
  if (n_a + n_b == 0) return MR_OK;
  launch_pdl(gather_titles_kernel, dim3((unsigned)ceil_div(n_a + n_b, 8)), dim3(256), 0, as_stream(stream), tok_ids, tok_mask, n_rows, (int)L,
             nid_a, n_a, nid_b, n_b, nid_i64, out_ids, out_mask);
  MR_CHECK_LAUNCH("gather_titles_kernel");
  return MR_OK;
}

int mr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, int64_t step, double lr, double beta1,
                 double beta2, double eps, double grad_scale, void* shadow_bf16, int64_t row_len, int64_t shadow_ld,
                 void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(p && g && m && v, MR_ERR_NULL, "mr_adam_step: null pointer");
  MR_REQUIRE(n >= 0 && step >= 1, MR_ERR_BAD_SHAPE, "mr_adam_step: n=%lld step=%lld", (long long)n, (long long)step);
  if (n == 0) return MR_OK;
  if (shadow_bf16) MR_REQUIRE(row_len > 0 && shadow_ld >= row_len && n % row_len == 0, MR_ERR_BAD_SHAPE,
                              "mr_adam_step: bad shadow geometry");
  double bc1 = 1.0 - pow(beta1, (double)step);
  double bc2 = 1.0 - pow(beta2, (double)step);
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  adam_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(p, g, m, v, n, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)),
                                                              (float)beta1, (float)(1.0 - beta1), (float)beta2,
                                                              (float)(1.0 - beta2), (float)eps, (float)grad_scale,
                                                              static_cast<__nv_bfloat16*>(shadow_bf16), row_len, shadow_ld);
  MR_CHECK_LAUNCH("adam_kernel");
  return MR_OK;
}

int mr_adam_step_multi(int n_tensors, float* const* p, const float* const* g, float* const* m, float* const* v,
                       const int64_t* numel, const double* lr, int64_t step, double beta1, double beta2, double eps,
                       double grad_scale, int shadow_tensor, void* shadow_bf16, int64_t row_len, int64_t shadow_ld, void* stream) {
  return mr_adam_step_multi_dyn(n_tensors, p, g, m, v, numel, lr, step, beta1, beta2, eps, grad_scale, shadow_tensor, shadow_bf16, row_len,
                                shadow_ld, nullptr, stream);
}

int mr_adam_step_multi_dyn(int n_tensors, float* const* p, const float* const* g, float* const* m, float* const* v,
                           const int64_t* numel, const double* lr, int64_t step, double beta1, double beta2, double eps,
                           double grad_scale, int shadow_tensor, void* shadow_bf16, int64_t row_len, int64_t shadow_ld,
                           const float* dyn_device, void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(n_tensors >= 0 && n_tensors <= AD_MAX_TENSORS, MR_ERR_BAD_SHAPE, "mr_adam_step_multi: %d tensors (max %d per call)", n_tensors,
             AD_MAX_TENSORS);
  MR_REQUIRE(step >= 1, MR_ERR_BAD_SHAPE, "mr_adam_step_multi: step=%lld", (long long)step);
  if (n_tensors == 0) return MR_OK;
  MR_REQUIRE(p && g && m && v && numel && lr, MR_ERR_NULL, "mr_adam_step_multi: null pointer");
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  AdamMultiArgs a{};
  int64_t blocks = 0;
  for (int i = 0; i < n_tensors; ++i) {
    MR_REQUIRE(p[i] && g[i] && m[i] && v[i] && numel[i] >= 0, MR_ERR_NULL, "mr_adam_step_multi: tensor %d", i);
    a.p[i] = p[i]; a.g[i] = g[i]; a.m[i] = m[i]; a.v[i] = v[i]; a.n[i] = numel[i];
    a.lr_over_bc1[i] = dyn_device != nullptr ? (float)lr[i] : (float)(lr[i] / bc1);
    a.blk_off[i] = (int32_t)blocks;
    blocks += ceil_div(numel[i], (int64_t)AD_CHUNK);
    MR_REQUIRE(blocks < (1ll << 30), MR_ERR_BAD_SHAPE, "mr_adam_step_multi: too many elements");
  }
  a.blk_off[n_tensors] = (int32_t)blocks;
  a.n_tensors = n_tensors;
  a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  a.beta1 = (float)beta1; a.omb1 = (float)(1.0 - beta1); a.beta2 = (float)beta2; a.omb2 = (float)(1.0 - beta2);
  a.eps = (float)eps; a.grad_scale = (float)grad_scale;
  a.dyn = dyn_device;
  a.shadow_tensor = shadow_bf16 ? shadow_tensor : -1;
  a.shadow = static_cast<__nv_bfloat16*>(shadow_bf16); a.row_len = row_len; a.shadow_ld = shadow_ld;
  if (a.shadow_tensor >= 0)
    MR_REQUIRE(shadow_tensor < n_tensors && row_len > 0 && shadow_ld >= row_len && numel[shadow_tensor] % row_len == 0, MR_ERR_BAD_SHAPE,
               "mr_adam_step_multi: bad shadow geometry");
  if (blocks == 0) return MR_OK;
  launch_pdl(adam_multi_kernel, dim3((unsigned)blocks), dim3(256), 0, as_stream(stream), a);
  MR_CHECK_LAUNCH("adam_multi_kernel");
  return MR_OK;
}

int mr_cast_pad_bf16(const float* src, void* dst, int64_t rows, int64_t cols, int64_t ld, void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(src && dst, MR_ERR_NULL, "mr_cast_pad_bf16: null pointer");
  MR_REQUIRE(rows >= 0 && cols > 0 && ld >= cols, MR_ERR_BAD_SHAPE, "mr_cast_pad_bf16: bad shape");
  if (rows == 0) return MR_OK;
  int64_t blocks = ceil_div(rows * ld, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  cast_pad_bf16_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(src, static_cast<__nv_bfloat16*>(dst), rows, cols, ld);
  MR_CHECK_LAUNCH("cast_pad_bf16_kernel");
  return MR_OK;
}

}  // extern "C"
