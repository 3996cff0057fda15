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
  colsum_small_kernel<<<(unsigned)ceil_div(C, 32), 1024, 0, st>>>(x, ld, out, R, C);
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

__global__ void cast_pad_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows,
                                     int64_t cols, int64_t ld) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = rows * ld, stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int64_t r = i / ld, c = i - r * ld;
    dst[i] = __float2bfloat16(c < cols ? src[r * cols + c] : 0.f);
  }
}

}  // namespace mr

extern "C" {

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
