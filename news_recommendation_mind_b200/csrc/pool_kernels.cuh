// Additive-attention pooling of the CNN news encoder (CNN.py:44-46 -> Attention.py:5-30,56-80),
// shared by the fp32 and the bf16 path (templated on the storage type of c / key).
//   s[l]   = <q, key[l,:]> / sqrt(H)
//   p      = masked softmax over l  (all-masked title -> all zeros, XSoftmax semantics)
//   news   = sum_l p[l] c[l,:]
// One warp per title; warp-shuffle max / sum for the softmax.  L <= 32*PL_MAXR.
#pragma once
#include "common.cuh"

namespace mr {

constexpr int PL_MAXR = 8;   // titles up to 256 tokens (reference max signal_length is 512 but
                             // TwoTower runs 30..100; checked on the host)

template <class T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <class T>
__global__ void __launch_bounds__(256)
cnn_pool_fwd_kernel(const T* __restrict__ c, const T* __restrict__ key, int64_t ld, const void* __restrict__ mask,
                    int mask_i64, const float* __restrict__ q, float* __restrict__ prob, float* __restrict__ news,
                    int64_t N, int L, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float inv = rsqrtf((float)H);
  const T* kn = key + n * L * ld;
  const T* cn = c + n * L * ld;
  float s[PL_MAXR];
  bool keep[PL_MAXR];
#pragma unroll
  for (int r = 0; r < PL_MAXR; ++r) { s[r] = 0.f; keep[r] = false; }
  for (int l = 0; l < L; ++l) {
    float d = 0.f;
    for (int h = lane; h < H; h += 32) d = fmaf(__ldg(q + h), to_f(kn[(int64_t)l * ld + h]), d);
    d = warp_sum(d) * inv;
    if ((l & 31) == lane) {
#pragma unroll
      for (int r = 0; r < PL_MAXR; ++r)
        if (r == (l >> 5)) {
          s[r] = d;
          keep[r] = mask ? (load_index(mask, mask_i64, n * L + l) != 0) : true;
        }
    }
  }
  float mx = -INFINITY;
#pragma unroll
  for (int r = 0; r < PL_MAXR; ++r)
    if (keep[r]) mx = fmaxf(mx, s[r]);
  mx = warp_max(mx);
  float e[PL_MAXR], sum = 0.f;
#pragma unroll
  for (int r = 0; r < PL_MAXR; ++r) { e[r] = keep[r] ? expf(s[r] - mx) : 0.f; sum += e[r]; }
  sum = warp_sum(sum);
  const float rs = sum > 0.f ? 1.f / sum : 0.f;
#pragma unroll
  for (int r = 0; r < PL_MAXR; ++r) {
    e[r] *= rs;
    int l = r * 32 + lane;
    if (l < L) prob[n * L + l] = e[r];
  }
  for (int h0 = 0; h0 < H; h0 += 32) {
    int h = h0 + lane;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < PL_MAXR; ++r) {
      if (r * 32 >= L) break;
      for (int j = 0; j < 32 && r * 32 + j < L; ++j) {
        float pj = __shfl_sync(0xffffffffu, e[r], j);
        if (h < H) acc = fmaf(pj, to_f(cn[(int64_t)(r * 32 + j) * ld + h]), acc);
      }
    }
    if (h < H) news[n * H + h] = acc;
  }
}

// backward: d_news [N,H] -> dkp (grad wrt proj pre-activation) and dc_pool, both TO [.., ldo];
// dq partial per title.
template <class T, class TO>
__global__ void __launch_bounds__(256)
cnn_pool_bwd_kernel(const T* __restrict__ c, const T* __restrict__ key, int64_t ld, const float* __restrict__ prob,
                    const float* __restrict__ q, const float* __restrict__ d_news, TO* __restrict__ dkp,
                    TO* __restrict__ dc_pool, int64_t ldo, float* __restrict__ dq_partial, int64_t N, int L, int H);

template <class TO> __device__ __forceinline__ TO from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

template <class T, class TO>
__global__ void __launch_bounds__(256)
cnn_pool_bwd_kernel(const T* __restrict__ c, const T* __restrict__ key, int64_t ld, const float* __restrict__ prob,
                    const float* __restrict__ q, const float* __restrict__ d_news, TO* __restrict__ dkp,
                    TO* __restrict__ dc_pool, int64_t ldo, float* __restrict__ dq_partial, int64_t N, int L, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float inv = rsqrtf((float)H);
  const T* kn = key + n * L * ld;
  const T* cn = c + n * L * ld;
  TO* dk = dkp + n * L * ldo;
  TO* dcp = dc_pool + n * L * ldo;
  const float* dn = d_news + n * H;
  // dp[l] = <d_news, c[l]>, kept in lane l%32 / register l/32
  float p[PL_MAXR], dp[PL_MAXR];
#pragma unroll
  for (int r = 0; r < PL_MAXR; ++r) {
    int l = r * 32 + lane;
    p[r] = l < L ? prob[n * L + l] : 0.f;
    dp[r] = 0.f;
  }
  for (int l = 0; l < L; ++l) {
    float d = 0.f;
    for (int h = lane; h < H; h += 32) d = fmaf(__ldg(dn + h), to_f(cn[(int64_t)l * ld + h]), d);
    d = warp_sum(d);
    if ((l & 31) == lane) {
#pragma unroll
      for (int r = 0; r < PL_MAXR; ++r)
        if (r == (l >> 5)) dp[r] = d;
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int r = 0; r < PL_MAXR; ++r) dot = fmaf(p[r], dp[r], dot);
  dot = warp_sum(dot);
  float ds[PL_MAXR];
#pragma unroll
  for (int r = 0; r < PL_MAXR; ++r) ds[r] = p[r] * (dp[r] - dot) * inv;   // softmax bwd (Attention.py:79) and 1/sqrt(H)
  for (int h0 = 0; h0 < H; h0 += 32) {
    int h = h0 + lane;
    float qh = h < H ? __ldg(q + h) : 0.f;
    float dnh = h < H ? __ldg(dn + h) : 0.f;
    float dq = 0.f;
#pragma unroll
    for (int r = 0; r < PL_MAXR; ++r) {
      if (r * 32 >= L) break;
      for (int j = 0; j < 32 && r * 32 + j < L; ++j) {
        int l = r * 32 + j;
        float dsl = __shfl_sync(0xffffffffu, ds[r], j);
        float pl = __shfl_sync(0xffffffffu, p[r], j);
        if (h < H) {
          float k = to_f(kn[(int64_t)l * ld + h]);
          dq = fmaf(dsl, k, dq);
          dk[(int64_t)l * ldo + h] = from_f<TO>(dsl * qh * (1.f - k * k));
          dcp[(int64_t)l * ldo + h] = from_f<TO>(pl * dnh);
        }
      }
    }
    if (h < H) dq_partial[n * H + h] = dq;
  }
  // padding columns [H, ldo) feed tensor-core GEMMs in the bf16 path: keep them exactly zero
  for (int64_t i = lane; i < (int64_t)L * (ldo - H); i += 32) {
    const int64_t l = i / (ldo - H), h = H + i % (ldo - H);
    dk[l * ldo + h] = from_f<TO>(0.f);
    dcp[l * ldo + h] = from_f<TO>(0.f);
  }
}

}  // namespace mr
