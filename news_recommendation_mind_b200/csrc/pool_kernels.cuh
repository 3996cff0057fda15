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

// backward: d_news [N,H] -> dkp (grad wrt proj pre-activation) and, when dc_pool != nullptr, dc_pool = p * d_news,
// both TO [.., ldo]; per-title partials of dq and (when dbq_partial != nullptr) of the projection-bias gradient
// sum_l dkp[l, :].

template <class TO> __device__ __forceinline__ TO from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

template <class T, class TO>
__global__ void __launch_bounds__(256)
cnn_pool_bwd_kernel(const T* __restrict__ c, const T* __restrict__ key, int64_t ld, const float* __restrict__ prob,
                    const float* __restrict__ q, const float* __restrict__ d_news, TO* __restrict__ dkp,
                    TO* __restrict__ dc_pool, int64_t ldo, float* __restrict__ dq_partial, float* __restrict__ dbq_partial,
                    int64_t N, int L, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float inv = rsqrtf((float)H);
  const T* kn = key + n * L * ld;
  const T* cn = c + n * L * ld;
  TO* dk = dkp + n * L * ldo;
  TO* dcp = dc_pool ? dc_pool + n * L * ldo : nullptr;
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
    float dq = 0.f, db = 0.f;
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
          const float o = dsl * qh * (1.f - k * k);
          db += o;
          dk[(int64_t)l * ldo + h] = from_f<TO>(o);
          if (dcp) dcp[(int64_t)l * ldo + h] = from_f<TO>(pl * dnh);
        }
      }
    }
    if (h < H) {
      dq_partial[n * H + h] = dq;
      if (dbq_partial) dbq_partial[n * H + h] = db;
    }
  }
  // padding columns [H, ldo) feed tensor-core GEMMs in the bf16 path: keep them exactly zero
  for (int64_t i = lane; i < (int64_t)L * (ldo - H); i += 32) {
    const int64_t l = i / (ldo - H), h = H + i % (ldo - H);
    dk[l * ldo + h] = from_f<TO>(0.f);
    if (dcp) dcp[l * ldo + h] = from_f<TO>(0.f);
  }
}

}  // namespace mr

// ------------------------------------------------------------------------------------------------------
// bf16 fast path (titles of at most 32 tokens, row pitch ld = Hp a multiple of 8 and <= 256 columns).
// One warp per title.  Row dot products run 4 rows at a time with 8 lanes per row and 16-byte loads (a row
// group reads 128 contiguous bytes), reduced with 3 shuffle levels; the pooled sum / the per-row outputs give
// each of the first Hp/8 lanes one 8-column piece, so every row is one coalesced 16-byte load per lane and
// the 32 row loads of a title are independent (memory-level parallelism instead of a shuffle chain per row).
// ------------------------------------------------------------------------------------------------------
namespace mr {

__device__ __forceinline__ void bf8_to_f(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__device__ __forceinline__ uint4 f_to_bf8(const float* f) {
  uint4 o;
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(f[4], f[5]), d = __floats2bfloat162_rn(f[6], f[7]);
  o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
  return o;
}

// row scores: out (valid in every lane l < L) = sum_h vec[h] * x[l, h];  vec is given as this lane's slices
// for the "8 lanes per row" mapping: vr[i][0..7] = vec[(part + 8 i) * 8 ..], part = lane & 7
// MASK: also writes one byte per 8-column piece to mask_out[l * 32 + piece], bit e = (x[l, 8 piece + e] > 0)
template <int MAXP, bool MASK = false>     // MAXP = ceil(pieces / 8) <= 4
__device__ __forceinline__ float row_dots_bf16(const __nv_bfloat16* __restrict__ x, int64_t ld, int L, int pieces,
                                               const float (&vr)[MAXP][8], int lane, uint8_t* __restrict__ mask_out = nullptr) {
  const int rr = lane >> 3, part = lane & 7;
  float mine = 0.f;
  for (int it = 0; it < 8; ++it) {
    const int l = it * 4 + rr;
    float d = 0.f;
    if (l < L) {
      const uint4* row = reinterpret_cast<const uint4*>(x + (int64_t)l * ld);
      uint4 v[MAXP];
#pragma unroll
      for (int i = 0; i < MAXP; ++i)
        if (part + 8 * i < pieces) v[i] = __ldg(row + part + 8 * i);
#pragma unroll
      for (int i = 0; i < MAXP; ++i)
        if (part + 8 * i < pieces) {
          float f[8];
          bf8_to_f(v[i], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) d = fmaf(vr[i][e], f[e], d);
          if (MASK) {
            uint32_t bits = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) bits |= (f[e] > 0.f ? 1u : 0u) << e;
            mask_out[(int64_t)l * 32 + part + 8 * i] = (uint8_t)bits;
          }
        }
    }
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    const float sv = __shfl_sync(0xffffffffu, d, (lane & 3) * 8);
    if ((lane >> 2) == it) mine = sv;
  }
  return mine;
}

template <int MAXP>
__global__ void __launch_bounds__(256)
cnn_pool_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ c, const __nv_bfloat16* __restrict__ key, int64_t ld,
                         const void* __restrict__ mask, int mask_i64, const float* __restrict__ q, float* __restrict__ prob,
                         float* __restrict__ news, int64_t N, int L, int H) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int pieces = (int)(ld / 8);
  const __nv_bfloat16* kn = key + n * L * ld;
  const __nv_bfloat16* cn = c + n * L * ld;
  float qr[MAXP][8];
#pragma unroll
  for (int i = 0; i < MAXP; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int h = ((lane & 7) + 8 * i) * 8 + e;
      qr[i][e] = h < H ? __ldg(q + h) : 0.f;
    }
  const float s = row_dots_bf16<MAXP>(kn, ld, L, pieces, qr, lane) * rsqrtf((float)H);
  const bool keep = lane < L && (mask ? (load_index(mask, mask_i64, n * L + lane) != 0) : true);
  const float mx = warp_max(keep ? s : -INFINITY);
  const float ex = keep ? expf(s - mx) : 0.f;
  const float sum = warp_sum(ex);
  const float p = sum > 0.f ? ex / sum : 0.f;             // all-masked title -> zeros (XSoftmax, Attention.py:66-74)
  if (lane < L) prob[n * L + lane] = p;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const uint4* cp = reinterpret_cast<const uint4*>(cn) + lane;
#pragma unroll 8
  for (int l = 0; l < L; ++l) {
    const float pl = __shfl_sync(0xffffffffu, p, l);
    if (lane < pieces) {
      float f[8];
      bf8_to_f(__ldg(cp + (int64_t)l * pieces), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(pl, f[e], acc[e]);
    }
  }
  if (lane < pieces) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int h = lane * 8 + e;
      if (h < H) news[n * H + h] = acc[e];
    }
  }
}

// Backward.  Writes dkp = grad wrt the projection pre-activation (bf16 [T, ld]), dnp = d_news padded to the row
// pitch ld (fp32 [N, ld]) and cmask = one bit per element of c (c > 0; 32 bytes per token row, ld <= 256): the
// RELUGRAD_POOL epilogue of the tap GEMM forms relu'(c) * (p[t] * d_news[n, :] + ...) from them -- dc_pool is never
// materialised and c is not read again.  The gradients of the pooling query and of the projection bias
// (sum over tokens of dkp, taken before the bf16 rounding) are reduced over the CTA's 8 titles in a fixed order and
// written as one partial row per CTA: part[blockIdx][0][ld] = dq, part[blockIdx][1][ld] = dbq.
template <int MAXP>
__global__ void __launch_bounds__(256)
cnn_pool_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ c, const __nv_bfloat16* __restrict__ key, int64_t ld,
                         const float* __restrict__ prob, const float* __restrict__ q, const float* __restrict__ d_news,
                         __nv_bfloat16* __restrict__ dkp, float* __restrict__ dnp, float* __restrict__ part,
                         uint8_t* __restrict__ cmask, int64_t N, int L, int H) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][2][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  const int pieces = (int)(ld / 8);
  float dq[8], db[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { dq[e] = 0.f; db[e] = 0.f; }
  if (n < N) {
    const __nv_bfloat16* kn = key + n * L * ld;
    const __nv_bfloat16* cn = c + n * L * ld;
    const float* dn = d_news + n * H;
    float dr[MAXP][8];
#pragma unroll
    for (int i = 0; i < MAXP; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int h = ((lane & 7) + 8 * i) * 8 + e;
        dr[i][e] = h < H ? __ldg(dn + h) : 0.f;
      }
    // <d_news, c[l]>; the same pass over c writes its sign mask
    const float dp = row_dots_bf16<MAXP, true>(cn, ld, L, pieces, dr, lane, cmask + n * L * 32);
    const float p = lane < L ? prob[n * L + lane] : 0.f;
    const float dot = warp_sum(p * dp);
    const float ds = p * (dp - dot) * rsqrtf((float)H);                       // softmax backward (Attention.py:77-80)
    // this lane's 8-column piece of q and d_news
    float qv[8], dv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int h = lane * 8 + e;
      qv[e] = (lane < pieces && h < H) ? __ldg(q + h) : 0.f;
      dv[e] = (lane < pieces && h < H) ? __ldg(dn + h) : 0.f;
    }
    if (lane < pieces) {
      float4* o = reinterpret_cast<float4*>(dnp + n * ld + lane * 8);
      o[0] = make_float4(dv[0], dv[1], dv[2], dv[3]);
      o[1] = make_float4(dv[4], dv[5], dv[6], dv[7]);
    }
    const uint4* kp = reinterpret_cast<const uint4*>(kn) + lane;
    uint4* dko = reinterpret_cast<uint4*>(dkp + n * L * ld) + lane;
#pragma unroll 4
    for (int l = 0; l < L; ++l) {
      const float dsl = __shfl_sync(0xffffffffu, ds, l);
      if (lane < pieces) {
        float k[8], o1[8];
        bf8_to_f(__ldg(kp + (int64_t)l * pieces), k);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          dq[e] = fmaf(dsl, k[e], dq[e]);
          o1[e] = dsl * qv[e] * (1.f - k[e] * k[e]);
          db[e] += o1[e];
        }
        dko[(int64_t)l * pieces] = f_to_bf8(o1);
      }
    }
  }
  if (lane < pieces) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[warp][0][lane * 8 + e] = dq[e];
      red[warp][1][lane * 8 + e] = db[e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * (int)ld; i += 256) {
    const int which = i >= (int)ld, col = i - which * (int)ld;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][which][col];
    part[(int64_t)blockIdx.x * 2 * ld + i] = s;
  }
}

// part [rows][2][ld] -> d_query[h] = sum_rows part[.][0][h],  d_proj_b[h] = sum_rows part[.][1][h]   (h < H), fixed order
static __global__ void __launch_bounds__(1024)
cnn_pool_bwd_final_kernel(const float* __restrict__ part, int64_t rows, int ld, int H, float* __restrict__ d_query,
                          float* __restrict__ d_proj_b) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sm[32][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (col < 2 * ld)
    for (int64_t r = ry; r < rows; r += 32) s += part[r * 2 * ld + col];
  sm[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && col < 2 * ld) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += sm[i][cx];
    const int which = col >= ld, h = col - which * ld;
    if (h < H) (which ? d_proj_b : d_query)[h] = t;
  }
}

}  // namespace mr
