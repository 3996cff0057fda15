// fp32 SIMT GEMM used by the MR_F32 (verification) path and by the small GEMMs of both paths.
//
//   C[m,n] = sum_k A(m,k) * B(k,n)         m < M, n < N, k in this split's K range
//
// A and B are supplied as functors so the same kernel serves plain matrices, the gathered /
// tap-shifted im2col view of the title tokens (conv forward, dgrad, wgrad) and transposed views.
// Tile 128x64x16, 256 threads, 8x4 outputs per thread, smem double buffered by register prefetch.
// Split-K: blockIdx.z owns K range [z*k_split, (z+1)*k_split) and writes a partial tile to
// `partial[z][M][N]`; splitk_reduce() then sums the partials in a fixed order (bit-reproducible,
// no atomics) and applies the epilogue.
#pragma once
#include "common.cuh"

namespace mr {

constexpr int GBM = 128, GBN = 64, GBK = 16, GTM = 8, GTN = 4, GTHREADS = 256;

// Loader concept:  float operator()(int64 m_or_k, int64 k_or_n) const;  static bool FAST_IS_K
//   A loaders: (m,k) -> value, A_K_FAST says consecutive k are contiguous in memory
//   B loaders: (k,n) -> value, B_N_FAST says consecutive n are contiguous in memory

template <class ALoad, class BLoad, class Epi, bool A_K_FAST, bool B_N_FAST>
__global__ void __launch_bounds__(GTHREADS)
gemm_simt_kernel(int64_t M, int64_t N, int64_t K, int64_t k_split, ALoad a, BLoad b, Epi epi, float* partial) {
  __shared__ float As[GBK][GBM + 4];
  __shared__ float Bs[GBK][GBN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * GBM;
  const int64_t n0 = (int64_t)blockIdx.y * GBN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_split;
  const int64_t kend = min(K, kbeg + k_split);
  const int ty = tid / (GBN / GTN);   // 0..15 -> rows ty*8
  const int tx = tid % (GBN / GTN);   // 0..15 -> cols tx*4

  float acc[GTM][GTN];
#pragma unroll
  for (int i = 0; i < GTM; ++i)
#pragma unroll
    for (int j = 0; j < GTN; ++j) acc[i][j] = 0.f;

  float ra[GBM * GBK / GTHREADS];     // 8
  float rb[GBN * GBK / GTHREADS];     // 4

  auto fetch = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < GBM * GBK / GTHREADS; ++i) {
      int e = tid + i * GTHREADS;
      int mm, kk;
      if (A_K_FAST) { mm = e / GBK; kk = e % GBK; } else { kk = e / GBM; mm = e % GBM; }
      int64_t m = m0 + mm, k = k0 + kk;
      ra[i] = (m < M && k < kend) ? a(m, k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < GBN * GBK / GTHREADS; ++i) {
      int e = tid + i * GTHREADS;
      int nn, kk;
      if (B_N_FAST) { kk = e / GBN; nn = e % GBN; } else { nn = e / GBK; kk = e % GBK; }
      int64_t n = n0 + nn, k = k0 + kk;
      rb[i] = (n < N && k < kend) ? b(k, n) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < GBM * GBK / GTHREADS; ++i) {
      int e = tid + i * GTHREADS;
      int mm, kk;
      if (A_K_FAST) { mm = e / GBK; kk = e % GBK; } else { kk = e / GBM; mm = e % GBM; }
      As[kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < GBN * GBK / GTHREADS; ++i) {
      int e = tid + i * GTHREADS;
      int nn, kk;
      if (B_N_FAST) { kk = e / GBN; nn = e % GBN; } else { nn = e / GBK; kk = e % GBK; }
      Bs[kk][nn] = rb[i];
    }
  };

  if (kbeg < kend) fetch(kbeg);
  for (int64_t k0 = kbeg; k0 < kend; k0 += GBK) {
    stash();
    __syncthreads();
    if (k0 + GBK < kend) fetch(k0 + GBK);
#pragma unroll
    for (int kk = 0; kk < GBK; ++kk) {
      float av[GTM], bv[GTN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * GTM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * GTM + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * GTN]);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
      av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
      bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
#pragma unroll
      for (int i = 0; i < GTM; ++i)
#pragma unroll
        for (int j = 0; j < GTN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < GTM; ++i) {
    int64_t m = m0 + ty * GTM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < GTN; ++j) {
      int64_t n = n0 + tx * GTN + j;
      if (n >= N) continue;
      if (partial != nullptr) partial[((int64_t)blockIdx.z * M + m) * N + n] = acc[i][j];
      else epi(m, n, acc[i][j]);
    }
  }
}

template <class Epi>
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int64_t MN, int64_t N, int splits, Epi epi) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(int64_t)z * MN + i];
  epi(i / N, i % N, s);
}

// Host launcher.  `splits` > 1 needs `partial` with splits*M*N floats.
template <bool A_K_FAST, bool B_N_FAST, class ALoad, class BLoad, class Epi>
inline cudaError_t gemm_simt(int64_t M, int64_t N, int64_t K, ALoad a, BLoad b, Epi epi, int splits, float* partial,
                             cudaStream_t st) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (splits < 1) splits = 1;
  int64_t k_split = ceil_div(ceil_div(K, splits), GBK) * GBK;
  if (k_split <= 0) k_split = GBK;
  splits = (int)ceil_div(K, k_split);
  if (splits < 1) splits = 1;
  dim3 grid((unsigned)ceil_div(M, GBM), (unsigned)ceil_div(N, GBN), (unsigned)splits);
  gemm_simt_kernel<ALoad, BLoad, Epi, A_K_FAST, B_N_FAST><<<grid, GTHREADS, 0, st>>>(
      M, N, K, k_split, a, b, epi, splits > 1 ? partial : nullptr);
  count_launch();
  if (splits > 1) {
    int64_t MN = M * N;
    splitk_reduce_kernel<Epi><<<(unsigned)ceil_div(MN, 256), 256, 0, st>>>(partial, MN, N, splits, epi);
    count_launch();
  }
  return cudaGetLastError();
}

// pick a split count so that a K-heavy GEMM with few output tiles still fills the 148 SMs
inline int pick_splits(int64_t M, int64_t N, int64_t K, int64_t max_splits = 64) {
  int64_t tiles = ceil_div(M, GBM) * ceil_div(N, GBN);
  int64_t want = ceil_div(148 * 2, tiles);
  int64_t by_k = ceil_div(K, 8 * GBK);
  int64_t s = want < by_k ? want : by_k;
  if (s > max_splits) s = max_splits;
  return s < 1 ? 1 : (int)s;
}

// ---- common loaders / epilogues ---------------------------------------------------------------
struct RowMajor {           // X[r, c] with leading dimension ld ; call as (r, c)
  const float* p; int64_t ld;
  __device__ __forceinline__ float operator()(int64_t r, int64_t c) const { return __ldg(p + r * ld + c); }
};
struct Transposed {         // view (r, c) -> X[c, r]
  const float* p; int64_t ld;
  __device__ __forceinline__ float operator()(int64_t r, int64_t c) const { return __ldg(p + c * ld + r); }
};
struct StoreEpi {
  float* out; int64_t ld;
  __device__ __forceinline__ void operator()(int64_t m, int64_t n, float v) const { out[m * ld + n] = v; }
};
struct BiasActEpi {         // out = act(v + bias[n]); act 0 none, 1 relu, 2 tanh
  float* out; int64_t ld; const float* bias; int act;
  __device__ __forceinline__ void operator()(int64_t m, int64_t n, float v) const {
    if (bias) v += __ldg(bias + n);
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 2) v = tanhf(v);
    out[m * ld + n] = v;
  }
};

// column sums of a [R, C] matrix (bias gradients), two fixed-order levels, no atomics.
// `partial` needs colsum_chunks(R) * C floats.
int64_t colsum_chunks(int64_t R);
cudaError_t colsum(const float* x, float* out, int64_t R, int64_t C, float* partial, cudaStream_t st);
// out[c] = sum_r x[r*ld + c], R small (one launch)
cudaError_t colsum_small(const float* x, int64_t ld, float* out, int64_t R, int64_t C, cudaStream_t st);
cudaError_t colsum_bf16(const __nv_bfloat16* x, int64_t ld, float* out, int64_t R, int64_t C, float* partial, cudaStream_t st);

}  // namespace mr
