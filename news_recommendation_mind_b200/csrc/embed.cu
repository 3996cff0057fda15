// Token-embedding gather and the atomic-free embedding-table gradient.
//   forward : models/Embeddings/BERT.py:39            out[t,:] = W[ids[t],:]
//   backward: autograd of the above (embedding_dense_backward with padding_idx = 0)
//
// Gradient algorithm (deterministic, no floating-point atomics):
//   1. keys = token ids, values = token positions; stable LSB radix sort on the low
//      ceil(log2 V) bits (cub::DeviceRadixSort -- integer plumbing from the CUDA toolkit);
//   2. segment boundaries of equal ids via a flag + lower-bound search per vocabulary row;
//   3. level 1: one warp per chunk of <= SEG_CHUNK sorted rows sums those rows of d_emb
//      (float4 / bf16x8 coalesced row reads) -> single-chunk segments write d_table directly,
//      longer ones write a partial row;
//   4. level 2: one warp per multi-chunk segment adds its partial rows in order.
//   Rows with no occurrence and the padding row are written as zeros (dense gradient, as the
//   reference produces, so that a dense Adam step sees every row).
#include <cub/cub.cuh>
#include "common.cuh"
#include "embed.cuh"

namespace mr {

constexpr int SEG_CHUNK = 32;

__global__ void gather_rows_kernel(const void* __restrict__ ids, int is64, const float* __restrict__ table,
                                   float* __restrict__ out, int64_t T, int64_t E, int64_t V) {
  // one warp per row, float4 when E % 4 == 0
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  int64_t row = load_index(ids, is64, t);
  row = row < 0 ? 0 : (row >= V ? V - 1 : row);
  const float* src = table + row * E;
  float* dst = out + t * E;
  if ((E & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int64_t i = lane; i < E / 4; i += 32) d4[i] = __ldg(s4 + i);
  } else {
    for (int64_t i = lane; i < E; i += 32) dst[i] = __ldg(src + i);
  }
}

__global__ void make_keys_kernel(const void* __restrict__ ids, int is64, int32_t* __restrict__ keys,
                                 int32_t* __restrict__ vals, int64_t T, int64_t V) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  int64_t row = load_index(ids, is64, t);
  row = row < 0 ? 0 : (row >= V ? V - 1 : row);
  keys[t] = (int32_t)row;
  vals[t] = (int32_t)t;
}

// seg_start[v] = first sorted position whose key >= v   (v in [0, V]); binary search per row
__global__ void seg_bounds_kernel(const int32_t* __restrict__ sorted_keys, int32_t* __restrict__ seg_start, int64_t T,
                                  int64_t V) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v > V) return;
  int64_t lo = 0, hi = T;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted_keys[mid] < v) lo = mid + 1; else hi = mid;
  }
  seg_start[v] = (int32_t)lo;
}

// chunks per row (0 for empty / padding rows)
__global__ void chunk_count_kernel(const int32_t* __restrict__ seg_start, int32_t* __restrict__ nchunk, int64_t V,
                                   int64_t padding_idx) {
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v > V) return;
  if (v == V) { nchunk[v] = 0; return; }                // sentinel so the scan yields the total at [V]
  int32_t cnt = seg_start[v + 1] - seg_start[v];
  nchunk[v] = (v == padding_idx) ? 0 : (cnt + SEG_CHUNK - 1) / SEG_CHUNK;
}

// chunk -> row table: chunk ch belongs to the last row v with chunk_off[v] <= ch (binary search; a per-row loop would
// serialise the thousands of chunks of the PAD token in one thread)
__global__ void chunk_fill_kernel(const int32_t* __restrict__ chunk_off, int32_t* __restrict__ chunk_row, int64_t V,
                                  int64_t max_chunks) {
  const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= max_chunks || ch >= chunk_off[V]) return;
  int64_t lo = 0, hi = V;                      // invariant: chunk_off[lo] <= ch < chunk_off[hi]
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (chunk_off[mid] <= ch) lo = mid; else hi = mid;
  }
  chunk_row[ch] = (int32_t)lo;
}

template <class T> struct RowReader;
template <> struct RowReader<float> {
  static constexpr int VEC = 4;
  __device__ static void load(const float* row, int64_t i, float* o) {
    float4 v = __ldg(reinterpret_cast<const float4*>(row) + i);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  __device__ static float load1(const float* row, int64_t i) { return __ldg(row + i); }
};
template <> struct RowReader<__nv_bfloat16> {
  static constexpr int VEC = 8;
  __device__ static void load(const __nv_bfloat16* row, int64_t i, float* o) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(row) + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); o[2 * j] = f.x; o[2 * j + 1] = f.y; }
  }
  __device__ static float load1(const __nv_bfloat16* row, int64_t i) { return __bfloat162float(row[i]); }
};

// level 1: warp per chunk.  MAXV = vectors per lane (E <= 32 * VEC * MAXV on the vector path)
template <class T, int MAXV>
__global__ void __launch_bounds__(256)
seg_reduce_l1_kernel(const T* __restrict__ d_emb, int64_t ld, const int32_t* __restrict__ sorted_pos,
                     const int32_t* __restrict__ seg_start, const int32_t* __restrict__ chunk_off,
                     const int32_t* __restrict__ chunk_row, const int32_t* __restrict__ nchunk,
                     float* __restrict__ d_table, float* __restrict__ partial, int64_t n_chunks, int64_t E, int64_t V) {
  // partial rows have pitch Ep = E rounded up to 4 (16-byte rows: the heavy-row kernels read float4 whatever E is)
  const int64_t Ep = (E + 3) & ~(int64_t)3;
  using R = RowReader<T>;
  constexpr int VEC = R::VEC;
  const int lane = threadIdx.x & 31;
  const int64_t ch = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= n_chunks || ch >= chunk_off[V]) return;
  const int32_t v = chunk_row[ch];
  const int32_t j = (int32_t)ch - chunk_off[v];
  const int32_t beg = seg_start[v] + j * SEG_CHUNK;
  const int32_t end = min(seg_start[v + 1], beg + SEG_CHUNK);
  float* dst = (nchunk[v] == 1) ? d_table + (int64_t)v * E : partial + ch * Ep;
  // vector path: rows are read in VEC-wide pieces up to the row pitch (columns in [E, ld) are padding the
  // producer keeps at zero); only the first E sums are written back.
  const int64_t nv = (E + VEC - 1) / VEC;
  const bool vec_ok = (ld % VEC == 0) && (nv * VEC <= ld) && (nv <= 32 * MAXV);
  if (vec_ok) {
    float acc[MAXV][VEC];
#pragma unroll
    for (int a = 0; a < MAXV; ++a)
#pragma unroll
      for (int b = 0; b < VEC; ++b) acc[a][b] = 0.f;
    for (int32_t r = beg; r < end; ++r) {
      const T* row = d_emb + (int64_t)sorted_pos[r] * ld;
#pragma unroll
      for (int a = 0; a < MAXV; ++a) {
        int64_t i = lane + 32 * a;
        if (i < nv) {
          float t[VEC];
          R::load(row, i, t);
#pragma unroll
          for (int b = 0; b < VEC; ++b) acc[a][b] += t[b];
        }
      }
    }
#pragma unroll
    for (int a = 0; a < MAXV; ++a) {
      int64_t i = lane + 32 * a;
      if (i < nv) {
#pragma unroll
        for (int b = 0; b < VEC; ++b)
          if (i * VEC + b < E) dst[i * VEC + b] = acc[a][b];
      }
    }
  } else {
    for (int64_t e = lane; e < E; e += 32) {
      float s = 0.f;
      for (int32_t r = beg; r < end; ++r) s += R::load1(d_emb + (int64_t)sorted_pos[r] * ld, e);
      dst[e] = s;
    }
  }
}

// level 2: warp per vocabulary row; zero-fills empty rows, folds the partial rows of short multi-chunk
// segments; rows with more than SEG_HEAVY chunks (frequent tokens: [CLS], [SEP], the head of the Zipf
// distribution) are queued for the heavy path below instead of being summed by one warp.
constexpr int SEG_HEAVY = 8;

template <class TO> __device__ __forceinline__ TO seg_out(float v);
template <> __device__ __forceinline__ float seg_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 seg_out<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// TO = float: the dense table gradient;  TO = bf16: the grouped conv-gradient rows S (row pitch ldo)
template <class TO>
__global__ void __launch_bounds__(256)
seg_reduce_l2_kernel(const int32_t* __restrict__ chunk_off, const int32_t* __restrict__ nchunk,
                     const float* __restrict__ partial, TO* __restrict__ d_table, int64_t ldo, int64_t V, int64_t E,
                     int32_t* __restrict__ heavy_count, int32_t* __restrict__ heavy_rows, int32_t max_heavy, int32_t heavy_thr) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= V) return;
  const int32_t n = nchunk[v];
  if (n == 1) return;                                   // written by level 1
  TO* dst = d_table + v * ldo;
  if (n == 0) {
    for (int64_t e = lane; e < E; e += 32) dst[e] = seg_out<TO>(0.f);
    return;
  }
  if (n > heavy_thr) {
    if (lane == 0) {
      const int32_t slot = atomicAdd(heavy_count, 1);   // integer bookkeeping only: the order of the list does
      if (slot < max_heavy) heavy_rows[slot] = (int32_t)v;   // not influence any floating-point sum
    }
    return;
  }
  const int64_t Ep = (E + 3) & ~(int64_t)3;
  const float* src = partial + (int64_t)chunk_off[v] * Ep;
  for (int64_t e = lane; e < E; e += 32) {
    float s = 0.f;
    for (int32_t j = 0; j < n; ++j) s += src[(int64_t)j * Ep + e];
    dst[e] = seg_out<TO>(s);
  }
}

// heavy rows (more than SEG_HEAVY chunks: PAD, [CLS], [SEP], the head of the Zipf distribution).  Work item = one
// slice of <= SEG_SLICE consecutive partial rows of one heavy row; items are numbered by a prefix sum over the heavy
// list (recomputed per CTA in shared memory -- the list has at most a few hundred entries).  A CTA sums a slice with
// 2 row groups x 128 column vectors, 4 independent accumulators each, combined in a fixed order; heavy_final adds the
// slice sums of a row in slice order.  The (atomic) order of the heavy list only decides which CTA does what.
constexpr int SEG_SLICE = 64;
constexpr int SEG_MAX_HEAVY_SMEM = 2048;

__device__ __forceinline__ int heavy_prefix(const int32_t* __restrict__ heavy_rows, const int32_t* __restrict__ nchunk, int nh,
                                            int32_t* off /* smem [nh + 1] */) {
  // off[h] = first item of heavy row h; computed by one warp with a shuffle scan
  if (threadIdx.x < 32) {
    int32_t run = 0;
    for (int base = 0; base < nh; base += 32) {
      const int h = base + threadIdx.x;
      int32_t c = h < nh ? (nchunk[heavy_rows[h]] + SEG_SLICE - 1) / SEG_SLICE : 0;
      int32_t x = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if ((int)threadIdx.x >= d) x += y;
      }
      if (h < nh) off[h] = run + x - c;
      run += __shfl_sync(0xffffffffu, x, 31);
    }
    if (threadIdx.x == 0) off[nh] = run;
  }
  __syncthreads();
  return off[nh];
}

__global__ void __launch_bounds__(256)
seg_reduce_heavy_kernel(const int32_t* __restrict__ chunk_off, const int32_t* __restrict__ nchunk,
                        const float* __restrict__ partial, float* __restrict__ partial2,
                        const int32_t* __restrict__ heavy_count, const int32_t* __restrict__ heavy_rows, int64_t E,
                        int32_t max_items) {
  pdl_trigger();
  pdl_wait();
  __shared__ int32_t off[SEG_MAX_HEAVY_SMEM + 1];
  __shared__ float4 red[128];
  const int nh = min(*heavy_count, SEG_MAX_HEAVY_SMEM);
  if (nh == 0) return;
  const int n_items = min(heavy_prefix(heavy_rows, nchunk, nh, off), max_items);
  const int rg = threadIdx.x >> 7, cv = threadIdx.x & 127;
  const int64_t Ep = (E + 3) & ~(int64_t)3;              // pitch of the partial rows (columns E..Ep-1 are never read back)
  const int64_t nv = Ep >> 2;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    int lo = 0, hi = nh;                                 // off[lo] <= item < off[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (off[mid] <= item) lo = mid; else hi = mid;
    }
    const int32_t v = heavy_rows[lo];
    const int32_t n = nchunk[v];
    const int32_t j0 = (item - off[lo]) * SEG_SLICE, j1 = min(n, j0 + SEG_SLICE);
    const float* src = partial + (int64_t)chunk_off[v] * Ep;
    for (int64_t c0 = 0; c0 < nv; c0 += 128) {
      const int64_t c = c0 + cv;
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
      if (c < nv) {
        int32_t j = j0 + rg;
        for (; j + 6 < j1; j += 8) {
          const float4 x0 = *reinterpret_cast<const float4*>(src + (int64_t)j * Ep + 4 * c);
          const float4 x1 = *reinterpret_cast<const float4*>(src + (int64_t)(j + 2) * Ep + 4 * c);
          const float4 x2 = *reinterpret_cast<const float4*>(src + (int64_t)(j + 4) * Ep + 4 * c);
          const float4 x3 = *reinterpret_cast<const float4*>(src + (int64_t)(j + 6) * Ep + 4 * c);
          a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
          a1.x += x1.x; a1.y += x1.y; a1.z += x1.z; a1.w += x1.w;
          a2.x += x2.x; a2.y += x2.y; a2.z += x2.z; a2.w += x2.w;
          a3.x += x3.x; a3.y += x3.y; a3.z += x3.z; a3.w += x3.w;
        }
        for (; j < j1; j += 2) {
          const float4 x0 = *reinterpret_cast<const float4*>(src + (int64_t)j * Ep + 4 * c);
          a0.x += x0.x; a0.y += x0.y; a0.z += x0.z; a0.w += x0.w;
        }
      }
      const float4 t = make_float4((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y), (a0.z + a1.z) + (a2.z + a3.z),
                                   (a0.w + a1.w) + (a2.w + a3.w));
      if (rg == 1) red[cv] = t;
      __syncthreads();
      if (rg == 0 && c < nv) {
        const float4 u = red[cv];
        *reinterpret_cast<float4*>(partial2 + (int64_t)item * Ep + 4 * c) = make_float4(t.x + u.x, t.y + u.y, t.z + u.z, t.w + u.w);
      }
      __syncthreads();
    }
  }
}

// final: CTA (h, column block of 64) adds the slice sums of heavy row h in slice order -- 4 slice groups x 64 columns
// per CTA (4 independent chains each), combined through shared memory in a fixed order
template <class TO>
__global__ void __launch_bounds__(256)
seg_reduce_heavy_final_kernel(const float* __restrict__ partial2, TO* __restrict__ d_table, int64_t ldo,
                              const int32_t* __restrict__ nchunk, const int32_t* __restrict__ heavy_count,
                              const int32_t* __restrict__ heavy_rows, int64_t E, int32_t max_items) {
  pdl_trigger();
  pdl_wait();
  __shared__ int32_t off[SEG_MAX_HEAVY_SMEM + 1];
  __shared__ float red[4][64];
  const int nh = min(*heavy_count, SEG_MAX_HEAVY_SMEM);
  if (nh == 0) return;
  heavy_prefix(heavy_rows, nchunk, nh, off);
  const int64_t Ep = (E + 3) & ~(int64_t)3;
  const int cx = threadIdx.x & 63, sg = threadIdx.x >> 6;
  const int64_t e = (int64_t)blockIdx.y * 64 + cx;
  for (int h = blockIdx.x; h < nh; h += gridDim.x) {
    const int32_t v = heavy_rows[h];
    const int i0 = off[h], i1 = min(off[h + 1], max_items);
    // slice group sg takes a contiguous quarter of the slices
    const int per = (i1 - i0 + 3) / 4;
    const int a = min(i1, i0 + sg * per), b = min(i1, a + per);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (e < E) {
      int i = a;
      for (; i + 3 < b; i += 4) {
        a0 += partial2[(int64_t)i * Ep + e];
        a1 += partial2[(int64_t)(i + 1) * Ep + e];
        a2 += partial2[(int64_t)(i + 2) * Ep + e];
        a3 += partial2[(int64_t)(i + 3) * Ep + e];
      }
      for (; i < b; ++i) a0 += partial2[(int64_t)i * Ep + e];
    }
    red[sg][cx] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (sg == 0 && e < E) d_table[(int64_t)v * ldo + e] = seg_out<TO>((red[0][cx] + red[1][cx]) + (red[2][cx] + red[3][cx]));
    __syncthreads();
  }
}

struct EmbedGradPlan {
  int64_t T, E, V, max_chunks, max_heavy, max_items;
  int32_t heavy_thr;             // rows with more chunks than this go to the heavy path
  size_t cub_sort_bytes, cub_scan_bytes;
};

static EmbedGradPlan plan_embed_grad(int64_t T, int64_t E, int64_t V) {
  EmbedGradPlan p{T, E, V, 0, 0, 0, 0, 0, 0};
  p.max_chunks = ceil_div(T, SEG_CHUNK) + V;
  // heavy threshold: at least SEG_HEAVY chunks, raised so that at most SEG_MAX_HEAVY_SMEM rows can exceed it; the heavy
  // kernels read float4 columns, so a row length that is not a multiple of 4 keeps everything on level 2
  int64_t thr = SEG_HEAVY;
  while (T / ((int64_t)SEG_CHUNK * thr) + 1 > 2048) thr *= 2;
  p.heavy_thr = (int32_t)thr;
  p.max_heavy = T / ((int64_t)SEG_CHUNK * thr) + 1;
  p.max_items = p.max_chunks / 64 + p.max_heavy;                // slices of <= SEG_SLICE partial rows
  int bits = 1;
  while ((1ll << bits) < V) ++bits;
  cub::DeviceRadixSort::SortPairs(nullptr, p.cub_sort_bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)T, 0, bits);
  cub::DeviceScan::ExclusiveSum(nullptr, p.cub_scan_bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)V + 1);
  return p;
}

// ---- token-grouped conv gradient ("sum before multiply") -------------------------------------------------------------
// Every use of a vocabulary row v in the conv input contributes  x_v (outer) dconv[t - tap + 1]  to the filter gradient
// and  dconv[t - tap + 1] W_tap^T  to the gradient of row v.  Both are linear in dconv, so the rows are first summed per
// vocabulary row:   S[v, tap, :] = sum over tokens t with ids[t] = v of dconv[t + 1 - tap, :]   (same title only),
// and the two GEMMs then run over V rows instead of T tokens (news_cnn_tc.cu).  Same machinery as the table gradient
// above (sorted token positions, <= 32-row chunks per warp, fixed-order partial sums, no atomics); a token's "row" is
// the 3 x Hp concatenation of its right neighbour, itself and its left neighbour in dconv.
template <int MAXV>     // 16-byte vectors per lane: 3 * Hp / 8 <= 32 * MAXV
__global__ void __launch_bounds__(256)
group_taps_l1_kernel(const __nv_bfloat16* __restrict__ dconv, int64_t ld, int L, int pieces, const int32_t* __restrict__ sorted_pos,
                     const int4* __restrict__ chunk_desc, __nv_bfloat16* __restrict__ S, int64_t lds, float* __restrict__ partial,
                     int64_t n_chunks) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t ch = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ch >= n_chunks) return;
  const int4 cd = __ldg(chunk_desc + ch);
  if (cd.x < 0) return;
  const int32_t v = cd.x, beg = cd.y, end = cd.z;
  const int nvec = 3 * pieces;
  int shift[MAXV], col[MAXV];             // vector i = lane + 32 a: tap = i / pieces reads row t + 1 - tap, columns 8 (i % pieces)..
#pragma unroll
  for (int a = 0; a < MAXV; ++a) {
    const int i = lane + 32 * a;
    const int tap = i / pieces;
    shift[a] = 1 - tap;
    col[a] = (i - tap * pieces) * 8;
  }
  float acc[MAXV][8];
#pragma unroll
  for (int a = 0; a < MAXV; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  // my = this lane's sorted position (one coalesced load per chunk), broadcast per row
  const int32_t my = (beg + lane < end) ? sorted_pos[beg + lane] : 0;
#pragma unroll 4
  for (int32_t r = 0; r < end - beg; ++r) {
    const int32_t t = __shfl_sync(0xffffffffu, my, r);
    const int l = t % L;
#pragma unroll
    for (int a = 0; a < MAXV; ++a) {
      const int i = lane + 32 * a;
      const int ls = l + shift[a];
      if (i < nvec && ls >= 0 && ls < L) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(dconv + (int64_t)(t + shift[a]) * ld + col[a]));
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const float2 f = __bfloat1622float2(h[w]);
          acc[a][2 * w] += f.x;
          acc[a][2 * w + 1] += f.y;
        }
      }
    }
  }
  const bool direct = cd.w != 0;
#pragma unroll
  for (int a = 0; a < MAXV; ++a) {
    const int i = lane + 32 * a;
    if (i < nvec) {
      if (direct) {
        uint4 o;
        __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[a][0], acc[a][1]), p1 = __floats2bfloat162_rn(acc[a][2], acc[a][3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(acc[a][4], acc[a][5]), p3 = __floats2bfloat162_rn(acc[a][6], acc[a][7]);
        o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
        o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
        *reinterpret_cast<uint4*>(S + (int64_t)v * lds + i * 8) = o;
      } else {
        float4* d = reinterpret_cast<float4*>(partial + ch * (int64_t)(nvec * 8) + i * 8);
        d[0] = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
        d[1] = make_float4(acc[a][4], acc[a][5], acc[a][6], acc[a][7]);
      }
    }
  }
}

// ---- the grouping PLAN depends on the token ids only (not on any gradient): it can be built while the forward pass
// is still running (ops.py queues it on a side stream) and is consumed by the backward --------------------------
struct GroupPlanLayout {
  int64_t svals, seg_start, nchunk, chunk_off, chunk_desc, scratch, total;     // byte offsets into the plan buffer
  int64_t max_chunks;
  size_t cub_sort_bytes, cub_scan_bytes;
};
static GroupPlanLayout group_plan_layout(int64_t T, int64_t V) {
  GroupPlanLayout l{};
  EmbedGradPlan p = plan_embed_grad(T, 4, V);
  l.max_chunks = p.max_chunks;
  l.cub_sort_bytes = p.cub_sort_bytes;
  l.cub_scan_bytes = p.cub_scan_bytes;
  int64_t o = 0;
  l.svals = o; o += arena_bytes(T, 4);
  l.seg_start = o; o += arena_bytes(V + 1, 4);
  l.nchunk = o; o += arena_bytes(V + 1, 4);
  l.chunk_off = o; o += arena_bytes(V + 1, 4);
  l.chunk_desc = o; o += arena_bytes(p.max_chunks * 4, 4);
  l.scratch = o;                                     // keys, vals, sorted keys, cub temporaries (dead after the build)
  o += 3 * arena_bytes(T, 4) + arena_bytes((int64_t)p.cub_sort_bytes, 1) + arena_bytes((int64_t)p.cub_scan_bytes, 1);
  l.total = o + 256;
  return l;
}

// chunk ch -> {vocabulary row, first sorted position, end, 1 if it is the row's only chunk}: one 16-byte load in the
// reduction kernel instead of a chain of dependent lookups
__global__ void chunk_desc_kernel(const int32_t* __restrict__ chunk_off, const int32_t* __restrict__ seg_start,
                                  const int32_t* __restrict__ nchunk, int4* __restrict__ desc, int64_t V, int64_t max_chunks) {
  const int64_t ch = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= max_chunks) return;
  if (ch >= chunk_off[V]) { desc[ch] = make_int4(-1, 0, 0, 0); return; }
  int64_t lo = 0, hi = V;
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (chunk_off[mid] <= ch) lo = mid; else hi = mid;
  }
  const int32_t v = (int32_t)lo;
  const int32_t beg = seg_start[v] + ((int32_t)ch - chunk_off[v]) * SEG_CHUNK;
  desc[ch] = make_int4(v, beg, min(seg_start[v + 1], beg + SEG_CHUNK), nchunk[v] == 1 ? 1 : 0);
}

int64_t token_group_plan_bytes(int64_t T, int64_t V) {
  if (T < 0 || V < 1 || T >= (1ll << 31)) return -1;
  return group_plan_layout(T, V).total;
}

int token_group_plan_build(const void* ids, int ids_i64, int64_t T, int64_t V, void* plan, int64_t plan_bytes, cudaStream_t st) {
  const GroupPlanLayout l = group_plan_layout(T, V);
  MR_REQUIRE(plan != nullptr && plan_bytes >= l.total, MR_ERR_WORKSPACE, "token grouping plan: buffer too small (%lld given, %lld needed)",
             (long long)plan_bytes, (long long)l.total);
  uint8_t* base = static_cast<uint8_t*>(plan);
  int32_t* svals = reinterpret_cast<int32_t*>(base + l.svals);
  int32_t* seg_start = reinterpret_cast<int32_t*>(base + l.seg_start);
  int32_t* nchunk = reinterpret_cast<int32_t*>(base + l.nchunk);
  int32_t* chunk_off = reinterpret_cast<int32_t*>(base + l.chunk_off);
  int4* desc = reinterpret_cast<int4*>(base + l.chunk_desc);
  Arena ar(base + l.scratch, plan_bytes - l.scratch);
  int32_t* keys = ar.take<int32_t>(T);
  int32_t* vals = ar.take<int32_t>(T);
  int32_t* skeys = ar.take<int32_t>(T);
  void* cub_sort = ar.take<char>((int64_t)l.cub_sort_bytes);
  void* cub_scan = ar.take<char>((int64_t)l.cub_scan_bytes);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "token grouping plan: buffer too small");
  make_keys_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, st>>>(ids, ids_i64, keys, vals, T, V);
  MR_CHECK_LAUNCH("make_keys_kernel");
  int bits = 1;
  while ((1ll << bits) < V) ++bits;
  size_t sb = l.cub_sort_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_sort, sb, keys, skeys, vals, svals, (int)T, 0, bits, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "radix sort: %s", cudaGetErrorString(e));
  count_launch(2 * ((bits + 7) / 8) + 1);
  seg_bounds_kernel<<<(unsigned)ceil_div(V + 1, 256), 256, 0, st>>>(skeys, seg_start, T, V);
  MR_CHECK_LAUNCH("seg_bounds_kernel");
  chunk_count_kernel<<<(unsigned)ceil_div(V + 1, 256), 256, 0, st>>>(seg_start, nchunk, V, -1);     // every row, padding row included
  MR_CHECK_LAUNCH("chunk_count_kernel");
  size_t cb = l.cub_scan_bytes;
  e = cub::DeviceScan::ExclusiveSum(cub_scan, cb, nchunk, chunk_off, (int)V + 1, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "scan: %s", cudaGetErrorString(e));
  count_launch(2);
  chunk_desc_kernel<<<(unsigned)ceil_div(l.max_chunks, 256), 256, 0, st>>>(chunk_off, seg_start, nchunk, desc, V, l.max_chunks);
  MR_CHECK_LAUNCH("chunk_desc_kernel");
  return MR_OK;
}

int64_t token_group_workspace_bytes(int64_t T, int64_t Hp, int64_t V) {
  if (T < 0 || V < 1 || T >= (1ll << 31)) return -1;
  const int64_t E = 3 * Hp;
  EmbedGradPlan p = plan_embed_grad(T, E, V);
  int64_t b = 0;
  b += arena_bytes(group_plan_layout(T, V).total, 1);           // plan built in place when the caller brings none
  b += arena_bytes(p.max_chunks * E, 4);
  b += arena_bytes(p.max_heavy + 1, 4) + arena_bytes(p.max_items * E, 4);
  return b + 256;
}

int token_group_taps(const void* ids, int ids_i64, const void* plan_in, const __nv_bfloat16* dconv, int64_t ld, int L, int64_t T,
                     int64_t V, __nv_bfloat16* S, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
  const int64_t Hp = ld, E = 3 * Hp;
  MR_REQUIRE(Hp % 8 == 0 && 3 * Hp / 8 <= 64, MR_ERR_UNSUPPORTED, "token grouping: row pitch %lld", (long long)Hp);
  EmbedGradPlan p = plan_embed_grad(T, E, V);
  const GroupPlanLayout l = group_plan_layout(T, V);
  Arena ar(workspace, workspace_bytes);
  uint8_t* own_plan = ar.take<uint8_t>(l.total);
  float* partial = ar.take<float>(p.max_chunks * E);
  int32_t* heavy = ar.take<int32_t>(p.max_heavy + 1);
  float* partial2 = ar.take<float>(p.max_items * E);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "token grouping: workspace too small (%lld given)", (long long)workspace_bytes);
  const uint8_t* base = static_cast<const uint8_t*>(plan_in);
  if (base == nullptr) {
    if (int rc = token_group_plan_build(ids, ids_i64, T, V, own_plan, l.total, st)) return rc;
    base = own_plan;
  }
  const int32_t* svals = reinterpret_cast<const int32_t*>(base + l.svals);
  const int32_t* nchunk = reinterpret_cast<const int32_t*>(base + l.nchunk);
  const int32_t* chunk_off = reinterpret_cast<const int32_t*>(base + l.chunk_off);
  const int4* desc = reinterpret_cast<const int4*>(base + l.chunk_desc);
  const int pieces = (int)(Hp / 8);
  if (3 * pieces <= 32)
    launch_pdl(group_taps_l1_kernel<1>, dim3((unsigned)ceil_div(p.max_chunks, 8)), dim3(256), 0, st, dconv, ld, L, pieces, svals, desc, S, E, partial, p.max_chunks);
  else
    launch_pdl(group_taps_l1_kernel<2>, dim3((unsigned)ceil_div(p.max_chunks, 8)), dim3(256), 0, st, dconv, ld, L, pieces, svals, desc, S, E, partial, p.max_chunks);
  MR_CHECK_LAUNCH("group_taps_l1_kernel");
  cudaMemsetAsync(heavy, 0, sizeof(int32_t), st);
  launch_pdl(seg_reduce_l2_kernel<__nv_bfloat16>, dim3((unsigned)ceil_div(V, 8)), dim3(256), 0, st, chunk_off, nchunk, partial, S, E, V, E, heavy, heavy + 1,
                                                                                (int32_t)p.max_heavy, p.heavy_thr);
  MR_CHECK_LAUNCH("seg_reduce_l2_kernel");
  launch_pdl(seg_reduce_heavy_kernel, dim3(148 * 4), dim3(256), 0, st, chunk_off, nchunk, partial, partial2, heavy, heavy + 1, E, (int32_t)p.max_items);
  MR_CHECK_LAUNCH("seg_reduce_heavy_kernel");
  launch_pdl(seg_reduce_heavy_final_kernel<__nv_bfloat16>, dim3(64, (unsigned)ceil_div(E, 64)), dim3(256), 0, st, partial2, S, E, nchunk, heavy, heavy + 1, E, (int32_t)p.max_items);
  MR_CHECK_LAUNCH("seg_reduce_heavy_final_kernel");
  return MR_OK;
}

}  // namespace mr

extern "C" {

int mr_embed_gather_f32(const void* ids, int ids_i64, const float* table, float* out, int64_t T, int64_t E, int64_t V,
                        void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(ids && table && out, MR_ERR_NULL, "mr_embed_gather_f32: null pointer");
  MR_REQUIRE(T >= 0 && E >= 1 && V >= 1, MR_ERR_BAD_SHAPE, "mr_embed_gather_f32: T=%lld E=%lld V=%lld", (long long)T,
             (long long)E, (long long)V);
  if (T == 0) return MR_OK;
  gather_rows_kernel<<<(unsigned)ceil_div(T, 8), 256, 0, as_stream(stream)>>>(ids, ids_i64, table, out, T, E, V);
  MR_CHECK_LAUNCH("gather_rows_kernel");
  return MR_OK;
}

int64_t mr_embed_grad_workspace_bytes(int64_t T, int64_t E, int64_t V) {
  using namespace mr;
  if (T < 0 || E < 1 || V < 1 || T >= (1ll << 31)) return -1;
  EmbedGradPlan p = plan_embed_grad(T, E, V);
  int64_t b = 0;
  b += 4 * arena_bytes(T, 4);                 // keys, vals, sorted keys, sorted vals
  b += arena_bytes(V + 1, 4);                 // seg_start
  b += 2 * arena_bytes(V + 1, 4);             // nchunk, chunk_off
  b += arena_bytes(p.max_chunks, 4);          // chunk_row
  b += arena_bytes(p.max_chunks * align_up(E, 4), 4);      // partial rows (16-byte pitch)
  b += arena_bytes(p.max_heavy + 1, 4) + arena_bytes(p.max_items * align_up(E, 4), 4);   // heavy list, heavy partials
  b += arena_bytes((int64_t)p.cub_sort_bytes, 1) + arena_bytes((int64_t)p.cub_scan_bytes, 1);
  return b + 256;
}

int mr_embed_grad_segreduce(const void* ids, int ids_i64, const void* d_emb, int d_emb_dtype, int64_t d_emb_ld,
                            float* d_table, int64_t T, int64_t E, int64_t V, int64_t padding_idx, void* workspace,
                            int64_t workspace_bytes, void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(ids && d_emb && d_table, MR_ERR_NULL, "mr_embed_grad_segreduce: null pointer");
  MR_REQUIRE(T >= 0 && E >= 1 && V >= 1 && T < (1ll << 31), MR_ERR_BAD_SHAPE, "mr_embed_grad_segreduce: T=%lld E=%lld V=%lld",
             (long long)T, (long long)E, (long long)V);
  MR_REQUIRE(d_emb_dtype == MR_F32 || d_emb_dtype == MR_BF16, MR_ERR_UNSUPPORTED, "mr_embed_grad_segreduce: dtype %d", d_emb_dtype);
  cudaStream_t st = as_stream(stream);
  if (T == 0) {
    cudaMemsetAsync(d_table, 0, sizeof(float) * V * E, st);
    return MR_OK;
  }
  EmbedGradPlan p = plan_embed_grad(T, E, V);
  Arena ar(workspace, workspace_bytes);
  int32_t* keys = ar.take<int32_t>(T);
  int32_t* vals = ar.take<int32_t>(T);
  int32_t* skeys = ar.take<int32_t>(T);
  int32_t* svals = ar.take<int32_t>(T);
  int32_t* seg_start = ar.take<int32_t>(V + 1);
  int32_t* nchunk = ar.take<int32_t>(V + 1);
  int32_t* chunk_off = ar.take<int32_t>(V + 1);
  int32_t* chunk_row = ar.take<int32_t>(p.max_chunks);
  float* partial = ar.take<float>(p.max_chunks * align_up(E, 4));
  int32_t* heavy = ar.take<int32_t>(p.max_heavy + 1);            // [0] = count, [1..] = rows
  float* partial2 = ar.take<float>(p.max_items * align_up(E, 4));
  void* cub_sort = ar.take<char>((int64_t)p.cub_sort_bytes);
  void* cub_scan = ar.take<char>((int64_t)p.cub_scan_bytes);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "mr_embed_grad_segreduce: workspace too small (%lld given)", (long long)workspace_bytes);

  make_keys_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, st>>>(ids, ids_i64, keys, vals, T, V);
  MR_CHECK_LAUNCH("make_keys_kernel");
  int bits = 1;
  while ((1ll << bits) < V) ++bits;
  size_t sb = p.cub_sort_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_sort, sb, keys, skeys, vals, svals, (int)T, 0, bits, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "radix sort: %s", cudaGetErrorString(e));
  count_launch(2 * ((bits + 7) / 8) + 1);
  seg_bounds_kernel<<<(unsigned)ceil_div(V + 1, 256), 256, 0, st>>>(skeys, seg_start, T, V);
  MR_CHECK_LAUNCH("seg_bounds_kernel");
  chunk_count_kernel<<<(unsigned)ceil_div(V + 1, 256), 256, 0, st>>>(seg_start, nchunk, V, padding_idx);
  MR_CHECK_LAUNCH("chunk_count_kernel");
  size_t cb = p.cub_scan_bytes;
  e = cub::DeviceScan::ExclusiveSum(cub_scan, cb, nchunk, chunk_off, (int)V + 1, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "scan: %s", cudaGetErrorString(e));
  count_launch(2);
  chunk_fill_kernel<<<(unsigned)ceil_div(p.max_chunks, 256), 256, 0, st>>>(chunk_off, chunk_row, V, p.max_chunks);
  MR_CHECK_LAUNCH("chunk_fill_kernel");
  // The true chunk count (chunk_off[V]) is only known on the device; launch level 1 for the bound
  // ceil(T/SEG_CHUNK)+V and let surplus warps exit (no host sync on this path).
  const int64_t E4 = d_emb_ld > 0 ? d_emb_ld : E;
  MR_REQUIRE(E4 >= E, MR_ERR_BAD_SHAPE, "mr_embed_grad_segreduce: row pitch %lld < E=%lld", (long long)E4, (long long)E);
  if (d_emb_dtype == MR_F32 && E > 384)          // up to 512 fp32 columns on the vector path (MHA projection gradients: 450)
    seg_reduce_l1_kernel<float, 4><<<(unsigned)ceil_div(p.max_chunks, 8), 256, 0, st>>>(
        static_cast<const float*>(d_emb), E4, svals, seg_start, chunk_off, chunk_row, nchunk, d_table, partial, p.max_chunks, E, V);
  else if (d_emb_dtype == MR_F32)
    seg_reduce_l1_kernel<float, 3><<<(unsigned)ceil_div(p.max_chunks, 8), 256, 0, st>>>(
        static_cast<const float*>(d_emb), E4, svals, seg_start, chunk_off, chunk_row, nchunk, d_table, partial, p.max_chunks, E, V);
  else
    seg_reduce_l1_kernel<__nv_bfloat16, 3><<<(unsigned)ceil_div(p.max_chunks, 8), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(d_emb), E4, svals, seg_start, chunk_off, chunk_row, nchunk, d_table, partial, p.max_chunks, E, V);
  MR_CHECK_LAUNCH("seg_reduce_l1_kernel");
  cudaMemsetAsync(heavy, 0, sizeof(int32_t), st);
  launch_pdl(seg_reduce_l2_kernel<float>, dim3((unsigned)ceil_div(V, 8)), dim3(256), 0, st, chunk_off, nchunk, partial, d_table, E, V, E, heavy, heavy + 1,
                                                                        (int32_t)p.max_heavy, p.heavy_thr);
  MR_CHECK_LAUNCH("seg_reduce_l2_kernel");
  launch_pdl(seg_reduce_heavy_kernel, dim3(148 * 4), dim3(256), 0, st, chunk_off, nchunk, partial, partial2, heavy, heavy + 1, E, (int32_t)p.max_items);
  MR_CHECK_LAUNCH("seg_reduce_heavy_kernel");
  launch_pdl(seg_reduce_heavy_final_kernel<float>, dim3(64, (unsigned)ceil_div(E, 64)), dim3(256), 0, st, partial2, d_table, E, nchunk, heavy, heavy + 1, E, (int32_t)p.max_items);
  MR_CHECK_LAUNCH("seg_reduce_heavy_final_kernel");
  return MR_OK;
}

}  // extern "C"
