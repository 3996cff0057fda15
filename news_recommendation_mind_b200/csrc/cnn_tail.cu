// Fused tail of the CNN news encoder on tcgen05 (cnn_tail.cuh; models/Encoders/CNN.py:44-48, Attention.py:5-30,56-80).
//
// Round 1 ran the backward of (projection + tanh + additive-attention pooling) as three kernels that each streamed
// [T, Hp] bf16 matrices through HBM: pooling backward (read c, key; write dkp, mask), the projection-weight gradient
// (token-reduction GEMM: read c, dkp) and the RELUGRAD_POOL tap GEMM (read dkp, mask; write dconv) -- 1.0 GB of traffic
// and 315 us for work whose operands fit a tile.  Here one persistent CTA per SM walks tiles of G titles (<= 128 token
// rows, position-major r = l*G + g like every other tile in this library):
//
//   TMA warp    c tile and key tile (each nblk blocks of [128 rows x 64 B], SWIZZLE_64B) into a 2-stage ring; Wq resident.
//   warps 0-3   "phase A" (one title per warp at a time): dp = <d_news, c[l]>, softmax backward -> ds, then
//               dkp = ds q (1 - key^2) written IN PLACE over the key tile (bf16), the relu'(c) sign bits to a side buffer,
//               d_query / d_proj_b partial sums in registers.
//   MMA warp    D2 += dkp^T c   (both operands MN-major views of the staged tiles, K = the tile's token rows; accumulated in
//               TMEM over all tiles of the CTA: the projection-weight gradient), then D1 = dkp Wq (dkp as the K-major A
//               operand, Wq from its resident panel image).
//   warps 4-7   epilogue: dconv = relu'(c) * (p d_news + D1) -> bf16 -> global, column sums for the conv bias.
// HBM traffic per step: read c, key (288 MB), write dconv (144 MB).  Rounding points are those of the three-kernel path
// (dkp and dconv bf16, everything else fp32), so the emulation in tests/test_gpu_tc.py holds for both.
#include "cnn_tail.cuh"
#include "gemm_simt.cuh"
#include "ktiming.cuh"
#include "pool_kernels.cuh"
#include "tapgemm.cuh"
#include "tma.cuh"
#include "tokred.cuh"

namespace mr {

constexpr int CT_THREADS = 576;     // warps 0-7 phase A (two groups), 8-15 epilogue, 16 MMA issuer, 17 TMA / weights
constexpr uint32_t CT_BLK = 8192;   // one 32-column block of a tile: 128 rows x 64 B, SWIZZLE_64B

struct CnnTailBwdArgs {
  int64_t n_titles, n_tiles;
  int L, G, H, Hp, nblk, n_mt, n_side;      // n_side: stages of the side buffers (sign mask, d_news rows): 3 when they fit
  const float* prob;
  const float* d_news;
  const float* query;
  const uint8_t* wq_img;
  const uint8_t* cmask;    // [T][32]: sign bits of c, written by the forward
  __nv_bfloat16* dconv;
  float* part_qb;      // [grid][2][Hp]
  float* part_w;       // [n_mt][grid][128][Hp]
  float* csum;         // [grid * 4][Hp]
  uint32_t smem_bytes;
  long long* dbg;      // optional [grid][4 roles][5] cycle counters (mr_debug_tapgemm_counters)
};

#define CT_TIMED(slot, stmt)                                   \
  do {                                                         \
    if (p.dbg != nullptr) {                                    \
      const long long t0__ = clock64();                        \
      stmt;                                                    \
      dbg_acc[slot] += clock64() - t0__;                       \
    } else {                                                   \
      stmt;                                                    \
    }                                                          \
  } while (0)

// byte offset of the 16-byte piece `piece` (columns 8 piece .. 8 piece + 7) of tile row r
__device__ __forceinline__ uint32_t ct_off(int r, int piece) {
  return (uint32_t)(piece >> 2) * CT_BLK + (uint32_t)r * 64u + ((((uint32_t)piece & 3u) ^ (((uint32_t)r >> 1) & 3u)) << 4);
}

// 18 warps: 96 registers per thread (the register file is allocated in units of 512 per warp)
__global__ void __launch_bounds__(CT_THREADS, 1)
cnn_tail_bwd_kernel(const CnnTailBwdArgs p, const __grid_constant__ CUtensorMap cmap, const __grid_constant__ CUtensorMap kmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  long long dbg_acc[4] = {0, 0, 0, 0};
  const long long dbg_t0 = clock64();
  const uint32_t tile_bytes = (uint32_t)p.nblk * CT_BLK;
  const uint32_t w_bytes = (uint32_t)(p.Hp / 8) * (uint32_t)p.Hp * 16u;
  const int dnp = p.Hp + 4;                               // +16 B per d_news row: the G rows read by one LDS.128 fall into different banks
  uint8_t* sC = smem;
  uint8_t* sK = sC + 2 * tile_bytes;
  uint8_t* sW = sK + 2 * tile_bytes;
  float* sDn = reinterpret_cast<float*>(sW + w_bytes);
  float* sDs = sDn + p.n_side * p.G * dnp;                  // [2 phase-A groups][128]: softmax-backward factor of every tile row
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDs + 2 * 128);
  uint64_t* full = bars;               // [2] TMA: c + key tile landed
  uint64_t* slot_free = bars + 2;      // [2] MMAs that read the stage have completed
  uint64_t* dkp_ready = bars + 4;      // [2] phase A wrote dkp / mask / d_news rows of the stage (128 arrivals)
  uint64_t* e_done = bars + 6;         // [n_side] epilogue is done with the side stage's mask / d_news rows (256 arrivals)
  uint64_t* d1_full = bars + 9;
  uint64_t* d1_empty = bars + 10;      // 256 arrivals
  uint64_t* d2_done = bars + 11;
  uint64_t* w_ready = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = p.G, L = p.L, Hp = p.Hp, H = p.H;
  const int pieces = Hp >> 3;

  pdl_trigger();
  for (uint32_t i = tid * 16; i < 4 * tile_bytes; i += CT_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&slot_free[i], 1);
      tc::mbar_init(&dkp_ready[i], 128);
    }
    for (int i = 0; i < 3; ++i) tc::mbar_init(&e_done[i], 256);
    tc::mbar_init(d1_full, 1);
    tc::mbar_init(d1_empty, 256);
    tc::mbar_init(d2_done, 1);
    tc::mbar_init(w_ready, 1);
    tc::fence_barrier_init();
  }
  if (warp == 16) tc::tmem_alloc(tmem_slot, 512);
  pdl_wait();                          // everything above touched this CTA's shared memory / TMEM only
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // =================================== phase A: pooling backward of the tile ===============================
    // two groups of four warps: group a owns stage a, i.e. every other tile of the CTA (two tiles in flight: the per-title
    // chains below are latency bound, a second warp per scheduler fills the issue slots)
    const uint32_t a = (uint32_t)warp >> 2;
    const int wq = warp & 3;
    const float inv = rsqrtf((float)H);
    const float* prob = pdl_acquire(p.prob);
    const float* d_news = pdl_acquire(p.d_news);
    // second pass (dkp over key): the group's 128 lanes own (row group, piece) units -- lane gl works on the 16-byte piece
    // gl % pieces of the rows rg, rg + RG, ...  (RG = 128 / pieces row groups; 120 of 128 lanes busy for 20 pieces, where
    // one title per warp with one piece per lane kept only 20 of 32)
    const int gl = wq * 32 + lane, RG = 128 / pieces;
    const bool unit_on = gl < RG * pieces;
    const int upiece = gl % pieces, urg = gl / pieces;
    float dq[8], db[8], qv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int h = upiece * 8 + e;
      dq[e] = 0.f; db[e] = 0.f;
      qv[e] = (unit_on && h < H) ? __ldg(p.query + h) : 0.f;
    }
    float* ds_s = sDs + a * 128;
    const int rr = lane >> 3, part = lane & 7;
    const uint8_t* cT = sC + (size_t)a * tile_bytes;
    uint8_t* kT = sK + (size_t)a * tile_bytes;
    uint32_t u = 0;
    for (int64_t tile = blockIdx.x + (int64_t)a * gridDim.x; tile < p.n_tiles; tile += 2 * (int64_t)gridDim.x, ++u) {
      const uint32_t ph = u & 1u, i = 2u * u + a;
      const uint32_t ss = i % (uint32_t)p.n_side, su = i / (uint32_t)p.n_side;          // side stage and its use count
      float* dn_s = sDn + (size_t)ss * G * dnp;
      CT_TIMED(0, tc::mbar_wait(&e_done[ss], (su & 1u) ^ 1u));   // the epilogue of the side stage's previous tile no longer reads it
      for (int g = wq; g < G; g += 4) {
        const int64_t n = tile * G + g;
        for (int h = lane; h < dnp; h += 32) dn_s[g * dnp + h] = (n < p.n_titles && h < H) ? __ldg(d_news + n * H + h) : 0.f;
      }
      __syncwarp();
      CT_TIMED(1, tc::mbar_wait(&full[a], ph));
      asm volatile("bar.sync %0, 128;" ::"r"(3 + (int)a) : "memory");         // nobody of the group still reads the previous tile's ds_s
      const long long tA0 = clock64();
      for (int g = wq; g < G; g += 4) {
        const int64_t n = tile * G + g;
        const float pl = (n < p.n_titles && lane < L) ? __ldg(prob + n * L + lane) : 0.f;
        float dr[3][8];
#pragma unroll
        for (int v = 0; v < 3; ++v)
#pragma unroll
          for (int e = 0; e < 8; ++e) dr[v][e] = (part + 8 * v < pieces) ? dn_s[g * dnp + (part + 8 * v) * 8 + e] : 0.f;
        // dp[l] = <d_news, c[l]>: 8 lanes per row, 4 rows at a time; the same pass writes the sign bits of c
        // (c = relu(.) >= +0, so "c > 0" is "the bf16 bits are not zero")
        float mine = 0.f;
        // the three 16-byte loads of an iteration are issued before anything is stored, so that their latencies overlap
        // (the compiler cannot move a shared-memory load above a store it cannot disambiguate)
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
          const int l = it * 4 + rr, r = l * G + g;
          uint4 raw[3];
#pragma unroll
          for (int v = 0; v < 3; ++v)
            raw[v] = (l < L && part + 8 * v < pieces) ? *reinterpret_cast<const uint4*>(cT + ct_off(r, part + 8 * v)) : make_uint4(0, 0, 0, 0);
          float d = 0.f;
#pragma unroll
          for (int v = 0; v < 3; ++v) {
            float f[8];
            bf8_to_f(raw[v], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) d = fmaf(dr[v][e], f[e], d);
          }
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          const float sv = __shfl_sync(0xffffffffu, d, (lane & 3) * 8);
          if ((lane >> 2) == it) mine = sv;
        }
        const float dot = warp_sum(pl * mine);
        if (p.dbg != nullptr) dbg_acc[2] += clock64() - tA0;
        const float ds = pl * (mine - dot) * inv;                              // softmax backward (Attention.py:77-80) and 1/sqrt(H)
        ds_s[lane * G + g] = ds;                                               // row r = l * G + g (lanes past L hold 0)
      }
      asm volatile("bar.sync %0, 128;" ::"r"(3 + (int)a) : "memory");         // the four warps of this group: every row's factor is there
      // dkp over key, same bytes; four rows per pass, loads first (see above)
      if (unit_on) {
#pragma unroll 1
        for (int r0 = urg; r0 < 128; r0 += 4 * RG) {
          uint4 kraw[4];
          float dsr[4];
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int r = r0 + w * RG;
            kraw[w] = r < 128 ? *reinterpret_cast<const uint4*>(kT + ct_off(r, upiece)) : make_uint4(0, 0, 0, 0);
            dsr[w] = r < 128 ? ds_s[r] : 0.f;
          }
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int r = r0 + w * RG;
            float k[8], o1[8];
            bf8_to_f(kraw[w], k);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              dq[e] = fmaf(dsr[w], k[e], dq[e]);
              o1[e] = dsr[w] * (qv[e] - qv[e] * k[e] * k[e]);
              db[e] += o1[e];
            }
            if (r < 128) *reinterpret_cast<uint4*>(kT + ct_off(r, upiece)) = f_to_bf8(o1);
          }
        }
      }
      tc::fence_proxy_async();                                                 // generic-proxy writes of dkp before the tensor core reads them
      tc::mbar_arrive(&dkp_ready[a]);
      if (p.dbg != nullptr) dbg_acc[3] += clock64() - tA0;
    }
    // d_query / d_proj_b partials of this CTA: combined below (after every role is done with shared memory)
    __syncthreads();
    float* red = reinterpret_cast<float*>(sC);                                 // [256 lanes][16]: (dq | db) of each lane's piece
    {
      float* mine = red + (size_t)(a * 128 + gl) * 16;
#pragma unroll
      for (int e = 0; e < 8; ++e) { mine[e] = dq[e]; mine[8 + e] = db[e]; }
    }
    __syncthreads();
    for (int c = tid; c < 2 * Hp; c += 256) {
      const int which = c >= Hp, col = c - which * Hp, pc = col >> 3, e = col & 7;
      float v = 0.f;
      for (int grp = 0; grp < 2; ++grp)
        for (int rg = 0; rg < RG; ++rg) v += red[(size_t)(grp * 128 + rg * pieces + pc) * 16 + which * 8 + e];
      p.part_qb[(size_t)blockIdx.x * 2 * Hp + c] = v;
    }
  } else if (warp < 16) {
    // =================================== epilogue: dconv rows =============================================
    // eight warps: warp (q4, half) reads TMEM lanes 32 q4 .. 32 q4 + 31 (one token row per thread) and the column chunks of its half
    const int q4 = warp & 3, half = (warp >> 2) & 1, r = q4 * 32 + lane;
    const int g = r % G, l = r / G;
    const bool row_ok = l < L;
    const int nc = (Hp + 31) >> 5;
    const int c_begin = half ? (nc + 1) / 2 : 0, c_end = half ? nc : (nc + 1) / 2;
    const float* prob = pdl_acquire(p.prob);
    const uint8_t* cmask_g = pdl_acquire(p.cmask);
    float colacc[3] = {0.f, 0.f, 0.f};
    const uint32_t tb = tmem + ((uint32_t)(q4 * 32) << 16);
    uint32_t i = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
      const uint32_t s = i & 1u, ph = (i >> 1) & 1u;
      const int64_t n = tile * G + g, t = n * L + l;
      const bool valid = row_ok && n < p.n_titles;
      const float pt = valid ? __ldg(prob + t) : 0.f;
      uint32_t cm[3];
      {
        const uint32_t* mrow = reinterpret_cast<const uint32_t*>(cmask_g + t * 32);     // requested before the waits below
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) cm[cj] = (valid && c_begin + cj < c_end) ? __ldg(mrow + c_begin + cj) : 0u;
      }
      CT_TIMED(0, tc::mbar_wait(&dkp_ready[s], ph));
      const uint32_t ss = i % (uint32_t)p.n_side;
      const float4* dn_row = reinterpret_cast<const float4*>(sDn + (size_t)ss * G * dnp + g * dnp);
      CT_TIMED(1, tc::mbar_wait(d1_full, i & 1u));
      tc::tc_fence_after();
      const long long tE0 = clock64();
      // NOT unrolled (instruction-cache footprint, see tapgemm.cu); per-chunk registers are picked with selects
#pragma unroll 1
      for (int cj = 0; cj < 3; ++cj) {
        const int ci = c_begin + cj;
        if (ci >= c_end) break;
        const int n0 = ci * 32;
        const bool wide = Hp - n0 >= 32;
        const uint32_t cmask_c = cj == 0 ? cm[0] : (cj == 1 ? cm[1] : cm[2]);
        uint32_t v[32];
        if (wide) {
          tc::tmem_ld32(tb + n0, v);
        } else {
          uint32_t h16[16];
          tc::tmem_ld16(tb + n0, h16);
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = h16[j]; v[16 + j] = 0u; }
        }
        float4 dn[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) dn[w] = (w < 4 || wide) ? dn_row[ci * 8 + w] : make_float4(0.f, 0.f, 0.f, 0.f);
        tc::tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          const float4 d4 = dn[w];
          const int j = 4 * w;
          f[j + 0] = ((cmask_c >> (j + 0)) & 1u) ? fmaf(pt, d4.x, __uint_as_float(v[j + 0])) : 0.f;
          f[j + 1] = ((cmask_c >> (j + 1)) & 1u) ? fmaf(pt, d4.y, __uint_as_float(v[j + 1])) : 0.f;
          f[j + 2] = ((cmask_c >> (j + 2)) & 1u) ? fmaf(pt, d4.z, __uint_as_float(v[j + 2])) : 0.f;
          f[j + 3] = ((cmask_c >> (j + 3)) & 1u) ? fmaf(pt, d4.w, __uint_as_float(v[j + 3])) : 0.f;
        }
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(p.dconv + t * Hp + n0);
#pragma unroll
          for (int w = 0; w < 4; w += 2) {
            if (w < 2 || wide) {
              uint4 o[2];
#pragma unroll
              for (int x = 0; x < 2; ++x) {
                o[x].x = tc::pack_bf16(f[(w + x) * 8 + 0], f[(w + x) * 8 + 1]);
                o[x].y = tc::pack_bf16(f[(w + x) * 8 + 2], f[(w + x) * 8 + 3]);
                o[x].z = tc::pack_bf16(f[(w + x) * 8 + 4], f[(w + x) * 8 + 5]);
                o[x].w = tc::pack_bf16(f[(w + x) * 8 + 6], f[(w + x) * 8 + 7]);
              }
              tc::st_global_256(dst + w, o[0], o[1]);
            }
          }
        }
        const float csum = tc::warp_colsum32(f, lane);
#pragma unroll
        for (int ck = 0; ck < 3; ++ck)
          if (ck == cj) colacc[ck] += csum;
      }
      tc::tc_fence_before();
      tc::mbar_arrive(d1_empty);
      tc::mbar_arrive(&e_done[ss]);
      if (p.dbg != nullptr) dbg_acc[2] += clock64() - tE0;
    }
    {
      float* cs = p.csum + ((size_t)blockIdx.x * 4 + q4) * Hp;
#pragma unroll
      for (int cj = 0; cj < 3; ++cj)
        if (c_begin + cj < c_end && (c_begin + cj) * 32 + lane < Hp) cs[(c_begin + cj) * 32 + lane] = colacc[cj];
    }
    // projection-weight gradient of this CTA's tiles: TMEM -> fp32 partial (each half its share of the columns)
    tc::mbar_wait(d2_done, 0);
    tc::tc_fence_after();
    const int n16 = Hp >> 4;
    const int h_begin = half ? n16 / 2 : 0, h_end = half ? n16 : n16 / 2;
    for (int mt = 0; mt < p.n_mt; ++mt) {
      float* dst = p.part_w + (((size_t)mt * gridDim.x + blockIdx.x) * 128 + r) * Hp;
      for (int hc = h_begin; hc < h_end; ++hc) {
        const int c0 = hc * 16;
        uint32_t v[16];
        tc::tmem_ld16(tb + (uint32_t)((1 + mt) * Hp + c0), v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int w = 0; w < 4; ++w)
          *reinterpret_cast<uint4*>(dst + c0 + 4 * w) = make_uint4(v[4 * w], v[4 * w + 1], v[4 * w + 2], v[4 * w + 3]);
      }
    }
    tc::tc_fence_before();
    __syncthreads();
    __syncthreads();
  } else if (warp == 16) {
    // =================================== MMA issuer =======================================================
    const uint32_t idesc1 = tc::make_idesc(128, Hp, 0, 0), idesc2 = tc::make_idesc(128, Hp, 1, 1);
    const uint32_t b_ps = (uint32_t)Hp * 16u;
    const uint32_t wbase = tc::smem_u32(sW);
    const int nks1 = Hp >> 4;
    tc::mbar_wait(w_ready, 0);
    uint32_t i = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
      const uint32_t s = i & 1u, ph = (i >> 1) & 1u;
      CT_TIMED(0, tc::mbar_wait(&dkp_ready[s], ph));
      tc::tc_fence_after();
      const uint32_t kbase = tc::smem_u32(sK) + s * tile_bytes, cbase = tc::smem_u32(sC) + s * tile_bytes;
      if (tc::elect_one()) {
        // D2[mt] (+)= dkp[:, 128 mt ..]^T c : token rows are the K index, 16 per MMA
        for (int mt = 0; mt < p.n_mt; ++mt) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t da = tc::make_desc_sw(kbase + (uint32_t)mt * 4u * CT_BLK + (uint32_t)ks * 1024u, CT_BLK, 512, 4, 0);
            const uint64_t dbb = tc::make_desc_sw(cbase + (uint32_t)ks * 1024u, CT_BLK, 512, 4, 0);
            tc::umma(tmem + (uint32_t)((1 + mt) * Hp), da, dbb, idesc2, (i | (uint32_t)ks) ? 1u : 0u);
          }
        }
      }
      __syncwarp();
      CT_TIMED(1, tc::mbar_wait(d1_empty, (i & 1u) ^ 1u)); // the epilogue has drained D1 of the previous tile
      tc::tc_fence_after();
      if (tc::elect_one()) {
        for (int ks = 0; ks < nks1; ++ks) {
          const uint64_t da = tc::make_desc_sw(kbase + (uint32_t)(ks >> 1) * CT_BLK + (uint32_t)(ks & 1) * 32u, 16, 512, 4, 0);
          const uint64_t dbb = tc::make_desc(wbase + 2u * (uint32_t)ks * b_ps, b_ps, 128);
          tc::umma(tmem, da, dbb, idesc1, ks > 0 ? 1u : 0u);
        }
        tc::umma_commit(&slot_free[s]);
        tc::umma_commit(d1_full);
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::umma_commit(d2_done);
    __syncwarp();
    __syncthreads();
    __syncthreads();
  } else {
    // =================================== TMA producer =====================================================
    if (lane == 0) {
      tc::tma_prefetch_desc(&cmap);
      tc::tma_prefetch_desc(&kmap);
      const uint8_t* wsrc = pdl_acquire(p.wq_img);
      tc::mbar_arrive_expect_tx(w_ready, w_bytes);
      for (uint32_t off = 0; off < w_bytes; off += 32768) {
        const uint32_t nb = w_bytes - off < 32768 ? w_bytes - off : 32768;
        tc::bulk_g2s(tc::smem_u32(sW) + off, wsrc + off, nb, w_ready);
      }
      const uint32_t stage_tx = 2u * (uint32_t)p.nblk * (uint32_t)(G * L) * 64u;
      uint32_t i = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        const uint32_t s = i & 1u, ph = (i >> 1) & 1u;
        CT_TIMED(0, tc::mbar_wait(&slot_free[s], ph ^ 1u));
        tc::mbar_arrive_expect_tx(&full[s], stage_tx);
        const uint32_t cdst = tc::smem_u32(sC) + s * tile_bytes, kdst = tc::smem_u32(sK) + s * tile_bytes;
        for (int b = 0; b < p.nblk; ++b) {
          tc::tma_load_3d(cdst + (uint32_t)b * CT_BLK, &cmap, b * 32, (int)(tile * G), 0, &full[s]);
          tc::tma_load_3d(kdst + (uint32_t)b * CT_BLK, &kmap, b * 32, (int)(tile * G), 0, &full[s]);
        }
      }
    }
    __syncthreads();
    __syncthreads();
  }
  if (p.dbg != nullptr && lane == 0 && (warp == 0 || warp == 8 || warp == 16 || warp == 17)) {
    // roles: 0 phase A {e_done wait, full wait, row dots, whole tile}, 1 epilogue {dkp_ready wait, d1_full wait, body},
    //        2 MMA {dkp_ready wait, d1_empty wait}, 3 TMA {slot_free wait}
    long long* d = p.dbg + ((size_t)blockIdx.x * 4 + (warp == 0 ? 0 : warp == 8 ? 1 : warp == 16 ? 2 : 3)) * 5;
    d[0] = dbg_acc[0]; d[1] = dbg_acc[1]; d[2] = dbg_acc[2]; d[3] = dbg_acc[3]; d[4] = clock64() - dbg_t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 16) tc::tmem_dealloc(tmem, 512);
}

// ---- host ------------------------------------------------------------------------------------------------------------
static bool use_tail() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MINDREC_CNN_TAIL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

bool cnn_tail_supported(int64_t L, int64_t Hp) { return use_tail() && L >= 16 && L <= 32 && Hp >= 16 && Hp <= 160 && Hp % 16 == 0; }

// titles per tile: a power of two (4 for 17..32 tokens, 8 for 16), so that a title's rows sit at the same positions modulo 4 whatever
// its slot g -- the tensor-core sums over token rows then see the same addends in the same order (batch-invariant results)
static inline int tail_G(int64_t L) { return L <= 16 ? 8 : 4; }

static int tail_grid(int64_t n_titles, int64_t L) {
  const int64_t G = tail_G(L), n_tiles = ceil_div(n_titles, G);
  return (int)(n_tiles < sm_count() ? (n_tiles < 1 ? 1 : n_tiles) : sm_count());
}

int64_t cnn_tail_bwd_workspace_bytes(int64_t n_titles, int64_t L, int64_t Hp) {
  const int64_t grid = sm_count();                 // upper bound of tail_grid (the device may differ at planning time)
  (void)n_titles; (void)L;
  return arena_bytes(grid * 2 * Hp, 4) + arena_bytes(2 * grid * 128 * Hp, 4) + arena_bytes(grid * 4 * Hp, 4) + 256;
}

int cnn_tail_bwd(int64_t n_titles, int64_t L, int64_t H, const __nv_bfloat16* c, const __nv_bfloat16* key, const uint8_t* cmask, const float* prob,
                 const float* d_news, const float* query, const uint8_t* wq_img, __nv_bfloat16* dconv, float* d_proj_w,
                 float* d_proj_b, float* d_query, float* d_conv_b, void* ws, int64_t wsb, cudaStream_t st) {
  const int64_t Hp = align_up(H, 16);
  MR_REQUIRE(cnn_tail_supported(L, Hp), MR_ERR_UNSUPPORTED, "cnn_tail_bwd: L=%lld Hp=%lld not supported", (long long)L, (long long)Hp);
  if (n_titles <= 0) return MR_OK;
  CnnTailBwdArgs a{};
  a.n_titles = n_titles; a.L = (int)L; a.G = tail_G(L); a.H = (int)H; a.Hp = (int)Hp;
  a.n_tiles = ceil_div(n_titles, (int64_t)a.G);
  a.nblk = (int)ceil_div(Hp, 32);
  a.n_mt = (int)ceil_div(Hp, 128);
  const int grid = tail_grid(n_titles, L);
  Arena ar(ws, wsb);
  a.part_qb = ar.take<float>((int64_t)grid * 2 * Hp);
  a.part_w = ar.take<float>((int64_t)a.n_mt * grid * 128 * Hp);
  a.csum = ar.take<float>((int64_t)grid * 4 * Hp);
  MR_REQUIRE(ar.ok(), MR_ERR_WORKSPACE, "cnn_tail_bwd: workspace too small (%lld given)", (long long)wsb);
  a.prob = prob; a.d_news = d_news; a.query = query; a.wq_img = wq_img; a.dconv = dconv; a.cmask = cmask;
  const size_t tile_bytes = (size_t)a.nblk * CT_BLK, w_bytes = (size_t)(Hp / 8) * Hp * 16;
  const size_t side = (size_t)a.G * (Hp + 4) * 4;
  a.n_side = 4 * tile_bytes + w_bytes + 3 * side + 2 * 128 * 4 + 13 * 8 + 16 <= 227 * 1024 ? 3 : 2;
  size_t smem = 4 * tile_bytes + w_bytes + a.n_side * side + 2 * 128 * 4 + 13 * 8 + 16;
  const size_t a_reach = 3 * tile_bytes + (size_t)a.n_mt * 4 * CT_BLK;       // the MN-major A operand reads whole 128-column groups
  if (smem < a_reach) smem = a_reach;
  MR_REQUIRE(smem <= 227 * 1024, MR_ERR_UNSUPPORTED, "cnn_tail_bwd: %zu bytes of shared memory", smem);
  a.smem_bytes = (uint32_t)smem;
  a.dbg = g_tapgemm_dbg;
  if (g_tapgemm_dbg != nullptr) g_tapgemm_dbg += 148 * 4 * 5;
  alignas(64) CUtensorMap cmap, kmap;
  if (int rc = tma_encode_3d(&cmap, c, (uint64_t)Hp, (uint64_t)n_titles, (uint64_t)L, (uint64_t)L * Hp * 2, (uint64_t)Hp * 2, 32,
                             (uint32_t)a.G, (uint32_t)L, 64))
    return rc;
  if (int rc = tma_encode_3d(&kmap, key, (uint64_t)Hp, (uint64_t)n_titles, (uint64_t)L, (uint64_t)L * Hp * 2, (uint64_t)Hp * 2, 32,
                             (uint32_t)a.G, (uint32_t)L, 64))
    return rc;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(cnn_tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "cnn_tail_bwd: shared-memory opt-in failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  {
    TimedLaunch tl(2, st);
    launch_pdl(cnn_tail_bwd_kernel, dim3((unsigned)grid), dim3(CT_THREADS), smem, st, a, cmap, kmap);
  }
  MR_CHECK_LAUNCH("cnn_tail_bwd_kernel");
  // fixed-order reductions of the per-CTA partials
  launch_pdl(cnn_pool_bwd_final_kernel, dim3((unsigned)ceil_div(2 * Hp, 32)), dim3(1024), 0, st, (const float*)a.part_qb, (int64_t)grid, (int)Hp,
             (int)H, d_query, d_proj_b);
  MR_CHECK_LAUNCH("cnn_pool_bwd_final_kernel");
  {
    TokRedPlan plan{};
    plan.args.partial = a.part_w; plan.args.S = grid; plan.args.taps = 1; plan.args.NQ = (int)Hp;
    if (int rc = tokred_reduce(plan, d_proj_w, (int)H, (int)H, 1, H, 0, st)) return rc;      // dst[k + n*H]: i = n (dkp column), j = k
  }
  cudaError_t e = colsum_small(a.csum, Hp, d_conv_b, (int64_t)grid * 4, H, st);
  MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "cnn_tail_bwd: colsum: %s", cudaGetErrorString(e));
  return MR_OK;
}

// =====================================================================================================================
// Forward: key = tanh(c Wq^T + bq) (bf16, saved for the backward), p = masked softmax over the title of <q, key> / sqrt(H)
// (all-masked title -> zeros, XSoftmax), news = sum_l p[l] c[l, :]  -- CNN.py:44-48 in one pass over c.
//   TMA warp    c tiles (3-stage ring), Wq resident
//   MMA warp    D = c Wq^T (A = c tile K-major, double-buffered accumulator), then the pooled sum as a second MMA:
//               Dn^T [Hp x 16] = c^T P with P[g, r] = p[r] for the rows r of title g (bf16 high part in row g, low part in
//               row G + g, so that p keeps 16 mantissa bits) -- the c tile is the MN-major A operand, K = token rows
//   warps 0-7   epilogue (token row per thread, two column halves): tanh, key -> global, partial <q, key>; the half-0 warps
//               then run the per-title softmax through shared memory and write prob and the P operand
//   warps 8-11  news[g, h] = Dn[h, g] + Dn[h, G + g] -> global
// =====================================================================================================================
constexpr int CF_THREADS = 448;     // warps 0-7 epilogue, 8-11 news rows, 12 MMA issuer, 13 TMA / weights
constexpr int CF_STAGES = 3;

struct CnnTailFwdArgs {
  int64_t n_titles, n_tiles;
  int L, G, H, Hp, nblk, n_mt;
  const void* mask; int mask_i64;
  const float* query;
  const float* proj_b;
  const uint8_t* wq_img;
  __nv_bfloat16* key;
  float* prob;
  float* news;
  long long* dbg;
};

__device__ __forceinline__ float ct_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(CF_THREADS, 1)
cnn_tail_fwd_kernel(const CnnTailFwdArgs p, const __grid_constant__ CUtensorMap cmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  long long dbg_acc[4] = {0, 0, 0, 0};
  const long long dbg_t0 = clock64();
  const uint32_t tile_bytes = (uint32_t)p.nblk * CT_BLK;
  const uint32_t w_bytes = (uint32_t)(p.Hp / 8) * (uint32_t)p.Hp * 16u;
  uint8_t* sC = smem;                                                  // [CF_STAGES][tile_bytes]
  uint8_t* sW = sC + CF_STAGES * tile_bytes;
  uint8_t* sP = sW + w_bytes;                                          // [2][16 panels][16 rows][16 B]
  float* sScore = reinterpret_cast<float*>(sP + 2 * 4096);             // [2 stages][2 halves][128]
  float* sE = sScore + 2 * 2 * 128;                                    // [2 stages][128]
  float* sQ = sE + 2 * 128;                                            // [Hp] (query / sqrt(H)),
  float* sBias = sQ + p.Hp;                                            // [Hp]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + p.Hp);
  uint64_t* full = bars;                 // [3] c tile landed
  uint64_t* c_free = bars + 3;           // [3] both MMAs that read the stage have completed
  uint64_t* d_full = bars + 6;           // [2]
  uint64_t* d_empty = bars + 8;          // [2] 256 arrivals
  uint64_t* p_ready = bars + 10;         // [2] 128 arrivals (half-0 epilogue threads)
  uint64_t* n_full = bars + 12;          // [2]
  uint64_t* n_empty = bars + 14;         // [2] 128 arrivals
  uint64_t* w_ready = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = p.G, L = p.L, Hp = p.Hp, H = p.H;

  pdl_trigger();
  for (uint32_t i = tid * 16; i < CF_STAGES * tile_bytes; i += CF_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  for (uint32_t i = tid * 16; i < 2 * 4096; i += CF_THREADS * 16) *reinterpret_cast<uint4*>(sP + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int i = 0; i < CF_STAGES; ++i) {
      tc::mbar_init(&full[i], 1);
      tc::mbar_init(&c_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&d_full[i], 1);
      tc::mbar_init(&d_empty[i], 256);
      tc::mbar_init(&p_ready[i], 128);
      tc::mbar_init(&n_full[i], 1);
      tc::mbar_init(&n_empty[i], 128);
    }
    tc::mbar_init(w_ready, 1);
    tc::fence_barrier_init();
  }
  if (warp == 12) tc::tmem_alloc(tmem_slot, 512);
  pdl_wait();                          // everything above touched this CTA's shared memory / TMEM only
  {
    const float inv = rsqrtf((float)H);
    for (int h = tid; h < Hp; h += CF_THREADS) {
      sQ[h] = h < H ? p.query[h] * inv : 0.f;
      sBias[h] = h < H ? p.proj_b[h] : 0.f;
    }
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t DN_COL = 416;                               // D stages at columns 0 and 256 (<= 160 wide), Dn^T stages from here (32 each)

  if (warp < 8) {
    // =================================== epilogue ==========================================================
    const int q4 = warp & 3, half = warp >> 2, r = q4 * 32 + lane;
    const int g = r % G, l = r / G;
    const bool row_ok = l < L;
    const int nc = (Hp + 31) >> 5;
    // half 0 also runs the softmax: it takes the smaller share of the column chunks
    const int c_begin = half ? nc / 2 : 0, c_end = half ? nc : nc / 2;
    uint32_t i = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
      const uint32_t as = i & 1u, aph = (i >> 1) & 1u;
      const int64_t n = tile * G + g, t = n * L + l;
      const bool valid = row_ok && n < p.n_titles;
      bool keep = false;
      if (half == 0 && valid) keep = p.mask ? (load_index(p.mask, p.mask_i64, t) != 0) : true;
      CT_TIMED(0, tc::mbar_wait(&d_full[as], aph));
      tc::tc_fence_after();
      const long long tE0 = clock64();
      const uint32_t tb = tmem + ((uint32_t)(q4 * 32) << 16) + as * 256u;
      float sc = 0.f;
#pragma unroll 1
      for (int ci = c_begin; ci < c_end; ++ci) {
        const int n0 = ci * 32;
        const bool wide = Hp - n0 >= 32;
        uint32_t v[32];
        if (wide) {
          tc::tmem_ld32(tb + n0, v);
        } else {
          uint32_t h16[16];
          tc::tmem_ld16(tb + n0, h16);
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = h16[j]; v[16 + j] = 0u; }
        }
        tc::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 b2 = *reinterpret_cast<const float2*>(sBias + n0 + 2 * j);
          pk[j] = tc::pack_bf16(ct_tanh(__uint_as_float(v[2 * j]) + b2.x), ct_tanh(__uint_as_float(v[2 * j + 1]) + b2.y));
        }
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(p.key + t * Hp + n0);
#pragma unroll
          for (int w = 0; w < 4; w += 2)
            if (w < 2 || wide)
              tc::st_global_256(dst + w, make_uint4(pk[4 * w], pk[4 * w + 1], pk[4 * w + 2], pk[4 * w + 3]),
                                make_uint4(pk[4 * w + 4], pk[4 * w + 5], pk[4 * w + 6], pk[4 * w + 7]));
        }
        // the score uses the bf16-rounded key (what the backward and the stand-alone pooling kernels read)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 q2 = *reinterpret_cast<const float2*>(sQ + n0 + 2 * j);
          sc = fmaf(__uint_as_float(pk[j] << 16), q2.x, sc);
          sc = fmaf(__uint_as_float(pk[j] & 0xFFFF0000u), q2.y, sc);
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&d_empty[as]);
      float* score_s = sScore + as * 256;
      score_s[half * 128 + r] = valid ? sc : 0.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (p.dbg != nullptr) dbg_acc[1] += clock64() - tE0;
      if (half == 0) {
        // ---- masked softmax over the title's rows r' = l' * G + g ----------------------------------------------
        // rows of title g inside this warp: the lanes with lane % G == g (G is a power of two); the four warps' partial
        // results meet in shared memory
        const float s_me = score_s[r] + score_s[128 + r];
        float* e_s = sE + as * 128;                                   // [0, 4G): partial maxima, [64, 64 + 4G): partial sums
        float mx = keep ? s_me : -INFINITY;
        for (int o = G; o < 32; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane < G) e_s[q4 * G + lane] = mx;
        asm volatile("bar.sync 2, 128;" ::: "memory");
        mx = fmaxf(fmaxf(e_s[g], e_s[G + g]), fmaxf(e_s[2 * G + g], e_s[3 * G + g]));
        const float ex = keep ? expf(s_me - mx) : 0.f;
        float sum = ex;
        for (int o = G; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane < G) e_s[64 + q4 * G + lane] = sum;
        asm volatile("bar.sync 2, 128;" ::: "memory");
        sum = (e_s[64 + g] + e_s[64 + G + g]) + (e_s[64 + 2 * G + g] + e_s[64 + 3 * G + g]);
        const float pr = (valid && sum > 0.f) ? ex / sum : 0.f;       // all-masked title -> zeros (XSoftmax, Attention.py:66-74)
        if (valid) p.prob[t] = pr;
        // P operand: bf16 high part in row g, low part in row G + g, column (K index) r
        const __nv_bfloat16 hi = __float2bfloat16(pr);
        const __nv_bfloat16 lo = __float2bfloat16(pr - __bfloat162float(hi));
        uint8_t* pp = sP + as * 4096 + (r >> 3) * 256 + (r & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(pp + g * 16) = hi;
        *reinterpret_cast<__nv_bfloat16*>(pp + (G + g) * 16) = lo;
        tc::fence_proxy_async();
        tc::mbar_arrive(&p_ready[as]);       // (e_s / score_s of this stage are rewritten two tiles later, behind the next tile's barriers)
        if (p.dbg != nullptr) dbg_acc[2] += clock64() - tE0;
      }
    }
  } else if (warp < 12) {
    // =================================== news rows =========================================================
    const int q4 = warp & 3;
    uint32_t i = 0;
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
      const uint32_t as = i & 1u, aph = (i >> 1) & 1u;
      CT_TIMED(0, tc::mbar_wait(&n_full[as], aph));
      tc::tc_fence_after();
      for (int mt = 0; mt < p.n_mt; ++mt) {
        const int h = mt * 128 + q4 * 32 + lane;
        if (mt * 128 + q4 * 32 >= Hp) continue;                      // warp-uniform
        uint32_t v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(q4 * 32) << 16) + DN_COL + as * 32u + (uint32_t)mt * 16u, v);
        tc::tmem_ld_wait();
        if (h < H) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int64_t n = tile * G + g;
            const uint32_t lo = G == 4 ? v[(4 + g) & 15] : v[(8 + g) & 15];
            if (g < G && n < p.n_titles) p.news[n * H + h] = __uint_as_float(v[g]) + __uint_as_float(lo);
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&n_empty[as]);
    }
  } else if (warp == 12) {
    // =================================== MMA issuer =======================================================
    const uint32_t idesc1 = tc::make_idesc(128, Hp, 0, 0), idesc2 = tc::make_idesc(128, 16, 1, 0);
    const uint32_t b_ps = (uint32_t)Hp * 16u;
    const uint32_t wbase = tc::smem_u32(sW), pbase = tc::smem_u32(sP), cbase0 = tc::smem_u32(sC);
    const int nks1 = Hp >> 4;
    tc::mbar_wait(w_ready, 0);
    const int64_t n_mine = p.n_tiles > (int64_t)blockIdx.x ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto mma1 = [&](uint32_t i) {
      const uint32_t cs = i % CF_STAGES, cph = (i / CF_STAGES) & 1u, as = i & 1u, aph = (i >> 1) & 1u;
      CT_TIMED(0, tc::mbar_wait(&full[cs], cph));
      CT_TIMED(1, tc::mbar_wait(&d_empty[as], aph ^ 1u));
      tc::tc_fence_after();
      const uint32_t cbase = cbase0 + cs * tile_bytes;
      if (tc::elect_one()) {
        for (int ks = 0; ks < nks1; ++ks) {
          const uint64_t da = tc::make_desc_sw(cbase + (uint32_t)(ks >> 1) * CT_BLK + (uint32_t)(ks & 1) * 32u, 16, 512, 4, 0);
          const uint64_t dbb = tc::make_desc(wbase + 2u * (uint32_t)ks * b_ps, b_ps, 128);
          tc::umma(tmem + as * 256u, da, dbb, idesc1, ks > 0 ? 1u : 0u);
        }
        tc::umma_commit(&d_full[as]);
      }
      __syncwarp();
    };
    if (n_mine > 0) mma1(0);
    for (uint32_t i = 0; i < (uint32_t)n_mine; ++i) {
      if (i + 1 < (uint32_t)n_mine) mma1(i + 1);
      const uint32_t cs = i % CF_STAGES, as = i & 1u, aph = (i >> 1) & 1u;
      CT_TIMED(2, tc::mbar_wait(&p_ready[as], aph));
      CT_TIMED(3, tc::mbar_wait(&n_empty[as], aph ^ 1u));
      tc::tc_fence_after();
      const uint32_t cbase = cbase0 + cs * tile_bytes;
      if (tc::elect_one()) {
        for (int mt = 0; mt < p.n_mt; ++mt) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t da = tc::make_desc_sw(cbase + (uint32_t)mt * 4u * CT_BLK + (uint32_t)ks * 1024u, CT_BLK, 512, 4, 0);
            const uint64_t dbb = tc::make_desc(pbase + as * 4096u + (uint32_t)ks * 512u, 256, 128);
            tc::umma(tmem + DN_COL + as * 32u + (uint32_t)mt * 16u, da, dbb, idesc2, ks > 0 ? 1u : 0u);
          }
        }
        tc::umma_commit(&n_full[as]);
        tc::umma_commit(&c_free[cs]);
      }
      __syncwarp();
    }
  } else {
    // =================================== TMA producer =====================================================
    if (lane == 0) {
      tc::tma_prefetch_desc(&cmap);
      const uint8_t* wsrc = pdl_acquire(p.wq_img);
      tc::mbar_arrive_expect_tx(w_ready, w_bytes);
      for (uint32_t off = 0; off < w_bytes; off += 32768) {
        const uint32_t nb = w_bytes - off < 32768 ? w_bytes - off : 32768;
        tc::bulk_g2s(tc::smem_u32(sW) + off, wsrc + off, nb, w_ready);
      }
      const uint32_t stage_tx = (uint32_t)p.nblk * (uint32_t)(G * L) * 64u;
      uint32_t i = 0;
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++i) {
        const uint32_t cs = i % CF_STAGES, cph = (i / CF_STAGES) & 1u;
        CT_TIMED(0, tc::mbar_wait(&c_free[cs], cph ^ 1u));
        tc::mbar_arrive_expect_tx(&full[cs], stage_tx);
        const uint32_t cdst = tc::smem_u32(sC) + cs * tile_bytes;
        for (int b = 0; b < p.nblk; ++b) tc::tma_load_3d(cdst + (uint32_t)b * CT_BLK, &cmap, b * 32, (int)(tile * G), 0, &full[cs]);
      }
    }
  }
  if (p.dbg != nullptr && lane == 0 && (warp == 0 || warp == 4 || warp == 12 || warp == 13)) {
    // roles: 0 epilogue half 0 {d_full wait, E1, E1 + softmax}, 1 epilogue half 1, 2 MMA {full, d_empty, p_ready, n_empty waits}, 3 TMA {c_free wait}
    long long* d = p.dbg + ((size_t)blockIdx.x * 4 + (warp == 0 ? 0 : warp == 4 ? 1 : warp == 12 ? 2 : 3)) * 5;
    d[0] = dbg_acc[0]; d[1] = dbg_acc[1]; d[2] = dbg_acc[2]; d[3] = dbg_acc[3]; d[4] = clock64() - dbg_t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 12) tc::tmem_dealloc(tmem, 512);
}

int cnn_tail_fwd(int64_t n_titles, int64_t L, int64_t H, const __nv_bfloat16* c, const void* mask, int mask_i64, const float* query,
                 const float* proj_b, const uint8_t* wq_img, __nv_bfloat16* key, float* prob, float* news, cudaStream_t st) {
  const int64_t Hp = align_up(H, 16);
  MR_REQUIRE(cnn_tail_supported(L, Hp), MR_ERR_UNSUPPORTED, "cnn_tail_fwd: L=%lld Hp=%lld not supported", (long long)L, (long long)Hp);
  if (n_titles <= 0) return MR_OK;
  CnnTailFwdArgs a{};
  a.n_titles = n_titles; a.L = (int)L; a.G = tail_G(L); a.H = (int)H; a.Hp = (int)Hp;
  a.n_tiles = ceil_div(n_titles, (int64_t)a.G);
  a.nblk = (int)ceil_div(Hp, 32);
  a.n_mt = (int)ceil_div(Hp, 128);
  a.mask = mask; a.mask_i64 = mask_i64; a.query = query; a.proj_b = proj_b; a.wq_img = wq_img; a.key = key; a.prob = prob; a.news = news;
  const int grid = tail_grid(n_titles, L);
  const size_t tile_bytes = (size_t)a.nblk * CT_BLK, w_bytes = (size_t)(Hp / 8) * Hp * 16;
  size_t smem = CF_STAGES * tile_bytes + w_bytes + 2 * 4096 + (2 * 2 * 128 + 2 * 128 + 2 * Hp) * 4 + 17 * 8 + 16;
  const size_t a_reach = (CF_STAGES - 1) * tile_bytes + (size_t)a.n_mt * 4 * CT_BLK;      // the MN-major A operand reads whole 128-column groups
  if (smem < a_reach) smem = a_reach;
  MR_REQUIRE(smem <= 227 * 1024, MR_ERR_UNSUPPORTED, "cnn_tail_fwd: %zu bytes of shared memory", smem);
  a.dbg = g_tapgemm_dbg;
  if (g_tapgemm_dbg != nullptr) g_tapgemm_dbg += 148 * 4 * 5;
  alignas(64) CUtensorMap cmap;
  if (int rc = tma_encode_3d(&cmap, c, (uint64_t)Hp, (uint64_t)n_titles, (uint64_t)L, (uint64_t)L * Hp * 2, (uint64_t)Hp * 2, 32,
                             (uint32_t)a.G, (uint32_t)L, 64))
    return rc;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(cnn_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "cnn_tail_fwd: shared-memory opt-in failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  launch_pdl(cnn_tail_fwd_kernel, dim3((unsigned)grid), dim3(CF_THREADS), smem, st, a, cmap);
  MR_CHECK_LAUNCH("cnn_tail_fwd_kernel");
  return MR_OK;
}

}  // namespace mr
