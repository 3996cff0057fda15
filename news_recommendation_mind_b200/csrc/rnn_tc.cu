// LSTM recurrence of the user encoder on tcgen05 (MR_BF16 path; models/Encoders/RNN.py:36-73, :76-104 for LSTUR).
//
// Per step every sequence needs  pre[4H] = W_hh[4H, H] . h[H].  Round 1 ran this as mma.sync m16n8k16 with W_hh in registers +
// shared memory; legacy HMMA issues at ~20 cycles per instruction on sm_100a, 400 instructions per step, so the MMA phase
// was ~3300 of the ~3900 cycles of a step (99 us for 50 steps).  Here the same product is ONE batch of tcgen05.mma per step:
//   A = W_hh, bf16, RESIDENT IN TENSOR MEMORY for the whole kernel (rows = 4H rounded up to 128 -> MT = rows/128 tiles of M = 128
//       lanes; a row's K elements are packed two per 32-bit column, KT * 8 columns per tile; H = 150: 5 x 80 = 400 of the 512
//       columns).  The first version kept W_hh in shared memory: every step then re-read 200 KB through the 128 B/clk
//       shared-memory port (32 cycles per MMA against the 8-cycle floor of M128 N16 K16) -- with the A operand in TMEM the
//       tensor core only fetches the 512-byte B tile per MMA;
//   B = the hidden states of the CTA's NSEQ sequences, 16 rows x H (panel layout, 5 KB): rows [0, NSEQ) the bf16 high parts,
//       rows [NSEQ, 2 NSEQ) the low parts (h = hi + lo keeps ~16 mantissa bits through the recurrence), remaining rows zero;
//   D = MT accumulators of 128 lanes x 16 columns in TMEM (fp32), behind the A columns.
// MT x KT = 5 x 10 MMAs of M128 N16 K16 per step: 8 cycles each at the tensor pipe's floor against ~3300 cycles before.
// A step:  (1) one elected lane issues the MMAs + tcgen05.commit;  (2) all 16 warps wait on the mbarrier, read their TMEM
// quadrant (tcgen05.ld 32x32b), add the high and low columns and park the pre-activations in shared memory;  (3) gate phase as
// before -- one thread per (sequence, unit): sigmoid / tanh, c and h update (only while s < len: packed-sequence semantics),
// saved tensors to global, new h (hi | lo) into the B tile;  fence.proxy.async + __syncthreads, next step.
// One CTA owns NSEQ = 2 / 4 / 8 sequences for all S steps (B = 256: 128 CTAs): no inter-CTA traffic, no per-step launch.
#include "rnn_res.cuh"
#include "tapgemm.cuh"   // sm_count()
#include "tc05.cuh"
#include <stdlib.h>

namespace mr {

constexpr int RT_THREADS = 512;
constexpr int RT_MAXH = 160;
constexpr int RT_MAXMT = 5;

__device__ __forceinline__ float rt_sigm(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float rt_tanh(float x) {
  const float t = __expf(-2.0f * fminf(fmaxf(x, -15.f), 15.f));
  return __fdividef(1.0f - t, 1.0f + t);
}

namespace tc {
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
}  // namespace tc

struct RTGeom {
  int GH, MT, KT, rows;
  uint32_t w_ps, w_bytes, h_bytes;
  size_t smem;
};
static inline RTGeom rt_geom(int H, int nseq) {
  RTGeom g;
  g.GH = 4 * H;
  g.MT = (g.GH + 127) / 128;
  g.KT = (H + 15) / 16;
  g.rows = g.MT * 128;
  g.w_ps = (uint32_t)g.rows * 16u;
  g.w_bytes = (uint32_t)g.KT * 2u * g.w_ps;
  g.h_bytes = (uint32_t)g.KT * 2u * 272u;      // panel stride 256 + 16: consecutive panels in different banks
  g.smem = (size_t)g.h_bytes + (size_t)g.rows * nseq * 4 + 128;
  return g;
}

static bool rnn_use_tc() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MINDREC_RNN_TC");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

bool rnn_tc_supported(int kind, int H) {
  if (kind != MR_RNN_LSTM || H < 8 || H > RT_MAXH || !rnn_use_tc()) return false;
  const RTGeom g = rt_geom(H, 8);
  return g.MT <= RT_MAXMT && g.MT * (g.KT * 8 + 16) <= 512 && g.smem <= 227 * 1024;
}

int64_t rnn_tc_scratch_bytes(int kind, int H) {
  if (kind != MR_RNN_LSTM || H > RT_MAXH) return 256;
  return (int64_t)rt_geom(H, 8).w_bytes + 256;
}

// W_hh [4H, H] fp32 -> bf16 panel image [KT*2 panels][rows][8], zero padded (the byte image the CTAs bulk-copy)
__global__ void rnn_tc_prep_kernel(const float* __restrict__ w_hh, __nv_bfloat16* __restrict__ img, int GH, int H, int rows, int kc) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * kc) return;
  const int panel = i / (rows * 8), rest = i - panel * rows * 8;
  const int n = rest >> 3, k = panel * 8 + (rest & 7);
  img[i] = __float2bfloat16((n < GH && k < H) ? w_hh[(int64_t)n * H + k] : 0.f);
}

template <int NSEQ>
__global__ void __launch_bounds__(RT_THREADS, 1)
rnn_tc_fwd_kernel(const float* __restrict__ xp, int ldx, const uint8_t* __restrict__ w_img, const float* __restrict__ h0,
                  const int32_t* __restrict__ lens, float* __restrict__ gates, float* __restrict__ hs, float* __restrict__ cs,
                  float* __restrict__ user, int B, int S, int H, int MT, int KT) {
  constexpr int G = 4;
  constexpr int PPT = (NSEQ * RT_MAXH + RT_THREADS - 1) / RT_THREADS;        // (sequence, unit) pairs per thread
  const int GH = G * H, rows = MT * 128;
  const uint32_t w_ps = (uint32_t)rows * 16u, h_ps = 272u, h_bytes = (uint32_t)KT * 2u * h_ps;
  extern __shared__ __align__(1024) uint8_t rt_smem[];
  uint8_t* Hsm = rt_smem;                                           // B operand: [KT*2 panels][16 rows][16 B]
  float* pre = reinterpret_cast<float*>(Hsm + h_bytes);             // [rows][NSEQ]
  uint64_t* bars = reinterpret_cast<uint64_t*>(pre + (size_t)rows * NSEQ);      // [0] weights landed, [1] MMAs of the step done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  int* len_s = reinterpret_cast<int*>(tmem_slot + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, b0 = blockIdx.x * NSEQ;

  pdl_trigger();
  for (uint32_t i = tid * 16; i < h_bytes; i += RT_THREADS * 16) *reinterpret_cast<uint4*>(Hsm + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], (uint32_t)MT);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  pdl_wait();                                  // everything above touched this CTA's shared memory / TMEM only
  w_img = pdl_acquire(w_img);
  xp = pdl_acquire(xp);
  h0 = pdl_acquire(h0);
  lens = pdl_acquire(lens);
  __syncthreads();
  if (tid < NSEQ) {
    const int b = b0 + tid;
    const int l = b < B ? (lens ? lens[b] : S) : 0;
    len_s[tid] = l < 0 ? 0 : (l > S ? S : l);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t d_col = (uint32_t)(MT * KT * 8);                    // accumulators behind the A columns
  // W_hh -> tensor memory: warp w fills lanes 32 (w % 4) .. + 31 of tiles w / 4, w / 4 + 4; a lane reads its row's 16-byte
  // pieces from the panel image (consecutive lanes = consecutive rows: coalesced) and stores 8 columns (K = 16) at a time
  for (int t = warp >> 2; t < MT; t += 4) {
    const int row = t * 128 + (warp & 3) * 32 + lane;
    const uint4* src = reinterpret_cast<const uint4*>(w_img + (size_t)row * 16);
    for (int ks = 0; ks < KT; ++ks) {
      const uint4 p0 = __ldg(src + (size_t)(2 * ks) * rows), p1 = __ldg(src + (size_t)(2 * ks + 1) * rows);
      const uint32_t v[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      tc::tmem_st8(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * KT * 8 + ks * 8), v);
    }
  }
  tc::tmem_st_wait();
  int max_len = 0;
#pragma unroll
  for (int i = 0; i < NSEQ; ++i) max_len = max(max_len, len_s[i]);

  // gate role: pair p = tid + q * THREADS -> (sequence n = p / H, unit j = p % H); h and c live in registers
  int pn[PPT], pj[PPT], plen[PPT];
  float hreg[PPT], creg[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int p = tid + q * RT_THREADS;
    const bool on = p < NSEQ * H;
    pn[q] = on ? p / H : 0;
    pj[q] = on ? p - pn[q] * H : 0;
    const int b = b0 + pn[q];
    plen[q] = (on && b < B) ? len_s[pn[q]] : 0;
    hreg[q] = (on && b < B && h0 != nullptr) ? h0[(int64_t)b * H + pj[q]] : 0.f;
    creg[q] = 0.f;
    if (on && h0 != nullptr) {
      const __nv_bfloat16 hi = __float2bfloat16(hreg[q]);
      uint8_t* dst = Hsm + (size_t)(pj[q] >> 3) * h_ps + (pj[q] & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(dst + pn[q] * 16) = hi;
      *reinterpret_cast<__nv_bfloat16*>(dst + (NSEQ + pn[q]) * 16) = __float2bfloat16(hreg[q] - __bfloat162float(hi));
    }
  }
  tc::fence_proxy_async();                     // the generic-proxy writes of the B tile, before the tensor core reads it
  tc::tc_fence_before();                       // ... and the tcgen05.st of W_hh before the MMAs of another warp
  __syncthreads();

  const uint32_t idesc = tc::make_idesc(128, 16, 0, 0);
  const uint64_t b_tmpl = tc::make_desc(tc::smem_u32(Hsm), h_ps, 128);
  const uint32_t b_lo0 = (uint32_t)b_tmpl, b_hi = (uint32_t)(b_tmpl >> 32);
  const uint32_t b_kstep = (2u * h_ps) >> 4;
  (void)w_ps;
  const int quad = warp & 3;
  uint32_t phase = 0;

  for (int s = 0; s < max_len; ++s) {
    float xv[PPT][G];
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      if (s < plen[q]) {
        const float* xps = xp + ((int64_t)(b0 + pn[q]) * S + s) * ldx + pj[q];
#pragma unroll
        for (int g = 0; g < G; ++g) xv[q][g] = __ldg(xps + g * H);          // consumed after the MMA phase: latency hidden
      }
    }
    // ---- (1) D[t] = W_hh[tile t] . [h_hi | h_lo]^T ------------------------------------------------------------------
    // one issuing warp per M tile (MT arrivals per step on the barrier): the tensor core takes ~22 cycles per M128 N16 K16 MMA
    // whoever issues it, but the issue loops of the MT independent accumulators overlap
    if (warp < MT) {
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t a_col = tmem + (uint32_t)(warp * KT * 8);
#pragma unroll
        for (int ks = 0; ks < RT_MAXH / 16; ++ks)
          if (ks < KT) tc::umma_ts(tmem + d_col + (uint32_t)warp * 16u, a_col + (uint32_t)ks * 8u, b_lo0 + (uint32_t)ks * b_kstep, b_hi, idesc, (uint32_t)ks);
        tc::umma_commit(&bars[1]);
      }
      __syncwarp();
    }
    // ---- (2) TMEM -> pre[row][n] = high + low ----------------------------------------------------------------------
    tc::mbar_wait(&bars[1], phase);
    phase ^= 1u;
    tc::tc_fence_after();
    for (int t = warp >> 2; t < MT; t += 4) {
      const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + d_col + (uint32_t)t * 16u;
      float* prow = pre + (size_t)(t * 128 + quad * 32 + lane) * NSEQ;
      if constexpr (NSEQ == 2) {
        uint32_t v[4];
        tc::tmem_ld4(taddr, v);
        tc::tmem_ld_wait();
        *reinterpret_cast<float2*>(prow) = make_float2(__uint_as_float(v[0]) + __uint_as_float(v[2]), __uint_as_float(v[1]) + __uint_as_float(v[3]));
      } else if constexpr (NSEQ == 4) {
        uint32_t v[8];
        tc::tmem_ld8(taddr, v);
        tc::tmem_ld_wait();
        *reinterpret_cast<float4*>(prow) = make_float4(__uint_as_float(v[0]) + __uint_as_float(v[4]), __uint_as_float(v[1]) + __uint_as_float(v[5]),
                                                       __uint_as_float(v[2]) + __uint_as_float(v[6]), __uint_as_float(v[3]) + __uint_as_float(v[7]));
      } else {
        uint32_t v[16];
        tc::tmem_ld16(taddr, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < 8; n += 4)
          *reinterpret_cast<float4*>(prow + n) = make_float4(__uint_as_float(v[n]) + __uint_as_float(v[8 + n]), __uint_as_float(v[n + 1]) + __uint_as_float(v[9 + n]),
                                                             __uint_as_float(v[n + 2]) + __uint_as_float(v[10 + n]), __uint_as_float(v[n + 3]) + __uint_as_float(v[11 + n]));
      }
    }
    tc::tc_fence_before();
    __syncthreads();
    // ---- (3) gates ----------------------------------------------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      if (s < plen[q]) {
        const int n = pn[q], j = pj[q], b = b0 + n;
        float pr[G];
#pragma unroll
        for (int g = 0; g < G; ++g) pr[g] = pre[(size_t)(g * H + j) * NSEQ + n];
        float* gs = gates + ((int64_t)b * S + s) * GH + j;
        const int64_t o = ((int64_t)b * S + s) * H + j;
        const float gi = rt_sigm(xv[q][0] + pr[0]), gf = rt_sigm(xv[q][1] + pr[1]);
        const float gg = rt_tanh(xv[q][2] + pr[2]), go = rt_sigm(xv[q][3] + pr[3]);
        const float c = gf * creg[q] + gi * gg;
        const float h = go * rt_tanh(c);
        gs[0] = gi; gs[H] = gf; gs[2 * H] = gg; gs[3 * H] = go;
        creg[q] = c;
        cs[o] = c; hs[o] = h;
        hreg[q] = h;
        const __nv_bfloat16 hi = __float2bfloat16(h);
        uint8_t* dst = Hsm + (size_t)(j >> 3) * h_ps + (j & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(dst + n * 16) = hi;
        *reinterpret_cast<__nv_bfloat16*>(dst + (NSEQ + n) * 16) = __float2bfloat16(h - __bfloat162float(hi));
      }
    }
    tc::fence_proxy_async();
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int p = tid + q * RT_THREADS;
    if (p < NSEQ * H && b0 + pn[q] < B) user[(int64_t)(b0 + pn[q]) * H + pj[q]] = hreg[q];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static int rt_nseq(int B) {
  const int sms = sm_count();
  if ((B + 1) / 2 <= sms) return 2;
  if ((B + 3) / 4 <= sms) return 4;
  return 8;
}

int rnn_tc_fwd(int kind, const float* xp, int ldx, const float* w_hh_f32, const float* h0, const int32_t* lens, float* gates, float* hs,
               float* cs, float* user, int B, int S, int H, void* scratch, cudaStream_t st) {
  MR_REQUIRE(kind == MR_RNN_LSTM, MR_ERR_UNSUPPORTED, "rnn_tc_fwd: LSTM only");
  const int nseq = rt_nseq(B);
  const RTGeom g = rt_geom(H, nseq);
  __nv_bfloat16* img = static_cast<__nv_bfloat16*>(scratch);
  const int kc = g.KT * 16;
  launch_pdl(rnn_tc_prep_kernel, dim3((unsigned)ceil_div((int64_t)g.rows * kc, 256)), dim3(256), 0, st, w_hh_f32, img, g.GH, H, g.rows, kc);
  MR_CHECK_LAUNCH("rnn_tc_prep_kernel");
  const unsigned grid = (unsigned)ceil_div(B, nseq);
  const uint8_t* wimg = reinterpret_cast<const uint8_t*>(img);
#define RT_LAUNCH_F(NS)                                                                                                  \
  {                                                                                                                      \
    cudaError_t e = cudaFuncSetAttribute(rnn_tc_fwd_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem); \
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "rnn_tc_fwd: shared-memory opt-in failed: %s", cudaGetErrorString(e));    \
    launch_pdl(rnn_tc_fwd_kernel<NS>, dim3(grid), dim3(RT_THREADS), g.smem, st, xp, ldx, wimg, h0, lens, gates, hs, cs, user, B, S, H, g.MT, g.KT); \
  }
  if (nseq == 2) RT_LAUNCH_F(2) else if (nseq == 4) RT_LAUNCH_F(4) else RT_LAUNCH_F(8)
#undef RT_LAUNCH_F
  MR_CHECK_LAUNCH("rnn_tc_fwd_kernel");
  return MR_OK;
}

// =====================================================================================================================
// Backward recurrence of the LSTM on tcgen05.  Per step every sequence needs  dh_prev[H] = W_hh^T[H, 4H] . d[4H]  (d = the gate
// pre-activation gradients of the step).  M = the hidden index k, K = the gate index n, N = sequences (bf16 high | low parts
// of d, so that d keeps ~16 mantissa bits).  W_hh^T is RESIDENT IN TENSOR MEMORY as the A operand:
//   tile 0   k < 128:           128 lanes x K0 = 4H (-> 16) gate rows: K0 / 2 columns (H = 150: 304)
//   tile 1   k = 128 .. H - 1:  a second 128-lane x K0 tile would not fit the 512 columns.  Its (at most 32) rows are folded:
//            lane m = 32 c + k' holds chunk c (of four, CL = K0/4 -> 16 gate rows each) of row 128 + k', and the B operand
//            carries the four chunks of d side by side in its N dimension (column block c' = chunk c' of every sequence).
//            D1[32 c + k', block c'] is the partial sum over chunk c when c' == c (the other blocks are ignored), and
//            dh_prev[128 + k'] is the sum of the four diagonal partials: CL / 2 = 80 columns instead of 304.
// A step: (1) one thread per (sequence, unit) turns the carried dh / dc and the saved gates into d, writes the bf16 rows the
// weight-gradient GEMMs read, its bias partial sums and the two B tiles; (2) one elected lane issues K0/16 + CL/16 MMAs of
// M128 N16 (N = 8 NSEQ for tile 1) K16; (3) eight warps read the accumulators (tcgen05.ld), add high and low parts and leave
// dh_prev in shared memory for step (1) of the next (earlier) time step.
// =====================================================================================================================
struct RBGeom {
  int GH, KT0, CL, KT1, NB1, a1_col, d0_col, d1_col, cols;
  uint32_t b0_bytes, b1_bytes;
  size_t smem;
};
static inline RBGeom rb_geom(int H, int nseq) {
  RBGeom g;
  g.GH = 4 * H;
  g.KT0 = (g.GH + 15) / 16;
  g.CL = ((g.KT0 * 16 + 3) / 4 + 15) / 16 * 16;
  g.KT1 = H > 128 ? g.CL / 16 : 0;
  g.NB1 = 8 * nseq;
  g.a1_col = g.KT0 * 8;
  g.d0_col = g.a1_col + g.KT1 * 8;
  g.d1_col = g.d0_col + 16;
  g.cols = g.d1_col + (g.KT1 ? g.NB1 : 0);
  g.b0_bytes = (uint32_t)g.KT0 * 2u * 272u;                        // panel strides + 16 bytes: consecutive panels in different banks
  g.b1_bytes = (uint32_t)g.KT1 * 2u * ((uint32_t)g.NB1 * 16u + 16u);
  g.smem = (size_t)g.b0_bytes + g.b1_bytes + (size_t)nseq * 128 * 4 + (size_t)4 * nseq * 32 * 4 + 256;
  return g;
}

bool rnn_tc_bwd_supported(int kind, int H, int B) {
  if (kind != MR_RNN_LSTM || H < 8 || H > RT_MAXH || !rnn_use_tc()) return false;
  const int nseq = (B + 1) / 2 <= sm_count() ? 2 : 4;
  const RBGeom g = rb_geom(H, nseq);
  return g.cols <= 512 && g.smem <= 227 * 1024;
}
int rnn_tc_bwd_rows(int B) { return (int)ceil_div(B, (B + 1) / 2 <= sm_count() ? 2 : 4); }
int64_t rnn_tc_bwd_scratch_bytes(int kind, int H) {
  if (kind != MR_RNN_LSTM || H > RT_MAXH) return 256;
  const RBGeom g = rb_geom(H, 2);
  return (int64_t)(g.KT0 + g.KT1) * 2 * 128 * 16 + 256;
}

// W_hh [4H, H] fp32 -> bf16 images [k-step panels][128 lanes][8]: img0[pn][k][e] = W_hh[8 pn + e][k],
// img1[pn][32 c + k'][e] = W_hh[CL c + 8 pn + e][128 + k']   (zero outside the matrix)
__global__ void rnn_tc_bwd_prep_kernel(const float* __restrict__ w_hh, __nv_bfloat16* __restrict__ img0, __nv_bfloat16* __restrict__ img1,
                                       int GH, int H, int KT0, int CL, int KT1) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n0 = KT0 * 2 * 128 * 8, n1 = KT1 * 2 * 128 * 8;
  if (i < n0) {
    const int e = i & 7, k = (i >> 3) & 127, pn = i >> 10;
    const int n = 8 * pn + e;
    img0[i] = __float2bfloat16((n < GH && k < H) ? w_hh[(int64_t)n * H + k] : 0.f);
  } else if (i < n0 + n1) {
    const int ii = i - n0;
    const int e = ii & 7, m = (ii >> 3) & 127, pn = ii >> 10;
    const int c = m >> 5, k = 128 + (m & 31), kk = 8 * pn + e, n = CL * c + kk;
    img1[ii] = __float2bfloat16((kk < CL && n < GH && k < H) ? w_hh[(int64_t)n * H + k] : 0.f);
  }
}

template <int NSEQ>
__global__ void __launch_bounds__(RT_THREADS, 1)
rnn_tc_bwd_kernel(const uint8_t* __restrict__ img0, const uint8_t* __restrict__ img1, const int32_t* __restrict__ lens,
                  const float* __restrict__ gates, const float* __restrict__ cs, const float* __restrict__ d_user,
                  __nv_bfloat16* __restrict__ gib, int GHp16, float* __restrict__ d_h0, float* __restrict__ bias_part, int B, int S, int H,
                  const RBGeom g, long long* dbg) {
  long long dbg_acc[4] = {0, 0, 0, 0}, dbg_t = 0;
  const long long dbg_t0 = clock64();
  constexpr int G = 4;
  constexpr int PPT = (NSEQ * RT_MAXH + RT_THREADS - 1) / RT_THREADS;
  constexpr int NB1 = 8 * NSEQ;
  const int GH = G * H;
  extern __shared__ __align__(1024) uint8_t rb_smem[];
  uint8_t* B0 = rb_smem;                                             // [KT0*2 panels][16 rows][16 B]: rows s / NSEQ + s = high / low part of sequence s
  uint8_t* B1 = B0 + g.b0_bytes;                                     // [KT1*2 panels][NB1 rows][16 B]: row 2 NSEQ c + (s | NSEQ + s), chunk c
  float* dh_s = reinterpret_cast<float*>(B1 + g.b1_bytes);           // [NSEQ][128]
  float* dh1_s = dh_s + NSEQ * 128;                                  // [4 chunks][NSEQ][32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dh1_s + 4 * NSEQ * 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  int* len_s = reinterpret_cast<int*>(tmem_slot + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, b0 = blockIdx.x * NSEQ;

  pdl_trigger();
  for (uint32_t i = tid * 16; i < g.b0_bytes + g.b1_bytes; i += RT_THREADS * 16) *reinterpret_cast<uint4*>(rb_smem + i) = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  pdl_wait();
  img0 = pdl_acquire(img0);
  img1 = pdl_acquire(img1);
  lens = pdl_acquire(lens);
  gates = pdl_acquire(gates);
  cs = pdl_acquire(cs);
  d_user = pdl_acquire(d_user);
  if (tid < NSEQ) {
    const int b = b0 + tid;
    const int l = b < B ? (lens ? lens[b] : S) : 0;
    len_s[tid] = l < 0 ? 0 : (l > S ? S : l);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // W_hh^T -> tensor memory (see rnn_tc_fwd_kernel): the four warps of a lane quadrant share the k-steps of both tiles
  {
    const int q4 = warp & 3, row = q4 * 32 + lane;
    const uint32_t tl = tmem + ((uint32_t)(q4 * 32) << 16);
    for (int ks = warp >> 2; ks < g.KT0 + g.KT1; ks += 4) {
      const bool t1 = ks >= g.KT0;
      const int kl = t1 ? ks - g.KT0 : ks;
      const uint4* src = reinterpret_cast<const uint4*>(t1 ? img1 : img0) + (size_t)(2 * kl) * 128 + row;
      const uint4 p0 = __ldg(src), p1 = __ldg(src + 128);
      const uint32_t v[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
      tc::tmem_st8(tl + (uint32_t)((t1 ? g.a1_col : 0) + kl * 8), v);
    }
    tc::tmem_st_wait();
  }
  int max_len = 0;
#pragma unroll
  for (int i = 0; i < NSEQ; ++i) max_len = max(max_len, len_s[i]);

  int pn[PPT], pj[PPT], plen[PPT];
  bool pon[PPT];
  float dh_c[PPT], dc_c[PPT], bsum[PPT][G];
  uint32_t off0[PPT][G], off1[PPT][G];
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int p = tid + q * RT_THREADS;
    pon[q] = p < NSEQ * H;
    pn[q] = pon[q] ? p / H : 0;
    pj[q] = pon[q] ? p - pn[q] * H : 0;
    plen[q] = (pon[q] && b0 + pn[q] < B) ? len_s[pn[q]] : 0;
    dh_c[q] = 0.f; dc_c[q] = 0.f;
#pragma unroll
    for (int gg = 0; gg < G; ++gg) {
      bsum[q][gg] = 0.f;
      // where this pair's gate gradients go in the two B tiles (high part; the low part is NSEQ rows further)
      const int gn = gg * H + pj[q], c = gn / g.CL, kk = gn - c * g.CL;
      off0[q][gg] = (uint32_t)((gn >> 3) * 272 + (gn & 7) * 2 + pn[q] * 16);
      off1[q][gg] = (uint32_t)((kk >> 3) * (NB1 * 16 + 16) + (2 * NSEQ * c + pn[q]) * 16 + (kk & 7) * 2);
    }
  }
  tc::tc_fence_before();
  __syncthreads();

  const uint32_t idesc0 = tc::make_idesc(128, 16, 0, 0), idesc1 = tc::make_idesc(128, NB1, 0, 0);
  const uint64_t b0_tmpl = tc::make_desc(tc::smem_u32(B0), 272, 128), b1_tmpl = tc::make_desc(tc::smem_u32(B1), NB1 * 16 + 16, 128);
  const uint32_t b0_lo = (uint32_t)b0_tmpl, b0_hi = (uint32_t)(b0_tmpl >> 32), b1_lo = (uint32_t)b1_tmpl, b1_hi = (uint32_t)(b1_tmpl >> 32);
  const uint32_t b0_kstep = (2u * 272u) >> 4, b1_kstep = (2u * (NB1 * 16u + 16u)) >> 4;
  uint32_t phase = 0;

  // saved tensors of the step, loaded one step ahead (they do not depend on the carried gradients)
  float sv[PPT][6];
  auto load_step = [&](int t) {
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      if (t >= 0 && t < plen[q]) {
        const int b = b0 + pn[q];
        const float* gs = gates + ((int64_t)b * S + t) * GH + pj[q];
        const int64_t o = ((int64_t)b * S + t) * H + pj[q];
        sv[q][0] = __ldg(gs); sv[q][1] = __ldg(gs + H); sv[q][2] = __ldg(gs + 2 * H); sv[q][3] = __ldg(gs + 3 * H);
        sv[q][4] = __ldg(cs + o);
        sv[q][5] = t > 0 ? __ldg(cs + o - H) : 0.f;
      }
    }
  };
  load_step(max_len - 1);

  for (int t = max_len - 1; t >= 0; --t) {
    if (dbg != nullptr) dbg_t = clock64();
    // ---- (1) gate gradients --------------------------------------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      if (!pon[q]) continue;
      const int n = pn[q], j = pj[q], b = b0 + n;
      float d[G] = {0.f, 0.f, 0.f, 0.f};
      if (t < plen[q]) {
        float dh = dh_c[q];
        if (t == plen[q] - 1) dh += __ldg(d_user + (int64_t)b * H + j);
        const float gi = sv[q][0], gf = sv[q][1], gg = sv[q][2], go = sv[q][3];
        const float tcv = tanhf(sv[q][4]);
        const float dc = dc_c[q] + dh * go * (1.f - tcv * tcv);
        d[0] = dc * gg * gi * (1.f - gi);
        d[1] = dc * sv[q][5] * gf * (1.f - gf);
        d[2] = dc * gi * (1.f - gg * gg);
        d[3] = dh * tcv * go * (1.f - go);
        dc_c[q] = dc * gf;
        __nv_bfloat16* go16 = gib + ((int64_t)b * S + t) * GHp16 + j;
#pragma unroll
        for (int x = 0; x < G; ++x) { go16[x * H] = __float2bfloat16(d[x]); bsum[q][x] += d[x]; }
      }
      if (dbg != nullptr && q == 0) { const long long c = clock64(); dbg_acc[0] += c - dbg_t; dbg_t = c; }
#pragma unroll
      for (int x = 0; x < G; ++x) {
        const __nv_bfloat16 hi = __float2bfloat16(d[x]);
        const __nv_bfloat16 lo = __float2bfloat16(d[x] - __bfloat162float(hi));
        *reinterpret_cast<__nv_bfloat16*>(B0 + off0[q][x]) = hi;
        *reinterpret_cast<__nv_bfloat16*>(B0 + off0[q][x] + NSEQ * 16) = lo;
        if (g.KT1) {
          *reinterpret_cast<__nv_bfloat16*>(B1 + off1[q][x]) = hi;
          *reinterpret_cast<__nv_bfloat16*>(B1 + off1[q][x] + NSEQ * 16) = lo;
        }
      }
    }
    if (dbg != nullptr) { const long long c = clock64(); dbg_acc[1] += c - dbg_t; dbg_t = c; }
    tc::fence_proxy_async();
    if (dbg != nullptr) { const long long c = clock64(); dbg_acc[2] += c - dbg_t; dbg_t = c; }
    __syncthreads();
    if (dbg != nullptr) { const long long c = clock64(); dbg_acc[3] += c - dbg_t; dbg_t = c; }
    // ---- (2) dh_prev = W_hh^T d ----------------------------------------------------------------------------------------
    if (warp == 0) {
      tc::tc_fence_after();
      if (tc::elect_one()) {
#pragma unroll 2
        for (int ks = 0; ks < g.KT0; ++ks)
          tc::umma_ts(tmem + (uint32_t)g.d0_col, tmem + (uint32_t)(ks * 8), b0_lo + (uint32_t)ks * b0_kstep, b0_hi, idesc0, (uint32_t)ks);
        for (int ks = 0; ks < g.KT1; ++ks)
          tc::umma_ts(tmem + (uint32_t)g.d1_col, tmem + (uint32_t)(g.a1_col + ks * 8), b1_lo + (uint32_t)ks * b1_kstep, b1_hi, idesc1, (uint32_t)ks);
        tc::umma_commit(&bars[0]);
      }
      __syncwarp();
    }
    load_step(t - 1);                          // in flight while the tensor core works
    // ---- (3) accumulators -> dh_prev in shared memory --------------------------------------------------------------
    tc::mbar_wait(&bars[0], phase);
    phase ^= 1u;
    tc::tc_fence_after();
    if (warp < 4) {
      uint32_t v[2 * NSEQ];
      const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)g.d0_col;
      if constexpr (NSEQ == 2) tc::tmem_ld4(ta, v); else tc::tmem_ld8(ta, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int n = 0; n < NSEQ; ++n) dh_s[n * 128 + warp * 32 + lane] = __uint_as_float(v[n]) + __uint_as_float(v[NSEQ + n]);
    } else if (warp < 8 && g.KT1) {
      const int c = warp & 3;
      uint32_t v[2 * NSEQ];
      const uint32_t ta = tmem + ((uint32_t)(c * 32) << 16) + (uint32_t)(g.d1_col + 2 * NSEQ * c);
      if constexpr (NSEQ == 2) tc::tmem_ld4(ta, v); else tc::tmem_ld8(ta, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int n = 0; n < NSEQ; ++n) dh1_s[(c * NSEQ + n) * 32 + lane] = __uint_as_float(v[n]) + __uint_as_float(v[NSEQ + n]);
    }
    tc::tc_fence_before();
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      if (pon[q] && t < plen[q]) {
        const int n = pn[q], j = pj[q];
        dh_c[q] = j < 128 ? dh_s[n * 128 + j]
                          : (dh1_s[(0 * NSEQ + n) * 32 + j - 128] + dh1_s[(1 * NSEQ + n) * 32 + j - 128]) +
                            (dh1_s[(2 * NSEQ + n) * 32 + j - 128] + dh1_s[(3 * NSEQ + n) * 32 + j - 128]);
      }
    }
  }
  if (dbg != nullptr && tid == 0) {
    // warp 0, whole kernel: {gate math, B-tile stores, fence.proxy.async, __syncthreads} of step (1)
    long long* d = dbg + (size_t)blockIdx.x * 4 * 5;
    d[0] = dbg_acc[0]; d[1] = dbg_acc[1]; d[2] = dbg_acc[2]; d[3] = dbg_acc[3]; d[4] = clock64() - dbg_t0;
  }
#pragma unroll
  for (int q = 0; q < PPT; ++q)
    if (d_h0 != nullptr && pon[q] && b0 + pn[q] < B) d_h0[(int64_t)(b0 + pn[q]) * H + pj[q]] = dh_c[q];
  // bias-gradient partials of this CTA: sum over its sequences in a fixed order through shared memory (the B0 tile is free now)
  __syncthreads();
  float* fold = reinterpret_cast<float*>(B0);                          // [NSEQ][GH]
#pragma unroll
  for (int q = 0; q < PPT; ++q)
    if (pon[q])
#pragma unroll
      for (int x = 0; x < G; ++x) fold[(size_t)pn[q] * GH + x * H + pj[q]] = bsum[q][x];
  __syncthreads();
  for (int n = tid; n < GH; n += RT_THREADS) {
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < NSEQ; ++q) a += fold[(size_t)q * GH + n];
    bias_part[((size_t)blockIdx.x * 2 + 0) * GH + n] = a;
    bias_part[((size_t)blockIdx.x * 2 + 1) * GH + n] = a;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int rnn_tc_bwd(int kind, const float* w_hh_f32, const int32_t* lens, const float* gates, const float* cs, const float* d_user,
               __nv_bfloat16* gib, int GHp16, float* d_h0, float* bias_part, int B, int S, int H, void* scratch, cudaStream_t st) {
  MR_REQUIRE(kind == MR_RNN_LSTM, MR_ERR_UNSUPPORTED, "rnn_tc_bwd: LSTM only");
  const int nseq = (B + 1) / 2 <= sm_count() ? 2 : 4;
  const RBGeom g = rb_geom(H, nseq);
  MR_REQUIRE(g.cols <= 512 && (size_t)nseq * g.GH * 4 <= g.b0_bytes, MR_ERR_UNSUPPORTED, "rnn_tc_bwd: H=%d does not fit", H);
  __nv_bfloat16* img0 = static_cast<__nv_bfloat16*>(scratch);
  __nv_bfloat16* img1 = img0 + (size_t)g.KT0 * 2 * 128 * 8;
  const int total = (g.KT0 + g.KT1) * 2 * 128 * 8;
  launch_pdl(rnn_tc_bwd_prep_kernel, dim3((unsigned)ceil_div(total, 256)), dim3(256), 0, st, w_hh_f32, img0, img1, g.GH, H, g.KT0, g.CL, g.KT1);
  MR_CHECK_LAUNCH("rnn_tc_bwd_prep_kernel");
  const unsigned grid = (unsigned)ceil_div(B, nseq);
  const uint8_t* i0 = reinterpret_cast<const uint8_t*>(img0);
  const uint8_t* i1 = reinterpret_cast<const uint8_t*>(img1);
  long long* dbg = g_tapgemm_dbg;                       // mr_debug_tapgemm_counters: [grid][4][5] cycle counters of this launch
  if (g_tapgemm_dbg != nullptr) g_tapgemm_dbg += 148 * 4 * 5;
#define RB_LAUNCH(NS)                                                                                                    \
  {                                                                                                                      \
    cudaError_t e = cudaFuncSetAttribute(rnn_tc_bwd_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem); \
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "rnn_tc_bwd: shared-memory opt-in failed: %s", cudaGetErrorString(e));    \
    launch_pdl(rnn_tc_bwd_kernel<NS>, dim3(grid), dim3(RT_THREADS), g.smem, st, i0, i1, lens, gates, cs, d_user, gib, GHp16, d_h0, bias_part, B, S, H, g, dbg); \
  }
  if (nseq == 2) RB_LAUNCH(2) else RB_LAUNCH(4)
#undef RB_LAUNCH
  MR_CHECK_LAUNCH("rnn_tc_bwd_kernel");
  return MR_OK;
}

}  // namespace mr
