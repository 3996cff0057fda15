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
  g.h_bytes = (uint32_t)g.KT * 2u * 256u;
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
  const uint32_t w_ps = (uint32_t)rows * 16u, h_ps = 256u, h_bytes = (uint32_t)KT * 2u * h_ps;
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
    tc::mbar_init(&bars[1], 1);
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
    if (warp == 0) {
      tc::tc_fence_after();
      if (tc::elect_one()) {
#pragma unroll 1
        for (int t = 0; t < MT; ++t) {
          const uint32_t a_col = tmem + (uint32_t)(t * KT * 8);
#pragma unroll
          for (int ks = 0; ks < RT_MAXH / 16; ++ks)
            if (ks < KT) tc::umma_ts(tmem + d_col + (uint32_t)t * 16u, a_col + (uint32_t)ks * 8u, b_lo0 + (uint32_t)ks * b_kstep, b_hi, idesc, (uint32_t)ks);
        }
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

}  // namespace mr
