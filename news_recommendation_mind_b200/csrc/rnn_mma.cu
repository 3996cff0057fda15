// Persistent LSTM / GRU recurrence on warp-level tensor-core MMAs, recurrent weights resident on chip
// (MR_BF16 path).  Reference: models/Encoders/RNN.py:36-73 (pack_padded_sequence + nn.LSTM / nn.GRU, h_n), :76-104 (LSTUR).
//
// The per-step product  pre[GH, n] = W_hh[GH, H] . h[H, n]  is latency critical (S dependent steps) and far too small
// for tcgen05 (N = a handful of sequences): it runs as mma.sync m16n8k16 (bf16 x bf16 -> fp32) with
//   * W_hh as the A operand, converted to bf16 once per launch; k-tiles < KREG live in REGISTERS for the whole
//     kernel (each of the 16 warps owns up to three 16-row tiles), the remaining k-tiles in shared memory and are
//     fetched with ldmatrix (conflict-free padded pitch) -- no per-step global or L2 traffic for the weights;
//   * h as the B operand, split into bf16 high + low parts (side by side in the 8 MMA columns when a CTA owns <= 4
//     sequences, two MMAs per A fragment otherwise), so the recurrence keeps ~16 mantissa bits of the fp32 hidden state; cell state, gate math and all saved tensors are fp32.
// One CTA owns NSEQ (2/4/8) sequences for all S steps: no inter-CTA traffic, no per-step launch.
// Measured (B200, H=150, 2 sequences per CTA): with one MMA per part (800 per step) the MMA phase was ~4000 of the ~4600
// cycles of a step -- legacy HMMA issues at ~20-33 cycles per m16n8k16 per SM sub-partition on sm_100a; packing the two
// parts into the columns of one MMA (400 per step) brought the step to ~3900 cycles; the gate phase is ~600 cycles.
// Step = (1) MMA phase -> pre-activations to shared memory, (2) gate phase: one thread per (sequence, unit).
#include "rnn_res.cuh"
#include "tapgemm.cuh"   // sm_count()

namespace mr {

constexpr int RM_THREADS = 512;
constexpr int RM_WARPS = 16;
constexpr int RM_MAXMT = 3;      // m-tiles per warp  -> G*H <= 768
constexpr int RM_MAXKT = 10;     // k-tiles           -> H <= 160
constexpr int RM_KREG = 4;       // k-tiles held in registers

__device__ __forceinline__ float rm_sigm(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// tanh through one ex2 + one rcp (|error| ~1e-7): the libm tanhf is a long instruction sequence on the critical path of every step
__device__ __forceinline__ float rm_tanh(float x) {
  const float t = __expf(-2.0f * fminf(fmaxf(x, -15.f), 15.f));
  return __fdividef(1.0f - t, 1.0f + t);
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t smem_addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_addr));
}

// W_hh [GH, H] fp32 -> bf16 image [MT*16][KT*16], zero padded
__global__ void rnn_mma_prep_kernel(const float* __restrict__ w_hh, __nv_bfloat16* __restrict__ img, int GH, int H, int rows, int cols) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int n = i / cols, k = i - n * cols;
  img[i] = __float2bfloat16((n < GH && k < H) ? w_hh[(int64_t)n * H + k] : 0.f);
}

struct RMGeom {
  int G, GH, MT, KT, wp, hp;          // wp / hp: pitches (elements) of the shared-memory W part and of h
  size_t smem;
};
static inline RMGeom rm_geom(int kind, int H, int nseq) {
  RMGeom g;
  g.G = kind == MR_RNN_LSTM ? 4 : 3;
  g.GH = g.G * H;
  g.MT = (g.GH + 15) / 16;
  g.KT = (H + 15) / 16;
  g.wp = g.KT > RM_KREG ? (g.KT - RM_KREG) * 16 + 8 : 0;
  g.hp = g.KT * 16 + 8;
  g.smem = (size_t)g.MT * 16 * g.wp * 2 + (size_t)2 * 8 * g.hp * 2 + (size_t)g.MT * 16 * nseq * 4 + 64;
  return g;
}

bool rnn_mma_supported(int kind, int H) {
  const RMGeom g = rm_geom(kind, H, 8);
  return g.MT <= RM_MAXMT * RM_WARPS && g.KT <= RM_MAXKT && g.smem <= 227 * 1024;
}

int64_t rnn_mma_scratch_bytes(int kind, int H) {
  const RMGeom g = rm_geom(kind, H, 8);
  return (int64_t)g.MT * 16 * g.KT * 16 * 2 + 256;
}

template <int KIND, int NSEQ>
__global__ void __launch_bounds__(RM_THREADS, 1)
rnn_mma_fwd_kernel(const float* __restrict__ xp, int ldx, const __nv_bfloat16* __restrict__ w_img, const float* __restrict__ b_hh,
                   const float* __restrict__ h0, const int32_t* __restrict__ lens, float* __restrict__ gates,
                   float* __restrict__ hs, float* __restrict__ cs, float* __restrict__ user, int B, int S, int H, int MT, int KT) {
  pdl_trigger();
  pdl_wait();
  w_img = pdl_acquire(w_img);      // written by rnn_mma_prep_kernel, the grid this one programmatically depends on (see common.cuh)
  xp = pdl_acquire(xp);
  h0 = pdl_acquire(h0);
  lens = pdl_acquire(lens);
  constexpr int G = KIND == 0 ? 4 : 3;
  constexpr int PPT = (NSEQ * RM_MAXKT * 16 + RM_THREADS - 1) / RM_THREADS;       // (sequence, unit) pairs per thread
  const int GH = G * H;
  const int wp = KT > RM_KREG ? (KT - RM_KREG) * 16 + 8 : 0, hp = KT * 16 + 8, kc = KT * 16;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Wsm = reinterpret_cast<__nv_bfloat16*>(smem_raw);                              // [MT*16][wp]   k-tiles >= KREG
  __nv_bfloat16* hsm = Wsm + (size_t)MT * 16 * wp;                                              // [2][8][hp]    h high / low parts
  float* pre = reinterpret_cast<float*>(hsm + 2 * 8 * hp);                                      // [MT*16][NSEQ]
  int* len_s = reinterpret_cast<int*>(pre + (size_t)MT * 16 * NSEQ);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, b0 = blockIdx.x * NSEQ;
  const int gq = lane >> 2, tq = lane & 3;

  // ---- one-time: weights on chip, initial state ----------------------------------------------------------------
  uint32_t areg[RM_MAXMT][RM_KREG][4];
#pragma unroll
  for (int i = 0; i < RM_MAXMT; ++i) {
    const int mt = warp + RM_WARPS * i;
#pragma unroll
    for (int kt = 0; kt < RM_KREG; ++kt) {
#pragma unroll
      for (int r = 0; r < 4; ++r) areg[i][kt][r] = 0u;
      if (mt < MT && kt < KT) {
        const __nv_bfloat16* base = w_img + (size_t)(mt * 16 + gq) * kc + kt * 16 + 2 * tq;
        areg[i][kt][0] = *reinterpret_cast<const uint32_t*>(base);
        areg[i][kt][1] = *reinterpret_cast<const uint32_t*>(base + (size_t)8 * kc);
        areg[i][kt][2] = *reinterpret_cast<const uint32_t*>(base + 8);
        areg[i][kt][3] = *reinterpret_cast<const uint32_t*>(base + (size_t)8 * kc + 8);
      }
    }
  }
  if (wp > 0) {
    const int cpr = (KT - RM_KREG) * 2;                     // 16-byte chunks per row kept in shared memory
    for (int i = tid; i < MT * 16 * cpr; i += RM_THREADS) {
      const int row = i / cpr, c = i - row * cpr;
      *reinterpret_cast<uint4*>(Wsm + (size_t)row * wp + c * 8) =
          *reinterpret_cast<const uint4*>(w_img + (size_t)row * kc + RM_KREG * 16 + c * 8);
    }
  }
  for (int i = tid; i < 2 * 8 * hp; i += RM_THREADS) hsm[i] = __float2bfloat16(0.f);
  if (tid < NSEQ) {
    const int b = b0 + tid;
    const int l = b < B ? (lens ? lens[b] : S) : 0;
    len_s[tid] = l < 0 ? 0 : (l > S ? S : l);
  }
  __syncthreads();
  int max_len = 0;
#pragma unroll
  for (int i = 0; i < NSEQ; ++i) max_len = max(max_len, len_s[i]);

  // gate role: pair p = tid + q * THREADS  ->  (sequence n = p / H, unit j = p % H); state lives in registers
  int pn[PPT], pj[PPT], plen[PPT];
  float hreg[PPT], creg[PPT], bh[PPT][G];
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int p = tid + q * RM_THREADS;
    const bool on = p < NSEQ * H;
    pn[q] = on ? p / H : 0;
    pj[q] = on ? p - pn[q] * H : 0;
    const int b = b0 + pn[q];
    plen[q] = (on && b < B) ? len_s[pn[q]] : 0;
    hreg[q] = (on && b < B && h0 != nullptr) ? h0[(int64_t)b * H + pj[q]] : 0.f;
    creg[q] = 0.f;
#pragma unroll
    for (int g = 0; g < G; ++g) bh[q][g] = (KIND == 1 && on) ? __ldg(b_hh + g * H + pj[q]) : 0.f;
    if (on && h0 != nullptr) {
      const __nv_bfloat16 hi = __float2bfloat16(hreg[q]);
      hsm[pn[q] * hp + pj[q]] = hi;
      hsm[8 * hp + pn[q] * hp + pj[q]] = __float2bfloat16(hreg[q] - __bfloat162float(hi));
    }
  }
  __syncthreads();

  const uint32_t wsm_u = (uint32_t)__cvta_generic_to_shared(Wsm);
  // ldmatrix row address of this lane inside a 16 x 16 tile: row (lane & 7) + 8 * ((lane >> 3) & 1), 16-byte chunk lane >> 4
  const uint32_t lm_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * wp + (lane >> 4) * 8) * 2u;

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
    // ---- (1) pre = W_hh . (h_hi + h_lo) ------------------------------------------------------------------------
    // NSEQ <= 4: the high and the low part of h sit side by side in the N = 8 columns of ONE MMA (columns [0, NSEQ) high,
    // [NSEQ, 2 NSEQ) low) -- half the MMAs of the two-pass form, and the MMA count is what bounds a step; NSEQ = 8 needs
    // all eight columns for the sequences and issues one MMA per part into two independent accumulator chains
    constexpr bool PACK = NSEQ <= 4;
    // (splitting the k-tiles over two accumulator chains did not change the step time: the phase is HMMA-throughput bound)
    float acc[RM_MAXMT][4], acl[RM_MAXMT][4];
#pragma unroll
    for (int i = 0; i < RM_MAXMT; ++i)
#pragma unroll
      for (int r = 0; r < 4; ++r) { acc[i][r] = 0.f; acl[i][r] = 0.f; }
    // B-fragment row of this lane's column gq: PACK -> high part of sequence gq, low part of sequence gq - NSEQ, or a zero row
    const __nv_bfloat16* hrow = !PACK ? hsm + gq * hp
                                      : (gq < NSEQ ? hsm + gq * hp : (gq < 2 * NSEQ ? hsm + 8 * hp + (gq - NSEQ) * hp : hsm + 7 * hp));
#pragma unroll
    for (int kt = 0; kt < RM_MAXKT; ++kt) {
      if (kt < KT) {
        const __nv_bfloat16* hb = hrow + kt * 16 + 2 * tq;
        const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(hb), bh1 = *reinterpret_cast<const uint32_t*>(hb + 8);
        uint32_t bl0 = 0, bl1 = 0;
        if (!PACK) {
          bl0 = *reinterpret_cast<const uint32_t*>(hb + 8 * hp);
          bl1 = *reinterpret_cast<const uint32_t*>(hb + 8 * hp + 8);
        }
#pragma unroll
        for (int i = 0; i < RM_MAXMT; ++i) {
          const int mt = warp + RM_WARPS * i;
          if (mt < MT) {
            if (kt < RM_KREG) {
              mma_bf16(acc[i], areg[i][kt < RM_KREG ? kt : 0], bh0, bh1);
              if (!PACK) mma_bf16(acl[i], areg[i][kt < RM_KREG ? kt : 0], bl0, bl1);
            } else {
              uint32_t a[4];
              ldmatrix_x4(a, wsm_u + (uint32_t)(mt * 16 * wp + (kt - RM_KREG) * 16) * 2u + lm_off);
              mma_bf16(acc[i], a, bh0, bh1);
              if (!PACK) mma_bf16(acl[i], a, bl0, bl1);
            }
          }
        }
      }
    }
    // accumulator (row gq / gq + 8, columns 2 tq, 2 tq + 1) -> pre[row][n] = high + low
#pragma unroll
    for (int i = 0; i < RM_MAXMT; ++i) {
      const int mt = warp + RM_WARPS * i;
      float v0, v1, v2, v3;
      v0 = acc[i][0] + acl[i][0]; v1 = acc[i][1] + acl[i][1];
      v2 = acc[i][2] + acl[i][2]; v3 = acc[i][3] + acl[i][3];
      if (PACK) {        // the low-part columns live NSEQ / 2 lanes further (same row group)
        v0 += __shfl_xor_sync(0xffffffffu, v0, NSEQ / 2);
        v1 += __shfl_xor_sync(0xffffffffu, v1, NSEQ / 2);
        v2 += __shfl_xor_sync(0xffffffffu, v2, NSEQ / 2);
        v3 += __shfl_xor_sync(0xffffffffu, v3, NSEQ / 2);
      }
      if (mt < MT && 2 * tq < NSEQ) {
        float* p0 = pre + (size_t)(mt * 16 + gq) * NSEQ + 2 * tq;
        *reinterpret_cast<float2*>(p0) = make_float2(v0, v1);
        *reinterpret_cast<float2*>(p0 + 8 * NSEQ) = make_float2(v2, v3);
      }
    }
    __syncthreads();
    // ---- (2) gates ----------------------------------------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < PPT; ++q) {
      if (s < plen[q]) {
        const int n = pn[q], j = pj[q], b = b0 + n;
        float pr[G];
#pragma unroll
        for (int g = 0; g < G; ++g) pr[g] = pre[(size_t)(g * H + j) * NSEQ + n] + bh[q][g];
        float* gs = gates + ((int64_t)b * S + s) * GH + j;
        const int64_t o = ((int64_t)b * S + s) * H + j;
        float h;
        if (KIND == 0) {
          const float gi = rm_sigm(xv[q][0] + pr[0]), gf = rm_sigm(xv[q][1] + pr[1]);
          const float gg = rm_tanh(xv[q][2] + pr[2]), go = rm_sigm(xv[q][3] + pr[3]);
          const float c = gf * creg[q] + gi * gg;
          h = go * rm_tanh(c);
          gs[0] = gi; gs[H] = gf; gs[2 * H] = gg; gs[3 * H] = go;
          creg[q] = c;
          cs[o] = c; hs[o] = h;
        } else {
          const float r = rm_sigm(xv[q][0] + pr[0]), z = rm_sigm(xv[q][1] + pr[1]);
          const float hn = pr[2];
          const float nn = rm_tanh(xv[q][2] + r * hn);
          h = (1.f - z) * nn + z * hreg[q];
          gs[0] = r; gs[H] = z; gs[2 * H] = nn;
          cs[o] = hn; hs[o] = h;
        }
        hreg[q] = h;
        const __nv_bfloat16 hi = __float2bfloat16(h);
        hsm[n * hp + j] = hi;
        hsm[8 * hp + n * hp + j] = __float2bfloat16(h - __bfloat162float(hi));
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int p = tid + q * RM_THREADS;
    if (p < NSEQ * H && b0 + pn[q] < B) user[(int64_t)(b0 + pn[q]) * H + pj[q]] = hreg[q];
  }
}

// sequences per CTA: the smallest of 2 / 4 / 8 that covers the batch with one wave of CTAs
static int rm_nseq(int B) {
  const int sms = sm_count();
  if ((B + 1) / 2 <= sms) return 2;
  if ((B + 3) / 4 <= sms) return 4;
  return 8;
}

template <int KIND>
static int rm_launch_fwd(int nseq, const float* xp, int ldx, const __nv_bfloat16* img, const float* b_hh, const float* h0,
                         const int32_t* lens, float* gates, float* hs, float* cs, float* user, int B, int S, int H, cudaStream_t st) {
  const RMGeom g = rm_geom(KIND == 0 ? MR_RNN_LSTM : MR_RNN_GRU, H, nseq);
  const unsigned grid = (unsigned)ceil_div(B, nseq);
#define RM_LAUNCH_F(NS)                                                                                              \
  {                                                                                                                  \
    cudaFuncSetAttribute(rnn_mma_fwd_kernel<KIND, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);     \
    launch_pdl(rnn_mma_fwd_kernel<KIND, NS>, dim3(grid), dim3(RM_THREADS), g.smem, st, xp, ldx, img, b_hh, h0, lens, gates, hs, cs, user, B, S, H, g.MT, g.KT); \
  }
  if (nseq == 2) RM_LAUNCH_F(2) else if (nseq == 4) RM_LAUNCH_F(4) else RM_LAUNCH_F(8)
#undef RM_LAUNCH_F
  MR_CHECK_LAUNCH("rnn_mma_fwd_kernel");
  return MR_OK;
}

int rnn_mma_fwd(int kind, const float* xp, int ldx, const float* w_hh_f32, const float* b_hh, const float* h0, const int32_t* lens,
                float* gates, float* hs, float* cs, float* user, int B, int S, int H, void* scratch, cudaStream_t st) {
  const int nseq = rm_nseq(B);
  const RMGeom g = rm_geom(kind, H, nseq);
  __nv_bfloat16* img = static_cast<__nv_bfloat16*>(scratch);
  const int rows = g.MT * 16, cols = g.KT * 16;
  launch_pdl(rnn_mma_prep_kernel, dim3((unsigned)ceil_div((int64_t)rows * cols, 256)), dim3(256), 0, st, w_hh_f32, img, g.GH, H, rows, cols);
  MR_CHECK_LAUNCH("rnn_mma_prep_kernel");
  return kind == MR_RNN_LSTM ? rm_launch_fwd<0>(nseq, xp, ldx, img, b_hh, h0, lens, gates, hs, cs, user, B, S, H, st)
                             : rm_launch_fwd<1>(nseq, xp, ldx, img, b_hh, h0, lens, gates, hs, cs, user, B, S, H, st);
}

}  // namespace mr
