// Host side of the tap GEMM (tapgemm.cuh): shared-memory carve-up, weight packing, launch.
#include "tapgemm.cuh"
#include <stdlib.h>
#include <string.h>

namespace mr {

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// optional in-kernel wait accounting (args.dbg != nullptr): per CTA, cycles each role spent blocked
#define TG_TIMED(slot, stmt)                                   \
  do {                                                         \
    if (p.dbg != nullptr) {                                    \
      const long long t0__ = clock64();                        \
      stmt;                                                    \
      dbg_acc[slot] += clock64() - t0__;                       \
    } else {                                                   \
      stmt;                                                    \
    }                                                          \
  } while (0)

// ring position + phase bit, advanced without integer division (the role loops are latency bound:
// every instruction on their critical path shows up in the tile rate)
struct Ring {
  uint32_t pos, phase, n;
  __device__ __forceinline__ explicit Ring(uint32_t n_) : pos(0), phase(0), n(n_) {}
  __device__ __forceinline__ void next() {
    if (++pos == n) { pos = 0; phase ^= 1u; }
  }
};

// POOL = true: the instantiation whose epilogue is TG_EPI_RELUGRAD_POOL (kept apart so that neither epilogue pays for the
// other's registers and instructions).
// EPI2 = true (dense A staged by TMA, N <= 256): 12 warps, warps 8-11 are a SECOND epilogue group.  Group w owns
// accumulator stage w, i.e. every other tile of the CTA: an epilogue is a long dependent chain per thread (TMEM load,
// convert, activation, pack, store), so one warp per scheduler leaves most issue slots empty; two tiles in flight fill them.
template <bool POOL, bool EPI2>
__global__ void __launch_bounds__(EPI2 ? TG_THREADS2 : TG_THREADS, 1) tapgemm_kernel(const TapGemmArgs p, const __grid_constant__ CUtensorMap tmap) {
  constexpr int NT = EPI2 ? TG_THREADS2 : TG_THREADS;
  long long dbg_acc[4] = {0, 0, 0, 0};
  const long long dbg_t0 = clock64();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                      // A slots: [rows_alloc x 128 B], 1024-byte aligned (swizzle period)
  uint8_t* sB = sA + (size_t)p.ns_a * p.a_slot_bytes;
  float* sbias = reinterpret_cast<float*>(sB + (size_t)p.ns_b * p.b_slot_bytes);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sbias + 512);
  uint64_t* a_empty = a_full + TG_MAX_SLOTS;
  uint64_t* b_full = a_empty + TG_MAX_SLOTS;
  uint64_t* b_empty = b_full + TG_MAX_SLOTS;
  uint64_t* t_full = b_empty + TG_MAX_SLOTS;
  uint64_t* t_empty = t_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  uint8_t* hot = reinterpret_cast<uint8_t*>(tmem_slot + 4);          // [n_hot][hot_pitch] bf16 rows (gather mode)
  const uint32_t hot_pitch = (uint32_t)((p.K * 2 + 15) / 16 * 16);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_total = p.nsz[0] + (p.n_sub > 1 ? p.nsz[1] : 0);
  const uint32_t acc_stages = n_total <= 256 ? 2u : 1u;
  const int n_chunks = (p.K + TG_KC - 1) / TG_KC;
  const int last_kc = p.K - (n_chunks - 1) * TG_KC;
  // Every CTA streams the same weight blocks; in lockstep all 148 SMs would hit the same L2 lines at the
  // same moment.  So each CTA walks the (k-chunk, tap) blocks in its own rotated order and reads its own
  // replica of the packed weights.
  // (rotation changes the fp32 accumulation order per CTA, i.e. a title's result would depend on which CTA
  //  encodes it; it is off by default so that the encoder is batch-invariant -- p.rotate enables it)
  const int rot_c = p.rotate ? (int)(blockIdx.x % (unsigned)n_chunks) : 0;
  const int rot_t = p.rotate ? (int)((blockIdx.x / (unsigned)n_chunks) % (unsigned)p.taps) : 0;
  const uint8_t* wrep = p.wpack + (size_t)(blockIdx.x % (unsigned)p.w_reps) * p.w_rep_stride;

  // ---- one-time setup ----------------------------------------------------------------------
  pdl_trigger();
  {
    const uint32_t a_bytes = (uint32_t)p.ns_a * p.a_slot_bytes;
    for (uint32_t i = tid * 16; i < a_bytes; i += NT * 16) *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
      for (int i = 0; i < TG_MAX_SLOTS; ++i) {
        tc::mbar_init(&a_full[i], p.use_tma ? 1 : 128);
        tc::mbar_init(&a_empty[i], 1);
        tc::mbar_init(&b_full[i], 1);
        tc::mbar_init(&b_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&t_full[i], 1);
        tc::mbar_init(&t_empty[i], 128);
      }
      tc::fence_barrier_init();
    }
    if (warp == 4) tc::tmem_alloc(tmem_slot, 512);
    pdl_wait();                        // everything above touched this CTA's shared memory / TMEM only
    for (int i = tid; i < 512; i += NT)
      sbias[i] = (p.bias != nullptr && i < p.n_valid) ? p.bias[i] + (p.bias2 != nullptr ? p.bias2[i] : 0.f) : 0.f;
    if (p.ids != nullptr)
      for (uint32_t i = tid; i < (uint32_t)p.n_hot * (hot_pitch / 16); i += NT) {
        const uint32_t h = i / (hot_pitch / 16), q = i % (hot_pitch / 16);
        reinterpret_cast<uint4*>(hot + (size_t)h * hot_pitch)[q] = __ldg(reinterpret_cast<const uint4*>(p.a + p.hot_ids[h] * p.lda) + q);
      }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }
  const uint32_t tmem = *tmem_slot;

  if (warp < 4 || (EPI2 && warp >= 8)) {
    // =================================== epilogue ===========================================
    const int wg = EPI2 ? (warp >> 3) : 0;                   // epilogue group (EPI2: group w <-> accumulator stage w)
    const int r = tid & 127;
    const int g = r % p.G, l = r / p.G;
    const bool row_ok = l < p.L;
    const int64_t row_off = (int64_t)g * p.L + l;            // token offset inside the tile's G titles
    Ring acc(acc_stages);
    if (EPI2) acc.pos = (uint32_t)wg;
    auto acc_advance = [&]() {
      if (EPI2) acc.phase ^= 1u;                              // same stage every time, next use
      else acc.next();
    };
    const int64_t tile_step = (int64_t)gridDim.x * (EPI2 ? 2 : 1);
    float colacc[8];                      // column sums of the stored rows: lane j, slot ci <-> column 32 ci + j
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) colacc[ci] = 0.f;
    const int lane = tid & 31;
    float* sdn = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(hot) + 15) & ~(uintptr_t)15);     // POOL: [2][G][dn_pitch]
    const int dn_pitch = n_total + 4;        // +16 B per row: the G rows read by one LDS.128 fall into different banks
    for (int64_t tile = blockIdx.x + (int64_t)wg * gridDim.x; tile < p.n_tiles; tile += tile_step) {
      const int64_t t = tile * p.G * p.L + row_off;
      const bool valid = row_ok && (tile * p.G + g < p.n_titles) && t < p.n_rows;
      if constexpr (POOL) {
        // ---- dconv = relu'(c) * (acc + p[t] d_news[title] (+ e0)) ; the c row (<= 160 columns) and p[t] are requested
        // BEFORE the accumulator is waited for, so one memory latency per tile is exposed at most ----------------
        uint32_t cm[5] = {0u, 0u, 0u, 0u, 0u};       // bit j of cm[ci]: c[t, 32 ci + j] > 0 (sign mask written by the pooling backward)
        float pt = 0.f;
        if (valid) {
          const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(p.cmask + t * 32));
          cm[4] = __ldg(reinterpret_cast<const uint32_t*>(p.cmask + t * 32 + 16));
          cm[0] = m0.x; cm[1] = m0.y; cm[2] = m0.z; cm[3] = m0.w;
          pt = __ldg(p.prob + t);
        }
        // the tile's G rows of d_news go through shared memory (two stages, one per accumulator stage)
        float* sd_stage = sdn + (size_t)acc.pos * p.G * dn_pitch;
        {
          const int pieces = n_total >> 2;
          for (int i = r; i < p.G * pieces; i += 128) {
            const int gg = i / pieces, q = i - gg * pieces;
            const int64_t title = tile * p.G + gg;
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (title < p.n_titles) v4 = __ldg(reinterpret_cast<const float4*>(p.dnp + title * p.ldn) + q);
            *reinterpret_cast<float4*>(sd_stage + gg * dn_pitch + 4 * q) = v4;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
        }
        const float4* dn_row = reinterpret_cast<const float4*>(sd_stage + g * dn_pitch);
        TG_TIMED(0, tc::mbar_wait(&t_full[acc.pos], acc.phase));
        tc::tc_fence_after();
        const uint32_t tb = tmem + ((uint32_t)((warp & 3) * 32) << 16) + acc.pos * 256u;
        // NOT unrolled: five copies of this body (~2k instructions) overflow the instruction cache, and the epilogue warps
        // then stall on instruction fetch; the per-chunk mask / column-sum registers are picked with selects instead
#pragma unroll 1
        for (int ci = 0; ci < 5; ++ci) {
          if (ci * 32 < n_total) {
            const int n0 = ci * 32;
            const bool wide = n_total - n0 >= 32;
            const uint32_t cmask_c = ci == 0 ? cm[0] : (ci == 1 ? cm[1] : (ci == 2 ? cm[2] : (ci == 3 ? cm[3] : cm[4])));
            uint32_t v[32];
            if (wide) {
              tc::tmem_ld32(tb + n0, v);
            } else {
              uint32_t h[16];
              tc::tmem_ld16(tb + n0, h);
#pragma unroll
              for (int j = 0; j < 16; ++j) { v[j] = h[j]; v[16 + j] = 0u; }
            }
            float4 dn[8];
            uint4 x0[4];
            if (valid) {
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (u < 4 || wide) dn[u] = dn_row[ci * 8 + u];
              if (p.e0 != nullptr) {
                const uint4* q0 = reinterpret_cast<const uint4*>(p.e0 + t * p.lde + n0);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (u < 2 || wide) x0[u] = __ldg(q0 + u);
              }
            }
            tc::tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const bool on = valid && (u < 2 || wide);
              const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&x0[u]);
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                const int j = u * 8 + 2 * w;
                float a0 = 0.f, a1 = 0.f;
                if (on) {
                  const float4 d4 = dn[j >> 2];
                  const float d0 = (j & 2) ? d4.z : d4.x, d1 = (j & 2) ? d4.w : d4.y;
                  a0 = fmaf(pt, d0, __uint_as_float(v[j]));
                  a1 = fmaf(pt, d1, __uint_as_float(v[j + 1]));
                  if (p.e0 != nullptr) {
                    const float2 e2 = __bfloat1622float2(h0[w]);
                    a0 += e2.x; a1 += e2.y;
                  }
                  a0 = ((cmask_c >> j) & 1u) ? a0 : 0.f;
                  a1 = ((cmask_c >> (j + 1)) & 1u) ? a1 : 0.f;
                }
                f[j] = a0; f[j + 1] = a1;
              }
            }
            if (valid) {
              uint4* dst = reinterpret_cast<uint4*>(p.out + t * p.ldo + n0);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (u < 2 || wide) {
                  uint4 o;
                  o.x = tc::pack_bf16(f[u * 8 + 0], f[u * 8 + 1]);
                  o.y = tc::pack_bf16(f[u * 8 + 2], f[u * 8 + 3]);
                  o.z = tc::pack_bf16(f[u * 8 + 4], f[u * 8 + 5]);
                  o.w = tc::pack_bf16(f[u * 8 + 6], f[u * 8 + 7]);
                  dst[u] = o;
                }
              }
            }
            if (p.colsum_out != nullptr) {
              const float csum = tc::warp_colsum32(f, lane);
#pragma unroll
              for (int cj = 0; cj < 5; ++cj)
                if (cj == ci) colacc[cj] += csum;
            }
          }
        }
        tc::tc_fence_before();
        tc::mbar_arrive(&t_empty[acc.pos]);
        acc_advance();
        continue;
      }
      if constexpr (!POOL) {
      TG_TIMED(0, tc::mbar_wait(&t_full[acc.pos], acc.phase));
      tc::tc_fence_after();
      const uint32_t tbase = tmem + ((uint32_t)((warp & 3) * 32) << 16) + acc.pos * 256u;
      int col0 = 0;
      for (int sub = 0; sub < p.n_sub; ++sub) {
        const int nsub = sub == 0 ? p.nsz[0] : p.nsz[1];
        for (int c0 = 0; c0 < nsub; c0 += 32) {
          uint32_t v[32];
          const int n0 = col0 + c0;
          const bool wide = nsub - c0 >= 32;
          if (wide) {
            tc::tmem_ld32(tbase + n0, v);
          } else {
            uint32_t h[16];
            tc::tmem_ld16(tbase + n0, h);
#pragma unroll
            for (int j = 0; j < 16; ++j) { v[j] = h[j]; v[16 + j] = 0u; }
          }
          // operands of the RELUGRAD epilogue are fetched while the TMEM load is in flight
          uint4 x0[4], x1[4];
          if (p.epi == TG_EPI_RELUGRAD && valid) {
            const uint4* q0 = reinterpret_cast<const uint4*>(p.e0 + t * p.lde + n0);
            const uint4* q1 = reinterpret_cast<const uint4*>(p.e1 + t * p.lde + n0);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (u < 2 || wide) { x0[u] = __ldg(q0 + u); x1[u] = __ldg(q1 + u); }
          }
          tc::tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = valid ? __uint_as_float(v[j]) : 0.f;
          if (valid) {
            if (p.epi == TG_EPI_BIAS_RELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j] + sbias[n0 + j], 0.f);
            } else if (p.epi == TG_EPI_BIAS_TANH) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = tanh_fast(f[j] + sbias[n0 + j]);
            } else if (p.epi == TG_EPI_RELUGRAD) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (u < 2 || wide) {
                  const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&x0[u]);
                  const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&x1[u]);
#pragma unroll
                  for (int w = 0; w < 4; ++w) {
                    float2 d = __bfloat1622float2(h0[w]), c = __bfloat1622float2(h1[w]);
                    f[u * 8 + 2 * w] = c.x > 0.f ? f[u * 8 + 2 * w] + d.x : 0.f;
                    f[u * 8 + 2 * w + 1] = c.y > 0.f ? f[u * 8 + 2 * w + 1] + d.y : 0.f;
                  }
                }
              }
            }
            if (p.epi == TG_EPI_BIAS_F32) {
              float4* dst = reinterpret_cast<float4*>(p.out_f32 + t * p.ldo + n0);
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if ((u < 4 || wide) && n0 + 4 * u < p.n_store)
                  dst[u] = make_float4(f[4 * u] + sbias[n0 + 4 * u], f[4 * u + 1] + sbias[n0 + 4 * u + 1],
                                       f[4 * u + 2] + sbias[n0 + 4 * u + 2], f[4 * u + 3] + sbias[n0 + 4 * u + 3]);
              continue;
            }
            uint4* dst = reinterpret_cast<uint4*>(p.out + t * p.ldo + n0);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (u < 2 || wide) {
                uint4 o;
                o.x = tc::pack_bf16(f[u * 8 + 0], f[u * 8 + 1]);
                o.y = tc::pack_bf16(f[u * 8 + 2], f[u * 8 + 3]);
                o.z = tc::pack_bf16(f[u * 8 + 4], f[u * 8 + 5]);
                o.w = tc::pack_bf16(f[u * 8 + 6], f[u * 8 + 7]);
                dst[u] = o;
              }
            }
          }
          if (p.colsum_out != nullptr && p.epi == TG_EPI_RELUGRAD) {
            // column sums of what this warp stored (invalid rows count as zero); n_sub == 1, n_total <= 256
            const float cs = tc::warp_colsum32(f, lane);
#pragma unroll
            for (int ci = 0; ci < 8; ++ci)
              if (ci == (n0 >> 5)) colacc[ci] += cs;
          }
        }
        col0 += nsub;
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&t_empty[acc.pos]);
      acc_advance();
      }
    }
    if (p.colsum_out != nullptr) {
      float* cs = p.colsum_out + ((size_t)blockIdx.x * (EPI2 ? 8 : 4) + wg * 4 + (warp & 3)) * n_total;
#pragma unroll
      for (int ci = 0; ci < 8; ++ci)
        if (ci * 32 + lane < n_total) cs[ci * 32 + lane] = colacc[ci];
    }
  } else if (warp == 4) {
    // =================================== MMA issuer ==========================================
    // The whole warp walks the loop (warp-uniform control flow, so descriptor arithmetic stays in uniform
    // registers); one elected lane issues tcgen05.mma / tcgen05.commit.  Descriptors are built once per
    // block and advanced by adding to their low word (start address field, 16-byte units).
    const uint32_t idesc0 = tc::make_idesc(128, p.nsz[0], 0, 0);
    const uint32_t idesc1 = tc::make_idesc(128, p.n_sub > 1 ? p.nsz[1] : 16, 0, 0);
    const uint32_t n0 = (uint32_t)p.nsz[0];
    const bool two = p.n_sub > 1;
    const uint32_t b_ps = (uint32_t)n_total * 16;
    const uint64_t a_hi = tc::make_desc_sw(0, 16, 1024, 2, 0) & 0xFFFFFFFF00000000ull;
    const uint64_t b_desc0 = tc::make_desc(0, b_ps, 128);
    const uint32_t b_kstep = (2u * b_ps) >> 4;             // k-step of 16 elements = two panels
    const uint32_t sA0 = tc::smem_u32(sA), sB0 = tc::smem_u32(sB);
    // operand row offset (in 16-byte units of the start-address field) of each tap: (halo + shift) * 128 B
    const int ctr = (p.taps - 1) / 2;
    const uint32_t tap_off0 = (uint32_t)(p.halo + p.dir * (0 - ctr) * p.G) * 8u;
    const uint32_t tap_off1 = (uint32_t)(p.halo + p.dir * (1 - ctr) * p.G) * 8u;
    const uint32_t tap_off2 = (uint32_t)(p.halo + p.dir * (2 - ctr) * p.G) * 8u;
    Ring ra(p.ns_a), rb(p.ns_b), acc(acc_stages);
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      TG_TIMED(0, tc::mbar_wait(&t_empty[acc.pos], acc.phase ^ 1u));
      tc::tc_fence_after();
      const uint32_t dcol = tmem + acc.pos * 256u;
      uint32_t accum = 0;
      int c = rot_c;
      for (int cc = 0; cc < n_chunks; ++cc) {
        const int nks = (c == n_chunks - 1 ? last_kc : TG_KC) >> 4;
        TG_TIMED(1, tc::mbar_wait(&a_full[ra.pos], ra.phase));
        const uint32_t a_slot16 = (sA0 + ra.pos * p.a_slot_bytes) >> 4;
        int tap = rot_t;
        for (int tt = 0; tt < p.taps; ++tt) {
          TG_TIMED(2, tc::mbar_wait(&b_full[rb.pos], rb.phase));
          tc::tc_fence_after();
          const uint32_t b_slot = sB0 + rb.pos * p.b_slot_bytes;
          const uint32_t toff = tap == 0 ? tap_off0 : (tap == 1 ? tap_off1 : tap_off2);
          const uint64_t da0 = a_hi | (uint64_t)((a_slot16 + toff) & 0x3FFFu);
          const uint64_t db0 = b_desc0 | (uint64_t)((b_slot >> 4) & 0x3FFFu);
          const uint64_t db1 = b_desc0 | (uint64_t)(((b_slot + n0 * 16u) >> 4) & 0x3FFFu);
          const bool last_tap = tt == p.taps - 1;
          TG_TIMED(3, {
          if (tc::elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < nks) {
                tc::umma(dcol, da0 + (uint64_t)(2 * ks), db0 + (uint64_t)(ks * b_kstep), idesc0, accum | (uint32_t)ks);
                if (two) tc::umma(dcol + n0, da0 + (uint64_t)(2 * ks), db1 + (uint64_t)(ks * b_kstep), idesc1, accum | (uint32_t)ks);
              }
            }
            tc::umma_commit(&b_empty[rb.pos]);
            if (last_tap) tc::umma_commit(&a_empty[ra.pos]);
            if (last_tap && cc == n_chunks - 1) tc::umma_commit(&t_full[acc.pos]);
          }
          __syncwarp();
          });
          accum = 1;
          rb.next();
          if (++tap == p.taps) tap = 0;
        }
        ra.next();
        if (++c == n_chunks) c = 0;
      }
      acc.next();
    }
  } else if (warp == 5) {
    // =================================== weight producer ====================================
    if ((tid & 31) == 0) {
      const uint32_t sB0 = tc::smem_u32(sB);
      const uint32_t full_bytes = (uint32_t)(TG_KC / 8) * (uint32_t)n_total * 16u;
      const uint32_t last_bytes = (uint32_t)(last_kc / 8) * (uint32_t)n_total * 16u;
      Ring rb(p.ns_b);
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        int c = rot_c;
        for (int cc = 0; cc < n_chunks; ++cc) {
          const uint32_t bytes = c == n_chunks - 1 ? last_bytes : full_bytes;
          int tap = rot_t;
          for (int tt = 0; tt < p.taps; ++tt) {
            TG_TIMED(0, tc::mbar_wait(&b_empty[rb.pos], rb.phase ^ 1u));
            tc::mbar_arrive_expect_tx(&b_full[rb.pos], bytes);
            tc::bulk_g2s(sB0 + rb.pos * p.b_slot_bytes, wrep + (size_t)(c * p.taps + tap) * p.b_slot_bytes, bytes, &b_full[rb.pos]);
            rb.next();
            if (++tap == p.taps) tap = 0;
          }
          if (++c == n_chunks) c = 0;
        }
      }
    }
  } else if (p.use_tma) {
    // =================================== A producer (TMA) ===================================
    // One warp.  Dense activations: ONE 3-D tile load per k-chunk -- the tensor map views the [T, ld] matrix as
    // (column, title, position), so the box {64, G, L} arrives in shared memory already in the position-major
    // row order r = l*G + g, 128-byte swizzled by the copy engine.  Token table: lane i gathers rows 4i..4i+3
    // of the tile with one tile::gather4 (4 token ids -> 4 x 128 B); rows outside the titles use row index V,
    // which is out of bounds and therefore zero filled.
    if (warp == 6) {
      const int lane = tid & 31;
      if (lane == 0) tc::tma_prefetch_desc(&tmap);
      const uint32_t sA0 = tc::smem_u32(sA) + (uint32_t)p.halo * 128u;
      const uint32_t dense_bytes = (uint32_t)(p.G * p.L) * 128u;
      int64_t row_off[4];
      int row_g[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = 4 * lane + k;
        const int g = r % p.G, l = r / p.G;
        row_g[k] = g;
        row_off[k] = l < p.L ? (int64_t)g * p.L + l : -1;
      }
      auto row_id = [&](int64_t tile, int k) -> int {
        if (p.ids == nullptr || row_off[k] < 0 || tile >= p.n_tiles || tile * p.G + row_g[k] >= p.n_titles ||
            tile * p.G * p.L + row_off[k] >= p.n_rows)
          return (int)(p.V + (int64_t)p.n_hot * p.hot_reps);      // out of bounds -> zero filled
        int64_t id = load_index(p.ids, p.ids_i64, tile * p.G * p.L + row_off[k]);
        id = id < 0 ? 0 : (id >= p.V ? p.V - 1 : id);
        if (p.hot_reps > 0) {          // hot rows: read this CTA's replica instead of the one row every SM wants
#pragma unroll
          for (int h = 0; h < TG_MAX_HOT; ++h)
            if (h < p.n_hot && id == p.hot_ids[h]) id = p.V + (int64_t)h * p.hot_reps + (int64_t)(blockIdx.x % (unsigned)p.hot_reps);
        }
        return (int)id;
      };
      int nxt[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) nxt[k] = row_id(blockIdx.x, k);
      Ring ra(p.ns_a);
      for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        int rid[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) rid[k] = nxt[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) nxt[k] = row_id(tile + gridDim.x, k);      // ids of the next tile, in flight
        int c = rot_c;
        for (int cc = 0; cc < n_chunks; ++cc) {
          TG_TIMED(0, tc::mbar_wait(&a_empty[ra.pos], ra.phase ^ 1u));
          const uint32_t slot = sA0 + ra.pos * p.a_slot_bytes;
          if (p.ids != nullptr) {
            if (lane == 0) tc::mbar_arrive_expect_tx(&a_full[ra.pos], 128u * 128u);
            __syncwarp();
            tc::tma_gather4(slot + (uint32_t)lane * 512u, &tmap, c * TG_KC, rid[0], rid[1], rid[2], rid[3], &a_full[ra.pos]);
          } else if (lane == 0) {
            tc::mbar_arrive_expect_tx(&a_full[ra.pos], dense_bytes);
            tc::tma_load_3d(slot, &tmap, c * TG_KC, (int)(tile * p.G), 0, &a_full[ra.pos]);
          }
          ra.next();
          if (++c == n_chunks) c = 0;
        }
      }
    }
  } else if constexpr (!EPI2) {
    // =================================== A producers (cp.async) =============================
    const int ptid = tid - 192;            // 0..127
    const int rgrp = ptid >> 3, j = ptid & 7;
    const uint32_t depth = (uint32_t)(p.ns_a - 2 < 3 ? p.ns_a - 2 : 3);   // cp.async groups in flight behind the signalled one
    const uint32_t sA0 = tc::smem_u32(sA);
    // this thread stages the 16-byte piece j of rows r = rgrp + 16*s (s = 0..7) of every k-chunk.  Row r is
    // token (title = tile*G + r%G, position = r/G); the geometry is tile independent, so it is computed once.
    int64_t row_off[8];      // token offset inside the tile's G titles, -1 = permanent zero row
    int row_g[8];
    uint32_t dst_off[8];     // row-linear SWIZZLE_128B: row at (halo + r) * 128, piece j at ((j ^ row) & 7) * 16
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int r = rgrp + 16 * s;
      const int g = r % p.G, l = r / p.G;
      row_g[s] = g;
      row_off[s] = l < p.L ? (int64_t)g * p.L + l : -1;
      const uint32_t row = (uint32_t)(p.halo + r);
      dst_off[s] = row * 128u + ((((uint32_t)j ^ row) & 7u) << 4);
    }
    // token index of row s of a tile (-1 = zero row); the id load is issued one tile ahead and classified at the
    // top of the next iteration, so its latency overlaps the staging of the current tile
    auto row_token = [&](int64_t tile, int s) -> int64_t {
      if (row_off[s] < 0 || tile >= p.n_tiles || tile * p.G + row_g[s] >= p.n_titles) return -1;
      const int64_t t_ = tile * p.G * p.L + row_off[s];
      return t_ < p.n_rows ? t_ : -1;                 // rows past the problem (flat row lists padded to whole tiles): zero rows, no id read
    };
    auto raw_of = [&](int64_t t) -> int64_t { return (t < 0 || p.ids == nullptr) ? t : load_index(p.ids, p.ids_i64, t); };
    // -> source row (token id in gather mode, token index otherwise); -1 = zero row; -2-h = hot row h (smem copy)
    auto classify = [&](int64_t t, int64_t raw) -> int64_t {
      if (t < 0 || p.ids == nullptr) return t;
      int64_t id = raw < 0 ? 0 : (raw >= p.V ? p.V - 1 : raw);
#pragma unroll
      for (int h = 0; h < TG_MAX_HOT; ++h)
        if (h < p.n_hot && id == p.hot_ids[h]) id = -2 - h;
      return id;
    };
    int64_t raw[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) raw[s] = raw_of(row_token(blockIdx.x, s));
    Ring ra(p.ns_a), sig(p.ns_a);
    uint32_t pending = 0;                  // committed but not yet signalled k-chunks
    for (int64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      int64_t nxt[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) nxt[s] = classify(row_token(tile, s), raw[s]);
#pragma unroll
      for (int s = 0; s < 8; ++s) raw[s] = raw_of(row_token(tile + gridDim.x, s));      // in flight while this tile is staged
      const __nv_bfloat16* rowp[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        if (nxt[s] >= 0) rowp[s] = p.a + nxt[s] * p.lda + j * 8;
        else if (nxt[s] == -1) rowp[s] = nullptr;
        else rowp[s] = reinterpret_cast<const __nv_bfloat16*>(hot + (size_t)(-2 - nxt[s]) * hot_pitch) + j * 8;   // shared memory
      }
      uint32_t hot_mask = 0;
#pragma unroll
      for (int s = 0; s < 8; ++s) hot_mask |= (nxt[s] <= -2 ? 1u : 0u) << s;
      int c = rot_c;
      for (int cc = 0; cc < n_chunks; ++cc) {
        const int kc = c == n_chunks - 1 ? last_kc : TG_KC;
        TG_TIMED(0, tc::mbar_wait(&a_empty[ra.pos], ra.phase ^ 1u));
        TG_TIMED(3, {
          if (j * 8 < kc) {
            const uint32_t slot = sA0 + ra.pos * p.a_slot_bytes;
            const int col = c * TG_KC;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
              const __nv_bfloat16* q = rowp[s];
              if ((hot_mask >> s) & 1u) {          // hot token: 16-byte smem -> smem copy (covered by the fence below)
                const uint4 v = *reinterpret_cast<const uint4*>(q + col);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot + dst_off[s]), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
              } else {
                tc::cp_async16(slot + dst_off[s], q != nullptr ? (const void*)(q + col) : (const void*)p.a, q != nullptr ? 16u : 0u);
              }
            }
          }
          tc::cp_async_commit();
        });
        ra.next();
        if (++c == n_chunks) c = 0;
        if (++pending > depth) {
          TG_TIMED(1, {
            if (depth == 3) tc::cp_async_wait<3>();
            else if (depth == 2) tc::cp_async_wait<2>();
            else tc::cp_async_wait<1>();
            tc::fence_proxy_async();
          });
          tc::mbar_arrive(&a_full[sig.pos]);
          sig.next();
          --pending;
        }
      }
    }
    tc::cp_async_wait<0>();
    tc::fence_proxy_async();
    while (pending > 0) {
      tc::mbar_arrive(&a_full[sig.pos]);
      sig.next();
      --pending;
    }
  }

  if (p.dbg != nullptr && (tid == 0 || tid == 128 || tid == 160 || tid == 192)) {
    // rows: 0 epilogue {t_full}, 1 mma {t_empty, a_full, b_full}, 2 w-producer {b_empty}, 3 a-producer {a_empty, cp.wait, ids, cp.issue}
    long long* d = p.dbg + ((size_t)blockIdx.x * 4 + (tid == 0 ? 0 : tid == 128 ? 1 : tid == 160 ? 2 : 3)) * 5;
    d[0] = dbg_acc[0]; d[1] = dbg_acc[1]; d[2] = dbg_acc[2]; d[3] = dbg_acc[3]; d[4] = clock64() - dbg_t0;
  }
  // ---- teardown ----------------------------------------------------------------------------
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

// Packs fp32 weights into the [chunk][tap] panel blocks the kernel streams:
//   W[tap][n, k] = src[n*sn + k*sk + tap*st]  for n < n_valid, k < k_valid, else 0
// k_mod > 0: the K index is a concatenation of blocks of k_mod columns, k = b*k_mod + kk, read from
//   src[n*sn + kk*sk + b*sb + tap*st]  for kk < k_valid  (e.g. the three conv taps side by side, each padded to k_mod)
__global__ void tapgemm_pack_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int taps, int n_total, int K,
                                    int n_valid, int k_valid, int64_t sn, int64_t sk, int64_t st, uint32_t slot_bytes,
                                    int reps, int64_t rep_stride, int k_mod, int64_t sb) {
  pdl_trigger();
  pdl_wait();
  const int n_chunks = (K + TG_KC - 1) / TG_KC;
  const int64_t total = (int64_t)n_chunks * taps * (TG_KC / 8) * n_total;      // 16-byte units
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(u % n_total);
    int64_t rest = u / n_total;
    const int panel = (int)(rest % (TG_KC / 8));
    rest /= (TG_KC / 8);
    const int tap = (int)(rest % taps);
    const int c = (int)(rest / taps);
    const int k0 = c * TG_KC + panel * 8;
    if (k0 >= K) continue;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float lo = 0.f, hi = 0.f;
      const int k = k0 + 2 * e;
      if (k_mod > 0) {
        const int b0 = k / k_mod, kk0 = k - b0 * k_mod, b1 = (k + 1) / k_mod, kk1 = k + 1 - b1 * k_mod;
        if (n < n_valid && kk0 < k_valid) lo = src[n * sn + kk0 * sk + b0 * sb + tap * st];
        if (n < n_valid && kk1 < k_valid) hi = src[n * sn + kk1 * sk + b1 * sb + tap * st];
      } else {
        if (n < n_valid && k < k_valid) lo = src[n * sn + k * sk + tap * st];
        if (n < n_valid && k + 1 < k_valid) hi = src[n * sn + (k + 1) * sk + tap * st];
      }
      w[e] = tc::pack_bf16(lo, hi);
    }
    uint8_t* o = dst + (size_t)(c * taps + tap) * slot_bytes + (size_t)panel * n_total * 16 + (size_t)n * 16;
    for (int r = 0; r < reps; ++r) *reinterpret_cast<uint4*>(o + (size_t)r * rep_stride) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

long long* g_tapgemm_dbg = nullptr;

bool use_tma_default() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MINDREC_TMA");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

bool use_tma_gather() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MINDREC_TMA_GATHER");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 148;
}

constexpr size_t TG_SMEM_MAX = 227 * 1024;
constexpr size_t TG_SMEM_FIXED = 512 * 4 + (4 * TG_MAX_SLOTS + 4) * 8 + 16;   // bias + barriers + tmem slot

static int64_t g_hot_ids[TG_MAX_HOT] = {0, 0, 0, 0};
static int g_n_hot = 0;
static int g_hot_reps = 0;
int hot_replicas() { return g_hot_reps; }
int hot_tokens(int64_t* out) {
  for (int i = 0; i < g_n_hot; ++i) out[i] = g_hot_ids[i];
  return g_n_hot;
}

int tapgemm_plan(TapGemmArgs& a, TapGemmPlan* plan) {
  MR_REQUIRE(a.L >= 1 && a.L <= 128, MR_ERR_UNSUPPORTED, "tap gemm: signal_length %d not in [1,128] (bf16 path)", a.L);
  MR_REQUIRE(a.K >= 16 && a.K % 16 == 0, MR_ERR_BAD_SHAPE, "tap gemm: K=%d must be a positive multiple of 16", a.K);
  MR_REQUIRE(a.taps == 1 || a.taps == 3, MR_ERR_BAD_SHAPE, "tap gemm: taps=%d", a.taps);
  MR_REQUIRE(a.n_sub == 1 || a.n_sub == 2, MR_ERR_BAD_SHAPE, "tap gemm: n_sub=%d", a.n_sub);
  int n_total = 0;
  for (int s = 0; s < a.n_sub; ++s) {
    MR_REQUIRE(a.nsz[s] >= 16 && a.nsz[s] <= 256 && a.nsz[s] % 16 == 0, MR_ERR_BAD_SHAPE, "tap gemm: N sub-tile %d", a.nsz[s]);
    n_total += a.nsz[s];
  }
  MR_REQUIRE(n_total <= 512, MR_ERR_UNSUPPORTED, "tap gemm: N=%d > 512 TMEM columns", n_total);
  int G = 128 / a.L;
  if (G > 16) G = 16;
  a.G = G;
  // dense activations: TMA tile loads.  Token-table gather: cp.async + the hot-row shared-memory cache by default
  // (PAD/[CLS]/[SEP] rows hammer a few L2 lines when fetched by every SM; MINDREC_TMA_GATHER=1 selects gather4).
  a.hot_reps = a.ids != nullptr ? hot_replicas() : 0;
  a.use_tma = use_tma_default() && (a.ids == nullptr || use_tma_gather() || a.hot_reps > 0) ? 1 : 0;
  if (a.epi == TG_EPI_RELUGRAD_POOL) a.use_tma = 1;          // this epilogue only exists in the two-group (TMA) instantiation
  plan->epi2 = a.use_tma && a.ids == nullptr && n_total <= 256;
  plan->colsum_rows = 0;
  a.halo = a.taps > 1 ? (int)align_up(G, 8) : 0;      // multiple of 8 rows: tile rows start on a 1024-byte swizzle period
  a.a_ps = 0;
  a.a_slot_bytes = (uint32_t)(align_up(128 + 2 * a.halo, 8) * 128);      // one k-chunk of 64 columns, SWIZZLE_128B rows
  a.b_slot_bytes = (uint32_t)((TG_KC / 8) * n_total * 16);
  const int n_chunks = (a.K + TG_KC - 1) / TG_KC;
  int ns_a = n_chunks + 1;
  if (ns_a < 4) ns_a = 4;
  if (ns_a > 6) ns_a = 6;
  a.n_hot = 0;
  if (a.ids != nullptr) {
    int64_t hot[TG_MAX_HOT];
    const int n = hot_tokens(hot);
    for (int i = 0; i < n; ++i)
      if (hot[i] >= 0 && hot[i] < a.V) a.hot_ids[a.n_hot++] = hot[i];
  }
  size_t hot_bytes = (size_t)a.n_hot * ((size_t)(a.K * 2 + 15) / 16 * 16);
  if (a.epi == TG_EPI_RELUGRAD_POOL) {
    MR_REQUIRE(a.ids == nullptr && a.n_sub == 1 && n_total <= 160 && a.ldn % 4 == 0 && a.cmask != nullptr && a.prob != nullptr && a.dnp != nullptr,
               MR_ERR_BAD_SHAPE, "tap gemm: RELUGRAD_POOL epilogue needs a dense A, N <= 160 and its operands");
    hot_bytes = 16 + 2 * (size_t)G * (n_total + 4) * 4;       // two stages of the tile's d_news rows
  }
  size_t left = TG_SMEM_MAX - TG_SMEM_FIXED - hot_bytes - 128;
  while (ns_a > 4 && (size_t)ns_a * a.a_slot_bytes + 2 * (size_t)a.b_slot_bytes > left) --ns_a;
  MR_REQUIRE((size_t)ns_a * a.a_slot_bytes + 2 * (size_t)a.b_slot_bytes <= left, MR_ERR_UNSUPPORTED,
             "tap gemm: tile does not fit shared memory (N=%d)", n_total);
  int ns_b = (int)((left - (size_t)ns_a * a.a_slot_bytes) / a.b_slot_bytes);
  if (ns_b > TG_MAX_SLOTS) ns_b = TG_MAX_SLOTS;
  a.ns_a = ns_a;
  a.ns_b = ns_b;
  a.n_tiles = ceil_div(a.n_titles, (int64_t)G);
  if (a.n_rows <= 0) a.n_rows = a.n_titles * a.L;
  a.w_reps = TG_W_REPS;
  {
    static int rot = -1;
    if (rot < 0) { const char* e = getenv("MINDREC_ROTATE"); rot = (e != nullptr && e[0] == '1') ? 1 : 0; }
    a.rotate = rot;
  }
  a.w_rep_stride = tapgemm_pack_bytes(a.taps, n_total, a.K) / TG_W_REPS;
  if (a.use_tma) {
    if (a.ids != nullptr) {
      if (int rc = tma_encode_2d(&plan->tmap, a.a, (uint64_t)a.lda, (uint64_t)(a.V + (int64_t)a.n_hot * a.hot_reps),
                                 (uint64_t)a.lda * 2, TG_KC, 1, 128))
        return rc;
    } else {
      if (int rc = tma_encode_3d(&plan->tmap, a.a, (uint64_t)a.lda, (uint64_t)a.n_titles, (uint64_t)a.L, (uint64_t)a.L * a.lda * 2,
                                 (uint64_t)a.lda * 2, TG_KC, (uint32_t)G, (uint32_t)a.L, 128))
        return rc;
    }
  } else {
    memset(&plan->tmap, 0, sizeof(plan->tmap));
  }
  plan->args = a;
  plan->smem_bytes = (size_t)ns_a * a.a_slot_bytes + (size_t)ns_b * a.b_slot_bytes + TG_SMEM_FIXED + hot_bytes;
  int64_t g = a.n_tiles < sm_count() ? a.n_tiles : sm_count();
  plan->grid = (int)(g < 1 ? 1 : g);
  plan->colsum_rows = plan->grid * (plan->epi2 ? 8 : 4);
  return MR_OK;
}

int tapgemm_launch(const TapGemmPlan& plan, cudaStream_t stream) {
  if (plan.args.n_titles <= 0) return MR_OK;
  static thread_local size_t attr_set = 0;
  if (plan.smem_bytes > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TG_SMEM_MAX);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "tap gemm: cannot opt in to %zu bytes of shared memory: %s", TG_SMEM_MAX,
               cudaGetErrorString(e));
    attr_set = TG_SMEM_MAX;
  }
  TapGemmArgs args = plan.args;
  args.dbg = g_tapgemm_dbg;
  if (g_tapgemm_dbg != nullptr) g_tapgemm_dbg += 148 * 4 * 5;     // one record block per launch
  if (args.epi == TG_EPI_RELUGRAD_POOL) launch_pdl(tapgemm_kernel<true, true>, dim3(plan.grid), dim3(TG_THREADS2), plan.smem_bytes, stream, args, plan.tmap);
  else if (plan.epi2) launch_pdl(tapgemm_kernel<false, true>, dim3(plan.grid), dim3(TG_THREADS2), plan.smem_bytes, stream, args, plan.tmap);
  else launch_pdl(tapgemm_kernel<false, false>, dim3(plan.grid), dim3(TG_THREADS), plan.smem_bytes, stream, args, plan.tmap);
  MR_CHECK_LAUNCH("tapgemm_kernel");
  return MR_OK;
}

int tapgemm_pack_blocks(const float* src, uint8_t* dst, int taps, int n_total, int K, int n_valid, int k_valid, int64_t sn,
                        int64_t sk, int64_t st, int k_mod, int64_t sb, cudaStream_t stream) {
  const int64_t units = tapgemm_pack_bytes(taps, n_total, K) / TG_W_REPS / 16;
  const uint32_t slot = (uint32_t)((TG_KC / 8) * n_total * 16);
  int blocks = (int)ceil_div(units, 256);
  if (blocks > 1024) blocks = 1024;
  launch_pdl(tapgemm_pack_kernel, dim3(blocks), dim3(256), 0, stream, src, dst, taps, n_total, K, n_valid, k_valid, sn, sk, st, slot, (int)TG_W_REPS,
             (int64_t)(tapgemm_pack_bytes(taps, n_total, K) / TG_W_REPS), k_mod, sb);
  MR_CHECK_LAUNCH("tapgemm_pack_kernel");
  return MR_OK;
}

int tapgemm_pack(const float* src, uint8_t* dst, int taps, int n_total, int K, int n_valid, int k_valid, int64_t sn,
                 int64_t sk, int64_t st, cudaStream_t stream) {
  return tapgemm_pack_blocks(src, dst, taps, n_total, K, n_valid, k_valid, sn, sk, st, 0, 0, stream);
}

}  // namespace mr

extern "C" {
int mr_news_cnn_set_hot_replicas(int reps) {
  using namespace mr;
  MR_REQUIRE(reps >= 0 && reps <= 1024, MR_ERR_BAD_SHAPE, "mr_news_cnn_set_hot_replicas: reps=%d", reps);
  g_hot_reps = reps;
  return MR_OK;
}
int mr_news_cnn_set_hot_tokens(const int64_t* ids, int n) {
  using namespace mr;
  MR_REQUIRE(n >= 0 && n <= TG_MAX_HOT && (n == 0 || ids != nullptr), MR_ERR_BAD_SHAPE, "mr_news_cnn_set_hot_tokens: n=%d (max %d)", n, TG_MAX_HOT);
  for (int i = 0; i < n; ++i) g_hot_ids[i] = ids[i];
  g_n_hot = n;
  return MR_OK;
}
/* debug hook (not part of the reference-facing ABI): per-role wait counters of the next tap-GEMM launches */
__attribute__((visibility("default"))) void mr_debug_tapgemm_counters(long long* device_buffer) { mr::g_tapgemm_dbg = device_buffer; }
}
