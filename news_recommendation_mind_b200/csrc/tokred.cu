// Token-reduction GEMM (tokred.cuh): kernel, planning, launch and the fixed-order partial reduction.
#include "tokred.cuh"
#include "tapgemm.cuh"   // sm_count()
#include <string.h>

namespace mr {

constexpr int TR_MAX_STAGES = 4;

#define TR_TIMED(slot, stmt)                                   \
  do {                                                         \
    if (p.dbg != nullptr) {                                    \
      const long long t0__ = clock64();                        \
      stmt;                                                    \
      dbg_acc[slot] += clock64() - t0__;                       \
    } else {                                                   \
      stmt;                                                    \
    }                                                          \
  } while (0)

__global__ void __launch_bounds__(TR_THREADS, 1) tokred_kernel(const TokRedArgs p, const __grid_constant__ CUtensorMap pmap,
                                                              const __grid_constant__ CUtensorMap qmap) {
  long long dbg_acc[4] = {0, 0, 0, 0};
  const long long dbg_t0 = clock64();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * p.stage_bytes);
  uint64_t* empty = full + TR_MAX_STAGES;
  uint64_t* done = empty + TR_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  uint8_t* hot = reinterpret_cast<uint8_t*>(tmem_slot + 2);          // [n_hot][256 B]: this CTA's 128-column slice of the hot rows

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m = blockIdx.x / p.S, s = blockIdx.x % p.S;
  const bool has_tiles = (int64_t)s < p.n_tiles;

  pdl_trigger();
  {
    const uint32_t bytes = (uint32_t)p.n_stages * p.stage_bytes;
    for (uint32_t i = tid * 16; i < bytes; i += TR_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    pdl_wait();                        // first global read below
    if (p.ids != nullptr)
      for (int i = tid; i < p.n_hot * 16; i += TR_THREADS) {
        const int h = i >> 4, q = i & 15;
        const int col = (m * 16 + q) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (col < p.KP) v = __ldg(reinterpret_cast<const uint4*>(p.p + p.hot_ids[h] * p.ldp + col));
        reinterpret_cast<uint4*>(hot + (size_t)h * 256)[q] = v;
      }
    if (tid == 0) {
      for (int i = 0; i < TR_MAX_STAGES; ++i) {
        tc::mbar_init(&full[i], (p.p_tma ? 0 : 128) + (p.q_tma ? 1 : 0));
        tc::mbar_init(&empty[i], 1);
      }
      tc::mbar_init(done, 1);
      tc::fence_barrier_init();
    }
    if (warp == 4) tc::tmem_alloc(tmem_slot, 512);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // ---- epilogue: TMEM -> fp32 partial -------------------------------------------------------
    float* dst = p.partial + ((size_t)(m * p.S + s) * p.taps * 128 + tid) * p.NQ;
    if (has_tiles) {
      tc::mbar_wait(done, 0);
      tc::tc_fence_after();
    }
    for (int tap = 0; tap < p.taps; ++tap) {
      float* d = dst + (size_t)tap * 128 * p.NQ;
      for (int c0 = 0; c0 < p.NQ; c0 += 16) {
        uint32_t v[16];
        if (has_tiles) {
          tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tap * p.NQ + c0), v);
          tc::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(d + c0 + 4 * u) = make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
      }
    }
  } else if (warp == 4) {
    // ---- MMA issuer (warp-uniform loop, one elected lane issues) --------------------------------------
    if (has_tiles) {
      const uint32_t idesc = tc::make_idesc(128, p.NQ, 1, 1);
      const int ctr = (p.taps - 1) / 2;
      const uint64_t a_tmpl = tc::make_desc_sw(0, p.p_ps, 1024, 2, 0);
      const uint64_t b_tmpl = tc::make_desc_sw(0, p.q_ps, 8 * p.q_rb, p.q_layout, 0);
      const uint32_t a_kstep = (16u * 128u) >> 4, b_kstep = (16u * p.q_rb) >> 4;      // 16 token rows per k-step
      uint32_t it = 0;
      for (int64_t tile = s; tile < p.n_tiles; tile += p.S, ++it) {
        const uint32_t st = it % (uint32_t)p.n_stages;
        TR_TIMED(0, tc::mbar_wait(&full[st], (it / (uint32_t)p.n_stages) & 1u));
        tc::tc_fence_after();
        const uint32_t pbase = tc::smem_u32(smem) + st * p.stage_bytes;
        const uint32_t qbase = pbase + p.p_bytes;
        const uint64_t db0 = b_tmpl | (uint64_t)((qbase >> 4) & 0x3FFFu);
        TR_TIMED(1, {
        if (tc::elect_one()) {
          for (int tap = 0; tap < p.taps; ++tap) {
            const int shift = (tap - ctr) * p.G;
            const uint64_t da0 = a_tmpl | (uint64_t)(((pbase + (uint32_t)(p.halo + shift) * 128u) >> 4) & 0x3FFFu);
            const uint32_t dcol = tmem + (uint32_t)(tap * p.NQ);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              tc::umma(dcol, da0 + (uint64_t)(ks * a_kstep), db0 + (uint64_t)(ks * b_kstep), idesc, (it | (uint32_t)ks) ? 1u : 0u);
          }
          tc::umma_commit(&empty[st]);
        }
        __syncwarp();
        });
      }
      if (tc::elect_one()) tc::umma_commit(done);
      __syncwarp();
    }
  } else if (has_tiles) {
    // ---- producers: stage P (slice m of its columns) and Q for every tile of this CTA ---------------
    const int ptid = tid - 160;            // 0..127
    const int rgrp = ptid >> 3, j = ptid & 7;
    const int q_panels = p.NQ / 8;
    const uint32_t depth = p.n_stages >= 3 ? 1u : 0u;
    const uint32_t ppb_shift = p.q_rb == 128 ? 3u : (p.q_rb == 64 ? 2u : 1u);       // log2(16-byte pieces per Q row)
    const uint32_t ppb_mask = (1u << ppb_shift) - 1u;
    // tile-independent geometry of this thread's rows r = rgrp + 16*ss (token = tile*G*L + row_off)
    int64_t row_off[8];
    int row_g[8];
    uint32_t p_dst[8], q_dst[8], q_x[8];
#pragma unroll
    for (int ss = 0; ss < 8; ++ss) {
      const int r = rgrp + 16 * ss;
      const int g = r % p.G, l = r / p.G;
      row_g[ss] = g;
      row_off[ss] = l < p.L ? (int64_t)g * p.L + l : -1;
      const uint32_t prow_s = (uint32_t)(p.halo + r);
      p_dst[ss] = prow_s * 128u + ((((uint32_t)j ^ prow_s) & 7u) << 4);
      q_dst[ss] = (uint32_t)r * p.q_rb;
      q_x[ss] = (((uint32_t)r * p.q_rb) >> 7) & ppb_mask;
    }
    const int pcol0 = (m * 16 + j) * 8, pcol1 = pcol0 + 64;
    const bool pok0 = pcol0 < p.KP, pok1 = pcol1 < p.KP;
    auto token_of = [&](int64_t tile, int ss) -> int64_t {
      if (row_off[ss] < 0 || tile >= p.n_tiles || tile * p.G + row_g[ss] >= p.n_titles) return -1;
      return tile * p.G * p.L + row_off[ss];
    };
    // raw id of a token (the load is issued one tile ahead and only classified at the top of the next iteration,
    // so its latency overlaps the staging of the current tile)
    auto raw_of = [&](int64_t t) -> int64_t { return (t < 0 || p.ids == nullptr) ? t : load_index(p.ids, p.ids_i64, t); };
    // source row of P: token index (dense), token id (gather), -1 = zero row, -2-h = hot row h (shared memory)
    auto classify = [&](int64_t t, int64_t raw) -> int64_t {
      if (t < 0 || p.ids == nullptr) return t;
      int64_t id = raw < 0 ? 0 : (raw >= p.V ? p.V - 1 : raw);
#pragma unroll
      for (int h = 0; h < 4; ++h)
        if (h < p.n_hot && id == p.hot_ids[h]) id = -2 - h;
      return id;
    };
    int64_t raw[8];
#pragma unroll
    for (int ss = 0; ss < 8; ++ss) raw[ss] = raw_of(token_of(s, ss));
    uint32_t it = 0, signaled = 0;
    for (int64_t tile = s; tile < p.n_tiles; tile += p.S, ++it) {
      const uint32_t st = it % (uint32_t)p.n_stages;
      int64_t cur[8];
#pragma unroll
      for (int ss = 0; ss < 8; ++ss) cur[ss] = classify(token_of(tile, ss), raw[ss]);
#pragma unroll
      for (int ss = 0; ss < 8; ++ss) raw[ss] = raw_of(token_of(tile + p.S, ss));   // in flight while this tile is staged
      TR_TIMED(0, tc::mbar_wait(&empty[st], ((it / (uint32_t)p.n_stages) & 1u) ^ 1u));
      const long long t_issue0 = clock64();
      const uint32_t pbase = tc::smem_u32(smem) + st * p.stage_bytes;
      const uint32_t qbase = pbase + p.p_bytes;
      if (p.q_tma && ptid == 0) {
        // TMA: one 3-D tile load per Q block (and per P block when P is dense); the box arrives in the
        // position-major row order, swizzled by the copy engine
        const uint32_t rows = (uint32_t)(p.G * p.L);
        const uint32_t q_blocks = (uint32_t)(p.NQ * 2) / p.q_rb;
        tc::mbar_arrive_expect_tx(&full[st], rows * (q_blocks * p.q_rb + (p.p_tma ? 256u : 0u)));
        for (uint32_t b = 0; b < q_blocks; ++b)
          tc::tma_load_3d(qbase + b * p.q_ps, &qmap, (int)(b * (p.q_rb / 2)), (int)(tile * p.G), 0, &full[st]);
        if (p.p_tma) {
          tc::tma_load_3d(pbase + (uint32_t)p.halo * 128u, &pmap, m * 128, (int)(tile * p.G), 0, &full[st]);
          tc::tma_load_3d(pbase + p.p_ps + (uint32_t)p.halo * 128u, &pmap, m * 128 + 64, (int)(tile * p.G), 0, &full[st]);
        }
      }
      if (p.p_tma) continue;                       // nothing left for the cp.async threads
#pragma unroll
      for (int ss = 0; ss < 8; ++ss) {
        const int64_t t = token_of(tile, ss);
        const bool valid = t >= 0;
        const bool is_hot = valid && cur[ss] <= -2;
        const uint32_t d0 = pbase + p_dst[ss];
        if (is_hot) {                               // hot token: 16-byte smem -> smem copies
          const uint4* hrow = reinterpret_cast<const uint4*>(hot + (size_t)(-2 - cur[ss]) * 256);
          const uint4 v0 = hrow[j], v1 = hrow[j + 8];
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d0), "r"(v0.x), "r"(v0.y), "r"(v0.z), "r"(v0.w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d0 + p.p_ps), "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w) : "memory");
        } else {
          const __nv_bfloat16* prow = p.p + (valid ? cur[ss] : 0) * p.ldp;
          tc::cp_async16(d0, valid && pok0 ? (const void*)(prow + pcol0) : (const void*)p.q, valid && pok0 ? 16u : 0u);
          tc::cp_async16(d0 + p.p_ps, valid && pok1 ? (const void*)(prow + pcol1) : (const void*)p.q, valid && pok1 ? 16u : 0u);
        }
        if (!p.q_tma) {
          const __nv_bfloat16* qrow = p.q + (valid ? t : 0) * p.ldq;
          for (int jj = j; jj < q_panels; jj += 8) {
            const uint32_t blk = (uint32_t)jj >> ppb_shift, q = (uint32_t)jj & ppb_mask;
            tc::cp_async16(qbase + blk * p.q_ps + q_dst[ss] + ((q ^ q_x[ss]) << 4), valid ? (const void*)(qrow + jj * 8) : (const void*)p.q,
                           valid ? 16u : 0u);
          }
        }
      }
      tc::cp_async_commit();
      dbg_acc[2] += clock64() - t_issue0;
      if (it + 1 - signaled > depth) {
        TR_TIMED(1, {
        if (depth == 1) tc::cp_async_wait<1>();
        else tc::cp_async_wait<0>();
        tc::fence_proxy_async();
        });
        tc::mbar_arrive(&full[signaled % (uint32_t)p.n_stages]);
        ++signaled;
      }
    }
    if (!p.p_tma) {
      tc::cp_async_wait<0>();
      tc::fence_proxy_async();
      while (signaled < it) {
        tc::mbar_arrive(&full[signaled % (uint32_t)p.n_stages]);
        ++signaled;
      }
    }
  }

  if (p.dbg != nullptr && (tid == 128 || tid == 160)) {
    // rows: 1 mma {full wait, issue}, 3 producer {empty wait, cp.wait, issue}
    long long* d = p.dbg + ((size_t)blockIdx.x * 4 + (tid == 128 ? 1 : 3)) * 5;
    d[0] = dbg_acc[0]; d[1] = dbg_acc[1]; d[2] = dbg_acc[2]; d[3] = dbg_acc[3]; d[4] = clock64() - dbg_t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

// 256 threads = 64 outputs x 4 groups of partials (4 independent chains each); fixed combine order
__global__ void __launch_bounds__(256) tokred_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dst, int S, int taps, int NQ,
                                                            int i_valid, int j_valid, int64_t dj, int64_t di, int64_t dt) {
  __shared__ float red[4][64];
  pdl_trigger();
  pdl_wait();
  const int ox = threadIdx.x & 63, sg = threadIdx.x >> 6;
  const int64_t total = (int64_t)taps * i_valid * j_valid;
  const int64_t idx = (int64_t)blockIdx.x * 64 + ox;
  const bool on = idx < total;
  int j = 0, i = 0, tap = 0;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (on) {
    j = (int)(idx % j_valid);
    i = (int)((idx / j_valid) % i_valid);
    tap = (int)(idx / ((int64_t)j_valid * i_valid));
    const int m = i / 128, il = i % 128;
    const float* src = partial + (((size_t)m * S * taps + tap) * 128 + il) * NQ + j;
    const size_t stride = (size_t)taps * 128 * NQ;
    const int per = (S + 3) / 4;
    const int s0 = min(S, sg * per), s1 = min(S, s0 + per);
    int s = s0;
    for (; s + 3 < s1; s += 4) {
      a0 += src[(size_t)s * stride];
      a1 += src[(size_t)(s + 1) * stride];
      a2 += src[(size_t)(s + 2) * stride];
      a3 += src[(size_t)(s + 3) * stride];
    }
    for (; s < s1; ++s) a0 += src[(size_t)s * stride];
  }
  red[sg][ox] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (sg == 0 && on) dst[j * dj + i * di + tap * dt] = (red[0][ox] + red[1][ox]) + (red[2][ox] + red[3][ox]);
}

static void tokred_geometry(int64_t n_titles, int L, int taps, int KP, int* G, int* n_mtiles, int* S, int64_t* n_tiles) {
  int g = 128 / L;
  if (g > 16) g = 16;
  if (g < 1) g = 1;
  *G = g;
  *n_tiles = ceil_div(n_titles, (int64_t)g);
  *n_mtiles = (KP + 127) / 128;
  int64_t s = sm_count() / *n_mtiles;
  if (s > *n_tiles) s = *n_tiles;
  if (s < 1) s = 1;
  *S = (int)s;
}

int64_t tokred_partial_bytes(int64_t n_titles, int L, int taps, int KP, int NQ) {
  int G, n_mtiles, S;
  int64_t n_tiles;
  tokred_geometry(n_titles, L < 1 ? 1 : (L > 128 ? 128 : L), taps, KP, &G, &n_mtiles, &S, &n_tiles);
  return (int64_t)n_mtiles * S * taps * 128 * NQ * 4;
}

int tokred_plan(TokRedArgs& a, TokRedPlan* plan) {
  MR_REQUIRE(a.L >= 1 && a.L <= 128, MR_ERR_UNSUPPORTED, "token-reduction gemm: signal_length %d not in [1,128]", a.L);
  MR_REQUIRE(a.taps == 1 || a.taps == 3, MR_ERR_BAD_SHAPE, "token-reduction gemm: taps=%d", a.taps);
  MR_REQUIRE(a.KP >= 8 && a.KP % 8 == 0, MR_ERR_BAD_SHAPE, "token-reduction gemm: KP=%d", a.KP);
  MR_REQUIRE(a.NQ >= 16 && a.NQ % 16 == 0 && a.NQ <= 256 && a.taps * a.NQ <= 512, MR_ERR_UNSUPPORTED,
             "token-reduction gemm: taps*NQ = %d*%d exceeds the 512 TMEM columns", a.taps, a.NQ);
  tokred_geometry(a.n_titles, a.L, a.taps, a.KP, &a.G, &a.n_mtiles, &a.S, &a.n_tiles);
  a.q_tma = use_tma_default() ? 1 : 0;
  a.p_tma = (a.q_tma && (a.ids == nullptr)) ? 1 : 0;
  // TMA destinations start on a 1024-byte swizzle period (halo a multiple of 8 rows); cp.async staging has no such need
  a.halo = a.taps > 1 ? (a.p_tma ? (int)align_up(a.G, 8) : a.G) : 0;
  // P: two 64-column blocks of [rows x 128 B] (SWIZZLE_128B);  Q: NQ columns in blocks of 64 / 32 / 16 columns
  // (SWIZZLE_128B / 64B / 32B -- the widest row that divides NQ), 128 rows each; every block 1024-byte aligned
  a.p_ps = (uint32_t)(align_up(128 + 2 * a.halo, 8) * 128);
  a.p_bytes = 2 * a.p_ps;
  a.q_layout = a.NQ % 64 == 0 ? 2 : (a.NQ % 32 == 0 ? 4 : 6);
  a.q_rb = a.q_layout == 2 ? 128u : (a.q_layout == 4 ? 64u : 32u);
  a.q_ps = 128 * a.q_rb;
  a.stage_bytes = (uint32_t)align_up(a.p_bytes + (uint32_t)(a.NQ * 2 / a.q_rb) * a.q_ps, 1024);
  a.n_hot = 0;
  if (a.ids != nullptr) {
    int64_t hot[TG_MAX_HOT];
    const int n = hot_tokens(hot);
    for (int i = 0; i < n; ++i)
      if (hot[i] >= 0 && hot[i] < a.V) a.hot_ids[a.n_hot++] = hot[i];
  }
  const size_t fixed = (2 * TR_MAX_STAGES + 1) * 8 + 16 + 4 * 256;
  int ns = (int)((227 * 1024 - fixed - 128) / a.stage_bytes);
  if (ns > TR_MAX_STAGES) ns = TR_MAX_STAGES;
  MR_REQUIRE(ns >= 1, MR_ERR_UNSUPPORTED, "token-reduction gemm: stage of %u bytes does not fit shared memory", a.stage_bytes);
  a.n_stages = ns;
  memset(&plan->pmap, 0, sizeof(CUtensorMap));
  memset(&plan->qmap, 0, sizeof(CUtensorMap));
  if (a.q_tma)
    if (int rc = tma_encode_3d(&plan->qmap, a.q, (uint64_t)a.ldq, (uint64_t)a.n_titles, (uint64_t)a.L, (uint64_t)a.L * a.ldq * 2,
                               (uint64_t)a.ldq * 2, a.q_rb / 2, (uint32_t)a.G, (uint32_t)a.L, (int)a.q_rb))
      return rc;
  if (a.p_tma)
    if (int rc = tma_encode_3d(&plan->pmap, a.p, (uint64_t)a.ldp, (uint64_t)a.n_titles, (uint64_t)a.L, (uint64_t)a.L * a.ldp * 2,
                               (uint64_t)a.ldp * 2, 64, (uint32_t)a.G, (uint32_t)a.L, 128))
      return rc;
  plan->args = a;
  plan->smem_bytes = (size_t)ns * a.stage_bytes + fixed;
  plan->grid = a.n_mtiles * a.S;
  return MR_OK;
}

int tokred_launch(const TokRedPlan& plan, cudaStream_t stream) {
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tokred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "token-reduction gemm: shared-memory opt-in failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  TokRedArgs args = plan.args;
  args.dbg = g_tapgemm_dbg;
  if (g_tapgemm_dbg != nullptr) g_tapgemm_dbg += 148 * 4 * 5;
  launch_pdl(tokred_kernel, dim3(plan.grid), dim3(TR_THREADS), plan.smem_bytes, stream, args, plan.pmap, plan.qmap);
  MR_CHECK_LAUNCH("tokred_kernel");
  return MR_OK;
}

int tokred_reduce(const TokRedPlan& plan, float* dst, int i_valid, int j_valid, int64_t dj, int64_t di, int64_t dt,
                  cudaStream_t stream) {
  const TokRedArgs& a = plan.args;
  const int64_t total = (int64_t)a.taps * i_valid * j_valid;
  launch_pdl(tokred_reduce_kernel, dim3((unsigned)ceil_div(total, 64)), dim3(256), 0, stream, (const float*)a.partial, dst, a.S, a.taps, a.NQ,
             i_valid, j_valid, dj, di, dt);
  MR_CHECK_LAUNCH("tokred_reduce_kernel");
  return MR_OK;
}

}  // namespace mr
