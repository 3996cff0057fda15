// Two-CTA ("cta_group::2") variant of the tap GEMM for the conv forward: a CTA pair works on 256 token rows per
// tile (each CTA stages and drains its own 128 rows) and the packed conv weights are RESIDENT in shared memory,
// each CTA holding the half of the output channels it feeds to the pair's tcgen05.mma.  Compared with
// tapgemm_kernel this removes the per-tile weight stream (307 KB of shared-memory writes and L2 reads per tile)
// and halves the B-operand shared-memory reads per SM, which ncu showed to be what paces the single-CTA kernel.
//
//   rank 0 (leader)  warp 4 issues tcgen05.mma.cta_group::2 (M = 256, N = 160, K = 16) and multicast-commits to
//                    the barriers of both CTAs;           rank 1: warp 4 only reports its weights as loaded.
//   both ranks       warps 0-3 epilogue (own 128 TMEM lanes), warp 5 loads the resident weight half once,
//                    warps 6-9 stage the A rows (cp.async gather + hot-row cache) and arrive on the LEADER's
//                    a_full barrier (local arrive on rank 0, mapa + remote arrive on rank 1).
#include "tapgemm.cuh"
#include <stdlib.h>
#include <string.h>

namespace mr {

constexpr int TG2_THREADS = 448;     // warps 0-3 epilogue, 4 MMA / relay, 5 weights, 6-13 A producers
constexpr int TG2_PROD = 256;        // producer threads
constexpr int TG2_RPT = 128 * 8 / TG2_PROD;   // rows per producer thread (a thread owns one 16-byte piece column)

struct Ring2 {
  uint32_t pos, phase, n;
  __device__ __forceinline__ explicit Ring2(uint32_t n_) : pos(0), phase(0), n(n_) {}
  __device__ __forceinline__ void next() {
    if (++pos == n) { pos = 0; phase ^= 1u; }
  }
};

namespace tc {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > (1ll << 31)) {
      printf("libmindrec: cluster mbarrier wait timed out (block %d thread %d smem 0x%x parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on `bar` (same offset) in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
}  // namespace tc

struct TapGemm2Args {
  TapGemmArgs g;                 // problem description (gather mode, n_sub == 1)
  const uint8_t* wpack2;         // [rank][chunk][tap] blocks of b_half_slot bytes: panels x (N/2) rows x 16 B
  uint32_t b_half_slot, b_half_total;
  int64_t n_pairs;
};

#define TG2_TIMED(slot, stmt)                                  \
  do {                                                         \
    if (p.dbg != nullptr) {                                    \
      const long long t0__ = clock64();                        \
      stmt;                                                    \
      dbg_acc[slot] += clock64() - t0__;                       \
    } else {                                                   \
      stmt;                                                    \
    }                                                          \
  } while (0)

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TG2_THREADS, 1) tapgemm2_kernel(const TapGemm2Args q) {
  const TapGemmArgs& p = q.g;
  long long dbg_acc[4] = {0, 0, 0, 0};
  const long long dbg_t0 = clock64();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)p.ns_a * p.a_slot_bytes;                  // resident weights (this CTA's N/2 rows)
  float* sbias = reinterpret_cast<float*>(sB + q.b_half_total);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sbias + 512);         // local: 128 producer arrivals
  uint64_t* a_empty = a_full + TG_MAX_SLOTS;                           // multicast commit
  uint64_t* peer_full = a_empty + TG_MAX_SLOTS;                        // leader: 1 remote arrival per chunk (peer's relay)
  uint64_t* t_full = peer_full + TG_MAX_SLOTS;                         // multicast commit
  uint64_t* t_empty = t_full + 2;                                      // leader: 2 arrivals (one per CTA's epilogue)
  uint64_t* b_ready = t_empty + 2;                                     // local bulk-copy completion
  uint64_t* peer_ready = b_ready + 1;                                  // leader: 1 remote arrival
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_ready + 1);
  uint8_t* hot = reinterpret_cast<uint8_t*>(tmem_slot + 4);
  const uint32_t hot_pitch = (uint32_t)((p.K * 2 + 15) / 16 * 16);

  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = tc::cluster_ctarank();
  const int64_t pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;
  const int N = p.nsz[0], NH = N / 2;
  const int n_chunks = (p.K + TG_KC - 1) / TG_KC;
  const int last_kc = p.K - (n_chunks - 1) * TG_KC;

  pdl_trigger();
  {
    const uint32_t a_bytes = (uint32_t)p.ns_a * p.a_slot_bytes;
    for (uint32_t i = tid * 16; i < a_bytes; i += TG2_THREADS * 16) *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
    pdl_wait();                        // first global read below; the stores above are to this CTA's shared memory
    for (int i = tid; i < 512; i += TG2_THREADS) sbias[i] = (p.bias != nullptr && i < p.n_valid) ? p.bias[i] : 0.f;
    for (uint32_t i = tid; i < (uint32_t)p.n_hot * (hot_pitch / 16); i += TG2_THREADS) {
      const uint32_t h = i / (hot_pitch / 16), pc = i % (hot_pitch / 16);
      reinterpret_cast<uint4*>(hot + (size_t)h * hot_pitch)[pc] = __ldg(reinterpret_cast<const uint4*>(p.a + p.hot_ids[h] * p.lda) + pc);
    }
    if (tid == 0) {
      for (int i = 0; i < TG_MAX_SLOTS; ++i) {
        tc::mbar_init(&a_full[i], TG2_PROD);
        tc::mbar_init(&a_empty[i], 1);
        tc::mbar_init(&peer_full[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        tc::mbar_init(&t_full[i], 1);
        tc::mbar_init(&t_empty[i], 2);
      }
      tc::mbar_init(b_ready, 1);
      tc::mbar_init(peer_ready, 1);
      tc::fence_barrier_init();
    }
    if (warp == 4) tc::tmem_alloc2(tmem_slot, 512);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::cluster_sync_all();          // barriers of both CTAs exist before anyone arrives remotely
    tc::tc_fence_after();
  }
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // =================================== epilogue (own 128 rows) ================================
    const int r = tid;
    const int g = r % p.G, l = r / p.G;
    const bool row_ok = l < p.L;
    const int64_t row_off = (int64_t)g * p.L + l;
    const uint32_t t_empty_leader = tc::map_to_rank(tc::smem_u32(t_empty), 0);
    Ring2 acc(2);
    for (int64_t pair = pair0; pair < q.n_pairs; pair += pair_step) {
      const int64_t tile = 2 * pair + rank;
      TG2_TIMED(0, tc::mbar_wait_cluster(&t_full[acc.pos], acc.phase));
      tc::tc_fence_after();
      const int64_t t = tile * p.G * p.L + row_off;
      const bool valid = row_ok && tile < p.n_tiles && (tile * p.G + g < p.n_titles) && t < p.n_rows;
      const uint32_t tbase = tmem + ((uint32_t)(warp * 32) << 16) + acc.pos * 256u;
      uint32_t cm0 = 0, cm1 = 0, cm2 = 0, cm3 = 0, cm4 = 0;        // sign bits of the row's (<= 160) stored columns
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const bool wide = N - c0 >= 32;
        if (wide) {
          tc::tmem_ld32(tbase + c0, v);
        } else {
          uint32_t h[16];
          tc::tmem_ld16(tbase + c0, h);
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = h[j]; v[16 + j] = 0u; }
        }
        tc::tmem_ld_wait();
        if (valid) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(__uint_as_float(v[j]) + sbias[c0 + j], 0.f);      // bias + ReLU
          uint4* dst = reinterpret_cast<uint4*>(p.out + t * p.ldo + c0);
          const bool al32 = ((reinterpret_cast<uintptr_t>(dst) & 31) == 0);
#pragma unroll
          for (int u = 0; u < 4; u += 2) {
            if (u < 2 || wide) {
              uint4 o[2];
#pragma unroll
              for (int x = 0; x < 2; ++x) {
                o[x].x = tc::pack_bf16(f[(u + x) * 8 + 0], f[(u + x) * 8 + 1]);
                o[x].y = tc::pack_bf16(f[(u + x) * 8 + 2], f[(u + x) * 8 + 3]);
                o[x].z = tc::pack_bf16(f[(u + x) * 8 + 4], f[(u + x) * 8 + 5]);
                o[x].w = tc::pack_bf16(f[(u + x) * 8 + 6], f[(u + x) * 8 + 7]);
              }
              if (al32) tc::st_global_256(dst + u, o[0], o[1]);      // one 32-byte store per row piece: half the LSU wavefronts
              else { dst[u] = o[0]; dst[u + 1] = o[1]; }
              if (p.cmask_out != nullptr) {
                // relu output >= +0: "stored value > 0" is "its bf16 bits are not zero" (h + 0x7fff carries into bit 15 for h >= 1)
                const uint32_t w8[8] = {o[0].x, o[0].y, o[0].z, o[0].w, o[1].x, o[1].y, o[1].z, o[1].w};
                uint32_t bits = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const uint32_t m = ((w8[e] + 0x7FFF7FFFu) & 0x80008000u) >> 15;      // bit 0: low half, bit 16: high half
                  bits |= ((m | (m >> 15)) & 3u) << (2 * e);
                }
                const uint32_t sh = bits << (16 * (u >> 1));
                const int ci = c0 >> 5;
                cm0 |= ci == 0 ? sh : 0u; cm1 |= ci == 1 ? sh : 0u; cm2 |= ci == 2 ? sh : 0u; cm3 |= ci == 3 ? sh : 0u; cm4 |= ci == 4 ? sh : 0u;
              }
            }
          }
        }
      }
      if (p.cmask_out != nullptr && valid) {
        uint8_t* mrow = p.cmask_out + t * 32;
        *reinterpret_cast<uint4*>(mrow) = make_uint4(cm0, cm1, cm2, cm3);
        *reinterpret_cast<uint32_t*>(mrow + 16) = cm4;
      }
      tc::tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");               // the four epilogue warps of this CTA
      if (tid == 0) tc::mbar_arrive_cluster(t_empty_leader + acc.pos * 8u);
      acc.next();
    }
  } else if (warp == 4) {
    if (rank == 0) {
      // ================================= MMA issuer (leader) ===================================
      const uint32_t idesc = tc::make_idesc(256, N, 0, 0);
      const uint32_t b_ps = (uint32_t)NH * 16;
      const uint64_t a_hi = tc::make_desc_sw(0, 16, 1024, 2, 0) & 0xFFFFFFFF00000000ull;
      const uint64_t b_desc0 = tc::make_desc(0, b_ps, 128);
      const uint32_t b_kstep = (2u * b_ps) >> 4;
      const uint32_t sA0 = tc::smem_u32(sA), sB0 = tc::smem_u32(sB);
      const int ctr = (p.taps - 1) / 2;
      const uint32_t tap_off0 = (uint32_t)(p.halo + p.dir * (0 - ctr) * p.G) * 8u;
      const uint32_t tap_off1 = (uint32_t)(p.halo + p.dir * (1 - ctr) * p.G) * 8u;
      const uint32_t tap_off2 = (uint32_t)(p.halo + p.dir * (2 - ctr) * p.G) * 8u;
      tc::mbar_wait(b_ready, 0);                    // own weight half
      tc::mbar_wait_cluster(peer_ready, 0);         // the peer's
      Ring2 ra(p.ns_a), acc(2);
      for (int64_t pair = pair0; pair < q.n_pairs; pair += pair_step) {
        TG2_TIMED(0, tc::mbar_wait_cluster(&t_empty[acc.pos], acc.phase ^ 1u));
        tc::tc_fence_after();
        const uint32_t dcol = tmem + acc.pos * 256u;
        uint32_t accum = 0;
        for (int c = 0; c < n_chunks; ++c) {
          const int nks = (c == n_chunks - 1 ? last_kc : TG_KC) >> 4;
          TG2_TIMED(1, tc::mbar_wait(&a_full[ra.pos], ra.phase));
          TG2_TIMED(2, tc::mbar_wait_cluster(&peer_full[ra.pos], ra.phase));
          tc::tc_fence_after();
          const uint32_t a_slot16 = (sA0 + ra.pos * p.a_slot_bytes) >> 4;
          // all taps of the chunk in one elected section: up to 12 MMAs back to back, one commit
          const uint32_t b_chunk = sB0 + (uint32_t)(c * p.taps) * q.b_half_slot;
          const bool last_chunk = c == n_chunks - 1;
          TG2_TIMED(3, {
          if (tc::elect_one()) {
            for (int tap = 0; tap < p.taps; ++tap) {
              const uint32_t toff = tap == 0 ? tap_off0 : (tap == 1 ? tap_off1 : tap_off2);
              const uint64_t da0 = a_hi | (uint64_t)((a_slot16 + toff) & 0x3FFFu);
              const uint64_t db0 = b_desc0 | (uint64_t)(((b_chunk + (uint32_t)tap * q.b_half_slot) >> 4) & 0x3FFFu);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (ks < nks) tc::umma2(dcol, da0 + (uint64_t)(2 * ks), db0 + (uint64_t)(ks * b_kstep), idesc, accum | (uint32_t)(tap | ks));
            }
            tc::umma2_commit(&a_empty[ra.pos]);
            if (last_chunk) tc::umma2_commit(&t_full[acc.pos]);
          }
          __syncwarp();
          });
          accum = 1;
          ra.next();
        }
        acc.next();
      }
    } else {
      // ================================= relay (peer CTA) ======================================
      // forwards "my A chunk is staged" to the leader with ONE remote arrive per chunk (the 128 producer threads
      // arrive on the local barrier, which is cheap; cluster-scope arrives from every thread were not)
      tc::mbar_wait(b_ready, 0);
      if ((tid & 31) == 0) {
        tc::mbar_arrive_cluster(tc::map_to_rank(tc::smem_u32(peer_ready), 0));
        const uint32_t peer_full_leader = tc::map_to_rank(tc::smem_u32(peer_full), 0);
        Ring2 ra(p.ns_a);
        for (int64_t pair = pair0; pair < q.n_pairs; pair += pair_step)
          for (int c = 0; c < n_chunks; ++c) {
            tc::mbar_wait(&a_full[ra.pos], ra.phase);
            tc::mbar_arrive_cluster(peer_full_leader + ra.pos * 8u);
            ra.next();
          }
      }
    }
  } else if (warp == 5) {
    // =================================== resident weights ====================================
    if ((tid & 31) == 0) {
      const uint8_t* src = q.wpack2 + (size_t)rank * q.b_half_total;
      tc::mbar_arrive_expect_tx(b_ready, q.b_half_total);
      for (uint32_t off = 0; off < q.b_half_total; off += 32768) {
        const uint32_t n = q.b_half_total - off < 32768 ? q.b_half_total - off : 32768;
        tc::bulk_g2s(tc::smem_u32(sB) + off, src + off, n, b_ready);
      }
    }
  } else {
    // =================================== A producers (cp.async) =============================
    const int ptid = tid - 192;
    const int rgrp = ptid >> 3, j = ptid & 7;
    const uint32_t depth = (uint32_t)(p.ns_a - 2 < 3 ? p.ns_a - 2 : 3);
    const uint32_t sA0 = tc::smem_u32(sA);
    int64_t row_off[TG2_RPT];
    int row_g[TG2_RPT];
    uint32_t dst_off[TG2_RPT];
#pragma unroll
    for (int s = 0; s < TG2_RPT; ++s) {
      const int r = rgrp + (TG2_PROD / 8) * s;
      const int g = r % p.G, l = r / p.G;
      row_g[s] = g;
      row_off[s] = l < p.L ? (int64_t)g * p.L + l : -1;
      const uint32_t row = (uint32_t)(p.halo + r);
      dst_off[s] = row * 128u + ((((uint32_t)j ^ row) & 7u) << 4);
    }
    auto row_token = [&](int64_t tile, int s) -> int64_t {
      if (row_off[s] < 0 || tile >= p.n_tiles || tile * p.G + row_g[s] >= p.n_titles) return -1;
      const int64_t t_ = tile * p.G * p.L + row_off[s];
      return t_ < p.n_rows ? t_ : -1;                 // rows past the problem (flat row lists padded to whole tiles): zero rows, no id read
    };
    auto raw_of = [&](int64_t t) -> int64_t { return (t < 0 || p.ids == nullptr) ? t : load_index(p.ids, p.ids_i64, t); };
    auto classify = [&](int64_t t, int64_t raw) -> int64_t {
      if (t < 0 || p.ids == nullptr) return t;
      int64_t id = raw < 0 ? 0 : (raw >= p.V ? p.V - 1 : raw);
#pragma unroll
      for (int h = 0; h < TG_MAX_HOT; ++h)
        if (h < p.n_hot && id == p.hot_ids[h]) id = -2 - h;
      return id;
    };
    int64_t raw[TG2_RPT];
#pragma unroll
    for (int s = 0; s < TG2_RPT; ++s) raw[s] = raw_of(row_token(2 * pair0 + rank, s));
    Ring2 ra(p.ns_a), sig(p.ns_a);
    uint32_t pending = 0;
    for (int64_t pair = pair0; pair < q.n_pairs; pair += pair_step) {
      const int64_t tile = 2 * pair + rank;
      int64_t cur[TG2_RPT];
#pragma unroll
      for (int s = 0; s < TG2_RPT; ++s) cur[s] = classify(row_token(tile, s), raw[s]);
#pragma unroll
      for (int s = 0; s < TG2_RPT; ++s) raw[s] = raw_of(row_token(tile + 2 * pair_step, s));
      const __nv_bfloat16* rowp[TG2_RPT];
      uint32_t hot_mask = 0;
#pragma unroll
      for (int s = 0; s < TG2_RPT; ++s) {
        if (cur[s] >= 0) rowp[s] = p.a + cur[s] * p.lda + j * 8;
        else if (cur[s] == -1) rowp[s] = nullptr;
        else rowp[s] = reinterpret_cast<const __nv_bfloat16*>(hot + (size_t)(-2 - cur[s]) * hot_pitch) + j * 8;
        hot_mask |= (cur[s] <= -2 ? 1u : 0u) << s;
      }
      for (int c = 0; c < n_chunks; ++c) {
        const int kc = c == n_chunks - 1 ? last_kc : TG_KC;
        TG2_TIMED(0, tc::mbar_wait_cluster(&a_empty[ra.pos], ra.phase ^ 1u));
        if (j * 8 < kc) {
          const uint32_t slot = sA0 + ra.pos * p.a_slot_bytes;
          const int col = c * TG_KC;
#pragma unroll
          for (int s = 0; s < TG2_RPT; ++s) {
            const __nv_bfloat16* src = rowp[s];
            if ((hot_mask >> s) & 1u) {
              const uint4 v = *reinterpret_cast<const uint4*>(src + col);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot + dst_off[s]), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            } else {
              tc::cp_async16(slot + dst_off[s], src != nullptr ? (const void*)(src + col) : (const void*)p.a, src != nullptr ? 16u : 0u);
            }
          }
        }
        tc::cp_async_commit();
        ra.next();
        if (++pending > depth) {
          TG2_TIMED(1, {
          if (depth == 3) tc::cp_async_wait<3>();
          else if (depth == 2) tc::cp_async_wait<2>();
          else tc::cp_async_wait<1>();
          tc::fence_proxy_async();
          });
          TG2_TIMED(2, tc::mbar_arrive(&a_full[sig.pos]));
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

  if (p.dbg != nullptr && (tid == 0 || tid == 128 || tid == 192)) {
    long long* d = p.dbg + ((size_t)blockIdx.x * 4 + (tid == 0 ? 0 : tid == 128 ? 1 : 3)) * 5;
    d[0] = dbg_acc[0]; d[1] = dbg_acc[1]; d[2] = dbg_acc[2]; d[3] = dbg_acc[3]; d[4] = clock64() - dbg_t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();            // the peer's shared memory / TMEM stay alive until the pair is done
  if (warp == 4) tc::tmem_dealloc2(tmem, 512);
}

// W[tap][n, k] packed per CTA rank: [rank][chunk][tap][panel][n % (N/2)][8 x bf16]
__global__ void tapgemm2_pack_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int taps, int N, int K, int n_valid,
                                     int k_valid, int64_t sn, int64_t sk, int64_t st, uint32_t half_slot, uint32_t half_total) {
  pdl_trigger();
  pdl_wait();
  const int n_chunks = (K + TG_KC - 1) / TG_KC;
  const int NH = N / 2;
  const int64_t total = (int64_t)n_chunks * taps * (TG_KC / 8) * N;
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(u % N);
    int64_t rest = u / N;
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
      if (n < n_valid && k < k_valid) lo = src[n * sn + k * sk + tap * st];
      if (n < n_valid && k + 1 < k_valid) hi = src[n * sn + (k + 1) * sk + tap * st];
      w[e] = tc::pack_bf16(lo, hi);
    }
    const int rank = n / NH, nl = n % NH;
    uint8_t* o = dst + (size_t)rank * half_total + (size_t)(c * taps + tap) * half_slot + (size_t)panel * NH * 16 + (size_t)nl * 16;
    *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---- host ------------------------------------------------------------------------------------------
static bool use_2cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MINDREC_2CTA");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

static void geom2(const TapGemmArgs& a, int* G, int* halo, uint32_t* a_slot, uint32_t* half_slot, uint32_t* half_total) {
  int g = 128 / a.L;
  if (g > 16) g = 16;
  if (g < 1) g = 1;
  *G = g;
  *halo = a.taps > 1 ? g : 0;
  *a_slot = (uint32_t)(align_up(128 + 2 * *halo, 8) * 128);
  const int n_chunks = (a.K + TG_KC - 1) / TG_KC;
  *half_slot = (uint32_t)((TG_KC / 8) * (a.nsz[0] / 2) * 16);
  *half_total = (uint32_t)(n_chunks * a.taps) * *half_slot;
}

constexpr size_t TG2_FIXED = 512 * 4 + (3 * TG_MAX_SLOTS + 6) * 8 + 16;

int64_t tapgemm2_pack_bytes(int taps, int N, int K) {
  return (int64_t)2 * ((K + TG_KC - 1) / TG_KC) * taps * (TG_KC / 8) * (N / 2) * 16;
}

// true when this problem can run on the two-CTA kernel (token-table gather, one N block, weights resident)
bool tapgemm2_supported(const TapGemmArgs& a) {
  if (!use_2cta() || a.ids == nullptr || a.n_sub != 1 || a.L < 1 || a.L > 128 || a.nsz[0] % 32 != 0 || a.nsz[0] > 256) return false;
  if (a.epi != TG_EPI_BIAS_RELU || (a.cmask_out != nullptr && a.nsz[0] > 160)) return false;
  int G, halo;
  uint32_t a_slot, half_slot, half_total;
  geom2(a, &G, &halo, &a_slot, &half_slot, &half_total);
  const size_t hot_bytes = (size_t)TG_MAX_HOT * ((size_t)(a.K * 2 + 15) / 16 * 16);
  return (size_t)half_total + 4 * (size_t)a_slot + TG2_FIXED + hot_bytes + 128 <= 227 * 1024;
}

int tapgemm2_pack(const float* src, uint8_t* dst, int taps, int N, int K, int n_valid, int k_valid, int64_t sn, int64_t sk,
                  int64_t st, cudaStream_t stream) {
  const int n_chunks = (K + TG_KC - 1) / TG_KC;
  const uint32_t half_slot = (uint32_t)((TG_KC / 8) * (N / 2) * 16);
  const uint32_t half_total = (uint32_t)(n_chunks * taps) * half_slot;
  const int64_t units = (int64_t)n_chunks * taps * (TG_KC / 8) * N;
  int blocks = (int)ceil_div(units, 256);
  if (blocks > 1024) blocks = 1024;
  launch_pdl(tapgemm2_pack_kernel, dim3(blocks), dim3(256), 0, stream, src, dst, taps, N, K, n_valid, k_valid, sn, sk, st, half_slot, half_total);
  MR_CHECK_LAUNCH("tapgemm2_pack_kernel");
  return MR_OK;
}

int tapgemm2_run(TapGemmArgs a, const uint8_t* wpack2, cudaStream_t stream) {
  TapGemm2Args q{};
  int G, halo;
  uint32_t a_slot, half_slot, half_total;
  geom2(a, &G, &halo, &a_slot, &half_slot, &half_total);
  a.G = G;
  a.halo = halo;
  a.a_slot_bytes = a_slot;
  a.use_tma = 0;
  a.n_hot = 0;
  int64_t hot[TG_MAX_HOT];
  const int nh = hot_tokens(hot);
  for (int i = 0; i < nh; ++i)
    if (hot[i] >= 0 && hot[i] < a.V) a.hot_ids[a.n_hot++] = hot[i];
  const size_t hot_bytes = (size_t)a.n_hot * ((size_t)(a.K * 2 + 15) / 16 * 16);
  size_t left = 227 * 1024 - TG2_FIXED - hot_bytes - 128 - half_total;
  int ns_a = (int)(left / a_slot);
  if (ns_a > 6) ns_a = 6;
  MR_REQUIRE(ns_a >= 4, MR_ERR_UNSUPPORTED, "two-CTA tap gemm: only %d A slots fit", ns_a);
  a.ns_a = ns_a;
  a.ns_b = 0;
  a.n_tiles = ceil_div(a.n_titles, (int64_t)G);
  if (a.n_rows <= 0) a.n_rows = a.n_titles * a.L;
  a.dbg = g_tapgemm_dbg;
  if (g_tapgemm_dbg != nullptr) g_tapgemm_dbg += 148 * 4 * 5;
  q.g = a;
  q.wpack2 = wpack2;
  q.b_half_slot = half_slot;
  q.b_half_total = half_total;
  q.n_pairs = ceil_div(a.n_tiles, (int64_t)2);
  if (a.n_titles <= 0) return MR_OK;
  const size_t smem = (size_t)ns_a * a_slot + half_total + TG2_FIXED + hot_bytes;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    MR_REQUIRE(e == cudaSuccess, MR_ERR_LAUNCH, "two-CTA tap gemm: shared-memory opt-in failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  int64_t clusters = sm_count() / 2;
  if (clusters > q.n_pairs) clusters = q.n_pairs;
  if (clusters < 1) clusters = 1;
  launch_pdl(tapgemm2_kernel, dim3((unsigned)(2 * clusters)), dim3(TG2_THREADS), smem, stream, q);
  MR_CHECK_LAUNCH("tapgemm2_kernel");
  return MR_OK;
}

}  // namespace mr
