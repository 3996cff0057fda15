// Self-test of the tcgen05 building blocks (descriptor conventions of tc05.cuh): one CTA stages two
// small fp32 matrices into the bf16 panel layout, runs D = A x B^T on the tensor core for every
// combination of operand majorness / row shift the production kernels rely on, and returns D.
// tests/test_gpu_tc.py compares it with a bf16-rounded fp32 matmul.
#include "common.cuh"
#include "tc05.cuh"

namespace mr {

struct SelfTestArgs {
  const float* a; int ra, ca;   // source A, row-major [ra, ca]
  const float* b; int rb, cb;   // source B, row-major [rb, cb]
  float* d;                     // out [128, N] fp32
  int a_mn, b_mn;               // 0: rows are the M/N index, cols are K;  1: rows are K, cols are M/N
  int N, K;                     // MMA N (multiple of 16, <= 256) and total K (multiple of 16)
  int a_shift, halo;            // A operand starts `a_shift` rows later (|a_shift| <= halo)
  int swap;                     // debug: swap the LBO / SBO fields
  int a_layout, b_layout;       // 0 = SWIZZLE_NONE panel layout; 2/4/6 = row-linear SWIZZLE_128B/64B/32B (tc05.cuh);
                                // a_layout 8: the A operand is read from TENSOR MEMORY (K-major, bf16 pairs per 32-bit column)
  int base_off_mode;            // swizzled operands: 0 = base_offset field 0, 1 = (start >> 7) & 7
};

// row-linear swizzled layout: rows of RB = 128/64/32 bytes (CB = RB/2 bf16 columns per block), consecutive
// rows RB bytes apart, 16-byte chunks XOR-ed with the row's position inside the 1024-byte swizzle period
__device__ __forceinline__ int sw_rb(int layout) { return layout == 2 ? 128 : (layout == 4 ? 64 : 32); }
__device__ __forceinline__ size_t sw_off(int layout, size_t blk_stride, int row, int col) {
  const int rb = sw_rb(layout), cb = rb / 2;
  const int blk = col / cb, cc = col % cb;
  const size_t rowb = (size_t)row * rb;
  const int x = (int)((rowb >> 7) & (rb / 16 - 1));
  return (size_t)blk * blk_stride + rowb + (size_t)(((cc / 8) ^ x) * 16) + (cc % 8) * 2;
}

__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(SelfTestArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int a_rows = (p.ra + 2 * p.halo) | 1, b_rows = p.rb | 1;
  uint32_t a_ps = a_rows * 16, b_ps = b_rows * 16;                 // panel / block strides (bytes)
  int a_panels = p.ca / 8, b_panels = p.cb / 8;
  if (p.a_layout) { a_ps = (uint32_t)((p.ra + 2 * p.halo + 7) / 8 * 8) * sw_rb(p.a_layout); a_panels = (p.ca * 2 + sw_rb(p.a_layout) - 1) / sw_rb(p.a_layout); }
  if (p.b_layout) { b_ps = (uint32_t)((p.rb + 7) / 8 * 8) * sw_rb(p.b_layout); b_panels = (p.cb * 2 + sw_rb(p.b_layout) - 1) / sw_rb(p.b_layout); }
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((size_t)a_panels * a_ps + 1023) / 1024 * 1024;
  const size_t total = ((size_t)a_panels * a_ps + 1023) / 1024 * 1024 + (size_t)b_panels * b_ps;
  for (size_t i = tid * 16; i < total; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  for (int i = tid; p.a_layout != 8 && i < p.ra * p.ca; i += 128) {
    int r = i / p.ca, c = i % p.ca;
    const size_t off = p.a_layout ? sw_off(p.a_layout, a_ps, r + p.halo, c)
                                  : (size_t)(c / 8) * a_ps + (size_t)(r + p.halo) * 16 + (c % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sa + off) = __float2bfloat16(p.a[i]);
  }
  for (int i = tid; i < p.rb * p.cb; i += 128) {
    int r = i / p.cb, c = i % p.cb;
    const size_t off = p.b_layout ? sw_off(p.b_layout, b_ps, r, c) : (size_t)(c / 8) * b_ps + (size_t)r * 16 + (c % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sb + off) = __float2bfloat16(p.b[i]);
  }
  if (p.a_layout == 8) {
    // A -> TMEM columns [256 - K/2 .. 256): thread tid owns row tid
    const uint32_t tm = tmem_slot;
    for (int c0 = 0; c0 < p.K / 2; c0 += 8) {
      uint32_t v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tc::pack_bf16(p.a[(size_t)tid * p.ca + 2 * (c0 + j)], p.a[(size_t)tid * p.ca + 2 * (c0 + j) + 1]);
      tc::tmem_st8(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(256 - p.K / 2 + c0), v);
    }
    tc::tmem_st_wait();
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0 && p.a_layout == 8) {
    const uint32_t idesc = tc::make_idesc(128, p.N, 0, p.b_mn);
    const uint32_t bbase = tc::smem_u32(sb);
    for (int ks = 0; ks < p.K / 16; ++ks) {
      uint32_t b_addr, b_lbo, b_sbo;
      if (!p.b_mn) { b_addr = bbase + 2 * ks * b_ps; b_lbo = b_ps; b_sbo = 128; }
      else { b_addr = bbase + ks * 256; b_lbo = 128; b_sbo = b_ps; }
      const uint64_t db = tc::make_desc(b_addr, b_lbo, b_sbo);
      tc::umma_ts(tmem, tmem + (uint32_t)(256 - p.K / 2 + 8 * ks), (uint32_t)db, (uint32_t)(db >> 32), idesc, ks > 0);
    }
    tc::umma_commit(&bar);
  } else if (tid == 0) {
    const uint32_t idesc = tc::make_idesc(128, p.N, p.a_mn, p.b_mn);
    const uint32_t abase = tc::smem_u32(sa) + (uint32_t)(p.halo + p.a_shift) * 16, bbase = tc::smem_u32(sb);
    for (int ks = 0; ks < p.K / 16; ++ks) {
      uint32_t a_addr, a_lbo, a_sbo, b_addr, b_lbo, b_sbo;
      if (!p.a_mn) { a_addr = abase + 2 * ks * a_ps; a_lbo = a_ps; a_sbo = 128; }
      else { a_addr = abase + ks * 256; a_lbo = 128; a_sbo = a_ps; }
      if (!p.b_mn) { b_addr = bbase + 2 * ks * b_ps; b_lbo = b_ps; b_sbo = 128; }
      else { b_addr = bbase + ks * 256; b_lbo = 128; b_sbo = b_ps; }
      if (p.swap) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
      uint64_t da = tc::make_desc(a_addr, a_lbo, a_sbo), db = tc::make_desc(b_addr, b_lbo, b_sbo);
      if (p.a_layout) {
        const int rb = sw_rb(p.a_layout), kpb = rb / 32;        // k-steps of 16 elements per block
        uint32_t st;
        if (!p.a_mn) { st = tc::smem_u32(sa) + (uint32_t)(ks / kpb) * a_ps + (uint32_t)(p.halo + p.a_shift) * rb + (uint32_t)(ks % kpb) * 32; a_lbo = 16; a_sbo = 8 * rb; }
        else { st = tc::smem_u32(sa) + (uint32_t)(p.halo + p.a_shift + 16 * ks) * rb; a_lbo = a_ps; a_sbo = 8 * rb; }
        da = tc::make_desc_sw(st, a_lbo, a_sbo, p.a_layout, p.base_off_mode ? ((st >> 7) & 7u) : 0u);
      }
      if (p.b_layout) {
        const int rb = sw_rb(p.b_layout), kpb = rb / 32;
        uint32_t st;
        if (!p.b_mn) { st = tc::smem_u32(sb) + (uint32_t)(ks / kpb) * b_ps + (uint32_t)(ks % kpb) * 32; b_lbo = 16; b_sbo = 8 * rb; }
        else { st = tc::smem_u32(sb) + (uint32_t)(16 * ks) * rb; b_lbo = b_ps; b_sbo = 8 * rb; }
        db = tc::make_desc_sw(st, b_lbo, b_sbo, p.b_layout, p.base_off_mode ? ((st >> 7) & 7u) : 0u);
      }
      tc::umma(tmem, da, db, idesc, ks > 0);
    }
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    uint32_t v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) p.d[(size_t)tid * p.N + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace mr

extern "C" {

int mr_tc_selftest(const float* a, int64_t ra, int64_t ca, const float* b, int64_t rb, int64_t cb, float* d, int a_mn,
                   int b_mn, int64_t N, int64_t K, int64_t a_shift, int64_t halo, int swap, int a_layout, int b_layout,
                   int base_off_mode, void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  MR_REQUIRE(a && b && d, MR_ERR_NULL, "mr_tc_selftest: null pointer");
  MR_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16, MR_ERR_BAD_SHAPE, "mr_tc_selftest: N=%lld K=%lld",
             (long long)N, (long long)K);
  MR_REQUIRE(ca % 8 == 0 && cb % 8 == 0 && halo >= 0 && a_shift >= -halo && a_shift <= halo, MR_ERR_BAD_SHAPE,
             "mr_tc_selftest: bad operand shapes");
  if (!a_mn) MR_REQUIRE(ra == 128 && ca >= K, MR_ERR_BAD_SHAPE, "A (K-major) must be [128, >=K]");
  else MR_REQUIRE(ca == 128 && ra >= K, MR_ERR_BAD_SHAPE, "A (MN-major) must be [>=K, 128]");
  if (!b_mn) MR_REQUIRE(rb >= N && cb >= K, MR_ERR_BAD_SHAPE, "B (K-major) must be [>=N, >=K]");
  else MR_REQUIRE(cb >= N && rb >= K, MR_ERR_BAD_SHAPE, "B (MN-major) must be [>=K, >=N]");
  SelfTestArgs p{a, (int)ra, (int)ca, b, (int)rb, (int)cb, d, a_mn, b_mn, (int)N, (int)K, (int)a_shift, (int)halo, swap,
                 a_layout, b_layout, base_off_mode};
  auto okl = [](int l) { return l == 0 || l == 2 || l == 4 || l == 6; };
  MR_REQUIRE((okl(a_layout) || a_layout == 8) && okl(b_layout), MR_ERR_BAD_SHAPE, "mr_tc_selftest: layouts must be 0, 2, 4 or 6 (A: or 8 = tensor memory)");
  if (a_layout == 8) MR_REQUIRE(!a_mn && b_layout == 0 && a_shift == 0 && N + K / 2 <= 256 && K % 16 == 0, MR_ERR_BAD_SHAPE, "mr_tc_selftest: A from tensor memory needs a K-major A, N + K/2 <= 256");
  auto opbytes = [](int layout, int64_t rows, int64_t cols) -> size_t {
    if (!layout) return (size_t)(cols / 8) * ((rows | 1) * 16);
    const int64_t rbytes = layout == 2 ? 128 : (layout == 4 ? 64 : 32);
    return (size_t)((cols * 2 + rbytes - 1) / rbytes) * (size_t)((rows + 7) / 8 * 8) * rbytes;
  };
  size_t smem = (opbytes(a_layout == 8 ? 0 : a_layout, ra + 2 * halo, ca) + 1023) / 1024 * 1024 + opbytes(b_layout, rb, cb) + 1024;
  MR_REQUIRE(smem <= 220 * 1024, MR_ERR_BAD_SHAPE, "mr_tc_selftest: operands need %zu bytes of shared memory", smem);
  cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(p);
  MR_CHECK_LAUNCH("tc_selftest_kernel");
  return MR_OK;
}

}  // extern "C"

// ---- micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, bf16) as a function of N, operands fixed in smem ----
namespace mr {
__global__ void __launch_bounds__(128, 1) tc_mma_rate_kernel(int N, int iters, int b_layout, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid * 16; i < 64 * 1024; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc(128, N, 0, 0);
    const uint32_t a = tc::smem_u32(smem), b = a + 32 * 1024;
    const uint64_t da = tc::make_desc_sw(a, 16, 1024, 2, 0);
    const uint64_t db = b_layout == 2 ? tc::make_desc_sw(b, 16, 1024, 2, 0) : tc::make_desc(b, (uint32_t)N * 16, 128);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) tc::umma(tmem, da, db, idesc, 1);
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    out[0] = clock64() - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}
}  // namespace mr

extern "C" __attribute__((visibility("default"))) int mr_debug_mma_rate(int N, int iters, int b_layout, long long* out_device, void* stream) {
  using namespace mr;
  if (int rc = require_sm100()) return rc;
  cudaFuncSetAttribute(tc_mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  tc_mma_rate_kernel<<<1, 128, 64 * 1024, as_stream(stream)>>>(N, iters, b_layout, out_device);
  MR_CHECK_LAUNCH("tc_mma_rate_kernel");
  return MR_OK;
}
