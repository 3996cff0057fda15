// TMA (cp.async.bulk.tensor) helpers: host-side tensor-map construction through the driver entry point
// (no link-time dependency on libcuda) and the device-side issue wrappers used by the producers.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "tc05.cuh"

namespace mr {

// bf16 tensor maps.  rank 2: dims {cols, rows}, row pitch in bytes; rank 3: dims {d0, d1, d2} with byte strides
// s1, s2 of dims 1 and 2.  swizzle_bytes in {0, 32, 64, 128}; box0 * 2 bytes must not exceed swizzle_bytes.
int tma_encode_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t row_pitch_bytes, uint32_t box_cols,
                  uint32_t box_rows, int swizzle_bytes);
int tma_encode_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes, uint64_t s2_bytes,
                  uint32_t b0, uint32_t b1, uint32_t b2, int swizzle_bytes);

namespace tc {

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// 3-D tile load, completion on an mbarrier (bytes of the full box, out-of-bounds elements are zero filled)
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst_smem),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// gather of 4 rows (r0..r3) x box columns starting at column c0 of a 2-D tensor; the rows land consecutively
__device__ __forceinline__ void tma_gather4(uint32_t dst_smem, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          dst_smem),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

}  // namespace tc
}  // namespace mr
