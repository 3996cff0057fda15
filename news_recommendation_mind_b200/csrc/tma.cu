// Host-side CUtensorMap construction (cuTensorMapEncodeTiled fetched with cudaGetDriverEntryPoint).
#include "tma.cuh"

namespace mr {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

static CUtensorMapSwizzle swz(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                      : (bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE));
}

int tma_encode_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t row_pitch_bytes, uint32_t box_cols,
                  uint32_t box_rows, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  MR_REQUIRE(fn != nullptr, MR_ERR_LAUNCH, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MR_REQUIRE(r == CUDA_SUCCESS, MR_ERR_LAUNCH, "cuTensorMapEncodeTiled(2d) failed with %d (cols %llu rows %llu pitch %llu box %u x %u)",
             (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_pitch_bytes, box_cols, box_rows);
  return MR_OK;
}

int tma_encode_3d(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes, uint64_t s2_bytes,
                  uint32_t b0, uint32_t b1, uint32_t b2, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  MR_REQUIRE(fn != nullptr, MR_ERR_LAUNCH, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MR_REQUIRE(r == CUDA_SUCCESS, MR_ERR_LAUNCH,
             "cuTensorMapEncodeTiled(3d) failed with %d (dims %llu %llu %llu strides %llu %llu box %u %u %u)", (int)r,
             (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1_bytes,
             (unsigned long long)s2_bytes, b0, b1, b2);
  return MR_OK;
}

}  // namespace mr
