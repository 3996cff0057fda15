// Shared helpers for libmindrec.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/mindrec.h"

namespace mr {

// ---- error channel: thread-local message, integer status across the C ABI -------------------
char* err_buf();                       // defined in runtime.cu
int set_err(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define MR_REQUIRE(cond, code, ...)                       \
  do {                                                    \
    if (!(cond)) return ::mr::set_err((code), __VA_ARGS__); \
  } while (0)

#define MR_CHECK_LAUNCH(name)                                                          \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess)                                                            \
      return ::mr::set_err(MR_ERR_LAUNCH, "%s: %s", (name), cudaGetErrorString(e__)); \
    ::mr::count_launch();                                                              \
  } while (0)

int require_sm100();                   // MR_OK or MR_ERR_NOT_SM100 for the current device (cached)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t align_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// bump allocator over the caller's workspace
struct Arena {
  char* base;
  int64_t cap, off;
  Arena(void* p, int64_t n) : base(static_cast<char*>(p)), cap(n), off(0) {}
  template <class T>
  T* take(int64_t count) {
    int64_t bytes = align_up(count * (int64_t)sizeof(T), 256);
    if (base == nullptr || off + bytes > cap) { off = cap + 1; return nullptr; }
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= cap; }
};
inline int64_t arena_bytes(int64_t count, int64_t elt) { return align_up(count * elt, 256); }

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ int64_t load_index(const void* p, int is64, int64_t i) {
  return is64 ? static_cast<const int64_t*>(p)[i] : (int64_t) static_cast<const int32_t*>(p)[i];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace mr
