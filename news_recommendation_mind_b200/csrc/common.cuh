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

// ---- programmatic dependent launch ------------------------------------------------------------
// Kernels of the hot chain are launched with the programmatic-stream-serialization attribute: the next kernel's CTAs may
// become resident (and run their prologue: shared-memory clearing, barrier init, TMEM allocation, descriptor prefetch) as
// soon as the previous kernel's CTAs have retired from an SM, instead of after the whole grid has drained plus a launch
// latency.  Every such kernel executes pdl_trigger() at its top and pdl_wait() before its first global-memory access;
// pdl_wait() returns only when the preceding grid has completed and its writes are visible, so the ordering seen by
// the data is exactly that of ordinary stream serialization.  MINDREC_PDL=0 launches without the attribute.
bool pdl_enabled();                    // runtime.cu
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// ptxas treats ld.global.nc (what `const T* __restrict__` loads compile to) as movable across griddepcontrol.wait -- the "memory"
// clobber above does not order non-coherent loads -- and did hoist the first W_hh loads of rnn_mma_fwd_kernel above the wait,
// i.e. above the completion of the kernel that WRITES that buffer (seen as an intermittent wrong LSTM at >= 4 sequences per
// CTA).  Laundering the pointer through an asm volatile placed after the wait makes every address derived from it, and so
// every load, depend on an instruction that cannot move above the wait.  Use it for every buffer the preceding kernel wrote
// and this kernel reads through a `const __restrict__` pointer early in its prologue.
template <class T>
__device__ __forceinline__ T* pdl_acquire(T* p) {
  asm volatile("" : "+l"(p)::"memory");
  return p;
}

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
