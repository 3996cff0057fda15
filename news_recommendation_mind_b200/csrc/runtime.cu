// Library-wide runtime bits: error channel, device gate, launch counter.
#include "common.cuh"
#include <stdlib.h>

namespace mr {

std::atomic<int64_t> g_launches{0};

char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MINDREC_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int require_sm100() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return set_err(MR_ERR_NOT_SM100, "no CUDA device available (libmindrec has no CPU fallback)");
  }
  static thread_local int cached_dev = -2, cached_rc = 0;
  if (dev == cached_dev) return cached_rc;
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cached_dev = dev;
  cached_rc = (major == 10) ? MR_OK
                            : set_err(MR_ERR_NOT_SM100, "device %d is sm_%d0, libmindrec is built for sm_100a only",
                                      dev, major);
  return cached_rc;
}

}  // namespace mr

extern "C" {

int mr_version(void) { return 100; }

const char* mr_last_error(void) { return mr::err_buf(); }

int64_t mr_launch_count(void) { return mr::g_launches.load(); }

int mr_device_check(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return mr::set_err(MR_ERR_NOT_SM100, "CUDA device %d not present", device);
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (major != 10) return mr::set_err(MR_ERR_NOT_SM100, "device %d is compute capability %d.x, need 10.x", device, major);
  return MR_OK;
}

}  // extern "C"
