// Optional per-launch timing of the encoder's big kernels with CUDA events on the launching stream (bench hooks).
#pragma once
#include "common.cuh"

namespace mr {

constexpr int KT_KERNELS = 3;
constexpr int CONV_EVT_SLOTS = 256;
struct KernelTiming {
  bool enabled = false;
  cudaEvent_t beg[KT_KERNELS][CONV_EVT_SLOTS], end[KT_KERNELS][CONV_EVT_SLOTS];
  bool created = false;
  int64_t count[KT_KERNELS] = {0, 0, 0};
};
extern KernelTiming g_kt;
struct TimedLaunch {              // records the two events around a launch sequence (scope)
  int which, slot;
  cudaStream_t st;
  bool on;
  TimedLaunch(int w, cudaStream_t s) : which(w), slot(0), st(s), on(g_kt.enabled) {
    if (on) {
      slot = (int)(g_kt.count[w] % CONV_EVT_SLOTS);
      cudaEventRecord(g_kt.beg[w][slot], st);
    }
  }
  ~TimedLaunch() {
    if (on) {
      cudaEventRecord(g_kt.end[which][slot], st);
      ++g_kt.count[which];
    }
  }
};


}  // namespace mr
