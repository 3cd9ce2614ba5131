// Shared helpers for libnsc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/nsc_b200.h"

namespace nsc {

void set_error(const char* fmt, ...);

#define NSC_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      nsc::set_error(__VA_ARGS__);               \
      return NSC_E_INVALID;                      \
    }                                            \
  } while (0)

#define NSC_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      nsc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NSC_E_CUDA;                                                                          \
    }                                                                                             \
  } while (0)

#define NSC_LAUNCH_OK()                                                                       \
  do {                                                                                        \
    nsc::count_launch();                                                                      \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      nsc::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NSC_E_CUDA;                                                                      \
    }                                                                                         \
  } while (0)

#define NSC_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != NSC_OK) return _r; \
  } while (0)

constexpr float kLeakySlope = 0.2f;  // tf.nn.leaky_relu default (nn_core_operator.py:30)

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == NSC_ACT_TANH) return tanhf(v);
  if (act == NSC_ACT_LRELU) return v > 0.f ? v : kLeakySlope * v;
  return v;
}

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// TF SAME padding: out = ceil(L/s); pad = max((out-1)s + (k-1)d + 1 - L, 0); left = pad/2.
__host__ __device__ inline void same_padding(int L, int k, int d, int s, int* out, int* left) {
  int o = (L + s - 1) / s;
  int total = (o - 1) * s + (k - 1) * d + 1 - L;
  if (total < 0) total = 0;
  *out = o;
  *left = total / 2;
}

int sm_count();
void count_launch();

// Optional per-launch timing (nsc_profile_begin / nsc_profile_end): a ProfScope around a launch records a
// CUDA-event pair on the launching stream plus the launch's ALGORITHMIC flops and bytes.  Inactive = free.
struct ProfScope {
  ProfScope(cudaStream_t st, const char* name, double flops, double bytes);
  ~ProfScope();
  int slot;
  cudaStream_t st;
};

}  // namespace nsc
