// Shared host/device helpers for libgte_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "gte.h"

namespace gte {

// thread-local error text behind gte_last_error_string()
char* err_buf();
int fail(int code, const char* fmt, ...);

inline cudaStream_t as_stream(gte_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// SM count of the current device (cached per device; immutable after first query)
int sm_count();

// number of kernels launched by this library since load (diagnostic; gte_launch_count())
void note_launch();

#define GTE_CHECK_ARG(cond, ...)                                  \
  do {                                                            \
    if (!(cond)) return ::gte::fail(GTE_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define GTE_CHECK_LAUNCH(name)                                                            \
  do {                                                                                    \
    ::gte::note_launch();                                                                 \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess)                                                               \
      return ::gte::fail(GTE_ERR_CUDA, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
  } while (0)

#define GTE_CHECK_CUDA(call, name)                                                  \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess)                                                         \
      return ::gte::fail(GTE_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__));    \
  } while (0)

// ---------------------------------------------------------------- device --
#ifdef __CUDACC__
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif

}  // namespace gte
