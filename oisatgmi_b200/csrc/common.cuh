// Shared helpers for liboisat (sm_100a).  Compiled with --fmad=false: every
// a*b+c in this tree is two IEEE roundings unless written as fma(), because
// several kernels reproduce numpy/scipy results bit for bit.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/oisat.h"

namespace oisat {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define OISAT_CHECK_ARG(cond, msg)                         \
  do {                                                     \
    if (!(cond)) {                                         \
      ::oisat::set_error("%s: %s", __func__, msg);         \
      return OISAT_E_ARG;                                  \
    }                                                      \
  } while (0)

#define OISAT_CHECK_LAUNCH()                                                     \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      ::oisat::set_error("%s: CUDA launch failed: %s", __func__,                 \
                         cudaGetErrorString(e__));                               \
      return OISAT_E_CUDA;                                                       \
    }                                                                            \
    ::oisat::count_launch();                                                     \
  } while (0)

#define OISAT_CHECK_CUDA(expr)                                                   \
  do {                                                                           \
    cudaError_t e__ = (expr);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      ::oisat::set_error("%s: %s failed: %s", __func__, #expr,                   \
                         cudaGetErrorString(e__));                               \
      return OISAT_E_CUDA;                                                       \
    }                                                                            \
  } while (0)

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ double qnan() { return CUDART_NAN; }

// exact up-conversion of one element of a reader-typed array
__device__ __forceinline__ double load_as_double(const void* p, int dtype, int64_t i) {
  switch (dtype) {
    case OISAT_F16: return (double)__half2float(((const __half*)p)[i]);
    case OISAT_F32: return (double)((const float*)p)[i];
    default: return ((const double*)p)[i];
  }
}

// x**2 evaluated in the array's own dtype (numpy semantics: float16 is
// computed in float32 and rounded back to float16), then widened exactly.
__device__ __forceinline__ double load_square_native(const void* p, int dtype, int64_t i) {
  switch (dtype) {
    case OISAT_F16: {
      float v = __half2float(((const __half*)p)[i]);
      return (double)__half2float(__float2half_rn(__fmul_rn(v, v)));
    }
    case OISAT_F32: {
      float v = ((const float*)p)[i];
      return (double)__fmul_rn(v, v);
    }
    default: {
      double v = ((const double*)p)[i];
      return __dmul_rn(v, v);
    }
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace oisat
