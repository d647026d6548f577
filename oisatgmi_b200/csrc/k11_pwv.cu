// SSMIS precipitable-water path.
//
//   oisat_reader_ssmis   reader.py:1292-1297 (ssmis_reader_wv): the scaled map -> float32 water
//                        vapour (flag values above 250 -> NaN, x 0.3, >= 75 or inf -> NaN) and
//                        its 5 % uncertainty, every step in float32 like numpy.
//   oisat_pwv_partial    pwv_cal.py:62,68: delta_p * q / g / 10000 per model layer, float32
//                        arithmetic step by step (float32 arrays and Python floats, NEP 50).
//   oisat_pwv_column     pwv_cal.py:91-93: nansum over the layers of (partial / 1000), in the
//                        partial column's own dtype (float32 on the model grid, float64 after
//                        the model has been resampled to a satellite mesh), layer by layer in
//                        order -- numpy reduces axis 0 of a C-ordered block sequentially --
//                        NaN where the satellite map has no value (NaN or inf).
#include "common.cuh"

namespace oisat {

__global__ void __launch_bounds__(256)
rd_ssmis_kernel(const void* __restrict__ src, int dtype, int64_t n, float* __restrict__ pwv,
                float* __restrict__ unc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v;
  if (dtype == OISAT_U8) v = (float)reinterpret_cast<const uint8_t*>(src)[i];
  else if (dtype == OISAT_I32) v = (float)reinterpret_cast<const int32_t*>(src)[i];
  else v = (float)load_as_double(src, dtype, i);
  if (v > 250.0f) v = CUDART_NAN_F;
  v = __fmul_rn(v, 0.3f);
  if (v >= 75.0f || isinf(v)) v = CUDART_NAN_F;
  pwv[i] = v;
  unc[i] = __fmul_rn(v, 0.05f);
}

__global__ void __launch_bounds__(256)
pwv_partial_kernel(const float* __restrict__ dp, const float* __restrict__ q, int64_t n,
                   float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = __fdiv_rn(__fdiv_rn(__fmul_rn(dp[i], q[i]), 9.80665f), 10000.0f);
}

template <typename T>
__global__ void __launch_bounds__(256)
pwv_column_kernel(const T* __restrict__ pc, int n_lev, int64_t n, const double* __restrict__ vcd,
                  double* __restrict__ out) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  T acc = (T)0;
  for (int k = 0; k < n_lev; ++k) {
    const T v = pc[(int64_t)k * n + c] / (T)1000.0;
    if (v == v) acc = acc + v;
  }
  const double s = vcd[c];
  out[c] = (s != s || isinf(s)) ? qnan() : (double)acc;
}

}  // namespace oisat

using namespace oisat;

extern "C" int oisat_reader_ssmis(const void* src, int32_t dtype, int64_t n, float* pwv,
                                  float* uncertainty, void* stream) {
  if (n <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(src && pwv && uncertainty, "null pointer");
  OISAT_CHECK_ARG(dtype == OISAT_U8 || dtype == OISAT_I32 || dtype == OISAT_F16 ||
                      dtype == OISAT_F32 || dtype == OISAT_F64, "bad dtype");
  rd_ssmis_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(src, dtype, n, pwv,
                                                                              uncertainty);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_pwv_partial(const float* delta_p, const float* profile, int64_t n, float* out,
                                 void* stream) {
  if (n <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(delta_p && profile && out, "null pointer");
  pwv_partial_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(delta_p, profile,
                                                                                 n, out);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_pwv_column(const void* partial, int32_t dtype, int32_t n_lev, int64_t n_cell,
                                const double* sat_vcd, double* out, void* stream) {
  if (n_cell <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(partial && sat_vcd && out && n_lev >= 1, "bad arguments");
  OISAT_CHECK_ARG(dtype == OISAT_F32 || dtype == OISAT_F64, "partial column must be float32 or float64");
  const unsigned blocks = (unsigned)ceil_div(n_cell, 256);
  if (dtype == OISAT_F32)
    pwv_column_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)partial, n_lev,
                                                                      n_cell, sat_vcd, out);
  else
    pwv_column_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double*)partial, n_lev,
                                                                       n_cell, sat_vcd, out);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
