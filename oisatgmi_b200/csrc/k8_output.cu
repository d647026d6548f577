// K8: the data side of driver.write_to_nc (driver.py:156-227):
// the nine float32 fields the reference stores per month -- the means, the prior and
// posterior model column, the OI diagnostics and the emission scaling factor
//     posterior / prior, with NaN, +-inf and 0 replaced by 1        (driver.py:203-206)
// -- produced in one pass on the device, so that the month's device -> host copy is
// float32 (half the bytes) and already in file layout.  Writing the NetCDF container
// itself is file I/O and stays with the caller (netCDF4 is the reference's dependency).
#include "common.cuh"

namespace oisat {

struct OutputArgs {
  const double* src[8];   // sat_vcd, ctm_prior, ctm_posterior, sat_error, ak, error_OI, aux1, aux2
  float* out;             // [9][n]: the eight above in this order with scaling_factor at row 6
  int64_t n;
};

__global__ void __launch_bounds__(256)
output_fields_kernel(const __grid_constant__ OutputArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const double prior = a.src[1][i], post = a.src[2][i];
  double s = post / prior;                       // float64 division, like numpy
  if (s != s || isinf(s) || s == 0.0) s = 1.0;
  const int row_of[8] = {0, 1, 2, 3, 4, 5, 7, 8};
#pragma unroll
  for (int k = 0; k < 8; ++k) a.out[(int64_t)row_of[k] * a.n + i] = (float)a.src[k][i];
  a.out[6 * a.n + i] = (float)s;
}

}  // namespace oisat

extern "C" int oisat_output_fields(int64_t n, const double* sat_vcd, const double* ctm_prior,
                                   const double* ctm_posterior, const double* sat_error,
                                   const double* ak, const double* error_oi, const double* aux1,
                                   const double* aux2, float* out, void* stream) {
  using namespace oisat;
  if (n <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(sat_vcd && ctm_prior && ctm_posterior && sat_error && ak && error_oi && aux1 &&
                      aux2 && out, "null pointer");
  OutputArgs a;
  const double* src[8] = {sat_vcd, ctm_prior, ctm_posterior, sat_error, ak, error_oi, aux1, aux2};
  for (int k = 0; k < 8; ++k) a.src[k] = src[k];
  a.out = out;
  a.n = n;
  output_fields_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(a);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
