// K4: temporal accumulation of gridded granules (averaging.py:64-108, 11-24).
//
// The reference stacks every granule grid of the month in RAM and calls
// np.nanmean over the granule axis (a sequential axis-0 add), and loops over
// cells in Python for the error.  Here the month is a [10][n_cell] float64
// block of running sums and exact counts: constant memory in the number of
// granules, one all-reduce to merge ranks, and -- because granules are added in
// list order -- the same floating-point result as numpy's sequential reduction.
#include "common.cuh"

namespace oisat {

__device__ __forceinline__ void add_value(double* sum, double* cnt, double v) {
  if (v == v) {  // nanmean: NaN entries contribute 0 to the sum and nothing to the count
    *sum = __dadd_rn(*sum, v);
    *cnt = *cnt + 1.0;
  }
}

__global__ void __launch_bounds__(256)
accum_add_kernel(double* __restrict__ acc, int64_t n, const double* __restrict__ vcd,
                 const double* __restrict__ sigma, const double* __restrict__ ctm_vcd,
                 const double* __restrict__ aux1, const double* __restrict__ aux2,
                 int sigma_is_variance) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  if (vcd) {
    double v = vcd[c];
    if (isinf(v)) v = qnan();  // averaging.py:92
    add_value(&acc[0 * n + c], &acc[5 * n + c], v);
  }
  if (sigma) {
    const double s = sigma[c];
    // averaging.py:101 sat_chosen_error**2 -- squared here in float64, or handed over already
    // squared in the array's own dtype when that is narrower (numpy squares in the native dtype)
    double v = sigma_is_variance ? s : __dmul_rn(s, s);
    if (isinf(v)) v = qnan();    // averaging.py:19
    add_value(&acc[1 * n + c], &acc[6 * n + c], v);
  }
  if (ctm_vcd) add_value(&acc[2 * n + c], &acc[7 * n + c], ctm_vcd[c]);
  if (aux1) add_value(&acc[3 * n + c], &acc[8 * n + c], aux1[c]);
  if (aux2) add_value(&acc[4 * n + c], &acc[9 * n + c], aux2[c]);
}

__global__ void __launch_bounds__(256)
accum_finalize_kernel(const double* __restrict__ acc, int64_t n, double* __restrict__ sat_vcd,
                      double* __restrict__ sat_err, double* __restrict__ ctm_vcd,
                      double* __restrict__ aux1, double* __restrict__ aux2) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  // 0/0 = NaN is exactly what np.nanmean returns for an all-NaN column
  if (sat_vcd) sat_vcd[c] = acc[0 * n + c] / acc[5 * n + c];
  if (sat_err) {
    const double k = acc[6 * n + c];
    sat_err[c] = sqrt(acc[1 * n + c] / (k * k));  // averaging.py:21,23
  }
  if (ctm_vcd) ctm_vcd[c] = acc[2 * n + c] / acc[7 * n + c];
  if (aux1) aux1[c] = acc[3 * n + c] / acc[8 * n + c];
  if (aux2) aux2[c] = acc[4 * n + c] / acc[9 * n + c];
}

// Ordered segmented reduction: thread = model cell, walks its pair list (granule
// order) and adds the five staged values.  Deterministic, no atomics.
__global__ void __launch_bounds__(256)
accum_pairs_kernel(double* __restrict__ acc, int64_t n, const int64_t* __restrict__ seg_start,
                   const int64_t* __restrict__ seg_pair, const double* __restrict__ staged,
                   int64_t n_pairs) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int64_t b = seg_start[c], e = seg_start[c + 1];
  if (b == e) return;
  double s[5], k[5];
#pragma unroll
  for (int q = 0; q < 5; ++q) { s[q] = acc[q * n + c]; k[q] = acc[(5 + q) * n + c]; }
  for (int64_t i = b; i < e; ++i) {
    const int64_t p = seg_pair[i];
    double v0 = staged[0 * n_pairs + p];
    if (isinf(v0)) v0 = qnan();
    add_value(&s[0], &k[0], v0);
    const double sg = staged[1 * n_pairs + p];
    double v1 = __dmul_rn(sg, sg);
    if (isinf(v1)) v1 = qnan();
    add_value(&s[1], &k[1], v1);
    add_value(&s[2], &k[2], staged[2 * n_pairs + p]);
    add_value(&s[3], &k[3], staged[3 * n_pairs + p]);
    add_value(&s[4], &k[4], staged[4 * n_pairs + p]);
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) { acc[q * n + c] = s[q]; acc[(5 + q) * n + c] = k[q]; }
}

}  // namespace oisat

using namespace oisat;

extern "C" int oisat_accum_add(double* acc, int64_t n_cell, const double* vcd,
                               const double* sigma, const double* ctm_vcd, const double* aux1,
                               const double* aux2, void* stream) {
  OISAT_CHECK_ARG(acc && n_cell >= 0, "bad accumulator");
  if (n_cell == 0) return OISAT_OK;
  accum_add_kernel<<<(unsigned)ceil_div(n_cell, 256), 256, 0, (cudaStream_t)stream>>>(
      acc, n_cell, vcd, sigma, ctm_vcd, aux1, aux2, 0);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_accum_add_variance(double* acc, int64_t n_cell, const double* vcd,
                                        const double* variance, const double* ctm_vcd,
                                        const double* aux1, const double* aux2, void* stream) {
  OISAT_CHECK_ARG(acc && n_cell >= 0, "bad accumulator");
  if (n_cell == 0) return OISAT_OK;
  accum_add_kernel<<<(unsigned)ceil_div(n_cell, 256), 256, 0, (cudaStream_t)stream>>>(
      acc, n_cell, vcd, variance, ctm_vcd, aux1, aux2, 1);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_accum_finalize(const double* acc, int64_t n_cell, double* sat_vcd,
                                    double* sat_err, double* ctm_vcd, double* aux1, double* aux2,
                                    void* stream) {
  OISAT_CHECK_ARG(acc && n_cell >= 0, "bad accumulator");
  if (n_cell == 0) return OISAT_OK;
  accum_finalize_kernel<<<(unsigned)ceil_div(n_cell, 256), 256, 0, (cudaStream_t)stream>>>(
      acc, n_cell, sat_vcd, sat_err, ctm_vcd, aux1, aux2);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_accum_pairs(double* acc, int64_t n_cell, const int64_t* seg_start,
                                 const int64_t* seg_pair, const double* staged, int64_t n_pairs,
                                 void* stream) {
  OISAT_CHECK_ARG(acc && seg_start && n_cell >= 0, "bad accumulator");
  if (n_cell == 0 || n_pairs == 0) return OISAT_OK;
  OISAT_CHECK_ARG(seg_pair && staged, "null pointer");
  accum_pairs_kernel<<<(unsigned)ceil_div(n_cell, 256), 256, 0, (cudaStream_t)stream>>>(
      acc, n_cell, seg_start, seg_pair, staged, n_pairs);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
