// Warp-level building blocks of the per-cell vertical operators (K3).  One warp
// works on one model cell; per-warp scratch lives in shared memory.  The same
// routines serve the stand-alone kernels (k3_vertical.cu) and the fused month
// pipeline (fused_amf.cu), which is what keeps the two paths results-identical.
//
// Everything that numpy/scipy evaluate in a fixed order is evaluated in that
// order here (no FMA contraction in this tree, see common.cuh):
//   * np.sum / np.nansum of a contiguous 1-D array: pairwise summation with
//     eight running partial sums for n <= 128 (numpy loops_utils: *_pairwise_sum);
//   * scipy interp1d(kind='linear', assume_sorted=False): stable argsort of x,
//     searchsorted(side='left'), indices clipped to [1, n-1], and
//     y = ((x-x_lo)/(x_hi-x_lo))*y_hi + ((x_hi-x)/(x_hi-x_lo))*y_lo
//     (scipy/interpolate/_interpolate.py: interp1d._call_linear);
//   * numpy.interp for the float64 non-extrapolating case interp1d delegates to.
#pragma once

#include "common.cuh"

namespace oisat {

constexpr int kMaxSatLev = 96;   // TEMPO has 72 (reader.py:502-512)
constexpr int kMaxCtmLev = 128;  // numpy's pairwise block: one leaf

constexpr float kG0f = 9.80665f;
constexpr float kMairF = (float)28.97e-3;
constexpr float kNaF = (float)6.02214076e23;
constexpr float k1em4f = (float)1e-4;
constexpr float k1em15f = (float)1e-15;
constexpr float k1em9f = (float)1e-9;

// amf_recal.py:51-56 in float32, left to right
__device__ __forceinline__ float partial_column_f32(float dp, float x) {
  float v = __fmul_rn(dp, x);
  v = __fdiv_rn(v, kG0f);
  v = __fdiv_rn(v, kMairF);
  v = __fmul_rn(v, kNaF);
  v = __fmul_rn(v, k1em4f);
  v = __fmul_rn(v, k1em15f);
  v = __fmul_rn(v, 100.0f);
  v = __fmul_rn(v, k1em9f);
  return v;
}

// ak_conv_mopitt.py:68 in float32
__device__ __forceinline__ float air_column_f32(float dp) {
  float v = __fdiv_rn(dp, kG0f);
  v = __fdiv_rn(v, kMairF);
  v = __fmul_rn(v, kNaF);
  v = __fmul_rn(v, k1em4f);
  v = __fmul_rn(v, k1em15f);
  v = __fmul_rn(v, 100.0f);
  return v;
}

__device__ __forceinline__ bool nan_less(double a, double b) {
  // numpy's sort/searchsorted order: NaN is larger than everything
  return a < b || (b != b && a == a);
}

// Pairwise sum (numpy order) of vals[0..n), n <= 128, executed by a full warp.
// NaNs must already be replaced (nansum) by the caller.  Result on all lanes.
template <typename T>
__device__ __forceinline__ T warp_np_sum(const T* vals, int n, int lane) {
  T res;
  if (n < 8) {
    res = (T)0;
    for (int i = 0; i < n; ++i) res = res + vals[i];
    return res;
  }
  const int body = n - (n % 8);
  T r = (T)0;
  if (lane < 8) {
    r = vals[lane];
    for (int i = 8 + lane; i < body; i += 8) r = r + vals[i];
  }
  r = r + __shfl_xor_sync(0xffffffffu, r, 1);
  r = r + __shfl_xor_sync(0xffffffffu, r, 2);
  r = r + __shfl_xor_sync(0xffffffffu, r, 4);
  res = __shfl_sync(0xffffffffu, r, 0);
  for (int i = body; i < n; ++i) res = res + vals[i];
  return res;
}

// Stable ascending argsort of (x, y) pairs held in shared memory, NaN last
// (np.argsort(kind='mergesort')), written to xs/ys.  Fast paths for the two
// monotone cases every real profile falls into.
__device__ __forceinline__ void warp_sort_levels(const double* xr, const double* yr, int n,
                                                 double* xs, double* ys, int lane) {
  bool inc = true, dec = true;
  for (int i = lane; i + 1 < n; i += 32) {
    const double a = xr[i], b = xr[i + 1];
    inc = inc && (a <= b);  // equal keys keep their order under a stable sort
    dec = dec && (a > b);
  }
  inc = __all_sync(0xffffffffu, inc);
  dec = __all_sync(0xffffffffu, dec);
  if (inc) {
    for (int i = lane; i < n; i += 32) { xs[i] = xr[i]; ys[i] = yr[i]; }
  } else if (dec) {
    for (int i = lane; i < n; i += 32) { xs[n - 1 - i] = xr[i]; ys[n - 1 - i] = yr[i]; }
  } else {
    for (int i = lane; i < n; i += 32) {
      const double xi = xr[i];
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const double xj = xr[j];
        const bool eq = (xj == xi) || (xj != xj && xi != xi);
        rank += (nan_less(xj, xi) || (eq && j < i)) ? 1 : 0;
      }
      xs[rank] = xi;
      ys[rank] = yr[i];
    }
  }
  __syncwarp();
}

// np.searchsorted(xs, v, side='left')
__device__ __forceinline__ int searchsorted_left(const double* xs, int n, double v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (nan_less(xs[mid], v)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// interp1d._call_linear (linear extrapolation falls out of the index clipping)
__device__ __forceinline__ double interp1d_linear(const double* xs, const double* ys, int n,
                                                  double v) {
  int idx = searchsorted_left(xs, n, v);
  idx = idx < 1 ? 1 : idx;
  idx = idx > n - 1 ? n - 1 : idx;
  const double x_lo = xs[idx - 1], x_hi = xs[idx];
  const double y_lo = ys[idx - 1], y_hi = ys[idx];
  const double den = x_hi - x_lo;
  return ((v - x_lo) / den) * y_hi + ((x_hi - v) / den) * y_lo;
}

// numpy.interp (arr_interp in numpy/_core/src/multiarray/compiled_base.c) followed by
// interp1d._check_bounds with fill_value = NaN
__device__ __forceinline__ double np_interp_nanfill(const double* xp, const double* fp, int n,
                                                    double x) {
  if (x != x) return x;
  if (x < xp[0] || x > xp[n - 1]) return qnan();
  // j with xp[j] <= x < xp[j+1]; x == xp[n-1] gives j = n-1
  int lo = 0, hi = n;  // count of xp[i] <= x
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (xp[mid] <= x) lo = mid + 1; else hi = mid;
  }
  const int j = lo - 1;
  if (j >= n - 1) return fp[n - 1];
  if (xp[j] == x) return fp[j];
  const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
  double r = slope * (x - xp[j]) + fp[j];
  if (r != r) {
    r = slope * (x - xp[j + 1]) + fp[j + 1];
    if (r != r && fp[j] == fp[j + 1]) r = fp[j];
  }
  return r;
}

// float32 natural log as numpy evaluates np.log on a float32 array (the result
// is float32).  CUDA's logf is within 1 ulp; numpy's SIMD kernel within ~4 ulp;
// the two agree to float32 rounding noise, which is the reference's own noise
// floor for this term.
__device__ __forceinline__ double log_as_f32(float p) { return (double)logf(p); }

struct WarpScratch {
  double xr[kMaxSatLev > kMaxCtmLev ? kMaxSatLev : kMaxCtmLev];
  double yr[kMaxSatLev > kMaxCtmLev ? kMaxSatLev : kMaxCtmLev];
  double xs[kMaxSatLev > kMaxCtmLev ? kMaxSatLev : kMaxCtmLev];
  double ys[kMaxSatLev > kMaxCtmLev ? kMaxSatLev : kMaxCtmLev];
  double va[kMaxCtmLev];
  double vb[kMaxCtmLev];
};

// ---------------------------------------------------------------------------
// AMF recalculation for one cell (amf_recal.py:93-119).  The caller has put
// log(p_sat) / scattering weights into s.xr / s.yr (n_sat entries).  Model
// column accessors return level k of this cell.  CTM_F32: native float32 model
// fields (partial column, log and column sum in float32).
// Returns new_amf; *col receives the model column (nansum of partial columns).
// ---------------------------------------------------------------------------
template <bool CTM_F32, typename GetPmid, typename GetPc>
__device__ __forceinline__ double warp_amf_cell(WarpScratch& s, int n_sat, int n_ctm,
                                                bool has_trop, double trop, GetPmid pmid_at,
                                                GetPc pc_at, double* col, int lane) {
  warp_sort_levels(s.xr, s.yr, n_sat, s.xs, s.ys, lane);
  float* pcf = reinterpret_cast<float*>(s.vb);
  for (int k = lane; k < n_ctm; k += 32) {
    double pm, lp, pc;
    if (CTM_F32) {
      const float pmf = (float)pmid_at(k);
      pm = (double)pmf;
      lp = log_as_f32(pmf);
      pc = pc_at(k);  // float32 value widened exactly
    } else {
      pm = pmid_at(k);
      lp = log(pm);
      pc = pc_at(k);
    }
    double sw = interp1d_linear(s.xs, s.ys, n_sat, lp);
    if (isinf(sw)) sw = 0.0;
    if (has_trop && pm < trop) { sw = qnan(); pc = qnan(); }
    const double prod = sw * pc;
    s.va[k] = (prod != prod) ? 0.0 : prod;  // nansum: NaN -> 0
    if (CTM_F32) pcf[k] = (pc != pc) ? 0.0f : (float)pc; else s.vb[k] = (pc != pc) ? 0.0 : pc;
  }
  __syncwarp();
  const double scd = warp_np_sum<double>(s.va, n_ctm, lane);
  const double vcd_m = CTM_F32 ? (double)warp_np_sum<float>(pcf, n_ctm, lane)
                               : warp_np_sum<double>(s.vb, n_ctm, lane);
  __syncwarp();
  *col = vcd_m;
  return vcd_m != 0.0 ? scd / vcd_m : qnan();
}

}  // namespace oisat
