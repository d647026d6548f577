// Building blocks of the per-cell vertical operators (K3).  A GROUP of LANES
// threads (a full warp, or a half warp in the fused month kernel) works on one
// model cell; group scratch lives in shared memory.  The same routines serve the
// stand-alone kernels (k3_vertical.cu) and the fused pipeline (fused_amf.cu).
// All 32 lanes of a warp always execute these routines together (the two halves
// on different cells), so full-mask shuffles/votes are used and per-group
// results are extracted with the group's lane mask.
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
constexpr int kMaxCtmLev = 127;  // one leaf of numpy's pairwise sum, 7-step search

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

// float32 natural log as numpy evaluates np.log on a float32 array (the result
// is float32).  CUDA's logf is within 1 ulp; numpy's SIMD kernel within ~4 ulp;
// the two agree to float32 rounding noise, which is the reference's own noise
// floor for this term.
__device__ __forceinline__ float log_f32(float p) { return logf(p); }

__device__ __forceinline__ bool nan_less(double a, double b) {
  // numpy's sort/searchsorted order: NaN is larger than everything
  return a < b || (b != b && a == a);
}

template <int LANES>
__device__ __forceinline__ unsigned group_mask(int lane) {
  return LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << (lane & ~(LANES - 1)));
}

template <int LANES>
__device__ __forceinline__ bool group_all(bool pred, int lane) {
  const unsigned m = group_mask<LANES>(lane);
  return (__ballot_sync(0xffffffffu, pred) & m) == m;
}

// Pairwise sum (numpy order) of vals[0..n), n <= 128, by the first 8 lanes of the
// group.  NaNs must already be replaced (nansum).  Result on all lanes of the group.
template <typename T, int LANES>
__device__ __forceinline__ T group_np_sum(const T* vals, int n, int lane) {
  const int gl = lane & (LANES - 1);
  T res;
  if (n < 8) {
    res = (T)0;
    for (int i = 0; i < n; ++i) res = res + vals[i];
    return res;
  }
  const int body = n - (n % 8);
  T r = (T)0;
  if (gl < 8) {
    r = vals[gl];
    for (int i = 8 + gl; i < body; i += 8) r = r + vals[i];
  }
  r = r + __shfl_xor_sync(0xffffffffu, r, 1);
  r = r + __shfl_xor_sync(0xffffffffu, r, 2);
  r = r + __shfl_xor_sync(0xffffffffu, r, 4);
  res = __shfl_sync(0xffffffffu, r, 0, LANES);
  for (int i = body; i < n; ++i) res = res + vals[i];
  return res;
}

// Stable ascending argsort of (x, y) pairs held in shared memory, NaN last
// (np.argsort(kind='mergesort')), written to xs/ys.  Fast paths for the two
// monotone cases every real profile falls into.
template <int LANES>
__device__ __forceinline__ void group_sort_levels(const double* xr, const double* yr, int n,
                                                  double* xs, double* ys, int lane) {
  const int gl = lane & (LANES - 1);
  bool inc = true, dec = true;
  for (int i = gl; i + 1 < n; i += LANES) {
    const double a = xr[i], b = xr[i + 1];
    inc = inc && (a <= b);  // equal keys keep their order under a stable sort
    dec = dec && (a > b);
  }
  inc = group_all<LANES>(inc, lane);
  dec = group_all<LANES>(dec, lane);
  if (inc) {
    for (int i = gl; i < n; i += LANES) { xs[i] = xr[i]; ys[i] = yr[i]; }
  } else if (dec) {
    for (int i = gl; i < n; i += LANES) { xs[n - 1 - i] = xr[i]; ys[n - 1 - i] = yr[i]; }
  } else {
    for (int i = gl; i < n; i += LANES) {
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

// np.searchsorted(xs, v, side='left') for n <= 127: fixed trip count, no
// data-dependent branches (the two halves of a warp search different tables)
// A NaN table entry compares false, i.e. "larger", like numpy's NaN-last order; a
// NaN query gives position 0 where numpy gives n, but every caller turns a NaN
// query into a NaN result whatever the bracket, so the outputs are identical.
__device__ __forceinline__ int searchsorted_left(const double* xs, int n, double v) {
  int pos = 0;
  for (int step = 1 << (31 - __clz(n)); step >= 1; step >>= 1) {  // uniform trip count
    const int t = pos + step;
    const int probe = (t <= n ? t : n) - 1;
    const bool take = (t <= n) && (xs[probe] < v);
    pos = take ? t : pos;
  }
  return pos;
}

// interp1d._call_linear (linear extrapolation falls out of the index clipping).
// FAST: one reciprocal instead of two divisions (differs from scipy in the last
// bit; used by the fused kernel only).
template <bool FAST>
__device__ __forceinline__ double interp1d_linear(const double* xs, const double* ys, int n,
                                                  double v) {
  int idx = searchsorted_left(xs, n, v);
  idx = idx < 1 ? 1 : idx;
  idx = idx > n - 1 ? n - 1 : idx;
  const double x_lo = xs[idx - 1], x_hi = xs[idx];
  const double y_lo = ys[idx - 1], y_hi = ys[idx];
  const double den = x_hi - x_lo;
  if (FAST) {
    const double r = 1.0 / den;
    return ((v - x_lo) * r) * y_hi + ((x_hi - v) * r) * y_lo;
  }
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

// Group scratch.  In the stand-alone kernels the six arrays are distinct; the
// fused kernel aliases them to save shared memory: [xs|ys] doubles as the staging
// area of the gridded rows, and va/vb live in [xr|yr], which is dead once the
// table is sorted.
struct GroupScratch {
  double* xr;
  double* yr;
  double* xs;
  double* ys;
  double* va;
  double* vb;
};

__host__ __device__ inline int scratch_sorted_len(int n_sat, int n_rows) {
  const int half_rows = (n_rows + 1) / 2;
  return n_sat > half_rows ? n_sat : half_rows;
}
__host__ __device__ inline int scratch_raw_len(int n_sat, int n_ctm) {
  return n_sat > n_ctm ? n_sat : n_ctm;
}
__host__ __device__ inline int scratch_doubles(int n_sat, int n_ctm, int n_rows) {
  return 2 * scratch_sorted_len(n_sat, n_rows) + 2 * scratch_raw_len(n_sat, n_ctm);
}
__device__ __forceinline__ GroupScratch carve_scratch(double* base, int n_sat, int n_ctm,
                                                      int n_rows) {
  GroupScratch s;
  const int a = scratch_sorted_len(n_sat, n_rows), b = scratch_raw_len(n_sat, n_ctm);
  s.xs = base;
  s.ys = base + a;
  s.xr = base + 2 * a;
  s.yr = base + 2 * a + b;
  s.va = s.xr;
  s.vb = s.yr;
  return s;
}

// ---------------------------------------------------------------------------
// AMF recalculation for one cell (amf_recal.py:93-119).  The caller has put
// log(p_sat) / scattering weights into s.xr / s.yr (n_sat entries).  Accessors
// return, for model level k of this cell: pm(k) p_mid, lp(k) log p_mid and pc(k)
// the partial column -- float32 values widened exactly when CTM_F32 (native
// model fields: the column sum is then a float32 sum like numpy's), float64
// otherwise.  Returns new_amf; *col receives nansum(partial column).
// ---------------------------------------------------------------------------
template <int LANES, bool CTM_F32, bool FAST, typename GetPm, typename GetLp, typename GetPc>
__device__ __forceinline__ double group_amf_cell(const GroupScratch& s, int n_sat, int n_ctm,
                                                 bool has_trop, double trop, GetPm pm_at,
                                                 GetLp lp_at, GetPc pc_at, double* col, int lane) {
  const int gl = lane & (LANES - 1);
  group_sort_levels<LANES>(s.xr, s.yr, n_sat, s.xs, s.ys, lane);
  double* va = s.va;
  double* vb = s.vb;
  float* pcf = reinterpret_cast<float*>(vb);
  for (int k = gl; k < n_ctm; k += LANES) {
    double pc = pc_at(k);
    double sw = interp1d_linear<FAST>(s.xs, s.ys, n_sat, lp_at(k));
    if (isinf(sw)) sw = 0.0;
    if (has_trop && pm_at(k) < trop) { sw = qnan(); pc = qnan(); }
    const double prod = sw * pc;
    va[k] = (prod != prod) ? 0.0 : prod;  // nansum: NaN -> 0
    if (CTM_F32) pcf[k] = (pc != pc) ? 0.0f : (float)pc; else vb[k] = (pc != pc) ? 0.0 : pc;
  }
  __syncwarp();
  const double scd = group_np_sum<double, LANES>(va, n_ctm, lane);
  const double vcd_m = CTM_F32 ? (double)group_np_sum<float, LANES>(pcf, n_ctm, lane)
                               : group_np_sum<double, LANES>(vb, n_ctm, lane);
  __syncwarp();
  *col = vcd_m;
  return vcd_m != 0.0 ? scd / vcd_m : qnan();
}

}  // namespace oisat
