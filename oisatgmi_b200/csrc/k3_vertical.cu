// K3: stand-alone per-cell vertical operators (one warp per cell).
//   oisat_vertical_amf     amf_recal.py:93-119,175-183
//   oisat_vertical_column  amf_recal.py:160-171 (no scattering weights)
//   oisat_vertical_mopitt  ak_conv_mopitt.py:118-142
//   oisat_vertical_gosat   ak_conv_gosat.py:118-141
// The reference loops over cells in Python and builds a scipy interp1d object per
// cell (113 us/cell); columns here are strided views of level-major arrays, so a
// warp's loads of one level for neighbouring cells coalesce through L1/L2.
#include "vertical.cuh"

namespace oisat {

constexpr int kWarpsPerBlock = 4;
constexpr int kArr = 128;  // >= kMaxSatLev, kMaxCtmLev

struct WarpArrays {
  double xr[kArr], yr[kArr], xs[kArr], ys[kArr], va[kArr], vb[kArr];
  __device__ __forceinline__ GroupScratch view() {
    GroupScratch s;
    s.xr = xr; s.yr = yr; s.xs = xs; s.ys = ys; s.va = va; s.vb = vb;
    return s;
  }
};

struct CtmView {
  const void* pmid;
  const void* b;   // profile (mode 0) / partial column or profile (mode 1)
  const void* dp;  // delta_p (mode 0) / air column (mode 1, AK operators)
  int64_t stride;
};

__device__ __forceinline__ float ld_f32(const void* p, int64_t i) { return ((const float*)p)[i]; }
__device__ __forceinline__ double ld_f64(const void* p, int64_t i) { return ((const double*)p)[i]; }

template <bool CTM_F32>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
vertical_amf_kernel(int64_t n_items, const int32_t* __restrict__ sat_index,
                    const int32_t* __restrict__ ctm_index, const double* __restrict__ vcd,
                    const double* __restrict__ amf, const double* __restrict__ trop,
                    const double* __restrict__ p_sat, const double* __restrict__ sw, int n_sat,
                    int64_t sat_stride, CtmView ctm, int n_ctm, double* __restrict__ new_amf,
                    double* __restrict__ ctm_vcd, double* __restrict__ vcd_out) {
  __shared__ WarpArrays scratch[kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (item >= n_items) return;
  const int64_t si = sat_index ? sat_index[item] : item;
  const int64_t ci = ctm_index ? ctm_index[item] : si;
  const double v = vcd[si];
  if (v != v) {  // amf_recal.py:99-100 and :176
    if (lane == 0) { new_amf[si] = qnan(); ctm_vcd[si] = qnan(); vcd_out[si] = qnan(); }
    return;
  }
  const GroupScratch s = scratch[warp].view();
  for (int l = lane; l < n_sat; l += 32) {
    s.xr[l] = log(p_sat[(int64_t)l * sat_stride + si]);
    s.yr[l] = sw[(int64_t)l * sat_stride + si];
  }
  __syncwarp();
  const bool has_trop = trop != nullptr;
  const double tp = has_trop ? trop[si] : 0.0;
  double col;
  double namf;
  if (CTM_F32) {
    auto pm = [&](int k) { return (double)ld_f32(ctm.pmid, (int64_t)k * ctm.stride + ci); };
    namf = group_amf_cell<32, true, false>(
        s, n_sat, n_ctm, has_trop, tp, pm,
        [&](int k) { return (double)log_f32(ld_f32(ctm.pmid, (int64_t)k * ctm.stride + ci)); },
        [&](int k) {
          return (double)partial_column_f32(ld_f32(ctm.dp, (int64_t)k * ctm.stride + ci),
                                            ld_f32(ctm.b, (int64_t)k * ctm.stride + ci));
        },
        &col, lane);
  } else {
    auto pm = [&](int k) { return ld_f64(ctm.pmid, (int64_t)k * ctm.stride + ci); };
    namf = group_amf_cell<32, false, false>(
        s, n_sat, n_ctm, has_trop, tp, pm, [&](int k) { return log(pm(k)); },
        [&](int k) { return ld_f64(ctm.b, (int64_t)k * ctm.stride + ci); }, &col, lane);
  }
  if (lane == 0) {
    const double vnew = (amf[si] * v) / namf;  // amf_recal.py:179
    new_amf[si] = namf;
    vcd_out[si] = vnew;
    ctm_vcd[si] = (vnew != vnew || isinf(vnew)) ? qnan() : col;  // :180-181
  }
}

// model column only (O3 path): np.nansum(pc, axis=0) is a sequential level-by-level
// add in the array dtype; one thread per cell.
template <bool CTM_F32>
__global__ void __launch_bounds__(128)
vertical_column_kernel(int64_t n_items, const int32_t* __restrict__ sat_index,
                       const int32_t* __restrict__ ctm_index, const double* __restrict__ vcd,
                       const double* __restrict__ trop, CtmView ctm, int n_ctm,
                       double* __restrict__ ctm_vcd) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items) return;
  const int64_t si = sat_index ? sat_index[item] : item;
  const int64_t ci = ctm_index ? ctm_index[item] : si;
  const double v = vcd[si];
  const bool has_trop = trop != nullptr;
  const double tp = has_trop ? trop[si] : 0.0;
  double out;
  if (CTM_F32) {
    float acc = 0.0f;
    for (int k = 0; k < n_ctm; ++k) {
      const float pm = ld_f32(ctm.pmid, (int64_t)k * ctm.stride + ci);
      float pc = partial_column_f32(ld_f32(ctm.dp, (int64_t)k * ctm.stride + ci),
                                    ld_f32(ctm.b, (int64_t)k * ctm.stride + ci));
      if (has_trop && (double)pm < tp) pc = CUDART_NAN_F;
      acc = __fadd_rn(acc, pc != pc ? 0.0f : pc);
    }
    out = (double)acc;
  } else {
    double acc = 0.0;
    for (int k = 0; k < n_ctm; ++k) {
      const double pm = ld_f64(ctm.pmid, (int64_t)k * ctm.stride + ci);
      double pc = ld_f64(ctm.b, (int64_t)k * ctm.stride + ci);
      if (has_trop && pm < tp) pc = qnan();
      acc = acc + (pc != pc ? 0.0 : pc);
    }
    out = acc;
  }
  ctm_vcd[si] = (v != v) ? qnan() : out;
}

// ---------------------------------------------------------------------------
// Averaging-kernel operators.  Here the MODEL column is the interpolation table
// (sorted in log pressure) and the satellite levels are the query points.
// ---------------------------------------------------------------------------
template <bool CTM_F32>
__device__ __forceinline__ void load_ctm_table(const GroupScratch& s, const CtmView& ctm,
                                               int n_ctm, int64_t ci, int lane) {
  for (int k = lane; k < n_ctm; k += 32) {
    if (CTM_F32) {
      s.xr[k] = (double)log_f32(ld_f32(ctm.pmid, (int64_t)k * ctm.stride + ci));
      s.yr[k] = (double)ld_f32(ctm.b, (int64_t)k * ctm.stride + ci);
    } else {
      s.xr[k] = log(ld_f64(ctm.pmid, (int64_t)k * ctm.stride + ci));
      s.yr[k] = ld_f64(ctm.b, (int64_t)k * ctm.stride + ci);
    }
  }
  __syncwarp();
  group_sort_levels<32>(s.xr, s.yr, n_ctm, s.xs, s.ys, lane);
}

template <bool CTM_F32>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
vertical_mopitt_kernel(int64_t n_items, const int32_t* __restrict__ sat_index,
                       const int32_t* __restrict__ ctm_index, const double* __restrict__ vcd,
                       const double* __restrict__ ap_col, const double* __restrict__ ap_sfc,
                       const double* __restrict__ p_sat, const double* __restrict__ ak,
                       const double* __restrict__ ap_prof, int n_sat, int64_t sat_stride,
                       CtmView ctm, int n_ctm, double* __restrict__ ctm_vcd,
                       double* __restrict__ ctm_xcol) {
  __shared__ WarpArrays scratch[kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (item >= n_items) return;
  const int64_t si = sat_index ? sat_index[item] : item;
  const int64_t ci = ctm_index ? ctm_index[item] : si;
  const double v = vcd[si];
  if (v != v) {
    if (lane == 0) { ctm_vcd[si] = qnan(); ctm_xcol[si] = qnan(); }
    return;
  }
  const GroupScratch s = scratch[warp].view();
  load_ctm_table<CTM_F32>(s, ctm, n_ctm, ci, lane);
  // interpolate the model profile to the satellite levels and weight with AK[1:]
  for (int l = lane; l < n_sat; l += 32) {
    const double q = log(p_sat[(int64_t)l * sat_stride + si]);
    double xi;
    if (CTM_F32) {
      // float32 tables: scipy stays on interp1d._call_linear, then NaN outside the range
      xi = interp1d_linear<false>(s.xs, s.ys, n_ctm, q);
      if (q < s.xs[0] || q > s.xs[n_ctm - 1]) xi = qnan();
    } else {
      xi = np_interp_nanfill(s.xs, s.ys, n_ctm, q);
    }
    const double term = ak[(int64_t)(l + 1) * sat_stride + si] *
                        (log10(xi) - log10(ap_prof[(int64_t)l * sat_stride + si]));
    s.va[l] = (term != term) ? 0.0 : term;
  }
  // air column of the model cell
  float* airf = reinterpret_cast<float*>(s.vb);
  for (int k = lane; k < n_ctm; k += 32) {
    if (CTM_F32) {
      const float a = air_column_f32(ld_f32(ctm.dp, (int64_t)k * ctm.stride + ci));
      airf[k] = (a != a) ? 0.0f : a;
    } else {
      const double a = ld_f64(ctm.dp, (int64_t)k * ctm.stride + ci);
      s.vb[k] = (a != a) ? 0.0 : a;
    }
  }
  __syncwarp();
  const double part = group_np_sum<double, 32>(s.va, n_sat, lane);
  const double air = CTM_F32 ? (double)group_np_sum<float, 32>(airf, n_ctm, lane)
                             : group_np_sum<double, 32>(s.vb, n_ctm, lane);
  if (lane == 0) {
    // surface term uses the bottom model layer as stored (unsorted index 0)
    double x0;
    if (CTM_F32) x0 = (double)log10f(ld_f32(ctm.b, ci)); else x0 = log10(ld_f64(ctm.b, ci));
    const double prof_part = ap_col[si] + part;
    const double sfc_part = ak[si] * (x0 - log10(ap_sfc[si]));
    const double col = prof_part + sfc_part;
    ctm_xcol[si] = 1e6 * col / air;
    ctm_vcd[si] = isinf(v) ? qnan() : col;  // ak_conv_mopitt.py:141-142
  }
}

template <bool CTM_F32>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
vertical_gosat_kernel(int64_t n_items, const int32_t* __restrict__ sat_index,
                      const int32_t* __restrict__ ctm_index, const double* __restrict__ x_col,
                      const double* __restrict__ p_sat, const double* __restrict__ ak,
                      const double* __restrict__ ap_prof, const double* __restrict__ pw, int n_sat,
                      int64_t sat_stride, CtmView ctm, int n_ctm, double* __restrict__ ctm_xcol) {
  __shared__ WarpArrays scratch[kWarpsPerBlock];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  if (item >= n_items) return;
  const int64_t si = sat_index ? sat_index[item] : item;
  const int64_t ci = ctm_index ? ctm_index[item] : si;
  const double v = x_col[si];
  if (v != v || isinf(v)) {  // ak_conv_gosat.py:120-121,139-140
    if (lane == 0) ctm_xcol[si] = qnan();
    return;
  }
  const GroupScratch s = scratch[warp].view();
  load_ctm_table<CTM_F32>(s, ctm, n_ctm, ci, lane);
  for (int l = lane; l < n_sat; l += 32) {
    const int64_t e = (int64_t)l * sat_stride + si;
    const double xi = interp1d_linear<false>(s.xs, s.ys, n_ctm, log(p_sat[e]));
    const double ap = ap_prof[e];
    double t = ap + (xi - ap) * ak[e];
    t = t * pw[e];
    s.va[l] = (t <= 0.0 || t != t) ? 0.0 : t;  // t<=0 -> NaN -> dropped by nansum
  }
  __syncwarp();
  const double tot = group_np_sum<double, 32>(s.va, n_sat, lane);
  if (lane == 0) ctm_xcol[si] = tot;
}

static int check_levels(int n_sat, int n_ctm) {
  if (n_sat < 2 || n_sat > kMaxSatLev) {
    set_error("satellite levels must be in [2, %d]", kMaxSatLev);
    return OISAT_E_UNSUPPORTED;
  }
  if (n_ctm < 2 || n_ctm > kMaxCtmLev) {
    set_error("model levels must be in [2, %d]", kMaxCtmLev);
    return OISAT_E_UNSUPPORTED;
  }
  return OISAT_OK;
}

}  // namespace oisat

using namespace oisat;

extern "C" int oisat_vertical_amf(int64_t n_items, const int32_t* sat_index,
                                  const int32_t* ctm_index, const double* vcd, const double* amf,
                                  const double* trop, const double* p_sat, const double* sw,
                                  int32_t n_sat_lev, int64_t sat_stride, const void* ctm_pmid,
                                  const void* ctm_b, const void* ctm_dp, int32_t ctm_mode,
                                  int32_t n_ctm_lev, int64_t ctm_stride, double* new_amf,
                                  double* ctm_vcd, double* vcd_out, void* stream) {
  if (n_items == 0) return OISAT_OK;
  OISAT_CHECK_ARG(vcd && amf && p_sat && sw && ctm_pmid && ctm_b && new_amf && ctm_vcd && vcd_out,
                  "null pointer");
  OISAT_CHECK_ARG(ctm_mode == 0 || ctm_mode == 1, "ctm_mode must be 0 or 1");
  OISAT_CHECK_ARG(ctm_mode == 1 || ctm_dp, "native mode needs delta_p");
  int rc = check_levels(n_sat_lev, n_ctm_lev);
  if (rc) return rc;
  CtmView ctm{ctm_pmid, ctm_b, ctm_dp, ctm_stride};
  const unsigned blocks = (unsigned)ceil_div(n_items, kWarpsPerBlock);
  cudaStream_t s = (cudaStream_t)stream;
  if (ctm_mode == 0)
    vertical_amf_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, s>>>(
        n_items, sat_index, ctm_index, vcd, amf, trop, p_sat, sw, n_sat_lev, sat_stride, ctm,
        n_ctm_lev, new_amf, ctm_vcd, vcd_out);
  else
    vertical_amf_kernel<false><<<blocks, kWarpsPerBlock * 32, 0, s>>>(
        n_items, sat_index, ctm_index, vcd, amf, trop, p_sat, sw, n_sat_lev, sat_stride, ctm,
        n_ctm_lev, new_amf, ctm_vcd, vcd_out);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_vertical_column(int64_t n_items, const int32_t* sat_index,
                                     const int32_t* ctm_index, const double* vcd,
                                     const double* trop, const void* ctm_pmid, const void* ctm_b,
                                     const void* ctm_dp, int32_t ctm_mode, int32_t n_ctm_lev,
                                     int64_t ctm_stride, double* ctm_vcd, void* stream) {
  if (n_items == 0) return OISAT_OK;
  OISAT_CHECK_ARG(vcd && ctm_pmid && ctm_b && ctm_vcd, "null pointer");
  OISAT_CHECK_ARG(ctm_mode == 0 || ctm_mode == 1, "ctm_mode must be 0 or 1");
  OISAT_CHECK_ARG(ctm_mode == 1 || ctm_dp, "native mode needs delta_p");
  OISAT_CHECK_ARG(n_ctm_lev >= 1, "bad level count");
  CtmView ctm{ctm_pmid, ctm_b, ctm_dp, ctm_stride};
  const unsigned blocks = (unsigned)ceil_div(n_items, 128);
  cudaStream_t s = (cudaStream_t)stream;
  if (ctm_mode == 0)
    vertical_column_kernel<true><<<blocks, 128, 0, s>>>(n_items, sat_index, ctm_index, vcd, trop,
                                                       ctm, n_ctm_lev, ctm_vcd);
  else
    vertical_column_kernel<false><<<blocks, 128, 0, s>>>(n_items, sat_index, ctm_index, vcd, trop,
                                                        ctm, n_ctm_lev, ctm_vcd);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_vertical_mopitt(int64_t n_items, const int32_t* sat_index,
                                     const int32_t* ctm_index, const double* vcd,
                                     const double* ap_col, const double* ap_sfc,
                                     const double* p_sat, const double* ak, const double* ap_prof,
                                     int32_t n_sat_lev, int64_t sat_stride, const void* ctm_pmid,
                                     const void* ctm_prof, const void* ctm_dp_or_air,
                                     int32_t ctm_mode, int32_t n_ctm_lev, int64_t ctm_stride,
                                     double* ctm_vcd, double* ctm_xcol, void* stream) {
  if (n_items == 0) return OISAT_OK;
  OISAT_CHECK_ARG(vcd && ap_col && ap_sfc && p_sat && ak && ap_prof && ctm_pmid && ctm_prof &&
                      ctm_dp_or_air && ctm_vcd && ctm_xcol, "null pointer");
  OISAT_CHECK_ARG(ctm_mode == 0 || ctm_mode == 1, "ctm_mode must be 0 or 1");
  int rc = check_levels(n_sat_lev, n_ctm_lev);
  if (rc) return rc;
  CtmView ctm{ctm_pmid, ctm_prof, ctm_dp_or_air, ctm_stride};
  const unsigned blocks = (unsigned)ceil_div(n_items, kWarpsPerBlock);
  cudaStream_t s = (cudaStream_t)stream;
  if (ctm_mode == 0)
    vertical_mopitt_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, s>>>(
        n_items, sat_index, ctm_index, vcd, ap_col, ap_sfc, p_sat, ak, ap_prof, n_sat_lev,
        sat_stride, ctm, n_ctm_lev, ctm_vcd, ctm_xcol);
  else
    vertical_mopitt_kernel<false><<<blocks, kWarpsPerBlock * 32, 0, s>>>(
        n_items, sat_index, ctm_index, vcd, ap_col, ap_sfc, p_sat, ak, ap_prof, n_sat_lev,
        sat_stride, ctm, n_ctm_lev, ctm_vcd, ctm_xcol);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_vertical_gosat(int64_t n_items, const int32_t* sat_index,
                                    const int32_t* ctm_index, const double* x_col,
                                    const double* p_sat, const double* ak, const double* ap_prof,
                                    const double* pw, int32_t n_sat_lev, int64_t sat_stride,
                                    const void* ctm_pmid, const void* ctm_prof, int32_t ctm_mode,
                                    int32_t n_ctm_lev, int64_t ctm_stride, double* ctm_xcol,
                                    void* stream) {
  if (n_items == 0) return OISAT_OK;
  OISAT_CHECK_ARG(x_col && p_sat && ak && ap_prof && pw && ctm_pmid && ctm_prof && ctm_xcol,
                  "null pointer");
  OISAT_CHECK_ARG(ctm_mode == 0 || ctm_mode == 1, "ctm_mode must be 0 or 1");
  int rc = check_levels(n_sat_lev, n_ctm_lev);
  if (rc) return rc;
  CtmView ctm{ctm_pmid, ctm_prof, nullptr, ctm_stride};
  const unsigned blocks = (unsigned)ceil_div(n_items, kWarpsPerBlock);
  cudaStream_t s = (cudaStream_t)stream;
  if (ctm_mode == 0)
    vertical_gosat_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, s>>>(
        n_items, sat_index, ctm_index, x_col, p_sat, ak, ap_prof, pw, n_sat_lev, sat_stride, ctm,
        n_ctm_lev, ctm_xcol);
  else
    vertical_gosat_kernel<false><<<blocks, kWarpsPerBlock * 32, 0, s>>>(
        n_items, sat_index, ctm_index, x_col, p_sat, ak, ap_prof, pw, n_sat_lev, sat_stride, ctm,
        n_ctm_lev, ctm_xcol);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
