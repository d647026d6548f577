// K7: reader front-end -- what the reference's level-2
// readers do to the file variables between reading them and calling
// `interpolator`: unit scaling + float16 quantisation, quality flags, scattering-
// weight clean-up (and the pixel-major -> level-major transposition), mid-level
// pressures from hybrid coefficients, tropopause pressure from a layer index.
// reader.py:807-903 (OMI NO2), :906-983 (OMI HCHO), :707-804 (TROPOMI NO2).
//
// In the reference these are numpy expressions plus, for the OMI NO2 flags, a Python
// loop that formats every pixel's flag as a binary string (reader.py:862-869).
// Everything here is elementwise and bit-exact: each kernel evaluates the SAME
// roundings numpy's promotion rules produce (which operand is float16 / float32 /
// float64 at every step is spelled out next to the code).
#include "common.cuh"

namespace oisat {

__device__ __forceinline__ __half to_half(const void* p, int dtype, int64_t i) {
  switch (dtype) {
    case OISAT_F16: return reinterpret_cast<const __half*>(p)[i];
    case OISAT_F32: return __float2half_rn(reinterpret_cast<const float*>(p)[i]);
    case OISAT_F64: return __double2half(reinterpret_cast<const double*>(p)[i]);
    case OISAT_I32: return __int2half_rn(reinterpret_cast<const int32_t*>(p)[i]);
    default: return __ushort_as_half(0x7e00);
  }
}

// out = float16( (..(x * f0) * f1 ..) ) with the products in x's own dtype: a Python
// float times a float32 array stays float32 (reader.py:846-847, 752-754)
__global__ void __launch_bounds__(256)
rd_scale_kernel(const void* __restrict__ src, int dtype, int64_t n, double f0, double f1, double f2,
                int nf, __half* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (dtype == OISAT_F32) {
    float y = reinterpret_cast<const float*>(src)[i];
    if (nf > 0) y = __fmul_rn(y, (float)f0);
    if (nf > 1) y = __fmul_rn(y, (float)f1);
    if (nf > 2) y = __fmul_rn(y, (float)f2);
    out[i] = __float2half_rn(y);
  } else if (dtype == OISAT_F64) {
    double y = reinterpret_cast<const double*>(src)[i];
    if (nf > 0) y = __dmul_rn(y, f0);
    if (nf > 1) y = __dmul_rn(y, f1);
    if (nf > 2) y = __dmul_rn(y, f2);
    out[i] = __double2half(y);
  } else {
    out[i] = to_half(src, dtype, i);   // plain cast (nf must be 0)
  }
}

// mode 0, OMI NO2 (reader.py:849-870): usable unless bits 0 and 1 of int(float16(flag))
// are both set; times (float16(cloud) < float16(0.3)) times (float16(terrain) <
// float16(0.2)), the product in float64.  mode 1, OMI HCHO (:939-949):
// (float16(flag) == 0) * (float16(cloud) < float16(0.4)).
__global__ void __launch_bounds__(256)
rd_quality_kernel(int mode, const void* __restrict__ flags, int fdtype,
                  const void* __restrict__ cloud, int cdtype, const void* __restrict__ terrain,
                  int tdtype, int64_t n, __half cloud_max, __half terrain_max,
                  double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const __half raw = to_half(flags, fdtype, i);
  const double c = __hlt(to_half(cloud, cdtype, i), cloud_max) ? 1.0 : 0.0;
  if (mode == 0) {
    long long v = (long long)__half2float(raw);   // int(): toward zero
    if (v < 0) v = -v;
    const double q = (v & 3) == 3 ? -100.0 : 1.0;
    const double t = __hlt(to_half(terrain, tdtype, i), terrain_max) ? 1.0 : 0.0;
    out[i] = __dmul_rn(__dmul_rn(q, c), t);
  } else {
    const double q = __heq(raw, __float2half_rn(0.0f)) ? 1.0 : 0.0;
    out[i] = __dmul_rn(q, c);
  }
}

__device__ __forceinline__ __half clean_weight(__half w) {
  // reader.py:887-888: NaN, inf, > 100, < 0 -> 0 (comparisons on the float16 value)
  const float f = __half2float(w);
  return (f != f || isinf(f) || f > 100.0f || f < 0.0f) ? __float2half_rn(0.0f) : w;
}

// float16 weights, level-major out[l][p]; `scale` (may be null) multiplies the float16
// value in scale's dtype before the result is rounded back to float16 (TROPOMI:
// averaging kernel x total AMF, reader.py:773-774).  Pixel-major input ([p][l], OMI NO2
// and TROPOMI files) goes through shared memory so that both the loads and the stores
// are coalesced.
constexpr int kWeightPixels = 128;

__global__ void __launch_bounds__(256)
rd_weights_kernel(const void* __restrict__ src, int dtype, int pixel_major, int L, int64_t n_px,
                  const void* __restrict__ scale, int sdtype, __half* __restrict__ out) {
  auto finish = [&](__half w, int64_t p) {
    if (scale != nullptr) {
      if (sdtype == OISAT_F64)
        w = __double2half(__dmul_rn((double)__half2float(w), reinterpret_cast<const double*>(scale)[p]));
      else
        w = __float2half_rn(__fmul_rn(__half2float(w), reinterpret_cast<const float*>(scale)[p]));
    }
    return clean_weight(w);
  };
  if (!pixel_major) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)L * n_px) out[i] = finish(to_half(src, dtype, i), i % n_px);
    return;
  }
  // pixel-major: kWeightPixels pixels x L levels are ONE contiguous piece of the file
  // array: coalesced loads into shared memory, then level rows of kWeightPixels pixels
  // written two pixels per thread (128 bytes per level and warp)
  extern __shared__ __half wtile[];          // [kWeightPixels][L + 1]
  const int pitch = L + 1;
  const int64_t p0 = (int64_t)blockIdx.x * kWeightPixels;
  const int64_t left = n_px - p0;
  const int n_here = left < kWeightPixels ? (int)left : kWeightPixels;
  for (int i = threadIdx.x; i < n_here * L; i += blockDim.x) {
    const int px = i / L, l = i - px * L;
    wtile[px * pitch + l] = to_half(src, dtype, p0 * L + i);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L * (kWeightPixels / 2); i += blockDim.x) {
    const int l = i / (kWeightPixels / 2), px = 2 * (i - l * (kWeightPixels / 2));
    if (px + 1 < n_here && ((n_px & 1) == 0)) {
      const __half a = finish(wtile[px * pitch + l], p0 + px);
      const __half b = finish(wtile[(px + 1) * pitch + l], p0 + px + 1);
      *reinterpret_cast<__half2*>(out + (int64_t)l * n_px + p0 + px) = __halves2half2(a, b);
    } else {
      if (px < n_here) out[(int64_t)l * n_px + p0 + px] = finish(wtile[px * pitch + l], p0 + px);
      if (px + 1 < n_here)
        out[(int64_t)l * n_px + p0 + px + 1] = finish(wtile[(px + 1) * pitch + l], p0 + px + 1);
    }
  }
}

// mid-level pressures, level-major float16.
//   mode 0: out[l][p] = float16(a[l])                                (OMI NO2, reader.py:872-886)
//   mode 1: ps16 = float16(ps); 0.5*((a_l + b_l ps16) + (a_l+1 + b_l+1 ps16))   (OMI HCHO, :965-966)
//   mode 2: ps32 = float32(ps) / float32(ps_div); 0.5*(((a_l + b_l ps32) + a_l+1) + b_l+1 ps32)
//                                                                    (TROPOMI, :763,772-773)
//   mode 3: the same with float32 coefficients, i.e. the whole expression in float32 -- what
//           numpy does when the file stores tm5_constant_a/b as float32 (the elements are
//           then numpy float32 scalars; concatenating the Python 0 does not widen them)
// a, b are passed as float64; in modes 1 and 2 they ARE float64 in the reference (numpy
// float64 scalars promote the whole expression to float64).
__global__ void __launch_bounds__(256)
rd_pmid_kernel(int mode, const double* __restrict__ a, const double* __restrict__ b,
               const void* __restrict__ ps, int psdtype, double ps_div, int L, int64_t n_px,
               __half* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_px) return;
  double x = 0.0;
  if (mode == 1) {
    x = (double)__half2float(to_half(ps, psdtype, p));
  } else if (mode >= 2) {
    const float v = psdtype == OISAT_F64 ? (float)reinterpret_cast<const double*>(ps)[p]
                                         : reinterpret_cast<const float*>(ps)[p];
    x = (double)__fdiv_rn(v, (float)ps_div);
  }
  for (int l = 0; l < L; ++l) {
    double v;
    if (mode == 0) {
      v = a[l];
    } else if (mode == 1) {
      const double lo = __dadd_rn(a[l], __dmul_rn(b[l], x));
      const double hi = __dadd_rn(a[l + 1], __dmul_rn(b[l + 1], x));
      v = __dmul_rn(0.5, __dadd_rn(lo, hi));
    } else if (mode == 2) {
      double s = __dadd_rn(a[l], __dmul_rn(b[l], x));
      s = __dadd_rn(s, a[l + 1]);
      s = __dadd_rn(s, __dmul_rn(b[l + 1], x));
      v = __dmul_rn(0.5, s);
    } else {
      const float xf = (float)x;
      float s = __fadd_rn((float)a[l], __fmul_rn((float)b[l], xf));
      s = __fadd_rn(s, (float)a[l + 1]);
      s = __fadd_rn(s, __fmul_rn((float)b[l + 1], xf));
      out[(int64_t)l * n_px + p] = __float2half_rn(__fmul_rn(0.5f, s));
      continue;
    }
    out[(int64_t)l * n_px + p] = __double2half(v);
  }
}

// tropopause pressure = p_mid of the tropopause layer, NaN when the index is not in
// (0, L) (reader.py:781-789)
__global__ void __launch_bounds__(256)
rd_tropopause_kernel(const int32_t* __restrict__ layer, const __half* __restrict__ p_mid, int L,
                     int64_t n_px, __half* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_px) return;
  const int k = layer[p];
  out[p] = (k > 0 && k < L) ? p_mid[(int64_t)k * n_px + p] : __ushort_as_half(0x7e00);
}

static bool float_dtype(int d) { return d == OISAT_F16 || d == OISAT_F32 || d == OISAT_F64; }

// MOPITT / GOSAT clean-up (reader.py:1143-1203, 1228-1262): optional thresholds on the SOURCE
// value (bit 0: v <= 0 -> NaN, bit 1: inf -> NaN), up to three factors multiplied in the
// array's own dtype (NEP 50: a Python float times a float32 array stays float32), cast to the
// output dtype, optional threshold on the RESULT (bit 0: r <= 0 -> NaN, as for the a-priori
// column, :1180-1181), and the (pixel, level) -> (level, pixel) transposition of the files'
// profile variables.  Thread = output element; a NaN compares false, so it stays NaN.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
rd_clean_kernel(const TI* __restrict__ src, int64_t n_px, int n_lev, int pixel_major, int pre,
                double f0, double f1, double f2, int nf, int post, TO* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_px * n_lev) return;
  const int64_t lev = i / n_px, px = i - lev * n_px;
  TI v = src[pixel_major ? px * n_lev + lev : i];
  if (((pre & 1) && v <= (TI)0) || ((pre & 2) && isinf(v))) v = (TI)CUDART_NAN;
  if (nf > 0) v = v * (TI)f0;      // (--fmad=false: one rounding per multiply)
  if (nf > 1) v = v * (TI)f1;
  if (nf > 2) v = v * (TI)f2;
  TO r;
  if constexpr (sizeof(TO) == 2) {
    r = sizeof(TI) == 8 ? __double2half((double)v) : __float2half_rn((float)v);
    if ((post & 1) && __half2float(r) <= 0.0f) r = __float2half_rn(CUDART_NAN_F);
  } else {
    r = (TO)v;
    if ((post & 1) && r <= (TO)0) r = (TO)CUDART_NAN;
  }
  out[i] = r;
}

// x_col of the MOPITT reader (reader.py:1168): (1e6 * vcd / (dry * 1e-15)).astype(float32) with
// vcd float16 -- numpy forms 1e6 * vcd in float16 (it overflows to inf: kept, the reference
// does the same), dry * 1e-15 in float32, and the quotient in float32.
__global__ void __launch_bounds__(256)
rd_mopitt_xcol_kernel(const __half* __restrict__ vcd, const float* __restrict__ dry, int64_t n,
                      float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const __half num = __float2half_rn(__fmul_rn(__half2float(vcd[i]), __half2float(__float2half_rn(1e6f))));
  out[i] = __fdiv_rn(__half2float(num), __fmul_rn(dry[i], 1e-15f));
}

}  // namespace oisat

using namespace oisat;

extern "C" int oisat_reader_scale_f16(const void* src, int32_t dtype, int64_t n,
                                      const double* h_factors, int32_t n_factors, void* out,
                                      void* stream) {
  if (n <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(src && out, "null pointer");
  OISAT_CHECK_ARG(n_factors >= 0 && n_factors <= 3 && (n_factors == 0 || h_factors), "bad factors");
  OISAT_CHECK_ARG(float_dtype(dtype) || dtype == OISAT_I32, "bad dtype");
  OISAT_CHECK_ARG(n_factors == 0 || dtype == OISAT_F32 || dtype == OISAT_F64,
                  "scaling needs float32/float64 input");
  const double f0 = n_factors > 0 ? h_factors[0] : 1.0, f1 = n_factors > 1 ? h_factors[1] : 1.0,
               f2 = n_factors > 2 ? h_factors[2] : 1.0;
  rd_scale_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      src, dtype, n, f0, f1, f2, n_factors, (__half*)out);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_reader_quality(int32_t mode, const void* flags, int32_t flags_dtype,
                                    const void* cloud, int32_t cloud_dtype, const void* terrain,
                                    int32_t terrain_dtype, int64_t n, double* quality_flag,
                                    void* stream) {
  if (n <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(mode == 0 || mode == 1, "mode: 0 = OMI NO2, 1 = OMI HCHO");
  OISAT_CHECK_ARG(flags && cloud && quality_flag && (mode == 1 || terrain), "null pointer");
  OISAT_CHECK_ARG((float_dtype(flags_dtype) || flags_dtype == OISAT_I32) && float_dtype(cloud_dtype) &&
                      (mode == 1 || float_dtype(terrain_dtype)), "bad dtype");
  // the thresholds are Python floats compared with float16 arrays: numpy compares in float16
  const __half cloud_max = __double2half(mode == 0 ? 0.3 : 0.4);
  const __half terrain_max = __double2half(0.2);
  rd_quality_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      mode, flags, flags_dtype, cloud, cloud_dtype, terrain, terrain_dtype, n, cloud_max,
      terrain_max, quality_flag);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_reader_weights(const void* src, int32_t dtype, int32_t pixel_major,
                                    int32_t n_lev, int64_t n_px, const void* scale,
                                    int32_t scale_dtype, void* out, void* stream) {
  if (n_px <= 0 || n_lev <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(src && out, "null pointer");
  OISAT_CHECK_ARG(float_dtype(dtype), "bad dtype");
  OISAT_CHECK_ARG(!scale || scale_dtype == OISAT_F32 || scale_dtype == OISAT_F64, "bad scale dtype");
  cudaStream_t s = (cudaStream_t)stream;
  if (pixel_major) {
    const size_t smem = (size_t)kWeightPixels * (n_lev + 1) * sizeof(__half);
    OISAT_CHECK_ARG(smem <= 48 * 1024, "too many levels");
    rd_weights_kernel<<<(unsigned)ceil_div(n_px, kWeightPixels), 256, smem, s>>>(
        src, dtype, 1, n_lev, n_px, scale, scale_dtype, (__half*)out);
  } else {
    rd_weights_kernel<<<(unsigned)ceil_div((int64_t)n_lev * n_px, 256), 256, 0, s>>>(
        src, dtype, 0, n_lev, n_px, scale, scale_dtype, (__half*)out);
  }
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_reader_pmid(int32_t mode, const double* a, const double* b, const void* ps,
                                 int32_t ps_dtype, double ps_div, int32_t n_lev, int64_t n_px,
                                 void* out, void* stream) {
  if (n_px <= 0 || n_lev <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(mode >= 0 && mode <= 3, "mode: 0 constants, 1 OMI HCHO, 2/3 TROPOMI");
  OISAT_CHECK_ARG(a && out && (mode == 0 || (b && ps)), "null pointer");
  OISAT_CHECK_ARG(mode == 0 || float_dtype(ps_dtype), "bad dtype");
  OISAT_CHECK_ARG(mode < 2 || (ps_dtype != OISAT_F16 && ps_div != 0.0), "bad surface pressure");
  rd_pmid_kernel<<<(unsigned)ceil_div(n_px, 256), 256, 0, (cudaStream_t)stream>>>(
      mode, a, b, ps, ps_dtype, ps_div, n_lev, n_px, (__half*)out);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_reader_tropopause(const int32_t* layer, const void* p_mid, int32_t n_lev,
                                       int64_t n_px, void* out, void* stream) {
  if (n_px <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(layer && p_mid && out && n_lev >= 1, "null pointer");
  rd_tropopause_kernel<<<(unsigned)ceil_div(n_px, 256), 256, 0, (cudaStream_t)stream>>>(
      layer, (const __half*)p_mid, n_lev, n_px, (__half*)out);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

template <typename TI>
static int launch_clean(const void* src, int64_t n_px, int n_lev, int pixel_major, int pre, double f0,
                        double f1, double f2, int nf, int out_dtype, int post, void* out,
                        cudaStream_t s) {
  const unsigned blocks = (unsigned)ceil_div(n_px * n_lev, 256);
  if (out_dtype == OISAT_F16)
    rd_clean_kernel<TI, __half><<<blocks, 256, 0, s>>>((const TI*)src, n_px, n_lev, pixel_major, pre,
                                                       f0, f1, f2, nf, post, (__half*)out);
  else if (out_dtype == OISAT_F32)
    rd_clean_kernel<TI, float><<<blocks, 256, 0, s>>>((const TI*)src, n_px, n_lev, pixel_major, pre,
                                                      f0, f1, f2, nf, post, (float*)out);
  else
    rd_clean_kernel<TI, double><<<blocks, 256, 0, s>>>((const TI*)src, n_px, n_lev, pixel_major, pre,
                                                       f0, f1, f2, nf, post, (double*)out);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_reader_clean(const void* src, int32_t dtype, int64_t n_px, int32_t n_lev,
                                  int32_t pixel_major, int32_t pre, const double* h_factors,
                                  int32_t n_factors, int32_t out_dtype, int32_t post, void* out,
                                  void* stream) {
  if (n_px <= 0 || n_lev <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(src && out, "null pointer");
  OISAT_CHECK_ARG(dtype == OISAT_F32 || dtype == OISAT_F64, "source must be float32 or float64");
  OISAT_CHECK_ARG(float_dtype(out_dtype), "bad output dtype");
  OISAT_CHECK_ARG(n_factors >= 0 && n_factors <= 3 && (n_factors == 0 || h_factors), "bad factors");
  const double f0 = n_factors > 0 ? h_factors[0] : 1.0, f1 = n_factors > 1 ? h_factors[1] : 1.0,
               f2 = n_factors > 2 ? h_factors[2] : 1.0;
  cudaStream_t s = (cudaStream_t)stream;
  return dtype == OISAT_F32
             ? launch_clean<float>(src, n_px, n_lev, pixel_major, pre, f0, f1, f2, n_factors, out_dtype,
                                   post, out, s)
             : launch_clean<double>(src, n_px, n_lev, pixel_major, pre, f0, f1, f2, n_factors,
                                    out_dtype, post, out, s);
}

extern "C" int oisat_reader_mopitt_xcol(const void* vcd_f16, const float* dry_air, int64_t n,
                                        float* x_col, void* stream) {
  if (n <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(vcd_f16 && dry_air && x_col, "null pointer");
  rd_mopitt_xcol_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __half*)vcd_f16, dry_air, n, x_col);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
