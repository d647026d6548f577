// K2 (generic): apply a geometry plan to every field of a granule in one launch,
// and K6: model -> satellite-grid resampling.
//
// The reference grids one 2-D field at a time -- LinearNDInterpolator on the
// fine mesh, distance mask, convolve2d box mean, KD-tree nearest sampling --
// 74 times per OMI NO2 granule with identical geometry (interpolator.py:162-283).
// All of it is linear in the pixel values, so the plan turns
// it into one stencil per output cell, applied here to all rows at once.
#include "vertical.cuh"

namespace oisat {

constexpr int kMaxFields = 16;
constexpr int kRowChunk = 8;

struct FieldTable {
  oisat_field f[kMaxFields];
  int32_t row0[kMaxFields + 1];  // first global row of each field
  int32_t n_fields;
};

__device__ __forceinline__ double field_value(const oisat_field& f, int lev, int32_t v) {
  const int64_t i = (int64_t)lev * f.lev_stride + v;
  return f.op == OISAT_OP_SQUARE_NATIVE ? load_square_native(f.data, f.dtype, i)
                                        : load_as_double(f.data, f.dtype, i);
}

// thread = output cell, blockIdx.y = chunk of kRowChunk rows.
__global__ void __launch_bounds__(128)
interp_apply_kernel(const int32_t* __restrict__ vert, const double* __restrict__ w, int nwin,
                    int64_t n_cells, const uint8_t* __restrict__ good,
                    const __grid_constant__ FieldTable tab, double* __restrict__ out,
                    int64_t out_stride, const int32_t* __restrict__ out_index) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const int row_begin = blockIdx.y * kRowChunk;
  const int n_rows = tab.row0[tab.n_fields];
  // resolve the (field, level) of each row of this chunk once
  int fid[kRowChunk], lev[kRowChunk];
  int nr = 0;
  {
    int f = 0;
    for (int r = 0; r < kRowChunk; ++r) {
      const int row = row_begin + r;
      if (row >= n_rows) break;
      while (row >= tab.row0[f + 1]) ++f;
      fid[r] = f;
      lev[r] = row - tab.row0[f];
      nr = r + 1;
    }
  }
  double acc[kRowChunk], fine[kRowChunk];
#pragma unroll
  for (int r = 0; r < kRowChunk; ++r) acc[r] = 0.0;
  for (int k = 0; k < nwin; ++k) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int64_t e = (int64_t)(3 * k + j) * n_cells + c;
      const int32_t v = vert[e];
      const double wt = w[e];
      const bool ok = good == nullptr || good[v] != 0;
#pragma unroll
      for (int r = 0; r < kRowChunk; ++r) {
        if (r < nr) {
          const double z = ok ? field_value(tab.f[fid[r]], lev[r], v) : qnan();
          // scipy: out = 0; out += c_j * z_j  (three roundings per product/sum)
          const double prod = __dmul_rn(wt, z);
          fine[r] = j == 0 ? __dadd_rn(0.0, prod) : __dadd_rn(fine[r], prod);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kRowChunk; ++r)
      if (r < nr) acc[r] = __dadd_rn(acc[r], __dmul_rn(fine[r], tab.f[fid[r]].box_weight));
  }
  const int64_t o = out_index ? (int64_t)out_index[c] : c;
#pragma unroll
  for (int r = 0; r < kRowChunk; ++r) {
    if (r < nr) {
      double v = acc[r];
      if (tab.f[fid[r]].post == OISAT_POST_SQRT) v = sqrt(v);
      out[(int64_t)(row_begin + r) * out_stride + o] = v;
    }
  }
}

__device__ __forceinline__ int64_t reflect(int64_t i, int64_t n) {
  // scipy convolve2d boundary='symm'
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
  return i;
}

// float32 derivations live in vertical.cuh (shared with K3)
__device__ __forceinline__ double resample_source(const void* src, const void* src2, int op,
                                                  int dtype, int64_t i) {
  if (op == OISAT_SRC_VALUE) return load_as_double(src, dtype, i);
  const float dp = ((const float*)src)[i];
  if (op == OISAT_SRC_PARTIAL_COLUMN) return (double)partial_column_f32(dp, ((const float*)src2)[i]);
  return (double)air_column_f32(dp);
}

__global__ void __launch_bounds__(128)
grid_resample_kernel(const void* __restrict__ src, const void* __restrict__ src2, int op, int dtype,
                     int64_t H, int64_t W, int ky, int kx, double box_weight,
                     const int32_t* __restrict__ nn, const uint8_t* __restrict__ nn_ok,
                     int64_t n_out, double* __restrict__ out, int64_t out_stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int lev = blockIdx.y;
  double v = qnan();
  if (nn_ok == nullptr || nn_ok[i]) {
    const int64_t node = nn[i];
    const int64_t r0 = node / W, c0 = node % W;
    const int64_t base = (int64_t)lev * H * W;
    if (ky == 1 && kx == 1) {
      v = __dmul_rn(resample_source(src, src2, op, dtype, base + node), box_weight);
    } else {
      double acc = 0.0;
      for (int a = -(ky / 2); a < ky - ky / 2; ++a)
        for (int b = -(kx / 2); b < kx - kx / 2; ++b) {
          const int64_t rr = reflect(r0 + a, H), cc = reflect(c0 + b, W);
          acc = __dadd_rn(acc, __dmul_rn(resample_source(src, src2, op, dtype, base + rr * W + cc),
                                         box_weight));
        }
      v = acc;
    }
  }
  out[(int64_t)lev * out_stride + i] = v;
}

}  // namespace oisat

extern "C" int oisat_interp_apply(const int32_t* vert, const double* w, int32_t nwin,
                                  int64_t n_cells, const uint8_t* good,
                                  const oisat_field* h_fields, int32_t n_fields, double* out,
                                  int64_t out_stride, const int32_t* out_index, void* stream) {
  using namespace oisat;
  OISAT_CHECK_ARG(n_fields >= 1 && n_fields <= kMaxFields, "1..16 fields per call");
  OISAT_CHECK_ARG(nwin >= 1 && n_cells >= 0 && out_stride >= 0, "bad extent");
  if (n_cells == 0) return OISAT_OK;
  OISAT_CHECK_ARG(vert && w && out && h_fields, "null pointer");
  FieldTable tab;
  tab.n_fields = n_fields;
  int rows = 0;
  for (int i = 0; i < n_fields; ++i) {
    const oisat_field& f = h_fields[i];
    OISAT_CHECK_ARG(f.data != nullptr && f.nlev >= 1, "field without data");
    OISAT_CHECK_ARG(f.dtype == OISAT_F16 || f.dtype == OISAT_F32 || f.dtype == OISAT_F64,
                    "field dtype must be f16/f32/f64");
    tab.f[i] = f;
    tab.row0[i] = rows;
    rows += f.nlev;
  }
  tab.row0[n_fields] = rows;
  dim3 grid((unsigned)ceil_div(n_cells, 128), (unsigned)ceil_div(rows, kRowChunk));
  interp_apply_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(vert, w, nwin, n_cells, good, tab,
                                                             out, out_stride, out_index);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_grid_resample(const void* src, const void* src2, int32_t src_op,
                                   int32_t dtype, int32_t nlev, int64_t H, int64_t W, int32_t ky,
                                   int32_t kx, double box_weight, const int32_t* nn,
                                   const uint8_t* nn_ok, int64_t n_out, double* out,
                                   int64_t out_stride, void* stream) {
  using namespace oisat;
  OISAT_CHECK_ARG(dtype == OISAT_F32 || dtype == OISAT_F64, "source must be f32/f64");
  OISAT_CHECK_ARG(src_op >= OISAT_SRC_VALUE && src_op <= OISAT_SRC_AIR_COLUMN, "bad src_op");
  OISAT_CHECK_ARG(src_op == OISAT_SRC_VALUE || dtype == OISAT_F32,
                  "derived sources need float32 model fields");
  OISAT_CHECK_ARG(src_op != OISAT_SRC_PARTIAL_COLUMN || src2, "partial column needs src2");
  OISAT_CHECK_ARG(nlev >= 1 && H >= 1 && W >= 1 && ky >= 1 && kx >= 1, "bad extent");
  if (n_out == 0) return OISAT_OK;
  OISAT_CHECK_ARG(src && nn && out, "null pointer");
  dim3 grid((unsigned)ceil_div(n_out, 128), (unsigned)nlev);
  grid_resample_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(src, src2, src_op, dtype, H, W, ky,
                                                              kx, box_weight, nn, nn_ok, n_out,
                                                              out, out_stride);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
