// K0: geometry predicates that the reference evaluates with a KD-tree on the host.
//
//   oisat_distmask      NOT(dists > radius) of interpolator.py:145-150,16 and
//                       filler_gosat.py:132-137,17.  The reference runs an
//                       unbounded nearest-neighbour query for every fine-grid
//                       node (219 s at TROPOMI scale) although
//                       only the predicate is consumed; here every PIXEL marks
//                       the few nodes inside its radius (idempotent byte stores,
//                       no atomics, no sort).
//   oisat_nearest_pixel the nearest pixel of every node within the radius: what
//                       NearestNDInterpolator / cKDTree.query answer for the
//                       nearest-neighbour gridding modes (interpolator.py:17-20,
//                       28-33; types 2 and 4), restricted to the nodes that the
//                       distance mask keeps anyway.  Same pixel-centric scatter, two
//                       passes: minimum squared distance, then lowest pixel index
//                       among the pixels at that distance.
//   oisat_quality_mask  interpolator.py:126-128.
#include "common.cuh"

namespace oisat {

template <typename T>
__device__ __forceinline__ double coord_to_double(T v);
template <>
__device__ __forceinline__ double coord_to_double<float>(float v) { return (double)v; }
template <>
__device__ __forceinline__ double coord_to_double<double>(double v) { return v; }

template <typename T>
__global__ void __launch_bounds__(256)
distmask_kernel(const T* __restrict__ lon, const T* __restrict__ lat, int64_t n_px,
                const double* __restrict__ xs, int64_t W, const double* __restrict__ ys,
                int64_t H, double radius, uint8_t* __restrict__ keep) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_px) return;
  const double px = coord_to_double<T>(lon[p]);
  const double py = coord_to_double<T>(lat[p]);
  if (!(px == px) || !(py == py)) return;
  const double x0 = xs[0], y0 = ys[0];
  const double sx = W > 1 ? (xs[W - 1] - x0) / (double)(W - 1) : 1.0;
  const double sy = H > 1 ? (ys[H - 1] - y0) / (double)(H - 1) : 1.0;
  // candidate index window, one node of slack on each side; the test below is exact
  double fi0 = floor((px - radius - x0) / sx) - 1.0, fi1 = ceil((px + radius - x0) / sx) + 1.0;
  double fj0 = floor((py - radius - y0) / sy) - 1.0, fj1 = ceil((py + radius - y0) / sy) + 1.0;
  if (fi1 < 0.0 || fj1 < 0.0 || fi0 > (double)(W - 1) || fj0 > (double)(H - 1)) return;
  int64_t i0 = fi0 < 0.0 ? 0 : (int64_t)fi0, i1 = fi1 > (double)(W - 1) ? W - 1 : (int64_t)fi1;
  int64_t j0 = fj0 < 0.0 ? 0 : (int64_t)fj0, j1 = fj1 > (double)(H - 1) ? H - 1 : (int64_t)fj1;
  for (int64_t j = j0; j <= j1; ++j) {
    const double dy = ys[j] - py;
    const double dy2 = __dmul_rn(dy, dy);
    for (int64_t i = i0; i <= i1; ++i) {
      const double dx = xs[i] - px;
      // same operation order as scipy's squared-Euclidean kernel, then sqrt
      const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), dy2));
      if (d <= radius) keep[j * W + i] = 1;
    }
  }
}

// PASS 0: best_d2[node] = min over pixels of the squared distance (bit pattern of a
// non-negative double orders like the value); PASS 1: node_px[node] = lowest index of
// a pixel at exactly that distance.  Squared distance as scipy's KD-tree compares it:
// dx*dx + dy*dy in float64 without contraction.
template <typename T, int PASS>
__global__ void __launch_bounds__(256)
nearest_kernel(const T* __restrict__ lon, const T* __restrict__ lat, int64_t n_px,
               const double* __restrict__ xs, int64_t W, const double* __restrict__ ys, int64_t H,
               double radius, unsigned long long* __restrict__ best_d2,
               int32_t* __restrict__ node_px) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_px) return;
  const double px = coord_to_double<T>(lon[p]);
  const double py = coord_to_double<T>(lat[p]);
  if (!(px == px) || !(py == py)) return;
  const double x0 = xs[0], y0 = ys[0];
  const double sx = W > 1 ? (xs[W - 1] - x0) / (double)(W - 1) : 1.0;
  const double sy = H > 1 ? (ys[H - 1] - y0) / (double)(H - 1) : 1.0;
  double fi0 = floor((px - radius - x0) / sx) - 1.0, fi1 = ceil((px + radius - x0) / sx) + 1.0;
  double fj0 = floor((py - radius - y0) / sy) - 1.0, fj1 = ceil((py + radius - y0) / sy) + 1.0;
  if (fi1 < 0.0 || fj1 < 0.0 || fi0 > (double)(W - 1) || fj0 > (double)(H - 1)) return;
  int64_t i0 = fi0 < 0.0 ? 0 : (int64_t)fi0, i1 = fi1 > (double)(W - 1) ? W - 1 : (int64_t)fi1;
  int64_t j0 = fj0 < 0.0 ? 0 : (int64_t)fj0, j1 = fj1 > (double)(H - 1) ? H - 1 : (int64_t)fj1;
  for (int64_t j = j0; j <= j1; ++j) {
    const double dy = ys[j] - py;
    const double dy2 = __dmul_rn(dy, dy);
    for (int64_t i = i0; i <= i1; ++i) {
      const double dx = xs[i] - px;
      const double d2 = __dadd_rn(__dmul_rn(dx, dx), dy2);
      if (!(__dsqrt_rn(d2) <= radius)) continue;   // the distance mask's own test
      const unsigned long long bits = (unsigned long long)__double_as_longlong(d2);
      if (PASS == 0) atomicMin(&best_d2[j * W + i], bits);
      else if (best_d2[j * W + i] == bits) atomicMin(&node_px[j * W + i], (int32_t)p);
    }
  }
}

// thread = (kept cell, window node): a nearest-neighbour stencil entry is the triple
// (pixel, pixel, pixel) with weights (1, 0, 0), the shape the linear stencil has
__global__ void __launch_bounds__(256)
plan_fill_nearest_kernel(const int32_t* __restrict__ cells, int64_t n_cells,
                         const int32_t* __restrict__ window, int nwin,
                         const int32_t* __restrict__ node_px, int pair_major,
                         int32_t* __restrict__ vert, double* __restrict__ w) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_cells * nwin) return;
  const int64_t p = idx / nwin;
  const int k = (int)(idx - p * nwin);
  const int32_t f = window ? window[(int64_t)cells[p] * nwin + k] : cells[p];
  const int32_t v = node_px[f];
  const int S = 3 * nwin;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int64_t o = pair_major ? p * S + 3 * k + j : (int64_t)(3 * k + j) * n_cells + p;
    vert[o] = v;
    w[o] = j == 0 ? 1.0 : 0.0;
  }
}

__global__ void __launch_bounds__(256)
fill_i32_kernel(int32_t* __restrict__ a, int64_t n, int32_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

__global__ void __launch_bounds__(256)
quality_mask_kernel(const void* __restrict__ q, int dtype, int64_t n, double thresh,
                    uint8_t* __restrict__ good) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  good[p] = load_as_double(q, dtype, p) > thresh ? 1 : 0;
}

}  // namespace oisat

extern "C" int oisat_distmask(const void* px_lon, const void* px_lat, int32_t coord_dtype,
                              int64_t n_px, const double* xs, int64_t W, const double* ys,
                              int64_t H, double radius, uint8_t* keep, void* stream) {
  using namespace oisat;
  OISAT_CHECK_ARG(px_lon && px_lat && xs && ys && keep, "null pointer");
  OISAT_CHECK_ARG(W >= 1 && H >= 1 && n_px >= 0, "bad extent");
  OISAT_CHECK_ARG(coord_dtype == OISAT_F32 || coord_dtype == OISAT_F64, "coords must be f32/f64");
  if (n_px == 0) return OISAT_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int threads = 256;
  const unsigned blocks = (unsigned)ceil_div(n_px, threads);
  if (coord_dtype == OISAT_F32)
    distmask_kernel<float><<<blocks, threads, 0, s>>>((const float*)px_lon, (const float*)px_lat,
                                                     n_px, xs, W, ys, H, radius, keep);
  else
    distmask_kernel<double><<<blocks, threads, 0, s>>>((const double*)px_lon,
                                                      (const double*)px_lat, n_px, xs, W, ys, H,
                                                      radius, keep);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_nearest_pixel(const void* px_lon, const void* px_lat, int32_t coord_dtype,
                                   int64_t n_px, const double* xs, int64_t W, const double* ys,
                                   int64_t H, double radius, uint64_t* work, int32_t* node_px,
                                   void* stream) {
  using namespace oisat;
  OISAT_CHECK_ARG(px_lon && px_lat && xs && ys && work && node_px, "null pointer");
  OISAT_CHECK_ARG(W >= 1 && H >= 1 && n_px >= 0 && n_px < (int64_t)0x7fffffff, "bad extent");
  OISAT_CHECK_ARG(coord_dtype == OISAT_F32 || coord_dtype == OISAT_F64, "coords must be f32/f64");
  cudaStream_t s = (cudaStream_t)stream;
  OISAT_CHECK_CUDA(cudaMemsetAsync(work, 0xff, (size_t)(W * H) * sizeof(uint64_t), s));
  if (n_px == 0) return OISAT_OK;
  const unsigned blocks = (unsigned)ceil_div(n_px, 256);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(work);
  // INT32_MAX = no pixel within the radius (the convention oisat_plan_cells tests)
  fill_i32_kernel<<<(unsigned)ceil_div(W * H, 256), 256, 0, s>>>(node_px, W * H, 0x7fffffff);
  OISAT_CHECK_LAUNCH();
  if (coord_dtype == OISAT_F32) {
    const float* lo = (const float*)px_lon;
    const float* la = (const float*)px_lat;
    nearest_kernel<float, 0><<<blocks, 256, 0, s>>>(lo, la, n_px, xs, W, ys, H, radius, best, node_px);
    OISAT_CHECK_LAUNCH();
    nearest_kernel<float, 1><<<blocks, 256, 0, s>>>(lo, la, n_px, xs, W, ys, H, radius, best, node_px);
  } else {
    const double* lo = (const double*)px_lon;
    const double* la = (const double*)px_lat;
    nearest_kernel<double, 0><<<blocks, 256, 0, s>>>(lo, la, n_px, xs, W, ys, H, radius, best, node_px);
    OISAT_CHECK_LAUNCH();
    nearest_kernel<double, 1><<<blocks, 256, 0, s>>>(lo, la, n_px, xs, W, ys, H, radius, best, node_px);
  }
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_plan_fill_nearest(const int32_t* cells, int64_t n_cells, const int32_t* window,
                                       int32_t nwin, const int32_t* node_px, int32_t pair_major,
                                       int32_t* vert, double* w, void* stream) {
  using namespace oisat;
  if (n_cells <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(cells && node_px && vert && w && nwin >= 1, "null pointer");
  OISAT_CHECK_ARG(window || nwin == 1, "a window table is needed when nwin > 1");
  plan_fill_nearest_kernel<<<(unsigned)ceil_div(n_cells * nwin, 256), 256, 0,
                             (cudaStream_t)stream>>>(cells, n_cells, window, nwin, node_px,
                                                     pair_major, vert, w);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_quality_mask(const void* qflag, int32_t dtype, int64_t n_px, double thresh,
                                  uint8_t* good, void* stream) {
  using namespace oisat;
  OISAT_CHECK_ARG(qflag && good, "null pointer");
  OISAT_CHECK_ARG(dtype == OISAT_F16 || dtype == OISAT_F32 || dtype == OISAT_F64, "bad dtype");
  if (n_px <= 0) return OISAT_OK;
  quality_mask_kernel<<<(unsigned)ceil_div(n_px, 256), 256, 0, (cudaStream_t)stream>>>(
      qflag, dtype, n_px, thresh, good);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
