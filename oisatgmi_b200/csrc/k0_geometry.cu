// K0: geometry predicates that the reference evaluates with a KD-tree on the host.
//
//   oisat_distmask      NOT(dists > radius) of interpolator.py:145-150,16 and
//                       filler_gosat.py:132-137,17.  The reference runs an
//                       unbounded nearest-neighbour query for every fine-grid
//                       node (219 s at TROPOMI scale) although
//                       only the predicate is consumed; here every PIXEL marks
//                       the few nodes inside its radius (idempotent byte stores,
//                       no atomics, no sort).
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
