// K1: geometry plan on the GPU (plan builder v1).
//
//   oisat_locate      which Delaunay triangle contains each node of the working
//                     mesh -- what LinearNDInterpolator's directed walk answers
//                     point by point on the host (interpolator.py:13-15; 13 s per
//                     TROPOMI granule).  Here every TRIANGLE rasterises its own
//                     bounding box and claims the nodes it contains, with scipy's
//                     acceptance rule (all barycentric coordinates within
//                     [-eps, 1+eps], eps = 100*DBL_EPSILON, qhull._find_simplex).
//                     A node on a shared edge is claimed by the lowest triangle
//                     index; both candidates interpolate to the same value there.
//   oisat_plan_cells  per model cell: is every node of its box window located
//                     (inside the hull and within reach of a pixel)?
//   oisat_plan_fill   vertices and barycentric weights of every window node of
//                     every kept cell: the stencil K2 / the fused kernel consume.
#include <climits>

#include "common.cuh"

namespace oisat {

constexpr double kLocateEps = 100.0 * 2.220446049250313e-16;  // scipy: eps = 100 * DBL_EPSILON
constexpr int kSmallBox = 64;       // nodes a single thread rasterises on its own
constexpr int kWarpBox = 1 << 16;   // nodes a warp rasterises; larger boxes go node-centric

template <typename T>
struct Coords {
  const T* x;
  const T* y;
  __device__ __forceinline__ double px(int32_t i) const { return (double)x[i]; }
  __device__ __forceinline__ double py(int32_t i) const { return (double)y[i]; }
};

struct TriGeom {
  double x2, y2;          // reference vertex (scipy: r = last vertex)
  double t00, t01, t10, t11;  // inverse of [[x0-x2, x1-x2], [y0-y2, y1-y2]]
  bool ok;
};

__device__ __forceinline__ TriGeom tri_geom(double x0, double y0, double x1, double y1, double x2,
                                            double y2) {
  TriGeom g;
  const double a = x0 - x2, b = x1 - x2, c = y0 - y2, d = y1 - y2;
  const double det = a * d - b * c;
  const double r = 1.0 / det;
  g.x2 = x2; g.y2 = y2;
  g.t00 = d * r; g.t01 = -b * r;
  g.t10 = -c * r; g.t11 = a * r;
  g.ok = det != 0.0 && det == det;
  return g;
}

// barycentric coordinates as scipy evaluates them (qhull._barycentric_coordinates)
__device__ __forceinline__ void bary(const TriGeom& g, double qx, double qy, double* c) {
  const double d0 = qx - g.x2, d1 = qy - g.y2;
  c[0] = __dadd_rn(__dmul_rn(g.t00, d0), __dmul_rn(g.t01, d1));
  c[1] = __dadd_rn(__dmul_rn(g.t10, d0), __dmul_rn(g.t11, d1));
  c[2] = __dsub_rn(__dsub_rn(1.0, c[0]), c[1]);
}

__device__ __forceinline__ bool inside(const double* c) {
  return c[0] >= -kLocateEps && c[0] <= 1.0 + kLocateEps && c[1] >= -kLocateEps &&
         c[1] <= 1.0 + kLocateEps && c[2] >= -kLocateEps && c[2] <= 1.0 + kLocateEps;
}

struct Box {
  int i0, i1, j0, j1;
  __device__ __forceinline__ int64_t count() const {
    return (i1 < i0 || j1 < j0) ? 0 : (int64_t)(i1 - i0 + 1) * (j1 - j0 + 1);
  }
};

__device__ __forceinline__ Box node_box(double xmin, double xmax, double ymin, double ymax,
                                        const double* xs, int64_t W, const double* ys, int64_t H) {
  const double x0 = xs[0], y0 = ys[0];
  const double sx = W > 1 ? (xs[W - 1] - x0) / (double)(W - 1) : 1.0;
  const double sy = H > 1 ? (ys[H - 1] - y0) / (double)(H - 1) : 1.0;
  Box b;
  // one node of slack on each side; the inside test is what decides
  const double fi0 = floor((xmin - x0) / sx) - 1.0, fi1 = ceil((xmax - x0) / sx) + 1.0;
  const double fj0 = floor((ymin - y0) / sy) - 1.0, fj1 = ceil((ymax - y0) / sy) + 1.0;
  b.i0 = fi0 < 0.0 ? 0 : (fi0 > (double)(W - 1) ? (int)W : (int)fi0);
  b.i1 = fi1 > (double)(W - 1) ? (int)W - 1 : (fi1 < 0.0 ? -1 : (int)fi1);
  b.j0 = fj0 < 0.0 ? 0 : (fj0 > (double)(H - 1) ? (int)H : (int)fj0);
  b.j1 = fj1 > (double)(H - 1) ? (int)H - 1 : (fj1 < 0.0 ? -1 : (int)fj1);
  return b;
}

__device__ __forceinline__ void claim(const TriGeom& g, int32_t t, int i, int j,
                                      const double* xs, const double* ys, int64_t W,
                                      const uint8_t* keep, int32_t* node_tri) {
  const int64_t f = (int64_t)j * W + i;
  if (!keep[f]) return;
  double c[3];
  bary(g, xs[i], ys[j], c);
  if (inside(c)) atomicMin(&node_tri[f], t);
}

// Pass 1, thread = triangle.  A triangle whose bounding box holds few mesh nodes
// rasterises it.  Mid-sized boxes (high-latitude pixels are many mesh nodes wide,
// hull-pocket slivers are long) are queued for pass 2, a warp each; the few
// boxes of up to the whole mesh (date-line crossers) are queued for pass 3.
template <typename T>
__global__ void __launch_bounds__(256)
locate_kernel(const int32_t* __restrict__ tri, int64_t n_tri, Coords<T> P,
              const double* __restrict__ xs, int64_t W, const double* __restrict__ ys, int64_t H,
              const uint8_t* __restrict__ keep, int32_t* __restrict__ node_tri,
              int32_t* __restrict__ counts, int32_t* __restrict__ mid_list,
              int32_t* __restrict__ big_list) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tri) return;
  const int32_t v0 = tri[3 * t], v1 = tri[3 * t + 1], v2 = tri[3 * t + 2];
  const double x0 = P.px(v0), y0 = P.py(v0), x1 = P.px(v1), y1 = P.py(v1), x2 = P.px(v2),
               y2 = P.py(v2);
  const TriGeom g = tri_geom(x0, y0, x1, y1, x2, y2);
  if (!g.ok) return;
  const Box b = node_box(fmin(x0, fmin(x1, x2)), fmax(x0, fmax(x1, x2)), fmin(y0, fmin(y1, y2)),
                         fmax(y0, fmax(y1, y2)), xs, W, ys, H);
  const int64_t cnt = b.count();
  if (cnt == 0) return;
  if (cnt > kWarpBox) {
    big_list[atomicAdd(&counts[1], 1)] = (int32_t)t;
    return;
  }
  if (cnt > kSmallBox) {
    mid_list[atomicAdd(&counts[0], 1)] = (int32_t)t;
    return;
  }
  for (int j = b.j0; j <= b.j1; ++j)
    for (int i = b.i0; i <= b.i1; ++i) claim(g, (int32_t)t, i, j, xs, ys, W, keep, node_tri);
}

// Range of mesh columns the triangle can touch on the mesh row at height y: the crossings of
// the three edges with that line (a vertex within rounding of the line counts with its own x),
// one node of slack on each side; empty when the line misses the triangle.  The inside test
// decides, this only says where it can succeed: a node further out than the slack fails it.
__device__ __forceinline__ bool row_span(const double* vx, const double* vy, double y, double x0,
                                         double sx, int i_lo, int i_hi, int* ia, int* ib) {
  const double tol = 1e-9 * (fabs(y) + 1.0);
  double xl = CUDART_INF, xr = -CUDART_INF;
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const double ax = vx[e], ay = vy[e], bx = vx[(e + 1) % 3], by = vy[(e + 1) % 3];
    if (fabs(y - ay) <= tol) { xl = fmin(xl, ax); xr = fmax(xr, ax); }
    const double lo = fmin(ay, by), hi = fmax(ay, by);
    if (y < lo - tol || y > hi + tol || hi == lo) continue;
    double x = ax + (y - ay) / (by - ay) * (bx - ax);
    x = fmin(fmax(x, fmin(ax, bx)), fmax(ax, bx));
    xl = fmin(xl, x);
    xr = fmax(xr, x);
  }
  if (!(xl <= xr)) return false;
  const double fa = floor((xl - x0) / sx) - 1.0, fb = ceil((xr - x0) / sx) + 1.0;
  *ia = fa < (double)i_lo ? i_lo : (int)fa;
  *ib = fb > (double)i_hi ? i_hi : (int)fb;
  return *ia <= *ib;
}

// Pass 2, warp = queued mid-sized triangle (grid-stride over the queue), lane = mesh row of
// its box, and on a row only the columns between the triangle's own edges are tested.  The
// queue is mostly slivers -- the hull pockets of a swath are closed by triangles a few nodes
// thin and hundreds long, whose boxes hold 23-32 M nodes per OMI granule against 5.6 M for all
// the small triangles together; testing whole boxes made this pass 0.57 ms of the 0.8 ms a
// granule's plan costs on the device.
template <typename T>
__global__ void __launch_bounds__(256)
locate_mid_kernel(const int32_t* __restrict__ tri, Coords<T> P, const double* __restrict__ xs,
                  int64_t W, const double* __restrict__ ys, int64_t H,
                  const uint8_t* __restrict__ keep, int32_t* __restrict__ node_tri,
                  const int32_t* __restrict__ counts, const int32_t* __restrict__ mid_list) {
  const int n_mid = counts[0];
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const double x_first = xs[0];
  const double sx = W > 1 ? (xs[W - 1] - x_first) / (double)(W - 1) : 1.0;
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n_mid; k += warps) {
    const int32_t t = mid_list[k];
    const int32_t v0 = tri[3 * (int64_t)t], v1 = tri[3 * (int64_t)t + 1], v2 = tri[3 * (int64_t)t + 2];
    const double vx[3] = {P.px(v0), P.px(v1), P.px(v2)}, vy[3] = {P.py(v0), P.py(v1), P.py(v2)};
    const TriGeom g = tri_geom(vx[0], vy[0], vx[1], vy[1], vx[2], vy[2]);
    const Box b = node_box(fmin(vx[0], fmin(vx[1], vx[2])), fmax(vx[0], fmax(vx[1], vx[2])),
                           fmin(vy[0], fmin(vy[1], vy[2])), fmax(vy[0], fmax(vy[1], vy[2])), xs, W, ys, H);
    for (int j = b.j0 + lane; j <= b.j1; j += 32) {
      int ia, ib;
      if (!row_span(vx, vy, ys[j], x_first, sx, b.i0, b.i1, &ia, &ib)) continue;
      for (int i = ia; i <= ib; ++i) claim(g, t, i, j, xs, ys, W, keep, node_tri);
    }
  }
}

// Pass 3, thread = mesh node.  The few kept nodes that no rasterised triangle claimed
// (they sit in a hull pocket, or outside the hull) are tested against the short
// list of large triangles; node-centric, so the large boxes are never rasterised.
// A block that has such a node works out the triangles' geometry once, in shared memory
// (a chunk of 256 triangles at a time), instead of every node re-deriving it from the
// vertex table (three dependent loads and a division per triangle and node).
template <typename T>
__global__ void __launch_bounds__(256)
locate_big_kernel(const int32_t* __restrict__ tri, Coords<T> P, const double* __restrict__ xs,
                  int64_t W, const double* __restrict__ ys, int64_t H,
                  const uint8_t* __restrict__ keep, int32_t* __restrict__ node_tri,
                  const int32_t* __restrict__ counts, const int32_t* __restrict__ big_list) {
  __shared__ TriGeom geom[256];
  __shared__ int32_t index[256];
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool open = f < W * H && keep[f] && node_tri[f] == INT_MAX;
  if (!__syncthreads_or(open)) return;
  const int n_big = counts[1];
  const double qx = open ? xs[f % W] : 0.0, qy = open ? ys[f / W] : 0.0;
  int32_t best = INT_MAX;
  for (int k0 = 0; k0 < n_big; k0 += 256) {
    const int n = n_big - k0 < 256 ? n_big - k0 : 256;
    __syncthreads();
    if ((int)threadIdx.x < n) {
      const int32_t t = big_list[k0 + threadIdx.x];
      const int32_t v0 = tri[3 * (int64_t)t], v1 = tri[3 * (int64_t)t + 1], v2 = tri[3 * (int64_t)t + 2];
      geom[threadIdx.x] = tri_geom(P.px(v0), P.py(v0), P.px(v1), P.py(v1), P.px(v2), P.py(v2));
      index[threadIdx.x] = t;
    }
    __syncthreads();
    if (open)
      for (int k = 0; k < n; ++k) {
        double c[3];
        bary(geom[k], qx, qy, c);
        if (inside(c) && index[k] < best) best = index[k];
      }
  }
  if (best != INT_MAX) node_tri[f] = best;
}

__global__ void __launch_bounds__(256)
plan_cells_kernel(const int32_t* __restrict__ window, int nwin, const uint8_t* __restrict__ nn_ok,
                  int64_t n_cell, const int32_t* __restrict__ node_tri,
                  uint8_t* __restrict__ cell_ok) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cell) return;
  bool ok = nn_ok == nullptr || nn_ok[c] != 0;
  for (int k = 0; ok && k < nwin; ++k) ok = node_tri[window[c * nwin + k]] != INT_MAX;
  cell_ok[c] = ok ? 1 : 0;
}

// thread = (kept cell, window node)
template <typename T>
__global__ void __launch_bounds__(256)
plan_fill_kernel(const int32_t* __restrict__ cells, int64_t n_cells,
                 const int32_t* __restrict__ window, int nwin,
                 const int32_t* __restrict__ node_tri, const int32_t* __restrict__ tri,
                 Coords<T> P, const double* __restrict__ xs, int64_t W,
                 const double* __restrict__ ys, int pair_major, int32_t* __restrict__ vert,
                 double* __restrict__ w) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_cells * nwin) return;
  const int64_t p = idx / nwin;
  const int k = (int)(idx - p * nwin);
  const int32_t f = window[(int64_t)cells[p] * nwin + k];
  const int32_t t = node_tri[f];
  const int32_t v[3] = {tri[3 * (int64_t)t], tri[3 * (int64_t)t + 1], tri[3 * (int64_t)t + 2]};
  const TriGeom g = tri_geom(P.px(v[0]), P.py(v[0]), P.px(v[1]), P.py(v[1]), P.px(v[2]),
                             P.py(v[2]));
  double c[3];
  bary(g, xs[f % W], ys[f / W], c);
  const int S = 3 * nwin;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int64_t o = pair_major ? p * S + 3 * k + j : (int64_t)(3 * k + j) * n_cells + p;
    vert[o] = v[j];
    w[o] = c[j];
  }
}

// Near ties of a triangulation (the scan of count_near_ties in delaunay.cpp, same
// expressions in the same order -- both trees are compiled without FMA contraction):
// one thread per half-edge, the lower-numbered twin tests the quadrilateral.
template <typename T>
__global__ void __launch_bounds__(256)
near_ties_kernel(const int32_t* __restrict__ tri, const int32_t* __restrict__ half, int64_t n_half,
                 Coords<T> P, double tol, unsigned long long* __restrict__ ties,
                 uint8_t* __restrict__ tri_flag) {
  const int64_t a64 = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (a64 >= n_half) return;
  const int32_t a = (int32_t)a64;
  const int32_t b = half[a];
  if (b < a) return;  // hull edge (-1) or the twin does it
  const int32_t a0 = a - a % 3, b0 = b - b % 3;
  const int32_t p0 = tri[a0 + (a + 2) % 3], pr = tri[a], pl = tri[a0 + (a + 1) % 3],
                p1 = tri[b0 + (b + 2) % 3];
  const double x0 = P.px(p0), y0 = P.py(p0), xl = P.px(pl), yl = P.py(pl), xr = P.px(pr),
               yr = P.py(pr), x1 = P.px(p1), y1 = P.py(p1);
  const double adx = x0 - x1, ady = y0 - y1, bdx = xl - x1, bdy = yl - y1, cdx = xr - x1,
               cdy = yr - y1;
  const double det = (adx * adx + ady * ady) * (bdx * cdy - cdx * bdy) +
                     (bdx * bdx + bdy * bdy) * (cdx * ady - adx * cdy) +
                     (cdx * cdx + cdy * cdy) * (adx * bdy - bdx * ady);
  const double area2 = fabs((x0 - xr) * (yl - yr) - (y0 - yr) * (xl - xr));
  if (fabs(det) <= tol * area2) {
    atomicAdd(ties, 1ull);
    if (tri_flag) {   // both triangles of the quadrilateral: Qhull may split it the other way
      tri_flag[a / 3] = 1;
      tri_flag[b / 3] = 1;
    }
  }
}

// mesh nodes that K1 placed in a triangle next to a (near-)tied edge: the only nodes whose
// stencil could differ from the one Qhull's triangulation gives
__global__ void __launch_bounds__(256)
flagged_nodes_kernel(const int32_t* __restrict__ node_tri, int64_t n_nodes,
                     const uint8_t* __restrict__ tri_flag, unsigned long long* __restrict__ count) {
  const int64_t f = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (f >= n_nodes) return;
  const int32_t t = node_tri[f];
  if (t != INT_MAX && tri_flag[t]) atomicAdd(count, 1ull);
}

}  // namespace oisat

using namespace oisat;

extern "C" int oisat_flagged_nodes(const int32_t* node_tri, int64_t n_nodes, const uint8_t* tri_flag,
                                   uint64_t* n_flagged, void* stream) {
  OISAT_CHECK_ARG(n_flagged != nullptr, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  OISAT_CHECK_CUDA(cudaMemsetAsync(n_flagged, 0, sizeof(uint64_t), s));
  if (n_nodes <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(node_tri && tri_flag, "null pointer");
  flagged_nodes_kernel<<<(unsigned)ceil_div(n_nodes, 256), 256, 0, s>>>(
      node_tri, n_nodes, tri_flag, (unsigned long long*)n_flagged);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_near_ties(const int32_t* tri, const int32_t* half, int64_t n_tri, const void* px,
                               const void* py, int32_t coord_dtype, double max_abs_coord,
                               uint64_t* n_ties, uint8_t* tri_flag, void* stream) {
  OISAT_CHECK_ARG(n_ties != nullptr, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  OISAT_CHECK_CUDA(cudaMemsetAsync(n_ties, 0, sizeof(uint64_t), s));
  if (n_tri <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(tri && half && px && py, "null pointer");
  OISAT_CHECK_ARG(coord_dtype == OISAT_F32 || coord_dtype == OISAT_F64, "coords must be f32/f64");
  OISAT_CHECK_ARG(3 * n_tri < (int64_t)INT_MAX, "bad extent");
  const double m2 = max_abs_coord * max_abs_coord;
  const double tol = 2e-14 * (m2 > 1e-300 ? m2 : 1e-300);
  const unsigned blocks = (unsigned)ceil_div(3 * n_tri, 256);
  if (coord_dtype == OISAT_F32)
    near_ties_kernel<float><<<blocks, 256, 0, s>>>(tri, half, 3 * n_tri,
                                                   Coords<float>{(const float*)px, (const float*)py},
                                                   tol, (unsigned long long*)n_ties, tri_flag);
  else
    near_ties_kernel<double><<<blocks, 256, 0, s>>>(tri, half, 3 * n_tri,
                                                    Coords<double>{(const double*)px, (const double*)py},
                                                    tol, (unsigned long long*)n_ties, tri_flag);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_locate(const int32_t* tri, int64_t n_tri, const void* px, const void* py,
                            int32_t coord_dtype, const double* xs, int64_t W, const double* ys,
                            int64_t H, const uint8_t* keep, int32_t* node_tri, int32_t* work,
                            void* stream) {
  if (n_tri <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(tri && px && py && xs && ys && keep && node_tri && work, "null pointer");
  OISAT_CHECK_ARG(coord_dtype == OISAT_F32 || coord_dtype == OISAT_F64, "coords must be f32/f64");
  OISAT_CHECK_ARG(W >= 1 && H >= 1 && W * H < (int64_t)INT_MAX && n_tri < (int64_t)INT_MAX,
                  "bad extent");
  const unsigned blocks = (unsigned)ceil_div(n_tri, 256);
  const unsigned nblocks = (unsigned)ceil_div(W * H, 256);
  cudaStream_t s = (cudaStream_t)stream;
  int32_t* counts = work;          // work = [n_mid][n_big][mid list of n_tri][big list of n_tri]
  int32_t* mid_list = work + 2;
  int32_t* big_list = work + 2 + n_tri;
  const unsigned mid_blocks = 148 * 8;
  OISAT_CHECK_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), s));
  if (coord_dtype == OISAT_F32) {
    const Coords<float> P{(const float*)px, (const float*)py};
    locate_kernel<float><<<blocks, 256, 0, s>>>(tri, n_tri, P, xs, W, ys, H, keep, node_tri,
                                               counts, mid_list, big_list);
    OISAT_CHECK_LAUNCH();
    locate_mid_kernel<float><<<mid_blocks, 256, 0, s>>>(tri, P, xs, W, ys, H, keep, node_tri,
                                                       counts, mid_list);
    OISAT_CHECK_LAUNCH();
    locate_big_kernel<float><<<nblocks, 256, 0, s>>>(tri, P, xs, W, ys, H, keep, node_tri,
                                                    counts, big_list);
  } else {
    const Coords<double> P{(const double*)px, (const double*)py};
    locate_kernel<double><<<blocks, 256, 0, s>>>(tri, n_tri, P, xs, W, ys, H, keep, node_tri,
                                                counts, mid_list, big_list);
    OISAT_CHECK_LAUNCH();
    locate_mid_kernel<double><<<mid_blocks, 256, 0, s>>>(tri, P, xs, W, ys, H, keep, node_tri,
                                                        counts, mid_list);
    OISAT_CHECK_LAUNCH();
    locate_big_kernel<double><<<nblocks, 256, 0, s>>>(tri, P, xs, W, ys, H, keep, node_tri,
                                                     counts, big_list);
  }
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_plan_cells(const int32_t* window, int32_t nwin, const uint8_t* nn_ok,
                                int64_t n_cell, const int32_t* node_tri, uint8_t* cell_ok,
                                void* stream) {
  if (n_cell <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(window && node_tri && cell_ok && nwin >= 1, "null pointer");
  plan_cells_kernel<<<(unsigned)ceil_div(n_cell, 256), 256, 0, (cudaStream_t)stream>>>(
      window, nwin, nn_ok, n_cell, node_tri, cell_ok);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_plan_fill(const int32_t* cells, int64_t n_cells, const int32_t* window,
                               int32_t nwin, const int32_t* node_tri, const int32_t* tri,
                               const void* px, const void* py, int32_t coord_dtype,
                               const double* xs, int64_t W, const double* ys, int32_t pair_major,
                               int32_t* vert, double* w, void* stream) {
  if (n_cells <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(cells && window && node_tri && tri && px && py && xs && ys && vert && w,
                  "null pointer");
  OISAT_CHECK_ARG(coord_dtype == OISAT_F32 || coord_dtype == OISAT_F64, "coords must be f32/f64");
  const unsigned blocks = (unsigned)ceil_div(n_cells * nwin, 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (coord_dtype == OISAT_F32)
    plan_fill_kernel<float><<<blocks, 256, 0, s>>>(
        cells, n_cells, window, nwin, node_tri, tri,
        Coords<float>{(const float*)px, (const float*)py}, xs, W, ys, pair_major, vert, w);
  else
    plan_fill_kernel<double><<<blocks, 256, 0, s>>>(
        cells, n_cells, window, nwin, node_tri, tri,
        Coords<double>{(const double*)px, (const double*)py}, xs, W, ys, pair_major, vert, w);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
