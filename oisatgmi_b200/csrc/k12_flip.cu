// K12: the Delaunay triangulation of a swath finished on the device.
//
// The host part (oisat_h_delaunay_seed_parts, delaunay_seed.inl) classifies the lattice quads
// and triangulates what the lattice does not cover -- the pockets between the curved outline
// and the convex hull, the lips of a date-line tear: a few thousand triangles.  Here:
//   oisat_seed_assemble   the two triangles of every seeded quad and all twin pointers, one
//                         thread per quad (what SeedBuilder::assemble does serially: 5 of the
//                         host's 12 ms per granule, plus the upload of 4.7 MB of triangles);
//   oisat_flip_delaunay   Lawson's flips as rounds of independent flips (flip_rounds.h) in ONE
//                         cooperative launch: mark | grid barrier | apply | grid barrier, until
//                         a round flips nothing, then the check of every edge.
// The triangles never visit the host: the near-tie scan, K1 and the stencil fill read them
// where they are.  The predicate is the floating-point filter; an edge it cannot decide is
// counted and the caller gives that granule to the exact host builder.
#include <cooperative_groups.h>
#include <cstdlib>

#include "common.cuh"
#include "flip_rounds.h"

namespace cg = cooperative_groups;

namespace oisat {
namespace {

template <typename T>
struct FlipCoords {
  const T* x;
  const T* y;
  __device__ __forceinline__ double operator()(int v, int axis) const {
    return (double)(axis ? y[v] : x[v]);
  }
};

struct DeviceOps {
  __device__ __forceinline__ unsigned long long max(unsigned long long* p, unsigned long long v) const {
    return atomicMax(p, v);
  }
  // list appends: the threads of a warp that arrive here together reserve their slots with ONE
  // atomic (every caller adds 1 to the same counter of the round).  Millions of returning
  // atomics on a single address were most of the kernel's time: they are served one after
  // the other by one L2 slice.
  __device__ __forceinline__ unsigned int add(unsigned int* p, unsigned int v) const {
    const cg::coalesced_group g = cg::coalesced_threads();
    unsigned int base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(p, v * g.size());
    return g.shfl(base, 0) + v * g.thread_rank();
  }
  __device__ __forceinline__ int32_t exch(int32_t* p, int32_t v) const { return atomicExch(p, v); }
};

// half-edge of the seeded quad whose first triangle is t on the given side
// (0 top: row r, 1 left: column c, 2 right, 3 bottom); see SeedBuilder::slot
__device__ __forceinline__ int32_t quad_slot(int32_t t, int side, int sigma) {
  const int32_t t0 = 3 * t, t1 = t0 + 3;
  if (sigma > 0) return side == 0 ? t0 : side == 1 ? t0 + 2 : side == 2 ? t1 : t1 + 1;
  return side == 0 ? t0 + 2 : side == 1 ? t0 : side == 2 ? t1 + 2 : t1 + 1;
}

__global__ void __launch_bounds__(256)
seed_assemble_kernel(const int32_t* __restrict__ qtri, int64_t rows, int64_t cols, int sigma,
                     const int32_t* __restrict__ otri, const int32_t* __restrict__ ohalf,
                     int64_t n_out3, int64_t base, int32_t* __restrict__ tri,
                     int32_t* __restrict__ half) {
  const int64_t qc = cols - 1, nq = (rows - 1) * qc;
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < nq) {
    const int32_t t = qtri[i];
    if (t < 0) return;
    const int64_t r = i / qc, c = i - r * qc;
    const int32_t a = (int32_t)(r * cols + c), b = a + 1, cc = a + (int32_t)cols, d = cc + 1;
    const int32_t t0 = 3 * t, t1 = t0 + 3;
    if (sigma > 0) {
      tri[t0] = a; tri[t0 + 1] = b; tri[t0 + 2] = cc;
      tri[t1] = b; tri[t1 + 1] = d; tri[t1 + 2] = cc;
      half[t0 + 1] = t1 + 2;
      half[t1 + 2] = t0 + 1;
    } else {
      tri[t0] = a; tri[t0 + 1] = cc; tri[t0 + 2] = b;
      tri[t1] = b; tri[t1 + 1] = cc; tri[t1 + 2] = d;
      half[t0 + 1] = t1;
      half[t1] = t0 + 1;
    }
    // every thread writes its own four outer slots; a side without a seeded neighbour is a
    // seam edge, written by the thread of the outside triangle across it (or stays -1: hull)
    int32_t u;
    if (r > 0 && (u = qtri[i - qc]) >= 0) half[quad_slot(t, 0, sigma)] = quad_slot(u, 3, sigma);
    if (c > 0 && (u = qtri[i - 1]) >= 0) half[quad_slot(t, 1, sigma)] = quad_slot(u, 2, sigma);
    if (c + 1 < qc && (u = qtri[i + 1]) >= 0) half[quad_slot(t, 2, sigma)] = quad_slot(u, 1, sigma);
    if (r + 2 < rows && (u = qtri[i + qc]) >= 0) half[quad_slot(t, 3, sigma)] = quad_slot(u, 0, sigma);
    return;
  }
  const int64_t j = i - nq;
  if (j >= n_out3) return;
  tri[base + j] = otri[j];
  const int32_t g = ohalf[j];
  half[base + j] = g;
  if (g >= 0 && g < base) half[g] = (int32_t)(base + j);
}

// A batch of meshes (the granules of a day) goes through the rounds together: a round costs
// two barriers and a chain of dependent loads whatever the number of edges in it, so fifteen
// granules cost what one costs.  The table is a kernel parameter (constant bank).
constexpr int kMaxBatch = 32;
constexpr int kMaxRounds = 4096;
constexpr int kFlipUnroll = 1;      // 4 measured slower (3.22 vs 3.00 ms per day batch): the step is not latency bound per thread
constexpr int kFlipThreads = 1024;   // one block per SM: 64 registers per thread

struct Batch {
  int n;
  int32_t tri_end[kMaxBatch];        // running total of triangles: mesh g owns [tri_end[g-1], tri_end[g])
  oisat_flip::Mesh mesh[kMaxBatch];
  const void* x[kMaxBatch];
  const void* y[kMaxBatch];
};

__device__ __forceinline__ int mesh_of(const Batch& b, int32_t t) {
  int g = 0;
  while (g + 1 < b.n && t >= b.tri_end[g]) ++g;
  return g;
}

// The rounds.  While a round has more than `tail` triangles to look at, the whole grid works
// on it (two grid barriers per round); the long tail of short rounds -- an OMI granule needs
// ~80-100 rounds, most of them with fewer than 500 flips -- is run by block 0 alone with block
// barriers, the other blocks leave.  result[0] = rounds run, result[1] = flips (whole batch).
template <typename T>
__global__ void __launch_bounds__(kFlipThreads)
flip_rounds_kernel(const __grid_constant__ Batch b, oisat_flip::Lists l, int max_rounds,
                   unsigned int tail, unsigned long long* __restrict__ result) {
  cg::grid_group grid = cg::this_grid();
  int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const DeviceOps ops;
  const int64_t n_tri = b.tri_end[b.n - 1];
  bool solo = gridDim.x == 1;
  int round = 0;
  unsigned int flips = 0;
  for (; round < max_rounds; ++round) {
    const int64_t n_items = round == 0 ? n_tri : (int64_t)*(volatile unsigned int*)&l.n_listed[round - 1];
    const int32_t* list = (round & 1) ? l.tri_list[1] : l.tri_list[0];
    for (int64_t i0 = tid; i0 < 3 * n_items; i0 += kFlipUnroll * nthreads) {
      oisat_flip::Quad q[kFlipUnroll];
      int g[kFlipUnroll];
#pragma unroll
      for (int u = 0; u < kFlipUnroll; ++u) {
        const int64_t i = i0 + u * nthreads;
        q[u].ia = -1;
        g[u] = 0;
        if (i < 3 * n_items) {
          const int64_t k = i / 3;
          const int32_t t = round == 0 ? (int32_t)k : list[k];
          g[u] = mesh_of(b, t);
          const oisat_flip::Mesh& m = b.mesh[g[u]];
          q[u] = oisat_flip::mark_prepare(m, 3 * (t - m.tri_base) + (int32_t)(i - 3 * k), round);
        }
      }
      bool go[kFlipUnroll];
#pragma unroll
      for (int u = 0; u < kFlipUnroll; ++u) {
        const FlipCoords<T> P{(const T*)b.x[g[u]], (const T*)b.y[g[u]]};
        go[u] = q[u].ia >= 0 && oisat_flip::mark_decide(q[u], P);
      }
#pragma unroll
      for (int u = 0; u < kFlipUnroll; ++u)
        if (go[u]) oisat_flip::mark_claim(b.mesh[g[u]], l, q[u], round, ops);
    }
    if (solo) __syncthreads(); else grid.sync();
    const int64_t n_marked = *(volatile unsigned int*)&l.n_marked[round];
    if (n_marked == 0) { ++round; break; }
    for (int64_t i = tid; i < n_marked; i += nthreads) {
      const int32_t e = l.edge_list[i];
      const oisat_flip::Mesh& m = b.mesh[mesh_of(b, e / 3)];
      flips += oisat_flip::apply_edge(m, l, e - 3 * m.tri_base, round, ops);
    }
    if (solo) __syncthreads(); else grid.sync();
    if (!solo && *(volatile unsigned int*)&l.n_listed[round] <= tail) {
      if (blockIdx.x != 0) break;
      solo = true;
      tid = threadIdx.x;
      nthreads = blockDim.x;
    }
  }
  flips = __reduce_add_sync(0xffffffffu, flips);
  if ((threadIdx.x & 31) == 0 && flips) atomicAdd(&result[1], (unsigned long long)flips);
  if (blockIdx.x == 0 && threadIdx.x == 0) result[0] = (unsigned long long)round;
}

// every edge once more, per mesh g: result[4 g + 2] = edges certainly not Delaunay (only when
// the rounds were cut off), result[4 g + 3] = edges the filter cannot decide; the batch's
// rounds and flips are copied to every mesh's result[4 g + 0 / 1]
template <typename T>
__global__ void __launch_bounds__(256)
flip_check_kernel(const __grid_constant__ Batch b, unsigned long long* __restrict__ result) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t n_half = 3 * (int64_t)b.tri_end[b.n - 1];
  if (e < b.n && e > 0) {
    result[4 * e] = result[0];
    result[4 * e + 1] = result[1];
  }
  if (e >= n_half) return;
  const int g = mesh_of(b, (int32_t)(e / 3));
  const oisat_flip::Mesh& m = b.mesh[g];
  const FlipCoords<T> P{(const T*)b.x[g], (const T*)b.y[g]};
  const int c = oisat_flip::check_edge(m.tri, m.half, e - 3 * (int64_t)m.tri_base, P);
  if (c & 1) atomicAdd(&result[4 * g + 2], 1ull);
  if (c & 2) atomicAdd(&result[4 * g + 3], 1ull);
}

}  // namespace
}  // namespace oisat

using namespace oisat;

extern "C" int oisat_seed_assemble(const int32_t* qtri, int64_t n_rows, int64_t n_cols,
                                   int32_t sigma, int64_t n_quads, const int32_t* otri,
                                   const int32_t* ohalf, int64_t n_outside, int32_t* tri,
                                   int32_t* half, void* stream) {
  OISAT_CHECK_ARG(qtri && tri && half && (n_outside == 0 || (otri && ohalf)), "null pointer");
  OISAT_CHECK_ARG(n_rows >= 2 && n_cols >= 2 && n_quads >= 1 && n_outside >= 0 &&
                  (sigma == 1 || sigma == -1), "bad extent");
  const int64_t n_tri = 2 * n_quads + n_outside;
  OISAT_CHECK_ARG(3 * n_tri < (int64_t)INT_MAX && n_rows * n_cols < (int64_t)INT_MAX, "bad extent");
  cudaStream_t s = (cudaStream_t)stream;
  OISAT_CHECK_CUDA(cudaMemsetAsync(half, 0xff, (size_t)(3 * n_tri) * sizeof(int32_t), s));
  const int64_t work = (n_rows - 1) * (n_cols - 1) + 3 * n_outside;
  seed_assemble_kernel<<<(unsigned)ceil_div(work, 256), 256, 0, s>>>(
      qtri, n_rows, n_cols, sigma, otri, ohalf, 3 * n_outside, 6 * n_quads, tri, half);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

// workspace for N triangles in all: [owner u64 N][cand i32 3N][n_listed u32 R][n_marked u32 R]
// (zeroed) | [stamp i32 N] (set to -1) | [tri_list 2 x i32 N][edge_list i32 3N/2+1]
extern "C" int64_t oisat_flip_workspace_bytes(int64_t n_tri) {
  if (n_tri < 0) return OISAT_E_ARG;
  return 8 * n_tri + 12 * n_tri + 8 * (int64_t)kMaxRounds + 4 * n_tri + 8 * n_tri +
         4 * (3 * n_tri / 2 + 1) + 64;
}

namespace {
int flip_chunk(const oisat_flip_item* items, int n_items, int32_t coord_dtype, void* workspace,
               uint64_t* result, cudaStream_t s) {
  Batch b;
  b.n = n_items;
  int64_t N = 0;
  for (int g = 0; g < n_items; ++g) N += items[g].n_tri;
  OISAT_CHECK_ARG(3 * N < (int64_t)INT_MAX, "bad extent");
  char* w = (char*)workspace;
  const int64_t zeroed = 20 * N + 8 * (int64_t)kMaxRounds;
  unsigned long long* owner = (unsigned long long*)w;
  int32_t* cand = (int32_t*)(w + 8 * N);
  oisat_flip::Lists l;
  l.n_listed = (unsigned int*)(w + 20 * N);
  l.n_marked = l.n_listed + kMaxRounds;
  int32_t* stamp = (int32_t*)(w + zeroed);
  l.tri_list[0] = stamp + N;
  l.tri_list[1] = stamp + 2 * N;
  l.edge_list = stamp + 3 * N;
  int64_t base = 0;
  for (int g = 0; g < n_items; ++g) {
    oisat_flip::Mesh& m = b.mesh[g];
    m.tri = items[g].tri;
    m.half = items[g].half;
    m.n_half = 3 * items[g].n_tri;
    m.stamp = stamp + base;
    m.cand = cand + 3 * base;
    m.owner = owner + base;
    m.tri_base = (int32_t)base;
    b.x[g] = items[g].px;
    b.y[g] = items[g].py;
    base += items[g].n_tri;
    b.tri_end[g] = (int32_t)base;
  }
  OISAT_CHECK_CUDA(cudaMemsetAsync(w, 0, (size_t)zeroed, s));
  OISAT_CHECK_CUDA(cudaMemsetAsync(stamp, 0xff, (size_t)(4 * N), s));
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0, per_sm_f = 0, per_sm_d = 0;
    OISAT_CHECK_CUDA(cudaGetDevice(&dev));
    OISAT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f, flip_rounds_kernel<float>, kFlipThreads, 0));
    OISAT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_d, flip_rounds_kernel<double>, kFlipThreads, 0));
    OISAT_CHECK_ARG(per_sm_f >= 1 && per_sm_d >= 1, "flip kernel does not fit an SM");
    OISAT_CHECK_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  int max_rounds = kMaxRounds;
  if (const char* e = getenv("OISAT_FLIP_MAX_ROUNDS")) {   // profiling: cost of the first k rounds
    max_rounds = atoi(e);
    if (max_rounds < 1 || max_rounds > kMaxRounds) max_rounds = kMaxRounds;
  }
  unsigned int tail = 512;
  if (const char* e = getenv("OISAT_FLIP_TAIL")) tail = (unsigned int)atoi(e);
  unsigned long long* res = (unsigned long long*)result;
  // one block per SM at most (a grid barrier costs with the number of blocks), and no more
  // blocks than there is work for in the first round
  int64_t blocks = sm_count;
  const int64_t want = ceil_div(3 * N, 4 * kFlipThreads);
  if (blocks > want) blocks = want;
  void* args[] = {&b, &l, &max_rounds, &tail, &res};
  const void* fn = coord_dtype == OISAT_F32 ? (const void*)flip_rounds_kernel<float>
                                            : (const void*)flip_rounds_kernel<double>;
  OISAT_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)blocks), dim3(kFlipThreads), args, 0, s));
  OISAT_CHECK_LAUNCH();
  const unsigned cblocks = (unsigned)ceil_div(3 * N, 256);
  if (coord_dtype == OISAT_F32)
    flip_check_kernel<float><<<cblocks, 256, 0, s>>>(b, res);
  else
    flip_check_kernel<double><<<cblocks, 256, 0, s>>>(b, res);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
}  // namespace

extern "C" int oisat_flip_delaunay_batch(const oisat_flip_item* items, int32_t n_items,
                                         int32_t coord_dtype, void* workspace, uint64_t* result,
                                         void* stream) {
  if (n_items <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(items && result && workspace, "null pointer");
  OISAT_CHECK_ARG(coord_dtype == OISAT_F32 || coord_dtype == OISAT_F64, "coords must be f32/f64");
  cudaStream_t s = (cudaStream_t)stream;
  OISAT_CHECK_CUDA(cudaMemsetAsync(result, 0, (size_t)n_items * 4 * sizeof(uint64_t), s));
  for (int g = 0; g < n_items; ++g)
    OISAT_CHECK_ARG(items[g].n_tri >= 1 && items[g].tri && items[g].half && items[g].px && items[g].py,
                    "bad item");
  for (int g0 = 0; g0 < n_items; g0 += kMaxBatch) {
    const int n = n_items - g0 < kMaxBatch ? n_items - g0 : kMaxBatch;
    const int rc = flip_chunk(items + g0, n, coord_dtype, workspace, result + 4 * g0, s);
    if (rc != OISAT_OK) return rc;
  }
  return OISAT_OK;
}

extern "C" int oisat_flip_delaunay(int32_t* tri, int32_t* half, int64_t n_tri, const void* px,
                                   const void* py, int32_t coord_dtype, void* workspace,
                                   uint64_t* result, void* stream) {
  OISAT_CHECK_ARG(result != nullptr, "null pointer");
  if (n_tri <= 0) {
    OISAT_CHECK_CUDA(cudaMemsetAsync(result, 0, 4 * sizeof(uint64_t), (cudaStream_t)stream));
    return OISAT_OK;
  }
  const oisat_flip_item item{tri, half, n_tri, px, py};
  return oisat_flip_delaunay_batch(&item, 1, coord_dtype, workspace, result, stream);
}
