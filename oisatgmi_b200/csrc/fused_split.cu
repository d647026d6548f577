// Two forms of the fused month step for records of fewer than 16 chunks (every BASELINE
// product); oisat_fused_amf (fused_amf.cu, half warp per pair end to end) serves wider records.
//
//   oisat_fused_amf_tile   THE DEFAULT (second half of this file): one launch, a block owns 16
//                          consecutive pairs, gathers their gridded columns into shared memory
//                          and runs the vertical operator there on all of its threads -- no row
//                          buffer, no data-dependent control flow; builds with the BASELINE
//                          products' level counts and stencil size as compile-time constants.
//   oisat_fused_amf_split  the two-launch form it grew out of (below), kept as the fallback for
//                          more than 62 satellite levels and as the bit-for-bit cross-check of
//                          the tile form (tests/test_gpu_fused.py).
//
// Split form of the fused month kernel (round-1 profile: oisat_fused_amf spends
// ~2400 warp instructions per two pairs, two thirds of them overhead of running
// the vertical operator co-operatively on 16 lanes).  Same arithmetic, two launches:
//
//   oisat_gather_rows        half warp per (granule, cell) pair, exactly the gather
//                            stage of fused_amf.cu; the gridded column (SW, p, vcd,
//                            tropopause) is written to a row buffer laid out
//                            [tile of 16 pairs][row][16] so that ...
//   oisat_vertical_amf_rows  ... ONE THREAD per pair evaluates amf_recal.py:93-119
//                            with every load coalesced across the 32 pairs of a
//                            tile (model columns too: consecutive pairs are
//                            consecutive model cells), no shared memory, no
//                            shuffles.  The interp1d bracket search becomes a
//                            merge: model levels and satellite levels are both
//                            sorted in pressure, so the bracket index only moves
//                            one way and each satellite level (and its log) is
//                            visited once.  Profiles that are not strictly
//                            monotone (scipy would argsort them) take a slow
//                            in-thread path with the generic sort + search.
//
// Costs 2 x 8 bytes x rows per pair of extra HBM traffic for the row buffer and
// buys ~4x fewer instructions for the vertical operator.
#include <cmath>
#include <cstdlib>
#include <mutex>

#include "vertical.cuh"

namespace oisat {

__host__ __device__ inline int rec_rows(int L, int has_trop) { return 2 * L + 2 + (has_trop ? 1 : 0); }
__host__ __device__ inline int rec_chunks(int L, int has_trop) { return (rec_rows(L, has_trop) + 7) / 8; }

struct SplitParams {
  oisat_fused_args a;
  double* rows;              // [ceil(n_pairs/16)][nrow_out][16]
  int nrow, nchunk, nrow_out;
  int probe;                 // timing experiments only (OISAT_WS_PROBE): 1 = consumers idle, 2 = producers idle
};

__device__ __forceinline__ void h8_to_f64(const uint4& u, double* z) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    asm("{ .reg .b16 lo, hi;\n\t"
        "  mov.b32 {lo, hi}, %2;\n\t"
        "  cvt.f64.f16 %0, lo;\n\t"
        "  cvt.f64.f16 %1, hi; }"
        : "=d"(z[2 * i]), "=d"(z[2 * i + 1]) : "r"(w[i]));
  }
}

// rows of the buffer: [0, L) scattering weights, [L, 2L) satellite pressures,
// 2L gridded vcd, 2L+1 tropopause (when present)
constexpr int kGatherThreads = 256;

__global__ void __launch_bounds__(kGatherThreads, 4)
gather_rows_kernel(const __grid_constant__ SplitParams P) {
  const oisat_fused_args& A = P.a;
  const int lane = threadIdx.x & 31;
  const int gl = lane & 15;
  const int64_t pair_raw = ((int64_t)blockIdx.x * kGatherThreads + threadIdx.x) >> 4;
  const bool mine = pair_raw < A.n_pairs;
  const int64_t pair = mine ? pair_raw : A.n_pairs - 1;  // shadow work keeps the warp converged
  const int L = A.n_sat_lev;
  const int S = 3 * A.nwin;
  const int g = A.pair_granule[pair];
  const int64_t rec0 = A.gran_record0[g];
  const int64_t px0 = A.gran_px0[g];
  const uint4* records = reinterpret_cast<const uint4*>(A.records);
  extern __shared__ __align__(16) unsigned char gsm[];
  const int sweep = S < 15 ? S : 15;                                  // entries staged at a time
  uint4* stage = reinterpret_cast<uint4*>(gsm);                        // [16][sweep][nchunk] chunks
  double* tile = reinterpret_cast<double*>(gsm + (size_t)16 * sweep * P.nchunk * sizeof(uint4));  // [nrow_out][17]
  double acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.0;
  double acc_amf = 0.0;
  for (int base = 0; base < S; base += 15) {
    const int nk = (S - base) < 15 ? (S - base) : 15;
    uint32_t cix = 0;
    double wt = 0.0, za = 0.0;
    if (gl < nk) {
      const int32_t v = A.vert[pair * S + base + gl];
      wt = A.w[pair * S + base + gl];
      cix = (uint32_t)((rec0 + v) * P.nchunk);
      za = wt * A.amf_masked[px0 + v];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) za += __shfl_xor_sync(0xffffffffu, za, o, 16);
    acc_amf += za;
    // All nk records of the sweep are requested at once with asynchronous copies into
    // this lane's own slots of shared memory (cp.async: no registers held while they
    // are in flight).  The kernel is bound by the latency of these loads -- with
    // register-staged loads, three per lane, 49% of the stall samples were long-
    // scoreboard waits -- so bytes in flight per SM are what counts.
    uint4* slot = stage + ((threadIdx.x >> 4) * sweep) * P.nchunk + gl;   // [pair][entry][chunk]
    for (int e = 0; e < nk; ++e) {
      const uint32_t ck = __shfl_sync(0xffffffffu, cix, e, 16);
      if (gl < P.nchunk) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slot + e * P.nchunk);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(records + ck + gl)
                     : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    for (int e = 0; e < nk; ++e) {
      const double wk = __shfl_sync(0xffffffffu, wt, e, 16);
      if (gl < P.nchunk) {
        double z[8];
        h8_to_f64(slot[e * P.nchunk], z);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fma(wk, z[k], acc[k]);
      }
    }
  }
  // The 16 pairs of this block form one tile of the row buffer.  Rows are
  // transposed through shared memory ([row][pair], pitch 17: conflict-free for a
  // half warp writing consecutive rows) so that the global stores are full 128-byte
  // lines instead of one 8-byte store per row and pair.
  const int col = threadIdx.x >> 4;
  if (gl < P.nchunk) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int row = gl + P.nchunk * e;  // record row
      if (row >= P.nrow) continue;
      if (row == 2 * L + 1) {
        if (mine) A.staged[1 * A.n_pairs + pair] = sqrt(acc[e] * A.box_weight_err);  // interpolator.py:188
      } else {
        const int orow = row <= 2 * L ? row : row - 1;  // tropopause follows vcd in the buffer
        tile[orow * 17 + col] = acc[e] * A.box_weight;
      }
    }
  }
  if (mine && gl == 0) A.staged[4 * A.n_pairs + pair] = acc_amf * A.box_weight;
  __syncthreads();
  double* out = P.rows + (int64_t)blockIdx.x * P.nrow_out * 16;
  for (int i = threadIdx.x; i < P.nrow_out * 16; i += kGatherThreads)
    out[i] = tile[(i >> 4) * 17 + (i & 15)];
}

// ---------------------------------------------------------------------------
// natural logarithm, table-driven
// ---------------------------------------------------------------------------
// The merge below needs log p of every gridded satellite level (amf_recal.py:
// 95-97): L logarithms per pair, 36% of the kernel's instructions with the CUDA
// libm routine (~85 SASS instructions each, special cases included).  This one
// is ~25: with x = 2^e * m and c_i the midpoint of the 1/128-wide interval of m,
//     log x = e ln2 - log(r_i) + log1p(z),   r_i = fl(1 / c_i),  z = m r_i - 1,
// |z| <= 2^-8, log1p by its series up to z^6 (truncation < 2e-18), -log(r_i)
// tabulated for the ROUNDED r_i so the table absorbs its rounding.  Error
// ~2e-16 relative, the same class as libm's and numpy's (which differ from one
// another in the last bit as well); the parity bar for float64 fields is 1e-6.
struct LogTable {
  double r[128];
  double neg_log_r[128];
};
__device__ LogTable g_log_table;

__device__ __forceinline__ double table_log(double x, const LogTable* __restrict__ tab) {
  if (!(x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308)) return log(x);
  const long long bits = __double_as_longlong(x);
  const int e = (int)(bits >> 52) - 1023;
  const int i = (int)(bits >> 45) & 127;
  const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
  const double z = fma(m, tab->r[i], -1.0);
  double p = fma(z, -1.0 / 6.0, 0.2);
  p = fma(z, p, -0.25);
  p = fma(z, p, 1.0 / 3.0);
  p = fma(z, p, -0.5);
  const double l1p = fma(z * z, p, z);
  return fma((double)e, 0.6931471805599453, tab->neg_log_r[i] + l1p);
}

// same, with r_i and -log(r_i) side by side: one 16-byte shared-memory load per logarithm
// (the two separate look-ups are scattered over the table, ~6 wavefronts each)
__device__ __forceinline__ double table_log2(double x, const double2* __restrict__ tab) {
  if (!(x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e308)) return log(x);
  const long long bits = __double_as_longlong(x);
  const int e = (int)(bits >> 52) - 1023;
  const int i = (int)(bits >> 45) & 127;
  const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
  const double2 rn = tab[i];
  const double z = fma(m, rn.x, -1.0);
  double p = fma(z, -1.0 / 6.0, 0.2);
  p = fma(z, p, -0.25);
  p = fma(z, p, 1.0 / 3.0);
  p = fma(z, p, -0.5);
  const double l1p = fma(z * z, p, z);
  return fma((double)e, 0.6931471805599453, rn.y + l1p);
}

// ---------------------------------------------------------------------------
// one thread per pair
// ---------------------------------------------------------------------------
struct RowView {
  double* base;  // &staged rows[tile][0][pair % 16], in shared memory
  __device__ __forceinline__ double at(int row) const { return base[row * 16]; }
  __device__ __forceinline__ void set(int row, double v) const { base[row * 16] = v; }
};

// numpy pairwise sums for n <= 128 (slow path only)
__device__ __forceinline__ double np_sum_f64(const double* v, int n) {
  if (n < 8) { double s = 0.0; for (int i = 0; i < n; ++i) s += v[i]; return s; }
  double q[8];
  for (int j = 0; j < 8; ++j) q[j] = v[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) q[j] += v[i + j];
  double s = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
  for (; i < n; ++i) s += v[i];
  return s;
}
__device__ __forceinline__ float np_sum_f32(const float* v, int n) {
  if (n < 8) { float s = 0.f; for (int i = 0; i < n; ++i) s = __fadd_rn(s, v[i]); return s; }
  float q[8];
  for (int j = 0; j < 8; ++j) q[j] = v[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) q[j] = __fadd_rn(q[j], v[i + j]);
  float s = __fadd_rn(__fadd_rn(__fadd_rn(q[0], q[1]), __fadd_rn(q[2], q[3])),
                      __fadd_rn(__fadd_rn(q[4], q[5]), __fadd_rn(q[6], q[7])));
  for (; i < n; ++i) s = __fadd_rn(s, v[i]);
  return s;
}

// slow path: satellite levels that are not strictly monotone (ties, NaNs, kinks):
// the stable rank sort np.argsort(kind='mergesort') defines, then the generic
// search; single thread, local arrays.  Rare by construction.
__device__ __noinline__ double amf_slow_path(const RowView& r, int L, int n_ctm, bool has_trop,
                                             double trop, const float* lp, const float* pc,
                                             const float* pm, int64_t stride, double* col) {
  double xs[kMaxSatLev], ys[kMaxSatLev];
  for (int i = 0; i < L; ++i) {
    const double xi = r.at(L + i);
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const double xj = r.at(L + j);
      const bool eq = (xj == xi) || (xj != xj && xi != xi);
      rank += (nan_less(xj, xi) || (eq && j < i)) ? 1 : 0;
    }
    xs[rank] = xi;
    ys[rank] = r.at(i);
  }
  double va[kMaxCtmLev];
  float vb[kMaxCtmLev];
  for (int k = 0; k < n_ctm; ++k) {
    double p = (double)pc[(int64_t)k * stride];
    double sw = interp1d_linear<true>(xs, ys, L, (double)lp[(int64_t)k * stride]);
    if (isinf(sw)) sw = 0.0;
    if (has_trop && (double)pm[(int64_t)k * stride] < trop) { sw = qnan(); p = qnan(); }
    const double prod = sw * p;
    va[k] = prod != prod ? 0.0 : prod;
    vb[k] = p != p ? 0.0f : (float)p;
  }
  const double scd = np_sum_f64(va, n_ctm);
  const double vcd_m = (double)np_sum_f32(vb, n_ctm);
  *col = vcd_m;
  return vcd_m != 0.0 ? scd / vcd_m : qnan();
}

constexpr int kQuad = 8;                        // threads per pair (4: 10.6 ms, 8: 10.0, 16: 10.6)
constexpr int kVerticalThreads = 16 * kQuad;    // one 16-pair tile of the row buffer per block

// One block = one tile of the row buffer (16 pairs), FOUR threads per pair.
//
// The rows of the tile are one contiguous ~12 KB piece of the buffer: one bulk
// asynchronous copy brings it into shared memory.  (Before the rows were staged,
// every bracket shift of the walk waited for its own global load -- and a warp
// shifts whenever ANY of its lanes does, ~150 exposed DRAM round trips per warp.)
//
// The walk over the model column is serial in nature (a merge), and with one thread
// per pair the kernel ran 8 warps per SM waiting on float64 dependency chains (ncu:
// issue slots 31% busy, 42% of the stalls fixed-latency waits).  So the column is cut
// into kQuad contiguous pieces: each thread finds its starting bracket with one
// binary search (the same searchsorted the merge tracks incrementally) and walks its
// piece; the per-level products go to shared memory and ONE thread per pair adds them
// in numpy's order (eight running sums, then the tree), so the result does not
// depend on how the column was cut.  Thread t of pair p is threadIdx.x = 16 t + p: a
// half warp reads 16 different columns of the tile, each in its own bank pair.
template <bool HAS_TROP>
__global__ void __launch_bounds__(kVerticalThreads)
vertical_rows_kernel(const __grid_constant__ SplitParams P) {
  extern __shared__ __align__(128) unsigned char vsm[];
  __shared__ LogTable tab;
  __shared__ __align__(8) unsigned long long mbar;
  __shared__ int unsorted[16];
  __shared__ double part_a[8 * 16];
  __shared__ float part_b[8 * 16];
  const oisat_fused_args& A = P.a;
  const int L = A.n_sat_lev, n_ctm = A.n_ctm_lev;
  double* rows_s = reinterpret_cast<double*>(vsm);                 // [nrow_out][16]
  double* prod_a = rows_s + (size_t)P.nrow_out * 16;               // [n_ctm][16]
  float* prod_b = reinterpret_cast<float*>(prod_a + (size_t)n_ctm * 16);   // [n_ctm][16]
  const int p = threadIdx.x & 15, t = threadIdx.x >> 4;
  const int64_t tile = blockIdx.x;
  const uint32_t tile_bytes = (uint32_t)P.nrow_out * 16u * (uint32_t)sizeof(double);
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(bar), "r"(tile_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"((uint32_t)__cvta_generic_to_shared(rows_s)),
          "l"(P.rows + tile * (int64_t)P.nrow_out * 16), "r"(tile_bytes), "r"(bar)
        : "memory");
  }
  for (int i = threadIdx.x; i < 128; i += kVerticalThreads) {
    tab.r[i] = g_log_table.r[i];
    tab.neg_log_r[i] = g_log_table.neg_log_r[i];
  }
  if (threadIdx.x < 16) unsorted[threadIdx.x] = n_ctm >= 8 ? 0 : 1;
  __syncthreads();  // barrier initialised, table staged
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p;\n\t"
                   "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                   "  selp.b32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar) : "memory");
  }
  const int64_t pair = tile * 16 + p;
  const bool live = pair < A.n_pairs;
  RowView r{rows_s + p};
  const double vcd = r.at(2 * L);
  const bool work = live && vcd == vcd;  // amf_recal.py:99-100
  const double trop = HAS_TROP ? r.at(2 * L + 1) : 0.0;
  // scipy sorts the levels ascending in log p: xs[j], j = 0..L-1; for a monotone
  // profile that is the storage order or its reverse.
  const bool descending = r.at(L) > r.at(2 * L - 1);
  auto row_of = [&](int j) { return descending ? L - 1 - j : j; };
  auto xs = [&](int j) { return r.at(L + row_of(j)); };
  __syncthreads();  // every thread has read the pressures it needs before they turn into logs
  // Phase A: p -> log p in place, level j by thread j mod kQuad
  if (work)
    for (int j = t; j < L; j += kQuad) {
      const int row = L + row_of(j);
      r.set(row, table_log(r.at(row), &tab));
    }
  __syncthreads();
  // strict monotonicity of every level; ties and NaNs take scipy's argsort path
  if (work) {
    bool bad = false;
    for (int j = t; j < L; j += kQuad) {
      const double x = xs(j);
      const double prev = j > 0 ? xs(j - 1) : -CUDART_INF;
      bad = bad || !(prev < x);
    }
    if (bad) unsorted[p] = 1;
  }
  __syncthreads();
  const bool sorted = unsorted[p] == 0;
  // model column of this pair's cell: consecutive pairs are consecutive cells, so a
  // half warp reads (mostly) one 64-byte piece of a line per level
  const float* lp = A.ctm_logp;
  const float* pc = A.ctm_pcol;
  const float* pm = A.ctm_pmid;
  const int64_t stride = A.n_cell;
  if (work) {
    const int64_t off = (int64_t)A.gran_slot[A.pair_granule[pair]] * n_ctm * A.n_cell +
                        A.pair_cell[pair];
    lp += off;
    pc += off;
    pm = HAS_TROP ? pm + off : lp;
  }
  // Phase B: interp1d(xs, SW)(log p_model) over this thread's piece of the column.
  // Model levels run from the surface up, so the query only decreases and so does
  // idx = searchsorted(xs, v, 'left'); the bracket [c-1, c], c = clip(idx, 1, L-1),
  // lives in registers and only ever moves one way.
  const int per = (n_ctm + kQuad - 1) / kQuad;
  const int k_begin = t * per, k_end = (t + 1) * per < n_ctm ? (t + 1) * per : n_ctm;
  if (work && sorted && k_begin < k_end) {
    int idx = 0;
    {
      const double v0 = (double)__ldg(lp + (int64_t)k_begin * stride);
      for (int step = 1 << (31 - __clz(L)); step >= 1; step >>= 1) {
        const int tt = idx + step;
        const int probe = (tt <= L ? tt : L) - 1;
        idx = (tt <= L && xs(probe) < v0) ? tt : idx;
      }
    }
    int c = idx < 1 ? 1 : (idx > L - 1 ? L - 1 : idx);
    double x_hi = xs(c), x_lo = xs(c - 1);
    double y_hi = r.at(row_of(c)), y_lo = r.at(row_of(c - 1));
    double rden = __drcp_rn(x_hi - x_lo);
    constexpr int kBatch = 6;
    for (int k0 = k_begin; k0 < k_end; k0 += kBatch) {
      float lpv[kBatch], pcv8[kBatch], pmv[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {  // several levels of the model column in flight at once
        const int k = k0 + j < k_end ? k0 + j : k_end - 1;
        lpv[j] = __ldg(lp + (int64_t)k * stride);
        pcv8[j] = __ldg(pc + (int64_t)k * stride);
        pmv[j] = HAS_TROP ? __ldg(pm + (int64_t)k * stride) : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int k = k0 + j;
        if (k < k_end) {
          const double v = (double)lpv[j];
          double pcv = (double)pcv8[j];
          while (idx > 0) {
            const double xm = (idx - 1 == c) ? x_hi : x_lo;
            if (xm < v) break;
            --idx;
            const int nc = idx < 1 ? 1 : (idx > L - 1 ? L - 1 : idx);
            if (nc != c) {  // shift the bracket one level toward lower pressure
              c = nc;
              x_hi = x_lo;
              y_hi = y_lo;
              x_lo = xs(c - 1);
              y_lo = r.at(row_of(c - 1));
              rden = __drcp_rn(x_hi - x_lo);   // correctly rounded, = 1.0 / x bit for bit
            }
          }
          double sw = ((v - x_lo) * rden) * y_hi + ((x_hi - v) * rden) * y_lo;  // interp1d._call_linear
          if (isinf(sw)) sw = 0.0;
          if (HAS_TROP && (double)pmv[j] < trop) { sw = qnan(); pcv = qnan(); }
          const double prod = sw * pcv;
          prod_a[k * 16 + p] = prod != prod ? 0.0 : prod;        // nansum terms
          prod_b[k * 16 + p] = pcv != pcv ? 0.0f : (float)pcv;
        }
      }
    }
  }
  __syncthreads();
  // Phase C: the terms are added in numpy's order -- eight running sums over k mod 8
  // (thread t owns sum t), then the fixed tree and the scalar tail by one thread
  static_assert(kQuad == 8, "phase C maps numpy's eight running sums onto the eight threads");
  const int body = n_ctm - (n_ctm % 8);
  if (work && sorted) {
    double qa = prod_a[t * 16 + p];
    float qb = prod_b[t * 16 + p];
    for (int k = t + 8; k < body; k += 8) {
      qa = qa + prod_a[k * 16 + p];
      qb = __fadd_rn(qb, prod_b[k * 16 + p]);
    }
    part_a[t * 16 + p] = qa;
    part_b[t * 16 + p] = qb;
  }
  __syncthreads();
  if (t != 0 || !live) return;
  const double old_amf = A.staged[4 * A.n_pairs + pair];
  double new_amf = qnan(), vnew = qnan(), col = qnan();
  if (work) {
    double colsum = 0.0;
    if (sorted) {
      auto qa = [&](int j) { return part_a[j * 16 + p]; };
      auto qb = [&](int j) { return part_b[j * 16 + p]; };
      double scd = ((qa(0) + qa(1)) + (qa(2) + qa(3))) + ((qa(4) + qa(5)) + (qa(6) + qa(7)));
      float cs = __fadd_rn(__fadd_rn(__fadd_rn(qb(0), qb(1)), __fadd_rn(qb(2), qb(3))),
                           __fadd_rn(__fadd_rn(qb(4), qb(5)), __fadd_rn(qb(6), qb(7))));
      for (int k = body; k < n_ctm; ++k) {   // scalar tail after the tree
        scd = scd + prod_a[k * 16 + p];
        cs = __fadd_rn(cs, prod_b[k * 16 + p]);
      }
      colsum = (double)cs;
      new_amf = colsum != 0.0 ? scd / colsum : qnan();
    } else {
      new_amf = amf_slow_path(r, L, n_ctm, HAS_TROP, trop, lp, pc, pm, stride, &colsum);
    }
    vnew = (old_amf * vcd) / new_amf;                       // amf_recal.py:179
    col = (vnew != vnew || isinf(vnew)) ? qnan() : colsum;  // :180-181
  }
  A.staged[0 * A.n_pairs + pair] = vnew;
  A.staged[2 * A.n_pairs + pair] = col;
  A.staged[3 * A.n_pairs + pair] = new_amf;
}


// ---------------------------------------------------------------------------
// one kernel per 16-pair tile: gather, then the vertical operator from shared memory
// ---------------------------------------------------------------------------
// The two kernels above exchange the gridded columns through a row buffer in HBM
// (2 x 8 bytes x rows per pair: 11 GB of the 26 GB the pair of them moves for an OMI
// HCHO month).  Here the 16 pairs of a block never leave the SM: the gather phase is the
// one of gather_rows_kernel (same arithmetic, bit for bit), its tile stays in shared
// memory, and the vertical operator runs on all 256 threads of the block with no
// data-dependent control flow:
//   * phase A: log p of every gridded level by table (in place, and a copy in
//     ascending order padded with +inf to 63 rows), then the reciprocal width of every
//     bracket once per pair (the merge of vertical_rows_kernel recomputed it whenever
//     any lane of the warp shifted its bracket -- 13% of its instructions, 20% of its
//     stall samples);
//   * phase B: thread (t, p) evaluates interp1d at model levels k = (t mod 8) + 8 i of
//     pair p -- searchsorted as a six-step bisection over the padded copy (no bounds
//     test), the bracket from shared memory -- and adds the terms in registers in
//     numpy's order: level k belongs to running sum k mod 8, which threads t and t + 8
//     share (first and second half of i; the second half continues the first half's
//     partial sum after a barrier), so no per-level products are staged at all;
//   * phase C: the fixed tree over the eight running sums and the scalar tail.
// Profiles that are not strictly monotone take amf_slow_path, as before.
constexpr int kTileThreads = 256;
constexpr int kTP = 17;           // pitch (doubles) of the gathered tile: 16 pairs + 1 (the gather writes rows per lane)
constexpr int kXP = 16;           // pitch of the search arrays: pair p owns bank pair p, so ANY mix of rows
                                  // across the 16 pairs of a half warp is conflict-free (phase B's probes are
                                  // data dependent; with pitch 17 they cost ~3.5 wavefronts each instead of 2)
constexpr int kSearchRows = 63;   // six bisection steps reach row 62

template <int PITCH>
struct RowViewP {
  double* base;
  __device__ __forceinline__ double at(int row) const { return base[row * PITCH]; }
  __device__ __forceinline__ void set(int row, double v) const { base[row * PITCH] = v; }
};

// amf_slow_path for a tile of pitch kTP (same code, different view)
__device__ __noinline__ double amf_slow_path_tile(const RowViewP<kTP>& r, int L, int n_ctm,
                                                  bool has_trop, double trop, const float* lp,
                                                  const float* pc, const float* pm, int64_t stride,
                                                  double* col) {
  double xs[kMaxSatLev], ys[kMaxSatLev];
  for (int i = 0; i < L; ++i) {
    const double xi = r.at(L + i);
    int rank = 0;
    for (int j = 0; j < L; ++j) {
      const double xj = r.at(L + j);
      const bool eq = (xj == xi) || (xj != xj && xi != xi);
      rank += (nan_less(xj, xi) || (eq && j < i)) ? 1 : 0;
    }
    xs[rank] = xi;
    ys[rank] = r.at(i);
  }
  double va[kMaxCtmLev];
  float vb[kMaxCtmLev];
  for (int k = 0; k < n_ctm; ++k) {
    double p = (double)pc[(int64_t)k * stride];
    double sw = interp1d_linear<true>(xs, ys, L, (double)lp[(int64_t)k * stride]);
    if (isinf(sw)) sw = 0.0;
    if (has_trop && (double)pm[(int64_t)k * stride] < trop) { sw = qnan(); p = qnan(); }
    const double prod = sw * p;
    va[k] = prod != prod ? 0.0 : prod;
    vb[k] = p != p ? 0.0f : (float)p;
  }
  const double scd = np_sum_f64(va, n_ctm);
  const double vcd_m = (double)np_sum_f32(vb, n_ctm);
  *col = vcd_m;
  return vcd_m != 0.0 ? scd / vcd_m : qnan();
}

struct TileSmem {   // static part
  double2 tab[128];   // (r_i, -log r_i)
  union {
    struct {          // vertical phase
      double part_a[8 * 16];
      double tail_a[8 * 16];
      float part_b[8 * 16];
      float tail_b[8 * 16];
    };
    struct {          // packed gather: weight and first chunk of every (pair, entry) of a sweep
      double gw[16 * 15];
      uint32_t gcix[16 * 15];
    };
  };
  double old_amf[16];
  int unsorted[16];
  int pair_id[16];    // pair of tile slot p (alive_pairs[16 * tile + p])
};

// H = levels per thread and half (ceil(ceil(n_ctm / 8) / 2) <= H).  CL / CS / CN: the
// satellite level count, stencil size and model level count when known at compile time
// (the BASELINE products), 0 = read from the arguments: with constants the record and
// tile geometry fold into immediates and the loops unroll (the generic build spends more
// than half of its instructions outside the arithmetic).
template <bool HAS_TROP, int H, int CL, int CS, int CN, int SW, int MINB, bool PACKED, bool BULK = false>
__global__ void __launch_bounds__(kTileThreads, MINB)
fused_tile_kernel(const __grid_constant__ SplitParams P) {
  const oisat_fused_args& A = P.a;
  extern __shared__ __align__(16) unsigned char tsm[];
  __shared__ TileSmem sm;
  __shared__ __align__(8) unsigned long long gather_bar;   // BULK: completion of a sweep's record copies
  const int lane = threadIdx.x & 31;
  const int gl = lane & 15;
  const int col = threadIdx.x >> 4;                                    // pair of the tile (gather phase)
  const int L = CL > 0 ? CL : A.n_sat_lev;
  const int n_ctm = CN > 0 ? CN : A.n_ctm_lev;
  const int S = CS > 0 ? CS : 3 * A.nwin;
  const int nrow = CL > 0 ? rec_rows(CL, HAS_TROP) : P.nrow;
  const int nchunk = CL > 0 ? rec_chunks(CL, HAS_TROP) : P.nchunk;
  const int nrow_out = CL > 0 ? 2 * CL + 1 + (HAS_TROP ? 1 : 0) : P.nrow_out;
  const int sweep = S < SW ? S : SW;   // entries staged at a time (SW <= 15: one lane per entry)
  // dynamic shared memory: [tile | union(gather stage, {xs, rd})]
  double* tile = reinterpret_cast<double*>(tsm);                       // [nrow_out][kTP]
  unsigned char* uni = tsm + (((size_t)nrow_out * kTP * sizeof(double) + 15) & ~(size_t)15);
  uint4* stage = reinterpret_cast<uint4*>(uni);                        // [16][sweep][nchunk]
  double* xs_s = reinterpret_cast<double*>(uni);                       // [kSearchRows][kXP], ascending
  double* rd_s = xs_s + kSearchRows * kXP;                             // [L][kXP], 1 / (xs[c] - xs[c-1])
  for (int i = threadIdx.x; i < 128; i += kTileThreads) {
    sm.tab[i] = make_double2(g_log_table.r[i], g_log_table.neg_log_r[i]);
  }
  if (threadIdx.x < 16) sm.unsorted[threadIdx.x] = n_ctm >= 8 ? 0 : 1;
  if (BULK && threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(&gather_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // identity of this thread in the vertical phase (pair p = gather lane, t = gather pair):
  // known now, so the model column's offset is loaded early and the column itself is
  // prefetched into L2 while the records are in flight
  const int p = threadIdx.x & 15, t = threadIdx.x >> 4;
  // the tile's pairs come from the compact list of pairs without a masked pixel
  // (oisat_pair_alive); blocks beyond the list have nothing to do
  const int64_t n_act = *A.n_alive;
  if ((int64_t)blockIdx.x * 16 >= n_act) return;
  const int64_t vslot = (int64_t)blockIdx.x * 16 + p;
  const bool live = vslot < n_act;
  const int64_t vpair = A.alive_pairs[live ? vslot : n_act - 1];
  const uint32_t stride = (uint32_t)A.n_cell;
  uint32_t off = 0;
  if (live)
    off = A.pair_ctm_off ? A.pair_ctm_off[vpair]
                         : (uint32_t)A.gran_slot[A.pair_granule[vpair]] * (uint32_t)n_ctm * stride +
                               (uint32_t)A.pair_cell[vpair];
  const float* lp = A.ctm_logp;
  const float* pc = A.ctm_pcol;
  const float* pm = HAS_TROP ? A.ctm_pmid : A.ctm_logp;
  // levels of phase B: k = j8 + 8 (i0 + i), i < cnt
  const int body = n_ctm - (n_ctm % 8);
  const int nb = body >> 3;                 // terms per running sum
  const int half = (nb + 1) >> 1;           // <= H
  const int j8 = t & 7;
  const int i0 = t < 8 ? 0 : half;
  const int cnt = t < 8 ? half : nb - half;
  // ------------------------------------------------------------ gather phase
  if constexpr (PACKED) {
    // Packed lanes: a record of nchunk < 16 chunks leaves 16 - nchunk lanes of every half warp
    // idle in the loops below (4 of 16 for OMI HCHO, 6 for OMI NO2, 7 for TROPOMI).  Here the
    // 16 * nchunk (pair, chunk) work items of the tile are dealt to consecutive threads, so
    // the loops run on 6 / 5 / 4.5 warps instead of 8; weights and chunk indices of a sweep
    // come from a small shared table (filled in the half-warp layout, which also keeps the
    // AMF term's summation tree) instead of shuffles.  Same arithmetic, same order.
    const int64_t pair_raw = (int64_t)blockIdx.x * 16 + col;
    const bool mine = pair_raw < n_act;
    const int64_t pair = A.alive_pairs[mine ? pair_raw : n_act - 1];
    if (gl == 0) {
      sm.pair_id[col] = (int)pair;
      sm.old_amf[col] = A.staged[4 * A.n_pairs + pair];   // gridded by oisat_pair_alive
    }
    int64_t rec0;
    if (A.pair_record0) {
      rec0 = A.pair_record0[pair];
    } else {
      rec0 = A.gran_record0[A.pair_granule[pair]];
    }
    const uint4* records = reinterpret_cast<const uint4*>(A.records);
    const int tid = threadIdx.x;
    const bool active = tid < 16 * nchunk;
    const int pp = active ? tid / nchunk : 0;          // pair of the tile
    const int ch = active ? tid - pp * nchunk : 0;     // chunk of its records
    double acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0;
    uint4* slot = stage + (pp * sweep) * nchunk + ch;                  // [pair][entry][chunk]
#pragma unroll 1
    for (int base = 0; base < S; base += SW) {
      const int nk = (S - base) < SW ? (S - base) : SW;
      int32_t v = 0;
      double wt = 0.0;
      if (base > 0) __syncthreads();                   // the previous sweep is done with the table
      if (gl < nk) {
        v = A.vert[pair * S + base + gl];
        wt = A.w[pair * S + base + gl];
        sm.gcix[col * SW + gl] = (uint32_t)((rec0 + v) * nchunk);
        sm.gw[col * SW + gl] = wt;
      }
      __syncthreads();
      if constexpr (BULK) {
        // One bulk asynchronous copy (TMA engine, cp.async.bulk) per (pair, entry): a record is
        // nchunk * 16 contiguous bytes on both sides, so the 16 x nk records of the sweep are
        // 16 x nk instructions in the whole block instead of 16 x nk x nchunk lane copies with
        // their address arithmetic; completion is counted on an mbarrier.
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&gather_bar);
        const uint32_t rec_bytes = (uint32_t)nchunk * 16u;
        if (tid == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                       ::"r"(bar), "r"(16u * (uint32_t)nk * rec_bytes) : "memory");
        if (tid < 16 * SW) {
          const int p2 = tid / SW, e2 = tid - p2 * SW;
          if (e2 < nk) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage + (p2 * sweep + e2) * nchunk);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(dst), "l"(records + sm.gcix[p2 * SW + e2]), "r"(rec_bytes), "r"(bar) : "memory");
          }
        }
      } else {
        if (active) {
#pragma unroll
          for (int e = 0; e < SW; ++e) {
            if (e < nk) {
              const uint32_t ck = sm.gcix[pp * SW + e] + (uint32_t)ch;
              const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slot + e * nchunk);
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(records + ck)
                           : "memory");
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      if (base == 0 && live) {
#pragma unroll
        for (int i = 0; i < H; ++i) {
          if (i < cnt) {
            const uint32_t at = off + (uint32_t)(j8 + 8 * (i0 + i)) * stride;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(lp + at));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + at));
            if (HAS_TROP) asm volatile("prefetch.global.L2 [%0];" ::"l"(pm + at));
          }
        }
      }
      if constexpr (BULK) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&gather_bar);
        const uint32_t parity = (uint32_t)((base / SW) & 1);
        uint32_t done = 0;
        while (!done)
          asm volatile("{ .reg .pred p;\n\t"
                       "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                       "  selp.b32 %0, 1, 0, p; }"
                       : "=r"(done) : "r"(bar), "r"(parity) : "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      if (active) {
#pragma unroll
        for (int e = 0; e < SW; ++e) {
          if (e < nk) {
            const double wk = sm.gw[pp * SW + e];
            double z[8];
            h8_to_f64(slot[e * nchunk], z);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fma(wk, z[k], acc[k]);
          }
        }
      }
    }
    const int sig_row = 2 * L + 1;
    if (active) {
      double sig = 0.0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int row = ch + nchunk * e;
        if (row == sig_row) sig = acc[e];
        const int orow = row <= 2 * L ? row : row - 1;                 // tropopause follows vcd
        if (row < nrow && row != sig_row) tile[orow * kTP + pp] = acc[e] * A.box_weight;
      }
      if ((int64_t)blockIdx.x * 16 + pp < n_act && ch == sig_row % nchunk)
        A.staged[1 * A.n_pairs + sm.pair_id[pp]] = sqrt(sig * A.box_weight_err);   // interpolator.py:188
    }
  } else
  {
    const int64_t pair_raw = (int64_t)blockIdx.x * 16 + col;
    const bool mine = pair_raw < n_act;
    const int64_t pair = A.alive_pairs[mine ? pair_raw : n_act - 1];  // shadow work keeps the warp converged
    if (gl == 0) sm.old_amf[col] = A.staged[4 * A.n_pairs + pair];    // gridded by oisat_pair_alive
    int64_t rec0;
    if (A.pair_record0) {
      rec0 = A.pair_record0[pair];
    } else {
      rec0 = A.gran_record0[A.pair_granule[pair]];
    }
    const uint4* records = reinterpret_cast<const uint4*>(A.records);
    double acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0;
    uint4* slot = stage + (col * sweep) * nchunk + gl;                 // [pair][entry][chunk]
    const bool has_chunk = gl < nchunk;
#pragma unroll 1
    for (int base = 0; base < S; base += SW) {
      const int nk = (S - base) < SW ? (S - base) : SW;
      uint32_t cix = 0;
      int32_t v = 0;
      double wt = 0.0;
      if (gl < nk) {
        v = A.vert[pair * S + base + gl];
        wt = A.w[pair * S + base + gl];
        cix = (uint32_t)((rec0 + v) * nchunk);
      }
      // All nk records of the sweep are requested at once with asynchronous copies into this
      // lane's own slots (no registers held while they are in flight) ...
#pragma unroll
      for (int e = 0; e < SW; ++e) {
        if (e < nk) {
          const uint32_t ck = __shfl_sync(0xffffffffu, cix, e, 16) + (uint32_t)gl;   // 32-bit chunk index
          if (has_chunk) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slot + e * nchunk);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(records + ck)
                         : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      // ... and while they are, once, the prefetch of the model levels this thread evaluates
      // in phase B
      if (base == 0 && live) {
#pragma unroll
        for (int i = 0; i < H; ++i) {
          if (i < cnt) {
            const uint32_t at = off + (uint32_t)(j8 + 8 * (i0 + i)) * stride;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(lp + at));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + at));
            if (HAS_TROP) asm volatile("prefetch.global.L2 [%0];" ::"l"(pm + at));
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int e = 0; e < SW; ++e) {
        if (e < nk) {
          const double wk = __shfl_sync(0xffffffffu, wt, e, 16);
          if (has_chunk) {
            double z[8];
            h8_to_f64(slot[e * nchunk], z);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fma(wk, z[k], acc[k]);
          }
        }
      }
    }
    // rows of this lane: gl + nchunk * e; the variance row (2L+1) leaves as sigma
    const int sig_row = 2 * L + 1;
    double sig = 0.0;
    if (has_chunk) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int row = gl + nchunk * e;
        if (row == sig_row) sig = acc[e];
        const int orow = row <= 2 * L ? row : row - 1;                 // tropopause follows vcd
        if (row < nrow && row != sig_row) tile[orow * kTP + col] = acc[e] * A.box_weight;
      }
    }
    if (mine && gl == sig_row % nchunk)
      A.staged[1 * A.n_pairs + pair] = sqrt(sig * A.box_weight_err);   // interpolator.py:188
  }
  __syncthreads();   // tile complete; the stage area is free
  // ---------------------------------------------------------- vertical phase
  const int64_t pair = vpair;
  RowViewP<kTP> r{tile + p};
  const double vcd = r.at(2 * L);
  const bool work = live && vcd == vcd;  // amf_recal.py:99-100
  const double trop = HAS_TROP ? r.at(2 * L + 1) : 0.0;
  const bool descending = r.at(L) > r.at(2 * L - 1);
  __syncthreads();   // every thread has read the raw pressures it needs before they turn into logs
  // phase A: p -> log p in place + ascending copy; padding rows = +inf
#pragma unroll
  for (int row = t; row < kSearchRows; row += 16) {
    if (row < L) {
      if (work) {
        const double lg = table_log2(r.at(L + row), sm.tab);
        r.set(L + row, lg);
        xs_s[(descending ? L - 1 - row : row) * kXP + p] = lg;
      }
    } else {
      xs_s[row * kXP + p] = CUDART_INF;
    }
  }
  __syncthreads();
  // this thread's piece of the model column (prefetched into L2 during the gather) is
  // requested now, all of it at once, and arrives while the bracket widths are inverted
  float lpv[H], pcv8[H], pmv[H];
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const int ii = i < cnt ? i : (cnt > 0 ? cnt - 1 : 0);
    const uint32_t at = off + (uint32_t)(j8 + 8 * (i0 + ii)) * stride;
    const bool ok = work && cnt > 0;
    lpv[i] = ok ? __ldg(lp + at) : 0.0f;
    pcv8[i] = ok ? __ldg(pc + at) : 0.0f;
    pmv[i] = (HAS_TROP && ok) ? __ldg(pm + at) : 0.0f;
  }
  if (work) {
    bool bad = false;
#pragma unroll
    for (int j = t; j < L; j += 16) {
      const double x = xs_s[j * kXP + p];
      const double prev = j > 0 ? xs_s[(j - 1) * kXP + p] : -CUDART_INF;
      bad = bad || !(prev < x);
      if (j > 0) rd_s[j * kXP + p] = __drcp_rn(x - prev);              // = 1.0 / (x - prev) bit for bit
    }
    if (bad) sm.unsorted[p] = 1;
  }
  __syncthreads();
  const bool sorted = sm.unsorted[p] == 0;
  // phase B
  const bool go = work && sorted;
  // signed row stride of the scattering weights in ascending-pressure order
  const double* y0 = tile + (descending ? L - 1 : 0) * kTP + p;
  const int ystep = descending ? -kTP : kTP;
  auto term = [&](float lpf, float pcf, float pmf, double& ta, float& tb) {
    const double v = (double)lpf;
    double pcv = (double)pcf;
    int idx = 0;
#pragma unroll
    for (int step = 32; step >= 1; step >>= 1)
      idx += (xs_s[(idx + step - 1) * kXP + p] < v) ? step : 0;         // searchsorted(xs, v, 'left')
    const int c = idx < 1 ? 1 : (idx > L - 1 ? L - 1 : idx);
    const double x_hi = xs_s[c * kXP + p], x_lo = xs_s[(c - 1) * kXP + p];
    const double rden = rd_s[c * kXP + p];
    const double y_hi = y0[c * ystep], y_lo = y0[(c - 1) * ystep];
    double sw = ((v - x_lo) * rden) * y_hi + ((x_hi - v) * rden) * y_lo;  // interp1d._call_linear
    if (isinf(sw)) sw = 0.0;
    if (HAS_TROP && (double)pmf < trop) { sw = qnan(); pcv = qnan(); }
    const double prod = sw * pcv;
    ta = prod != prod ? 0.0 : prod;        // nansum terms
    tb = pcv != pcv ? 0.0f : (float)pcv;
  };
  double ta[H];
  float tb[H];
  if (go) {
#pragma unroll
    for (int i = 0; i < H; ++i) {
      ta[i] = 0.0;
      tb[i] = 0.0f;
      if (i < cnt) term(lpv[i], pcv8[i], pmv[i], ta[i], tb[i]);
    }
    // scalar tail of numpy's pairwise sum: levels body .. n_ctm-1, one per thread t < 8
    if (t < 8 && body + t < n_ctm) {
      const uint32_t at = off + (uint32_t)(body + t) * stride;
      double a;
      float b;
      term(__ldg(lp + at), __ldg(pc + at), HAS_TROP ? __ldg(pm + at) : 0.0f, a, b);
      sm.tail_a[t * 16 + p] = a;
      sm.tail_b[t * 16 + p] = b;
    }
    if (t < 8) {                            // running sum j8, first half: r = a[j]; r += a[j + 8 i]
      double qa = ta[0];
      float qb = tb[0];
#pragma unroll
      for (int i = 1; i < H; ++i)
        if (i < cnt) { qa = qa + ta[i]; qb = __fadd_rn(qb, tb[i]); }
      sm.part_a[j8 * 16 + p] = qa;
      sm.part_b[j8 * 16 + p] = qb;
    }
  }
  __syncthreads();
  if (go && t >= 8) {                       // second half continues the same running sum
    double qa = sm.part_a[j8 * 16 + p];
    float qb = sm.part_b[j8 * 16 + p];
#pragma unroll
    for (int i = 0; i < H; ++i)
      if (i < cnt) { qa = qa + ta[i]; qb = __fadd_rn(qb, tb[i]); }
    sm.part_a[j8 * 16 + p] = qa;
    sm.part_b[j8 * 16 + p] = qb;
  }
  __syncthreads();
  // phase C
  if (t != 0 || !live) return;
  const double old_amf = sm.old_amf[p];
  double new_amf = qnan(), vnew = qnan(), colv = qnan();
  if (work) {
    double colsum = 0.0;
    if (sorted) {
      auto qa = [&](int j) { return sm.part_a[j * 16 + p]; };
      auto qb = [&](int j) { return sm.part_b[j * 16 + p]; };
      double scd = ((qa(0) + qa(1)) + (qa(2) + qa(3))) + ((qa(4) + qa(5)) + (qa(6) + qa(7)));
      float cs = __fadd_rn(__fadd_rn(__fadd_rn(qb(0), qb(1)), __fadd_rn(qb(2), qb(3))),
                           __fadd_rn(__fadd_rn(qb(4), qb(5)), __fadd_rn(qb(6), qb(7))));
      for (int k = body; k < n_ctm; ++k) {   // scalar tail after the tree
        scd = scd + sm.tail_a[(k - body) * 16 + p];
        cs = __fadd_rn(cs, sm.tail_b[(k - body) * 16 + p]);
      }
      colsum = (double)cs;
      new_amf = colsum != 0.0 ? scd / colsum : qnan();
    } else {
      new_amf = amf_slow_path_tile(r, L, n_ctm, HAS_TROP, trop, lp + off, pc + off, pm + off,
                                     (int64_t)stride, &colsum);
    }
    vnew = (old_amf * vcd) / new_amf;                        // amf_recal.py:179
    colv = (vnew != vnew || isinf(vnew)) ? qnan() : colsum;  // :180-181
  }
  A.staged[0 * A.n_pairs + pair] = vnew;
  A.staged[2 * A.n_pairs + pair] = colv;
  A.staged[3 * A.n_pairs + pair] = new_amf;
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised, persistent form of the tile kernel (OISAT_TILE_WS=1; an experiment, see
// oisat_fused_amf_tile for what it measured)
// ---------------------------------------------------------------------------------------------
// The tile kernel above is latency bound (ncu, profiles/r02_*): a block starts with three
// DEPENDENT global loads (live-pair list -> stencil -> record index) before it can request a
// single record (18 % of its stall samples), then waits for the records, then all of its warps
// stand at barriers between the gather and the phases of the vertical operator (15 %).  Here a
// block lives for the whole launch (two per SM) and its warps have ONE job each:
//
//   7 producer warps   gather: stencil tables of the NEXT sweep / tile are requested while the
//                      records of this one are in flight (the dependent chain is always one
//                      step ahead, in registers, then in a double-buffered shared table), the
//                      model columns of the next tile are prefetched into L2, records arrive by
//                      cp.async into the producers' own stage, float16 -> float64, FMA in scipy's
//                      order; the gridded tile goes to one of TWO tile buffers;
//   8 consumer warps   the vertical operator of the previous tile on the other buffer, exactly
//                      the phases A / B / C of the tile kernel (same code, same order, same
//                      bits: tests/test_gpu_fused.py compares the two forms bit for bit).
//
// Hand-over with named barriers (bar.arrive / bar.sync, ids 2..5 = full / empty per buffer);
// producers and consumers each have a private barrier for their internal steps.  While the
// consumers run logarithms and bisections the producers keep ~80 KB of record requests per SM
// in flight: the memory pipe and the FP64 / issue slots are busy at the same time.
constexpr int kWsProd = 224;                 // 7 producer warps: 16 pairs x (<= 14) chunks
constexpr int kWsCons = 256;                 // 8 consumer warps: 16 level lanes x 16 pairs
constexpr int kWsThreads = kWsProd + kWsCons;
constexpr int kWsSW = 15;                    // stencil entries per sweep
enum { kBarProd = 1, kBarFull = 2, kBarEmpty = 4, kBarCons = 6 };

__device__ __forceinline__ void bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int n) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}

struct WsSmem {
  double2 tab[128];                 // (r_i, -log r_i)
  double part_a[8 * 16];
  double tail_a[8 * 16];
  float part_b[8 * 16];
  float tail_b[8 * 16];
  double gw[2][16 * kWsSW];         // weight of every (pair slot, entry) of a sweep, double-buffered
  uint32_t gcix[2][16 * kWsSW];     // first chunk of its record
  int pid[3][16];                   // pair ids of tiles it, it+1, it+2 (ring)
  double c_old_amf[2][16];          // per tile buffer: what the consumers need to know
  uint32_t c_off[2][16];
  int c_pair[2][16];
  int unsorted[16];
};

template <bool HAS_TROP, int H, int CL, int CS, int CN>
__global__ void __launch_bounds__(kWsThreads, 2)
fused_ws_kernel(const __grid_constant__ SplitParams P) {
  const oisat_fused_args& A = P.a;
  extern __shared__ __align__(16) unsigned char wsm[];
  __shared__ WsSmem sm;
  const int L = CL > 0 ? CL : A.n_sat_lev;
  const int n_ctm = CN > 0 ? CN : A.n_ctm_lev;
  const int S = CS > 0 ? CS : 3 * A.nwin;
  const int nrow = CL > 0 ? rec_rows(CL, HAS_TROP) : P.nrow;
  const int nchunk = CL > 0 ? rec_chunks(CL, HAS_TROP) : P.nchunk;
  const int nrow_out = CL > 0 ? 2 * CL + 1 + (HAS_TROP ? 1 : 0) : P.nrow_out;
  const int sweep = S < kWsSW ? S : kWsSW;
  // dynamic shared memory: [tile 0 | tile 1 | stage | xs | rd]
  const size_t tile_doubles = ((size_t)nrow_out * kTP + 1) & ~(size_t)1;
  double* tile0 = reinterpret_cast<double*>(wsm);
  uint4* stage = reinterpret_cast<uint4*>(tile0 + 2 * tile_doubles);   // [16][sweep][nchunk]
  double* xs_s = reinterpret_cast<double*>(stage + (size_t)16 * sweep * nchunk);   // [kSearchRows][kXP]
  double* rd_s = xs_s + kSearchRows * kXP;                             // [L][kXP]
  const int64_t n_act = *A.n_alive;
  const int64_t n_tiles = (n_act + 15) >> 4;
  const int64_t G = gridDim.x;
  if ((int64_t)blockIdx.x >= n_tiles) return;
  const uint32_t stride = (uint32_t)A.n_cell;
  const float* lp = A.ctm_logp;
  const float* pc = A.ctm_pcol;
  const float* pm = HAS_TROP ? A.ctm_pmid : A.ctm_logp;
  for (int i = threadIdx.x; i < 128; i += kWsThreads)
    sm.tab[i] = make_double2(g_log_table.r[i], g_log_table.neg_log_r[i]);
  __syncthreads();

  if (threadIdx.x < kWsProd) {
    // ================================================================== producers
    const int tid = threadIdx.x;
    const bool active = tid < 16 * nchunk;
    const int pp = active ? tid / nchunk : 0;          // pair slot of the tile
    const int ch = active ? tid - pp * nchunk : 0;     // chunk of its records
    const uint4* records = reinterpret_cast<const uint4*>(A.records);
    uint4* slot = stage + (pp * sweep) * nchunk + ch;  // [pair][entry][chunk]
    auto load_pid = [&](int64_t tile) -> int {         // tid < 16
      if (tile >= n_tiles) return 0;
      const int64_t s = tile * 16 + tid;
      return A.alive_pairs[s < n_act ? s : n_act - 1];  // shadow slots repeat the last pair
    };
    // table items of this thread: i = tid and (tid < 16) tid + 224, i = 15 * slot + entry
    const int c0 = tid / kWsSW, e0 = tid - c0 * kWsSW;
    const int c1 = (tid + kWsProd) / kWsSW, e1 = tid + kWsProd - c1 * kWsSW;
    const bool has1 = tid + kWsProd < 16 * kWsSW;
    uint32_t ncix0 = 0, ncix1 = 0;
    double nw0 = 0.0, nw1 = 0.0;
    auto fetch_tab = [&](int ring, int base, int nk) {
      if (e0 < nk) {
        const int64_t pr = sm.pid[ring][c0];
        const int32_t v = A.vert[pr * S + base + e0];
        nw0 = A.w[pr * S + base + e0];
        ncix0 = (uint32_t)((A.pair_record0[pr] + v) * nchunk);
      }
      if (has1 && e1 < nk) {
        const int64_t pr = sm.pid[ring][c1];
        const int32_t v = A.vert[pr * S + base + e1];
        nw1 = A.w[pr * S + base + e1];
        ncix1 = (uint32_t)((A.pair_record0[pr] + v) * nchunk);
      }
    };
    auto store_tab = [&](int sb) {
      sm.gcix[sb][tid] = ncix0;
      sm.gw[sb][tid] = nw0;
      if (has1) {
        sm.gcix[sb][tid + kWsProd] = ncix1;
        sm.gw[sb][tid + kWsProd] = nw1;
      }
    };
    // model columns of a tile into L2: thread = (pair slot, level lane), 14 level lanes
    auto prefetch_model = [&](int ring) {
      const uint32_t off = A.pair_ctm_off[sm.pid[ring][tid & 15]];
      for (int k = tid >> 4; k < n_ctm; k += kWsProd / 16) {
        const uint32_t at = off + (uint32_t)k * stride;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(lp + at));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + at));
        if (HAS_TROP) asm volatile("prefetch.global.L2 [%0];" ::"l"(pm + at));
      }
    };
    // prologue: pair ids of the first two tiles, tables of the first sweep
    if (tid < 16) {
      sm.pid[0][tid] = load_pid(blockIdx.x);
      sm.pid[1][tid] = load_pid(blockIdx.x + G);
    }
    bar_sync(kBarProd, kWsProd);
    fetch_tab(0, 0, sweep);
    store_tab(0);
    prefetch_model(0);
    bar_sync(kBarProd, kWsProd);
    int q = 0;                                         // sweeps done: parity = table buffer
    int it = 0;
    double acc[8];
    for (int64_t tile_i = blockIdx.x; tile_i < n_tiles; tile_i += G, ++it) {
      const int b = it & 1, ring = it % 3, ring1 = (it + 1) % 3, ring2 = (it + 2) % 3;
      const bool more = tile_i + G < n_tiles;
      int pid2 = 0, f_pair = 0;
      uint32_t f_off = 0;
      double f_old = 0.0;
      if (tid < 16) {
        pid2 = load_pid(tile_i + 2 * G);
        f_pair = sm.pid[ring][tid];
        f_off = A.pair_ctm_off[f_pair];
        f_old = A.staged[4 * A.n_pairs + f_pair];      // gridded by oisat_pair_alive
      }
      if (P.probe != 2 || it == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.0;
      }
      const bool gather = P.probe != 2 || it == 0;   // probe 2: every tile repeats the first one
#pragma unroll 1
      for (int base = 0; base < S; base += kWsSW, ++q) {
        const int nk = (S - base) < kWsSW ? (S - base) : kWsSW;
        const int sb = q & 1;
        if (active && gather) {
#pragma unroll
          for (int e = 0; e < kWsSW; ++e) {
            if (e < nk) {
              const uint32_t ck = sm.gcix[sb][pp * kWsSW + e] + (uint32_t)ch;
              const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slot + e * nchunk);
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(records + ck)
                           : "memory");
            }
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // while the records fly: the tables of the next sweep (of this tile or of the next one)
        const bool last = base + kWsSW >= S;
        if (!last) {
          const int nbase = base + kWsSW;
          fetch_tab(ring, nbase, (S - nbase) < kWsSW ? (S - nbase) : kWsSW);
        } else if (more) {
          fetch_tab(ring1, 0, sweep);
          prefetch_model(ring1);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (active && gather) {
#pragma unroll
          for (int e = 0; e < kWsSW; ++e) {
            if (e < nk) {
              const double wk = sm.gw[sb][pp * kWsSW + e];
              double z[8];
              h8_to_f64(slot[e * nchunk], z);
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[k] = fma(wk, z[k], acc[k]);
            }
          }
        }
        store_tab(sb ^ 1);
        bar_sync(kBarProd, kWsProd);
      }
      // ring slot (it + 2) % 3 == (it - 1) % 3: every producer is past the epilogue of tile
      // it - 1 (it has been through a barrier of this tile); read again only after the barrier below
      if (tid < 16) sm.pid[ring2][tid] = pid2;
      // the tile buffer is free once the consumers have finished tile it - 2
      bar_sync(kBarEmpty + b, kWsThreads);
      double* tile = tile0 + b * tile_doubles;
      const int sig_row = 2 * L + 1;
      if (active) {
        double sig = 0.0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int row = ch + nchunk * e;
          if (row == sig_row) sig = acc[e];
          const int orow = row <= 2 * L ? row : row - 1;                 // tropopause follows vcd
          if (row < nrow && row != sig_row) tile[orow * kTP + pp] = acc[e] * A.box_weight;
        }
        if (tile_i * 16 + pp < n_act && ch == sig_row % nchunk)
          A.staged[1 * A.n_pairs + sm.pid[ring][pp]] = sqrt(sig * A.box_weight_err);   // interpolator.py:188
      }
      if (tid < 16) {
        sm.c_pair[b][tid] = f_pair;
        sm.c_off[b][tid] = f_off;
        sm.c_old_amf[b][tid] = f_old;
      }
      __threadfence_block();
      bar_arrive(kBarFull + b, kWsThreads);
    }
    return;
  }

  // ==================================================================== consumers
  const int ct = threadIdx.x - kWsProd;
  const int p = ct & 15, t = ct >> 4;
  const int body = n_ctm - (n_ctm % 8);
  const int nb = body >> 3;                 // terms per running sum
  const int half = (nb + 1) >> 1;           // <= H
  const int j8 = t & 7;
  const int i0 = t < 8 ? 0 : half;
  const int cnt = t < 8 ? half : nb - half;
  bar_arrive(kBarEmpty + 0, kWsThreads);    // both tile buffers start empty
  bar_arrive(kBarEmpty + 1, kWsThreads);
  int it = 0;
  for (int64_t tile_i = blockIdx.x; tile_i < n_tiles; tile_i += G, ++it) {
    const int b = it & 1;
    bar_sync(kBarFull + b, kWsThreads);
    double* tile = tile0 + b * tile_doubles;
    const bool live = tile_i * 16 + p < n_act;
    const int64_t pair = sm.c_pair[b][p];
    const uint32_t off = sm.c_off[b][p];
    if (t == 0) sm.unsorted[p] = n_ctm >= 8 ? 0 : 1;
    RowViewP<kTP> r{tile + p};
    const double vcd = r.at(2 * L);
    const bool work = live && vcd == vcd && P.probe != 1;  // amf_recal.py:99-100
    const double trop = HAS_TROP ? r.at(2 * L + 1) : 0.0;
    const bool descending = r.at(L) > r.at(2 * L - 1);
    bar_sync(kBarCons, kWsCons);   // every thread has read the raw pressures before they turn into logs
    // phase A: p -> log p in place + ascending copy; padding rows = +inf
#pragma unroll
    for (int row = t; row < kSearchRows; row += 16) {
      if (row < L) {
        if (work) {
          const double lg = table_log2(r.at(L + row), sm.tab);
          r.set(L + row, lg);
          xs_s[(descending ? L - 1 - row : row) * kXP + p] = lg;
        }
      } else {
        xs_s[row * kXP + p] = CUDART_INF;
      }
    }
    bar_sync(kBarCons, kWsCons);
    float lpv[H], pcv8[H], pmv[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const int ii = i < cnt ? i : (cnt > 0 ? cnt - 1 : 0);
      const uint32_t at = off + (uint32_t)(j8 + 8 * (i0 + ii)) * stride;
      const bool ok = work && cnt > 0;
      lpv[i] = ok ? __ldg(lp + at) : 0.0f;
      pcv8[i] = ok ? __ldg(pc + at) : 0.0f;
      pmv[i] = (HAS_TROP && ok) ? __ldg(pm + at) : 0.0f;
    }
    if (work) {
      bool bad = false;
#pragma unroll
      for (int j = t; j < L; j += 16) {
        const double x = xs_s[j * kXP + p];
        const double prev = j > 0 ? xs_s[(j - 1) * kXP + p] : -CUDART_INF;
        bad = bad || !(prev < x);
        if (j > 0) rd_s[j * kXP + p] = __drcp_rn(x - prev);              // = 1.0 / (x - prev) bit for bit
      }
      if (bad) sm.unsorted[p] = 1;
    }
    bar_sync(kBarCons, kWsCons);
    const bool sorted = sm.unsorted[p] == 0;
    const bool go = work && sorted;
    const double* y0 = tile + (descending ? L - 1 : 0) * kTP + p;
    const int ystep = descending ? -kTP : kTP;
    auto term = [&](float lpf, float pcf, float pmf, double& ta, float& tb) {
      const double v = (double)lpf;
      double pcv = (double)pcf;
      int idx = 0;
#pragma unroll
      for (int step = 32; step >= 1; step >>= 1)
        idx += (xs_s[(idx + step - 1) * kXP + p] < v) ? step : 0;         // searchsorted(xs, v, 'left')
      const int c = idx < 1 ? 1 : (idx > L - 1 ? L - 1 : idx);
      const double x_hi = xs_s[c * kXP + p], x_lo = xs_s[(c - 1) * kXP + p];
      const double rden = rd_s[c * kXP + p];
      const double y_hi = y0[c * ystep], y_lo = y0[(c - 1) * ystep];
      double sw = ((v - x_lo) * rden) * y_hi + ((x_hi - v) * rden) * y_lo;  // interp1d._call_linear
      if (isinf(sw)) sw = 0.0;
      if (HAS_TROP && (double)pmf < trop) { sw = qnan(); pcv = qnan(); }
      const double prod = sw * pcv;
      ta = prod != prod ? 0.0 : prod;        // nansum terms
      tb = pcv != pcv ? 0.0f : (float)pcv;
    };
    double ta[H];
    float tb[H];
    if (go) {
#pragma unroll
      for (int i = 0; i < H; ++i) {
        ta[i] = 0.0;
        tb[i] = 0.0f;
        if (i < cnt) term(lpv[i], pcv8[i], pmv[i], ta[i], tb[i]);
      }
      if (t < 8 && body + t < n_ctm) {      // scalar tail of numpy's pairwise sum
        const uint32_t at = off + (uint32_t)(body + t) * stride;
        double a;
        float bb;
        term(__ldg(lp + at), __ldg(pc + at), HAS_TROP ? __ldg(pm + at) : 0.0f, a, bb);
        sm.tail_a[t * 16 + p] = a;
        sm.tail_b[t * 16 + p] = bb;
      }
      if (t < 8) {                            // running sum j8, first half
        double qa = ta[0];
        float qb = tb[0];
#pragma unroll
        for (int i = 1; i < H; ++i)
          if (i < cnt) { qa = qa + ta[i]; qb = __fadd_rn(qb, tb[i]); }
        sm.part_a[j8 * 16 + p] = qa;
        sm.part_b[j8 * 16 + p] = qb;
      }
    }
    bar_sync(kBarCons, kWsCons);
    if (go && t >= 8) {                       // second half continues the same running sum
      double qa = sm.part_a[j8 * 16 + p];
      float qb = sm.part_b[j8 * 16 + p];
#pragma unroll
      for (int i = 0; i < H; ++i)
        if (i < cnt) { qa = qa + ta[i]; qb = __fadd_rn(qb, tb[i]); }
      sm.part_a[j8 * 16 + p] = qa;
      sm.part_b[j8 * 16 + p] = qb;
    }
    bar_sync(kBarCons, kWsCons);
    // phase C
    if (t == 0 && live) {
      const double old_amf = sm.c_old_amf[b][p];
      double new_amf = qnan(), vnew = qnan(), colv = qnan();
      if (work) {
        double colsum = 0.0;
        if (sorted) {
          auto qa = [&](int j) { return sm.part_a[j * 16 + p]; };
          auto qb = [&](int j) { return sm.part_b[j * 16 + p]; };
          double scd = ((qa(0) + qa(1)) + (qa(2) + qa(3))) + ((qa(4) + qa(5)) + (qa(6) + qa(7)));
          float cs = __fadd_rn(__fadd_rn(__fadd_rn(qb(0), qb(1)), __fadd_rn(qb(2), qb(3))),
                               __fadd_rn(__fadd_rn(qb(4), qb(5)), __fadd_rn(qb(6), qb(7))));
          for (int k = body; k < n_ctm; ++k) {   // scalar tail after the tree
            scd = scd + sm.tail_a[(k - body) * 16 + p];
            cs = __fadd_rn(cs, sm.tail_b[(k - body) * 16 + p]);
          }
          colsum = (double)cs;
          new_amf = colsum != 0.0 ? scd / colsum : qnan();
        } else {
          new_amf = amf_slow_path_tile(r, L, n_ctm, HAS_TROP, trop, lp + off, pc + off, pm + off,
                                       (int64_t)stride, &colsum);
        }
        vnew = (old_amf * vcd) / new_amf;                        // amf_recal.py:179
        colv = (vnew != vnew || isinf(vnew)) ? qnan() : colsum;  // :180-181
      }
      A.staged[0 * A.n_pairs + pair] = vnew;
      A.staged[2 * A.n_pairs + pair] = colv;
      A.staged[3 * A.n_pairs + pair] = new_amf;
    }
    // phase C has read the partial sums and the tile: hand the buffer back, and keep the fast
    // threads out of the next tile's shared arrays until it is through
    bar_sync(kBarCons, kWsCons);
    bar_arrive(kBarEmpty + b, kWsThreads);
  }
}

}  // namespace oisat

using namespace oisat;

static int upload_log_table() {
  static std::once_flag once[64];
  static int status[64];
  int dev = 0;
  OISAT_CHECK_CUDA(cudaGetDevice(&dev));
  OISAT_CHECK_ARG(dev >= 0 && dev < 64, "device ordinal out of range");
  std::call_once(once[dev], [dev]() {
    LogTable h;
    for (int i = 0; i < 128; ++i) {
      const double c = 1.0 + (i + 0.5) / 128.0;
      h.r[i] = 1.0 / c;
      h.neg_log_r[i] = -std::log(h.r[i]);
    }
    status[dev] = (int)cudaMemcpyToSymbol(g_log_table, &h, sizeof(h));
  });
  OISAT_CHECK_CUDA((cudaError_t)status[dev]);
  return OISAT_OK;
}

extern "C" int64_t oisat_rows_per_pair(int32_t n_sat_lev, int32_t has_trop) {
  return 2 * (int64_t)n_sat_lev + 1 + (has_trop ? 1 : 0);
}

extern "C" int oisat_fused_amf_split(const oisat_fused_args* h_args, double* rows, void* stream) {
  OISAT_CHECK_ARG(h_args != nullptr, "null args");
  const oisat_fused_args& a = *h_args;
  if (a.n_pairs == 0) return OISAT_OK;
  OISAT_CHECK_ARG(a.vert && a.w && a.gran_record0 && a.gran_px0 && a.gran_slot && a.records &&
                      a.amf_masked && a.ctm_logp && a.ctm_pcol && a.staged && a.pair_cell &&
                      a.pair_granule && rows, "null pointer");
  OISAT_CHECK_ARG(!a.has_trop || a.ctm_pmid, "tropopause masking needs the model p_mid");
  OISAT_CHECK_ARG(a.nwin >= 1 && a.n_sat_lev >= 2 && a.n_sat_lev <= kMaxSatLev, "bad stencil");
  OISAT_CHECK_ARG(a.n_ctm_lev >= 2 && a.n_ctm_lev <= kMaxCtmLev, "bad model level count");
  SplitParams P;
  P.a = a;
  P.rows = rows;
  P.nrow = rec_rows(a.n_sat_lev, a.has_trop);
  P.nchunk = rec_chunks(a.n_sat_lev, a.has_trop);
  P.nrow_out = (int)oisat_rows_per_pair(a.n_sat_lev, a.has_trop);
  P.probe = 0;
  OISAT_CHECK_ARG(P.nchunk < 16, "record too wide for the half-warp gather: use oisat_fused_amf");
  OISAT_CHECK_ARG(a.n_records > 0 && a.n_records * P.nchunk < ((int64_t)1 << 32),
                  "record block too large for 32-bit chunk indices: split the batch");
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = upload_log_table()) return rc;
  const int sweep = 3 * a.nwin < 15 ? 3 * a.nwin : 15;
  const size_t tile_bytes = (size_t)16 * sweep * P.nchunk * sizeof(uint4) +
                            (size_t)P.nrow_out * 17 * sizeof(double);
  OISAT_CHECK_CUDA(cudaFuncSetAttribute(gather_rows_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));
  gather_rows_kernel<<<(unsigned)ceil_div(a.n_pairs * 16, kGatherThreads), kGatherThreads,
                       tile_bytes, s>>>(P);
  OISAT_CHECK_LAUNCH();
  const size_t stage_bytes = (size_t)P.nrow_out * 16 * sizeof(double) +
                             (size_t)a.n_ctm_lev * 16 * (sizeof(double) + sizeof(float));
  const unsigned vblocks = (unsigned)ceil_div(a.n_pairs, 16);
  if (a.has_trop) {
    OISAT_CHECK_CUDA(cudaFuncSetAttribute(vertical_rows_kernel<true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)stage_bytes));
    vertical_rows_kernel<true><<<vblocks, kVerticalThreads, stage_bytes, s>>>(P);
  } else {
    OISAT_CHECK_CUDA(cudaFuncSetAttribute(vertical_rows_kernel<false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)stage_bytes));
    vertical_rows_kernel<false><<<vblocks, kVerticalThreads, stage_bytes, s>>>(P);
  }
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

template <bool HAS_TROP, int H, int CL, int CS, int CN, bool PACKED = false, int SW = 15, int MINB = 4,
          bool BULK = false>
static int launch_tile(const SplitParams& P, cudaStream_t s) {
  const oisat_fused_args& a = P.a;
  const int sweep = 3 * a.nwin < SW ? 3 * a.nwin : SW;
  const size_t tile_bytes = ((size_t)P.nrow_out * kTP * sizeof(double) + 15) & ~(size_t)15;
  const size_t stage_bytes = (size_t)16 * sweep * P.nchunk * sizeof(uint4);
  const size_t vert_bytes = (size_t)(kSearchRows + a.n_sat_lev) * kXP * sizeof(double);
  const size_t smem = tile_bytes + (stage_bytes > vert_bytes ? stage_bytes : vert_bytes);
  OISAT_CHECK_CUDA(cudaFuncSetAttribute(fused_tile_kernel<HAS_TROP, H, CL, CS, CN, SW, MINB, PACKED, BULK>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fused_tile_kernel<HAS_TROP, H, CL, CS, CN, SW, MINB, PACKED, BULK>
      <<<(unsigned)ceil_div(P.a.n_pairs, 16), kTileThreads, smem, s>>>(P);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

template <bool HAS_TROP, int H, int CL, int CS, int CN>
static int launch_ws(const SplitParams& P, cudaStream_t s) {
  const oisat_fused_args& a = P.a;
  const int sweep = 3 * a.nwin < kWsSW ? 3 * a.nwin : kWsSW;
  const size_t tile_doubles = ((size_t)P.nrow_out * kTP + 1) & ~(size_t)1;
  const size_t smem = 2 * tile_doubles * sizeof(double) + (size_t)16 * sweep * P.nchunk * sizeof(uint4) +
                      (size_t)(kSearchRows + a.n_sat_lev) * kXP * sizeof(double);
  static int sms[64];
  int dev = 0;
  OISAT_CHECK_CUDA(cudaGetDevice(&dev));
  if (sms[dev & 63] == 0)
    OISAT_CHECK_CUDA(cudaDeviceGetAttribute(&sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  auto kern = fused_ws_kernel<HAS_TROP, H, CL, CS, CN>;
  OISAT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  OISAT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWsThreads, smem));
  OISAT_CHECK_ARG(per_sm >= 1, "the warp-specialised tile kernel does not fit on this device");
  int64_t grid = (int64_t)per_sm * sms[dev & 63];
  const int64_t tiles = ceil_div(a.n_pairs, 16);
  if (grid > tiles) grid = tiles;
  kern<<<(unsigned)grid, kWsThreads, smem, s>>>(P);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_fused_amf_tile(const oisat_fused_args* h_args, void* stream) {
  OISAT_CHECK_ARG(h_args != nullptr, "null args");
  const oisat_fused_args& a = *h_args;
  if (a.n_pairs == 0) return OISAT_OK;
  OISAT_CHECK_ARG(a.vert && a.w && a.gran_record0 && a.gran_px0 && a.gran_slot && a.records &&
                      a.ctm_logp && a.ctm_pcol && a.staged && a.pair_cell &&
                      a.pair_granule, "null pointer");
  OISAT_CHECK_ARG(a.alive_pairs && a.n_alive,
                  "the tile form runs over the live pairs: call oisat_pair_alive first");
  OISAT_CHECK_ARG(!a.has_trop || a.ctm_pmid, "tropopause masking needs the model p_mid");
  OISAT_CHECK_ARG(a.nwin >= 1 && a.n_sat_lev >= 2 && a.n_sat_lev <= kSearchRows - 1, "bad stencil");
  OISAT_CHECK_ARG(a.n_ctm_lev >= 2 && a.n_ctm_lev <= kMaxCtmLev, "bad model level count");
  SplitParams P;
  P.a = a;
  P.rows = nullptr;
  P.nrow = rec_rows(a.n_sat_lev, a.has_trop);
  P.nchunk = rec_chunks(a.n_sat_lev, a.has_trop);
  P.nrow_out = (int)oisat_rows_per_pair(a.n_sat_lev, a.has_trop);
  P.probe = 0;
  OISAT_CHECK_ARG(P.nchunk < 16, "record too wide for the half-warp gather: use oisat_fused_amf");
  OISAT_CHECK_ARG(a.n_records > 0 && a.n_records * P.nchunk < ((int64_t)1 << 32),
                  "record block too large for 32-bit chunk indices: split the batch");
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = upload_log_table()) return rc;
  OISAT_CHECK_ARG((int64_t)a.n_ctm_lev * a.n_cell * 8 < ((int64_t)1 << 31) && a.n_cell < ((int64_t)1 << 31),
                  "model fields too large for 32-bit element offsets: use oisat_fused_amf_split");
  const int S = 3 * a.nwin, L = a.n_sat_lev, N = a.n_ctm_lev;
  // the BASELINE products, compiled with their geometry as constants (OISAT_TILE_GENERIC=1
  // forces the run-time build: the tests compare the two)
  const char* gen = getenv("OISAT_TILE_GENERIC");
  const bool generic = gen && gen[0] == '1';
  const char* pk = getenv("OISAT_TILE_PACKED");   // "0": half-warp-per-pair gather lanes (A/B runs, tests)
  const bool packed = !(pk && pk[0] == '0');
  // records fetched with one bulk asynchronous copy (TMA engine) per (pair, entry) -- the default
  // for the packed builds; OISAT_TILE_BULK=0: 16-byte cp.async per lane (A/B runs, tests)
  const char* bk = getenv("OISAT_TILE_BULK");
  const bool bulk = (bk && bk[0] == '1') && (reinterpret_cast<uintptr_t>(a.records) & 15) == 0;
  // OISAT_TILE_WS=1 runs the warp-specialised persistent form (needs the per-pair tables and
  // records of at most 14 chunks).  Measured on the OMI HCHO month (profiles/r02_ws_probe.md):
  // 5.70 ms against 5.21 ms for the one-tile-per-block form -- each role alone needs ~4.3 ms
  // with its 14-16 warps per SM, both are latency bound, and the 32 interchangeable warps of
  // the one-tile-per-block form hide latency better than two fixed groups do -- so it is NOT
  // the default; the tests keep the two forms bit-identical.
  const char* prenv = getenv("OISAT_WS_PROBE");
  P.probe = prenv ? atoi(prenv) : 0;
  const char* wsenv = getenv("OISAT_TILE_WS");
  const bool ws = (wsenv && wsenv[0] == '1') && a.pair_record0 && a.pair_ctm_off && P.nchunk <= 14;
  if (ws) {
    if (generic) {
    } else if (!a.has_trop && L == 47 && S == 12 && N == 72) {
      return launch_ws<false, 5, 47, 12, 72>(P, s);
    } else if (a.has_trop && L == 35 && S == 12 && N == 72) {
      return launch_ws<true, 5, 35, 12, 72>(P, s);
    } else if (a.has_trop && L == 34 && S == 90 && N == 72) {
      return launch_ws<true, 5, 34, 90, 72>(P, s);
    }
    if (((N / 8) + 1) / 2 <= 5)
      return a.has_trop ? launch_ws<true, 5, 0, 0, 0>(P, s) : launch_ws<false, 5, 0, 0, 0>(P, s);
    return a.has_trop ? launch_ws<true, 8, 0, 0, 0>(P, s) : launch_ws<false, 8, 0, 0, 0>(P, s);
  }
  if (generic) {
  } else if (!a.has_trop && L == 47 && S == 12 && N == 72) {
    if (packed && bulk) return launch_tile<false, 5, 47, 12, 72, true, 15, 4, true>(P, s);
    return packed ? launch_tile<false, 5, 47, 12, 72, true>(P, s) : launch_tile<false, 5, 47, 12, 72>(P, s);
  } else if (a.has_trop && L == 35 && S == 12 && N == 72) {
    if (packed && bulk) return launch_tile<true, 5, 35, 12, 72, true, 15, 4, true>(P, s);
    return packed ? launch_tile<true, 5, 35, 12, 72, true>(P, s) : launch_tile<true, 5, 35, 12, 72>(P, s);
  } else if (a.has_trop && L == 34 && S == 90 && N == 72) {
    if (packed && bulk) return launch_tile<true, 5, 34, 90, 72, true, 15, 4, true>(P, s);
    return packed ? launch_tile<true, 5, 34, 90, 72, true>(P, s) : launch_tile<true, 5, 34, 90, 72>(P, s);
  }
  const int half = ((N / 8) + 1) / 2;
  if (half <= 5)
    return a.has_trop ? launch_tile<true, 5, 0, 0, 0>(P, s) : launch_tile<false, 5, 0, 0, 0>(P, s);
  return a.has_trop ? launch_tile<true, 8, 0, 0, 0>(P, s) : launch_tile<false, 8, 0, 0, 0>(P, s);
}
