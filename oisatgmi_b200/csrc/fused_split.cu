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
#include <mutex>

#include "vertical.cuh"

namespace oisat {

__host__ __device__ inline int rec_rows(int L, int has_trop) { return 2 * L + 2 + (has_trop ? 1 : 0); }
__host__ __device__ inline int rec_chunks(int L, int has_trop) { return (rec_rows(L, has_trop) + 7) / 8; }

struct SplitParams {
  oisat_fused_args a;
  double* rows;              // [ceil(n_pairs/16)][nrow_out][16]
  int nrow, nchunk, nrow_out;
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
    for (int node = 0; node < nk; node += 3) {
      uint4 u[3];
      double wk[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint32_t ck = __shfl_sync(0xffffffffu, cix, node + j, 16);
        wk[j] = __shfl_sync(0xffffffffu, wt, node + j, 16);
        if (gl < P.nchunk) u[j] = __ldg(&records[ck + gl]);
      }
      if (gl < P.nchunk) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double z[8];
          h8_to_f64(u[j], z);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fma(wk[j], z[e], acc[e]);
        }
      }
    }
  }
  // The 16 pairs of this block form one tile of the row buffer.  Rows are
  // transposed through shared memory ([row][pair], pitch 17: conflict-free for a
  // half warp writing consecutive rows) so that the global stores are full 128-byte
  // lines instead of one 8-byte store per row and pair.
  extern __shared__ double tile[];  // [nrow_out][17]
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

constexpr int kVerticalThreads = 64;   // four 16-pair tiles of the row buffer per block

// The rows of a block's pairs are one contiguous piece of the buffer (4 tiles x
// nrow_out x 16 doubles, ~48 KB for OMI HCHO): one bulk asynchronous copy brings
// it into shared memory while the other resident blocks compute.  Before this,
// every bracket shift of the walk below waited for its own global load -- and a
// warp shifts whenever ANY of its lanes does, ~150 exposed DRAM round trips per
// warp (ncu: 60% of the stall samples were long-scoreboard, IPC 0.3).
template <bool HAS_TROP>
__global__ void __launch_bounds__(kVerticalThreads)
vertical_rows_kernel(const __grid_constant__ SplitParams P) {
  extern __shared__ __align__(128) double rows_s[];  // [4][nrow_out][16]
  __shared__ LogTable tab;
  __shared__ __align__(8) unsigned long long mbar;
  const oisat_fused_args& A = P.a;
  const int L = A.n_sat_lev, n_ctm = A.n_ctm_lev;
  const int64_t tile0 = (int64_t)blockIdx.x * (kVerticalThreads / 16);
  const int64_t n_tiles = (A.n_pairs + 15) >> 4;
  const int tiles_here = (int)(n_tiles - tile0 < kVerticalThreads / 16 ? n_tiles - tile0
                                                                       : kVerticalThreads / 16);
  const uint32_t tile_bytes = (uint32_t)P.nrow_out * 16u * (uint32_t)sizeof(double);
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(bar), "r"(tile_bytes * (uint32_t)tiles_here) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"((uint32_t)__cvta_generic_to_shared(rows_s)),
          "l"(P.rows + tile0 * (int64_t)P.nrow_out * 16), "r"(tile_bytes * (uint32_t)tiles_here),
          "r"(bar)
        : "memory");
  }
  for (int i = threadIdx.x; i < 128; i += kVerticalThreads) {
    tab.r[i] = g_log_table.r[i];
    tab.neg_log_r[i] = g_log_table.neg_log_r[i];
  }
  __syncthreads();  // barrier initialised, table staged
  const int64_t pair = (int64_t)blockIdx.x * kVerticalThreads + threadIdx.x;
  if (pair >= A.n_pairs) return;
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p;\n\t"
                   "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                   "  selp.b32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar) : "memory");
  }
  RowView r{rows_s + (threadIdx.x >> 4) * (P.nrow_out * 16) + (threadIdx.x & 15)};
  const double vcd = r.at(2 * L);
  const double old_amf = A.staged[4 * A.n_pairs + pair];
  double new_amf = qnan(), vnew = qnan(), col = qnan();
  if (vcd == vcd) {  // amf_recal.py:99-100
    const double trop = HAS_TROP ? r.at(2 * L + 1) : 0.0;
    // model column of this pair's cell: consecutive pairs are consecutive cells, so
    // the 32 lanes of a warp read (mostly) one or two 128-byte lines per level
    const int64_t off = (int64_t)A.gran_slot[A.pair_granule[pair]] * n_ctm * A.n_cell +
                        A.pair_cell[pair];
    const float* lp = A.ctm_logp + off;
    const float* pc = A.ctm_pcol + off;
    const float* pm = HAS_TROP ? A.ctm_pmid + off : lp;
    const int64_t stride = A.n_cell;
    // scipy sorts the levels ascending in log p: xs[j], j = 0..L-1; for a monotone
    // profile that is the storage order or its reverse.
    const bool descending = r.at(L) > r.at(2 * L - 1);
    auto row_of = [&](int j) { return descending ? L - 1 - j : j; };
    // Phase A, the same instruction stream for every lane: p -> log p, in place in
    // this thread's own column of the staged rows (sorted order is the storage
    // order or its reverse).  Strict monotonicity is checked for every level; ties
    // and NaNs take scipy's argsort path (slow path).
    bool sorted = n_ctm >= 8;
    {
      double prev = -CUDART_INF;
#pragma unroll 4
      for (int j = 0; j < L; ++j) {
        const int row = L + row_of(j);
        const double x = table_log(r.at(row), &tab);
        r.set(row, x);
        sorted = sorted && (prev < x);
        prev = x;
      }
    }
    // Phase B: interp1d(xs, SW)(log p_model) as a merge: model levels run from the
    // surface up, so the query only decreases and so does idx = searchsorted(xs, v,
    // 'left').  The bracket [c-1, c], c = clip(idx, 1, L-1), lives in registers.
    int idx = L, c = L - 1;
    double x_hi = r.at(L + row_of(c)), x_lo = r.at(L + row_of(c - 1));
    double y_hi = r.at(row_of(c)), y_lo = r.at(row_of(c - 1));
    double rden = 1.0 / (x_hi - x_lo);
    double qa[8], colsum = 0.0, scd = 0.0;
    float qb[8];
    const int body = n_ctm - (n_ctm % 8);
    for (int k0 = 0; k0 < n_ctm && sorted; k0 += 8) {
      float lpv[8], pcv8[8], pmv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // eight levels of the model column in flight at once
        const int k = k0 + j < n_ctm ? k0 + j : n_ctm - 1;
        lpv[j] = __ldg(lp + (int64_t)k * stride);
        pcv8[j] = __ldg(pc + (int64_t)k * stride);
        pmv[j] = HAS_TROP ? __ldg(pm + (int64_t)k * stride) : 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + j;
        if (k < n_ctm) {
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
              x_lo = r.at(L + row_of(c - 1));
              y_lo = r.at(row_of(c - 1));
              rden = 1.0 / (x_hi - x_lo);
            }
          }
          double sw = ((v - x_lo) * rden) * y_hi + ((x_hi - v) * rden) * y_lo;  // interp1d._call_linear
          if (isinf(sw)) sw = 0.0;
          if (HAS_TROP && (double)pmv[j] < trop) { sw = qnan(); pcv = qnan(); }
          const double prod = sw * pcv;
          const double a = prod != prod ? 0.0 : prod;        // nansum
          const float b = pcv != pcv ? 0.0f : (float)pcv;
          if (k < body) {                                    // numpy's eight running sums
            if (k0 == 0) { qa[j] = a; qb[j] = b; }
            else { qa[j] = qa[j] + a; qb[j] = __fadd_rn(qb[j], b); }
          } else {                                           // scalar tail after the tree
            if (k == body) {
              scd = ((qa[0] + qa[1]) + (qa[2] + qa[3])) + ((qa[4] + qa[5]) + (qa[6] + qa[7]));
              colsum = (double)__fadd_rn(__fadd_rn(__fadd_rn(qb[0], qb[1]), __fadd_rn(qb[2], qb[3])),
                                         __fadd_rn(__fadd_rn(qb[4], qb[5]), __fadd_rn(qb[6], qb[7])));
            }
            scd = scd + a;
            colsum = (double)__fadd_rn((float)colsum, b);
          }
        }
      }
    }
    if (sorted && body == n_ctm) {
      scd = ((qa[0] + qa[1]) + (qa[2] + qa[3])) + ((qa[4] + qa[5]) + (qa[6] + qa[7]));
      colsum = (double)__fadd_rn(__fadd_rn(__fadd_rn(qb[0], qb[1]), __fadd_rn(qb[2], qb[3])),
                                 __fadd_rn(__fadd_rn(qb[4], qb[5]), __fadd_rn(qb[6], qb[7])));
    }
    if (sorted) {
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
  OISAT_CHECK_ARG(P.nchunk < 16, "record too wide for the half-warp gather: use oisat_fused_amf");
  OISAT_CHECK_ARG(a.n_records > 0 && a.n_records * P.nchunk < ((int64_t)1 << 32),
                  "record block too large for 32-bit chunk indices: split the batch");
  cudaStream_t s = (cudaStream_t)stream;
  if (int rc = upload_log_table()) return rc;
  const size_t tile_bytes = (size_t)P.nrow_out * 17 * sizeof(double);
  gather_rows_kernel<<<(unsigned)ceil_div(a.n_pairs * 16, kGatherThreads), kGatherThreads,
                       tile_bytes, s>>>(P);
  OISAT_CHECK_LAUNCH();
  const size_t stage_bytes = (size_t)(kVerticalThreads / 16) * P.nrow_out * 16 * sizeof(double);
  const unsigned vblocks = (unsigned)ceil_div(a.n_pairs, kVerticalThreads);
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
