// K5: element-wise optimal interpolation with the 99-factor regularisation sweep
// (optimal_interpolation.py:14-52; Sa/So/bias from driver.py:65-114).
//
// The reference makes 99 full passes over the grid and keeps 3x99 full-size
// arrays alive; only nanmean(AK_r) per factor feeds the knee detector.  Here one
// pass reads Sa/So once and produces all 99 sums.  The sums follow numpy's
// pairwise summation tree exactly (leaves of <=128 elements with eight running
// partial sums, combined by the same recursive halving), so the 99 means -- and
// therefore the discrete knee decision taken on the host -- are bit-identical
// to np.nanmean(AK.flatten()).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace oisat {

constexpr int kMaxFactors = 128;
constexpr int kLeafMax = 128;

struct Factors {
  double r[kMaxFactors];
  int n;
};

// AK_r for one element, operation by operation as numpy evaluates
// optimal_interpolation.py:26-30
__device__ __forceinline__ void oi_terms(double Sa, double So, double r, double* K, double* Sb,
                                         double* AK) {
  const double sr = __dmul_rn(Sa, r);
  const double k = __dmul_rn(sr, __drcp_rn(__dadd_rn(sr, So)));
  const double sb = __dmul_rn(__dmul_rn(__dsub_rn(1.0, k), Sa), r);
  *K = k;
  *Sb = sb;
  *AK = __dsub_rn(1.0, __ddiv_rn(sb, sr));
}

// Eight lanes per leaf; lane j owns elements j, j+8, ... of the leaf (numpy's r[j]).
__global__ void __launch_bounds__(256)
oi_sweep_leaf_kernel(const double* __restrict__ Sa, const double* __restrict__ So,
                     const int64_t* __restrict__ leaf_start, int64_t n_leaf,
                     const __grid_constant__ Factors fac, double* __restrict__ leaf_sum,
                     int64_t sum_stride, double* __restrict__ leaf_cnt) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;  // leaf id
  const int j = threadIdx.x & 7;
  const bool active = g < n_leaf;
  int64_t start = 0;
  int len = 0;
  if (active) { start = leaf_start[g]; len = (int)(leaf_start[g + 1] - start); }
  const int body = len < 8 ? 0 : len - (len % 8);
  double sa[kLeafMax / 8], so[kLeafMax / 8];
  int mine = 0;
  for (int i = j; i < body; i += 8) { sa[mine] = Sa[start + i]; so[mine] = So[start + i]; ++mine; }
  const unsigned group_mask = 0xffu << (threadIdx.x & 24);
  // the factors are dealt to gridDim.y blocks: one block per leaf group would leave most
  // SMs idle (1,625 leaves of 8 lanes = 51 blocks) behind a 99-factor serial loop
  for (int f = blockIdx.y; f < fac.n; f += gridDim.y) {
    const double r = fac.r[f];
    double acc = 0.0, cnt = 0.0, K, Sb, AK;
    for (int m = 0; m < mine; ++m) {
      oi_terms(sa[m], so[m], r, &K, &Sb, &AK);
      const bool fin = AK == AK;
      const double v = fin ? AK : 0.0;
      acc = m == 0 ? v : __dadd_rn(acc, v);
      cnt += fin ? 1.0 : 0.0;
    }
    // ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) within the 8-lane group
    acc = __dadd_rn(acc, __shfl_xor_sync(group_mask, acc, 1));
    acc = __dadd_rn(acc, __shfl_xor_sync(group_mask, acc, 2));
    acc = __dadd_rn(acc, __shfl_xor_sync(group_mask, acc, 4));
    cnt += __shfl_xor_sync(group_mask, cnt, 1);
    cnt += __shfl_xor_sync(group_mask, cnt, 2);
    cnt += __shfl_xor_sync(group_mask, cnt, 4);
    if (active && j == 0) {
      if (len < 8) acc = 0.0;
      for (int i = body; i < len; ++i) {  // numpy's scalar tail (and the n<8 loop)
        oi_terms(Sa[start + i], So[start + i], r, &K, &Sb, &AK);
        const bool fin = AK == AK;
        acc = __dadd_rn(acc, fin ? AK : 0.0);
        cnt += fin ? 1.0 : 0.0;
      }
      leaf_sum[(int64_t)f * sum_stride + g] = acc;
      leaf_cnt[(int64_t)f * n_leaf + g] = cnt;
    }
  }
}

// numpy's recursive halving over the leaves, level by level: the host lays the
// (irregular) summation tree out as a schedule -- internal node i adds nodes
// left[i] and right[i]; nodes are grouped by depth -- and one block per factor
// walks the levels, so the ~3000 additions of a 361x576 grid take ~12 steps
// instead of one thread's 3000 sequential adds.  Same tree => same rounding.
__global__ void __launch_bounds__(256)
oi_sweep_combine_kernel(double* __restrict__ val, int64_t val_stride,
                        const double* __restrict__ leaf_cnt, int64_t n_leaf,
                        const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                        const int32_t* __restrict__ level_start, int n_levels,
                        double* __restrict__ sums, double* __restrict__ counts) {
  const int f = blockIdx.x;
  double* v = val + (int64_t)f * val_stride;
  for (int lv = 0; lv < n_levels; ++lv) {
    for (int i = level_start[lv] + threadIdx.x; i < level_start[lv + 1]; i += blockDim.x)
      v[n_leaf + i] = __dadd_rn(v[left[i]], v[right[i]]);
    __syncthreads();
  }
  // counts are integers: any order
  __shared__ double part[256];
  double c = 0.0;
  for (int64_t i = threadIdx.x; i < n_leaf; i += blockDim.x) c += leaf_cnt[(int64_t)f * n_leaf + i];
  part[threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    counts[f] = part[0];
    sums[f] = v[n_leaf > 1 ? 2 * n_leaf - 2 : 0];  // the root is the last internal node
  }
}

__global__ void __launch_bounds__(256)
oi_prepare_kernel(const double* __restrict__ xa, double* __restrict__ y,
                  const double* __restrict__ sigma, int64_t n, double bias_a, double bias_b,
                  double err_pct, double* __restrict__ Sa, double* __restrict__ So) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double v = __ddiv_rn(__dsub_rn(y[c], bias_a), bias_b);  // driver.py:65-106
  if (v < 0.0) v = 0.0;                                    // optimal_interpolation.py:14
  y[c] = v;
  const double t = __ddiv_rn(__dmul_rn(xa[c], err_pct), 100.0);  // driver.py:111
  Sa[c] = __dmul_rn(t, t);
  const double s = sigma[c];
  So[c] = __dmul_rn(s, s);
}

__global__ void __launch_bounds__(256)
oi_apply_kernel(const double* __restrict__ xa, const double* __restrict__ y,
                const double* __restrict__ Sa, const double* __restrict__ So, int64_t n, double r,
                double* __restrict__ xb, double* __restrict__ ak, double* __restrict__ inc,
                double* __restrict__ err) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double K, Sb, AK;
  oi_terms(Sa[c], So[c], r, &K, &Sb, &AK);
  const double d = __dmul_rn(K, __dsub_rn(y[c], xa[c]));  // optimal_interpolation.py:49
  inc[c] = d;
  xb[c] = __dadd_rn(xa[c], d);
  ak[c] = AK;
  err[c] = sqrt(Sb);
}

static void leaf_bounds(int64_t begin, int64_t n, std::vector<int64_t>& starts) {
  if (n <= kLeafMax) { starts.push_back(begin); return; }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  leaf_bounds(begin, n2, starts);
  leaf_bounds(begin + n2, n - n2, starts);
}

// summation tree as a levelled schedule; returns the node id of the subtree's root
struct TreeNode { int32_t l, r, depth; };
static int32_t build_tree(int64_t n, int32_t* next_leaf, std::vector<TreeNode>& nodes,
                          int64_t n_leaf, int32_t* depth_out) {
  if (n <= kLeafMax) { *depth_out = 0; return (*next_leaf)++; }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  int32_t dl, dr;
  const int32_t l = build_tree(n2, next_leaf, nodes, n_leaf, &dl);
  const int32_t r = build_tree(n - n2, next_leaf, nodes, n_leaf, &dr);
  const int32_t d = 1 + (dl > dr ? dl : dr);
  nodes.push_back({l, r, d});
  *depth_out = d;
  return (int32_t)(n_leaf + nodes.size() - 1);   // provisional id, remapped below
}

static int64_t leaf_count(int64_t n) {
  if (n <= kLeafMax) return 1;
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  return leaf_count(n2) + leaf_count(n - n2);
}

}  // namespace oisat

using namespace oisat;

extern "C" int oisat_oi_prepare(const double* xa, double* y, const double* sigma, int64_t n,
                                double bias_a, double bias_b, double err_pct, double* Sa,
                                double* So, void* stream) {
  if (n == 0) return OISAT_OK;
  OISAT_CHECK_ARG(xa && y && sigma && Sa && So && n > 0, "null pointer");
  oi_prepare_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      xa, y, sigma, n, bias_a, bias_b, err_pct, Sa, So);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int64_t oisat_oi_sweep_workspace(int64_t n, int32_t n_factors) {
  if (n <= 0 || n_factors <= 0) return 0;
  const int64_t nl = leaf_count(n);
  // leaf starts | schedule (left, right, level starts) | per-factor node values | leaf counts
  return (nl + 1) * (int64_t)sizeof(int64_t) + (2 * nl + 72) * (int64_t)sizeof(int32_t) +
         3 * nl * n_factors * (int64_t)sizeof(double) + 64;
}

extern "C" int oisat_oi_sweep(const double* Sa, const double* So, int64_t n,
                              const double* h_factors, int32_t n_factors, double* sums,
                              double* counts, void* work, int64_t work_bytes, void* stream) {
  OISAT_CHECK_ARG(Sa && So && h_factors && sums && counts && work, "null pointer");
  OISAT_CHECK_ARG(n > 0 && n_factors >= 1 && n_factors <= kMaxFactors, "bad extent");
  OISAT_CHECK_ARG(work_bytes >= oisat_oi_sweep_workspace(n, n_factors), "workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<int64_t> starts;
  leaf_bounds(0, n, starts);
  starts.push_back(n);
  const int64_t nl = (int64_t)starts.size() - 1;
  // schedule of the summation tree: internal nodes sorted by depth (stable), ids remapped
  std::vector<TreeNode> nodes;
  int32_t next_leaf = 0, root_depth = 0;
  build_tree(n, &next_leaf, nodes, nl, &root_depth);
  const int32_t n_int = (int32_t)nodes.size();
  std::vector<int32_t> order(n_int), newid(n_int), left(n_int), right(n_int);
  for (int32_t i = 0; i < n_int; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(),
                   [&](int32_t a, int32_t b) { return nodes[a].depth < nodes[b].depth; });
  for (int32_t k = 0; k < n_int; ++k) newid[order[k]] = k;
  std::vector<int32_t> level_start;
  for (int32_t k = 0; k < n_int; ++k) {
    const TreeNode& t = nodes[order[k]];
    if (k == 0 || t.depth != nodes[order[k - 1]].depth) level_start.push_back(k);
    left[k] = t.l < nl ? t.l : (int32_t)(nl + newid[t.l - nl]);
    right[k] = t.r < nl ? t.r : (int32_t)(nl + newid[t.r - nl]);
  }
  level_start.push_back(n_int);
  const int n_levels = (int)level_start.size() - 1;
  OISAT_CHECK_ARG(n_levels <= 64, "summation tree too deep");
  // the root must be the last node of the last level: it is the only node of maximal depth
  int64_t* d_start = (int64_t*)work;
  int32_t* d_left = (int32_t*)(d_start + nl + 1);
  int32_t* d_right = d_left + nl;
  int32_t* d_level = d_right + nl;
  double* d_val = (double*)(((uintptr_t)(d_level + 72) + 63) / 64 * 64);
  double* d_cnt = d_val + 2 * nl * n_factors;
  // pageable sources: the runtime stages them before returning, the vectors may die
  OISAT_CHECK_CUDA(cudaMemcpyAsync(d_start, starts.data(), (nl + 1) * sizeof(int64_t),
                                   cudaMemcpyHostToDevice, s));
  if (n_int > 0) {
    OISAT_CHECK_CUDA(cudaMemcpyAsync(d_left, left.data(), n_int * sizeof(int32_t),
                                     cudaMemcpyHostToDevice, s));
    OISAT_CHECK_CUDA(cudaMemcpyAsync(d_right, right.data(), n_int * sizeof(int32_t),
                                     cudaMemcpyHostToDevice, s));
  }
  OISAT_CHECK_CUDA(cudaMemcpyAsync(d_level, level_start.data(),
                                   level_start.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                                   s));
  Factors fac;
  fac.n = n_factors;
  for (int i = 0; i < n_factors; ++i) fac.r[i] = h_factors[i];
  const unsigned leaf_blocks = (unsigned)ceil_div(nl * 8, 256);
  unsigned fac_blocks = leaf_blocks >= 592 ? 1 : (592 + leaf_blocks - 1) / leaf_blocks;  // ~4 per SM
  if (fac_blocks > (unsigned)n_factors) fac_blocks = (unsigned)n_factors;
  oi_sweep_leaf_kernel<<<dim3(leaf_blocks, fac_blocks), 256, 0, s>>>(Sa, So, d_start, nl, fac, d_val,
                                                                    2 * nl, d_cnt);
  OISAT_CHECK_LAUNCH();
  oi_sweep_combine_kernel<<<(unsigned)n_factors, 256, 0, s>>>(d_val, 2 * nl, d_cnt, nl, d_left,
                                                             d_right, d_level, n_levels, sums,
                                                             counts);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

// ---------------------------------------------------------------- knee ----
// The knee of the (factor, nanmean(AK_r)) curve on the device: the Kneedle algorithm
// (Satopaa et al., 2011) for the one configuration the reference uses -- concave,
// increasing, S = 1, first knee (optimal_interpolation.py:37-41; kneed 0.8.x, restated in
// oisatgmi_b200/kneedle.py, whose every numpy expression is repeated here in the same
// order: normalisation by min/max, difference curve, argrelextrema with clipped ends, the
// threshold from np.diff(xn).mean() -- numpy's pairwise sum for n < 128 -- and the scan).
// One thread, ~100 points: this is control flow, moved here so that the month step has no
// host round trip between the sweep and the update.
namespace oisat {
struct FactorList {
  double f[kMaxFactors];
};

__device__ double np_mean_small(const double* v, int n) {   // np.mean of n <= 128 values
  double s;
  if (n < 8) {
    s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i];
  } else {
    double q[8];
    for (int j = 0; j < 8; ++j) q[j] = v[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) q[j] += v[i + j];
    s = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
    for (; i < n; ++i) s += v[i];
  }
  return s / (double)n;
}

__global__ void oi_knee_kernel(FactorList X, int n, const double* __restrict__ sums,
                               const double* __restrict__ counts, int32_t* __restrict__ pick,
                               double* __restrict__ factor, double* __restrict__ means) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double y[kMaxFactors], xn[kMaxFactors], yd[kMaxFactors], d[kMaxFactors];
  bool any_nan = false;
  for (int i = 0; i < n; ++i) {
    y[i] = sums[i] / counts[i];            // np.nanmean = sum of finite / count of finite
    if (means) means[i] = y[i];
    any_nan = any_nan || y[i] != y[i];
  }
  int best = 0;                            // "no knee": the reference falls back to index 0
  if (!any_nan && n >= 2) {                // a NaN poisons min/max, every comparison is false
    double xmin = X.f[0], xmax = X.f[0], ymin = y[0], ymax = y[0];
    for (int i = 1; i < n; ++i) {
      xmin = X.f[i] < xmin ? X.f[i] : xmin;
      xmax = X.f[i] > xmax ? X.f[i] : xmax;
      ymin = y[i] < ymin ? y[i] : ymin;
      ymax = y[i] > ymax ? y[i] : ymax;
    }
    const double xr = xmax - xmin, yr = ymax - ymin;
    for (int i = 0; i < n; ++i) {
      xn[i] = (X.f[i] - xmin) / xr;
      yd[i] = (y[i] - ymin) / yr - xn[i];
    }
    for (int i = 0; i + 1 < n; ++i) d[i] = xn[i + 1] - xn[i];
    const double drop = fabs(np_mean_small(d, n - 1));   // S = 1
    auto is_max = [&](int i) {
      const double l = yd[i > 0 ? i - 1 : 0], r = yd[i + 1 < n ? i + 1 : n - 1];
      return yd[i] >= l && yd[i] >= r;
    };
    auto is_min = [&](int i) {
      const double l = yd[i > 0 ? i - 1 : 0], r = yd[i + 1 < n ? i + 1 : n - 1];
      return yd[i] <= l && yd[i] <= r;
    };
    int first = -1;
    for (int i = 0; i < n && first < 0; ++i)
      if (is_max(i)) first = i;
    if (first >= 0) {
      double threshold = 0.0;
      int threshold_index = 0;
      for (int i = first; i < n; ++i) {
        if (xn[i] == 1.0) break;
        if (is_max(i)) {
          threshold = yd[i] - drop;
          threshold_index = i;
        }
        if (is_min(i)) threshold = 0.0;
        if (i + 1 < n && yd[i + 1] < threshold) {
          // np.argwhere(x == knee)[0]: the first factor equal to the knee abscissa
          best = threshold_index;
          for (int k = 0; k < n; ++k)
            if (X.f[k] == X.f[threshold_index]) { best = k; break; }
          break;
        }
      }
    }
  }
  *pick = best;
  *factor = X.f[best];
}

__global__ void __launch_bounds__(256)
oi_apply_dev_kernel(const double* __restrict__ xa, const double* __restrict__ y,
                    const double* __restrict__ Sa, const double* __restrict__ So, int64_t n,
                    const double* __restrict__ factor, double* __restrict__ xb,
                    double* __restrict__ ak, double* __restrict__ inc, double* __restrict__ err) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const double r = *factor;
  double K, Sb, AK;
  oi_terms(Sa[c], So[c], r, &K, &Sb, &AK);
  const double d = __dmul_rn(K, __dsub_rn(y[c], xa[c]));  // optimal_interpolation.py:49
  inc[c] = d;
  xb[c] = __dadd_rn(xa[c], d);
  ak[c] = AK;
  err[c] = sqrt(Sb);
}
}  // namespace oisat

extern "C" int oisat_oi_knee(const double* h_factors, int32_t n_factors, const double* sums,
                             const double* counts, int32_t* pick, double* factor, double* means,
                             void* stream) {
  OISAT_CHECK_ARG(h_factors && sums && counts && pick && factor, "null pointer");
  OISAT_CHECK_ARG(n_factors >= 1 && n_factors <= kMaxFactors, "bad extent");
  FactorList X;
  for (int i = 0; i < kMaxFactors; ++i) X.f[i] = i < n_factors ? h_factors[i] : 0.0;
  oi_knee_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(X, n_factors, sums, counts, pick, factor, means);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_oi_apply_dev(const double* xa, const double* y, const double* Sa,
                                  const double* So, int64_t n, const double* factor, double* xb,
                                  double* ak, double* inc, double* err, void* stream) {
  if (n == 0) return OISAT_OK;
  OISAT_CHECK_ARG(xa && y && Sa && So && factor && xb && ak && inc && err && n > 0, "null pointer");
  oi_apply_dev_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      xa, y, Sa, So, n, factor, xb, ak, inc, err);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_oi_apply(const double* xa, const double* y, const double* Sa,
                              const double* So, int64_t n, double factor, double* xb, double* ak,
                              double* inc, double* err, void* stream) {
  if (n == 0) return OISAT_OK;
  OISAT_CHECK_ARG(xa && y && Sa && So && xb && ak && inc && err && n > 0, "null pointer");
  oi_apply_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      xa, y, Sa, So, n, factor, xb, ak, inc, err);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
