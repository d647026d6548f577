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
                     double* __restrict__ leaf_cnt) {
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
  for (int f = 0; f < fac.n; ++f) {
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
      leaf_sum[(int64_t)f * n_leaf + g] = acc;
      leaf_cnt[(int64_t)f * n_leaf + g] = cnt;
    }
  }
}

// numpy's recursive halving over the leaves (depth <= ~40 for any int64 n)
__device__ double combine_leaves(const double* leaf, int64_t n, int64_t* next) {
  if (n <= kLeafMax) return leaf[(*next)++];
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  const double a = combine_leaves(leaf, n2, next);
  const double b = combine_leaves(leaf, n - n2, next);
  return __dadd_rn(a, b);
}

__global__ void oi_sweep_combine_kernel(const double* __restrict__ leaf_sum,
                                        const double* __restrict__ leaf_cnt, int64_t n_leaf,
                                        int64_t n, int n_factors, double* __restrict__ sums,
                                        double* __restrict__ counts) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_factors) return;
  int64_t next = 0;
  sums[f] = combine_leaves(leaf_sum + (int64_t)f * n_leaf, n, &next);
  double c = 0.0;
  for (int64_t i = 0; i < n_leaf; ++i) c += leaf_cnt[(int64_t)f * n_leaf + i];
  counts[f] = c;
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
  return (nl + 1) * (int64_t)sizeof(int64_t) + 2 * nl * n_factors * (int64_t)sizeof(double);
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
  int64_t* d_start = (int64_t*)work;
  double* d_sum = (double*)(d_start + nl + 1);
  double* d_cnt = d_sum + nl * n_factors;
  // pageable source: the runtime stages it before returning, `starts` may die
  OISAT_CHECK_CUDA(cudaMemcpyAsync(d_start, starts.data(), (nl + 1) * sizeof(int64_t),
                                   cudaMemcpyHostToDevice, s));
  Factors fac;
  fac.n = n_factors;
  for (int i = 0; i < n_factors; ++i) fac.r[i] = h_factors[i];
  oi_sweep_leaf_kernel<<<(unsigned)ceil_div(nl * 8, 256), 256, 0, s>>>(Sa, So, d_start, nl, fac,
                                                                      d_sum, d_cnt);
  OISAT_CHECK_LAUNCH();
  oi_sweep_combine_kernel<<<(unsigned)ceil_div(n_factors, 32), 32, 0, s>>>(d_sum, d_cnt, nl, n,
                                                                          n_factors, sums, counts);
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
