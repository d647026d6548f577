// Which (granule, cell) pairs of a month can hold a value at all, decided before the fused
// gather runs.
//
// The quality mask multiplies EVERY field of a pixel by NaN (interpolator.py:126-128) and
// LinearNDInterpolator + the box filter propagate a NaN vertex whatever its weight (SURVEY
// A.2): a pair with one masked pixel among its 3*nwin stencil vertices is NaN in all five
// staged quantities, without any arithmetic.  Real OMI months flag more than half of their
// pixels (clouds, row anomaly); the synthetic BASELINE month flags 20 % of them, which kills
// 25 % of the pairs -- a quarter of the gather-interpolate + AMF work whose result is known
// beforehand.
//
//   oisat_pair_alive    one thread per pair: reads the stencil and the per-pixel mask bytes the
//                       pack step wrote; dead pairs get their five NaNs here; for live pairs the
//                       gridded old AMF (staged row 4) is evaluated here as well -- the same
//                       products and the same 16-leaf summation tree per sweep of 15 entries as
//                       the half-warp butterfly of the gather kernels, so the bits are theirs.
//                       Live pairs are appended to `alive_pairs`, one reservation per block of
//                       256 pairs (order inside a block kept: neighbours stay neighbours, which
//                       is what the L2 reuse of the gather lives on).
//
// The tile kernel then runs over the compact list only (oisat_fused_args::alive_pairs).
#include "common.cuh"

namespace oisat {

constexpr int kAliveThreads = 256;

__global__ void __launch_bounds__(kAliveThreads)
pair_alive_kernel(int64_t n_pairs, int S, int vec, const int32_t* __restrict__ vert,
                  const double* __restrict__ w, const int64_t* __restrict__ pair_record0,
                  const int32_t* __restrict__ pair_granule, const int64_t* __restrict__ gran_px0,
                  const uint8_t* __restrict__ px_bad, const double* __restrict__ amf_masked,
                  double box_weight, double* __restrict__ staged,
                  int32_t* __restrict__ alive_pairs, unsigned long long* __restrict__ n_alive) {
  const int64_t pair = (int64_t)blockIdx.x * kAliveThreads + threadIdx.x;
  const bool mine = pair < n_pairs;
  bool alive = mine;
  double old_amf = 0.0;
  if (mine && vec) {   // S == 12, 16-byte aligned tables
    // the 2 x 2 window of the OMI products: the whole stencil of a pair is three 16-byte loads
    // of vertices and six of weights, all issued before the first dependent mask byte is needed
    const int64_t px0 = pair_record0 ? pair_record0[pair] : gran_px0[pair_granule[pair]];
    const int4* v4 = reinterpret_cast<const int4*>(vert + pair * 12);
    const int4 va = v4[0], vb = v4[1], vc = v4[2];
    const int32_t v[12] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y, vc.z, vc.w};
    uint8_t bad[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) bad[e] = px_bad[px0 + v[e]];
#pragma unroll
    for (int e = 0; e < 12; ++e) alive = alive && (bad[e] == 0);
    if (alive) {
      const double2* w2 = reinterpret_cast<const double2*>(w + pair * 12);
      double wt[12], am[12];
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        const double2 t = w2[e];
        wt[2 * e] = t.x;
        wt[2 * e + 1] = t.y;
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) am[e] = amf_masked[px0 + v[e]];
      double z[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) z[e] = e < 12 ? wt[e] * am[e] : 0.0;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < o; ++i) z[i] = z[i] + z[i + o];
      old_amf = z[0];
      staged[4 * n_pairs + pair] = old_amf * box_weight;
    } else {
#pragma unroll
      for (int q = 0; q < 5; ++q) staged[q * n_pairs + pair] = qnan();
    }
  } else if (mine) {
    const int64_t px0 = pair_record0 ? pair_record0[pair] : gran_px0[pair_granule[pair]];
    const int32_t* v = vert + pair * S;
    const double* wt = w + pair * S;
    for (int e = 0; e < S; ++e) alive = alive && (px_bad[px0 + v[e]] == 0);
    if (alive) {
      for (int base = 0; base < S; base += 15) {
        const int nk = (S - base) < 15 ? (S - base) : 15;
        double z[16];
#pragma unroll
        for (int e = 0; e < 16; ++e)
          z[e] = e < nk ? wt[base + e] * amf_masked[px0 + v[base + e]] : 0.0;
        // lane 0 of: for (o = 8; o > 0; o >>= 1) za += shfl_xor(za, o, 16)
#pragma unroll
        for (int o = 8; o > 0; o >>= 1)
#pragma unroll
          for (int i = 0; i < o; ++i) z[i] = z[i] + z[i + o];
        old_amf += z[0];
      }
      staged[4 * n_pairs + pair] = old_amf * box_weight;
    } else {
#pragma unroll
      for (int q = 0; q < 5; ++q) staged[q * n_pairs + pair] = qnan();
    }
  }
  // ordered append of this block's live pairs
  __shared__ int warp_count[kAliveThreads / 32];
  __shared__ unsigned long long block_base;
  const unsigned ballot = __ballot_sync(0xffffffffu, alive);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) warp_count[wid] = __popc(ballot);
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int i = 0; i < kAliveThreads / 32; ++i) {
      const int c = warp_count[i];
      warp_count[i] = tot;
      tot += c;
    }
    block_base = tot ? atomicAdd(n_alive, (unsigned long long)tot) : 0ull;
  }
  __syncthreads();
  if (alive) {
    const int rank = warp_count[wid] + __popc(ballot & ((1u << lane) - 1u));
    alive_pairs[block_base + rank] = (int32_t)pair;
  }
}

// Wide stencils (TROPOMI: 90 entries): a half warp per pair, lane = entry of a sweep of 15 --
// the gather kernels' own layout, butterfly included -- so that the 90 dependent gathers of a
// pair are six rounds of 15 parallel ones instead of 90 in a row in one thread.
__global__ void __launch_bounds__(kAliveThreads)
pair_alive_hw_kernel(int64_t n_pairs, int S, const int32_t* __restrict__ vert,
                     const double* __restrict__ w, const int64_t* __restrict__ pair_record0,
                     const int32_t* __restrict__ pair_granule, const int64_t* __restrict__ gran_px0,
                     const uint8_t* __restrict__ px_bad, const double* __restrict__ amf_masked,
                     double box_weight, double* __restrict__ staged,
                     int32_t* __restrict__ alive_pairs, unsigned long long* __restrict__ n_alive) {
  const int gl = threadIdx.x & 15, slot = threadIdx.x >> 4;
  const int64_t pair_raw = (int64_t)blockIdx.x * (kAliveThreads / 16) + slot;
  const bool mine = pair_raw < n_pairs;
  const int64_t pair = mine ? pair_raw : n_pairs - 1;
  const int64_t px0 = pair_record0 ? pair_record0[pair] : gran_px0[pair_granule[pair]];
  const unsigned half_mask = 0xffffu << (threadIdx.x & 16);
  bool any_bad = false;
  double acc_amf = 0.0;
  for (int base = 0; base < S; base += 15) {
    const int nk = (S - base) < 15 ? (S - base) : 15;
    double za = 0.0;
    bool bad = false;
    if (gl < nk) {
      const int32_t v = vert[pair * S + base + gl];
      bad = px_bad[px0 + v] != 0;
      za = w[pair * S + base + gl] * amf_masked[px0 + v];
    }
    // (the ballot must be executed by every lane in every sweep: no short-circuit around it)
    const unsigned bad_lanes = __ballot_sync(0xffffffffu, bad);
    any_bad = any_bad || (bad_lanes & half_mask) != 0;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) za += __shfl_xor_sync(0xffffffffu, za, o, 16);
    acc_amf += za;
  }
  const bool alive = mine && !any_bad;
  if (mine && gl == 0) {
    if (alive) {
      staged[4 * n_pairs + pair] = acc_amf * box_weight;
    } else {
#pragma unroll
      for (int q = 0; q < 5; ++q) staged[q * n_pairs + pair] = qnan();
    }
  }
  __shared__ int flag[kAliveThreads / 16];
  __shared__ unsigned long long block_base;
  if (gl == 0) flag[slot] = alive ? 1 : 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int i = 0; i < kAliveThreads / 16; ++i) {
      const int c = flag[i];
      flag[i] = tot;
      tot += c;
    }
    block_base = tot ? atomicAdd(n_alive, (unsigned long long)tot) : 0ull;
  }
  __syncthreads();
  if (alive && gl == 0) alive_pairs[block_base + flag[slot]] = (int32_t)pair;
}

}  // namespace oisat

using namespace oisat;

extern "C" int oisat_pair_alive(int64_t n_pairs, int32_t nwin, const int32_t* vert, const double* w,
                                const int64_t* pair_record0, const int32_t* pair_granule,
                                const int64_t* gran_px0, const uint8_t* px_bad,
                                const double* amf_masked, double box_weight, double* staged,
                                int32_t* alive_pairs, int64_t* n_alive, void* stream) {
  if (n_pairs <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(vert && w && px_bad && amf_masked && staged && alive_pairs && n_alive,
                  "null pointer");
  OISAT_CHECK_ARG(pair_record0 || (pair_granule && gran_px0), "no pixel base of the pairs");
  OISAT_CHECK_ARG(nwin >= 1 && n_pairs < ((int64_t)1 << 31), "bad stencil / too many pairs");
  cudaStream_t s = (cudaStream_t)stream;
  OISAT_CHECK_CUDA(cudaMemsetAsync(n_alive, 0, sizeof(int64_t), s));
  if (3 * nwin > 15) {
    pair_alive_hw_kernel<<<(unsigned)ceil_div(n_pairs, kAliveThreads / 16), kAliveThreads, 0, s>>>(
        n_pairs, 3 * nwin, vert, w, pair_record0, pair_granule, gran_px0, px_bad, amf_masked,
        box_weight, staged, alive_pairs, reinterpret_cast<unsigned long long*>(n_alive));
    OISAT_CHECK_LAUNCH();
    return OISAT_OK;
  }
  pair_alive_kernel<<<(unsigned)ceil_div(n_pairs, kAliveThreads), kAliveThreads, 0, s>>>(
      n_pairs, 3 * nwin,
      (3 * nwin == 12 && ((reinterpret_cast<uintptr_t>(vert) | reinterpret_cast<uintptr_t>(w)) & 15) == 0) ? 1 : 0,
      vert, w, pair_record0, pair_granule, gran_px0, px_bad, amf_masked,
      box_weight, staged, alive_pairs, reinterpret_cast<unsigned long long*>(n_alive));
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
