// Fused month pipeline for float16 `satellite_amf` products (OMI NO2 / HCHO,
// TROPOMI NO2 -- the BASELINE configurations):
//
//   oisat_pack_granule  reader layout ([level][pixel] float16, level-major) ->
//                       one pixel-major record per pixel, so that gathering a
//                       stencil vertex is ONE coalesced 128-bit load per lane
//                       instead of ~2L scattered 2-byte loads (which would be
//                       bound by L1 wavefronts, not HBM: DESIGN.md section 4).
//   oisat_fused_amf     per (granule, model cell) pair: gather-interpolate all
//                       2L+2(+1) gridded quantities in float64 (interpolator.py:
//                       162-209 through the geometry plan), read the model column
//                       of the matched slot from a shared-memory tile, evaluate
//                       amf_recal.py:93-119,175-183 and stage the five values the
//                       temporal mean needs.  The 123 MB/granule of gridded
//                       intermediates the reference keeps in RAM never exist.
//
// Record layout: R = 8*nchunk halfs, nchunk = ceil(nrow/8), nrow = 2L+2(+1);
// rows = [SW_0..SW_{L-1}, p_0..p_{L-1}, vcd, sigma^2 (squared in float16,
// interpolator.py:186), tropopause?].  Row r sits in chunk r % nchunk at element
// r / nchunk: lane q of a warp loads chunk q (16 bytes) and owns rows
// q, q+nchunk, ..., so the later shared-memory transpose is conflict-free.
#include "vertical.cuh"

namespace oisat {

constexpr int kFusedWarps = 4;
constexpr int kTileCells = 32;

__host__ __device__ inline int record_rows(int L, int has_trop) { return 2 * L + 2 + (has_trop ? 1 : 0); }
__host__ __device__ inline int record_chunks(int L, int has_trop) { return (record_rows(L, has_trop) + 7) / 8; }

// ------------------------------------------------------------------ pack ----
constexpr int kPackPixels = 64;

__global__ void __launch_bounds__(256)
pack_kernel(const __half* __restrict__ sw, const __half* __restrict__ pmid, int L,
            const __half* __restrict__ vcd, const __half* __restrict__ sigma,
            const __half* __restrict__ trop, int64_t n_px, __half* __restrict__ records) {
  extern __shared__ __half tile[];  // [kPackPixels][R + 2]
  const int has_trop = trop != nullptr;
  const int nrow = record_rows(L, has_trop);
  const int nchunk = record_chunks(L, has_trop);
  const int R = 8 * nchunk, pitch = R + 2;
  const int64_t p0 = (int64_t)blockIdx.x * kPackPixels;
  const int px = threadIdx.x & (kPackPixels - 1);
  const int64_t p = p0 + px;
  for (int r = threadIdx.x / kPackPixels; r < R; r += blockDim.x / kPackPixels) {
    // r enumerates record slots; slot -> source row
    const int q = r >> 3, e = r & 7;
    const int row = q + nchunk * e;
    __half v = __float2half_rn(0.0f);
    if (p < n_px && row < nrow) {
      if (row < L) v = sw[(int64_t)row * n_px + p];
      else if (row < 2 * L) v = pmid[(int64_t)(row - L) * n_px + p];
      else if (row == 2 * L) v = vcd[p];
      else if (row == 2 * L + 1) {
        const float s = __half2float(sigma[p]);
        v = __float2half_rn(__fmul_rn(s, s));  // numpy float16 square
      } else v = trop[p];
    }
    tile[px * pitch + r] = v;
  }
  __syncthreads();
  // contiguous write-out, 4 bytes per thread step (pitch keeps rows 4-byte aligned)
  const int words_per_rec = R / 2;
  const int64_t n_here = (n_px - p0) < kPackPixels ? (n_px - p0) : kPackPixels;
  uint32_t* dst = reinterpret_cast<uint32_t*>(records + p0 * R);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(tile);
  for (int64_t i = threadIdx.x; i < n_here * words_per_rec; i += blockDim.x) {
    const int rec = (int)(i / words_per_rec), wd = (int)(i % words_per_rec);
    dst[i] = src[rec * (pitch / 2) + wd];
  }
}

// ----------------------------------------------------------------- fused ----
struct FusedParams {
  oisat_fused_args a;
  int nrow, nchunk, R;
  double box, box_err;
};

__device__ __forceinline__ void half8_to_double(const uint4& u, double* z) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    z[2 * i] = (double)f.x;
    z[2 * i + 1] = (double)f.y;
  }
}

__global__ void __launch_bounds__(kFusedWarps * 32)
fused_amf_kernel(const __grid_constant__ FusedParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const oisat_fused_args& A = P.a;
  const int n_ctm = A.n_ctm_lev;
  const int cpitch = n_ctm + 1;
  float* ctm_s = reinterpret_cast<float*>(smem_raw);                    // [3][32][cpitch]
  WarpScratch* scratch = reinterpret_cast<WarpScratch*>(
      smem_raw + ((3 * kTileCells * cpitch * sizeof(float) + 15) / 16) * 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tile = blockIdx.x;
  const int g = A.tile_granule[tile];
  const int cell0 = A.tile_cell0[tile];
  const int64_t pair0 = A.tile_pair0[tile];
  const uint32_t mask = A.tile_mask[tile];

  // ---- stage the model tile: 32 consecutive cells x n_ctm levels x 3 fields.
  // Global reads are coalesced along cells; the transposed, padded shared layout
  // [cell][level] makes the later per-cell column reads conflict-free.
  {
    const int64_t slot_off = (int64_t)A.gran_slot[g] * n_ctm * A.n_cell;
    const int c = lane;
    const bool in_row = (int64_t)cell0 + c < A.n_cell;
    for (int k = warp; k < n_ctm; k += kFusedWarps) {
      const int64_t src = slot_off + (int64_t)k * A.n_cell + cell0 + c;
      float pm = 0.f, pr = 0.f, dp = 0.f;
      if (in_row) { pm = A.ctm_pmid[src]; pr = A.ctm_prof[src]; dp = A.ctm_dp[src]; }
      ctm_s[(0 * kTileCells + c) * cpitch + k] = pm;
      ctm_s[(1 * kTileCells + c) * cpitch + k] = pr;
      ctm_s[(2 * kTileCells + c) * cpitch + k] = dp;
    }
  }
  __syncthreads();

  WarpScratch& s = scratch[warp];
  double* rows = s.xs;  // xs and ys are adjacent: 2*128 doubles >= nrow
  const int L = A.n_sat_lev;
  const int S = 3 * A.nwin;
  const int64_t rec0 = A.gran_record0[g];
  const int64_t px0 = A.gran_px0[g];
  const uint4* records = reinterpret_cast<const uint4*>(A.records);

  // set bits of the tile mask are dealt round-robin to the warps
  int ord = 0;
  for (uint32_t m = mask; m; m &= m - 1, ++ord) {
    if ((ord % kFusedWarps) != warp) continue;
    const int l = __ffs(m) - 1;  // cell offset inside the tile
    const int64_t pair = pair0 + ord;
    double acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0;
    double acc_amf = 0.0;
    bool alive = true;
    for (int base = 0; base < S; base += 30) {  // 30 = 10 window nodes per sweep
      const int nk = (S - base) < 30 ? (S - base) : 30;
      int32_t v = 0;
      double wt = 0.0;
      if (lane < nk) {
        v = A.vert[pair * S + base + lane];
        wt = A.w[pair * S + base + lane];
        alive = alive && (A.good[px0 + v] != 0);
      }
      double fine[8], fine_amf = 0.0;
      for (int k = 0; k < nk; ++k) {
        const int32_t vk = __shfl_sync(0xffffffffu, v, k);
        const double wk = __shfl_sync(0xffffffffu, wt, k);
        const int j = k % 3;
        if (lane < P.nchunk) {
          const uint4 u = __ldg(&records[(rec0 + vk) * P.nchunk + lane]);
          double z[8];
          half8_to_double(u, z);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const double prod = __dmul_rn(wk, z[e]);
            fine[e] = j == 0 ? __dadd_rn(0.0, prod) : __dadd_rn(fine[e], prod);
          }
        } else if (lane == P.nchunk) {
          const double z = load_as_double(A.amf, A.amf_dtype, px0 + vk);
          const double prod = __dmul_rn(wk, z);
          fine_amf = j == 0 ? __dadd_rn(0.0, prod) : __dadd_rn(fine_amf, prod);
        }
        if (j == 2) {
          if (lane < P.nchunk) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int row = lane + P.nchunk * e;
              const double bw = row == 2 * L + 1 ? P.box_err : P.box;
              acc[e] = __dadd_rn(acc[e], __dmul_rn(fine[e], bw));
            }
          } else if (lane == P.nchunk) {
            acc_amf = __dadd_rn(acc_amf, __dmul_rn(fine_amf, P.box));
          }
        }
      }
    }
    alive = __all_sync(0xffffffffu, alive);
    if (!alive) {  // a masked vertex poisons every field of the cell (interpolator.py:126-128)
      if (lane < 5) A.staged[(int64_t)lane * A.n_pairs + pair] = qnan();
      continue;
    }
    if (lane < P.nchunk) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int row = lane + P.nchunk * e;
        if (row < P.nrow) rows[row] = acc[e];
      }
    }
    const double old_amf = __shfl_sync(0xffffffffu, acc_amf, P.nchunk);
    __syncwarp();
    const double vcd = rows[2 * L];
    const double sigma = sqrt(rows[2 * L + 1]);  // interpolator.py:188
    const bool has_trop = A.has_trop != 0;
    const double trop = has_trop ? rows[2 * L + 2] : 0.0;
    for (int i = lane; i < L; i += 32) {
      s.xr[i] = log(rows[L + i]);
      s.yr[i] = rows[i];
    }
    __syncwarp();
    double new_amf = qnan(), col = qnan(), vnew = qnan();
    if (vcd == vcd) {  // amf_recal.py:99-100
      const float* cp = ctm_s + (0 * kTileCells + l) * cpitch;
      const float* cx = ctm_s + (1 * kTileCells + l) * cpitch;
      const float* cd = ctm_s + (2 * kTileCells + l) * cpitch;
      double colsum;
      new_amf = warp_amf_cell<true>(
          s, L, n_ctm, has_trop, trop, [&](int k) { return (double)cp[k]; },
          [&](int k) { return (double)partial_column_f32(cd[k], cx[k]); }, &colsum, lane);
      vnew = (old_amf * vcd) / new_amf;                       // amf_recal.py:179
      col = (vnew != vnew || isinf(vnew)) ? qnan() : colsum;  // :180-181
    }
    __syncwarp();
    if (lane == 0) {
      A.staged[0 * A.n_pairs + pair] = vnew;
      A.staged[1 * A.n_pairs + pair] = sigma;
      A.staged[2 * A.n_pairs + pair] = col;
      A.staged[3 * A.n_pairs + pair] = new_amf;
      A.staged[4 * A.n_pairs + pair] = old_amf;
    }
  }
}

static size_t fused_smem_bytes(int n_ctm) {
  const size_t ctm = ((size_t)3 * kTileCells * (n_ctm + 1) * sizeof(float) + 15) / 16 * 16;
  return ctm + kFusedWarps * sizeof(WarpScratch);
}

}  // namespace oisat

using namespace oisat;

extern "C" int64_t oisat_pack_record_halfs(int32_t n_sat_lev, int32_t has_trop) {
  return 8 * (int64_t)record_chunks(n_sat_lev, has_trop);
}

extern "C" int oisat_pack_granule(const void* sw, const void* p_mid, int32_t n_sat_lev,
                                  const void* vcd, const void* sigma, const void* trop,
                                  int64_t n_px, void* records, void* stream) {
  OISAT_CHECK_ARG(sw && p_mid && vcd && sigma && records, "null pointer");
  OISAT_CHECK_ARG(n_sat_lev >= 2 && n_sat_lev <= kMaxSatLev, "bad level count");
  if (n_px <= 0) return OISAT_OK;
  const int R = 8 * record_chunks(n_sat_lev, trop != nullptr);
  const size_t smem = (size_t)kPackPixels * (R + 2) * sizeof(__half);
  pack_kernel<<<(unsigned)ceil_div(n_px, kPackPixels), 256, smem, (cudaStream_t)stream>>>(
      (const __half*)sw, (const __half*)p_mid, n_sat_lev, (const __half*)vcd,
      (const __half*)sigma, (const __half*)trop, n_px, (__half*)records);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_fused_amf(const oisat_fused_args* h_args, void* stream) {
  OISAT_CHECK_ARG(h_args != nullptr, "null args");
  const oisat_fused_args& a = *h_args;
  if (a.n_tiles == 0 || a.n_pairs == 0) return OISAT_OK;
  OISAT_CHECK_ARG(a.tile_granule && a.tile_cell0 && a.tile_pair0 && a.tile_mask && a.vert && a.w &&
                      a.gran_record0 && a.gran_px0 && a.gran_slot && a.records && a.good &&
                      a.amf && a.ctm_pmid && a.ctm_prof && a.ctm_dp && a.staged,
                  "null pointer");
  OISAT_CHECK_ARG(a.nwin >= 1 && a.n_sat_lev >= 2 && a.n_sat_lev <= kMaxSatLev, "bad stencil");
  OISAT_CHECK_ARG(a.n_ctm_lev >= 2 && a.n_ctm_lev <= kMaxCtmLev, "bad model level count");
  OISAT_CHECK_ARG(a.amf_dtype == OISAT_F16 || a.amf_dtype == OISAT_F32 || a.amf_dtype == OISAT_F64,
                  "bad amf dtype");
  FusedParams P;
  P.a = a;
  P.nrow = record_rows(a.n_sat_lev, a.has_trop);
  P.nchunk = record_chunks(a.n_sat_lev, a.has_trop);
  P.R = 8 * P.nchunk;
  OISAT_CHECK_ARG(P.nchunk < 32, "record too wide");
  P.box = a.box_weight;
  P.box_err = a.box_weight_err;
  const size_t smem = fused_smem_bytes(a.n_ctm_lev);
  static size_t configured = 0;
  if (smem > configured) {
    OISAT_CHECK_CUDA(cudaFuncSetAttribute(fused_amf_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  fused_amf_kernel<<<(unsigned)a.n_tiles, kFusedWarps * 32, smem, (cudaStream_t)stream>>>(P);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
