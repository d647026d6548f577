// Fused month pipeline for float16 `satellite_amf` products (OMI NO2 / HCHO,
// TROPOMI NO2 -- the BASELINE configurations):
//
//   oisat_ctm_prepare   once per model slot: float32 log-pressure and partial
//                       column (amf_recal.py:51-56,108) so the per-pair work
//                       reads two derived fields instead of recomputing them
//                       for every granule like the reference does (:151-152).
//   oisat_pack_granule / oisat_pack_batch
//                       reader layout ([level][pixel] float16, level-major) ->
//                       one pixel-major record per pixel, so that gathering a
//                       stencil vertex is ONE coalesced 128-bit load per lane
//                       instead of ~2L scattered 2-byte loads (which would be
//                       bound by L1 wavefronts, not HBM: DESIGN.md section 4);
//                       the batch form also folds the quality mask
//                       (interpolator.py:126-128) and runs a whole month in one launch.
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
// r / nchunk: lane q of a group loads chunk q (16 bytes) and owns rows
// q, q+nchunk, ..., so the later shared-memory transpose is conflict-free.
#include "vertical.cuh"

namespace oisat {

constexpr int kTileCells = 32;
constexpr int kFusedThreads = 256;

__host__ __device__ inline int record_rows(int L, int has_trop) { return 2 * L + 2 + (has_trop ? 1 : 0); }
__host__ __device__ inline int record_chunks(int L, int has_trop) { return (record_rows(L, has_trop) + 7) / 8; }

// ----------------------------------------------------------------- prepare ----
__global__ void __launch_bounds__(256)
ctm_prepare_kernel(const float* __restrict__ pmid, const float* __restrict__ prof,
                   const float* __restrict__ dp, int64_t n, float* __restrict__ logp,
                   float* __restrict__ pcol) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 pm = *reinterpret_cast<const float4*>(pmid + i);
    const float4 pr = *reinterpret_cast<const float4*>(prof + i);
    const float4 d = *reinterpret_cast<const float4*>(dp + i);
    float4 lp, pc;
    lp.x = log_f32(pm.x); lp.y = log_f32(pm.y); lp.z = log_f32(pm.z); lp.w = log_f32(pm.w);
    pc.x = partial_column_f32(d.x, pr.x); pc.y = partial_column_f32(d.y, pr.y);
    pc.z = partial_column_f32(d.z, pr.z); pc.w = partial_column_f32(d.w, pr.w);
    *reinterpret_cast<float4*>(logp + i) = lp;
    *reinterpret_cast<float4*>(pcol + i) = pc;
  } else {
    for (int64_t j = i; j < n; ++j) {
      logp[j] = log_f32(pmid[j]);
      pcol[j] = partial_column_f32(dp[j], prof[j]);
    }
  }
}

// -------------------------------------------------------------------- pack ----
constexpr int kPackPixels = 256;

struct PackSrc {
  const __half* sw;
  const __half* pmid;
  const __half* vcd;
  const __half* sigma;
  const __half* trop;
  int64_t n_px;
};

__device__ __forceinline__ int record_slot(int row, int nchunk) {
  return 8 * (row % nchunk) + row / nchunk;
}

// One block transposes kPackPixels pixels x nrow rows through shared memory.
// Load phase: a thread moves two neighbouring pixels of one source row per
// 32-bit load (rows are contiguous along pixels in the reader layout), so a warp
// reads 128 contiguous bytes; the odd word pitch of the tile makes the
// transposing stores conflict-free.  Store phase: 16 lanes per record, one
// 16-byte chunk each, i.e. fully coalesced 128-bit stores.
__device__ __forceinline__ void pack_block(const PackSrc& s, int L, int has_trop, int64_t p0,
                                           __half* __restrict__ records, __half* tile,
                                           const unsigned char* bad) {
  const int nrow = record_rows(L, has_trop);
  const int nchunk = record_chunks(L, has_trop);
  const int R = 8 * nchunk, pitch = R + 2;
  const int64_t left = s.n_px - p0;
  const int n_here = left < kPackPixels ? (int)left : kPackPixels;
  const bool pairs_ok = ((s.n_px & 1) == 0) && (n_here == kPackPixels);
  const __half zero = __float2half_rn(0.0f);
  __shared__ unsigned char slot_tab[2 * kMaxSatLev + 16];  // row -> slot, no divisions in the loop
  for (int row = threadIdx.x; row < nrow; row += blockDim.x)
    slot_tab[row] = (unsigned char)record_slot(row, nchunk);
  __syncthreads();
  // padding slots of the record (rows >= nrow)
  for (int i = threadIdx.x; i < kPackPixels * (R - nrow); i += blockDim.x) {
    const int px = i / (R - nrow), row = nrow + i % (R - nrow);
    tile[px * pitch + record_slot(row, nchunk)] = zero;
  }
  if (pairs_ok) {
    const int pp = threadIdx.x & (kPackPixels / 2 - 1);   // pixel pair
    const int rl = threadIdx.x / (kPackPixels / 2);       // row lane
    const int rstep = blockDim.x / (kPackPixels / 2);
    __half* t0 = tile + (2 * pp) * pitch;
    __half* t1 = t0 + pitch;
    const int64_t col = p0 + 2 * pp;
    // all loads of a pass are issued before the first transposing store, so a
    // thread keeps 2*kIter independent 32-bit loads in flight (pure streaming kernel:
    // bytes in flight per SM are what reaches HBM bandwidth)
    constexpr int kIter = 12;
    const __half* psw = s.sw + (int64_t)rl * s.n_px + col;
    const __half* ppm = s.pmid + (int64_t)rl * s.n_px + col;
    const int64_t hop = (int64_t)rstep * s.n_px;
    for (int row0 = rl; row0 < L; row0 += kIter * rstep) {
      __half2 va[kIter], vb[kIter];
#pragma unroll
      for (int i = 0; i < kIter; ++i) {
        if (row0 + i * rstep < L) {
          va[i] = __ldcs(reinterpret_cast<const __half2*>(psw + i * hop));
          vb[i] = __ldcs(reinterpret_cast<const __half2*>(ppm + i * hop));
        }
      }
#pragma unroll
      for (int i = 0; i < kIter; ++i) {
        const int row = row0 + i * rstep;
        if (row < L) {
          const int sa = slot_tab[row], sb = slot_tab[L + row];
          t0[sa] = __low2half(va[i]); t1[sa] = __high2half(va[i]);
          t0[sb] = __low2half(vb[i]); t1[sb] = __high2half(vb[i]);
        }
      }
      psw += kIter * hop;
      ppm += kIter * hop;
    }
    if (rl == 0) {
      const __half2 v = *reinterpret_cast<const __half2*>(s.vcd + col);
      const int sv = record_slot(2 * L, nchunk);
      t0[sv] = __low2half(v); t1[sv] = __high2half(v);
      const float2 g = __half22float2(*reinterpret_cast<const __half2*>(s.sigma + col));
      const int ss = record_slot(2 * L + 1, nchunk);
      t0[ss] = __float2half_rn(__fmul_rn(g.x, g.x));   // numpy float16 square
      t1[ss] = __float2half_rn(__fmul_rn(g.y, g.y));
      if (has_trop) {
        const __half2 tr = *reinterpret_cast<const __half2*>(s.trop + col);
        const int st = record_slot(2 * L + 2, nchunk);
        t0[st] = __low2half(tr); t1[st] = __high2half(tr);
      }
    }
  } else {  // ragged tail / odd pixel count: one value per thread step
    for (int i = threadIdx.x; i < n_here * nrow; i += blockDim.x) {
      const int row = i / n_here, px = i - row * n_here;
      const int64_t p = p0 + px;
      __half v;
      if (row < L) v = s.sw[(int64_t)row * s.n_px + p];
      else if (row < 2 * L) v = s.pmid[(int64_t)(row - L) * s.n_px + p];
      else if (row == 2 * L) v = s.vcd[p];
      else if (row == 2 * L + 1) {
        const float g = __half2float(s.sigma[p]);
        v = __float2half_rn(__fmul_rn(g, g));
      } else v = s.trop[p];
      tile[px * pitch + record_slot(row, nchunk)] = v;
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(records + p0 * R);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(tile);
  const int q = threadIdx.x & 15;
  if (q < nchunk) {
    const uint32_t nan2 = 0x7e007e00u;  // two float16 quiet NaNs
    for (int rec = threadIdx.x >> 4; rec < n_here; rec += blockDim.x >> 4) {
      const uint32_t* w = src + rec * (pitch / 2) + 4 * q;
      // a masked pixel is a NaN vertex for EVERY field (interpolator.py:126-128,163)
      dst[rec * nchunk + q] = (bad != nullptr && bad[rec]) ? make_uint4(nan2, nan2, nan2, nan2)
                                                           : make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__global__ void __launch_bounds__(256)
pack_kernel(PackSrc s, int L, int has_trop, __half* __restrict__ records) {
  extern __shared__ __half tile[];
  pack_block(s, L, has_trop, (int64_t)blockIdx.x * kPackPixels, records, tile, nullptr);
}

// Staged row pitch of the bulk-copy path, in halfs: 528 bytes keeps every row
// 16-byte aligned (a bulk-copy requirement) and rotates the banks by 4 words per row.
constexpr int kTmaPitch = kPackPixels + 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// Whole month in one launch.  One block = one tile of kPackPixels pixels of one
// granule.  The reader layout is row-contiguous, so a tile is nrow contiguous
// 512-byte row segments: they are fetched with bulk asynchronous copies
// (cp.async.bulk, completion counted on an mbarrier) issued by one warp -- no
// per-element load/store instructions and no registers tied up while ~50 KB per
// block are in flight; with four blocks resident per SM that is what it takes to
// keep HBM busy (the LDG/STS version above stalls at half the bandwidth: 17
// instructions per element and no loads in flight during its store phase).  The
// transposition then happens on the way out: a thread assembles one 16-byte record
// chunk from eight staged rows and stores it, consecutive threads to consecutive
// chunks.  Granules whose arrays are not 16-byte aligned (or whose pixel count is
// not a multiple of 8) take pack_block.
__global__ void __launch_bounds__(256, 4)
pack_batch_kernel(const oisat_pack_item* __restrict__ items, int n_items, int L, int has_trop,
                  int qflag_dtype, double thresh, int amf_dtype, __half* __restrict__ records,
                  double* __restrict__ amf_masked, int use_bulk,
                  const int32_t* __restrict__ block_item, uint8_t* __restrict__ px_bad,
                  int skip_masked) {
  extern __shared__ __align__(128) __half tile[];
  __shared__ unsigned char bad[kPackPixels];
  __shared__ __align__(8) unsigned long long mbar;
  // granule of this block: from the caller's table, else the last item with block0 <= blockIdx.x
  // (a bisection is ~9 DEPENDENT global loads before the block can request a single byte)
  int lo = 0;
  if (block_item) {
    lo = block_item[blockIdx.x];
  } else {
    int hi = n_items - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (items[mid].block0 <= (int64_t)blockIdx.x) lo = mid; else hi = mid - 1;
    }
  }
  const oisat_pack_item it = items[lo];
  PackSrc s{(const __half*)it.sw, (const __half*)it.p_mid, (const __half*)it.vcd,
            (const __half*)it.sigma, (const __half*)it.trop, it.n_px};
  const int64_t p0 = ((int64_t)blockIdx.x - it.block0) * kPackPixels;
  // quality mask of the tile's pixels + masked AMF (interpolator.py:126-128)
  auto mask_pixels = [&]() {
    if (threadIdx.x < kPackPixels) {
      const int64_t p = p0 + threadIdx.x;
      bool is_bad = true;
      if (p < it.n_px) {
        is_bad = !(load_as_double(it.qflag, qflag_dtype, p) > thresh);
        amf_masked[it.px0 + p] = is_bad ? qnan() : load_as_double(it.amf, amf_dtype, p);
        if (px_bad) px_bad[it.px0 + p] = is_bad ? 1 : 0;
      }
      bad[threadIdx.x] = is_bad ? 1 : 0;
    }
  };
  const int nrow = record_rows(L, has_trop);
  const int nchunk = record_chunks(L, has_trop);
  const int R = 8 * nchunk;
  uintptr_t align = (uintptr_t)it.sw | (uintptr_t)it.p_mid | (uintptr_t)it.vcd | (uintptr_t)it.sigma;
  if (has_trop) align |= (uintptr_t)it.trop;
  if (!use_bulk || (align & 15) != 0 || (it.n_px & 7) != 0) {
    mask_pixels();
    // (pack_block starts with a __syncthreads that also publishes `bad`)
    pack_block(s, L, has_trop, p0, records + it.px0 * R, tile, bad);
    return;
  }
  const int64_t left = it.n_px - p0;
  const int n_here = left < kPackPixels ? (int)left : kPackPixels;
  const uint32_t row_bytes = (uint32_t)n_here * (uint32_t)sizeof(__half);  // multiple of 16
  const uint32_t bar = smem_u32(&mbar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();  // barrier initialised
  // the row copies are requested FIRST; the mask (its own loads and stores) is worked out
  // while they are in flight and published by the barrier before the store phase
  if (threadIdx.x >= 32) mask_pixels();
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                   ::"r"(bar), "r"(row_bytes * (uint32_t)nrow) : "memory");
    __syncwarp();
    for (int row = threadIdx.x; row < nrow; row += 32) {
      const __half* src = row < L          ? s.sw + (int64_t)row * s.n_px
                          : row < 2 * L    ? s.pmid + (int64_t)(row - L) * s.n_px
                          : row == 2 * L   ? s.vcd
                          : row == 2 * L + 1 ? s.sigma
                                             : s.trop;
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
          ::"r"(smem_u32(tile + row * kTmaPitch)), "l"(src + p0), "r"(row_bytes), "r"(bar)
          : "memory");
    }
    mask_pixels();
  }
  // rows nrow..R-1 are the zero padding of the record
  for (int i = threadIdx.x; i < (R - nrow) * kPackPixels; i += blockDim.x)
    tile[(nrow + i / kPackPixels) * kTmaPitch + (i % kPackPixels)] = __float2half_rn(0.0f);
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p;\n\t"
                   "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                   "  selp.b32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar) : "memory");
  }
  if (threadIdx.x < n_here) {  // sigma -> sigma^2, squared in float16 like numpy (interpolator.py:186)
    __half* sg = tile + (2 * L + 1) * kTmaPitch + threadIdx.x;
    const float g = __half2float(*sg);
    *sg = __float2half_rn(__fmul_rn(g, g));
  }
  __syncthreads();
  const unsigned short* t16 = reinterpret_cast<const unsigned short*>(tile);
  uint4* dst = reinterpret_cast<uint4*>(records + (it.px0 + p0) * R);
  const int total = n_here * nchunk;
  const int kstride = nchunk * kTmaPitch;  // rows q, q+nchunk, ... are the 8 elements of chunk q
  const int dpx = (int)blockDim.x / nchunk, dq = (int)blockDim.x % nchunk;
  int px = threadIdx.x / nchunk, q = threadIdx.x % nchunk;
  const uint32_t nan2 = 0x7e007e00u;  // two float16 quiet NaNs
  for (int c = threadIdx.x; c < total; c += blockDim.x) {
    const unsigned short* e = t16 + q * kTmaPitch + px;
    uint4 v;
    v.x = (uint32_t)e[0] | ((uint32_t)e[kstride] << 16);
    v.y = (uint32_t)e[2 * kstride] | ((uint32_t)e[3 * kstride] << 16);
    v.z = (uint32_t)e[4 * kstride] | ((uint32_t)e[5 * kstride] << 16);
    v.w = (uint32_t)e[6 * kstride] | ((uint32_t)e[7 * kstride] << 16);
    // a masked pixel is a NaN vertex for EVERY field (interpolator.py:126-128,163); when the
    // caller runs over the live pairs only (oisat_pair_alive) its record is never read and is
    // not written at all
    if (bad[px]) v = make_uint4(nan2, nan2, nan2, nan2);
    if (!(skip_masked && bad[px])) dst[c] = v;
    px += dpx;
    q += dq;
    if (q >= nchunk) { q -= nchunk; ++px; }
  }
}

// ------------------------------------------------------------------- fused ----
struct FusedParams {
  oisat_fused_args a;
  int nrow, nchunk, nfield;
  int scratch_doubles;  // per group
};

__device__ __forceinline__ void half8_to_double(const uint4& u, double* z) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // float16 -> float64 is exact; one conversion per value, no float32 stop-over
    asm("{ .reg .b16 lo, hi;\n\t"
        "  mov.b32 {lo, hi}, %2;\n\t"
        "  cvt.f64.f16 %0, lo;\n\t"
        "  cvt.f64.f16 %1, hi; }"
        : "=d"(z[2 * i]), "=d"(z[2 * i + 1]) : "r"(w[i]));
  }
}

// LANES threads per (granule, cell) pair: 16 when a record has <= 15 chunks
// (every BASELINE product), 32 otherwise.  Both halves of a warp run the same
// instruction stream on different pairs; an idle half re-does its neighbour's
// pair and discards the result, so the warp never diverges.
template <int LANES>
__global__ void __launch_bounds__(kFusedThreads, 4)
fused_amf_kernel(const __grid_constant__ FusedParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kGroups = kFusedThreads / LANES;
  const oisat_fused_args& A = P.a;
  const int n_ctm = A.n_ctm_lev;
  const int cpitch = n_ctm + 1;
  float* ctm_s = reinterpret_cast<float*>(smem_raw);  // [nfield][32][cpitch]
  double* scratch_base = reinterpret_cast<double*>(
      smem_raw + (((size_t)P.nfield * kTileCells * cpitch * sizeof(float) + 15) / 16) * 16);
  const int lane = threadIdx.x & 31;
  const int gl = lane & (LANES - 1);
  const int grp = threadIdx.x / LANES;
  const int64_t tile = blockIdx.x;
  const int g = A.tile_granule[tile];
  const int cell0 = A.tile_cell0[tile];
  const int64_t pair0 = A.tile_pair0[tile];
  const uint32_t mask = A.tile_mask[tile];
  const bool has_trop = A.has_trop != 0;

  const int L = A.n_sat_lev;
  const int S = 3 * A.nwin;
  const int64_t rec0 = A.gran_record0[g];
  const int64_t px0 = A.gran_px0[g];
  const uint4* records = reinterpret_cast<const uint4*>(A.records);
  const int n_in_tile = __popc(mask);
  constexpr int kSweep = (LANES == 16) ? 15 : 30;  // whole window nodes per sweep

  // Lane k of a group fetches stencil entry k of its pair: the chunk index of
  // the vertex record (host checks it fits 32 bits), the weight, and the AMF term
  // (AMF is not float16, hence not in the record; the pack step has NaN-masked it).
  auto load_entries = [&](int64_t pair, int base, int nk, uint32_t& cix, double& wt, double& za) {
    cix = 0;
    wt = 0.0;
    za = 0.0;
    if (gl < nk) {
      const int32_t v = A.vert[pair * S + base + gl];
      wt = A.w[pair * S + base + gl];
      cix = (uint32_t)((rec0 + v) * P.nchunk);
      za = wt * A.amf_masked[px0 + v];
    }
  };
  // first round's entries are requested before the model tile is staged, so the
  // two dependent global-load chains overlap
  uint32_t pre_cix;
  double pre_wt, pre_za;
  load_entries(pair0 + (grp < n_in_tile ? grp : n_in_tile - 1), 0, S < kSweep ? S : kSweep, pre_cix,
               pre_wt, pre_za);

  // ---- stage the model tile: 32 consecutive cells x n_ctm levels x nfield
  // derived fields.  Global reads are coalesced along cells; the transposed,
  // padded shared layout [cell][level] makes the per-cell column reads
  // conflict-free.
  {
    const int64_t slot_off = (int64_t)A.gran_slot[g] * n_ctm * A.n_cell;
    const int c = lane;
    const bool in_grid = (int64_t)cell0 + c < A.n_cell;
    for (int k = threadIdx.x >> 5; k < n_ctm; k += kFusedThreads / 32) {
      const int64_t src = slot_off + (int64_t)k * A.n_cell + cell0 + c;
      float lp = 0.f, pc = 0.f, pm = 0.f;
      if (in_grid) {
        lp = __ldg(A.ctm_logp + src);
        pc = __ldg(A.ctm_pcol + src);
        if (has_trop) pm = __ldg(A.ctm_pmid + src);
      }
      ctm_s[(0 * kTileCells + c) * cpitch + k] = lp;
      ctm_s[(1 * kTileCells + c) * cpitch + k] = pc;
      if (has_trop) ctm_s[(2 * kTileCells + c) * cpitch + k] = pm;
    }
  }
  __syncthreads();

  const GroupScratch s = carve_scratch(scratch_base + (size_t)grp * P.scratch_doubles, L, n_ctm,
                                       P.nrow);
  double* rows = s.xs;  // [xs|ys] is contiguous and holds >= nrow doubles

  const int rounds = (n_in_tile + kGroups - 1) / kGroups;
  for (int rd = 0; rd < rounds; ++rd) {
    int ord = rd * kGroups + grp;
    const bool mine = ord < n_in_tile;
    if (!__any_sync(0xffffffffu, mine)) continue;  // no block-wide barrier below: warps may skip
    if (!mine) ord = n_in_tile - 1;  // shadow work of an odd half warp, results discarded
    const int l = (int)__fns(mask, 0, ord + 1);  // cell offset of the ord-th set bit
    const int64_t pair = pair0 + ord;

    double acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0;
    double acc_amf = 0.0;
    for (int base = 0; base < S; base += kSweep) {
      const int nk = (S - base) < kSweep ? (S - base) : kSweep;
      uint32_t cix;
      double wt, za;
      if (rd == 0 && base == 0) { cix = pre_cix; wt = pre_wt; za = pre_za; }
      else load_entries(pair, base, nk, cix, wt, za);
#pragma unroll
      for (int o = LANES / 2; o > 0; o >>= 1) za += __shfl_xor_sync(0xffffffffu, za, o, LANES);
      acc_amf += za;
      for (int node = 0; node < nk; node += 3) {
        uint4 u[3];
        double wk[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {  // issue the three vertex loads together
          const uint32_t ck = __shfl_sync(0xffffffffu, cix, node + j, LANES);
          wk[j] = __shfl_sync(0xffffffffu, wt, node + j, LANES);
          if (gl < P.nchunk) u[j] = __ldg(&records[ck + gl]);
        }
        if (gl < P.nchunk) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            double z[8];
            half8_to_double(u[j], z);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fma(wk[j], z[e], acc[e]);
          }
        }
      }
    }
    // box mean: sum over the window first, scale once (interpolator.py:40-46,66-76)
    if (gl < P.nchunk) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int row = gl + P.nchunk * e;
        if (row < P.nrow) rows[row] = acc[e] * (row == 2 * L + 1 ? A.box_weight_err : A.box_weight);
      }
    }
    const double old_amf = acc_amf * A.box_weight;
    __syncwarp();
    const double vcd = rows[2 * L];
    const double sigma = sqrt(rows[2 * L + 1]);  // interpolator.py:188
    const double trop = has_trop ? rows[2 * L + 2] : 0.0;
    // amf_recal.py:99-100 skips cells without a retrieval.  Masked pixels arrive as
    // NaN records, so such cells are common (cloud fields are coherent): when both
    // halves of the warp are idle or dead the vertical operator is skipped, and a
    // dead half next to a live one runs it on a dummy monotone table (results
    // discarded) so that it stays on the sort's fast path.
    const bool dead = !(vcd == vcd);
    if (__all_sync(0xffffffffu, dead || !mine)) {
      if (mine && gl == 0) {
        A.staged[0 * A.n_pairs + pair] = qnan();
        A.staged[1 * A.n_pairs + pair] = sigma;
        A.staged[2 * A.n_pairs + pair] = qnan();
        A.staged[3 * A.n_pairs + pair] = qnan();
        A.staged[4 * A.n_pairs + pair] = old_amf;
      }
      __syncwarp();
      continue;
    }
    for (int i = gl; i < L; i += LANES) {
      s.xr[i] = dead ? (double)i : log(rows[L + i]);
      s.yr[i] = dead ? 0.0 : rows[i];
    }
    __syncwarp();
    const float* clp = ctm_s + (0 * kTileCells + l) * cpitch;
    const float* cpc = ctm_s + (1 * kTileCells + l) * cpitch;
    const float* cpm = ctm_s + (2 * kTileCells + l) * cpitch;
    double colsum;
    double new_amf = group_amf_cell<LANES, true, true>(
        s, L, n_ctm, has_trop, trop, [&](int k) { return (double)cpm[k]; },
        [&](int k) { return (double)clp[k]; }, [&](int k) { return (double)cpc[k]; }, &colsum,
        lane);
    double vnew = qnan(), col = qnan();
    if (vcd == vcd) {                                         // amf_recal.py:99-100
      vnew = (old_amf * vcd) / new_amf;                       // :179
      col = (vnew != vnew || isinf(vnew)) ? qnan() : colsum;  // :180-181
    } else {
      new_amf = qnan();                                       // :176
    }
    if (mine && gl == 0) {
      A.staged[0 * A.n_pairs + pair] = vnew;
      A.staged[1 * A.n_pairs + pair] = sigma;
      A.staged[2 * A.n_pairs + pair] = col;
      A.staged[3 * A.n_pairs + pair] = new_amf;
      A.staged[4 * A.n_pairs + pair] = old_amf;
    }
    __syncwarp();
  }
}

static size_t fused_smem_bytes(const FusedParams& P, int lanes) {
  const size_t ctm = (((size_t)P.nfield * kTileCells * (P.a.n_ctm_lev + 1) * sizeof(float)) + 15) /
                     16 * 16;
  return ctm + (size_t)(kFusedThreads / lanes) * P.scratch_doubles * sizeof(double);
}

}  // namespace oisat

using namespace oisat;

extern "C" int oisat_ctm_prepare(const float* pmid, const float* prof, const float* dp, int64_t n,
                                 float* logp, float* pcol, void* stream) {
  if (n <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(pmid && prof && dp && logp && pcol, "null pointer");
  OISAT_CHECK_ARG(((uintptr_t)pmid | (uintptr_t)prof | (uintptr_t)dp | (uintptr_t)logp |
                   (uintptr_t)pcol) % 16 == 0, "pointers must be 16-byte aligned");
  ctm_prepare_kernel<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, (cudaStream_t)stream>>>(
      pmid, prof, dp, n, logp, pcol);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

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
  PackSrc s{(const __half*)sw, (const __half*)p_mid, (const __half*)vcd, (const __half*)sigma,
            (const __half*)trop, n_px};
  OISAT_CHECK_CUDA(cudaFuncSetAttribute(pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
  pack_kernel<<<(unsigned)ceil_div(n_px, kPackPixels), 256, smem, (cudaStream_t)stream>>>(
      s, n_sat_lev, trop != nullptr, (__half*)records);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int64_t oisat_pack_blocks(int64_t n_px) { return ceil_div(n_px, kPackPixels); }

extern "C" int oisat_pack_batch_indexed(const oisat_pack_item* items, int32_t n_items,
                                        int64_t total_blocks, const int32_t* block_item,
                                        int32_t n_sat_lev, int32_t has_trop, int32_t qflag_dtype,
                                        double flag_thresh, int32_t amf_dtype, void* records,
                                        double* amf_masked, void* stream);

extern "C" int oisat_pack_batch(const oisat_pack_item* items, int32_t n_items,
                                int64_t total_blocks, int32_t n_sat_lev, int32_t has_trop,
                                int32_t qflag_dtype, double flag_thresh, int32_t amf_dtype,
                                void* records, double* amf_masked, void* stream) {
  return oisat_pack_batch_indexed(items, n_items, total_blocks, nullptr, n_sat_lev, has_trop,
                                  qflag_dtype, flag_thresh, amf_dtype, records, amf_masked, stream);
}

extern "C" int oisat_pack_batch_indexed(const oisat_pack_item* items, int32_t n_items,
                                        int64_t total_blocks, const int32_t* block_item,
                                        int32_t n_sat_lev, int32_t has_trop, int32_t qflag_dtype,
                                        double flag_thresh, int32_t amf_dtype, void* records,
                                        double* amf_masked, void* stream) {
  return oisat_pack_batch_masked(items, n_items, total_blocks, block_item, n_sat_lev, has_trop,
                                 qflag_dtype, flag_thresh, amf_dtype, records, amf_masked, nullptr,
                                 0, stream);
}

extern "C" int oisat_pack_batch_masked(const oisat_pack_item* items, int32_t n_items,
                                       int64_t total_blocks, const int32_t* block_item,
                                       int32_t n_sat_lev, int32_t has_trop, int32_t qflag_dtype,
                                       double flag_thresh, int32_t amf_dtype, void* records,
                                       double* amf_masked, uint8_t* px_bad,
                                       int32_t skip_masked_records, void* stream) {
  if (n_items <= 0 || total_blocks <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(!skip_masked_records || px_bad, "skipping masked records needs the mask output");
  OISAT_CHECK_ARG(items && records && amf_masked, "null pointer");
  OISAT_CHECK_ARG(amf_dtype == OISAT_F16 || amf_dtype == OISAT_F32 || amf_dtype == OISAT_F64,
                  "bad amf dtype");
  OISAT_CHECK_ARG(n_sat_lev >= 2 && n_sat_lev <= kMaxSatLev, "bad level count");
  OISAT_CHECK_ARG(qflag_dtype == OISAT_F16 || qflag_dtype == OISAT_F32 || qflag_dtype == OISAT_F64,
                  "bad quality-flag dtype");
  const int R = 8 * record_chunks(n_sat_lev, has_trop);
  const size_t smem_ldg = (size_t)kPackPixels * (R + 2) * sizeof(__half);
  const size_t smem_bulk = (size_t)R * kTmaPitch * sizeof(__half);
  const size_t smem = smem_ldg > smem_bulk ? smem_ldg : smem_bulk;
  const char* mode = std::getenv("OISAT_PACK");  // "ldg" forces the load/store path (A/B runs)
  const int use_bulk = !(mode && mode[0] == 'l');
  OISAT_CHECK_CUDA(cudaFuncSetAttribute(pack_batch_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pack_batch_kernel<<<(unsigned)total_blocks, 256, smem, (cudaStream_t)stream>>>(
      items, n_items, n_sat_lev, has_trop, qflag_dtype, flag_thresh, amf_dtype, (__half*)records,
      amf_masked, use_bulk, block_item, px_bad, skip_masked_records);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}

extern "C" int oisat_fused_amf(const oisat_fused_args* h_args, void* stream) {
  OISAT_CHECK_ARG(h_args != nullptr, "null args");
  const oisat_fused_args& a = *h_args;
  if (a.n_tiles == 0 || a.n_pairs == 0) return OISAT_OK;
  OISAT_CHECK_ARG(a.tile_granule && a.tile_cell0 && a.tile_pair0 && a.tile_mask && a.vert && a.w &&
                      a.gran_record0 && a.gran_px0 && a.gran_slot && a.records &&
                      a.amf_masked && a.ctm_logp && a.ctm_pcol && a.staged,
                  "null pointer");
  OISAT_CHECK_ARG(!a.has_trop || a.ctm_pmid, "tropopause masking needs the model p_mid");
  OISAT_CHECK_ARG(a.nwin >= 1 && a.n_sat_lev >= 2 && a.n_sat_lev <= kMaxSatLev, "bad stencil");
  OISAT_CHECK_ARG(a.n_ctm_lev >= 2 && a.n_ctm_lev <= kMaxCtmLev, "bad model level count");
  FusedParams P;
  P.a = a;
  P.nrow = record_rows(a.n_sat_lev, a.has_trop);
  P.nchunk = record_chunks(a.n_sat_lev, a.has_trop);
  P.nfield = a.has_trop ? 3 : 2;
  P.scratch_doubles = scratch_doubles(a.n_sat_lev, a.n_ctm_lev, P.nrow);
  OISAT_CHECK_ARG(P.nchunk < 32, "record too wide");
  OISAT_CHECK_ARG(a.n_records > 0 && a.n_records * P.nchunk < ((int64_t)1 << 32),
                  "record block too large for 32-bit chunk indices: split the batch");
  const int lanes = P.nchunk < 16 ? 16 : 32;
  const size_t smem = fused_smem_bytes(P, lanes);
  auto kern = lanes == 16 ? fused_amf_kernel<16> : fused_amf_kernel<32>;
  OISAT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
  kern<<<(unsigned)a.n_tiles, kFusedThreads, smem, (cudaStream_t)stream>>>(P);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
