// Pair tables of a month on the device, from the concatenated granule plans:
//
//   oisat_segment_tables  the pairs of every model cell in GRANULE ORDER (what the ordered
//                         accumulation walks, averaging.py:64-108): a counting sort by cell --
//                         histogram, one-block scan, scatter through per-cell cursors (arrival
//                         order), then every cell sorts its own short segment by pair index
//                         (a granule holds a cell at most once, so pair order IS granule order).
//                         Deterministic result although the scatter uses atomics: the last
//                         step fixes the order.  Replaces torch.sort + bincount + cumsum.
//   oisat_pair_tables     the per-pair copies of the per-granule facts the tile kernel starts
//                         from (first record of the pair's granule, element offset of its model
//                         column).  Replaces three torch indexing kernels.
#include "common.cuh"

namespace oisat {

__global__ void __launch_bounds__(256)
seg_count_kernel(const int32_t* __restrict__ cell, int64_t n, int32_t* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&cnt[cell[i]], 1);
}

constexpr int kScanThreads = 1024;

// one block: exclusive scan of cnt[0..n_cell) -> seg_start (int64), cursors left in cnt
__global__ void __launch_bounds__(kScanThreads)
seg_scan_kernel(int32_t* __restrict__ cnt, int64_t n_cell, int64_t* __restrict__ seg_start) {
  __shared__ long long part[kScanThreads];
  const int64_t per = (n_cell + kScanThreads - 1) / kScanThreads;
  const int64_t c0 = (int64_t)threadIdx.x * per;
  const int64_t c1 = c0 + per < n_cell ? c0 + per : n_cell;
  long long s = 0;
  for (int64_t c = c0; c < c1; ++c) s += cnt[c];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 1; o < kScanThreads; o <<= 1) {   // Hillis-Steele inclusive scan
    const long long v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  long long run = part[threadIdx.x] - s;          // exclusive prefix of this thread's chunk
  for (int64_t c = c0; c < c1; ++c) {
    const int k = cnt[c];
    seg_start[c] = run;
    cnt[c] = (int32_t)run;                        // cursor of the scatter
    run += k;
  }
  if (threadIdx.x == kScanThreads - 1) seg_start[n_cell] = part[kScanThreads - 1];
}

__global__ void __launch_bounds__(256)
seg_scatter_kernel(const int32_t* __restrict__ cell, int64_t n, int32_t* __restrict__ cursor,
                   int64_t* __restrict__ seg_pair) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) seg_pair[atomicAdd(&cursor[cell[i]], 1)] = i;
}

__global__ void __launch_bounds__(256)
seg_sort_kernel(const int64_t* __restrict__ seg_start, int64_t n_cell, int64_t* __restrict__ seg_pair) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cell) return;
  const int64_t b = seg_start[c], e = seg_start[c + 1];
  for (int64_t i = b + 1; i < e; ++i) {            // insertion sort: segments are short (<= granules)
    const int64_t v = seg_pair[i];
    int64_t j = i;
    while (j > b && seg_pair[j - 1] > v) {
      seg_pair[j] = seg_pair[j - 1];
      --j;
    }
    seg_pair[j] = v;
  }
}

__global__ void __launch_bounds__(256)
pair_tables_kernel(int64_t n, const int32_t* __restrict__ pair_granule,
                   const int32_t* __restrict__ pair_cell, const int64_t* __restrict__ gran_px0,
                   const int32_t* __restrict__ gran_slot, int64_t slot_elems,
                   int64_t* __restrict__ pair_record0, uint32_t* __restrict__ pair_ctm_off) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = pair_granule[i];
  pair_record0[i] = gran_px0[g];
  pair_ctm_off[i] = (uint32_t)((int64_t)gran_slot[g] * slot_elems + pair_cell[i]);
}

}  // namespace oisat

using namespace oisat;

extern "C" int oisat_segment_tables(const int32_t* pair_cell, int64_t n_pairs, int64_t n_cell,
                                    int64_t* seg_start, int64_t* seg_pair, int32_t* work,
                                    void* stream) {
  OISAT_CHECK_ARG(n_cell > 0 && seg_start && work, "bad cell table");
  OISAT_CHECK_ARG(n_pairs >= 0 && n_pairs < ((int64_t)1 << 31), "too many pairs");
  cudaStream_t s = (cudaStream_t)stream;
  OISAT_CHECK_CUDA(cudaMemsetAsync(work, 0, sizeof(int32_t) * (size_t)n_cell, s));
  if (n_pairs > 0) {
    OISAT_CHECK_ARG(pair_cell && seg_pair, "null pointer");
    seg_count_kernel<<<(unsigned)ceil_div(n_pairs, 256), 256, 0, s>>>(pair_cell, n_pairs, work);
    OISAT_CHECK_LAUNCH();
  }
  seg_scan_kernel<<<1, kScanThreads, 0, s>>>(work, n_cell, seg_start);
  OISAT_CHECK_LAUNCH();
  if (n_pairs > 0) {
    seg_scatter_kernel<<<(unsigned)ceil_div(n_pairs, 256), 256, 0, s>>>(pair_cell, n_pairs, work,
                                                                      seg_pair);
    OISAT_CHECK_LAUNCH();
    seg_sort_kernel<<<(unsigned)ceil_div(n_cell, 256), 256, 0, s>>>(seg_start, n_cell, seg_pair);
    OISAT_CHECK_LAUNCH();
  }
  return OISAT_OK;
}

extern "C" int oisat_pair_tables(int64_t n_pairs, const int32_t* pair_granule,
                                 const int32_t* pair_cell, const int64_t* gran_px0,
                                 const int32_t* gran_slot, int32_t n_ctm_lev, int64_t n_cell,
                                 int64_t n_slots, int64_t* pair_record0, uint32_t* pair_ctm_off,
                                 void* stream) {
  if (n_pairs <= 0) return OISAT_OK;
  OISAT_CHECK_ARG(pair_granule && pair_cell && gran_px0 && gran_slot && pair_record0 && pair_ctm_off,
                  "null pointer");
  OISAT_CHECK_ARG(n_slots * (int64_t)n_ctm_lev * n_cell < ((int64_t)1 << 32),
                  "model block too large for 32-bit element offsets: fewer time slots per batch");
  pair_tables_kernel<<<(unsigned)ceil_div(n_pairs, 256), 256, 0, (cudaStream_t)stream>>>(
      n_pairs, pair_granule, pair_cell, gran_px0, gran_slot, (int64_t)n_ctm_lev * n_cell,
      pair_record0, pair_ctm_off);
  OISAT_CHECK_LAUNCH();
  return OISAT_OK;
}
