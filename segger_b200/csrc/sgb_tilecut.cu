// Batched prediction-tile cut: TilePredictDataset._subset (/root/reference/src/segger/data/tile_dataset.py:218-246) for
// SEVERAL tiles at once, producing the collated batch directly (what `DataLoader(collate)` makes of the per-tile
// subgraphs, data_module.py:333-344).  The per-tile entry points of sgb_tiles.cu cost ~100 launches and one stream
// synchronisation per tile; at the reference's tile size (50k transcripts) that is all launch latency.  Here a group of
// T tiles ("slots") is cut with two flag kernels per node / edge type.
//
// Inputs come from the coarse index of segger_b200.tiles.TilePredictSet: node ids sorted by grid cell (`perm`), edge ids
// sorted by the cell of their source; a slot's candidates are <= 3 contiguous ranges of those arrays (the 3 x 3 cell
// neighbourhood of its tile, one range per grid row).  Candidate c of the concatenated ranges is handled by thread c.
//
//   nodes:  key = slot << 40 | id << 1 | inside-the-tile-itself, mask = inside the tile grown by the margin.  The host
//           compacts the flagged keys and sorts them: ascending (slot, id) IS the collated node order, the low bit is
//           `predict_mask`, `key >> 40` the `batch` vector.
//   edges:  an edge is kept for a slot iff both endpoints are among the slot's nodes (HeteroData.subgraph): binary
//           search of the endpoint ids in the slot's segment of the sorted node keys; the positions found ARE the
//           collated node numbers.  key = slot << 40 | edge id (sorted by the host -> original edge order per tile).
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/segger_b200.h"
#include "sgb_common.cuh"

namespace sgb {
namespace {

constexpr int kT = 256;
constexpr int kMaxRanges = 1024;

struct Ranges {
  const int64_t* start;   // [R] first element of the range in the sorted id array
  const int64_t* off;     // [R + 1] exclusive prefix sum of the range lengths (candidate numbering)
  const int32_t* slot;    // [R] tile slot the range belongs to
  int R;
};

// largest r with off[r] <= c (among equal offsets the last one: empty ranges are skipped)
__device__ __forceinline__ int find_range(const int64_t* off, int R, int64_t c) {
  int lo = 0, hi = R - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= c) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <typename PT>
__device__ __forceinline__ bool in_outer(const PT* pos, int64_t id, const double* box) {
  const PT x = pos[2 * id], y = pos[2 * id + 1];
  // bounds rounded to the position dtype first, as torch does for `pos >= scalar` (sgb_box_select)
  return x >= static_cast<PT>(box[0]) && x < static_cast<PT>(box[2]) && y >= static_cast<PT>(box[1]) && y < static_cast<PT>(box[3]);
}

template <typename PT>
__global__ void __launch_bounds__(kT) tilecut_nodes_kernel(const int32_t* __restrict__ perm, const PT* __restrict__ pos, Ranges rg,
                                                           int64_t C, const double* __restrict__ boxes,
                                                           int64_t* __restrict__ keys, uint8_t* __restrict__ mask,
                                                           int32_t* __restrict__ slot_counts) {
  __shared__ int64_t s_off[kMaxRanges + 1];
  for (int i = threadIdx.x; i <= rg.R; i += blockDim.x) s_off[i] = rg.off[i];
  __syncthreads();
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int r = find_range(s_off, rg.R, c);
  const int slot = rg.slot[r];
  const int64_t id = perm[rg.start[r] + (c - s_off[r])];
  const double* box = boxes + 8 * slot;
  const bool keep = in_outer(pos, id, box);
  mask[c] = keep ? 1 : 0;
  if (!keep) return;
  const PT x = pos[2 * id], y = pos[2 * id + 1];
  const bool inner = x >= static_cast<PT>(box[4]) && x <= static_cast<PT>(box[6]) && y >= static_cast<PT>(box[5]) &&
                     y <= static_cast<PT>(box[7]);                       // the tile itself: closed (tile_dataset.py:239-244)
  keys[c] = (static_cast<int64_t>(slot) << 40) | (id << 1) | (inner ? 1 : 0);
  atomicAdd(slot_counts + slot, 1);
}

// position of node `id` in keys[lo, hi) (ascending ids within the slot), or -1
__device__ __forceinline__ int64_t find_node(const int64_t* __restrict__ keys, int64_t lo, int64_t hi, int64_t slot, int64_t id) {
  const int64_t want = (slot << 40) | (id << 1);
  int64_t a = lo, b = hi;
  while (a < b) {
    const int64_t mid = (a + b) >> 1;
    if (__ldg(keys + mid) < want) a = mid + 1; else b = mid;
  }
  return (a < hi && (__ldg(keys + a) >> 1) == (want >> 1)) ? a : -1;
}

template <typename IT, typename PT>
__global__ void __launch_bounds__(kT) tilecut_edges_kernel(const int32_t* __restrict__ eperm, const IT* __restrict__ ei, int64_t row_stride,
                                                           int64_t col_stride, Ranges rg, int64_t C, const PT* __restrict__ src_pos,
                                                           const double* __restrict__ boxes, const int64_t* __restrict__ src_keys,
                                                           const int64_t* __restrict__ src_ptr, const int64_t* __restrict__ dst_keys,
                                                           const int64_t* __restrict__ dst_ptr, int64_t* __restrict__ ekeys,
                                                           int32_t* __restrict__ pu, int32_t* __restrict__ pv,
                                                           uint8_t* __restrict__ mask, int32_t* __restrict__ slot_counts) {
  __shared__ int64_t s_off[kMaxRanges + 1];
  for (int i = threadIdx.x; i <= rg.R; i += blockDim.x) s_off[i] = rg.off[i];
  __syncthreads();
  const int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int r = find_range(s_off, rg.R, c);
  const int64_t slot = rg.slot[r];
  const int64_t eid = eperm[rg.start[r] + (c - s_off[r])];
  const int64_t u = static_cast<int64_t>(ei[eid * col_stride]);
  mask[c] = 0;
  // most candidates (the 3 x 3 cells hold ~9x the tile) fail on their source's position: one gather instead of two searches
  if (!in_outer(src_pos, u, boxes + 8 * slot)) return;
  const int64_t a = find_node(src_keys, src_ptr[slot], src_ptr[slot + 1], slot, u);
  if (a < 0) return;
  const int64_t v = static_cast<int64_t>(ei[row_stride + eid * col_stride]);
  const int64_t b = find_node(dst_keys, dst_ptr[slot], dst_ptr[slot + 1], slot, v);
  if (b < 0) return;
  mask[c] = 1;
  ekeys[c] = (slot << 40) | eid;
  pu[c] = static_cast<int32_t>(a);
  pv[c] = static_cast<int32_t>(b);
  atomicAdd(slot_counts + slot, 1);
}

inline unsigned blocks_for(int64_t n) { return static_cast<unsigned>(ceil_div(n > 0 ? n : 1, kT)); }

}  // namespace
}  // namespace sgb

using namespace sgb;

extern "C" int sgb_tilecut_nodes(const int32_t* perm, const void* pos, int pos_f64, const int64_t* rng_start, const int64_t* rng_off,
                                 const int32_t* rng_slot, int R, int64_t C, const double* boxes, int T, int64_t* keys,
                                 uint8_t* mask, int32_t* slot_counts, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(R >= 0 && R <= kMaxRanges && T > 0 && C >= 0 && C < (int64_t(1) << 31), SGB_ERR_RANGE,
              "tilecut_nodes: R=%d (<= %d), T=%d, C=%lld out of range", R, kMaxRanges, T, (long long)C);
  SGB_REQUIRE(slot_counts, SGB_ERR_ARG, "tilecut_nodes: null slot_counts");
  cudaMemsetAsync(slot_counts, 0, sizeof(int32_t) * T, stream);
  if (C == 0) return check_launch("tilecut_nodes(empty)");
  SGB_REQUIRE(perm && pos && rng_start && rng_off && rng_slot && boxes && keys && mask, SGB_ERR_ARG, "tilecut_nodes: null argument");
  Ranges rg{rng_start, rng_off, rng_slot, R};
  if (pos_f64)
    tilecut_nodes_kernel<double><<<blocks_for(C), kT, 0, stream>>>(perm, static_cast<const double*>(pos), rg, C, boxes, keys, mask, slot_counts);
  else
    tilecut_nodes_kernel<float><<<blocks_for(C), kT, 0, stream>>>(perm, static_cast<const float*>(pos), rg, C, boxes, keys, mask, slot_counts);
  return check_launch("tilecut_nodes");
}

extern "C" int sgb_tilecut_edges(const int32_t* eperm, const void* edge_index, int idx_bytes, int64_t row_stride, int64_t col_stride,
                                 const int64_t* rng_start, const int64_t* rng_off, const int32_t* rng_slot, int R, int64_t C,
                                 const void* src_pos, int pos_f64, const double* boxes, int T, const int64_t* src_keys,
                                 const int64_t* src_ptr, const int64_t* dst_keys, const int64_t* dst_ptr, int64_t* ekeys,
                                 int32_t* pu, int32_t* pv, uint8_t* mask, int32_t* slot_counts, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "tilecut_edges: idx_bytes must be 4 or 8");
  SGB_REQUIRE(R >= 0 && R <= kMaxRanges && T > 0 && C >= 0 && C < (int64_t(1) << 31), SGB_ERR_RANGE,
              "tilecut_edges: R=%d (<= %d), T=%d, C=%lld out of range", R, kMaxRanges, T, (long long)C);
  SGB_REQUIRE(slot_counts, SGB_ERR_ARG, "tilecut_edges: null slot_counts");
  cudaMemsetAsync(slot_counts, 0, sizeof(int32_t) * T, stream);
  if (C == 0) return check_launch("tilecut_edges(empty)");
  SGB_REQUIRE(eperm && edge_index && rng_start && rng_off && rng_slot && src_pos && boxes && src_keys && src_ptr && dst_keys &&
                  dst_ptr && ekeys && pu && pv && mask, SGB_ERR_ARG, "tilecut_edges: null argument");
  Ranges rg{rng_start, rng_off, rng_slot, R};
#define SGB_TC_LAUNCH(IT, PT)                                                                                                  \
  tilecut_edges_kernel<IT, PT><<<blocks_for(C), kT, 0, stream>>>(eperm, static_cast<const IT*>(edge_index), row_stride,        \
                                                                 col_stride, rg, C, static_cast<const PT*>(src_pos), boxes,    \
                                                                 src_keys, src_ptr, dst_keys, dst_ptr, ekeys, pu, pv, mask, slot_counts)
  if (idx_bytes == 8) { if (pos_f64) SGB_TC_LAUNCH(int64_t, double); else SGB_TC_LAUNCH(int64_t, float); }
  else { if (pos_f64) SGB_TC_LAUNCH(int32_t, double); else SGB_TC_LAUNCH(int32_t, float); }
#undef SGB_TC_LAUNCH
  return check_launch("tilecut_edges");
}
