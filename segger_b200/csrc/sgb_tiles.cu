// Device-side tile slicing and batch assembly (SURVEY 8f rows N2 second half and N3) + the masked
// compaction at the end of predict_step.
//
// Replaces, on the device and without per-attribute host round trips:
//   * TilePredictDataset._subset   (/root/reference/src/segger/data/tile_dataset.py:218-246): four compares +
//     chunked nonzero per node type, then HeteroData.subgraph (PyG bipartite_subgraph with relabelling);
//   * PartitionDataset.__getitem__ (/root/reference/src/segger/data/partition/dataset.py:512-579) + the PyG
//     DataLoader collate: slices of tile-contiguous node / edge stores concatenated into one batch with
//     shifted edge indices and a `batch` vector;
//   * `x[mask].cpu()` x 4 at the end of LitISTEncoder.predict_step (models/lightning_model.py:294-298).
// Everything is integer / byte movement: flags -> exclusive scan -> ordered scatter (order-preserving, so the
// results equal the reference's nonzero / boolean-mask semantics element for element).
#include "sgb_api_internal.cuh"
#include "sgb_sort.cuh"

namespace sgb {
namespace {

constexpr int kT = 256;

__global__ void mask_flags_kernel(const uint8_t* __restrict__ mask, int64_t n, int32_t* __restrict__ flags) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = mask[i] ? 1 : 0;
}

// torch compares a float32 tensor with a python scalar in float32 (the scalar is rounded to the tensor's dtype);
// a float64 tensor compares in float64.  T = dtype of the positions.
template <typename T>
__global__ void box_flags_kernel(const T* __restrict__ pos, int64_t n, T x0, T y0, T x1, T y1,
                                 int32_t* __restrict__ flags) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T x = pos[2 * i], y = pos[2 * i + 1];
  flags[i] = (x >= x0 && x < x1 && y >= y0 && y < y1) ? 1 : 0;   // outer box: half-open (tile_dataset.py:233-238)
}

// sel[scan[i]] = i for flagged i; map[i] = scan[i] or -1; count = scan[n]
__global__ void select_scatter_kernel(const int32_t* __restrict__ flags, const int32_t* __restrict__ scan, int64_t n,
                                      int32_t* __restrict__ sel, int32_t* __restrict__ map,
                                      int32_t* __restrict__ count) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0 && count) *count = scan[n];
  if (i >= n) return;
  const bool f = flags[i] != 0;
  if (f && sel) sel[scan[i]] = static_cast<int32_t>(i);
  if (map) map[i] = f ? scan[i] : -1;
}

template <typename T>
__global__ void inner_mask_kernel(const T* __restrict__ pos, const int32_t* __restrict__ sel,
                                  const int32_t* __restrict__ count, T x0, T y0, T x1, T y1,
                                  uint8_t* __restrict__ inner) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= *count) return;
  const int64_t i = sel[k];
  const T x = pos[2 * i], y = pos[2 * i + 1];
  inner[k] = (x >= x0 && x <= x1 && y >= y0 && y <= y1) ? 1 : 0;   // inner box: closed (tile_dataset.py:239-244)
}

// out[k, :] = src[sel[k], :] for k < *count (or m when count == NULL); rows of `words` 4-byte words
__global__ void gather_rows_words_kernel(const uint32_t* __restrict__ src, int64_t words, const int32_t* __restrict__ sel,
                                         const int32_t* __restrict__ count, int64_t m, uint32_t* __restrict__ dst) {
  const int64_t lim = count ? static_cast<int64_t>(*count) : m;
  const int64_t total = lim * words;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t k = t / words, w = t - k * words;
    dst[t] = src[static_cast<int64_t>(sel[k]) * words + w];
  }
}
__global__ void gather_rows_bytes_kernel(const uint8_t* __restrict__ src, int64_t row_bytes, const int32_t* __restrict__ sel,
                                         const int32_t* __restrict__ count, int64_t m, uint8_t* __restrict__ dst) {
  const int64_t lim = count ? static_cast<int64_t>(*count) : m;
  const int64_t total = lim * row_bytes;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t k = t / row_bytes, w = t - k * row_bytes;
    dst[t] = src[static_cast<int64_t>(sel[k]) * row_bytes + w];
  }
}

template <typename IdxT>
__global__ void edge_flags_kernel(const IdxT* __restrict__ ei, int64_t row_stride, int64_t col_stride, int64_t E,
                                  const int32_t* __restrict__ map_src, int64_t n_src, const int32_t* __restrict__ map_dst,
                                  int64_t n_dst, int32_t* __restrict__ flags) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t s = static_cast<int64_t>(ei[e * col_stride]);
  const int64_t d = static_cast<int64_t>(ei[row_stride + e * col_stride]);
  const bool ok = s >= 0 && s < n_src && d >= 0 && d < n_dst && map_src[s] >= 0 && map_dst[d] >= 0;
  flags[e] = ok ? 1 : 0;
}
template <typename IdxT>
__global__ void edge_scatter_kernel(const IdxT* __restrict__ ei, int64_t row_stride, int64_t col_stride, int64_t E,
                                    const int32_t* __restrict__ map_src, const int32_t* __restrict__ map_dst,
                                    const int32_t* __restrict__ flags, const int32_t* __restrict__ scan,
                                    IdxT* __restrict__ out, int64_t ld_out, int32_t* __restrict__ kept_eid,
                                    int32_t* __restrict__ count) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e == 0 && count) *count = scan[E];
  if (e >= E || !flags[e]) return;
  const int64_t k = scan[e];
  out[k] = static_cast<IdxT>(map_src[ei[e * col_stride]]);
  out[ld_out + k] = static_cast<IdxT>(map_dst[ei[row_stride + e * col_stride]]);
  if (kept_eid) kept_eid[k] = static_cast<int32_t>(e);
}

// ---- batch assembly: K row ranges -> one contiguous block ------------------------------------------------------
__device__ __forceinline__ int find_range(const int64_t* __restrict__ out_off, int K, int64_t r) {
  int lo = 0, hi = K;      // largest k with out_off[k] <= r
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (out_off[mid] <= r) lo = mid; else hi = mid;
  }
  return lo;
}
__global__ void ranges_gather_words_kernel(const uint32_t* __restrict__ src, int64_t words, const int64_t* __restrict__ starts,
                                           const int64_t* __restrict__ out_off, int K, uint32_t* __restrict__ dst) {
  const int64_t total = out_off[K] * words;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = t / words, w = t - r * words;
    const int k = find_range(out_off, K, r);
    dst[t] = src[(starts[k] + (r - out_off[k])) * words + w];
  }
}
__global__ void ranges_gather_bytes_kernel(const uint8_t* __restrict__ src, int64_t row_bytes, const int64_t* __restrict__ starts,
                                           const int64_t* __restrict__ out_off, int K, uint8_t* __restrict__ dst) {
  const int64_t total = out_off[K] * row_bytes;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = t / row_bytes, w = t - r * row_bytes;
    const int k = find_range(out_off, K, r);
    dst[t] = src[(starts[k] + (r - out_off[k])) * row_bytes + w];
  }
}
template <typename IdxT>
__global__ void edges_collate_kernel(const IdxT* __restrict__ ei, int64_t ld_in, const int64_t* __restrict__ e_starts,
                                     const int64_t* __restrict__ e_out_off, const int64_t* __restrict__ src_starts,
                                     const int64_t* __restrict__ src_out_off, const int64_t* __restrict__ dst_starts,
                                     const int64_t* __restrict__ dst_out_off, int K, IdxT* __restrict__ out,
                                     int64_t ld_out) {
  const int64_t total = e_out_off[K];
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int k = find_range(e_out_off, K, t);
    const int64_t e = e_starts[k] + (t - e_out_off[k]);
    out[t] = static_cast<IdxT>(static_cast<int64_t>(ei[e]) - src_starts[k] + src_out_off[k]);
    out[ld_out + t] = static_cast<IdxT>(static_cast<int64_t>(ei[ld_in + e]) - dst_starts[k] + dst_out_off[k]);
  }
}
__global__ void batch_vector_kernel(const int64_t* __restrict__ out_off, int K, int64_t* __restrict__ batch) {
  const int64_t total = out_off[K];
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x)
    batch[t] = find_range(out_off, K, t);
}

// ---- predict_step epilogue -------------------------------------------------------------------------------------
template <typename GeneT>
__global__ void compact_predictions_kernel(const int32_t* __restrict__ flags, const int32_t* __restrict__ scan, int64_t n,
                                           const int64_t* __restrict__ src_idx, const int64_t* __restrict__ seg_idx,
                                           const float* __restrict__ max_sim, const GeneT* __restrict__ gene,
                                           int64_t* __restrict__ o_src, int64_t* __restrict__ o_seg,
                                           float* __restrict__ o_sim, GeneT* __restrict__ o_gene,
                                           int32_t* __restrict__ count) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0) *count = scan[n];
  if (i >= n || !flags[i]) return;
  const int64_t k = scan[i];
  o_src[k] = src_idx[i];
  o_seg[k] = seg_idx[i];
  o_sim[k] = max_sim[i];
  o_gene[k] = gene[i];
}

// ---- tile partitioning (PartitionDataset.__init__) ---------------------------------------------------------------
template <typename IdxT>
__global__ void labels_to_u32_kernel(const IdxT* __restrict__ labels, int64_t n, int64_t range, uint32_t* __restrict__ keys,
                                     int32_t* __restrict__ status) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t v = static_cast<int64_t>(labels[i]);
  if (v < 0 || v >= range) {
    if (status) atomicOr(status, 1);
    v = v < 0 ? 0 : range - 1;
  }
  keys[i] = static_cast<uint32_t>(v);
}
__global__ void invert_perm_kernel(const int32_t* __restrict__ perm, int64_t n, int32_t* __restrict__ inv) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) inv[perm[i]] = static_cast<int32_t>(i);
}
// key = partition of the edge's source if both endpoints are in the same partition, else P (sorted to the end, dropped)
template <typename IdxT, typename LabT>
__global__ void edge_part_keys_kernel(const IdxT* __restrict__ ei, int64_t row_stride, int64_t col_stride, int64_t E,
                                      const LabT* __restrict__ lab_src, int64_t n_src, const LabT* __restrict__ lab_dst,
                                      int64_t n_dst, int64_t P, uint32_t* __restrict__ keys, int32_t* __restrict__ status) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t s = static_cast<int64_t>(ei[e * col_stride]), d = static_cast<int64_t>(ei[row_stride + e * col_stride]);
  uint32_t k = static_cast<uint32_t>(P);
  if (s < 0 || s >= n_src || d < 0 || d >= n_dst) {
    if (status) atomicOr(status, 1);
  } else {
    const int64_t ls = static_cast<int64_t>(lab_src[s]), ld = static_cast<int64_t>(lab_dst[d]);
    if (ls == ld && ls >= 0 && ls < P) k = static_cast<uint32_t>(ls);
  }
  keys[e] = k;
}
template <typename IdxT>
__global__ void edge_part_fill_kernel(const IdxT* __restrict__ ei, int64_t row_stride, int64_t col_stride,
                                      const uint32_t* __restrict__ order, const int32_t* __restrict__ rowptr, int64_t P,
                                      const int32_t* __restrict__ inv_src, const int32_t* __restrict__ inv_dst,
                                      IdxT* __restrict__ out, int64_t ld_out, int32_t* __restrict__ kept_eid) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= rowptr[P]) return;
  const int64_t e = order[k];
  out[k] = static_cast<IdxT>(inv_src[ei[e * col_stride]]);
  out[ld_out + k] = static_cast<IdxT>(inv_dst[ei[row_stride + e * col_stride]]);
  if (kept_eid) kept_eid[k] = static_cast<int32_t>(e);
}

// map[sel[k]] = k (fill < 0) or map[sel[k]] = fill for k < *count: the global -> tile-local node map of a tile subset
__global__ void scatter_rank_kernel(int32_t* __restrict__ map, const int32_t* __restrict__ sel, const int32_t* __restrict__ count,
                                    int64_t m, int fill_mode, int32_t fill) {
  const int64_t lim = count ? static_cast<int64_t>(*count) : m;
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k < lim) map[sel[k]] = fill_mode ? fill : static_cast<int32_t>(k);
}

struct SelWs {
  int32_t *flags, *scan;
  void* scan_ws;
  size_t scan_bytes, total;
};
SelWs carve_sel(void* ws, int64_t n) {
  SelWs c{};
  const size_t a = align_up(static_cast<size_t>(n > 0 ? n : 1) * 4), b = align_up(static_cast<size_t>(n + 1) * 4);
  char* p = static_cast<char*>(ws);
  c.flags = reinterpret_cast<int32_t*>(p); p += a;
  c.scan = reinterpret_cast<int32_t*>(p); p += b;
  c.scan_ws = p;
  c.scan_bytes = scan_workspace_bytes(n);
  c.total = a + b + c.scan_bytes;
  return c;
}
inline unsigned blocks_for(int64_t n) { return static_cast<unsigned>(ceil_div(n > 0 ? n : 1, kT)); }
inline unsigned stride_blocks(int64_t n) {
  return static_cast<unsigned>(std::min<int64_t>(ceil_div(n > 0 ? n : 1, kT), static_cast<int64_t>(sm_count()) * 16));
}

int finish_select(const SelWs& c, int64_t n, int32_t* sel, int32_t* map, int32_t* count, cudaStream_t stream) {
  int rc = exclusive_scan_i32(c.flags, c.scan, n, c.scan_ws, c.scan_bytes, stream);
  if (rc != SGB_OK) return rc;
  select_scatter_kernel<<<blocks_for(n), kT, 0, stream>>>(c.flags, c.scan, n, sel, map, count);
  return check_launch("select");
}

}  // namespace
}  // namespace sgb

using namespace sgb;

extern "C" size_t sgb_select_workspace_bytes(int64_t n) { return carve_sel(nullptr, n).total; }

// workspace: keys + sorted keys + sort scratch
extern "C" size_t sgb_argsort_workspace_bytes(int64_t n) {
  return 2 * align_up(static_cast<size_t>(n > 0 ? n : 1) * 4) + sort_pairs_workspace_bytes(n);
}

extern "C" int sgb_argsort_stable(const void* labels, int idx_bytes, int64_t n, int64_t range, int32_t* perm,
                                  int32_t* rowptr, int32_t* status, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "argsort_stable: idx_bytes must be 4 or 8");
  SGB_REQUIRE(n >= 0 && n < (int64_t(1) << 31) && range >= 1 && range < (int64_t(1) << 31), SGB_ERR_RANGE, "argsort_stable: size out of range");
  SGB_REQUIRE(ws && ws_bytes >= sgb_argsort_workspace_bytes(n), SGB_ERR_WORKSPACE, "argsort_stable: workspace too small");
  if (n == 0) {
    if (rowptr) cudaMemsetAsync(rowptr, 0, static_cast<size_t>(range + 1) * 4, stream);
    return check_launch("argsort_stable(empty)");
  }
  SGB_REQUIRE(labels && perm, SGB_ERR_ARG, "argsort_stable: null argument");
  const size_t nb = align_up(static_cast<size_t>(n) * 4);
  uint32_t* keys = static_cast<uint32_t*>(ws);
  uint32_t* skeys = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + nb);
  void* sort_ws = static_cast<char*>(ws) + 2 * nb;
  if (idx_bytes == 8) labels_to_u32_kernel<int64_t><<<blocks_for(n), kT, 0, stream>>>(static_cast<const int64_t*>(labels), n, range, keys, status);
  else labels_to_u32_kernel<int32_t><<<blocks_for(n), kT, 0, stream>>>(static_cast<const int32_t*>(labels), n, range, keys, status);
  int rc = sort_pairs(keys, nullptr, skeys, reinterpret_cast<uint32_t*>(perm), n, bits_for(range), sort_ws,
                      sort_pairs_workspace_bytes(n), stream, true);
  if (rc != SGB_OK) return rc;
  if (rowptr) {
    rc = rowptr_from_sorted(skeys, n, rowptr, range, stream);
    if (rc != SGB_OK) return rc;
  }
  return check_launch("argsort_stable");
}

extern "C" int sgb_invert_permutation(const int32_t* perm, int64_t n, int32_t* inv, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(n >= 0 && (n == 0 || (perm && inv)), SGB_ERR_ARG, "invert_permutation: bad argument");
  if (n == 0) return SGB_OK;
  invert_perm_kernel<<<blocks_for(n), kT, 0, stream>>>(perm, n, inv);
  return check_launch("invert_permutation");
}

extern "C" int sgb_partition_edges(const void* edge_index, int idx_bytes, int64_t row_stride, int64_t col_stride, int64_t E,
                                   const void* lab_src, const void* lab_dst, int lab_bytes, int64_t n_src, int64_t n_dst,
                                   int64_t P, const int32_t* inv_src, const int32_t* inv_dst, void* out_edge_index,
                                   int64_t ld_out, int32_t* kept_eid, int32_t* edge_rowptr /*[P+2]*/, int32_t* status,
                                   void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE((idx_bytes == 4 || idx_bytes == 8) && (lab_bytes == 4 || lab_bytes == 8), SGB_ERR_ARG, "partition_edges: index width must be 4 or 8");
  SGB_REQUIRE(E >= 0 && E < (int64_t(1) << 31) && P >= 1 && P < (int64_t(1) << 30), SGB_ERR_RANGE, "partition_edges: size out of range");
  SGB_REQUIRE(edge_rowptr && ws && ws_bytes >= sgb_argsort_workspace_bytes(E) + align_up(static_cast<size_t>(E > 0 ? E : 1) * 4), SGB_ERR_WORKSPACE,
              "partition_edges: workspace too small");
  if (E == 0) {
    cudaMemsetAsync(edge_rowptr, 0, static_cast<size_t>(P + 2) * 4, stream);
    return check_launch("partition_edges(empty)");
  }
  SGB_REQUIRE(edge_index && lab_src && lab_dst && inv_src && inv_dst && out_edge_index, SGB_ERR_ARG, "partition_edges: null argument");
  const size_t nb = align_up(static_cast<size_t>(E) * 4);
  uint32_t* keys = static_cast<uint32_t*>(ws);
  uint32_t* skeys = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + nb);
  uint32_t* order = reinterpret_cast<uint32_t*>(static_cast<char*>(ws) + 2 * nb);
  void* sort_ws = static_cast<char*>(ws) + 3 * nb;
#define SGB_KEYS(IT, LT)                                                                                          \
  edge_part_keys_kernel<IT, LT><<<blocks_for(E), kT, 0, stream>>>(static_cast<const IT*>(edge_index), row_stride, \
      col_stride, E, static_cast<const LT*>(lab_src), n_src, static_cast<const LT*>(lab_dst), n_dst, P, keys, status)
  if (idx_bytes == 8 && lab_bytes == 8) SGB_KEYS(int64_t, int64_t);
  else if (idx_bytes == 8) SGB_KEYS(int64_t, int32_t);
  else if (lab_bytes == 8) SGB_KEYS(int32_t, int64_t);
  else SGB_KEYS(int32_t, int32_t);
#undef SGB_KEYS
  int rc = sort_pairs(keys, nullptr, skeys, order, E, bits_for(P + 1), sort_ws, sort_pairs_workspace_bytes(E), stream, true);
  if (rc != SGB_OK) return rc;
  rc = rowptr_from_sorted(skeys, E, edge_rowptr, P + 1, stream);     // rowptr[P] = kept edges, rowptr[P+1] = E
  if (rc != SGB_OK) return rc;
  if (idx_bytes == 8)
    edge_part_fill_kernel<int64_t><<<blocks_for(E), kT, 0, stream>>>(static_cast<const int64_t*>(edge_index), row_stride, col_stride,
                                                                     order, edge_rowptr, P, inv_src, inv_dst,
                                                                     static_cast<int64_t*>(out_edge_index), ld_out, kept_eid);
  else
    edge_part_fill_kernel<int32_t><<<blocks_for(E), kT, 0, stream>>>(static_cast<const int32_t*>(edge_index), row_stride, col_stride,
                                                                     order, edge_rowptr, P, inv_src, inv_dst,
                                                                     static_cast<int32_t*>(out_edge_index), ld_out, kept_eid);
  return check_launch("partition_edges");
}

extern "C" int sgb_mask_select(const uint8_t* mask, int64_t n, int32_t* sel, int32_t* map, int32_t* count, void* ws,
                               size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(n >= 0 && n < (int64_t(1) << 31), SGB_ERR_RANGE, "mask_select: n out of range");
  SGB_REQUIRE(count && (n == 0 || mask), SGB_ERR_ARG, "mask_select: null argument");
  SGB_REQUIRE(ws && ws_bytes >= sgb_select_workspace_bytes(n), SGB_ERR_WORKSPACE, "mask_select: workspace too small");
  if (n == 0) { cudaMemsetAsync(count, 0, 4, stream); return check_launch("mask_select(empty)"); }
  SelWs c = carve_sel(ws, n);
  mask_flags_kernel<<<blocks_for(n), kT, 0, stream>>>(mask, n, c.flags);
  return finish_select(c, n, sel, map, count, stream);
}

extern "C" int sgb_box_select(const void* pos, int pos_f64, int64_t n, const double* outer /*x0,y0,x1,y1*/,
                              const double* inner /*x0,y0,x1,y1 or NULL*/, int32_t* sel, int32_t* map,
                              uint8_t* inner_mask, int32_t* count, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(n >= 0 && n < (int64_t(1) << 31), SGB_ERR_RANGE, "box_select: n out of range");
  SGB_REQUIRE(count && outer && (n == 0 || (pos && sel)), SGB_ERR_ARG, "box_select: null argument");
  SGB_REQUIRE(!inner || inner_mask, SGB_ERR_ARG, "box_select: inner box given without inner_mask output");
  SGB_REQUIRE(ws && ws_bytes >= sgb_select_workspace_bytes(n), SGB_ERR_WORKSPACE, "box_select: workspace too small");
  if (n == 0) { cudaMemsetAsync(count, 0, 4, stream); return check_launch("box_select(empty)"); }
  SelWs c = carve_sel(ws, n);
  if (pos_f64)
    box_flags_kernel<double><<<blocks_for(n), kT, 0, stream>>>(static_cast<const double*>(pos), n, outer[0], outer[1],
                                                               outer[2], outer[3], c.flags);
  else
    box_flags_kernel<float><<<blocks_for(n), kT, 0, stream>>>(static_cast<const float*>(pos), n,
                                                              static_cast<float>(outer[0]), static_cast<float>(outer[1]),
                                                              static_cast<float>(outer[2]), static_cast<float>(outer[3]), c.flags);
  int rc = finish_select(c, n, sel, map, count, stream);
  if (rc != SGB_OK) return rc;
  if (inner) {
    if (pos_f64)
      inner_mask_kernel<double><<<blocks_for(n), kT, 0, stream>>>(static_cast<const double*>(pos), sel, count, inner[0],
                                                                  inner[1], inner[2], inner[3], inner_mask);
    else
      inner_mask_kernel<float><<<blocks_for(n), kT, 0, stream>>>(static_cast<const float*>(pos), sel, count,
                                                                 static_cast<float>(inner[0]), static_cast<float>(inner[1]),
                                                                 static_cast<float>(inner[2]), static_cast<float>(inner[3]),
                                                                 inner_mask);
  }
  return check_launch("box_select");
}

extern "C" int sgb_scatter_rank(int32_t* map, const int32_t* sel, const int32_t* count, int64_t m, int fill_mode, int32_t fill,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(m >= 0 && (m == 0 || (map && sel)), SGB_ERR_ARG, "scatter_rank: bad argument");
  if (m == 0) return SGB_OK;
  scatter_rank_kernel<<<blocks_for(m), kT, 0, stream>>>(map, sel, count, m, fill_mode, fill);
  return check_launch("scatter_rank");
}

extern "C" int sgb_gather_rows_bytes(const void* src, int64_t row_bytes, const int32_t* sel, const int32_t* count,
                                     int64_t m, void* dst, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(row_bytes > 0 && m >= 0, SGB_ERR_ARG, "gather_rows_bytes: bad size");
  if (m == 0) return SGB_OK;
  SGB_REQUIRE(src && sel && dst, SGB_ERR_ARG, "gather_rows_bytes: null argument");
  const bool words = row_bytes % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3u) == 0 &&
                     (reinterpret_cast<uintptr_t>(dst) & 3u) == 0;
  if (words)
    gather_rows_words_kernel<<<stride_blocks(m * (row_bytes / 4)), kT, 0, stream>>>(
        static_cast<const uint32_t*>(src), row_bytes / 4, sel, count, m, static_cast<uint32_t*>(dst));
  else
    gather_rows_bytes_kernel<<<stride_blocks(m * row_bytes), kT, 0, stream>>>(static_cast<const uint8_t*>(src), row_bytes,
                                                                             sel, count, m, static_cast<uint8_t*>(dst));
  return check_launch("gather_rows_bytes");
}

extern "C" int sgb_edge_subset(const void* edge_index, int idx_bytes, int64_t row_stride, int64_t col_stride, int64_t E,
                               const int32_t* map_src, int64_t n_src, const int32_t* map_dst, int64_t n_dst,
                               void* out_edge_index, int64_t ld_out, int32_t* kept_eid, int32_t* count, void* ws,
                               size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "edge_subset: idx_bytes must be 4 or 8");
  SGB_REQUIRE(E >= 0 && E < (int64_t(1) << 31), SGB_ERR_RANGE, "edge_subset: E out of range");
  SGB_REQUIRE(count && (E == 0 || (edge_index && map_src && map_dst && out_edge_index)), SGB_ERR_ARG,
              "edge_subset: null argument");
  SGB_REQUIRE(ws && ws_bytes >= sgb_select_workspace_bytes(E), SGB_ERR_WORKSPACE, "edge_subset: workspace too small");
  if (E == 0) { cudaMemsetAsync(count, 0, 4, stream); return check_launch("edge_subset(empty)"); }
  SelWs c = carve_sel(ws, E);
  if (idx_bytes == 8)
    edge_flags_kernel<int64_t><<<blocks_for(E), kT, 0, stream>>>(static_cast<const int64_t*>(edge_index), row_stride,
                                                                 col_stride, E, map_src, n_src, map_dst, n_dst, c.flags);
  else
    edge_flags_kernel<int32_t><<<blocks_for(E), kT, 0, stream>>>(static_cast<const int32_t*>(edge_index), row_stride,
                                                                 col_stride, E, map_src, n_src, map_dst, n_dst, c.flags);
  int rc = exclusive_scan_i32(c.flags, c.scan, E, c.scan_ws, c.scan_bytes, stream);
  if (rc != SGB_OK) return rc;
  if (idx_bytes == 8)
    edge_scatter_kernel<int64_t><<<blocks_for(E), kT, 0, stream>>>(static_cast<const int64_t*>(edge_index), row_stride,
                                                                   col_stride, E, map_src, map_dst, c.flags, c.scan,
                                                                   static_cast<int64_t*>(out_edge_index), ld_out, kept_eid, count);
  else
    edge_scatter_kernel<int32_t><<<blocks_for(E), kT, 0, stream>>>(static_cast<const int32_t*>(edge_index), row_stride,
                                                                   col_stride, E, map_src, map_dst, c.flags, c.scan,
                                                                   static_cast<int32_t*>(out_edge_index), ld_out, kept_eid, count);
  return check_launch("edge_subset");
}

extern "C" int sgb_ranges_gather(const void* src, int64_t row_bytes, const int64_t* starts, const int64_t* out_off, int K,
                                 int64_t total_rows, void* dst, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(row_bytes > 0 && K >= 0 && total_rows >= 0, SGB_ERR_ARG, "ranges_gather: bad size");
  if (K == 0 || total_rows == 0) return SGB_OK;
  SGB_REQUIRE(src && starts && out_off && dst, SGB_ERR_ARG, "ranges_gather: null argument");
  const bool words = row_bytes % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3u) == 0 &&
                     (reinterpret_cast<uintptr_t>(dst) & 3u) == 0;
  if (words)
    ranges_gather_words_kernel<<<stride_blocks(total_rows * (row_bytes / 4)), kT, 0, stream>>>(
        static_cast<const uint32_t*>(src), row_bytes / 4, starts, out_off, K, static_cast<uint32_t*>(dst));
  else
    ranges_gather_bytes_kernel<<<stride_blocks(total_rows * row_bytes), kT, 0, stream>>>(
        static_cast<const uint8_t*>(src), row_bytes, starts, out_off, K, static_cast<uint8_t*>(dst));
  return check_launch("ranges_gather");
}

extern "C" int sgb_edges_collate(const void* edge_index, int idx_bytes, int64_t ld_in, const int64_t* e_starts,
                                 const int64_t* e_out_off, const int64_t* src_starts, const int64_t* src_out_off,
                                 const int64_t* dst_starts, const int64_t* dst_out_off, int K, int64_t total_edges,
                                 void* out, int64_t ld_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "edges_collate: idx_bytes must be 4 or 8");
  SGB_REQUIRE(K >= 0 && total_edges >= 0, SGB_ERR_ARG, "edges_collate: bad size");
  if (K == 0 || total_edges == 0) return SGB_OK;
  SGB_REQUIRE(edge_index && e_starts && e_out_off && src_starts && src_out_off && dst_starts && dst_out_off && out,
              SGB_ERR_ARG, "edges_collate: null argument");
  if (idx_bytes == 8)
    edges_collate_kernel<int64_t><<<stride_blocks(total_edges), kT, 0, stream>>>(
        static_cast<const int64_t*>(edge_index), ld_in, e_starts, e_out_off, src_starts, src_out_off, dst_starts,
        dst_out_off, K, static_cast<int64_t*>(out), ld_out);
  else
    edges_collate_kernel<int32_t><<<stride_blocks(total_edges), kT, 0, stream>>>(
        static_cast<const int32_t*>(edge_index), ld_in, e_starts, e_out_off, src_starts, src_out_off, dst_starts,
        dst_out_off, K, static_cast<int32_t*>(out), ld_out);
  return check_launch("edges_collate");
}

extern "C" int sgb_batch_vector(const int64_t* out_off, int K, int64_t total_rows, int64_t* batch, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(K >= 0 && total_rows >= 0, SGB_ERR_ARG, "batch_vector: bad size");
  if (K == 0 || total_rows == 0) return SGB_OK;
  SGB_REQUIRE(out_off && batch, SGB_ERR_ARG, "batch_vector: null argument");
  batch_vector_kernel<<<stride_blocks(total_rows), kT, 0, stream>>>(out_off, K, batch);
  return check_launch("batch_vector");
}

extern "C" int sgb_compact_predictions(const uint8_t* mask, int64_t n, const int64_t* src_idx, const int64_t* seg_idx,
                                       const float* max_sim, const void* gene, int gene_bytes, int64_t* out_src,
                                       int64_t* out_seg, float* out_sim, void* out_gene, int32_t* count, void* ws,
                                       size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(n >= 0 && n < (int64_t(1) << 31), SGB_ERR_RANGE, "compact_predictions: n out of range");
  SGB_REQUIRE(gene_bytes == 4 || gene_bytes == 8, SGB_ERR_ARG, "compact_predictions: gene ids must be int32 or int64");
  SGB_REQUIRE(count && (n == 0 || (mask && src_idx && seg_idx && max_sim && gene && out_src && out_seg && out_sim && out_gene)),
              SGB_ERR_ARG, "compact_predictions: null argument");
  SGB_REQUIRE(ws && ws_bytes >= sgb_select_workspace_bytes(n), SGB_ERR_WORKSPACE, "compact_predictions: workspace too small");
  if (n == 0) { cudaMemsetAsync(count, 0, 4, stream); return check_launch("compact_predictions(empty)"); }
  SelWs c = carve_sel(ws, n);
  mask_flags_kernel<<<blocks_for(n), kT, 0, stream>>>(mask, n, c.flags);
  int rc = exclusive_scan_i32(c.flags, c.scan, n, c.scan_ws, c.scan_bytes, stream);
  if (rc != SGB_OK) return rc;
  if (gene_bytes == 8)
    compact_predictions_kernel<int64_t><<<blocks_for(n), kT, 0, stream>>>(
        c.flags, c.scan, n, src_idx, seg_idx, max_sim, static_cast<const int64_t*>(gene), out_src, out_seg, out_sim,
        static_cast<int64_t*>(out_gene), count);
  else
    compact_predictions_kernel<int32_t><<<blocks_for(n), kT, 0, stream>>>(
        c.flags, c.scan, n, src_idx, seg_idx, max_sim, static_cast<const int32_t*>(gene), out_src, out_seg, out_sim,
        static_cast<int32_t*>(out_gene), count);
  return check_launch("compact_predictions");
}
