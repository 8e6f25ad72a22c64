// COO edge_index -> destination-sorted CSR (+ source-sorted transposed CSR) for the fused GATv2
// kernels.  Stable: edges of one row keep their original relative order, so every floating-point
// segment reduction downstream has a fixed summation order (deterministic, no atomics).
//
// Replaces the implicit scatter/gather indexing of PyG's MessagePassing.propagate
// (call site /root/reference/src/segger/models/ist_encoder.py:183-189).
#include "sgb_api_internal.cuh"
#include "sgb_sort.cuh"

namespace sgb {
namespace {

template <typename IdxT>
__global__ void csr_convert_kernel(const IdxT* __restrict__ ei, int64_t row_stride, int64_t col_stride,
                                   int64_t E, int64_t n_src, int64_t n_dst, uint32_t* __restrict__ src32,
                                   uint32_t* __restrict__ dst32, int32_t* __restrict__ status) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int64_t s = static_cast<int64_t>(ei[e * col_stride]);
  int64_t d = static_cast<int64_t>(ei[row_stride + e * col_stride]);
  if (s < 0 || s >= n_src || d < 0 || d >= n_dst) {
    if (status) atomicOr(status, 1);
    s = s < 0 ? 0 : (s >= n_src ? n_src - 1 : s);
    d = d < 0 ? 0 : (d >= n_dst ? n_dst - 1 : d);
  }
  src32[e] = static_cast<uint32_t>(s);
  dst32[e] = static_cast<uint32_t>(d);
}

__global__ void csr_gather_kernel(const uint32_t* __restrict__ eid, const uint32_t* __restrict__ other,
                                  int32_t* __restrict__ col, int64_t E) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < E) col[i] = static_cast<int32_t>(other[eid[i]]);
}

__global__ void csr_inverse_kernel(const int32_t* __restrict__ dst_eid, uint32_t* __restrict__ pos_of, int64_t E) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < E) pos_of[dst_eid[i]] = static_cast<uint32_t>(i);
}

__global__ void csr_transpose_fill_kernel(const uint32_t* __restrict__ src_eid, const uint32_t* __restrict__ dst32,
                                          const uint32_t* __restrict__ pos_of, int32_t* __restrict__ t_dst,
                                          int32_t* __restrict__ t_pos, int64_t E) {
  const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k < E) {
    const uint32_t e = src_eid[k];
    t_dst[k] = static_cast<int32_t>(dst32[e]);
    t_pos[k] = static_cast<int32_t>(pos_of[e]);
  }
}

struct CsrWs {
  uint32_t *src32, *dst32, *skeys, *seid, *pos_of;
  void* sort_ws;
  size_t sort_bytes, total;
};

CsrWs carve(void* ws, int64_t E) {
  CsrWs c{};
  const size_t eb = align_up(static_cast<size_t>(E > 0 ? E : 1) * 4);
  char* p = static_cast<char*>(ws);
  c.src32 = reinterpret_cast<uint32_t*>(p); p += eb;
  c.dst32 = reinterpret_cast<uint32_t*>(p); p += eb;
  c.skeys = reinterpret_cast<uint32_t*>(p); p += eb;
  c.seid = reinterpret_cast<uint32_t*>(p); p += eb;
  c.pos_of = reinterpret_cast<uint32_t*>(p); p += eb;
  c.sort_ws = p;
  c.sort_bytes = sort_pairs_workspace_bytes(E);
  c.total = 5 * eb + c.sort_bytes;
  return c;
}

}  // namespace
}  // namespace sgb

using namespace sgb;

extern "C" size_t sgb_csr_workspace_bytes(int64_t E) { return carve(nullptr, E).total; }

extern "C" int sgb_csr_build(const void* edge_index, int idx_bytes, int64_t row_stride, int64_t col_stride,
                             int64_t E, int64_t n_src, int64_t n_dst, int32_t* dst_rowptr, int32_t* dst_col,
                             int32_t* dst_eid, int32_t* src_rowptr, int32_t* src_dst, int32_t* src_pos,
                             int32_t* status, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "csr_build: idx_bytes must be 4 or 8");
  SGB_REQUIRE(E >= 0 && E < (int64_t(1) << 31), SGB_ERR_RANGE, "csr_build: E=%lld exceeds 2^31-1 per call", (long long)E);
  SGB_REQUIRE(n_src >= 0 && n_src < (int64_t(1) << 31) && n_dst >= 0 && n_dst < (int64_t(1) << 31), SGB_ERR_RANGE,
              "csr_build: node count out of range");
  SGB_REQUIRE(dst_rowptr && (E == 0 || (dst_col && dst_eid && edge_index)), SGB_ERR_ARG, "csr_build: null output");
  const bool want_t = src_rowptr != nullptr;
  SGB_REQUIRE(!want_t || E == 0 || (src_dst && src_pos), SGB_ERR_ARG, "csr_build: null transposed output");
  SGB_REQUIRE(ws && ws_bytes >= sgb_csr_workspace_bytes(E), SGB_ERR_WORKSPACE, "csr_build: workspace too small");
  if (E == 0) {
    cudaMemsetAsync(dst_rowptr, 0, static_cast<size_t>(n_dst + 1) * 4, stream);
    if (want_t) cudaMemsetAsync(src_rowptr, 0, static_cast<size_t>(n_src + 1) * 4, stream);
    return check_launch("csr_build(empty)");
  }
  CsrWs c = carve(ws, E);
  const unsigned blocks = static_cast<unsigned>(ceil_div(E, 256));
  if (idx_bytes == 8)
    csr_convert_kernel<int64_t><<<blocks, 256, 0, stream>>>(static_cast<const int64_t*>(edge_index), row_stride,
                                                            col_stride, E, n_src, n_dst, c.src32, c.dst32, status);
  else
    csr_convert_kernel<int32_t><<<blocks, 256, 0, stream>>>(static_cast<const int32_t*>(edge_index), row_stride,
                                                            col_stride, E, n_src, n_dst, c.src32, c.dst32, status);
  int rc = sort_pairs(c.dst32, nullptr, c.skeys, reinterpret_cast<uint32_t*>(dst_eid), E, bits_for(n_dst), c.sort_ws,
                      c.sort_bytes, stream, true);
  if (rc != SGB_OK) return rc;
  rc = rowptr_from_sorted(c.skeys, E, dst_rowptr, n_dst, stream);
  if (rc != SGB_OK) return rc;
  csr_gather_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(dst_eid), c.src32, dst_col, E);
  if (want_t) {
    rc = sort_pairs(c.src32, nullptr, c.skeys, c.seid, E, bits_for(n_src), c.sort_ws, c.sort_bytes, stream, true);
    if (rc != SGB_OK) return rc;
    rc = rowptr_from_sorted(c.skeys, E, src_rowptr, n_src, stream);
    if (rc != SGB_OK) return rc;
    csr_inverse_kernel<<<blocks, 256, 0, stream>>>(dst_eid, c.pos_of, E);
    csr_transpose_fill_kernel<<<blocks, 256, 0, stream>>>(c.seid, c.dst32, c.pos_of, src_dst, src_pos, E);
  }
  return check_launch("csr_build");
}
