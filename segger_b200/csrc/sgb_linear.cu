// C ABI of the dense projections: argument validation + back-end selection.
#include "sgb_api_internal.cuh"
#include "sgb_linear.cuh"
#include <algorithm>

using namespace sgb;

static int check_dims(const char* fn, int64_t M, int64_t N, int64_t K) {
  SGB_REQUIRE(M >= 0 && N >= 0 && K >= 0, SGB_ERR_ARG, "%s: negative dimension", fn);
  SGB_REQUIRE(M < (int64_t(1) << 31) && N < (int64_t(1) << 24) && K < (int64_t(1) << 24), SGB_ERR_RANGE,
              "%s: dimension out of range (M=%lld N=%lld K=%lld)", fn, (long long)M, (long long)N, (long long)K);
  return SGB_OK;
}

// workspace of sgb_linear_fwd (weights [N,K]) and sgb_linear_dgrad (same weights: pass the same N, K)
extern "C" size_t sgb_linear_workspace_bytes(int64_t N, int64_t K) {
  if (N <= 0 || K <= 0 || !tc_enabled()) return 0;
  const size_t a = tc_linear_workspace_bytes(N, K), b = tc_linear_workspace_bytes(K, N);
  return a > b ? a : b;
}

extern "C" int sgb_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* b, int64_t M,
                              int64_t N, int64_t K, float* y, int64_t ldy, int act, float* y_act, int64_t ldya,
                              int exact, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_dims("linear_fwd", M, N, K);
  if (rc != SGB_OK) return rc;
  if (M == 0 || N == 0) return SGB_OK;
  SGB_REQUIRE(y && (K == 0 || (x && w)), SGB_ERR_ARG, "linear_fwd: null tensor");
  SGB_REQUIRE(ldx >= K && ldw >= K && ldy >= N && (!y_act || ldya >= N), SGB_ERR_ARG, "linear_fwd: leading dimension too small");
  SGB_REQUIRE(act >= SGB_ACT_NONE && act <= SGB_ACT_SILU, SGB_ERR_ARG, "linear_fwd: unknown activation %d", act);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (skinny_linear_fwd_ok(x, ldx, M, N, K, y, ldy, y_act, ldya))   // K <= 16: HBM streaming on the fp32 pipe
    return skinny_linear_fwd(x, ldx, w, ldw, b, M, N, K, y, ldy, act, y_act, ldya, st);
  if (exact != 2 && tc_linear_fwd_ok(x, ldx, w, ldw, M, N, K)) {
    SGB_REQUIRE(ws && ws_bytes >= tc_linear_workspace_bytes(N, K), SGB_ERR_WORKSPACE, "linear_fwd: workspace too small");
    return tc_linear_fwd(x, ldx, w, ldw, b, M, N, K, y, ldy, act, y_act, ldya, exact, ws, st);
  }
  return simt_linear_fwd(x, ldx, w, ldw, b, M, N, K, y, ldy, act, y_act, ldya, st);
}

// y = x w^T + b + table[ids]: the forward projection of an input whose first columns are a row of a small table
// selected by an integer id (ISTEncoder: GELU(Embedding[gene id]), /root/reference/src/segger/models/ist_encoder.py:312-320).
// x holds only the remaining (dense) columns; table = act(Embedding) w_first^T is computed once per call by the caller.
namespace {
template <typename IdxT>
__global__ void rows_gather_add_kernel(float* __restrict__ y, int64_t ldy, int64_t M, int64_t N, const IdxT* __restrict__ ids,
                                       const float* __restrict__ tab, int64_t ldt) {
  const int64_t total = M * N;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t m = t / N, n = t - m * N;
    y[m * ldy + n] += __ldg(tab + static_cast<int64_t>(ids[m]) * ldt + n);
  }
}
}  // namespace

extern "C" int sgb_linear_fwd_gather(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* b, int64_t M,
                                     int64_t N, int64_t K, float* y, int64_t ldy, const void* ids, int idx_bytes,
                                     const float* table, int64_t ld_table, int exact, void* ws, size_t ws_bytes,
                                     void* stream) {
  int rc = check_dims("linear_fwd_gather", M, N, K);
  if (rc != SGB_OK) return rc;
  if (M == 0 || N == 0) return SGB_OK;
  SGB_REQUIRE(y && x && w && ids && table && K > 0, SGB_ERR_ARG, "linear_fwd_gather: null tensor");
  SGB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "linear_fwd_gather: idx_bytes must be 4 or 8");
  SGB_REQUIRE(ldx >= K && ldw >= K && ldy >= N && ld_table >= N, SGB_ERR_ARG, "linear_fwd_gather: leading dimension too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (exact != 2 && tc_linear_fwd_ok(x, ldx, w, ldw, M, N, K) && !skinny_linear_fwd_ok(x, ldx, M, N, K, y, ldy, nullptr, 0)) {
    SGB_REQUIRE(ws && ws_bytes >= tc_linear_workspace_bytes(N, K), SGB_ERR_WORKSPACE, "linear_fwd_gather: workspace too small");
    return tc_linear_fwd(x, ldx, w, ldw, b, M, N, K, y, ldy, SGB_ACT_NONE, nullptr, 0, exact, ws, st, ids, idx_bytes, table, ld_table);
  }
  rc = sgb_linear_fwd(x, ldx, w, ldw, b, M, N, K, y, ldy, SGB_ACT_NONE, nullptr, 0, exact, ws, ws_bytes, stream);
  if (rc != SGB_OK) return rc;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(M * N, 256), static_cast<int64_t>(sm_count()) * 16));
  if (idx_bytes == 8) rows_gather_add_kernel<int64_t><<<blocks, 256, 0, st>>>(y, ldy, M, N, static_cast<const int64_t*>(ids), table, ld_table);
  else rows_gather_add_kernel<int32_t><<<blocks, 256, 0, st>>>(y, ldy, M, N, static_cast<const int32_t*>(ids), table, ld_table);
  return check_launch("linear_fwd_gather");
}

extern "C" int sgb_linear_dgrad(const float* dy, int64_t ldy, const float* w, int64_t ldw, int64_t M, int64_t N,
                                int64_t K, float* dx, int64_t ldx, int accumulate, int act, const float* act_pre,
                                int64_t ld_pre, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_dims("linear_dgrad", M, N, K);
  if (rc != SGB_OK) return rc;
  if (M == 0 || K == 0) return SGB_OK;
  SGB_REQUIRE(dx && (N == 0 || (dy && w)), SGB_ERR_ARG, "linear_dgrad: null tensor");
  SGB_REQUIRE(ldy >= N && ldw >= K && ldx >= K && (!act_pre || ld_pre >= K), SGB_ERR_ARG, "linear_dgrad: leading dimension too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (tc_linear_dgrad_ok(dy, ldy, w, ldw, M, N, K)) {
    SGB_REQUIRE(ws && ws_bytes >= tc_linear_workspace_bytes(K, N), SGB_ERR_WORKSPACE, "linear_dgrad: workspace too small");
    return tc_linear_dgrad(dy, ldy, w, ldw, M, N, K, dx, ldx, accumulate, act, act_pre, ld_pre, ws, st);
  }
  return simt_linear_dgrad(dy, ldy, w, ldw, M, N, K, dx, ldx, accumulate, act, act_pre, ld_pre, st);
}

static size_t wgrad_gemm_ws(int64_t M, int64_t N, int64_t K) {
  const size_t a = simt_linear_wgrad_workspace_bytes(M, N, K);
  const size_t b = tc_enabled() ? tc_linear_wgrad_workspace_bytes(M, N, K) : 0;
  const size_t c = (K >= 4 && K <= 16 && N <= 256) ? skinny_linear_wgrad_workspace_bytes(M, N, K) : 0;
  return std::max(a, std::max(b, c));
}

extern "C" size_t sgb_linear_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  return wgrad_gemm_ws(M, N, K) + linear_colsum_workspace_bytes(M, N);
}

extern "C" int sgb_linear_wgrad(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N,
                                int64_t K, float* dw, int64_t lddw, float* db, int accumulate, void* ws,
                                size_t ws_bytes, void* stream) {
  int rc = check_dims("linear_wgrad", M, N, K);
  if (rc != SGB_OK) return rc;
  if (N == 0) return SGB_OK;
  SGB_REQUIRE((K == 0 || dw) && (M == 0 || (dy && (K == 0 || x))), SGB_ERR_ARG, "linear_wgrad: null tensor");
  SGB_REQUIRE(ldy >= N && ldx >= K && lddw >= K, SGB_ERR_ARG, "linear_wgrad: leading dimension too small");
  SGB_REQUIRE(ws && ws_bytes >= sgb_linear_wgrad_workspace_bytes(M, N, K), SGB_ERR_WORKSPACE, "linear_wgrad: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (K > 0) {
    if (skinny_linear_wgrad_ok(dy, ldy, x, ldx, M, N, K))   // K <= 16: dw and db in one streaming pass
      return skinny_linear_wgrad(dy, ldy, x, ldx, M, N, K, dw, lddw, db, accumulate, ws, st);
    if (tc_linear_wgrad_ok(dy, ldy, x, ldx, M, N, K))   // the tensor-core kernel also produces db (fused column sums)
      return tc_linear_wgrad(dy, ldy, x, ldx, M, N, K, dw, lddw, db, accumulate, ws, st);
    rc = simt_linear_wgrad(dy, ldy, x, ldx, M, N, K, dw, lddw, accumulate, ws, st);
    if (rc != SGB_OK) return rc;
  }
  if (db) rc = linear_colsum(dy, ldy, M, N, db, accumulate, static_cast<char*>(ws) + wgrad_gemm_ws(M, N, K), st);
  return rc;
}
