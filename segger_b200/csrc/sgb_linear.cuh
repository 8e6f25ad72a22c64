// Internal GEMM back ends behind sgb_linear_{fwd,dgrad,wgrad}.
#pragma once
#include "sgb_common.cuh"

namespace sgb {

// exact fp32 SIMT path (sgb_linear_simt.cu)
int simt_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* b, int64_t M, int64_t N,
                    int64_t K, float* y, int64_t ldy, int act, float* y_act, int64_t ldya, cudaStream_t stream);
int simt_linear_dgrad(const float* dy, int64_t ldy, const float* w, int64_t ldw, int64_t M, int64_t N, int64_t K,
                      float* dx, int64_t ldx, int accumulate, int act, const float* act_pre, int64_t ld_pre,
                      cudaStream_t stream);
size_t simt_linear_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K);
int simt_linear_wgrad(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K,
                      float* dw, int64_t lddw, int accumulate, void* ws, cudaStream_t stream);
size_t linear_colsum_workspace_bytes(int64_t M, int64_t N);
int linear_colsum(const float* dy, int64_t ldy, int64_t M, int64_t N, float* db, int accumulate, void* ws,
                  cudaStream_t stream);

// tall-and-skinny fp32 path for reduction depth K <= 16 (sgb_linear_skinny.cu)
bool skinny_linear_fwd_ok(const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K, const float* y, int64_t ldy,
                          const float* y_act, int64_t ldya);
int skinny_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* b, int64_t M, int64_t N,
                      int64_t K, float* y, int64_t ldy, int act, float* y_act, int64_t ldya, cudaStream_t stream);
bool skinny_linear_wgrad_ok(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K);
size_t skinny_linear_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K);
int skinny_linear_wgrad(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K,
                        float* dw, int64_t lddw, float* db, int accumulate, void* ws, cudaStream_t stream);

// tcgen05 3xTF32 path (sgb_linear_tc.cu)
bool tc_enabled();
bool tc_linear_fwd_ok(const float* x, int64_t ldx, const float* w, int64_t ldw, int64_t M, int64_t N, int64_t K);
size_t tc_linear_workspace_bytes(int64_t N, int64_t K);   // packed (split + swizzled) weights of one fwd / dgrad call
int tc_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* b, int64_t M, int64_t N, int64_t K,
                  float* y, int64_t ldy, int act, float* y_act, int64_t ldya, int exact, void* ws, cudaStream_t stream,
                  const void* gids = nullptr, int gid_bytes = 0, const float* gtab = nullptr, int64_t ld_gtab = 0);
bool tc_linear_dgrad_ok(const float* dy, int64_t ldy, const float* w, int64_t ldw, int64_t M, int64_t N, int64_t K);
int tc_linear_dgrad(const float* dy, int64_t ldy, const float* w, int64_t ldw, int64_t M, int64_t N, int64_t K, float* dx,
                    int64_t ldx, int accumulate, int act, const float* act_pre, int64_t ld_pre, void* ws, cudaStream_t stream);
bool tc_linear_wgrad_ok(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K);
size_t tc_linear_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K);
int tc_linear_wgrad(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K, float* dw,
                    int64_t lddw, float* db, int accumulate, void* ws, cudaStream_t stream);

}  // namespace sgb
