// Shared pieces of the fused GATv2 kernels (sgb_gatv2.cu: row-per-warp / generic paths,
// sgb_gatv2_quad.cu: cp.async-pipelined sub-warp-per-row path).
#pragma once
#include "sgb_api_internal.cuh"

namespace sgb {

__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ float lrelu(float z, float slope) { return z > 0.f ? z : slope * z; }
__device__ __forceinline__ float4 lrelu4(const float4 z, float slope) {
  return make_float4(lrelu(z.x, slope), lrelu(z.y, slope), lrelu(z.z, slope), lrelu(z.w, slope));
}
__device__ __forceinline__ float4 add4(const float4 a, const float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ void fma4(float4& acc, float w, const float4 x) {
  acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y); acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
}
__device__ __forceinline__ void scale4(float4& a, float s) { a.x *= s; a.y *= s; a.z *= s; a.w *= s; }

struct GatParams {
  const float *x_l, *x_r, *att, *bias;
  int64_t ld_l, ld_r;
  const int32_t *rowptr, *col, *eid;
  int64_t n_dst, n_src;
  int H, C;
  float slope;
  int training;
  uint32_t drop_thr;
  float keep_scale;
  uint64_t seed;
  const uint64_t* seed_dev;   // optional device word added to `seed` (CUDA-graph replays draw new masks); see gat_seed()
  // forward outputs / backward saved inputs
  float *out, *out_act;
  int64_t ld_out, ld_act;
  float *stat_max, *stat_den;
  float* e_logit;            // optional [E, H] raw attention logits in dst-CSR order: written by the sub-warp forward,
                             // read by the sub-warp backward (which then skips the att . lrelu(z) recomputation)
  // backward
  const float* grad_out;
  int64_t ld_g;
  int gelu_fused;
  float* g_buf;
  float *e_delta, *e_alpha;  // [E,H] per-edge scalars in dst-CSR order
  float *grad_x_l, *grad_x_r;
  int64_t ld_gl, ld_gr;
  float* partial;            // per-CTA (vector path) / per-warp (generic path) partial sums
  const int32_t *t_rowptr, *t_dst, *t_pos;
};

// Effective dropout seed of a launch: the by-value seed plus, when given, a device-resident word.  A captured CUDA graph
// freezes by-value arguments, so a replayed training step advances the device word instead (segger_b200/graphs.py).
__device__ __forceinline__ uint64_t gat_seed(const GatParams& p) {
  return p.seed_dev ? p.seed + __ldg(reinterpret_cast<const unsigned long long*>(p.seed_dev)) : p.seed;
}



// Per-edge scalar record written by the dst pass for the src pass, in dst-CSR order:
// rec[pos][0..H) = delta, rec[pos][H..2H) = alpha' ; row stride = 2H rounded up to 4 floats.
static inline __host__ __device__ int rec_stride(int H) { return (2 * H + 3) / 4 * 4; }

// sub-warp-per-row path (sgb_gatv2_quad.cu); each returns false if the shape is not covered
bool quad_fwd_launch(const GatParams& p, cudaStream_t stream);
size_t quad_bwd_partial_floats(int H, int C);
size_t quad_bwd_record_bytes(int H, int C);   // bytes of one per-edge record of the sub-warp backward (0: shape not covered)
bool quad_bwd_launch(const GatParams& p, float* grad_att, float* grad_bias, cudaStream_t stream);
bool quad_supported(const GatParams& p);   // shape (H, C) and slope covered by the sub-warp kernels

}  // namespace sgb
