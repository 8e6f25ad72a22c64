// tx <-> cell similarity scoring fused with the per-transcript max / arg-max and the cell-id lookup.
// Replaces, in LitISTEncoder.predict_step (/root/reference/src/segger/models/lightning_model.py:275-293):
//   sim = torch.cosine_similarity(emb_tx[src], emb_bd[dst])      (two [E,D] gathers materialised)
//   max_sim, max_idx = torch_scatter.scatter_max(sim, src, dim_size=N_tx)   (atomicMax + arg pass)
//   seg_idx[valid] = bd.index[dst[max_idx[valid]]]
// with one pass over a transcript-sorted candidate CSR: a group of G lanes owns one transcript, keeps
// its (normalised) embedding in registers and streams its 0..3 candidate cell rows (L2-resident).
// cosine_similarity follows ATen: sum((x1 / max(|x1|, eps)) * (x2 / max(|x2|, eps))).
// Ties -> lowest original edge id (stable CSR + strict '>').  No candidates -> (0, E, -1).
#include "sgb_api_internal.cuh"

namespace sgb {
namespace {

struct ScoreParams {
  const float *emb_tx, *emb_bd;
  int64_t ld_tx, ld_bd;
  int D;
  const int32_t *rowptr, *col, *eid;
  int64_t n_tx, E;
  float eps;
  const void* bd_index;
  int bd_index_bytes;
  float min_sim;
  int use_min;
  float* max_sim;
  int64_t *arg_edge, *seg_idx;
};

template <int G>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int o = G / 2; o >= 1; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
  return x;
}

__device__ __forceinline__ void finish(const ScoreParams& p, int64_t row, float best, int best_pos) {
  int64_t arg = p.E, seg = -1;
  float out = 0.f;
  if (best_pos >= 0) {
    arg = p.eid ? static_cast<int64_t>(p.eid[best_pos]) : static_cast<int64_t>(best_pos);
    out = best;
    bool valid = true;
    if (p.use_min) valid = best >= p.min_sim;
    if (valid) {
      const int c = p.col[best_pos];
      if (p.bd_index == nullptr) seg = c;
      else if (p.bd_index_bytes == 8) seg = static_cast<const int64_t*>(p.bd_index)[c];
      else seg = static_cast<const int32_t*>(p.bd_index)[c];
    }
  }
  p.max_sim[row] = out;
  if (p.arg_edge) p.arg_edge[row] = arg;
  if (p.seg_idx) p.seg_idx[row] = seg;
}

// vector path: D = 4 * G * T floats, G lanes per transcript
template <int G, int T>
__global__ void __launch_bounds__(256) score_vec_kernel(const ScoreParams p) {
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;
  const int64_t row = ((static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * (32 / G) + lane / G;
  const bool live = row < p.n_tx;
  const int beg = live ? p.rowptr[row] : 0, end = live ? p.rowptr[row + 1] : 0;
  // all lanes of a warp must take part in the shuffles: iterate to the warp-wide max degree
  int deg = end - beg;
  int maxdeg = deg;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(kFull, maxdeg, o));
  if (maxdeg == 0) {
    if (live && gl == 0) finish(p, row, 0.f, -1);
    return;
  }
  float4 a[T];
  float ssa = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    a[t] = live ? ldg4(p.emb_tx + row * p.ld_tx + (gl + G * t) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    ssa += a[t].x * a[t].x + a[t].y * a[t].y + a[t].z * a[t].z + a[t].w * a[t].w;
  }
  ssa = group_sum<G>(ssa);
  const float na = fmaxf(sqrtf(ssa), p.eps);
#pragma unroll
  for (int t = 0; t < T; ++t) { a[t].x /= na; a[t].y /= na; a[t].z /= na; a[t].w /= na; }
  float best = -INFINITY;
  int best_pos = -1;
  for (int k = 0; k < maxdeg; ++k) {
    const bool has = k < deg;
    float4 b[T];
    float ssb = 0.f;
    const int c = has ? __ldg(p.col + beg + k) : 0;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      b[t] = has ? ldg4(p.emb_bd + static_cast<int64_t>(c) * p.ld_bd + (gl + G * t) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      ssb += b[t].x * b[t].x + b[t].y * b[t].y + b[t].z * b[t].z + b[t].w * b[t].w;
    }
    ssb = group_sum<G>(ssb);
    const float nb = fmaxf(sqrtf(ssb), p.eps);
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t)
      s += a[t].x * (b[t].x / nb) + a[t].y * (b[t].y / nb) + a[t].z * (b[t].z / nb) + a[t].w * (b[t].w / nb);
    s = group_sum<G>(s);
    if (has && (best_pos < 0 || s > best)) { best = s; best_pos = beg + k; }
  }
  if (live && gl == 0) finish(p, row, best, best_pos);
}

// generic path: one warp per transcript, scalar, any D
__global__ void __launch_bounds__(256) score_gen_kernel(const ScoreParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= p.n_tx) return;
  const int beg = p.rowptr[row], end = p.rowptr[row + 1];
  if (beg == end) {
    if (lane == 0) finish(p, row, 0.f, -1);
    return;
  }
  float ssa = 0.f;
  for (int c = lane; c < p.D; c += 32) { const float v = p.emb_tx[row * p.ld_tx + c]; ssa += v * v; }
  const float na = fmaxf(sqrtf(group_sum<32>(ssa)), p.eps);
  float best = -INFINITY;
  int best_pos = -1;
  for (int k = beg; k < end; ++k) {
    const float* b = p.emb_bd + static_cast<int64_t>(p.col[k]) * p.ld_bd;
    float ssb = 0.f;
    for (int c = lane; c < p.D; c += 32) { const float v = __ldg(b + c); ssb += v * v; }
    const float nb = fmaxf(sqrtf(group_sum<32>(ssb)), p.eps);
    float s = 0.f;
    for (int c = lane; c < p.D; c += 32) s += (p.emb_tx[row * p.ld_tx + c] / na) * (__ldg(b + c) / nb);
    s = group_sum<32>(s);
    if (best_pos < 0 || s > best) { best = s; best_pos = k; }
  }
  if (lane == 0) finish(p, row, best, best_pos);
}

}  // namespace
}  // namespace sgb

using namespace sgb;

extern "C" int sgb_score_argmax(const float* emb_tx, int64_t ld_tx, const float* emb_bd, int64_t ld_bd, int D,
                                const int32_t* cand_rowptr, const int32_t* cand_col, const int32_t* cand_eid,
                                int64_t n_tx, int64_t E, float eps, const void* bd_index, int bd_index_bytes,
                                float min_similarity, float* max_sim, int64_t* arg_edge, int64_t* seg_idx,
                                void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(n_tx >= 0 && n_tx < (int64_t(1) << 31) && E >= 0 && E < (int64_t(1) << 31), SGB_ERR_RANGE, "score_argmax: size out of range");
  SGB_REQUIRE(D >= 1, SGB_ERR_ARG, "score_argmax: D must be >= 1");
  if (n_tx == 0) return SGB_OK;
  SGB_REQUIRE(emb_tx && cand_rowptr && max_sim && (E == 0 || (emb_bd && cand_col)), SGB_ERR_ARG, "score_argmax: null tensor");
  SGB_REQUIRE(!bd_index || bd_index_bytes == 4 || bd_index_bytes == 8, SGB_ERR_ARG, "score_argmax: bd_index_bytes must be 4 or 8");
  ScoreParams p{};
  p.emb_tx = emb_tx; p.emb_bd = emb_bd; p.ld_tx = ld_tx; p.ld_bd = ld_bd; p.D = D;
  p.rowptr = cand_rowptr; p.col = cand_col; p.eid = cand_eid; p.n_tx = n_tx; p.E = E; p.eps = eps;
  p.bd_index = bd_index; p.bd_index_bytes = bd_index_bytes;
  p.use_min = (min_similarity == min_similarity) ? 1 : 0;  // NaN disables
  p.min_sim = min_similarity;
  p.max_sim = max_sim; p.arg_edge = arg_edge; p.seg_idx = seg_idx;
  const bool vec = D % 4 == 0 && aligned16(emb_tx) && aligned16(emb_bd) && ld_tx % 4 == 0 && ld_bd % 4 == 0;
  const int q = D / 4;
  if (vec && q == 16) {
    score_vec_kernel<16, 1><<<static_cast<unsigned>(ceil_div(n_tx, 16)), 256, 0, stream>>>(p);
  } else if (vec && q == 8) {
    score_vec_kernel<8, 1><<<static_cast<unsigned>(ceil_div(n_tx, 32)), 256, 0, stream>>>(p);
  } else if (vec && q == 32) {
    score_vec_kernel<32, 1><<<static_cast<unsigned>(ceil_div(n_tx, 8)), 256, 0, stream>>>(p);
  } else if (vec && q == 64) {
    score_vec_kernel<32, 2><<<static_cast<unsigned>(ceil_div(n_tx, 8)), 256, 0, stream>>>(p);
  } else {
    score_gen_kernel<<<static_cast<unsigned>(ceil_div(n_tx, 8)), 256, 0, stream>>>(p);
  }
  return check_launch("score_argmax");
}
