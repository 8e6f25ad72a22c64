// Training losses of segger: triplet sampling, triplet-margin loss, metric (cosine / MSE) loss and the
// segmentation BCE, each as a fused gather + reduce kernel with a deterministic (fixed-order) mean and a backward that
// writes per-triplet row gradients (the caller segment-sums them by target row with sgb_embedding_bwd: no atomics).
//
// Replaces, on the `segger segment` training path,
//   FastTripletSelector.sample_triplets   /root/reference/src/segger/models/triplet_loss.py:88-125
//   TripletLoss / TripletMarginLoss       triplet_loss.py:128-160, lightning_model.py:116,181-186
//   MetricLoss                            triplet_loss.py:163-204
//   BCEWithLogitsLoss over tx.bd logits   lightning_model.py:188-205
#include "sgb_api_internal.cuh"

namespace sgb {
namespace {

constexpr int kLossThreads = 256;
constexpr int kLossWarps = kLossThreads / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// first index in [0, n) with cdf[idx] >= u (torch.searchsorted, right=False); n if none
__device__ __forceinline__ int lower_bound_f32(const float* __restrict__ cdf, int n, float u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(cdf + mid) < u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void triplet_sample_kernel(const int64_t* __restrict__ labels, int64_t N, int C, int P,
                                      const int64_t* __restrict__ present_idx, const int64_t* __restrict__ present,
                                      const int64_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                                      const int64_t* __restrict__ sorted_idx, const float* __restrict__ cdf_pos,
                                      const float* __restrict__ cdf_neg, const float* __restrict__ similarity,
                                      const float* __restrict__ u_pos, const float* __restrict__ u2,
                                      const float* __restrict__ u_neg, const float* __restrict__ u3,
                                      int64_t* __restrict__ positives, int64_t* __restrict__ negatives,
                                      float* __restrict__ dists_pos, float* __restrict__ dists_neg) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t lab = labels[i];
  const int64_t row = present_idx[lab];
  // positive: cluster by the similarity CDF row of the anchor's cluster, member uniformly inside it
  int pp = lower_bound_f32(cdf_pos + row * P, P, u_pos[i]);
  pp = pp < P ? pp : P - 1;
  const int64_t pc = present[pp];
  const int64_t ppos = static_cast<int64_t>(floorf(u2[i] * static_cast<float>(counts[pc])));
  const int64_t pos = sorted_idx[offsets[pc] + ppos];
  int np = lower_bound_f32(cdf_neg + row * P, P, u_neg[i]);
  np = np < P ? np : P - 1;
  const int64_t nc = present[np];
  const int64_t npos = static_cast<int64_t>(floorf(u3[i] * static_cast<float>(counts[nc])));
  const int64_t neg = sorted_idx[offsets[nc] + npos];
  positives[i] = pos;
  negatives[i] = neg;
  dists_pos[i] = 1.0f - similarity[lab * C + labels[pos]];
  dists_neg[i] = 1.0f - similarity[lab * C + labels[neg]];
}

__device__ __forceinline__ const float* row_ptr(const float* __restrict__ t, int64_t ld, const int64_t* __restrict__ idx, int64_t i) {
  return t + (idx ? idx[i] : i) * ld;
}

// block partial sums of per-item losses, fixed order inside the block
__device__ __forceinline__ void block_partial(float item_loss, bool lane0, float* __restrict__ partial) {
  __shared__ float s[kLossWarps];
  const int warp = threadIdx.x >> 5;
  if (lane0) s[warp] = item_loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kLossWarps; ++w) t += s[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024)
mean_reduce_kernel(const float* __restrict__ partial, int64_t n, float inv_count, float* __restrict__ out) {
  // one CTA: fixed strided accumulation per thread, ordered butterfly per warp, ordered sum over the 32 warps:
  // deterministic for a given n
  __shared__ float s[32];
  float t = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) t += partial[i];
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) a += s[w];
    *out = a * inv_count;
  }
}

// warp per triplet: d(x, y) = || x - y + eps ||_2 (torch.pairwise_distance), loss = max(margin + d_ap - d_an, 0)
__global__ void __launch_bounds__(kLossThreads)
triplet_fwd_kernel(const float* __restrict__ ta, int64_t lda, const int64_t* __restrict__ ia, const float* __restrict__ tp,
                   int64_t ldp, const int64_t* __restrict__ ip, const float* __restrict__ tn, int64_t ldn,
                   const int64_t* __restrict__ in_, int64_t T, int D, float margin, float eps, float* __restrict__ d_ap,
                   float* __restrict__ d_an, float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  float loss = 0.f;
  if (i < T) {
    const float* a = row_ptr(ta, lda, ia, i);
    const float* p = row_ptr(tp, ldp, ip, i);
    const float* n = row_ptr(tn, ldn, in_, i);
    float sp = 0.f, sn = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float av = a[c];
      const float dp = av - p[c] + eps, dn = av - n[c] + eps;
      sp = fmaf(dp, dp, sp);
      sn = fmaf(dn, dn, sn);
    }
    const float dap = sqrtf(warp_sum(sp)), dan = sqrtf(warp_sum(sn));
    if (lane == 0) { d_ap[i] = dap; d_an[i] = dan; }
    loss = fmaxf(margin + dap - dan, 0.f);
  }
  block_partial(loss, lane == 0, partial);
}

__global__ void __launch_bounds__(kLossThreads)
triplet_bwd_kernel(const float* __restrict__ ta, int64_t lda, const int64_t* __restrict__ ia, const float* __restrict__ tp,
                   int64_t ldp, const int64_t* __restrict__ ip, const float* __restrict__ tn, int64_t ldn,
                   const int64_t* __restrict__ in_, int64_t T, int D, float margin, float eps, const float* __restrict__ d_ap,
                   const float* __restrict__ d_an, const float* __restrict__ grad, float* __restrict__ ga,
                   float* __restrict__ gp, float* __restrict__ gn) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (i >= T) return;
  const float dap = d_ap[i], dan = d_an[i];
  const bool active = margin + dap - dan > 0.f;
  const float g = active ? __ldg(grad) / static_cast<float>(T) : 0.f;
  const float cp = dap > 0.f ? g / dap : 0.f, cn = dan > 0.f ? g / dan : 0.f;
  const float* a = row_ptr(ta, lda, ia, i);
  const float* p = row_ptr(tp, ldp, ip, i);
  const float* n = row_ptr(tn, ldn, in_, i);
  for (int c = lane; c < D; c += 32) {
    const float av = a[c];
    const float up = (av - p[c] + eps) * cp, un = (av - n[c] + eps) * cn;
    ga[i * D + c] = up - un;
    gp[i * D + c] = -up;
    gn[i * D + c] = un;
  }
}

// 128-bit variants of the two kernels above (D in {32, 64, 128}, 16-byte aligned rows): LPR = D/4 lanes per triplet, a
// warp handles 32/LPR triplets.  Same per-element arithmetic; the squared distances are summed four elements per lane and
// then across the LPR lanes (a different -- still fixed -- order than the 32-bit kernel's).
template <int LPR>
__global__ void __launch_bounds__(kLossThreads)
triplet_fwd_vec_kernel(const float* __restrict__ ta, int64_t lda, const int64_t* __restrict__ ia, const float* __restrict__ tp,
                       int64_t ldp, const int64_t* __restrict__ ip, const float* __restrict__ tn, int64_t ldn,
                       const int64_t* __restrict__ in_, int64_t T, float margin, float eps, float* __restrict__ d_ap,
                       float* __restrict__ d_an, float* __restrict__ partial) {
  constexpr int G = 32 / LPR;
  const int lane = threadIdx.x & 31, s = lane % LPR, gq = lane / LPR;
  const int64_t i = ((static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * G + gq;
  float sp = 0.f, sn = 0.f;
  if (i < T) {
    const float4 a = ldg4(row_ptr(ta, lda, ia, i) + 4 * s);
    const float4 p = ldg4(row_ptr(tp, ldp, ip, i) + 4 * s);
    const float4 n = ldg4(row_ptr(tn, ldn, in_, i) + 4 * s);
    float d;
    d = a.x - p.x + eps; sp = fmaf(d, d, sp); d = a.y - p.y + eps; sp = fmaf(d, d, sp);
    d = a.z - p.z + eps; sp = fmaf(d, d, sp); d = a.w - p.w + eps; sp = fmaf(d, d, sp);
    d = a.x - n.x + eps; sn = fmaf(d, d, sn); d = a.y - n.y + eps; sn = fmaf(d, d, sn);
    d = a.z - n.z + eps; sn = fmaf(d, d, sn); d = a.w - n.w + eps; sn = fmaf(d, d, sn);
  }
#pragma unroll
  for (int o = LPR / 2; o >= 1; o >>= 1) {
    sp += __shfl_xor_sync(kFull, sp, o);
    sn += __shfl_xor_sync(kFull, sn, o);
  }
  float loss = 0.f;
  if (i < T) {
    const float dap = sqrtf(sp), dan = sqrtf(sn);
    if (s == 0) { d_ap[i] = dap; d_an[i] = dan; }
    loss = fmaxf(margin + dap - dan, 0.f);
  }
  float wl = 0.f;      // the warp's triplets, in order
#pragma unroll
  for (int g2 = 0; g2 < G; ++g2) wl += __shfl_sync(kFull, loss, g2 * LPR);
  block_partial(wl, lane == 0, partial);
}

template <int LPR>
__global__ void __launch_bounds__(kLossThreads)
triplet_bwd_vec_kernel(const float* __restrict__ ta, int64_t lda, const int64_t* __restrict__ ia, const float* __restrict__ tp,
                       int64_t ldp, const int64_t* __restrict__ ip, const float* __restrict__ tn, int64_t ldn,
                       const int64_t* __restrict__ in_, int64_t T, float margin, float eps, const float* __restrict__ d_ap,
                       const float* __restrict__ d_an, const float* __restrict__ grad, float* __restrict__ ga,
                       float* __restrict__ gp, float* __restrict__ gn) {
  constexpr int D = 4 * LPR;
  const int s = threadIdx.x % LPR;
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LPR;
  if (i >= T) return;
  const float dap = d_ap[i], dan = d_an[i];
  const bool active = margin + dap - dan > 0.f;
  const float g = active ? __ldg(grad) / static_cast<float>(T) : 0.f;
  const float cp = dap > 0.f ? g / dap : 0.f, cn = dan > 0.f ? g / dan : 0.f;
  const float4 a = ldg4(row_ptr(ta, lda, ia, i) + 4 * s);
  const float4 p = ldg4(row_ptr(tp, ldp, ip, i) + 4 * s);
  const float4 n = ldg4(row_ptr(tn, ldn, in_, i) + 4 * s);
  const float4 up = make_float4((a.x - p.x + eps) * cp, (a.y - p.y + eps) * cp, (a.z - p.z + eps) * cp, (a.w - p.w + eps) * cp);
  const float4 un = make_float4((a.x - n.x + eps) * cn, (a.y - n.y + eps) * cn, (a.z - n.z + eps) * cn, (a.w - n.w + eps) * cn);
  st4(ga + i * D + 4 * s, make_float4(up.x - un.x, up.y - un.y, up.z - un.z, up.w - un.w));
  st4(gp + i * D + 4 * s, make_float4(-up.x, -up.y, -up.z, -up.w));
  st4(gn + i * D + 4 * s, un);
}

// Row-owner backward of the triplet loss when anchors, positives and negatives are rows of ONE table and triplet t has
// anchor row t (TripletLoss over an embedding matrix): LPR = D/4 lanes own row r and sum, in a fixed order,
//   the anchor term of triplet r, then -u_p(t) for every triplet t that sampled r as its positive (CSR over ip, t
//   increasing), then +u_n(t) for every t that sampled r as its negative,
// each term computed exactly as triplet_bwd_kernel writes it -- but nothing of width D is materialised per triplet, sorted
// or segment-summed: per row ~2 + |P(r)| + |N(r)| gathered rows instead of 3 written + 3 re-read + 2 sorted.
template <int LPR>
__global__ void __launch_bounds__(kLossThreads)
triplet_self_bwd_kernel(const float* __restrict__ emb, int64_t ld, const int64_t* __restrict__ ip,
                        const int64_t* __restrict__ in_, int64_t T, float margin, float eps,
                        const float* __restrict__ d_ap, const float* __restrict__ d_an, const float* __restrict__ grad,
                        const int32_t* __restrict__ p_rowptr, const int32_t* __restrict__ p_tid,
                        const int32_t* __restrict__ n_rowptr, const int32_t* __restrict__ n_tid,
                        float* __restrict__ gout, int64_t ldg) {
  const int s = threadIdx.x % LPR;
  const int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LPR;
  if (r >= T) return;
  const float g0 = __ldg(grad) / static_cast<float>(T);
  // coefficient pair of triplet t: (g / d_ap, g / d_an) if the hinge is active, else 0 (triplet_bwd_kernel's rule)
  auto coef = [&](int64_t t, float& cp, float& cn) {
    const float dap = __ldg(d_ap + t), dan = __ldg(d_an + t);
    const float g = (margin + dap - dan > 0.f) ? g0 : 0.f;
    cp = dap > 0.f ? g / dap : 0.f;
    cn = dan > 0.f ? g / dan : 0.f;
  };
  const float4 e = ldg4(emb + r * ld + 4 * s);
  float4 acc;
  {
    float cp, cn;
    coef(r, cp, cn);
    const float4 p = ldg4(emb + __ldg(ip + r) * ld + 4 * s);
    const float4 n = ldg4(emb + __ldg(in_ + r) * ld + 4 * s);
    acc.x = (e.x - p.x + eps) * cp - (e.x - n.x + eps) * cn;
    acc.y = (e.y - p.y + eps) * cp - (e.y - n.y + eps) * cn;
    acc.z = (e.z - p.z + eps) * cp - (e.z - n.z + eps) * cn;
    acc.w = (e.w - p.w + eps) * cp - (e.w - n.w + eps) * cn;
  }
  // a row that many triplets sampled (a small cluster next to large ones) sums thousands of terms: compensated (Kahan)
  // accumulation keeps the left fold as accurate as the chunked segment sums it replaces, at 3 adds per term
  float4 cmp = make_float4(0.f, 0.f, 0.f, 0.f);
  auto kadd = [](float& sum, float& c, float term) {
    const float y = __fsub_rn(term, c);
    const float tsum = __fadd_rn(sum, y);
    c = __fsub_rn(__fsub_rn(tsum, sum), y);
    sum = tsum;
  };
  for (int k = __ldg(p_rowptr + r), ke = __ldg(p_rowptr + r + 1); k < ke; ++k) {
    const int64_t t = __ldg(p_tid + k);
    float cp, cn;
    coef(t, cp, cn);
    const float4 a = ldg4(emb + t * ld + 4 * s);
    kadd(acc.x, cmp.x, -((a.x - e.x + eps) * cp)); kadd(acc.y, cmp.y, -((a.y - e.y + eps) * cp));
    kadd(acc.z, cmp.z, -((a.z - e.z + eps) * cp)); kadd(acc.w, cmp.w, -((a.w - e.w + eps) * cp));
  }
  for (int k = __ldg(n_rowptr + r), ke = __ldg(n_rowptr + r + 1); k < ke; ++k) {
    const int64_t t = __ldg(n_tid + k);
    float cp, cn;
    coef(t, cp, cn);
    const float4 a = ldg4(emb + t * ld + 4 * s);
    kadd(acc.x, cmp.x, (a.x - e.x + eps) * cn); kadd(acc.y, cmp.y, (a.y - e.y + eps) * cn);
    kadd(acc.z, cmp.z, (a.z - e.z + eps) * cn); kadd(acc.w, cmp.w, (a.w - e.w + eps) * cn);
  }
  st4(gout + r * ldg + 4 * s, acc);
}

// pair losses: mode 0 = mse(cosine_similarity(a, b), target) (ATen formula: sum (a / max(|a|, eps)) (b / max(|b|, eps)));
//              mode 1 = binary_cross_entropy_with_logits(a . b, target)
__global__ void __launch_bounds__(kLossThreads)
pair_fwd_kernel(const float* __restrict__ ta, int64_t lda, const int64_t* __restrict__ ia, const float* __restrict__ tb,
                int64_t ldb, const int64_t* __restrict__ ib, const float* __restrict__ target, int64_t T, int D, int mode,
                float eps, float* __restrict__ val, float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  float loss = 0.f;
  if (i < T) {
    const float* a = row_ptr(ta, lda, ia, i);
    const float* b = row_ptr(tb, ldb, ib, i);
    float saa = 0.f, sbb = 0.f, sab = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float av = a[c], bv = b[c];
      saa = fmaf(av, av, saa); sbb = fmaf(bv, bv, sbb); sab = fmaf(av, bv, sab);
    }
    saa = warp_sum(saa); sbb = warp_sum(sbb); sab = warp_sum(sab);
    const float t = target[i];
    float v;
    if (mode == 0) {
      v = sab / (fmaxf(sqrtf(saa), eps) * fmaxf(sqrtf(sbb), eps));
      loss = (v - t) * (v - t);
    } else {
      v = sab;
      loss = fmaxf(v, 0.f) - v * t + log1pf(expf(-fabsf(v)));
    }
    if (lane == 0) val[i] = v;
  }
  block_partial(loss, lane == 0, partial);
}

__global__ void __launch_bounds__(kLossThreads)
pair_bwd_kernel(const float* __restrict__ ta, int64_t lda, const int64_t* __restrict__ ia, const float* __restrict__ tb,
                int64_t ldb, const int64_t* __restrict__ ib, const float* __restrict__ target, int64_t T, int D, int mode,
                float eps, const float* __restrict__ val, const float* __restrict__ grad, float* __restrict__ gA,
                float* __restrict__ gB) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (i >= T) return;
  const float* a = row_ptr(ta, lda, ia, i);
  const float* b = row_ptr(tb, ldb, ib, i);
  const float g = __ldg(grad) / static_cast<float>(T);
  const float v = val[i], t = target[i];
  if (mode == 1) {
    const float dv = g * (1.0f / (1.0f + expf(-v)) - t);
    for (int c = lane; c < D; c += 32) {
      gA[i * D + c] = dv * b[c];
      gB[i * D + c] = dv * a[c];
    }
    return;
  }
  float saa = 0.f, sbb = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float av = a[c], bv = b[c];
    saa = fmaf(av, av, saa); sbb = fmaf(bv, bv, sbb);
  }
  const float ra = sqrtf(warp_sum(saa)), rb = sqrtf(warp_sum(sbb));
  const float na = fmaxf(ra, eps), nb = fmaxf(rb, eps);
  const bool fa = ra > eps, fb = rb > eps;          // norm not clamped: it depends on the row
  const float dv = g * 2.0f * (v - t);
  for (int c = lane; c < D; c += 32) {
    const float ah = a[c] / na, bh = b[c] / nb;
    gA[i * D + c] = dv * (bh - (fa ? v * ah : 0.f)) / na;
    gB[i * D + c] = dv * (ah - (fb ? v * bh : 0.f)) / nb;
  }
}

// the 128-bit triplet kernels: D in {32, 64, 128}, rows of all three tables 16-byte aligned
bool triplet_vec_ok(int D, const float* ta, int64_t lda, const float* tp, int64_t ldp, const float* tn, int64_t ldn) {
  return (D == 32 || D == 64 || D == 128) && aligned16(ta) && aligned16(tp) && aligned16(tn) && lda % 4 == 0 && ldp % 4 == 0 &&
         ldn % 4 == 0;
}
unsigned warp_blocks(int64_t T) { return static_cast<unsigned>(ceil_div(T > 0 ? T : 1, kLossWarps)); }

}  // namespace
}  // namespace sgb

using namespace sgb;

extern "C" int sgb_triplet_sample(const int64_t* labels, int64_t N, int C, int P, const int64_t* present_idx,
                                  const int64_t* present, const int64_t* counts, const int64_t* offsets,
                                  const int64_t* sorted_idx, const float* cdf_pos, const float* cdf_neg,
                                  const float* similarity, const float* u_pos, const float* u2, const float* u_neg,
                                  const float* u3, int64_t* positives, int64_t* negatives, float* dists_pos,
                                  float* dists_neg, void* stream) {
  SGB_REQUIRE(N >= 0 && C >= 1 && P >= 1 && P <= C, SGB_ERR_ARG, "triplet_sample: bad size");
  if (N == 0) return SGB_OK;
  SGB_REQUIRE(labels && present_idx && present && counts && offsets && sorted_idx && cdf_pos && cdf_neg && similarity && u_pos &&
                  u2 && u_neg && u3 && positives && negatives && dists_pos && dists_neg, SGB_ERR_ARG, "triplet_sample: null tensor");
  triplet_sample_kernel<<<static_cast<unsigned>(ceil_div(N, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      labels, N, C, P, present_idx, present, counts, offsets, sorted_idx, cdf_pos, cdf_neg, similarity, u_pos, u2, u_neg, u3,
      positives, negatives, dists_pos, dists_neg);
  return check_launch("triplet_sample");
}

extern "C" size_t sgb_loss_workspace_bytes(int64_t T) { return align_up(static_cast<size_t>(warp_blocks(T)) * sizeof(float)); }

extern "C" int sgb_triplet_margin_fwd(const float* ta, int64_t lda, const int64_t* ia, const float* tp, int64_t ldp,
                                      const int64_t* ip, const float* tn, int64_t ldn, const int64_t* in_, int64_t T, int D,
                                      float margin, float eps, float* d_ap, float* d_an, float* loss, void* ws,
                                      size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(T >= 0 && D >= 1 && loss, SGB_ERR_ARG, "triplet_margin_fwd: bad argument");
  if (T == 0) { cudaMemsetAsync(loss, 0, sizeof(float), stream); return check_launch("triplet_margin_fwd(empty)"); }
  SGB_REQUIRE(ta && tp && tn && d_ap && d_an && lda >= D && ldp >= D && ldn >= D, SGB_ERR_ARG, "triplet_margin_fwd: null tensor");
  SGB_REQUIRE(ws && ws_bytes >= sgb_loss_workspace_bytes(T), SGB_ERR_WORKSPACE, "triplet_margin_fwd: workspace too small");
  float* partial = static_cast<float*>(ws);
  if (triplet_vec_ok(D, ta, lda, tp, ldp, tn, ldn)) {
    const int lpr = D / 4, per_block = kLossWarps * (32 / lpr);
    const unsigned nbv = static_cast<unsigned>(ceil_div(T, per_block));       // <= warp_blocks(T): the workspace fits
#define SGB_TFV(L) triplet_fwd_vec_kernel<L><<<nbv, kLossThreads, 0, stream>>>(ta, lda, ia, tp, ldp, ip, tn, ldn, in_, T, margin, eps, d_ap, d_an, partial)
    if (lpr == 8) SGB_TFV(8); else if (lpr == 16) SGB_TFV(16); else SGB_TFV(32);
#undef SGB_TFV
    mean_reduce_kernel<<<1, 1024, 0, stream>>>(partial, nbv, 1.0f / static_cast<float>(T), loss);
    return check_launch("triplet_margin_fwd(vec)");
  }
  const unsigned nb = warp_blocks(T);
  triplet_fwd_kernel<<<nb, kLossThreads, 0, stream>>>(ta, lda, ia, tp, ldp, ip, tn, ldn, in_, T, D, margin, eps, d_ap, d_an, partial);
  mean_reduce_kernel<<<1, 1024, 0, stream>>>(partial, nb, 1.0f / static_cast<float>(T), loss);
  return check_launch("triplet_margin_fwd");
}

extern "C" int sgb_triplet_margin_bwd(const float* ta, int64_t lda, const int64_t* ia, const float* tp, int64_t ldp,
                                      const int64_t* ip, const float* tn, int64_t ldn, const int64_t* in_, int64_t T, int D,
                                      float margin, float eps, const float* d_ap, const float* d_an, const float* grad,
                                      float* ga, float* gp, float* gn, void* stream) {
  SGB_REQUIRE(T >= 0 && D >= 1, SGB_ERR_ARG, "triplet_margin_bwd: bad argument");
  if (T == 0) return SGB_OK;
  SGB_REQUIRE(ta && tp && tn && d_ap && d_an && grad && ga && gp && gn, SGB_ERR_ARG, "triplet_margin_bwd: null tensor");
  if (triplet_vec_ok(D, ta, lda, tp, ldp, tn, ldn) && aligned16(ga) && aligned16(gp) && aligned16(gn)) {
    const int lpr = D / 4;
    const unsigned nbv = static_cast<unsigned>(ceil_div(T * lpr, kLossThreads));
#define SGB_TBV(L) triplet_bwd_vec_kernel<L><<<nbv, kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(ta, lda, ia, tp, ldp, ip, tn, ldn, in_, T, margin, eps, d_ap, d_an, grad, ga, gp, gn)
    if (lpr == 8) SGB_TBV(8); else if (lpr == 16) SGB_TBV(16); else SGB_TBV(32);
#undef SGB_TBV
    return check_launch("triplet_margin_bwd(vec)");
  }
  triplet_bwd_kernel<<<warp_blocks(T), kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      ta, lda, ia, tp, ldp, ip, tn, ldn, in_, T, D, margin, eps, d_ap, d_an, grad, ga, gp, gn);
  return check_launch("triplet_margin_bwd");
}

extern "C" int sgb_triplet_self_bwd_supported(int D) { return (D == 32 || D == 64 || D == 128) ? 1 : 0; }

extern "C" int sgb_triplet_self_bwd(const float* emb, int64_t ld, const int64_t* ip, const int64_t* in_, int64_t T, int D,
                                    float margin, float eps, const float* d_ap, const float* d_an, const float* grad,
                                    const int32_t* p_rowptr, const int32_t* p_tid, const int32_t* n_rowptr,
                                    const int32_t* n_tid, float* gout, int64_t ldg, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(T >= 0 && T < (int64_t(1) << 31), SGB_ERR_RANGE, "triplet_self_bwd: T out of range");
  SGB_REQUIRE(sgb_triplet_self_bwd_supported(D), SGB_ERR_ARG, "triplet_self_bwd: D must be 32, 64 or 128 (got %d)", D);
  if (T == 0) return SGB_OK;
  SGB_REQUIRE(emb && ip && in_ && d_ap && d_an && grad && p_rowptr && p_tid && n_rowptr && n_tid && gout, SGB_ERR_ARG,
              "triplet_self_bwd: null tensor");
  SGB_REQUIRE(aligned16(emb) && ld % 4 == 0 && aligned16(gout) && ldg % 4 == 0, SGB_ERR_ALIGN,
              "triplet_self_bwd: rows must be 16-byte aligned");
  const int lpr = D / 4;
  const unsigned blocks = static_cast<unsigned>(ceil_div(T * lpr, kLossThreads));
#define SGB_TSB(L)                                                                                                      \
  triplet_self_bwd_kernel<L><<<blocks, kLossThreads, 0, stream>>>(emb, ld, ip, in_, T, margin, eps, d_ap, d_an, grad, p_rowptr, \
                                                                  p_tid, n_rowptr, n_tid, gout, ldg)
  if (lpr == 8) SGB_TSB(8); else if (lpr == 16) SGB_TSB(16); else SGB_TSB(32);
#undef SGB_TSB
  return check_launch("triplet_self_bwd");
}

extern "C" int sgb_pair_loss_fwd(const float* ta, int64_t lda, const int64_t* ia, const float* tb, int64_t ldb,
                                 const int64_t* ib, const float* target, int64_t T, int D, int mode, float eps, float* val,
                                 float* loss, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(T >= 0 && D >= 1 && loss && (mode == 0 || mode == 1), SGB_ERR_ARG, "pair_loss_fwd: bad argument");
  if (T == 0) { cudaMemsetAsync(loss, 0, sizeof(float), stream); return check_launch("pair_loss_fwd(empty)"); }
  SGB_REQUIRE(ta && tb && target && val && lda >= D && ldb >= D, SGB_ERR_ARG, "pair_loss_fwd: null tensor");
  SGB_REQUIRE(ws && ws_bytes >= sgb_loss_workspace_bytes(T), SGB_ERR_WORKSPACE, "pair_loss_fwd: workspace too small");
  float* partial = static_cast<float*>(ws);
  const unsigned nb = warp_blocks(T);
  pair_fwd_kernel<<<nb, kLossThreads, 0, stream>>>(ta, lda, ia, tb, ldb, ib, target, T, D, mode, eps, val, partial);
  mean_reduce_kernel<<<1, 1024, 0, stream>>>(partial, nb, 1.0f / static_cast<float>(T), loss);
  return check_launch("pair_loss_fwd");
}

extern "C" int sgb_pair_loss_bwd(const float* ta, int64_t lda, const int64_t* ia, const float* tb, int64_t ldb,
                                 const int64_t* ib, const float* target, int64_t T, int D, int mode, float eps,
                                 const float* val, const float* grad, float* gA, float* gB, void* stream) {
  SGB_REQUIRE(T >= 0 && D >= 1 && (mode == 0 || mode == 1), SGB_ERR_ARG, "pair_loss_bwd: bad argument");
  if (T == 0) return SGB_OK;
  SGB_REQUIRE(ta && tb && target && val && grad && gA && gB, SGB_ERR_ARG, "pair_loss_bwd: null tensor");
  pair_bwd_kernel<<<warp_blocks(T), kLossThreads, 0, static_cast<cudaStream_t>(stream)>>>(ta, lda, ia, tb, ldb, ib, target, T, D,
                                                                                       mode, eps, val, grad, gA, gB);
  return check_launch("pair_loss_bwd");
}
