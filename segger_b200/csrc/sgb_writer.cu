// Post-processing of the predictions: what ISTSegmentationWriter.assign_transcripts_to_cells does
// with polars + scikit-image on the host (/root/reference/src/segger/data/writer.py:132-253,
// data/utils/threshold.py:3-11), on the device:
//   * de-duplication: one row per transcript, the one with the highest similarity
//     (`.sort(by=[row_index, similarity], descending=[False, True]).unique(row_index, keep="first")`, writer.py:199-203);
//     exact ties -> lowest cell id (the reference's order among ties is unspecified);
//   * per-gene similarity thresholds min(Yen, Li) over the assigned transcripts of each gene (writer.py:209-240):
//     Yen = scikit-image threshold_yen (256-bin histogram over [min, max], argmax of
//     log((P1_sq * P2_sq)^-1 * (P1 (1 - P1))^2)), Li = threshold_li's iteration
//     t <- (mean_back - mean_fore) / (log mean_back - log mean_fore) from t0 = mean, tolerance = half the smallest gap
//     between distinct values, at most `max_iter` callbacks (threshold_li_custom).  scikit-image is a third-party
//     dependency absent from /root/reference: the published algorithms are restated (fp64 arithmetic), see
//     oracle/writer_ref.py.
// Integer work is exact; the thresholds are floating point (compared with a tolerance in the tests).
#include "sgb_api_internal.cuh"
#include "sgb_sort.cuh"
#include <cfloat>
#include <cmath>

namespace sgb {
namespace {

constexpr int kT = 256;
constexpr int kBins = 256;

inline unsigned nblk(int64_t n) { return static_cast<unsigned>(ceil_div(n > 0 ? n : 1, kT)); }

// ---- dedupe --------------------------------------------------------------------------------------------------------
__global__ void row_keys_kernel(const int64_t* __restrict__ row, const uint32_t* __restrict__ perm, int64_t n, int shift,
                                uint32_t* __restrict__ keys) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t r = row[perm ? perm[i] : i];
  keys[i] = static_cast<uint32_t>(static_cast<uint64_t>(r) >> shift);
}

// thread at the first entry of every run of equal rows picks the winner of the run
__global__ void run_best_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ seg, const float* __restrict__ sim,
                                const uint32_t* __restrict__ perm, int64_t n, int32_t* __restrict__ flags,
                                int32_t* __restrict__ winner) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t r = row[perm[i]];
  const bool start = (i == 0) || (row[perm[i - 1]] != r);
  flags[i] = start ? 1 : 0;
  if (!start) return;
  uint32_t best = perm[i];
  float bs = sim[best];
  int64_t bg = seg[best];
  for (int64_t j = i + 1; j < n; ++j) {
    const uint32_t e = perm[j];
    if (row[e] != r) break;
    const float s = sim[e];
    const int64_t g = seg[e];
    if (s > bs || (s == bs && g < bg)) { best = e; bs = s; bg = g; }
  }
  winner[i] = static_cast<int32_t>(best);
}
__global__ void compact_winners_kernel(const int32_t* __restrict__ flags, const int32_t* __restrict__ scan,
                                       const int32_t* __restrict__ winner, int64_t n, int32_t* __restrict__ order,
                                       int32_t* __restrict__ count) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i == 0) *count = scan[n];
  if (i < n && flags[i]) order[scan[i]] = winner[i];
}

// ---- thresholds ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_order_key(float f) {   // ascending order of floats as unsigned ints
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void sim_keys_kernel(const float* __restrict__ sim, int64_t n, uint32_t* __restrict__ keys) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = float_order_key(sim[i]);
}
template <typename GeneT>
__global__ void gene_keys_kernel(const GeneT* __restrict__ gene, const int64_t* __restrict__ seg, const uint32_t* __restrict__ perm,
                                 int64_t n, int n_genes, uint32_t* __restrict__ keys) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t e = perm[i];
  const int64_t g = static_cast<int64_t>(gene[e]);
  keys[i] = (seg[e] >= 0 && g >= 0 && g < n_genes) ? static_cast<uint32_t>(g) : static_cast<uint32_t>(n_genes);
}
__global__ void gather_f32_kernel(const float* __restrict__ src, const uint32_t* __restrict__ perm, int64_t n, float* __restrict__ dst) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

__device__ double block_sum(double v, double* red) {
  __syncthreads();
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kT / 32; ++w) t += red[w];   // fixed order
  return t;
}

// one CTA per gene; v = ascending similarities of the gene's assigned transcripts
__global__ void __launch_bounds__(kT) gene_threshold_kernel(const float* __restrict__ vals, const int32_t* __restrict__ rowptr,
                                                            int n_genes, int max_iter, double* __restrict__ thr_yen,
                                                            double* __restrict__ thr_li, int32_t* __restrict__ li_iters,
                                                            int32_t* __restrict__ counts) {
  __shared__ int hist[kBins];
  __shared__ double red[kT / 32];
  __shared__ float s_tol;
  const int g = blockIdx.x;
  const int b = rowptr[g], e = rowptr[g + 1];
  const int n = e - b;
  const float* v = vals + b;
  if (threadIdx.x == 0) counts[g] = n;
  if (n == 0) {
    if (threadIdx.x == 0) { thr_yen[g] = nan(""); thr_li[g] = nan(""); li_iters[g] = 0; }
    return;
  }
  const float vmin = v[0], vmax = v[n - 1];
  if (!(vmax > vmin)) {                    // all values equal: both thresholds are that value
    if (threadIdx.x == 0) { thr_yen[g] = vmin; thr_li[g] = vmin; li_iters[g] = 0; }
    return;
  }
  // ---- Yen: 256-bin histogram over [vmin, vmax] (numpy.histogram: right edge closed)
  for (int i = threadIdx.x; i < kBins; i += kT) hist[i] = 0;
  __syncthreads();
  const double lo = vmin, hi = vmax, scale = static_cast<double>(kBins) / (hi - lo), width = (hi - lo) / kBins;
  for (int i = threadIdx.x; i < n; i += kT) {
    const double x = v[i];
    int k = static_cast<int>((x - lo) * scale);
    if (k >= kBins) k = kBins - 1;
    // numpy corrects the floor against the actual edges lo + k * width
    if (k > 0 && x < lo + k * width) --k;
    else if (k < kBins - 1 && x >= lo + (k + 1) * width) ++k;
    atomicAdd(&hist[k], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // P1 = cumsum(pmf), P1_sq = cumsum(pmf^2), P2_sq = reversed cumsum of pmf^2 from the end
    double p1 = 0.0, p1sq = 0.0;
    double tot_sq = 0.0;
    for (int k = 0; k < kBins; ++k) { const double p = static_cast<double>(hist[k]) / n; tot_sq += p * p; }
    double best = -DBL_MAX;
    int arg = 0;
    double p2sq_after = tot_sq;            // sum_{j >= k} pmf_j^2
    for (int k = 0; k < kBins - 1; ++k) {
      const double p = static_cast<double>(hist[k]) / n;
      p1 += p; p1sq += p * p;
      p2sq_after -= p * p;                 // now sum_{j >= k+1}
      const double c = log((1.0 / (p1sq * p2sq_after)) * ((p1 * (1.0 - p1)) * (p1 * (1.0 - p1))));
      if (c > best) { best = c; arg = k; } // first maximum, like argmax (NaN never wins)
    }
    thr_yen[g] = lo + (arg + 0.5) * width; // bin centre
  }
  // ---- Li
  // tolerance = half the smallest gap between distinct values of (v - vmin), float32 arithmetic as numpy performs it
  float tol = FLT_MAX;
  for (int i = threadIdx.x; i + 1 < n; i += kT) {
    const float d = __fsub_rn(__fsub_rn(v[i + 1], vmin), __fsub_rn(v[i], vmin));
    if (d > 0.f && d < tol) tol = d;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) tol = fminf(tol, __shfl_xor_sync(kFull, tol, o));
  __shared__ float s_wtol[kT / 32];
  if ((threadIdx.x & 31) == 0) s_wtol[threadIdx.x >> 5] = tol;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = s_wtol[0];
    for (int w = 1; w < kT / 32; ++w) t = fminf(t, s_wtol[w]);
    s_tol = t * 0.5f;
  }
  __syncthreads();
  const double tolerance = s_tol;
  double part = 0.0;
  for (int i = threadIdx.x; i < n; i += kT) part += static_cast<double>(__fsub_rn(v[i], vmin));
  double t_next = block_sum(part, red) / n;
  double t_curr = -2.0 * tolerance;
  int iters = 0;                           // completed loop iterations; callbacks so far = iters + 1
  bool converged = true;
  while (fabs(t_next - t_curr) > tolerance) {
    if (iters + 2 > max_iter) { converged = false; break; }   // the next callback would be number max_iter + 1
    t_curr = t_next;
    double sf = 0.0, sb = 0.0, cf = 0.0;
    for (int i = threadIdx.x; i < n; i += kT) {
      const double w = static_cast<double>(__fsub_rn(v[i], vmin));
      if (w > t_curr) { sf += w; cf += 1.0; } else sb += w;
    }
    sf = block_sum(sf, red); sb = block_sum(sb, red); cf = block_sum(cf, red);
    const double mean_fore = sf / cf, mean_back = sb / (n - cf);
    if (mean_back == 0.0) break;
    t_next = (mean_back - mean_fore) / (log(mean_back) - log(mean_fore));
    ++iters;
  }
  if (threadIdx.x == 0) {
    thr_li[g] = t_next + static_cast<double>(vmin);
    li_iters[g] = converged ? iters : -1;
  }
}

}  // namespace
}  // namespace sgb

using namespace sgb;

extern "C" size_t sgb_dedupe_workspace_bytes(int64_t n) {
  const size_t a = align_up(static_cast<size_t>(n > 0 ? n : 1) * 4);
  return 6 * a + align_up(static_cast<size_t>(n + 1) * 4) + sort_pairs_workspace_bytes(n) + scan_workspace_bytes(n);
}

extern "C" int sgb_dedupe_max(const int64_t* row, const int64_t* seg, const float* sim, int64_t n, int row_bits, int32_t* order,
                              int32_t* count, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(n >= 0 && n < (int64_t(1) << 31), SGB_ERR_RANGE, "dedupe_max: n out of range");
  SGB_REQUIRE(row_bits >= 1 && row_bits <= 63, SGB_ERR_ARG, "dedupe_max: row_bits must be in [1, 63]");
  SGB_REQUIRE(count && ws && ws_bytes >= sgb_dedupe_workspace_bytes(n), SGB_ERR_WORKSPACE, "dedupe_max: workspace too small");
  if (n == 0) { cudaMemsetAsync(count, 0, 4, stream); return check_launch("dedupe_max(empty)"); }
  SGB_REQUIRE(row && seg && sim && order, SGB_ERR_ARG, "dedupe_max: null argument");
  const size_t a = align_up(static_cast<size_t>(n) * 4);
  char* p = static_cast<char*>(ws);
  uint32_t* keys = reinterpret_cast<uint32_t*>(p); p += a;
  uint32_t* skeys = reinterpret_cast<uint32_t*>(p); p += a;
  uint32_t* perm1 = reinterpret_cast<uint32_t*>(p); p += a;
  uint32_t* perm2 = reinterpret_cast<uint32_t*>(p); p += a;
  int32_t* flags = reinterpret_cast<int32_t*>(p); p += a;
  int32_t* winner = reinterpret_cast<int32_t*>(p); p += a;
  int32_t* scan = reinterpret_cast<int32_t*>(p); p += align_up(static_cast<size_t>(n + 1) * 4);
  void* sort_ws = p; p += sort_pairs_workspace_bytes(n);
  void* scan_ws = p;
  row_keys_kernel<<<nblk(n), kT, 0, stream>>>(row, nullptr, n, 0, keys);
  int rc = sort_pairs(keys, nullptr, skeys, perm1, n, row_bits < 32 ? row_bits : 32, sort_ws, sort_pairs_workspace_bytes(n), stream, true);
  if (rc != SGB_OK) return rc;
  const uint32_t* perm = perm1;
  if (row_bits > 32) {
    row_keys_kernel<<<nblk(n), kT, 0, stream>>>(row, perm1, n, 32, keys);
    rc = sort_pairs(keys, perm1, skeys, perm2, n, row_bits - 32, sort_ws, sort_pairs_workspace_bytes(n), stream, true);
    if (rc != SGB_OK) return rc;
    perm = perm2;
  }
  run_best_kernel<<<nblk(n), kT, 0, stream>>>(row, seg, sim, perm, n, flags, winner);
  rc = exclusive_scan_i32(flags, scan, n, scan_ws, scan_workspace_bytes(n), stream);
  if (rc != SGB_OK) return rc;
  compact_winners_kernel<<<nblk(n), kT, 0, stream>>>(flags, scan, winner, n, order, count);
  return check_launch("dedupe_max");
}

extern "C" size_t sgb_gene_threshold_workspace_bytes(int64_t n, int n_genes) {
  const size_t a = align_up(static_cast<size_t>(n > 0 ? n : 1) * 4);
  return 5 * a + align_up(static_cast<size_t>(n_genes + 2) * 4) + sort_pairs_workspace_bytes(n);
}

extern "C" int sgb_gene_thresholds(const void* gene, int gene_bytes, const int64_t* seg, const float* sim, int64_t n, int n_genes,
                                   int max_iter, double* thr_yen, double* thr_li, int32_t* li_iters, int32_t* counts, void* ws,
                                   size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(gene_bytes == 4 || gene_bytes == 8, SGB_ERR_ARG, "gene_thresholds: gene ids must be int32 or int64");
  SGB_REQUIRE(n >= 0 && n < (int64_t(1) << 31) && n_genes >= 1 && n_genes < (1 << 24), SGB_ERR_RANGE, "gene_thresholds: size out of range");
  SGB_REQUIRE(thr_yen && thr_li && li_iters && counts && ws && ws_bytes >= sgb_gene_threshold_workspace_bytes(n, n_genes),
              SGB_ERR_WORKSPACE, "gene_thresholds: null output or workspace too small");
  SGB_REQUIRE(n == 0 || (gene && seg && sim), SGB_ERR_ARG, "gene_thresholds: null input");
  const size_t a = align_up(static_cast<size_t>(n > 0 ? n : 1) * 4);
  char* p = static_cast<char*>(ws);
  uint32_t* keys = reinterpret_cast<uint32_t*>(p); p += a;
  uint32_t* skeys = reinterpret_cast<uint32_t*>(p); p += a;
  uint32_t* perm1 = reinterpret_cast<uint32_t*>(p); p += a;
  uint32_t* perm2 = reinterpret_cast<uint32_t*>(p); p += a;
  float* vals = reinterpret_cast<float*>(p); p += a;
  int32_t* rowptr = reinterpret_cast<int32_t*>(p); p += align_up(static_cast<size_t>(n_genes + 2) * 4);
  void* sort_ws = p;
  if (n == 0) {
    cudaMemsetAsync(rowptr, 0, static_cast<size_t>(n_genes + 2) * 4, stream);
  } else {
    // LSD order: by similarity first, then (stably) by gene -> ascending similarities inside every gene
    sim_keys_kernel<<<nblk(n), kT, 0, stream>>>(sim, n, keys);
    int rc = sort_pairs(keys, nullptr, skeys, perm1, n, 32, sort_ws, sort_pairs_workspace_bytes(n), stream);
    if (rc != SGB_OK) return rc;
    if (gene_bytes == 8) gene_keys_kernel<int64_t><<<nblk(n), kT, 0, stream>>>(static_cast<const int64_t*>(gene), seg, perm1, n, n_genes, keys);
    else gene_keys_kernel<int32_t><<<nblk(n), kT, 0, stream>>>(static_cast<const int32_t*>(gene), seg, perm1, n, n_genes, keys);
    rc = sort_pairs(keys, perm1, skeys, perm2, n, bits_for(n_genes + 1), sort_ws, sort_pairs_workspace_bytes(n), stream);
    if (rc != SGB_OK) return rc;
    rc = rowptr_from_sorted(skeys, n, rowptr, n_genes + 1, stream);
    if (rc != SGB_OK) return rc;
    gather_f32_kernel<<<nblk(n), kT, 0, stream>>>(sim, perm2, n, vals);
  }
  gene_threshold_kernel<<<n_genes, kT, 0, stream>>>(vals, rowptr, n_genes, max_iter, thr_yen, thr_li, li_iters, counts);
  return check_launch("gene_thresholds");
}
