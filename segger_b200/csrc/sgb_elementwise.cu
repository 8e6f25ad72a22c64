// Element-wise / row-wise stages of ISTEncoder.forward and their backward passes:
// activations, embedding gather + deterministic segment-sum backward, positional min/max
// normalisation + sinusoid features, L2 normalisation.
// Reference: /root/reference/src/segger/models/ist_encoder.py:22-31 (sinusoid), :57-79 (per-tile
// normalisation), :312 (Embedding), :320,325 (GELU), :331-332 (F.normalize).
#include "sgb_api_internal.cuh"
#include "sgb_sort.cuh"
#include <algorithm>

namespace sgb {
namespace {

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == SGB_ACT_GELU) return gelu_erf(x);
  if (act == SGB_ACT_SILU) return x / (1.0f + __expf(-x));
  return x;
}
__device__ __forceinline__ float act_grad(float x, int act) {
  if (act == SGB_ACT_GELU) return gelu_erf_grad(x);
  if (act == SGB_ACT_SILU) {
    const float s = 1.0f / (1.0f + __expf(-x));
    return s * (1.0f + x * (1.0f - s));
  }
  return 1.0f;
}

__global__ void act_fwd_kernel(const float* __restrict__ x, int64_t ldx, int64_t M, int64_t N, int act,
                               float* __restrict__ y, int64_t ldy) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int64_t r = i / N, c = i % N;
  y[r * ldy + c] = act_apply(x[r * ldx + c], act);
}
__global__ void act_bwd_kernel(const float* __restrict__ dy, int64_t ldy, const float* __restrict__ x, int64_t ldx,
                               int64_t M, int64_t N, int act, float* __restrict__ dx, int64_t lddx) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int64_t r = i / N, c = i % N;
  dx[r * lddx + c] = dy[r * ldy + c] * act_grad(x[r * ldx + c], act);
}

// 128-bit variants (N % 4 == 0, all leading dimensions % 4 == 0, 16-byte aligned bases); Q = N/4
__global__ void __launch_bounds__(256)
act_fwd_vec_kernel(const float* __restrict__ x, int64_t ldx, int64_t M, int Q, int act, float* __restrict__ y, int64_t ldy) {
  const int64_t total = M * Q;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / Q;
    const int c = static_cast<int>(i - r * Q) * 4;
    const float4 v = ldg4(x + r * ldx + c);
    st4(y + r * ldy + c, make_float4(act_apply(v.x, act), act_apply(v.y, act), act_apply(v.z, act), act_apply(v.w, act)));
  }
}
__global__ void __launch_bounds__(256)
act_bwd_vec_kernel(const float* __restrict__ dy, int64_t ldy, const float* __restrict__ x, int64_t ldx, int64_t M, int Q,
                   int act, float* __restrict__ dx, int64_t lddx) {
  const int64_t total = M * Q;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / Q;
    const int c = static_cast<int>(i - r * Q) * 4;
    const float4 g = ldg4(dy + r * ldy + c), v = ldg4(x + r * ldx + c);
    st4(dx + r * lddx + c, make_float4(g.x * act_grad(v.x, act), g.y * act_grad(v.y, act), g.z * act_grad(v.z, act),
                                       g.w * act_grad(v.w, act)));
  }
}

template <typename IdxT>
__global__ void embedding_fwd_kernel(const float* __restrict__ table, int64_t n_rows, int D, const IdxT* __restrict__ ids,
                                     int64_t N, float* __restrict__ out, int64_t ldo, float* __restrict__ out_act,
                                     int64_t lda, int act) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N * D) return;
  const int64_t r = i / D;
  const int c = static_cast<int>(i % D);
  int64_t g = static_cast<int64_t>(ids[r]);
  g = g < 0 ? 0 : (g >= n_rows ? n_rows - 1 : g);
  const float v = __ldg(table + g * D + c);
  if (out) out[r * ldo + c] = v;
  if (out_act) out_act[r * lda + c] = act_apply(v, act);
}

// D % 4 == 0, 16-byte aligned rows: one thread per 128-bit chunk of an output row (Q = D/4 chunks per row)
template <typename IdxT>
__global__ void __launch_bounds__(256)
embedding_fwd_vec_kernel(const float* __restrict__ table, int64_t n_rows, int D, int Q, const IdxT* __restrict__ ids,
                         int64_t N, float* __restrict__ out, int64_t ldo, float* __restrict__ out_act, int64_t lda, int act) {
  const int64_t total = N * Q;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / Q;
    const int c = static_cast<int>(i - r * Q) * 4;
    int64_t g = static_cast<int64_t>(__ldg(ids + r));
    g = g < 0 ? 0 : (g >= n_rows ? n_rows - 1 : g);
    const float4 v = ldg4(table + g * D + c);
    if (out) st4(out + r * ldo + c, v);
    if (out_act)
      st4(out_act + r * lda + c, make_float4(act_apply(v.x, act), act_apply(v.y, act), act_apply(v.z, act), act_apply(v.w, act)));
  }
}

template <typename IdxT>
__global__ void ids_to_u32_kernel(const IdxT* __restrict__ ids, int64_t N, int64_t n_rows, uint32_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int64_t g = static_cast<int64_t>(ids[i]);
  g = g < 0 ? 0 : (g >= n_rows ? n_rows - 1 : g);
  out[i] = static_cast<uint32_t>(g);
}

constexpr int kEmbChunk = 64;  // sorted rows per warp in the segment-sum backward

// Stage 1: warp per chunk of the id-sorted row list.  Complete segments are written straight to
// grad_table; the (at most two) segments that continue into neighbouring chunks go to part[c][0|1].
__global__ void __launch_bounds__(256)
embedding_bwd_stage1_kernel(const float* __restrict__ dy, int64_t ldy, const uint32_t* __restrict__ sid,
                            const uint32_t* __restrict__ perm, const int32_t* __restrict__ rowptr, int64_t N, int D,
                            const float* __restrict__ table, int act, float* __restrict__ grad_table,
                            float* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const int64_t chunk = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t s0 = chunk * kEmbChunk;
  if (s0 >= N) return;
  const int64_t s1 = min(N, s0 + kEmbChunk);
  for (int d0 = 0; d0 < D; d0 += 256) {
    float acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = 0.f;
    int64_t run_start = s0;
    uint32_t g = sid[s0];
    for (int64_t s = s0; s <= s1; ++s) {
      const bool at_end = (s == s1);
      const uint32_t gs = at_end ? 0xffffffffu : sid[s];
      if (at_end || gs != g) {
        // flush run [run_start, s) of gene g
        const bool complete = (rowptr[g] == run_start) && (rowptr[g + 1] == s);
        float* dst = complete ? grad_table + static_cast<int64_t>(g) * D
                              : part + (chunk * 2 + (run_start == s0 ? 0 : 1)) * D;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int c = d0 + lane + 32 * t;
          if (c < D) dst[c] = acc[t];
          acc[t] = 0.f;
        }
        if (at_end) break;
        g = gs;
        run_start = s;
      }
      const int64_t r = perm[s];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int c = d0 + lane + 32 * t;
        if (c < D) {
          float v = __ldg(dy + r * ldy + c);
          if (act != SGB_ACT_NONE) v *= act_grad(__ldg(table + static_cast<int64_t>(g) * D + c), act);
          acc[t] += v;
        }
      }
    }
  }
}

// Vector form of stage 1 for plain segment sums (no activation): D % 4 == 0, 16-byte aligned rows.  The chunk's 64
// (id, row) pairs are loaded once into registers and broadcast by shuffle, and the rows are gathered U at a time
// (U x V independent 128-bit loads in flight per lane) instead of one dependent perm -> row chain per iteration;
// the sum inside a segment still runs in sorted order, so the result is bit-identical to the scalar kernel.
template <int V, int U>
__global__ void __launch_bounds__(256)
segment_sum_stage1_vec_kernel(const float* __restrict__ dy, int64_t ldy, const uint32_t* __restrict__ sid,
                              const uint32_t* __restrict__ perm, const int32_t* __restrict__ rowptr, int64_t N, int D,
                              int d_base, float* __restrict__ grad_table, float* __restrict__ part) {
  const int lane = threadIdx.x & 31;
  const int64_t chunk = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t s0 = chunk * kEmbChunk;
  if (s0 >= N) return;
  const int n = static_cast<int>(min(static_cast<int64_t>(kEmbChunk), N - s0));
  static_assert(kEmbChunk == 64, "two register-resident (id, row) pairs per lane");
  const uint32_t id_a = lane < n ? sid[s0 + lane] : 0xffffffffu, id_b = 32 + lane < n ? sid[s0 + 32 + lane] : 0xffffffffu;
  const uint32_t pr_a = lane < n ? perm[s0 + lane] : 0u, pr_b = 32 + lane < n ? perm[s0 + 32 + lane] : 0u;
  const int D4 = D >> 2;
  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t g = __shfl_sync(kFull, id_a, 0);
  int run_start = 0;
  auto flush = [&](int s_end) {
    const int64_t gs = s0 + run_start, ge = s0 + s_end;
    const bool complete = (rowptr[g] == gs) && (rowptr[g + 1] == ge);
    float* dst = complete ? grad_table + static_cast<int64_t>(g) * D : part + (chunk * 2 + (run_start == 0 ? 0 : 1)) * D;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int q = (d_base >> 2) + lane + 32 * v;
      if (q < D4) *reinterpret_cast<float4*>(dst + q * 4) = acc[v];
      acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  for (int s = 0; s < n; s += U) {
    float4 x[U][V];
    uint32_t ids[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = s + u;
      const uint32_t r = k < 32 ? __shfl_sync(kFull, pr_a, k & 31) : __shfl_sync(kFull, pr_b, k & 31);
      ids[u] = k < 32 ? __shfl_sync(kFull, id_a, k & 31) : __shfl_sync(kFull, id_b, k & 31);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int q = (d_base >> 2) + lane + 32 * v;
        x[u][v] = (k < n && q < D4) ? ldg4(dy + static_cast<int64_t>(r) * ldy + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int k = s + u;
      if (k >= n) break;
      if (ids[u] != g) {
        flush(k);
        g = ids[u];
        run_start = k;
      }
#pragma unroll
      for (int v = 0; v < V; ++v) { acc[v].x += x[u][v].x; acc[v].y += x[u][v].y; acc[v].z += x[u][v].z; acc[v].w += x[u][v].w; }
    }
  }
  flush(n);
}

// Stage 2: one warp per (table row, 32-column group); segments spanning several chunks are summed in chunk order (the
// order per column is what it was with one warp per row, so results are bit-identical -- but a hot gene's few hundred
// chunks are now walked by D / 32 warps side by side instead of one warp looping over the column groups).
__global__ void __launch_bounds__(256)
embedding_bwd_stage2_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows, int D, int col_groups,
                            const float* __restrict__ part, float* __restrict__ grad_table) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t g = w / col_groups;
  const int c = static_cast<int>(w % col_groups) * 32 + lane;
  if (g >= n_rows || c >= D) return;
  const int64_t s = rowptr[g], e = rowptr[g + 1];
  float* dst = grad_table + g * D;
  if (s == e) {
    dst[c] = 0.f;
    return;
  }
  const int64_t c_first = s / kEmbChunk, c_last = (e - 1) / kEmbChunk;
  if (c_first == c_last) return;  // written by stage 1
  float acc = 0.f;
  for (int64_t ch = c_first; ch <= c_last; ++ch) {
    const int slot = (ch == c_first && s != ch * kEmbChunk) ? 1 : 0;
    acc += part[(ch * 2 + slot) * D + c];
  }
  dst[c] = acc;
}

// ---- positional features ----
__device__ __forceinline__ int f2ord(float f) {  // order-preserving float -> int
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void minmax_init_kernel(int* __restrict__ mm, int64_t n_batches) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_batches * 4) return;
  mm[i] = (i % 4 < 2) ? f2ord(INFINITY) : f2ord(-INFINITY);  // [b][minx,miny,maxx,maxy]
}

template <typename IdxT>
__global__ void minmax_kernel(const float* __restrict__ pos, int64_t N, const IdxT* __restrict__ batch,
                              int64_t n_batches, int* __restrict__ mm) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool valid = i < N;
  int64_t b = 0;
  float x = 0.f, y = 0.f;
  if (valid) {
    b = batch ? static_cast<int64_t>(batch[i]) : 0;
    b = b < 0 ? 0 : (b >= n_batches ? n_batches - 1 : b);
    x = pos[2 * i];
    y = pos[2 * i + 1];
  }
  // warp-aggregate when the whole warp is one tile (the common, tile-major case)
  const unsigned act = __ballot_sync(kFull, valid);
  if (!valid) return;
  const int b0 = __shfl_sync(act, static_cast<int>(b), __ffs(act) - 1);
  const bool uniform = __all_sync(act, static_cast<int>(b) == b0);
  if (uniform) {
    float mnx = x, mny = y, mxx = x, mxy = y;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      const float a = __shfl_xor_sync(act, mnx, o), c = __shfl_xor_sync(act, mny, o);
      const float d = __shfl_xor_sync(act, mxx, o), e = __shfl_xor_sync(act, mxy, o);
      const bool ok = (act >> ((threadIdx.x & 31) ^ o)) & 1u;
      if (ok) { mnx = fminf(mnx, a); mny = fminf(mny, c); mxx = fmaxf(mxx, d); mxy = fmaxf(mxy, e); }
    }
    if ((threadIdx.x & 31) == __ffs(act) - 1) {
      atomicMin(mm + b * 4 + 0, f2ord(mnx)); atomicMin(mm + b * 4 + 1, f2ord(mny));
      atomicMax(mm + b * 4 + 2, f2ord(mxx)); atomicMax(mm + b * 4 + 3, f2ord(mxy));
    }
  } else {
    atomicMin(mm + b * 4 + 0, f2ord(x)); atomicMin(mm + b * 4 + 1, f2ord(y));
    atomicMax(mm + b * 4 + 2, f2ord(x)); atomicMax(mm + b * 4 + 3, f2ord(y));
  }
}

// feat[d][i][0:half] = cos(p*freq), feat[d][i][half:2half] = sin(p*freq), optional zero pad
template <typename IdxT>
__global__ void posfreq_kernel(const float* __restrict__ pos, int64_t N, const IdxT* __restrict__ batch,
                               int64_t n_batches, const int* __restrict__ mm, const float* __restrict__ freqs, int dim,
                               float* __restrict__ feat, int64_t ldf) {
  const int half = dim / 2;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= 2 * N * half) return;
  const int k = static_cast<int>(i % half);
  const int64_t node2 = i / half;  // d * N + node
  const int d = static_cast<int>(node2 / N);
  const int64_t node = node2 % N;
  int64_t b = batch ? static_cast<int64_t>(batch[node]) : 0;
  b = b < 0 ? 0 : (b >= n_batches ? n_batches - 1 : b);
  const float mn = ord2f(mm[b * 4 + d]), mx = ord2f(mm[b * 4 + 2 + d]);
  const float v = pos[2 * node + d];
  // batch given: (pos - min) / (max - min + 1e-8);  batch None: (pos - min) / max(pos - min)
  const float pn = batch ? (v - mn) / (mx - mn + 1e-8f) : (v - mn) / (mx - mn);
  const float arg = pn * __ldg(freqs + k);
  float sn, cs;
  sincosf(arg, &sn, &cs);
  float* row = feat + node2 * ldf;
  row[k] = cs;
  row[half + k] = sn;
  if ((dim & 1) && k == 0) row[dim - 1] = 0.f;
}

// Vectorised variant (half % 4 == 0, ldf % 4 == 0, 16-byte aligned feat): one thread per 4 consecutive
// frequencies of one (axis, node) row -> two 128-bit stores; a warp covers whole rows, so the per-row loads
// (tile id, min/max, coordinate) are warp-uniform broadcasts and the row index needs one 64-bit division
// per thread on a power-of-two-friendly quotient instead of three.
template <typename IdxT>
__global__ void __launch_bounds__(256)
posfreq_vec_kernel(const float* __restrict__ pos, int64_t N, const IdxT* __restrict__ batch, int64_t n_batches,
                   const int* __restrict__ mm, const float* __restrict__ freqs, int half, int q_per_row,
                   float* __restrict__ feat, int64_t ldf) {
  const int64_t total = 2 * N * q_per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t node2 = i / q_per_row;            // d * N + node
    const int k = static_cast<int>(i - node2 * q_per_row) * 4;
    const int d = node2 >= N ? 1 : 0;
    const int64_t node = node2 - (d ? N : 0);
    int64_t b = batch ? static_cast<int64_t>(__ldg(batch + node)) : 0;
    b = b < 0 ? 0 : (b >= n_batches ? n_batches - 1 : b);
    const float mn = ord2f(__ldg(mm + b * 4 + d)), mx = ord2f(__ldg(mm + b * 4 + 2 + d));
    const float v = __ldg(pos + 2 * node + d);
    const float pn = batch ? (v - mn) / (mx - mn + 1e-8f) : (v - mn) / (mx - mn);
    const float4 f = ldg4(freqs + k);
    float4 sn, cs;
    sincosf(pn * f.x, &sn.x, &cs.x);
    sincosf(pn * f.y, &sn.y, &cs.y);
    sincosf(pn * f.z, &sn.z, &cs.z);
    sincosf(pn * f.w, &sn.w, &cs.w);
    float* row = feat + node2 * ldf;
    st4(row + k, cs);
    st4(row + half + k, sn);
  }
}

// Chebyshev basis of the normalised coordinates: out[d*N + node][n] = T_n(2p - 1), n < deg, p as in posfreq.
// The sinusoid features cos(p f_k), sin(p f_k) (f_k <= 1, p in [0,1]) are entire functions of p whose Chebyshev
// coefficients decay like (f_k/4)^n / n!, so a dozen basis columns reproduce all 256 feature columns to fp32
// rounding: feat = T M with a constant [deg, 256] coefficient matrix M.  The first positional Linear and its
// weight gradient then contract over deg instead of 256 columns, and the [2N, 256] feature matrix is never
// materialised (ops.InputStageFn).  One thread per (axis, node) row; deg % 4 == 0 -> 128-bit stores.
template <typename IdxT>
__global__ void __launch_bounds__(256)
poscheb_kernel(const float* __restrict__ pos, int64_t N, const IdxT* __restrict__ batch, int64_t n_batches,
               const int* __restrict__ mm, int deg, float* __restrict__ out, int64_t ldo) {
  for (int64_t node2 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; node2 < 2 * N;
       node2 += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int d = node2 >= N ? 1 : 0;
    const int64_t node = node2 - (d ? N : 0);
    int64_t b = batch ? static_cast<int64_t>(__ldg(batch + node)) : 0;
    b = b < 0 ? 0 : (b >= n_batches ? n_batches - 1 : b);
    const float mn = ord2f(__ldg(mm + b * 4 + d)), mx = ord2f(__ldg(mm + b * 4 + 2 + d));
    const float v = __ldg(pos + 2 * node + d);
    const float pn = batch ? (v - mn) / (mx - mn + 1e-8f) : (v - mn) / (mx - mn);
    const float t = 2.0f * pn - 1.0f, t2 = 2.0f * t;
    float* row = out + node2 * ldo;
    float a = 1.0f, bb = t;                      // T_n, T_{n+1}
    for (int n = 0; n < deg; n += 4) {
      float4 o;
      o.x = a; o.y = bb;
      const float c = fmaf(t2, bb, -a), e = fmaf(t2, c, -bb);
      o.z = c; o.w = e;
      st4(row + n, o);
      a = fmaf(t2, e, -c);
      bb = fmaf(t2, a, -e);
    }
  }
}

// ---- row subset helpers (projection of the tx-belongs-bd sources only) ----
// flag[0] = 1 iff ids[0] < ids[1] < ... < ids[n-1] (set to 1 by the init launch, cleared by any violating pair)
template <typename IdxT>
__global__ void strictly_increasing_kernel(const IdxT* __restrict__ ids, int64_t n, int32_t* __restrict__ flag) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i + 1 < n && !(ids[i] < ids[i + 1])) *flag = 0;
}
__global__ void flag_set_kernel(int32_t* flag, int32_t v) { *flag = v; }

// dst[ids[k], :] += src[k, :]; ids are unique (strictly increasing), so no two threads touch the same element.
template <typename IdxT>
__global__ void __launch_bounds__(256)
rows_add_kernel(float* __restrict__ dst, int64_t ldd, int64_t n_dst_rows, const IdxT* __restrict__ ids, int64_t n, int Q,
                const float* __restrict__ src, int64_t lds) {
  const int64_t total = n * Q;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t k = i / Q;
    const int c = static_cast<int>(i - k * Q) * 4;
    const int64_t r = static_cast<int64_t>(__ldg(ids + k));
    if (r < 0 || r >= n_dst_rows) continue;
    const float4 a = ldg4(src + k * lds + c);
    float* d = dst + r * ldd + c;
    float4 o = *reinterpret_cast<const float4*>(d);
    o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    *reinterpret_cast<float4*>(d) = o;
  }
}

// ---- L2 normalise ----
__global__ void __launch_bounds__(256)
l2norm_fwd_kernel(const float* __restrict__ x, int64_t ldx, int64_t M, int D, float eps, float* __restrict__ y,
                  int64_t ldy, float* __restrict__ norm) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= M) return;
  float ss = 0.f;
  for (int c = lane; c < D; c += 32) { const float v = x[r * ldx + c]; ss += v * v; }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
  const float nrm = sqrtf(ss);
  const float inv = 1.0f / fmaxf(nrm, eps);
  for (int c = lane; c < D; c += 32) y[r * ldy + c] = x[r * ldx + c] * inv;
  if (lane == 0 && norm) norm[r] = nrm;
}
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ dy, int64_t ldy, const float* __restrict__ y, int64_t ldyy,
                  const float* __restrict__ norm, int64_t M, int D, float eps, float* __restrict__ dx, int64_t lddx) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= M) return;
  const float nrm = norm[r];
  float dot = 0.f;
  for (int c = lane; c < D; c += 32) dot += dy[r * ldy + c] * y[r * ldyy + c];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) dot += __shfl_xor_sync(kFull, dot, o);
  const bool clamped = !(nrm > eps);
  const float inv = 1.0f / fmaxf(nrm, eps);
  for (int c = lane; c < D; c += 32) {
    const float g = dy[r * ldy + c];
    dx[r * lddx + c] = clamped ? g * inv : (g - y[r * ldyy + c] * dot) * inv;
  }
}

// D = 4 * LPR (32, 64 or 128 columns): LPR lanes own a row, one 128-bit chunk each -> a warp covers 32/LPR rows
template <int LPR>
__global__ void __launch_bounds__(256)
l2norm_fwd_vec_kernel(const float* __restrict__ x, int64_t ldx, int64_t M, float eps, float* __restrict__ y,
                      int64_t ldy, float* __restrict__ norm) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, s = lane % LPR;
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t r = w * RPW + lane / LPR;
  const bool ok = r < M;
  const float4 v = ok ? ldg4(x + r * ldx + s * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
  for (int o = LPR / 2; o >= 1; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
  const float nrm = sqrtf(ss);
  const float inv = 1.0f / fmaxf(nrm, eps);
  if (ok) {
    st4(y + r * ldy + s * 4, make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv));
    if (s == 0 && norm) norm[r] = nrm;
  }
}
template <int LPR>
__global__ void __launch_bounds__(256)
l2norm_bwd_vec_kernel(const float* __restrict__ dy, int64_t ldy, const float* __restrict__ y, int64_t ldyy,
                      const float* __restrict__ norm, int64_t M, float eps, float* __restrict__ dx, int64_t lddx) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, s = lane % LPR;
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t r = w * RPW + lane / LPR;
  const bool ok = r < M;
  const float4 g = ok ? ldg4(dy + r * ldy + s * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 e = ok ? ldg4(y + r * ldyy + s * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float nrm = ok ? __ldg(norm + r) : 1.f;
  float dot = g.x * e.x + g.y * e.y + g.z * e.z + g.w * e.w;
#pragma unroll
  for (int o = LPR / 2; o >= 1; o >>= 1) dot += __shfl_xor_sync(kFull, dot, o);
  const bool clamped = !(nrm > eps);
  const float inv = 1.0f / fmaxf(nrm, eps);
  if (ok)
    st4(dx + r * lddx + s * 4, clamped ? make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv)
                                       : make_float4((g.x - e.x * dot) * inv, (g.y - e.y * dot) * inv,
                                                     (g.z - e.z * dot) * inv, (g.w - e.w * dot) * inv));
}

}  // namespace
}  // namespace sgb

using namespace sgb;

extern "C" int sgb_act_fwd(const float* x, int64_t ldx, int64_t M, int64_t N, int act, float* y, int64_t ldy, void* stream) {
  SGB_REQUIRE(M >= 0 && N >= 0, SGB_ERR_ARG, "act_fwd: negative size");
  if (M * N == 0) return SGB_OK;
  SGB_REQUIRE(x && y && ldx >= N && ldy >= N, SGB_ERR_ARG, "act_fwd: bad argument");
  if (N % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && aligned16(x) && aligned16(y)) {
    const unsigned vb = static_cast<unsigned>(std::min<int64_t>(ceil_div(M * (N / 4), 256), static_cast<int64_t>(sm_count()) * 32));
    act_fwd_vec_kernel<<<vb, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, M, static_cast<int>(N / 4), act, y, ldy);
    return check_launch("act_fwd");
  }
  act_fwd_kernel<<<static_cast<unsigned>(ceil_div(M * N, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, M, N, act, y, ldy);
  return check_launch("act_fwd");
}
extern "C" int sgb_act_bwd(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N, int act,
                           float* dx, int64_t lddx, void* stream) {
  SGB_REQUIRE(M >= 0 && N >= 0, SGB_ERR_ARG, "act_bwd: negative size");
  if (M * N == 0) return SGB_OK;
  SGB_REQUIRE(x && dy && dx && ldx >= N && ldy >= N && lddx >= N, SGB_ERR_ARG, "act_bwd: bad argument");
  if (N % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && lddx % 4 == 0 && aligned16(x) && aligned16(dy) && aligned16(dx)) {
    const unsigned vb = static_cast<unsigned>(std::min<int64_t>(ceil_div(M * (N / 4), 256), static_cast<int64_t>(sm_count()) * 32));
    act_bwd_vec_kernel<<<vb, 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, ldy, x, ldx, M, static_cast<int>(N / 4), act, dx, lddx);
    return check_launch("act_bwd");
  }
  act_bwd_kernel<<<static_cast<unsigned>(ceil_div(M * N, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, ldy, x, ldx, M, N, act, dx, lddx);
  return check_launch("act_bwd");
}

extern "C" int sgb_embedding_fwd(const float* table, int64_t n_rows, int D, const void* ids, int idx_bytes, int64_t N,
                                 float* out, int64_t ldo, float* out_act, int64_t lda, int act, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "embedding_fwd: idx_bytes must be 4 or 8");
  SGB_REQUIRE(N >= 0 && D >= 1 && n_rows >= 1, SGB_ERR_ARG, "embedding_fwd: bad size");
  if (N == 0) return SGB_OK;
  SGB_REQUIRE(table && ids && (out || out_act), SGB_ERR_ARG, "embedding_fwd: null tensor");
  const bool vec = D % 4 == 0 && aligned16(table) && (!out || (aligned16(out) && ldo % 4 == 0)) &&
                   (!out_act || (aligned16(out_act) && lda % 4 == 0));
  if (vec) {
    const int Q = D / 4;
    const unsigned vb = static_cast<unsigned>(std::min<int64_t>(ceil_div(N * Q, 256), static_cast<int64_t>(sm_count()) * 32));
    if (idx_bytes == 8)
      embedding_fwd_vec_kernel<int64_t><<<vb, 256, 0, stream>>>(table, n_rows, D, Q, static_cast<const int64_t*>(ids), N, out, ldo, out_act, lda, act);
    else
      embedding_fwd_vec_kernel<int32_t><<<vb, 256, 0, stream>>>(table, n_rows, D, Q, static_cast<const int32_t*>(ids), N, out, ldo, out_act, lda, act);
    return check_launch("embedding_fwd");
  }
  const unsigned blocks = static_cast<unsigned>(ceil_div(N * D, 256));
  if (idx_bytes == 8)
    embedding_fwd_kernel<int64_t><<<blocks, 256, 0, stream>>>(table, n_rows, D, static_cast<const int64_t*>(ids), N, out, ldo, out_act, lda, act);
  else
    embedding_fwd_kernel<int32_t><<<blocks, 256, 0, stream>>>(table, n_rows, D, static_cast<const int32_t*>(ids), N, out, ldo, out_act, lda, act);
  return check_launch("embedding_fwd");
}

namespace {
struct EmbWs { uint32_t *ids32, *sid, *perm; int32_t* rowptr; float* part; void* sort_ws; size_t sort_bytes, total; };
EmbWs emb_carve(void* ws, int64_t N, int64_t n_rows, int D) {
  EmbWs e{};
  const size_t nb = align_up(static_cast<size_t>(N > 0 ? N : 1) * 4);
  char* p = static_cast<char*>(ws);
  e.ids32 = reinterpret_cast<uint32_t*>(p); p += nb;
  e.sid = reinterpret_cast<uint32_t*>(p); p += nb;
  e.perm = reinterpret_cast<uint32_t*>(p); p += nb;
  e.rowptr = reinterpret_cast<int32_t*>(p); p += align_up(static_cast<size_t>(n_rows + 1) * 4);
  const size_t chunks = static_cast<size_t>(ceil_div(N > 0 ? N : 1, kEmbChunk));
  e.part = reinterpret_cast<float*>(p); p += align_up(chunks * 2 * D * sizeof(float));
  e.sort_ws = p;
  e.sort_bytes = sort_pairs_workspace_bytes(N);
  e.total = static_cast<size_t>(p - static_cast<char*>(ws)) + e.sort_bytes;
  return e;
}
}  // namespace

extern "C" size_t sgb_embedding_bwd_workspace_bytes(int64_t N, int D, int64_t n_rows) {
  return emb_carve(nullptr, N, n_rows, D).total;
}

extern "C" int sgb_embedding_bwd(const float* dy, int64_t ldy, const void* ids, int idx_bytes, int64_t N, int D,
                                 int64_t n_rows, const float* table, int act, float* grad_table, void* ws,
                                 size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "embedding_bwd: idx_bytes must be 4 or 8");
  SGB_REQUIRE(N >= 0 && N < (int64_t(1) << 31) && D >= 1 && D <= 8192 && n_rows >= 1 && n_rows < (int64_t(1) << 31), SGB_ERR_RANGE,
              "embedding_bwd: size out of range");
  SGB_REQUIRE(grad_table && ws, SGB_ERR_ARG, "embedding_bwd: null tensor");
  EmbWs e = emb_carve(ws, N, n_rows, D);
  SGB_REQUIRE(ws_bytes >= e.total, SGB_ERR_WORKSPACE, "embedding_bwd: workspace too small");
  if (N == 0) {
    cudaMemsetAsync(grad_table, 0, static_cast<size_t>(n_rows) * D * sizeof(float), stream);
    return check_launch("embedding_bwd(empty)");
  }
  SGB_REQUIRE(dy && ids, SGB_ERR_ARG, "embedding_bwd: null tensor");
  const unsigned nblk = static_cast<unsigned>(ceil_div(N, 256));
  if (idx_bytes == 8) ids_to_u32_kernel<int64_t><<<nblk, 256, 0, stream>>>(static_cast<const int64_t*>(ids), N, n_rows, e.ids32);
  else ids_to_u32_kernel<int32_t><<<nblk, 256, 0, stream>>>(static_cast<const int32_t*>(ids), N, n_rows, e.ids32);
  int rc = sort_pairs(e.ids32, nullptr, e.sid, e.perm, N, bits_for(n_rows), e.sort_ws, e.sort_bytes, stream);
  if (rc != SGB_OK) return rc;
  rc = rowptr_from_sorted(e.sid, N, e.rowptr, n_rows, stream);
  if (rc != SGB_OK) return rc;
  const int64_t chunks = ceil_div(N, kEmbChunk);
  const unsigned s1_blocks = static_cast<unsigned>(ceil_div(chunks, 8));
  if (!table && D % 4 == 0 && ldy % 4 == 0 && aligned16(dy) && aligned16(grad_table)) {
    // plain segment sum: 128-bit gathers, several rows in flight; 1024 columns per pass
    for (int d0 = 0; d0 < D; d0 += 1024) {
      const int w = std::min(D - d0, 1024);
      if (w <= 128) segment_sum_stage1_vec_kernel<1, 8><<<s1_blocks, 256, 0, stream>>>(dy, ldy, e.sid, e.perm, e.rowptr, N, D, d0, grad_table, e.part);
      else if (w <= 256) segment_sum_stage1_vec_kernel<2, 4><<<s1_blocks, 256, 0, stream>>>(dy, ldy, e.sid, e.perm, e.rowptr, N, D, d0, grad_table, e.part);
      else if (w <= 512) segment_sum_stage1_vec_kernel<4, 2><<<s1_blocks, 256, 0, stream>>>(dy, ldy, e.sid, e.perm, e.rowptr, N, D, d0, grad_table, e.part);
      else segment_sum_stage1_vec_kernel<8, 2><<<s1_blocks, 256, 0, stream>>>(dy, ldy, e.sid, e.perm, e.rowptr, N, D, d0, grad_table, e.part);
    }
  } else {
    embedding_bwd_stage1_kernel<<<s1_blocks, 256, 0, stream>>>(
        dy, ldy, e.sid, e.perm, e.rowptr, N, D, table, table ? act : SGB_ACT_NONE, grad_table, e.part);
  }
  const int col_groups = (D + 31) / 32;
  embedding_bwd_stage2_kernel<<<static_cast<unsigned>(ceil_div(n_rows * col_groups, 8)), 256, 0, stream>>>(e.rowptr, n_rows, D, col_groups,
                                                                                                  e.part, grad_table);
  return check_launch("embedding_bwd");
}

extern "C" size_t sgb_posfreq_workspace_bytes(int64_t n_batches) {
  return align_up(static_cast<size_t>(n_batches > 0 ? n_batches : 1) * 4 * sizeof(int));
}

extern "C" int sgb_posfreq_fwd(const float* pos, int64_t N, const void* batch, int idx_bytes, int64_t n_batches,
                               int dim, const float* freqs, float* feat, int64_t ldf, void* ws, size_t ws_bytes,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(N >= 0 && dim >= 2 && n_batches >= 1, SGB_ERR_ARG, "posfreq_fwd: bad size");
  SGB_REQUIRE(!batch || idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "posfreq_fwd: idx_bytes must be 4 or 8");
  if (N == 0) return SGB_OK;
  SGB_REQUIRE(pos && freqs && feat && ws && ldf >= dim, SGB_ERR_ARG, "posfreq_fwd: null tensor");
  SGB_REQUIRE(ws_bytes >= sgb_posfreq_workspace_bytes(n_batches), SGB_ERR_WORKSPACE, "posfreq_fwd: workspace too small");
  int* mm = static_cast<int*>(ws);
  minmax_init_kernel<<<static_cast<unsigned>(ceil_div(n_batches * 4, 256)), 256, 0, stream>>>(mm, n_batches);
  const unsigned nb = static_cast<unsigned>(ceil_div(N, 256));
  const int64_t total = 2 * N * (dim / 2);
  const unsigned fb = static_cast<unsigned>(ceil_div(total, 256));
  const int half = dim / 2;
  const bool vec = (dim % 2 == 0) && (half % 4 == 0) && (ldf % 4 == 0) && aligned16(feat) && aligned16(freqs);
  const int qpr = half / 4;
  const unsigned vb = static_cast<unsigned>(std::min<int64_t>(ceil_div(2 * N * (vec ? qpr : 1), 256), static_cast<int64_t>(sm_count()) * 64));
  if (vec && batch && idx_bytes == 8) {
    minmax_kernel<int64_t><<<nb, 256, 0, stream>>>(pos, N, static_cast<const int64_t*>(batch), n_batches, mm);
    posfreq_vec_kernel<int64_t><<<vb, 256, 0, stream>>>(pos, N, static_cast<const int64_t*>(batch), n_batches, mm, freqs, half, qpr, feat, ldf);
  } else if (vec) {
    const int32_t* b32 = static_cast<const int32_t*>(batch);
    minmax_kernel<int32_t><<<nb, 256, 0, stream>>>(pos, N, b32, n_batches, mm);
    posfreq_vec_kernel<int32_t><<<vb, 256, 0, stream>>>(pos, N, b32, n_batches, mm, freqs, half, qpr, feat, ldf);
  } else if (batch && idx_bytes == 8) {
    minmax_kernel<int64_t><<<nb, 256, 0, stream>>>(pos, N, static_cast<const int64_t*>(batch), n_batches, mm);
    posfreq_kernel<int64_t><<<fb, 256, 0, stream>>>(pos, N, static_cast<const int64_t*>(batch), n_batches, mm, freqs, dim, feat, ldf);
  } else {
    const int32_t* b32 = static_cast<const int32_t*>(batch);
    minmax_kernel<int32_t><<<nb, 256, 0, stream>>>(pos, N, b32, n_batches, mm);
    posfreq_kernel<int32_t><<<fb, 256, 0, stream>>>(pos, N, b32, n_batches, mm, freqs, dim, feat, ldf);
  }
  return check_launch("posfreq_fwd");
}

extern "C" int sgb_poscheb_fwd(const float* pos, int64_t N, const void* batch, int idx_bytes, int64_t n_batches,
                               int deg, float* out, int64_t ldo, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(N >= 0 && deg >= 4 && deg % 4 == 0 && n_batches >= 1, SGB_ERR_ARG, "poscheb_fwd: bad size (deg % 4 == 0)");
  SGB_REQUIRE(!batch || idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "poscheb_fwd: idx_bytes must be 4 or 8");
  if (N == 0) return SGB_OK;
  SGB_REQUIRE(pos && out && ws && ldo >= deg, SGB_ERR_ARG, "poscheb_fwd: null tensor");
  SGB_REQUIRE(ldo % 4 == 0 && aligned16(out), SGB_ERR_ALIGN, "poscheb_fwd: out must be 16-byte aligned with ldo % 4 == 0");
  SGB_REQUIRE(ws_bytes >= sgb_posfreq_workspace_bytes(n_batches), SGB_ERR_WORKSPACE, "poscheb_fwd: workspace too small");
  int* mm = static_cast<int*>(ws);
  minmax_init_kernel<<<static_cast<unsigned>(ceil_div(n_batches * 4, 256)), 256, 0, stream>>>(mm, n_batches);
  const unsigned nb = static_cast<unsigned>(ceil_div(N, 256));
  const unsigned cb = static_cast<unsigned>(std::min<int64_t>(ceil_div(2 * N, 256), static_cast<int64_t>(sm_count()) * 32));
  if (batch && idx_bytes == 8) {
    minmax_kernel<int64_t><<<nb, 256, 0, stream>>>(pos, N, static_cast<const int64_t*>(batch), n_batches, mm);
    poscheb_kernel<int64_t><<<cb, 256, 0, stream>>>(pos, N, static_cast<const int64_t*>(batch), n_batches, mm, deg, out, ldo);
  } else {
    const int32_t* b32 = static_cast<const int32_t*>(batch);
    minmax_kernel<int32_t><<<nb, 256, 0, stream>>>(pos, N, b32, n_batches, mm);
    poscheb_kernel<int32_t><<<cb, 256, 0, stream>>>(pos, N, b32, n_batches, mm, deg, out, ldo);
  }
  return check_launch("poscheb_fwd");
}

extern "C" int sgb_index_strictly_increasing(const void* ids, int idx_bytes, int64_t n, int32_t* flag, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "index_strictly_increasing: idx_bytes must be 4 or 8");
  SGB_REQUIRE(n >= 0 && flag && (n == 0 || ids), SGB_ERR_ARG, "index_strictly_increasing: bad argument");
  flag_set_kernel<<<1, 1, 0, stream>>>(flag, 1);
  if (n > 1) {
    const unsigned blocks = static_cast<unsigned>(ceil_div(n, 256));
    if (idx_bytes == 8) strictly_increasing_kernel<int64_t><<<blocks, 256, 0, stream>>>(static_cast<const int64_t*>(ids), n, flag);
    else strictly_increasing_kernel<int32_t><<<blocks, 256, 0, stream>>>(static_cast<const int32_t*>(ids), n, flag);
  }
  return check_launch("index_strictly_increasing");
}

extern "C" int sgb_rows_add(float* dst, int64_t ldd, int64_t n_dst_rows, const void* ids, int idx_bytes, int64_t n, int D,
                            const float* src, int64_t lds, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(idx_bytes == 4 || idx_bytes == 8, SGB_ERR_ARG, "rows_add: idx_bytes must be 4 or 8");
  SGB_REQUIRE(n >= 0 && D >= 4 && D % 4 == 0 && n_dst_rows >= 0, SGB_ERR_ARG, "rows_add: bad size (D % 4 == 0)");
  if (n == 0) return SGB_OK;
  SGB_REQUIRE(dst && ids && src && ldd >= D && lds >= D, SGB_ERR_ARG, "rows_add: null tensor");
  SGB_REQUIRE(ldd % 4 == 0 && lds % 4 == 0 && aligned16(dst) && aligned16(src), SGB_ERR_ALIGN, "rows_add: 16-byte alignment required");
  const int Q = D / 4;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(n * Q, 256), static_cast<int64_t>(sm_count()) * 32));
  if (idx_bytes == 8) rows_add_kernel<int64_t><<<blocks, 256, 0, stream>>>(dst, ldd, n_dst_rows, static_cast<const int64_t*>(ids), n, Q, src, lds);
  else rows_add_kernel<int32_t><<<blocks, 256, 0, stream>>>(dst, ldd, n_dst_rows, static_cast<const int32_t*>(ids), n, Q, src, lds);
  return check_launch("rows_add");
}

extern "C" int sgb_l2norm_fwd(const float* x, int64_t ldx, int64_t M, int D, float eps, float* y, int64_t ldy,
                              float* norm, void* stream) {
  SGB_REQUIRE(M >= 0 && D >= 1, SGB_ERR_ARG, "l2norm_fwd: bad size");
  if (M == 0) return SGB_OK;
  SGB_REQUIRE(x && y, SGB_ERR_ARG, "l2norm_fwd: null tensor");
  if ((D == 32 || D == 64 || D == 128) && ldx % 4 == 0 && ldy % 4 == 0 && aligned16(x) && aligned16(y)) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rpw = 32 / (D / 4);
    const unsigned vb = static_cast<unsigned>(ceil_div(ceil_div(M, rpw), 8));
    if (D == 32) l2norm_fwd_vec_kernel<8><<<vb, 256, 0, st>>>(x, ldx, M, eps, y, ldy, norm);
    else if (D == 64) l2norm_fwd_vec_kernel<16><<<vb, 256, 0, st>>>(x, ldx, M, eps, y, ldy, norm);
    else l2norm_fwd_vec_kernel<32><<<vb, 256, 0, st>>>(x, ldx, M, eps, y, ldy, norm);
    return check_launch("l2norm_fwd");
  }
  l2norm_fwd_kernel<<<static_cast<unsigned>(ceil_div(M, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, M, D, eps, y, ldy, norm);
  return check_launch("l2norm_fwd");
}
extern "C" int sgb_l2norm_bwd(const float* dy, int64_t ldy, const float* y, int64_t ldyy, const float* norm, int64_t M,
                              int D, float eps, float* dx, int64_t lddx, void* stream) {
  SGB_REQUIRE(M >= 0 && D >= 1, SGB_ERR_ARG, "l2norm_bwd: bad size");
  if (M == 0) return SGB_OK;
  SGB_REQUIRE(dy && y && norm && dx, SGB_ERR_ARG, "l2norm_bwd: null tensor");
  if ((D == 32 || D == 64 || D == 128) && ldy % 4 == 0 && ldyy % 4 == 0 && lddx % 4 == 0 && aligned16(dy) && aligned16(y) &&
      aligned16(dx)) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rpw = 32 / (D / 4);
    const unsigned vb = static_cast<unsigned>(ceil_div(ceil_div(M, rpw), 8));
    if (D == 32) l2norm_bwd_vec_kernel<8><<<vb, 256, 0, st>>>(dy, ldy, y, ldyy, norm, M, eps, dx, lddx);
    else if (D == 64) l2norm_bwd_vec_kernel<16><<<vb, 256, 0, st>>>(dy, ldy, y, ldyy, norm, M, eps, dx, lddx);
    else l2norm_bwd_vec_kernel<32><<<vb, 256, 0, st>>>(dy, ldy, y, ldyy, norm, M, eps, dx, lddx);
    return check_launch("l2norm_bwd");
  }
  l2norm_bwd_kernel<<<static_cast<unsigned>(ceil_div(M, 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy, ldy, y, ldyy, norm, M, D, eps, dx, lddx);
  return check_launch("l2norm_bwd");
}
