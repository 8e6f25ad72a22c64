// Fused GATv2 attention / segment-softmax / aggregation kernels (forward + deterministic backward).
//
// What it replaces: everything torch_geometric's GATv2Conv.forward does after the two linear
// projections -- index_select x2, add, leaky_relu, mul+sum with att, scatter-amax / exp /
// scatter-add softmax, dropout, mul, scatter-add aggregation, bias -- as configured by segger at
// /root/reference/src/segger/models/ist_encoder.py:111-131 (math: SURVEY.md Appendix A.1 / D).
//
// Design (HBM/L2-bound gather work, no tensor cores):
//  * dst-sorted CSR, one warp owns one destination row: the row's softmax statistics, output and
//    grad_x_r are plain register reductions -> no atomics, fixed summation order.
//  * a feature row is H*C fp32 = 128*VEC floats; lane l holds float4 #(v*32+l), so every gather of
//    x_l[j] is VEC fully-coalesced 512-byte warp transactions (128-bit per lane).
//  * edges are processed in chunks: all gathers of a chunk are issued before the first use
//    (CH*VEC independent 128-bit loads in flight per lane), logits of the chunk are reduced with
//    warp shuffles, then a chunked online softmax rescales the running accumulator.
//  * nothing per-edge of width C is ever written; the backward emits two per-edge *scalars*
//    per head (delta, alpha') from the dst pass for the src-sorted (transposed) pass.
//  * arbitrary (H, C) fall back to a warp-per-(row, head) scalar kernel with the same structure.
#include "sgb_api_internal.cuh"
#include "sgb_gatv2.cuh"
#include <cstdlib>

namespace sgb {
namespace {

// ------------------------------------------------------------------------------------------------
// Head <-> lane mapping for the vector path.  CV = float4 per head (C = 4*CV), F = 128*VEC.
// ------------------------------------------------------------------------------------------------
template <int VEC, int CV>
struct HeadMap {
  static constexpr bool kSub = (CV <= 32);            // a head spans CV <= 32 lanes of one vector
  static constexpr int kVPH = kSub ? 1 : CV / 32;     // vectors per head otherwise
  static constexpr int kSlots = VEC / kVPH;           // softmax states each lane tracks
  static_assert(kSub ? (32 % CV == 0) : (CV % 32 == 0 && VEC % (CV / 32) == 0), "unsupported head layout");

  __device__ __forceinline__ static int slot_of(int v) { return v / kVPH; }
  __device__ __forceinline__ static int head(int slot, int lane) {
    return kSub ? (slot * 32 + lane) / CV : slot;
  }
  __device__ __forceinline__ static bool writer(int lane) { return kSub ? (lane % CV) == 0 : lane == 0; }

  // per-vector partial dot products -> per-slot totals, broadcast to every lane of the head group
  __device__ __forceinline__ static void reduce(const float (&part)[VEC], float (&out)[kSlots]) {
    if constexpr (kSub) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float x = part[v];
#pragma unroll
        for (int o = CV / 2; o >= 1; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
        out[v] = x;
      }
    } else {
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        float x = 0.f;
#pragma unroll
        for (int t = 0; t < kVPH; ++t) x += part[s * kVPH + t];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
        out[s] = x;
      }
    }
  }
};

template <int VEC> struct Chunk { static constexpr int value = VEC == 1 ? 8 : (VEC == 2 ? 4 : 2); };

// ================================================================================================
// Forward, vector path
// ================================================================================================
template <int VEC, int CV>
__global__ void __launch_bounds__(256) gatv2_fwd_vec_kernel(const GatParams p) {
  const uint64_t seed_eff = p.training ? gat_seed(p) : 0;
  using HM = HeadMap<VEC, CV>;
  constexpr int S = HM::kSlots;
  constexpr int CH = Chunk<VEC>::value;
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= p.n_dst) return;

  float4 a4[VEC], r4[VEC], acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int off = (v * 32 + lane) * 4;
    a4[v] = ldg4(p.att + off);
    r4[v] = ldg4(p.x_r + row * p.ld_r + off);
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float m[S], s[S];
#pragma unroll
  for (int q = 0; q < S; ++q) { m[q] = -INFINITY; s[q] = 0.f; }

  const int beg = p.rowptr[row], end = p.rowptr[row + 1];
  for (int base = beg; base < end; base += CH) {
    const int n = min(CH, end - base);
    int mycol = 0, myeid = 0;
    if (lane < n) {
      mycol = __ldg(p.col + base + lane);
      if (p.training) myeid = __ldg(p.eid + base + lane);
    }
    float4 x[CH][VEC];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int j = __shfl_sync(kFull, mycol, c);
      if (c < n) {
        const float* src = p.x_l + static_cast<int64_t>(j) * p.ld_l;
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[c][v] = ldg4(src + (v * 32 + lane) * 4);
      }
    }
    float lg[CH][S];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (c < n) {
        float part[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) part[v] = dot4(a4[v], lrelu4(add4(x[c][v], r4[v]), p.slope));
        HM::reduce(part, lg[c]);
      } else {
#pragma unroll
        for (int q = 0; q < S; ++q) lg[c][q] = -INFINITY;
      }
    }
    float sc[S];
#pragma unroll
    for (int q = 0; q < S; ++q) {
      float mc = lg[0][q];
#pragma unroll
      for (int c = 1; c < CH; ++c) mc = fmaxf(mc, lg[c][q]);
      const float mn = fmaxf(m[q], mc);
      sc[q] = __expf(m[q] - mn);
      s[q] *= sc[q];
      m[q] = mn;
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) scale4(acc[v], sc[HM::slot_of(v)]);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (c < n) {
        const int e = __shfl_sync(kFull, myeid, c);
        float w[S];
#pragma unroll
        for (int q = 0; q < S; ++q) {
          const float pe = __expf(lg[c][q] - m[q]);
          s[q] += pe;
          w[q] = pe;
          if (p.training)
            w[q] = rng_keep(seed_eff, e, p.H, HM::head(q, lane), p.drop_thr) ? pe * p.keep_scale : 0.f;
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) fma4(acc[v], w[HM::slot_of(v)], x[c][v]);
      }
    }
  }

  float inv[S];
#pragma unroll
  for (int q = 0; q < S; ++q) {
    const float den = s[q] + 1e-16f;
    inv[q] = 1.0f / den;
    if (HM::writer(lane)) {
      const int h = HM::head(q, lane);
      p.stat_max[row * p.H + h] = (beg == end) ? 0.f : m[q];
      p.stat_den[row * p.H + h] = den;
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int off = (v * 32 + lane) * 4;
    float4 o = acc[v];
    scale4(o, inv[HM::slot_of(v)]);
    if (p.bias) o = add4(o, ldg4(p.bias + off));
    if (p.out) st4(p.out + row * p.ld_out + off, o);
    if (p.out_act)
      st4(p.out_act + row * p.ld_act + off,
          make_float4(gelu_erf(o.x), gelu_erf(o.y), gelu_erf(o.z), gelu_erf(o.w)));
  }
}

// ================================================================================================
// Backward, dst-CSR pass (vector path): grad_x_r, per-edge scalars, partial grad_att / grad_bias
// ================================================================================================
template <int VEC, int CV>
__global__ void __launch_bounds__(256) gatv2_bwd_dst_vec_kernel(const GatParams p) {
  const uint64_t seed_eff = p.training ? gat_seed(p) : 0;
  using HM = HeadMap<VEC, CV>;
  constexpr int S = HM::kSlots;
  constexpr int CH = Chunk<VEC>::value;
  constexpr int F4 = VEC * 32;
  __shared__ float4 red[2][8][F4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * 8;

  float4 a4[VEC], gatt[VEC], gbias[VEC], b4[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const int off = (v * 32 + lane) * 4;
    a4[v] = ldg4(p.att + off);
    b4[v] = p.bias ? ldg4(p.bias + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    gatt[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    gbias[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp; row < p.n_dst; row += warps_total) {
    float4 g4[VEC], r4[VEC], gr[VEC];
    float part[VEC], cdot[S], m[S], inv[S];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int off = (v * 32 + lane) * 4;
      g4[v] = ldg4(p.grad_out + row * p.ld_g + off);
      const float4 o4 = ldg4(p.out + row * p.ld_out + off);
      if (p.gelu_fused) {
        g4[v].x *= gelu_erf_grad(o4.x); g4[v].y *= gelu_erf_grad(o4.y);
        g4[v].z *= gelu_erf_grad(o4.z); g4[v].w *= gelu_erf_grad(o4.w);
        st4(p.g_buf + row * p.ld_g + off, g4[v]);
      }
      r4[v] = ldg4(p.x_r + row * p.ld_r + off);
      gr[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      gbias[v] = add4(gbias[v], g4[v]);
      part[v] = dot4(g4[v], make_float4(o4.x - b4[v].x, o4.y - b4[v].y, o4.z - b4[v].z, o4.w - b4[v].w));
    }
    HM::reduce(part, cdot);   // c_i = g_i . o_i per head
#pragma unroll
    for (int q = 0; q < S; ++q) {
      const int h = HM::head(q, lane);
      m[q] = __ldg(p.stat_max + row * p.H + h);
      inv[q] = 1.0f / __ldg(p.stat_den + row * p.H + h);
    }
    const int beg = p.rowptr[row], end = p.rowptr[row + 1];
    for (int base = beg; base < end; base += CH) {
      const int n = min(CH, end - base);
      int mycol = 0, myeid = 0;
      if (lane < n) {
        mycol = __ldg(p.col + base + lane);
        if (p.training) myeid = __ldg(p.eid + base + lane);
      }
      float4 x[CH][VEC];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int j = __shfl_sync(kFull, mycol, c);
        if (c < n) {
          const float* src = p.x_l + static_cast<int64_t>(j) * p.ld_l;
#pragma unroll
          for (int v = 0; v < VEC; ++v) x[c][v] = ldg4(src + (v * 32 + lane) * 4);
        }
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (c < n) {
          const int e = __shfl_sync(kFull, myeid, c);
          float4 z[VEC];
          float pl[VEC], pd[VEC], lg[S], dd[S];
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            z[v] = add4(x[c][v], r4[v]);
            pl[v] = dot4(a4[v], lrelu4(z[v], p.slope));
            pd[v] = dot4(g4[v], x[c][v]);
          }
          HM::reduce(pl, lg);
          HM::reduce(pd, dd);
          float delta[S];
#pragma unroll
          for (int q = 0; q < S; ++q) {
            const int h = HM::head(q, lane);
            const float alpha = __expf(lg[q] - m[q]) * inv[q];
            float ks = 1.0f;
            if (p.training) ks = rng_keep(seed_eff, e, p.H, h, p.drop_thr) ? p.keep_scale : 0.f;
            delta[q] = alpha * (dd[q] * ks - cdot[q]);
            if (HM::writer(lane)) {
              const int64_t idx = static_cast<int64_t>(base + c) * p.H + h;
              p.e_delta[idx] = delta[q];
              p.e_alpha[idx] = alpha * ks;
            }
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float d = delta[HM::slot_of(v)];
            const float4 zz = z[v];
            const float4 dz = make_float4(d * a4[v].x * (zz.x > 0.f ? 1.f : p.slope), d * a4[v].y * (zz.y > 0.f ? 1.f : p.slope),
                                          d * a4[v].z * (zz.z > 0.f ? 1.f : p.slope), d * a4[v].w * (zz.w > 0.f ? 1.f : p.slope));
            gr[v] = add4(gr[v], dz);
            fma4(gatt[v], d, lrelu4(zz, p.slope));
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) st4(p.grad_x_r + row * p.ld_gr + (v * 32 + lane) * 4, gr[v]);
  }

  // CTA-level ordered reduction of the per-warp accumulators -> partial[blockIdx][2][F]
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    red[0][warp][v * 32 + lane] = gatt[v];
    red[1][warp][v * 32 + lane] = gbias[v];
  }
  __syncthreads();
  if (threadIdx.x < 2 * F4) {
    const int which = threadIdx.x / F4, f = threadIdx.x % F4;
    float4 t = red[which][0][f];
#pragma unroll
    for (int w = 1; w < 8; ++w) t = add4(t, red[which][w][f]);
    st4(p.partial + (static_cast<int64_t>(blockIdx.x) * 2 + which) * (F4 * 4) + f * 4, t);
  }
}

// out[c] = sum_b partial[b][c], fixed order.  cols = 2F (att | bias)
__global__ void colsum_partials_kernel(const float* __restrict__ partial, int64_t nrows, int cols, int F,
                                       float* __restrict__ grad_att, float* __restrict__ grad_bias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int64_t b = 0;
  for (; b + 4 <= nrows; b += 4) {
    s0 += partial[(b + 0) * cols + c];
    s1 += partial[(b + 1) * cols + c];
    s2 += partial[(b + 2) * cols + c];
    s3 += partial[(b + 3) * cols + c];
  }
  for (; b < nrows; ++b) s0 += partial[b * cols + c];
  const float s = (s0 + s1) + (s2 + s3);
  if (c < F) grad_att[c] = s;
  else if (grad_bias) grad_bias[c - F] = s;
}

// ================================================================================================
// Backward, src-CSR (transposed) pass (vector path): grad_x_l
// ================================================================================================
template <int VEC, int CV>
__global__ void __launch_bounds__(256) gatv2_bwd_src_vec_kernel(const GatParams p) {
  using HM = HeadMap<VEC, CV>;
  constexpr int CH = Chunk<VEC>::value >= 4 ? Chunk<VEC>::value / 2 : 2;   // two gathered rows per edge
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= p.n_src) return;
  const int beg = p.t_rowptr[row], end = p.t_rowptr[row + 1];
  float4 acc[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (beg < end) {
    float4 a4[VEC], l4[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int off = (v * 32 + lane) * 4;
      a4[v] = ldg4(p.att + off);
      l4[v] = ldg4(p.x_l + row * p.ld_l + off);
    }
    const float* gsrc = p.gelu_fused ? p.g_buf : p.grad_out;
    for (int base = beg; base < end; base += CH) {
      const int n = min(CH, end - base);
      int mydst = 0, mypos = 0;
      if (lane < n) {
        mydst = __ldg(p.t_dst + base + lane);
        mypos = __ldg(p.t_pos + base + lane);
      }
      float4 r[CH][VEC], g[CH][VEC];
      float dl[CH][VEC], al[CH][VEC];
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int i = __shfl_sync(kFull, mydst, c);
        const int pos = __shfl_sync(kFull, mypos, c);
        if (c < n) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const int off = (v * 32 + lane) * 4;
            r[c][v] = ldg4(p.x_r + static_cast<int64_t>(i) * p.ld_r + off);
            g[c][v] = ldg4(gsrc + static_cast<int64_t>(i) * p.ld_g + off);
            const int64_t idx = static_cast<int64_t>(pos) * p.H + HM::head(HM::slot_of(v), lane);
            dl[c][v] = __ldg(p.e_delta + idx);
            al[c][v] = __ldg(p.e_alpha + idx);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (c < n) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            const float4 z = add4(l4[v], r[c][v]);
            const float d = dl[c][v];
            acc[v].x += d * a4[v].x * (z.x > 0.f ? 1.f : p.slope);
            acc[v].y += d * a4[v].y * (z.y > 0.f ? 1.f : p.slope);
            acc[v].z += d * a4[v].z * (z.z > 0.f ? 1.f : p.slope);
            acc[v].w += d * a4[v].w * (z.w > 0.f ? 1.f : p.slope);
            fma4(acc[v], al[c][v], g[c][v]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) st4(p.grad_x_l + row * p.ld_gl + (v * 32 + lane) * 4, acc[v]);
}

// ================================================================================================
// Generic path: one warp per (row, head); lane handles channels c = lane + 32*t, t < TC.
// ================================================================================================
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
  return x;
}

template <int TC>
__global__ void __launch_bounds__(256) gatv2_fwd_gen_kernel(const GatParams p) {
  const uint64_t seed_eff = p.training ? gat_seed(p) : 0;
  const int lane = threadIdx.x & 31;
  const int64_t item = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (item >= p.n_dst * p.H) return;
  const int64_t row = item / p.H;
  const int h = static_cast<int>(item % p.H);
  const int C = p.C;
  const int fo = h * C;
  float a[TC], r[TC], acc[TC];
#pragma unroll
  for (int t = 0; t < TC; ++t) {
    const int c = lane + 32 * t;
    a[t] = c < C ? __ldg(p.att + fo + c) : 0.f;
    r[t] = c < C ? __ldg(p.x_r + row * p.ld_r + fo + c) : 0.f;
    acc[t] = 0.f;
  }
  float m = -INFINITY, s = 0.f;
  const int beg = p.rowptr[row], end = p.rowptr[row + 1];
  for (int k = beg; k < end; ++k) {
    const int j = __ldg(p.col + k);
    float x[TC], part = 0.f;
#pragma unroll
    for (int t = 0; t < TC; ++t) {
      const int c = lane + 32 * t;
      x[t] = c < C ? __ldg(p.x_l + static_cast<int64_t>(j) * p.ld_l + fo + c) : 0.f;
      part += a[t] * lrelu(x[t] + r[t], p.slope);
    }
    const float lg = warp_sum(part);
    const float mn = fmaxf(m, lg);
    const float sc = __expf(m - mn);
    const float pe = __expf(lg - mn);
    s = s * sc + pe;
    float w = pe;
    if (p.training) w = rng_keep(seed_eff, __ldg(p.eid + k), p.H, h, p.drop_thr) ? pe * p.keep_scale : 0.f;
#pragma unroll
    for (int t = 0; t < TC; ++t) acc[t] = acc[t] * sc + w * x[t];
    m = mn;
  }
  const float den = s + 1e-16f;
  const float inv = 1.0f / den;
  if (lane == 0) {
    p.stat_max[row * p.H + h] = (beg == end) ? 0.f : m;
    p.stat_den[row * p.H + h] = den;
  }
#pragma unroll
  for (int t = 0; t < TC; ++t) {
    const int c = lane + 32 * t;
    if (c < C) {
      float o = acc[t] * inv;
      if (p.bias) o += __ldg(p.bias + fo + c);
      if (p.out) p.out[row * p.ld_out + fo + c] = o;
      if (p.out_act) p.out_act[row * p.ld_act + fo + c] = gelu_erf(o);
    }
  }
}

// warps_total must be a multiple of H so that every warp keeps one head for its whole life.
template <int TC>
__global__ void __launch_bounds__(256) gatv2_bwd_dst_gen_kernel(const GatParams p) {
  const uint64_t seed_eff = p.training ? gat_seed(p) : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t wg = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * 8;
  const int h = static_cast<int>(wg % p.H);
  const int C = p.C, fo = h * C;
  float a[TC], b[TC], gatt[TC], gbias[TC];
#pragma unroll
  for (int t = 0; t < TC; ++t) {
    const int c = lane + 32 * t;
    a[t] = c < C ? __ldg(p.att + fo + c) : 0.f;
    b[t] = (c < C && p.bias) ? __ldg(p.bias + fo + c) : 0.f;
    gatt[t] = 0.f;
    gbias[t] = 0.f;
  }
  const int64_t items = p.n_dst * p.H;
  for (int64_t item = wg; item < items; item += warps_total) {
    const int64_t row = item / p.H;
    float g[TC], r[TC], gr[TC], part = 0.f;
#pragma unroll
    for (int t = 0; t < TC; ++t) {
      const int c = lane + 32 * t;
      g[t] = 0.f; r[t] = 0.f; gr[t] = 0.f;
      if (c < C) {
        g[t] = __ldg(p.grad_out + row * p.ld_g + fo + c);
        const float o = __ldg(p.out + row * p.ld_out + fo + c);
        if (p.gelu_fused) {
          g[t] *= gelu_erf_grad(o);
          p.g_buf[row * p.ld_g + fo + c] = g[t];
        }
        r[t] = __ldg(p.x_r + row * p.ld_r + fo + c);
        gbias[t] += g[t];
        part += g[t] * (o - b[t]);
      }
    }
    const float cdot = warp_sum(part);
    const float m = __ldg(p.stat_max + row * p.H + h);
    const float inv = 1.0f / __ldg(p.stat_den + row * p.H + h);
    const int beg = p.rowptr[row], end = p.rowptr[row + 1];
    for (int k = beg; k < end; ++k) {
      const int j = __ldg(p.col + k);
      float x[TC], z[TC], pl = 0.f, pd = 0.f;
#pragma unroll
      for (int t = 0; t < TC; ++t) {
        const int c = lane + 32 * t;
        x[t] = c < C ? __ldg(p.x_l + static_cast<int64_t>(j) * p.ld_l + fo + c) : 0.f;
        z[t] = x[t] + r[t];
        pl += a[t] * lrelu(z[t], p.slope);
        pd += g[t] * x[t];
      }
      const float lg = warp_sum(pl), dd = warp_sum(pd);
      const float alpha = __expf(lg - m) * inv;
      float ks = 1.0f;
      if (p.training) ks = rng_keep(seed_eff, __ldg(p.eid + k), p.H, h, p.drop_thr) ? p.keep_scale : 0.f;
      const float delta = alpha * (dd * ks - cdot);
      if (lane == 0) {
        p.e_delta[static_cast<int64_t>(k) * p.H + h] = delta;
        p.e_alpha[static_cast<int64_t>(k) * p.H + h] = alpha * ks;
      }
#pragma unroll
      for (int t = 0; t < TC; ++t) {
        gr[t] += delta * a[t] * (z[t] > 0.f ? 1.f : p.slope);
        gatt[t] += delta * lrelu(z[t], p.slope);
      }
    }
#pragma unroll
    for (int t = 0; t < TC; ++t) {
      const int c = lane + 32 * t;
      if (c < C) p.grad_x_r[row * p.ld_gr + fo + c] = gr[t];
    }
  }
  // per-warp partials: partial[wg][2][C]
#pragma unroll
  for (int t = 0; t < TC; ++t) {
    const int c = lane + 32 * t;
    if (c < C) {
      p.partial[(wg * 2 + 0) * C + c] = gatt[t];
      p.partial[(wg * 2 + 1) * C + c] = gbias[t];
    }
  }
}

// generic second stage: column (h, c) = ordered sum over warps wg with wg % H == h
__global__ void colsum_partials_gen_kernel(const float* __restrict__ partial, int64_t warps_total, int H, int C,
                                           float* __restrict__ grad_att, float* __restrict__ grad_bias) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= H * C) return;
  const int h = f / C, c = f % C;
  float sa = 0.f, sb = 0.f;
  for (int64_t wg = h; wg < warps_total; wg += H) {
    sa += partial[(wg * 2 + 0) * C + c];
    sb += partial[(wg * 2 + 1) * C + c];
  }
  grad_att[f] = sa;
  if (grad_bias) grad_bias[f] = sb;
}

template <int TC>
__global__ void __launch_bounds__(256) gatv2_bwd_src_gen_kernel(const GatParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t item = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (item >= p.n_src * p.H) return;
  const int64_t row = item / p.H;
  const int h = static_cast<int>(item % p.H);
  const int C = p.C, fo = h * C;
  const float* gsrc = p.gelu_fused ? p.g_buf : p.grad_out;
  float a[TC], l[TC], acc[TC];
#pragma unroll
  for (int t = 0; t < TC; ++t) {
    const int c = lane + 32 * t;
    a[t] = c < C ? __ldg(p.att + fo + c) : 0.f;
    l[t] = c < C ? __ldg(p.x_l + row * p.ld_l + fo + c) : 0.f;
    acc[t] = 0.f;
  }
  const int beg = p.t_rowptr[row], end = p.t_rowptr[row + 1];
  for (int k = beg; k < end; ++k) {
    const int i = __ldg(p.t_dst + k);
    const int64_t idx = static_cast<int64_t>(__ldg(p.t_pos + k)) * p.H + h;
    const float d = __ldg(p.e_delta + idx), al = __ldg(p.e_alpha + idx);
#pragma unroll
    for (int t = 0; t < TC; ++t) {
      const int c = lane + 32 * t;
      if (c < C) {
        const float z = l[t] + __ldg(p.x_r + static_cast<int64_t>(i) * p.ld_r + fo + c);
        const float g = __ldg(gsrc + static_cast<int64_t>(i) * p.ld_g + fo + c);
        acc[t] += d * a[t] * (z > 0.f ? 1.f : p.slope) + al * g;
      }
    }
  }
#pragma unroll
  for (int t = 0; t < TC; ++t) {
    const int c = lane + 32 * t;
    if (c < C) p.grad_x_l[row * p.ld_gl + fo + c] = acc[t];
  }
}

// alpha [E,H] in original edge order (pre-dropout), warp per (row, head)
__global__ void __launch_bounds__(256) gatv2_alpha_kernel(const GatParams p, float* __restrict__ alpha_out) {
  const int lane = threadIdx.x & 31;
  const int64_t item = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (item >= p.n_dst * p.H) return;
  const int64_t row = item / p.H;
  const int h = static_cast<int>(item % p.H);
  const int C = p.C, fo = h * C;
  const float m = __ldg(p.stat_max + row * p.H + h);
  const float inv = 1.0f / __ldg(p.stat_den + row * p.H + h);
  const int beg = p.rowptr[row], end = p.rowptr[row + 1];
  for (int k = beg; k < end; ++k) {
    const int j = __ldg(p.col + k);
    float part = 0.f;
    for (int c = lane; c < C; c += 32)
      part += __ldg(p.att + fo + c) *
              lrelu(__ldg(p.x_l + static_cast<int64_t>(j) * p.ld_l + fo + c) + __ldg(p.x_r + row * p.ld_r + fo + c), p.slope);
    const float lg = warp_sum(part);
    if (lane == 0) alpha_out[static_cast<int64_t>(__ldg(p.eid + k)) * p.H + h] = __expf(lg - m) * inv;
  }
}

__global__ void dropout_mask_kernel(uint64_t seed, int64_t n, int H, uint32_t thr, uint8_t* __restrict__ mask) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) mask[i] = rng_keep(seed, i / H, H, static_cast<int>(i % H), thr) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
enum class Path { kVec, kGen, kNone };

struct Shape {
  Path path;
  int vec, cv, tc;
};

#define SGB_VEC_COMBOS(X) \
  X(1, 8) X(1, 16) X(1, 32) X(2, 8) X(2, 16) X(2, 32) X(2, 64) X(3, 32) X(4, 16) X(4, 32) X(4, 64) X(4, 128)

Shape classify(int H, int C, bool aligned) {
  const int F = H * C;
  if (aligned && F % 128 == 0 && C % 4 == 0) {
    const int vec = F / 128, cv = C / 4;
#define X(V, Cv) if (vec == V && cv == Cv) return {Path::kVec, V, Cv, 0};
    SGB_VEC_COMBOS(X)
#undef X
  }
  if (C <= 256) {
    const int tc = C <= 32 ? 1 : (C <= 64 ? 2 : (C <= 128 ? 4 : 8));
    return {Path::kGen, 0, 0, tc};
  }
  return {Path::kNone, 0, 0, 0};
}

bool ld_ok(int64_t ld) { return ld % 4 == 0; }

// SEGGER_B200_GAT=legacy forces the first-generation row-per-warp kernels (debug / A-B profiling)
bool legacy_path() {
  const char* e = getenv("SEGGER_B200_GAT");   // read per call so a profiling script can A/B in one process
  return e && e[0] == 'l';
}

int gen_bwd_blocks(int H) {  // CTA count for the persistent generic dst pass: warps_total % H == 0
  int nb = sm_count() * 2;
  nb = (nb + H - 1) / H * H;
  return nb;
}
int vec_bwd_blocks(int64_t n_dst) {
  const int64_t want = ceil_div(n_dst > 0 ? n_dst : 1, 8);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  return static_cast<int>(want < cap ? want : cap);
}

}  // namespace
}  // namespace sgb

using namespace sgb;

static int validate_common(const char* fn, const float* x_l, const float* x_r, const float* att, int64_t n_dst,
                           int64_t E, int H, int C) {
  SGB_REQUIRE(H >= 1 && C >= 1, SGB_ERR_ARG, "%s: H and C must be >= 1", fn);
  SGB_REQUIRE(n_dst >= 0 && n_dst < (int64_t(1) << 31) && E >= 0 && E < (int64_t(1) << 31), SGB_ERR_RANGE,
              "%s: n_dst/E exceed 2^31-1 per call", fn);
  SGB_REQUIRE(att && (n_dst == 0 || x_r) && (E == 0 || x_l), SGB_ERR_ARG, "%s: null tensor", fn);
  return SGB_OK;
}

extern "C" int sgb_gatv2_fwd(const float* x_l, int64_t ld_l, const float* x_r, int64_t ld_r, const float* att,
                             const float* bias, const int32_t* dst_rowptr, const int32_t* dst_col,
                             const int32_t* dst_eid, int64_t n_dst, int64_t E, int H, int C, float negative_slope,
                             float p_drop, uint64_t seed, const uint64_t* seed_dev, int training, float* out, int64_t ld_out, float* out_act,
                             int64_t ld_act, float* stat_max, float* stat_den, float* e_logit, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = validate_common("gatv2_fwd", x_l, x_r, att, n_dst, E, H, C);
  if (rc != SGB_OK) return rc;
  if (n_dst == 0) return SGB_OK;
  SGB_REQUIRE(dst_rowptr && (out || out_act) && stat_max && stat_den && (E == 0 || dst_col), SGB_ERR_ARG, "gatv2_fwd: null argument");
  const bool train = training && p_drop > 0.f;
  SGB_REQUIRE(!train || E == 0 || dst_eid, SGB_ERR_ARG, "gatv2_fwd: training dropout needs dst_eid");
  SGB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, SGB_ERR_ARG, "gatv2_fwd: dropout p must be in [0,1)");
  const bool aligned = aligned16(x_l) && aligned16(x_r) && aligned16(att) && aligned16(bias) && aligned16(out) &&
                       aligned16(out_act) && ld_ok(ld_l) && ld_ok(ld_r) && (!out || ld_ok(ld_out)) && (!out_act || ld_ok(ld_act));
  const Shape sh = classify(H, C, aligned);
  SGB_REQUIRE(sh.path != Path::kNone, SGB_ERR_ARG, "gatv2_fwd: unsupported shape H=%d C=%d (aligned=%d)", H, C, (int)aligned);
  GatParams p{};
  p.x_l = x_l; p.x_r = x_r; p.att = att; p.bias = bias; p.ld_l = ld_l; p.ld_r = ld_r;
  p.rowptr = dst_rowptr; p.col = dst_col; p.eid = dst_eid; p.n_dst = n_dst; p.H = H; p.C = C;
  p.slope = negative_slope; p.training = train ? 1 : 0; p.drop_thr = drop_threshold(p_drop);
  p.keep_scale = 1.0f / (1.0f - p_drop); p.seed = seed; p.seed_dev = seed_dev;
  p.out = out; p.out_act = out_act; p.ld_out = ld_out; p.ld_act = ld_act; p.stat_max = stat_max; p.stat_den = stat_den;
  p.e_logit = e_logit;
  if (aligned && !legacy_path() && quad_fwd_launch(p, stream)) return check_launch("gatv2_fwd(quad)");
  SGB_REQUIRE(!e_logit, SGB_ERR_ARG, "gatv2_fwd: e_logit is written by the sub-warp kernels only (sgb_gatv2_quad_supported, "
              "16-byte aligned operands)");
  if (sh.path == Path::kVec) {
    const unsigned blocks = static_cast<unsigned>(ceil_div(n_dst, 8));
#define X(V, Cv) if (sh.vec == V && sh.cv == Cv) gatv2_fwd_vec_kernel<V, Cv><<<blocks, 256, 0, stream>>>(p);
    SGB_VEC_COMBOS(X)
#undef X
  } else {
    const unsigned blocks = static_cast<unsigned>(ceil_div(n_dst * H, 8));
    switch (sh.tc) {
      case 1: gatv2_fwd_gen_kernel<1><<<blocks, 256, 0, stream>>>(p); break;
      case 2: gatv2_fwd_gen_kernel<2><<<blocks, 256, 0, stream>>>(p); break;
      case 4: gatv2_fwd_gen_kernel<4><<<blocks, 256, 0, stream>>>(p); break;
      default: gatv2_fwd_gen_kernel<8><<<blocks, 256, 0, stream>>>(p); break;
    }
  }
  return check_launch("gatv2_fwd");
}

extern "C" int sgb_gatv2_quad_supported(int H, int C) {
  GatParams p{};
  p.H = H; p.C = C; p.slope = 0.2f;
  return quad_supported(p) ? 1 : 0;
}

extern "C" int sgb_gatv2_alpha(const float* x_l, int64_t ld_l, const float* x_r, int64_t ld_r, const float* att,
                               const int32_t* dst_rowptr, const int32_t* dst_col, const int32_t* dst_eid,
                               int64_t n_dst, int64_t E, int H, int C, float negative_slope, const float* stat_max,
                               const float* stat_den, float* alpha, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = validate_common("gatv2_alpha", x_l, x_r, att, n_dst, E, H, C);
  if (rc != SGB_OK) return rc;
  if (n_dst == 0 || E == 0) return SGB_OK;
  SGB_REQUIRE(dst_rowptr && dst_col && dst_eid && stat_max && stat_den && alpha, SGB_ERR_ARG, "gatv2_alpha: null argument");
  GatParams p{};
  p.x_l = x_l; p.x_r = x_r; p.att = att; p.ld_l = ld_l; p.ld_r = ld_r; p.rowptr = dst_rowptr; p.col = dst_col;
  p.eid = dst_eid; p.n_dst = n_dst; p.H = H; p.C = C; p.slope = negative_slope;
  p.stat_max = const_cast<float*>(stat_max); p.stat_den = const_cast<float*>(stat_den);
  SGB_REQUIRE(C <= (1 << 20), SGB_ERR_ARG, "gatv2_alpha: C out of range");
  gatv2_alpha_kernel<<<static_cast<unsigned>(ceil_div(n_dst * H, 8)), 256, 0, stream>>>(p, alpha);
  return check_launch("gatv2_alpha");
}

static size_t bwd_partial_floats(int64_t n_dst, int H, int C) {
  const int F = H * C;
  const size_t vec = static_cast<size_t>(vec_bwd_blocks(n_dst)) * 2 * F;
  const size_t gen = static_cast<size_t>(gen_bwd_blocks(H)) * 8 * 2 * C;
  const size_t quad = quad_bwd_partial_floats(H, C);
  const size_t m = vec > gen ? vec : gen;
  return m > quad ? m : quad;
}
// bytes of ONE of the two per-edge scalar arrays; two of them also hold the quad path's [E][rec_stride] records
static size_t bwd_edge_bytes(int64_t E, int H) {
  return align_up(static_cast<size_t>(E > 0 ? E : 1) * (rec_stride(H) / 2) * sizeof(float));
}

// the two per-edge arrays double as the sub-warp path's record array: [E][quad_bwd_record_bytes]
static size_t bwd_edge_region(int64_t E, int H, int C) {
  const size_t two = 2 * bwd_edge_bytes(E, H);
  const size_t quad = align_up(static_cast<size_t>(E > 0 ? E : 1) * quad_bwd_record_bytes(H, C));
  return two > quad ? two : quad;
}

extern "C" size_t sgb_gatv2_bwd_workspace_bytes(int64_t n_dst, int64_t E, int H, int C) {
  return bwd_edge_region(E, H, C) + align_up(bwd_partial_floats(n_dst, H, C) * sizeof(float));
}

extern "C" int sgb_gatv2_bwd(const float* x_l, int64_t ld_l, const float* x_r, int64_t ld_r, const float* att,
                             const float* bias, const float* out, int64_t ld_out, const float* grad_out, int64_t ld_g,
                             int gelu_fused, float* g_buf, const int32_t* dst_rowptr, const int32_t* dst_col,
                             const int32_t* dst_eid, const int32_t* src_rowptr, const int32_t* src_dst,
                             const int32_t* src_pos, int64_t n_src, int64_t n_dst, int64_t E, int H, int C,
                             float negative_slope, float p_drop, uint64_t seed, const uint64_t* seed_dev, int training, const float* stat_max,
                             const float* stat_den, const float* e_logit, float* grad_x_l, int64_t ld_gl, float* grad_x_r, int64_t ld_gr,
                             float* grad_att, float* grad_bias, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = validate_common("gatv2_bwd", x_l, x_r, att, n_dst, E, H, C);
  if (rc != SGB_OK) return rc;
  SGB_REQUIRE(n_src >= 0 && n_src < (int64_t(1) << 31), SGB_ERR_RANGE, "gatv2_bwd: n_src out of range");
  SGB_REQUIRE(grad_att && (n_src == 0 || grad_x_l) && (n_dst == 0 || (grad_x_r && out && grad_out && stat_max && stat_den)),
              SGB_ERR_ARG, "gatv2_bwd: null argument");
  // src_rowptr == NULL: "one source per edge" -- dst_col is a permutation of [0, E) (n_src == E), so every grad_x_l row has
  // exactly one contribution and the dst pass writes it directly (no transposed pass, no per-edge records read back)
  const bool direct_src = src_rowptr == nullptr;
  SGB_REQUIRE(dst_rowptr && (E == 0 || dst_col) && (direct_src ? n_src == E : (E == 0 || (src_dst && src_pos))), SGB_ERR_ARG,
              "gatv2_bwd: null CSR");
  SGB_REQUIRE(!gelu_fused || g_buf, SGB_ERR_ARG, "gatv2_bwd: gelu_fused requires g_buf");
  SGB_REQUIRE(ws && ws_bytes >= sgb_gatv2_bwd_workspace_bytes(n_dst, E, H, C), SGB_ERR_WORKSPACE, "gatv2_bwd: workspace too small");
  const bool train = training && p_drop > 0.f;
  SGB_REQUIRE(!train || E == 0 || dst_eid, SGB_ERR_ARG, "gatv2_bwd: training dropout needs dst_eid");
  const int F = H * C;
  const bool aligned = aligned16(x_l) && aligned16(x_r) && aligned16(att) && aligned16(bias) && aligned16(out) &&
                       aligned16(grad_out) && aligned16(g_buf) && aligned16(grad_x_l) && aligned16(grad_x_r) &&
                       ld_ok(ld_l) && ld_ok(ld_r) && ld_ok(ld_out) && ld_ok(ld_g) && ld_ok(ld_gl) && ld_ok(ld_gr);
  const Shape sh = classify(H, C, aligned);
  SGB_REQUIRE(sh.path != Path::kNone, SGB_ERR_ARG, "gatv2_bwd: unsupported shape H=%d C=%d (aligned=%d)", H, C, (int)aligned);

  const size_t edge = bwd_edge_bytes(E, H);
  char* w = static_cast<char*>(ws);
  GatParams p{};
  p.x_l = x_l; p.x_r = x_r; p.att = att; p.bias = bias; p.ld_l = ld_l; p.ld_r = ld_r;
  p.rowptr = dst_rowptr; p.col = dst_col; p.eid = dst_eid; p.n_dst = n_dst; p.n_src = n_src; p.H = H; p.C = C;
  p.slope = negative_slope; p.training = train ? 1 : 0; p.drop_thr = drop_threshold(p_drop);
  p.keep_scale = 1.0f / (1.0f - p_drop); p.seed = seed; p.seed_dev = seed_dev;
  p.out = const_cast<float*>(out); p.ld_out = ld_out;
  p.stat_max = const_cast<float*>(stat_max); p.stat_den = const_cast<float*>(stat_den);
  p.grad_out = grad_out; p.ld_g = ld_g; p.gelu_fused = gelu_fused; p.g_buf = g_buf;
  p.e_delta = reinterpret_cast<float*>(w); p.e_alpha = reinterpret_cast<float*>(w + edge);
  p.partial = reinterpret_cast<float*>(w + bwd_edge_region(E, H, C));
  p.grad_x_l = grad_x_l; p.grad_x_r = grad_x_r; p.ld_gl = ld_gl; p.ld_gr = ld_gr;
  p.t_rowptr = src_rowptr; p.t_dst = src_dst; p.t_pos = src_pos;
  p.e_logit = const_cast<float*>(e_logit);

  if (n_dst > 0 && aligned && !legacy_path() && quad_bwd_launch(p, grad_att, grad_bias, stream))
    return check_launch("gatv2_bwd(quad)");
  SGB_REQUIRE(!e_logit || n_dst == 0, SGB_ERR_ARG, "gatv2_bwd: e_logit is read by the sub-warp kernels only");
  SGB_REQUIRE(!direct_src, SGB_ERR_ARG, "gatv2_bwd: the one-source-per-edge form (src_rowptr == NULL) needs a shape covered by "
              "the sub-warp kernels (sgb_gatv2_quad_supported)");
  if (n_dst == 0) {
    cudaMemsetAsync(grad_att, 0, sizeof(float) * F, stream);
    if (grad_bias) cudaMemsetAsync(grad_bias, 0, sizeof(float) * F, stream);
  } else if (sh.path == Path::kVec) {
    const int nb = vec_bwd_blocks(n_dst);
#define X(V, Cv) if (sh.vec == V && sh.cv == Cv) gatv2_bwd_dst_vec_kernel<V, Cv><<<nb, 256, 0, stream>>>(p);
    SGB_VEC_COMBOS(X)
#undef X
    colsum_partials_kernel<<<static_cast<unsigned>(ceil_div(2 * F, 128)), 128, 0, stream>>>(p.partial, nb, 2 * F, F, grad_att, grad_bias);
  } else {
    const int nb = gen_bwd_blocks(H);
    switch (sh.tc) {
      case 1: gatv2_bwd_dst_gen_kernel<1><<<nb, 256, 0, stream>>>(p); break;
      case 2: gatv2_bwd_dst_gen_kernel<2><<<nb, 256, 0, stream>>>(p); break;
      case 4: gatv2_bwd_dst_gen_kernel<4><<<nb, 256, 0, stream>>>(p); break;
      default: gatv2_bwd_dst_gen_kernel<8><<<nb, 256, 0, stream>>>(p); break;
    }
    colsum_partials_gen_kernel<<<static_cast<unsigned>(ceil_div(F, 128)), 128, 0, stream>>>(
        p.partial, static_cast<int64_t>(nb) * 8, H, C, grad_att, grad_bias);
  }
  if (n_src > 0) {
    if (sh.path == Path::kVec) {
      const unsigned blocks = static_cast<unsigned>(ceil_div(n_src, 8));
#define X(V, Cv) if (sh.vec == V && sh.cv == Cv) gatv2_bwd_src_vec_kernel<V, Cv><<<blocks, 256, 0, stream>>>(p);
      SGB_VEC_COMBOS(X)
#undef X
    } else {
      const unsigned blocks = static_cast<unsigned>(ceil_div(n_src * H, 8));
      switch (sh.tc) {
        case 1: gatv2_bwd_src_gen_kernel<1><<<blocks, 256, 0, stream>>>(p); break;
        case 2: gatv2_bwd_src_gen_kernel<2><<<blocks, 256, 0, stream>>>(p); break;
        case 4: gatv2_bwd_src_gen_kernel<4><<<blocks, 256, 0, stream>>>(p); break;
        default: gatv2_bwd_src_gen_kernel<8><<<blocks, 256, 0, stream>>>(p); break;
      }
    }
  }
  return check_launch("gatv2_bwd");
}

extern "C" int sgb_dropout_mask(uint64_t seed, int64_t E, int H, float p_drop, uint8_t* mask, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SGB_REQUIRE(E >= 0 && H >= 1 && (E == 0 || mask), SGB_ERR_ARG, "dropout_mask: bad argument");
  SGB_REQUIRE(p_drop >= 0.f && p_drop < 1.f, SGB_ERR_ARG, "dropout_mask: p must be in [0,1)");
  if (E == 0) return SGB_OK;
  const int64_t n = E * H;
  dropout_mask_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, stream>>>(seed, n, H, drop_threshold(p_drop), mask);
  return check_launch("dropout_mask");
}
