// Fused GATv2 kernels, sub-warp-per-row ("quad") path with a cp.async shared-memory gather pipeline.
//
// Same math as sgb_gatv2.cu (replaces the PyG GATv2Conv message passing
// configured at /root/reference/src/segger/models/ist_encoder.py:111-131), restructured for the
// instruction-issue and latency limits the first kernels hit on B200 (profiles/r1_summary.md):
//
//  * LPR lanes own one row (LPR = 8 for F = 128: a warp advances G = 4 destination rows in lock
//    step), each lane holds V float4 of the row (float4 index t*LPR + s).  All per-edge scalar work
//    (logit reduction, exp, dropout hash, softmax bookkeeping, loop control) is paid once per
//    warp instruction for G edges, and the head reduction is log2(LPR) shuffles instead of 4-5.
//  * every lane stages exactly the float4s it will consume itself through a D-deep ring in shared
//    memory with cp.async (LDGSTS, 16 B, L2-only): the gathers of step K+D-1 are in flight while
//    step K is computed, across row boundaries, with no register cost and no cross-lane
//    synchronisation (a lane only reads back its own copies, so cp.async.wait_group suffices).
//    Destination-side rows (x_r, grad_out, out) travel through the same ring as "header" steps.
//  * the column index for step K+D is loaded one iteration before its copy is issued, so no
//    dependent global load sits on the issue path.
//  * LeakyReLU is max(z, slope*z) (FMUL + FMNMX per element, rounded like the reference's leaky_relu).
//  * online softmax with a lazily updated maximum: the running accumulator is only rescaled when a
//    logit exceeds the reference maximum by more than kTau (warp-uniform rare branch).  The saved
//    (stat_max, stat_den) pair is self-consistent, which is all the backward needs.
//  * row-owner reductions everywhere (dst pass: grad_x_r, src pass: grad_x_l), fixed summation order
//    -> bit-reproducible, no atomics.  The dst pass leaves one 16-byte record (delta, alpha') per
//    edge for the src pass; nothing of width C is ever written per edge.
#include "sgb_gatv2.cuh"

namespace sgb {
namespace {

constexpr int kQW = 4;                  // warps per CTA
constexpr int kQThreads = kQW * 32;
constexpr float kTau = 12.0f;           // lazy-max slack: exp(logit - m) <= e^12

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float4 sub4(const float4 a, const float4 b) {
  return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ float4 mul4(const float4 a, const float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
// LeakyReLU for 0 <= slope <= 1 as max(z, slope*z): FMUL + FMNMX, the product rounded per element like
// the reference's leaky_relu (a split  c1*sum(a z) + c2*sum(a |z|)  saves one instruction per element but
// measured no faster end to end, so the reference's own rounding is kept).
__device__ __forceinline__ float4 lrelu_max4(const float4 z, float slope) {
  return make_float4(fmaxf(z.x, slope * z.x), fmaxf(z.y, slope * z.y), fmaxf(z.z, slope * z.z), fmaxf(z.w, slope * z.w));
}

// Packed fp32 pairs (sm_100 FADD2 / FMUL2 / FFMA2: two IEEE round-to-nearest operations per issue slot).  These
// kernels are bound by instruction issue, not by the fp32 pipes, so every per-element add / multiply / fma of the
// per-edge step works on the (x, y) and (z, w) halves of a float4.
__device__ __forceinline__ float2 lo2(const float4 a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ float2 hi2(const float4 a) { return make_float2(a.z, a.w); }
__device__ __forceinline__ float4 cat2(const float2 a, const float2 b) { return make_float4(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ float4 add4p(const float4 a, const float4 b) {
  return cat2(__fadd2_rn(lo2(a), lo2(b)), __fadd2_rn(hi2(a), hi2(b)));
}
__device__ __forceinline__ float4 mul4p(const float4 a, const float4 b) {
  return cat2(__fmul2_rn(lo2(a), lo2(b)), __fmul2_rn(hi2(a), hi2(b)));
}
// acc += w * x
__device__ __forceinline__ void fma4p(float4& acc, const float w, const float4 x) {
  const float2 w2 = make_float2(w, w);
  acc = cat2(__ffma2_rn(w2, lo2(x), lo2(acc)), __ffma2_rn(w2, hi2(x), hi2(acc)));
}
// acc += a (.) b
__device__ __forceinline__ void fma4v(float4& acc, const float4 a, const float4 b) {
  acc = cat2(__ffma2_rn(lo2(a), lo2(b), lo2(acc)), __ffma2_rn(hi2(a), hi2(b), hi2(acc)));
}
__device__ __forceinline__ void scale4p(float4& a, const float s) {
  const float2 s2 = make_float2(s, s);
  a = cat2(__fmul2_rn(lo2(a), s2), __fmul2_rn(hi2(a), s2));
}
// two-lane running dot product: p += a (.) b over both halves; the caller adds p.x + p.y at the end
__device__ __forceinline__ void dot4p(float2& p, const float4 a, const float4 b) {
  p = __ffma2_rn(lo2(a), lo2(b), p);
  p = __ffma2_rn(hi2(a), hi2(b), p);
}
// lrelu(x + r) for 0 <= slope <= 1 as max(z, slope*z), packed add / multiply
__device__ __forceinline__ float4 lrelu_sum4p(const float4 x, const float4 r, const float slope) {
  const float2 s2 = make_float2(slope, slope);
  const float2 z0 = __fadd2_rn(lo2(x), lo2(r)), z1 = __fadd2_rn(hi2(x), hi2(r));
  const float2 m0 = __fmul2_rn(z0, s2), m1 = __fmul2_rn(z1, s2);
  return make_float4(fmaxf(z0.x, m0.x), fmaxf(z0.y, m0.y), fmaxf(z1.x, m1.x), fmaxf(z1.y, m1.y));
}

template <int LPR>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int o = LPR / 2; o >= 1; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
  return x;
}

// The rows of a warp's chunk live in lanes: lane l holds (first CSR position, degree, row offset inside the chunk) of
// the row it will hand to lane group l % G of quad l / G.  A quad advances its G rows in lock step for max-degree
// steps, so rows of unequal degree waste issue slots (kNN in-degrees: 1.32 x E/G steps in chunk order).  The chunk's
// rows are therefore dealt out in order of decreasing degree (1.06 x): results per row are unchanged (a row's edges
// keep their CSR order), only the grouping of rows into quads -- and with it the summation order of the
// grad_att / grad_bias partial sums -- differs.
constexpr int kNoRow = 0x7fffffff;   // row offset of an empty lane: wrow0 + kNoRow >= n for every n < 2^31

template <int LPR>
__device__ __forceinline__ void chunk_rows(const int32_t* __restrict__ rowptr, int64_t wrow0, int nrows, int64_t n, int lane,
                                           bool sort, int& rp_b, int& rp_d, int& rp_o) {
  constexpr int G = 32 / LPR;
  const int b = __ldg(rowptr + min(wrow0 + lane, n));
  const int e = __ldg(rowptr + min(wrow0 + lane + 1, n));
  const bool valid = lane < nrows;
  rp_b = b;
  rp_d = valid ? e - b : 0;
  rp_o = valid ? lane : kNoRow;
  // (rows of equal degree -- the out-degrees of a kNN graph -- have nothing to gain: skip)
  const bool ragged = __reduce_max_sync(kFull, valid ? rp_d : 0) != __reduce_min_sync(kFull, valid ? rp_d : 0x7fffffff);
  if (G > 1 && sort && nrows > G && ragged) {
    // distinct keys: decreasing (clamped) degree, then lane; empty lanes last.  Any permutation is correct.
    const uint32_t dk = static_cast<uint32_t>(min(rp_d, 0x3ffffff));
    const uint32_t key = valid ? ((0x3ffffffu - dk) << 5 | static_cast<uint32_t>(lane)) : (0xffffffe0u | static_cast<uint32_t>(lane));
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) rank += __shfl_sync(kFull, key, j) < key ? 1 : 0;
    int inv = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) inv = __shfl_sync(kFull, rank, j) == lane ? j : inv;
    rp_b = __shfl_sync(kFull, rp_b, inv);
    rp_d = __shfl_sync(kFull, rp_d, inv);
    rp_o = __shfl_sync(kFull, rp_o, inv);
  }
}

// rows of this lane group in quad q: CSR range [beg, beg+deg), row offset off, mx = max degree over the warp's groups
template <int LPR>
__device__ __forceinline__ void quad_info(int rp_b, int rp_d, int rp_o, int q, int g, int& beg, int& deg, int& off, int& mx) {
  constexpr int G = 32 / LPR;
  const int idx = q * G + g;
  beg = __shfl_sync(kFull, rp_b, idx);
  deg = __shfl_sync(kFull, rp_d, idx);
  off = __shfl_sync(kFull, rp_o, idx);
  mx = (G == 1) ? deg : __reduce_max_sync(kFull, deg);
}

// ------------------------------------------------------------------------------------------------
// Gather pipeline: a warp-private ring of D steps x SLOTS x (32 lanes x 16 B).
// The kernel supplies  gen(step descriptor)  through the Src policy:
//   header steps (NH per quad) then one step per edge position k < max degree of the quad.
// ------------------------------------------------------------------------------------------------
template <int LPR, int D, int SLOTS, int NH>
struct Pipe {
  static constexpr int G = 32 / LPR;
  uint32_t ring;        // shared-space byte address of this lane's column of the ring
  int pslot = 0, cslot = 0;
  // producer cursor
  int pq = 0, pk = 0, p_beg = 0, p_deg = 0, p_off = 0, p_n = 0;
  int nq = 0;
  int rp_b = 0, rp_d = 0, rp_o = 0;
  int g = 0;

  __device__ __forceinline__ void start(int nq_, int rp_b_, int rp_d_, int rp_o_, int g_) {
    nq = nq_; rp_b = rp_b_; rp_d = rp_d_; rp_o = rp_o_; g = g_;
    pq = 0; pk = 0;
    pslot = 0; cslot = 0;   // a previous chunk leaves only empty groups behind: restart the ring in phase
    if (nq > 0) {
      int mx;
      quad_info<LPR>(rp_b, rp_d, rp_o, 0, g, p_beg, p_deg, p_off, mx);
      p_n = mx + NH;
    }
  }
  // advance the producer cursor by one step (after the caller described the current one)
  __device__ __forceinline__ void advance() {
    if (++pk == p_n) {
      pk = 0;
      ++pq;
      if (pq < nq) {
        int mx;
        quad_info<LPR>(rp_b, rp_d, rp_o, pq, g, p_beg, p_deg, p_off, mx);
        p_n = mx + NH;
      }
    }
  }
  __device__ __forceinline__ uint32_t issue_addr(int slot_in_step) const {
    return ring + static_cast<uint32_t>((pslot * SLOTS + slot_in_step) * 512);
  }
  __device__ __forceinline__ void committed() {
    cp_commit();
    pslot = (pslot + 1 == D) ? 0 : pslot + 1;
  }
  __device__ __forceinline__ uint32_t read_addr(int slot_in_step) const {
    return ring + static_cast<uint32_t>((cslot * SLOTS + slot_in_step) * 512);
  }
  __device__ __forceinline__ void consumed() { cslot = (cslot + 1 == D) ? 0 : cslot + 1; }
};

// Per-edge record the dst pass leaves for the src pass (dst-CSR order), RSB bytes:
//   [delta_0..delta_{H-1}, alpha'_0..alpha'_{H-1}, pad to SH floats][one u16 of lrelu'(z) bits per lane of the row group]
// Lane s of the group owns bit 4t+c for element c of its float4 t (the same lane layout in both passes), so the src pass
// recovers lrelu'(z_e) without re-gathering x_r[i] and x_l[j].
__host__ __device__ constexpr int rec_scalars(int H) { return (2 * H + 3) / 4 * 4; }
__host__ __device__ constexpr int rec_bytes(int H, int LPR) { return (rec_scalars(H) * 4 + 2 * LPR + 15) / 16 * 16; }

__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
  return v;
}

__device__ __forceinline__ float4 lds4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts4(uint32_t saddr, const float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ================================================================================================
// Forward
// ================================================================================================
template <int V, int LPR, int H, int D>
__global__ void __launch_bounds__(kQThreads) gatv2_fwd_quad_kernel(const GatParams p, const int rpw, const int64_t nchunks, const int sort) {
  const uint64_t seed_eff = p.training ? gat_seed(p) : 0;
  constexpr int G = 32 / LPR, VPH = V / H;
  static_assert(V % H == 0, "a lane must hold whole heads");
  extern __shared__ float4 q_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = lane % LPR, g = lane / LPR;
  using P = Pipe<LPR, D, V, 1>;
  P pipe;
  pipe.ring = static_cast<uint32_t>(__cvta_generic_to_shared(q_smem + warp * (D * V * 32) + lane));
#pragma unroll
  for (int i = 0; i < D * V; ++i) sts4(pipe.ring + i * 512, make_float4(0.f, 0.f, 0.f, 0.f));

  float4 a[V];
#pragma unroll
  for (int t = 0; t < V; ++t) a[t] = ldg4(p.att + (t * LPR + s) * 4);
  const float slope = p.slope;
  const bool training = p.training != 0;

  for (int64_t chunk = static_cast<int64_t>(blockIdx.x) * kQW + warp; chunk < nchunks;
       chunk += static_cast<int64_t>(gridDim.x) * kQW) {
    const int64_t wrow0 = chunk * rpw;
    const int nrows = static_cast<int>(min(static_cast<int64_t>(rpw), p.n_dst - wrow0));
    const int nq = (nrows + G - 1) / G;
    int rp_b, rp_d, rp_o;
    chunk_rows<LPR>(p.rowptr, wrow0, nrows, p.n_dst, lane, sort != 0, rp_b, rp_d, rp_o);
    pipe.start(nq, rp_b, rp_d, rp_o, g);

    const float* nsrc = nullptr;
    bool nact = false;
    auto gen = [&]() {
      nact = false;
      if (pipe.pq < pipe.nq) {
        if (pipe.pk == 0) {
          const int64_t row = wrow0 + pipe.p_off;
          nact = row < p.n_dst;
          nsrc = p.x_r + row * p.ld_r;
        } else {
          const int k = pipe.pk - 1;
          if (k < pipe.p_deg) {
            nact = true;
            nsrc = p.x_l + static_cast<int64_t>(__ldg(p.col + pipe.p_beg + k)) * p.ld_l;
          }
        }
        pipe.advance();
      }
    };
    auto issue = [&]() {
      if (nact) {
#pragma unroll
        for (int t = 0; t < V; ++t) cp_async16(pipe.issue_addr(t), nsrc + (t * LPR + s) * 4);
      }
      pipe.committed();
    };
    gen();
#pragma unroll
    for (int i = 0; i < D - 1; ++i) { issue(); gen(); }

    for (int q = 0; q < nq; ++q) {
      int c_beg, c_deg, c_off, c_mx;
      quad_info<LPR>(rp_b, rp_d, rp_o, q, g, c_beg, c_deg, c_off, c_mx);
      const int64_t row = wrow0 + c_off;
      const bool rvalid = row < p.n_dst;

      issue(); gen(); cp_wait<D - 1>();
      float4 r[V], acc[V];
#pragma unroll
      for (int t = 0; t < V; ++t) {
        r[t] = lds4(pipe.read_addr(t));
        acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      pipe.consumed();
      float m[H], ss[H];
#pragma unroll
      for (int h = 0; h < H; ++h) { m[h] = 0.f; ss[h] = 0.f; }
      int eid_next = (training && c_deg > 0) ? __ldg(p.eid + c_beg) : 0;

      for (int k = 0; k < c_mx; ++k) {
        issue(); gen(); cp_wait<D - 1>();
        const bool act = k < c_deg;
        float4 x[V];
#pragma unroll
        for (int t = 0; t < V; ++t) x[t] = lds4(pipe.read_addr(t));
        pipe.consumed();
        float lg[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float2 p1 = make_float2(0.f, 0.f);
#pragma unroll
          for (int u = 0; u < VPH; ++u) {
            const int t = h * VPH + u;
            dot4p(p1, a[t], lrelu_sum4p(x[t], r[t], slope));
          }
          lg[h] = group_sum<LPR>(p1.x + p1.y);
        }
        const int e = eid_next;
        if (training && k + 1 < c_deg) eid_next = __ldg(p.eid + c_beg + k + 1);
        if (p.e_logit != nullptr && act && s < H) {   // raw logits in dst-CSR order: the backward skips att . lrelu(z)
          float lv = lg[0];
#pragma unroll
          for (int h = 1; h < H; ++h) lv = s == h ? lg[h] : lv;
          p.e_logit[static_cast<int64_t>(c_beg + k) * H + s] = lv;
        }

        if (k == 0) {
#pragma unroll
          for (int h = 0; h < H; ++h) m[h] = act ? lg[h] : 0.f;
        } else {
          bool need = false;
#pragma unroll
          for (int h = 0; h < H; ++h) need |= act && (lg[h] > m[h] + kTau);
          if (__any_sync(kFull, need)) {
#pragma unroll
            for (int h = 0; h < H; ++h) {
              const bool nh = act && (lg[h] > m[h] + kTau);
              const float mn = nh ? lg[h] : m[h];
              const float sc = __expf(m[h] - mn);
              ss[h] *= sc;
#pragma unroll
              for (int u = 0; u < VPH; ++u) scale4p(acc[h * VPH + u], sc);
              m[h] = mn;
            }
          }
        }
        uint32_t eh = 0;
        if (training) eh = rng_edge(seed_eff, static_cast<uint32_t>(e));
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float pe = act ? __expf(lg[h] - m[h]) : 0.f;
          ss[h] += pe;
          float w = pe;
          if (training) w = rng_head(eh, seed_eff, h) >= p.drop_thr ? pe * p.keep_scale : 0.f;
#pragma unroll
          for (int u = 0; u < VPH; ++u) fma4p(acc[h * VPH + u], w, x[h * VPH + u]);
        }
      }

      float inv[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float den = ss[h] + 1e-16f;
        inv[h] = 1.0f / den;
        if (s == 0 && rvalid) {
          p.stat_max[row * H + h] = m[h];
          p.stat_den[row * H + h] = den;
        }
      }
      if (rvalid) {
#pragma unroll
        for (int t = 0; t < V; ++t) {
          const int off = (t * LPR + s) * 4;
          float4 o = acc[t];
          scale4(o, inv[t / VPH]);
          if (p.bias) o = add4(o, ldg4(p.bias + off));
          if (p.out) st4(p.out + row * p.ld_out + off, o);
          if (p.out_act)
            st4(p.out_act + row * p.ld_act + off,
                make_float4(gelu_erf(o.x), gelu_erf(o.y), gelu_erf(o.z), gelu_erf(o.w)));
        }
      }
    }
  }
}

// ================================================================================================
// Backward, dst-CSR pass: grad_x_r, per-edge records (delta, alpha'), partial grad_att / grad_bias
// ================================================================================================
// MINB = resident CTAs per SM the register allocation is sized for (the persistent grid is MINB x #SMs, so
// that every CTA is resident from the start: a grid of 4 x #SMs with only 3 CTAs fitting runs a second,
// two-thirds-empty wave).
template <int V, int LPR, int H, int D, int MINB>
__global__ void __launch_bounds__(kQThreads, MINB) gatv2_bwd_dst_quad_kernel(const GatParams p, const int rpw, const int64_t nchunks, const int sort) {
  const uint64_t seed_eff = p.training ? gat_seed(p) : 0;
  constexpr int G = 32 / LPR, VPH = V / H, F4 = V * LPR;
  constexpr int SH = rec_scalars(H), RSB = rec_bytes(H, LPR);
  extern __shared__ float4 q_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = lane % LPR, g = lane / LPR;
  using P = Pipe<LPR, D, V, 3>;
  P pipe;
  // per warp: ring D*V*512 B, then V*512 B of grad_bias accumulators (lane-private)
  float4* wbase = q_smem + warp * ((D + 1) * V * 32);
  pipe.ring = static_cast<uint32_t>(__cvta_generic_to_shared(wbase + lane));
  const uint32_t gb_addr = pipe.ring + D * V * 512;
#pragma unroll
  for (int i = 0; i < (D + 1) * V; ++i) sts4(pipe.ring + i * 512, make_float4(0.f, 0.f, 0.f, 0.f));

  float4 a[V], gatt[V];
#pragma unroll
  for (int t = 0; t < V; ++t) {
    a[t] = ldg4(p.att + (t * LPR + s) * 4);
    gatt[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float slope = p.slope;
  const bool training = p.training != 0;
  const bool direct_src = p.t_rowptr == nullptr;   // one source per edge: this pass also writes grad_x_l

  for (int64_t chunk = static_cast<int64_t>(blockIdx.x) * kQW + warp; chunk < nchunks;
       chunk += static_cast<int64_t>(gridDim.x) * kQW) {
    const int64_t wrow0 = chunk * rpw;
    const int nrows = static_cast<int>(min(static_cast<int64_t>(rpw), p.n_dst - wrow0));
    const int nq = (nrows + G - 1) / G;
    int rp_b, rp_d, rp_o;
    chunk_rows<LPR>(p.rowptr, wrow0, nrows, p.n_dst, lane, sort != 0, rp_b, rp_d, rp_o);
    pipe.start(nq, rp_b, rp_d, rp_o, g);

    const float* nsrc = nullptr;
    bool nact = false;
    auto gen = [&]() {
      nact = false;
      if (pipe.pq < pipe.nq) {
        if (pipe.pk < 3) {
          const int64_t row = wrow0 + pipe.p_off;
          nact = row < p.n_dst;
          nsrc = pipe.pk == 0 ? p.x_r + row * p.ld_r : (pipe.pk == 1 ? p.grad_out + row * p.ld_g : p.out + row * p.ld_out);
        } else {
          const int k = pipe.pk - 3;
          if (k < pipe.p_deg) {
            nact = true;
            nsrc = p.x_l + static_cast<int64_t>(__ldg(p.col + pipe.p_beg + k)) * p.ld_l;
          }
        }
        pipe.advance();
      }
    };
    auto issue = [&]() {
      if (nact) {
#pragma unroll
        for (int t = 0; t < V; ++t) cp_async16(pipe.issue_addr(t), nsrc + (t * LPR + s) * 4);
      }
      pipe.committed();
    };
    gen();
#pragma unroll
    for (int i = 0; i < D - 1; ++i) { issue(); gen(); }

    for (int q = 0; q < nq; ++q) {
      int c_beg, c_deg, c_off, c_mx;
      quad_info<LPR>(rp_b, rp_d, rp_o, q, g, c_beg, c_deg, c_off, c_mx);
      const int64_t row = wrow0 + c_off;
      const bool rvalid = row < p.n_dst;
      float m[H], inv[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        m[h] = rvalid ? __ldg(p.stat_max + row * H + h) : 0.f;
        inv[h] = rvalid ? __ldg(p.stat_den + row * H + h) : 1.f;
      }
      int eid_next = (training && c_deg > 0) ? __ldg(p.eid + c_beg) : 0;

      float4 r[V], g4[V], gr[V];
      issue(); gen(); cp_wait<D - 1>();
#pragma unroll
      for (int t = 0; t < V; ++t) r[t] = lds4(pipe.read_addr(t));
      pipe.consumed();
      issue(); gen(); cp_wait<D - 1>();
#pragma unroll
      for (int t = 0; t < V; ++t) g4[t] = lds4(pipe.read_addr(t));
      pipe.consumed();
      issue(); gen(); cp_wait<D - 1>();
      float cdot[H];
#pragma unroll
      for (int h = 0; h < H; ++h) cdot[h] = 0.f;
#pragma unroll
      for (int t = 0; t < V; ++t) {
        const int off = (t * LPR + s) * 4;
        const float4 o4 = lds4(pipe.read_addr(t));
        if (!rvalid) g4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.gelu_fused) {
          g4[t].x *= gelu_erf_grad(o4.x); g4[t].y *= gelu_erf_grad(o4.y);
          g4[t].z *= gelu_erf_grad(o4.z); g4[t].w *= gelu_erf_grad(o4.w);
          if (rvalid) st4(p.g_buf + row * p.ld_g + off, g4[t]);
        }
        const float4 b4 = p.bias ? ldg4(p.bias + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        cdot[t / VPH] += dot4(g4[t], sub4(o4, b4));
        const uint32_t ga = gb_addr + t * 512;
        sts4(ga, add4(lds4(ga), g4[t]));
        gr[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      pipe.consumed();
#pragma unroll
      for (int h = 0; h < H; ++h) {
        cdot[h] = group_sum<LPR>(cdot[h]);
        inv[h] = 1.0f / inv[h];
      }

      for (int k = 0; k < c_mx; ++k) {
        issue(); gen(); cp_wait<D - 1>();
        const bool act = k < c_deg;
        float4 z[V];   // holds u = lrelu(z): sign(u) == sign(z), so lrelu'(z) is recovered from u
        float lg[H], dd[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float2 p1 = make_float2(0.f, 0.f), pd = make_float2(0.f, 0.f);
#pragma unroll
          for (int u = 0; u < VPH; ++u) {
            const int t = h * VPH + u;
            const float4 x = lds4(pipe.read_addr(t));
            dot4p(pd, g4[t], x);
            z[t] = lrelu_sum4p(x, r[t], slope);
            dot4p(p1, a[t], z[t]);
          }
          lg[h] = group_sum<LPR>(p1.x + p1.y);
          dd[h] = group_sum<LPR>(pd.x + pd.y);
        }
        pipe.consumed();
        const int e = eid_next;
        if (training && k + 1 < c_deg) eid_next = __ldg(p.eid + c_beg + k + 1);
        const int jcol = (direct_src && act) ? __ldg(p.col + c_beg + k) : 0;
        uint32_t eh = 0;
        if (training) eh = rng_edge(seed_eff, static_cast<uint32_t>(e));
        float delta[H], alk[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float alpha = act ? __expf(lg[h] - m[h]) * inv[h] : 0.f;
          float ks = 1.0f;
          if (training) ks = rng_head(eh, seed_eff, h) >= p.drop_thr ? p.keep_scale : 0.f;
          delta[h] = alpha * (dd[h] * ks - cdot[h]);
          alk[h] = alpha * ks;
        }
        char* rec_ptr = reinterpret_cast<char*>(p.e_delta) + static_cast<int64_t>(c_beg + k) * RSB;
        if (act && !direct_src && s < SH / 4) {
          // record: [delta_0..delta_{H-1}, alpha'_0..alpha'_{H-1}, pad]
          float rec[SH];
#pragma unroll
          for (int i = 0; i < SH; ++i) rec[i] = i < H ? delta[i] : (i < 2 * H ? alk[i - H] : 0.f);
          float4 out4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < SH / 4; ++i)
            if (s == i) out4 = make_float4(rec[4 * i], rec[4 * i + 1], rec[4 * i + 2], rec[4 * i + 3]);
          st4(reinterpret_cast<float*>(rec_ptr) + s * 4, out4);
        }
        uint32_t zbits = 0;
#pragma unroll
        for (int t = 0; t < V; ++t) {
          const float d = delta[t / VPH], ds = d * slope;     // delta * lrelu'(z): d where z > 0, d * slope elsewhere
          const float4 zz = z[t];
          if (!direct_src) {           // (warp-uniform) the common form: accumulate dL/dz_e into grad_x_r with FMAs
            const bool px = zz.x > 0.f, py = zz.y > 0.f, pz = zz.z > 0.f, pw = zz.w > 0.f;
            fma4v(gr[t], make_float4(px ? d : ds, py ? d : ds, pz ? d : ds, pw ? d : ds), a[t]);
            zbits |= (px ? 1u : 0u) << (4 * t) | (py ? 2u : 0u) << (4 * t) | (pz ? 4u : 0u) << (4 * t) | (pw ? 8u : 0u) << (4 * t);
          } else {                     // one source per edge: grad_x_l[j] = dL/dz_e + alpha'_e g_i is written here too
            const float4 dz = make_float4((zz.x > 0.f ? d : ds) * a[t].x, (zz.y > 0.f ? d : ds) * a[t].y,
                                          (zz.z > 0.f ? d : ds) * a[t].z, (zz.w > 0.f ? d : ds) * a[t].w);
            gr[t] = add4(gr[t], dz);
            if (act) {
              const float ak = alk[t / VPH];
              st4(p.grad_x_l + static_cast<int64_t>(jcol) * p.ld_gl + (t * LPR + s) * 4,
                  make_float4(fmaf(ak, g4[t].x, dz.x), fmaf(ak, g4[t].y, dz.y), fmaf(ak, g4[t].z, dz.z), fmaf(ak, g4[t].w, dz.w)));
            }
          }
          fma4p(gatt[t], d, zz);
        }
        if (act && !direct_src) *reinterpret_cast<uint16_t*>(rec_ptr + SH * 4 + 2 * s) = static_cast<uint16_t>(zbits);
      }
      if (rvalid) {
#pragma unroll
        for (int t = 0; t < V; ++t) st4(p.grad_x_r + row * p.ld_gr + (t * LPR + s) * 4, gr[t]);
      }
    }
  }

  // ---- CTA-level ordered reduction -> partial[blockIdx][2][F] ------------------------------------
  float4 gb[V];
#pragma unroll
  for (int t = 0; t < V; ++t) {
    gb[t] = lds4(gb_addr + t * 512);
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {   // sum over the G lane groups (same columns, different rows)
      gatt[t].x += __shfl_xor_sync(kFull, gatt[t].x, o); gatt[t].y += __shfl_xor_sync(kFull, gatt[t].y, o);
      gatt[t].z += __shfl_xor_sync(kFull, gatt[t].z, o); gatt[t].w += __shfl_xor_sync(kFull, gatt[t].w, o);
      gb[t].x += __shfl_xor_sync(kFull, gb[t].x, o); gb[t].y += __shfl_xor_sync(kFull, gb[t].y, o);
      gb[t].z += __shfl_xor_sync(kFull, gb[t].z, o); gb[t].w += __shfl_xor_sync(kFull, gb[t].w, o);
    }
  }
  __syncthreads();   // every warp is done with its ring; reuse the start of shared memory
  float4* red = q_smem;   // [2][kQW][F4]
  if (g == 0) {
#pragma unroll
    for (int t = 0; t < V; ++t) {
      red[(0 * kQW + warp) * F4 + t * LPR + s] = gatt[t];
      red[(1 * kQW + warp) * F4 + t * LPR + s] = gb[t];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * F4; i += kQThreads) {
    const int which = i / F4, f = i % F4;
    float4 tsum = red[(which * kQW + 0) * F4 + f];
#pragma unroll
    for (int w = 1; w < kQW; ++w) tsum = add4(tsum, red[(which * kQW + w) * F4 + f]);
    st4(p.partial + (static_cast<int64_t>(blockIdx.x) * 2 + which) * (F4 * 4) + f * 4, tsum);
  }
}

// ------------------------------------------------------------------------------------------------
// dst pass, saved-logit form (p.e_logit != NULL: the forward left the raw logit of every (edge, head) in dst-CSR order).
// Same results as the kernel above up to rounding; per edge and element it no longer evaluates
// lrelu(z) and att . lrelu(z) (the logit is read back, 4H bytes per edge) and it keeps neither `att` nor a copy of z in
// registers:
//   dL/dz_e = delta_e * lrelu'(z_e) (.) att            sel_e := delta_e * lrelu'(z_e)   (delta or delta * slope per element)
//   grad_x_r[i] = att (.) sum_e sel_e                   (att applied once per row, from shared memory)
//   grad_att   += sel_e (.) z_e                         (lrelu(z) = lrelu'(z) z)
// Two sweeps over the staged source row per edge (g . x first: delta needs it), both from the ring slot.
// ------------------------------------------------------------------------------------------------
// LEAN: x_r[i] and the grad_att accumulators live in lane-private shared memory as well and att is read through L1 at the
// row end, which frees ~32 registers per thread for a fourth resident CTA per SM (the pass is bound by dependent-issue
// latency at 3 warps per scheduler, not by instruction count): 8 more shared-memory accesses per 4-edge step.
template <int V, int LPR, int H, int D, int MINB, bool LEAN>
__global__ void __launch_bounds__(kQThreads, MINB) gatv2_bwd_dst_lg_quad_kernel(const GatParams p, const int rpw, const int64_t nchunks, const int sort) {
  const uint64_t seed_eff = p.training ? gat_seed(p) : 0;
  constexpr int G = 32 / LPR, VPH = V / H, F4 = V * LPR;
  constexpr int SH = rec_scalars(H), RSB = rec_bytes(H, LPR);
  extern __shared__ float4 q_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = lane % LPR, g = lane / LPR;
  using P = Pipe<LPR, D, V, 3>;
  P pipe;
  // per warp (lane-private columns): ring D*V*512 B, V*512 B of grad_bias accumulators, then
  //   !LEAN: V*512 B holding this lane's slice of att;   LEAN: V*512 B for x_r[i] and V*512 B of grad_att accumulators
  constexpr int kExtra = LEAN ? 3 : 2;
  float4* wbase = q_smem + warp * ((D + kExtra) * V * 32);
  pipe.ring = static_cast<uint32_t>(__cvta_generic_to_shared(wbase + lane));
  const uint32_t gb_addr = pipe.ring + D * V * 512;
  const uint32_t att_addr = gb_addr + V * 512;        // !LEAN
  const uint32_t r_addr = gb_addr + V * 512;          // LEAN
  const uint32_t gatt_addr = r_addr + V * 512;        // LEAN
#pragma unroll
  for (int i = 0; i < (D + kExtra) * V; ++i) sts4(pipe.ring + i * 512, make_float4(0.f, 0.f, 0.f, 0.f));
  float4 gatt[LEAN ? 1 : V];
  if constexpr (!LEAN) {
#pragma unroll
    for (int t = 0; t < V; ++t) {
      sts4(att_addr + t * 512, ldg4(p.att + (t * LPR + s) * 4));
      gatt[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  auto att_of = [&](int t) { return LEAN ? ldg4(p.att + (t * LPR + s) * 4) : lds4(att_addr + t * 512); };
  const float slope = p.slope;
  const bool training = p.training != 0;
  const bool direct_src = p.t_rowptr == nullptr;   // one source per edge: this pass also writes grad_x_l

  for (int64_t chunk = static_cast<int64_t>(blockIdx.x) * kQW + warp; chunk < nchunks;
       chunk += static_cast<int64_t>(gridDim.x) * kQW) {
    const int64_t wrow0 = chunk * rpw;
    const int nrows = static_cast<int>(min(static_cast<int64_t>(rpw), p.n_dst - wrow0));
    const int nq = (nrows + G - 1) / G;
    int rp_b, rp_d, rp_o;
    chunk_rows<LPR>(p.rowptr, wrow0, nrows, p.n_dst, lane, sort != 0, rp_b, rp_d, rp_o);
    pipe.start(nq, rp_b, rp_d, rp_o, g);

    const float* nsrc = nullptr;
    bool nact = false;
    auto gen = [&]() {
      nact = false;
      if (pipe.pq < pipe.nq) {
        if (pipe.pk < 3) {
          const int64_t row = wrow0 + pipe.p_off;
          nact = row < p.n_dst;
          nsrc = pipe.pk == 0 ? p.x_r + row * p.ld_r : (pipe.pk == 1 ? p.grad_out + row * p.ld_g : p.out + row * p.ld_out);
        } else {
          const int k = pipe.pk - 3;
          if (k < pipe.p_deg) {
            nact = true;
            nsrc = p.x_l + static_cast<int64_t>(__ldg(p.col + pipe.p_beg + k)) * p.ld_l;
          }
        }
        pipe.advance();
      }
    };
    auto issue = [&]() {
      if (nact) {
#pragma unroll
        for (int t = 0; t < V; ++t) cp_async16(pipe.issue_addr(t), nsrc + (t * LPR + s) * 4);
      }
      pipe.committed();
    };
    gen();
#pragma unroll
    for (int i = 0; i < D - 1; ++i) { issue(); gen(); }

    for (int q = 0; q < nq; ++q) {
      int c_beg, c_deg, c_off, c_mx;
      quad_info<LPR>(rp_b, rp_d, rp_o, q, g, c_beg, c_deg, c_off, c_mx);
      const int64_t row = wrow0 + c_off;
      const bool rvalid = row < p.n_dst;
      float m[H], inv[H], lgn[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        m[h] = rvalid ? __ldg(p.stat_max + row * H + h) : 0.f;
        inv[h] = rvalid ? __ldg(p.stat_den + row * H + h) : 1.f;
        lgn[h] = c_deg > 0 ? __ldg(p.e_logit + static_cast<int64_t>(c_beg) * H + h) : 0.f;
      }
      int eid_next = (training && c_deg > 0) ? __ldg(p.eid + c_beg) : 0;

      float4 r[LEAN ? 1 : V], g4[V], sacc[V];
      issue(); gen(); cp_wait<D - 1>();
#pragma unroll
      for (int t = 0; t < V; ++t) {
        if constexpr (LEAN) sts4(r_addr + t * 512, lds4(pipe.read_addr(t)));
        else r[t] = lds4(pipe.read_addr(t));
      }
      pipe.consumed();
      issue(); gen(); cp_wait<D - 1>();
#pragma unroll
      for (int t = 0; t < V; ++t) g4[t] = lds4(pipe.read_addr(t));
      pipe.consumed();
      issue(); gen(); cp_wait<D - 1>();
      float cdot[H];
#pragma unroll
      for (int h = 0; h < H; ++h) cdot[h] = 0.f;
#pragma unroll
      for (int t = 0; t < V; ++t) {
        const int off = (t * LPR + s) * 4;
        const float4 o4 = lds4(pipe.read_addr(t));
        if (!rvalid) g4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.gelu_fused) {
          g4[t].x *= gelu_erf_grad(o4.x); g4[t].y *= gelu_erf_grad(o4.y);
          g4[t].z *= gelu_erf_grad(o4.z); g4[t].w *= gelu_erf_grad(o4.w);
          if (rvalid) st4(p.g_buf + row * p.ld_g + off, g4[t]);
        }
        const float4 b4 = p.bias ? ldg4(p.bias + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        cdot[t / VPH] += dot4(g4[t], sub4(o4, b4));
        const uint32_t ga = gb_addr + t * 512;
        sts4(ga, add4(lds4(ga), g4[t]));
        sacc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      pipe.consumed();
#pragma unroll
      for (int h = 0; h < H; ++h) {
        cdot[h] = group_sum<LPR>(cdot[h]);
        inv[h] = 1.0f / inv[h];
      }

      for (int k = 0; k < c_mx; ++k) {
        issue(); gen(); cp_wait<D - 1>();
        const bool act = k < c_deg;
        float dd[H], lg[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {           // sweep 1: g_i . x_l[j] per head
          float2 pd = make_float2(0.f, 0.f);
#pragma unroll
          for (int u = 0; u < VPH; ++u) {
            const int t = h * VPH + u;
            dot4p(pd, g4[t], lds4(pipe.read_addr(t)));
          }
          dd[h] = group_sum<LPR>(pd.x + pd.y);
          lg[h] = lgn[h];
        }
        if (k + 1 < c_deg) {
#pragma unroll
          for (int h = 0; h < H; ++h) lgn[h] = __ldg(p.e_logit + static_cast<int64_t>(c_beg + k + 1) * H + h);
        }
        const int e = eid_next;
        if (training && k + 1 < c_deg) eid_next = __ldg(p.eid + c_beg + k + 1);
        const int jcol = (direct_src && act) ? __ldg(p.col + c_beg + k) : 0;
        uint32_t eh = 0;
        if (training) eh = rng_edge(seed_eff, static_cast<uint32_t>(e));
        float delta[H], alk[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float alpha = act ? __expf(lg[h] - m[h]) * inv[h] : 0.f;
          float ks = 1.0f;
          if (training) ks = rng_head(eh, seed_eff, h) >= p.drop_thr ? p.keep_scale : 0.f;
          delta[h] = alpha * (dd[h] * ks - cdot[h]);
          alk[h] = alpha * ks;
        }
        char* rec_ptr = reinterpret_cast<char*>(p.e_delta) + static_cast<int64_t>(c_beg + k) * RSB;
        if (act && !direct_src && s < SH / 4) {
          float rec[SH];
#pragma unroll
          for (int i = 0; i < SH; ++i) rec[i] = i < H ? delta[i] : (i < 2 * H ? alk[i - H] : 0.f);
          float4 out4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < SH / 4; ++i)
            if (s == i) out4 = make_float4(rec[4 * i], rec[4 * i + 1], rec[4 * i + 2], rec[4 * i + 3]);
          st4(reinterpret_cast<float*>(rec_ptr) + s * 4, out4);
        }
        uint32_t zbits = 0;
#pragma unroll
        for (int t = 0; t < V; ++t) {           // sweep 2: z = x_l[j] + x_r[i], sel = delta * lrelu'(z)
          const float d = delta[t / VPH], ds = d * slope;
          float4 rt;
          if constexpr (LEAN) rt = lds4(r_addr + t * 512); else rt = r[t];
          const float4 zz = add4p(lds4(pipe.read_addr(t)), rt);
          const bool px = zz.x > 0.f, py = zz.y > 0.f, pz = zz.z > 0.f, pw = zz.w > 0.f;
          const float4 sel = make_float4(px ? d : ds, py ? d : ds, pz ? d : ds, pw ? d : ds);
          sacc[t] = add4p(sacc[t], sel);
          if constexpr (LEAN) {
            float4 ga = lds4(gatt_addr + t * 512);
            fma4v(ga, sel, zz);
            sts4(gatt_addr + t * 512, ga);
          } else {
            fma4v(gatt[t], sel, zz);
          }
          if (!direct_src) {
            zbits |= (px ? 1u : 0u) << (4 * t) | (py ? 2u : 0u) << (4 * t) | (pz ? 4u : 0u) << (4 * t) | (pw ? 8u : 0u) << (4 * t);
          } else if (act) {                     // one source per edge: grad_x_l[j] = dL/dz_e + alpha'_e g_i
            const float4 dz = mul4p(sel, att_of(t));
            const float ak = alk[t / VPH];
            st4(p.grad_x_l + static_cast<int64_t>(jcol) * p.ld_gl + (t * LPR + s) * 4,
                make_float4(fmaf(ak, g4[t].x, dz.x), fmaf(ak, g4[t].y, dz.y), fmaf(ak, g4[t].z, dz.z), fmaf(ak, g4[t].w, dz.w)));
          }
        }
        pipe.consumed();
        if (act && !direct_src) *reinterpret_cast<uint16_t*>(rec_ptr + SH * 4 + 2 * s) = static_cast<uint16_t>(zbits);
      }
      if (rvalid) {
#pragma unroll
        for (int t = 0; t < V; ++t)
          st4(p.grad_x_r + row * p.ld_gr + (t * LPR + s) * 4, mul4p(att_of(t), sacc[t]));
      }
    }
  }

  // ---- CTA-level ordered reduction -> partial[blockIdx][2][F] ------------------------------------
  float4 gb[V], gsum[V];
#pragma unroll
  for (int t = 0; t < V; ++t) {
    if constexpr (LEAN) gsum[t] = lds4(gatt_addr + t * 512); else gsum[t] = gatt[t];
  }
#define gatt gsum
#pragma unroll
  for (int t = 0; t < V; ++t) {
    gb[t] = lds4(gb_addr + t * 512);
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {   // sum over the G lane groups (same columns, different rows)
      gatt[t].x += __shfl_xor_sync(kFull, gatt[t].x, o); gatt[t].y += __shfl_xor_sync(kFull, gatt[t].y, o);
      gatt[t].z += __shfl_xor_sync(kFull, gatt[t].z, o); gatt[t].w += __shfl_xor_sync(kFull, gatt[t].w, o);
      gb[t].x += __shfl_xor_sync(kFull, gb[t].x, o); gb[t].y += __shfl_xor_sync(kFull, gb[t].y, o);
      gb[t].z += __shfl_xor_sync(kFull, gb[t].z, o); gb[t].w += __shfl_xor_sync(kFull, gb[t].w, o);
    }
  }
  __syncthreads();   // every warp is done with its ring; reuse the start of shared memory
  float4* red = q_smem;   // [2][kQW][F4]
  if (g == 0) {
#pragma unroll
    for (int t = 0; t < V; ++t) {
      red[(0 * kQW + warp) * F4 + t * LPR + s] = gatt[t];
      red[(1 * kQW + warp) * F4 + t * LPR + s] = gb[t];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * F4; i += kQThreads) {
    const int which = i / F4, f = i % F4;
    float4 tsum = red[(which * kQW + 0) * F4 + f];
#pragma unroll
    for (int w = 1; w < kQW; ++w) tsum = add4(tsum, red[(which * kQW + w) * F4 + f]);
    st4(p.partial + (static_cast<int64_t>(blockIdx.x) * 2 + which) * (F4 * 4) + f * 4, tsum);
  }
#undef gatt
}

// fixed-order column sums of partial[nb][cols] -> grad_att | grad_bias.  block = (32, 8)
__global__ void quad_colsum_kernel(const float* __restrict__ partial, int nb, int cols, int F,
                                   float* __restrict__ grad_att, float* __restrict__ grad_bias) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < cols)
    for (int b = threadIdx.y; b < nb; b += 8) acc += partial[static_cast<int64_t>(b) * cols + c];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = red[0][threadIdx.x];
#pragma unroll
    for (int y = 1; y < 8; ++y) t += red[y][threadIdx.x];
    if (c < F) grad_att[c] = t;
    else if (grad_bias) grad_bias[c - F] = t;
  }
}

// ================================================================================================
// Backward, src-CSR (transposed) pass: grad_x_l
// ================================================================================================
template <int V, int LPR, int H, int D>
__global__ void __launch_bounds__(kQThreads) gatv2_bwd_src_quad_kernel(const GatParams p, const int rpw, const int64_t nchunks, const int sort) {
  // grad_x_l[j] = sum over out-edges e = (j -> i) of  delta_e * att (.) lrelu'(z_e)  +  alpha'_e * g_i
  // Per edge the ring carries the V float4 of g_i this lane consumes and the edge's record (scalars + lrelu' bits):
  // neither x_r[i] nor the row's own x_l is read.
  constexpr int G = 32 / LPR, VPH = V / H;
  constexpr int SH = rec_scalars(H), RSB = rec_bytes(H, LPR), RCH = RSB / 16;
  constexpr int SLOTS = V + 1;
  static_assert(RCH <= LPR, "the record of a row group is staged by the lanes of that group");
  extern __shared__ float4 q_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = lane % LPR, g = lane / LPR;
  using P = Pipe<LPR, D, SLOTS, 1>;
  P pipe;
  pipe.ring = static_cast<uint32_t>(__cvta_generic_to_shared(q_smem + warp * (D * SLOTS * 32) + lane));
#pragma unroll
  for (int i = 0; i < D * SLOTS; ++i) sts4(pipe.ring + i * 512, make_float4(0.f, 0.f, 0.f, 0.f));
  // the record of this lane's group starts at lane g*LPR of slot V
  const uint32_t rec_off = static_cast<uint32_t>(V * 512) - static_cast<uint32_t>(s * 16);

  float4 a[V];
#pragma unroll
  for (int t = 0; t < V; ++t) a[t] = ldg4(p.att + (t * LPR + s) * 4);
  const float slope = p.slope;
  const float* gsrc = p.gelu_fused ? p.g_buf : p.grad_out;
  const char* recs = reinterpret_cast<const char*>(p.e_delta);

  for (int64_t chunk = static_cast<int64_t>(blockIdx.x) * kQW + warp; chunk < nchunks;
       chunk += static_cast<int64_t>(gridDim.x) * kQW) {
    const int64_t wrow0 = chunk * rpw;
    const int nrows = static_cast<int>(min(static_cast<int64_t>(rpw), p.n_src - wrow0));
    const int nq = (nrows + G - 1) / G;
    int rp_b, rp_d, rp_o;
    chunk_rows<LPR>(p.t_rowptr, wrow0, nrows, p.n_src, lane, sort != 0, rp_b, rp_d, rp_o);
    pipe.start(nq, rp_b, rp_d, rp_o, g);

    const float* nsrc = nullptr;
    const char* nrec = nullptr;
    bool nact = false;
    auto gen = [&]() {
      nact = false;
      if (pipe.pq < pipe.nq) {
        if (pipe.pk > 0) {                     // step 0 of a quad is an empty header (keeps the cursor arithmetic uniform)
          const int k = pipe.pk - 1;
          if (k < pipe.p_deg) {
            nact = true;
            const int64_t i = __ldg(p.t_dst + pipe.p_beg + k);
            const int64_t pos = __ldg(p.t_pos + pipe.p_beg + k);
            nsrc = gsrc + i * p.ld_g;
            nrec = recs + pos * RSB;
          }
        }
        pipe.advance();
      }
    };
    auto issue = [&]() {
      if (nact) {
#pragma unroll
        for (int t = 0; t < V; ++t) cp_async16(pipe.issue_addr(t), nsrc + (t * LPR + s) * 4);
        if (s < RCH) cp_async16(pipe.issue_addr(V), nrec + s * 16);
      }
      pipe.committed();
    };
    gen();
#pragma unroll
    for (int i = 0; i < D - 1; ++i) { issue(); gen(); }

    for (int q = 0; q < nq; ++q) {
      int c_beg, c_deg, c_off, c_mx;
      quad_info<LPR>(rp_b, rp_d, rp_o, q, g, c_beg, c_deg, c_off, c_mx);
      const int64_t row = wrow0 + c_off;
      const bool rvalid = row < p.n_src;

      issue(); gen(); cp_wait<D - 1>();        // header step: nothing to read
      pipe.consumed();
      float4 acc[V];
#pragma unroll
      for (int t = 0; t < V; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);

      for (int k = 0; k < c_mx; ++k) {
        issue(); gen(); cp_wait<D - 1>();
        __syncwarp();   // the record was staged by other lanes of the group
        const bool act = k < c_deg;
        float rec[SH];
#pragma unroll
        for (int i = 0; i < SH / 4; ++i) {
          const float4 v = lds4(pipe.read_addr(0) + rec_off + i * 16);
          rec[4 * i] = v.x; rec[4 * i + 1] = v.y; rec[4 * i + 2] = v.z; rec[4 * i + 3] = v.w;
        }
        const uint32_t zb = lds_u16(pipe.read_addr(0) + rec_off + SH * 4 + 2 * s);
#pragma unroll
        for (int t = 0; t < V; ++t) {
          const float4 gg = lds4(pipe.read_addr(t));
          const float d = act ? rec[t / VPH] : 0.f;
          const float al = act ? rec[H + t / VPH] : 0.f;
          const float ds = d * slope;
          fma4v(acc[t], make_float4((zb >> (4 * t)) & 1u ? d : ds, (zb >> (4 * t + 1)) & 1u ? d : ds,
                                    (zb >> (4 * t + 2)) & 1u ? d : ds, (zb >> (4 * t + 3)) & 1u ? d : ds), a[t]);
          if (act) fma4p(acc[t], al, gg);
        }
        pipe.consumed();
        __syncwarp();   // all lanes are done with this step's record before its slot is refilled
      }
      if (rvalid) {
#pragma unroll
        for (int t = 0; t < V; ++t) st4(p.grad_x_l + row * p.ld_gl + (t * LPR + s) * 4, acc[t]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
#define SGB_QUAD_COMBOS(X) \
  X(4, 8, 1) X(4, 8, 2) X(4, 8, 4) X(3, 8, 3) X(3, 8, 1) X(2, 8, 1) X(2, 8, 2) X(2, 32, 1) X(2, 32, 2) X(4, 32, 1) X(4, 32, 2) X(4, 32, 4)

struct QShape {
  int v, lpr;
};

// F = 16*V*LPR/4 ... a lane holds V float4, LPR lanes per row: F = 4*V*LPR; heads must not split a
// float4 slice: C % (4*LPR) == 0.
bool quad_shape(int H, int C, QShape& qs) {
  const int F = H * C;
  const int lprs[2] = {8, 32};
  for (int i = 0; i < 2; ++i) {
    const int lpr = lprs[i];
    if (F % (4 * lpr) != 0 || C % (4 * lpr) != 0) continue;
    const int v = F / (4 * lpr);
    if (v > 4 || v % H != 0) continue;
#define X(V, L, Hh) if (v == V && lpr == L && H == Hh) { qs = {V, L}; return true; }
    SGB_QUAD_COMBOS(X)
#undef X
  }
  return false;
}

int pick_rpw(int64_t n_rows, int G) {
  // rows per warp: enough warps to fill the machine several times over, at most 32 rows (rowptr lives in lanes)
  const int64_t target_warps = static_cast<int64_t>(sm_count()) * 16 * 6;
  int64_t rpw = n_rows / target_warps;
  rpw = rpw / G * G;
  if (rpw < G) rpw = G;
  if (rpw > 32) rpw = 32;
  // a whole warp per row (F = 512: 2 KB rows): shorter chunks keep the rows in flight -- and the source rows their
  // 20 neighbours share -- inside the L2 (measured on configs[3]: 8 rows 8.50 / 17.7 ms fwd / bwd, 32 rows 8.93 / 19.0)
  if (G == 1 && rpw > 8) rpw = 8;
  static const int forced = [] { const char* e = getenv("SEGGER_B200_GAT_RPW"); return e ? atoi(e) : 0; }();
  if (forced > 0) { rpw = forced / G * G; if (rpw < G) rpw = G; if (rpw > 32) rpw = 32; }
  return static_cast<int>(rpw);
}

constexpr int kDFwd = 4, kDDst = 4, kDSrc = 4;
constexpr int kDDstLean = 3;   // 4 resident CTAs per SM: (3 + 3) * V * 512 B per warp = 48 KB per CTA at V = 4

// SEGGER_B200_GAT_SORT=0: rows keep their chunk order inside a warp (A/B of the degree-ordered quads); read per call
int quad_sort() {
  const char* e = getenv("SEGGER_B200_GAT_SORT");
  return (e && e[0] == '0') ? 0 : 1;
}

template <typename K>
bool set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)) == cudaSuccess;
}

}  // namespace

// max(z, slope*z) and the sign test on lrelu(z) in the backward need 0 <= slope <= 1
static bool quad_slope_ok(float slope) { return slope >= 0.f && slope <= 1.f; }

bool quad_supported(const GatParams& p) {
  QShape qs;
  return quad_slope_ok(p.slope) && quad_shape(p.H, p.C, qs);
}

bool quad_fwd_launch(const GatParams& p, cudaStream_t stream) {
  QShape qs;
  if (!quad_slope_ok(p.slope) || !quad_shape(p.H, p.C, qs)) return false;
  const int G = 32 / qs.lpr;
  const int rpw = pick_rpw(p.n_dst, G);
  const int64_t nchunks = ceil_div(p.n_dst, rpw);
  const unsigned blocks = static_cast<unsigned>(ceil_div(nchunks, kQW));
#define X(V, L, Hh)                                                                                   \
  if (qs.v == V && qs.lpr == L && p.H == Hh) {                                                        \
    const size_t smem = static_cast<size_t>(kQW) * kDFwd * V * 512;                                   \
    auto kern = gatv2_fwd_quad_kernel<V, L, Hh, kDFwd>;                                               \
    if (smem > 48 * 1024 && !set_smem(kern, smem)) return false;                                      \
    kern<<<blocks, kQThreads, smem, stream>>>(p, rpw, nchunks, quad_sort());                                       \
    return true;                                                                                      \
  }
  SGB_QUAD_COMBOS(X)
#undef X
  return false;
}

// resident CTAs per SM of the dst pass: 3 (no register cap) or 4 (128 registers, a few spilled scalars)
static int quad_dst_minb() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SEGGER_B200_GAT_DST_OCC");
    v = (e && e[0] == '4') ? 4 : 3;
  }
  return v;
}
// The saved-logit dst pass runs its LEAN form (4 resident CTAs per SM) on two-pass graphs: measured on the cfg-2
// tx-neighbors-tx conv 1.475 -> 1.422 ms for dst + src; the one-source-per-edge form (tx-belongs-bd) needs att on every
// step and is faster with it in shared memory (0.205 against 0.220 ms).  SEGGER_B200_GAT_DST_OCC=3 / 4 forces a form.
static bool quad_dst_lean(const GatParams& p) {
  static const int forced = [] { const char* e = getenv("SEGGER_B200_GAT_DST_OCC"); return e ? atoi(e) : 0; }();
  if (p.e_logit == nullptr) return false;
  return forced == 4 || (forced != 3 && p.t_rowptr != nullptr);
}
static int quad_dst_blocks(int64_t n_dst, int rpw, int per_sm) {
  const int64_t nchunks = ceil_div(n_dst > 0 ? n_dst : 1, rpw);
  const int64_t want = ceil_div(nchunks, kQW);
  const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
  return static_cast<int>(want < cap ? want : cap);
}

size_t quad_bwd_record_bytes(int H, int C) {
  QShape qs;
  if (!quad_shape(H, C, qs)) return 0;
  return static_cast<size_t>(rec_bytes(H, qs.lpr));
}

size_t quad_bwd_partial_floats(int H, int C) {
  QShape qs;
  if (!quad_shape(H, C, qs)) return 0;
  return static_cast<size_t>(sm_count()) * 4 * 2 * H * C;
}

bool quad_bwd_launch(const GatParams& p, float* grad_att, float* grad_bias, cudaStream_t stream) {
  QShape qs;
  if (!quad_slope_ok(p.slope) || !quad_shape(p.H, p.C, qs)) return false;
  const int G = 32 / qs.lpr;
  const int F = p.H * p.C;
  if (p.n_dst > 0) {
    const int rpw = pick_rpw(p.n_dst, G);
    const int64_t nchunks = ceil_div(p.n_dst, rpw);
    const int nb = quad_dst_blocks(p.n_dst, rpw, quad_dst_lean(p) ? 4 : quad_dst_minb());
#define X(V, L, Hh)                                                                                   \
  if (qs.v == V && qs.lpr == L && p.H == Hh && p.e_logit != nullptr) {                                \
    const size_t smem = static_cast<size_t>(kQW) * (kDDst + 2) * V * 512;                             \
    if (quad_dst_lean(p)) {                                                                           \
      const size_t smem4 = static_cast<size_t>(kQW) * (kDDstLean + 3) * V * 512;                      \
      auto kern = gatv2_bwd_dst_lg_quad_kernel<V, L, Hh, kDDstLean, 4, true>;                         \
      if (smem4 > 48 * 1024 && !set_smem(kern, smem4)) return false;                                  \
      kern<<<nb, kQThreads, smem4, stream>>>(p, rpw, nchunks, quad_sort());                           \
    } else {                                                                                          \
      auto kern = gatv2_bwd_dst_lg_quad_kernel<V, L, Hh, kDDst, 3, false>;                            \
      if (smem > 48 * 1024 && !set_smem(kern, smem)) return false;                                    \
      kern<<<nb, kQThreads, smem, stream>>>(p, rpw, nchunks, quad_sort());                            \
    }                                                                                                 \
  } else if (qs.v == V && qs.lpr == L && p.H == Hh) {                                                 \
    const size_t smem = static_cast<size_t>(kQW) * (kDDst + 1) * V * 512;                             \
    if (quad_dst_minb() == 4) {                                                                       \
      auto kern = gatv2_bwd_dst_quad_kernel<V, L, Hh, kDDst, 4>;                                      \
      if (smem > 48 * 1024 && !set_smem(kern, smem)) return false;                                    \
      kern<<<nb, kQThreads, smem, stream>>>(p, rpw, nchunks, quad_sort());                                         \
    } else {                                                                                          \
      auto kern = gatv2_bwd_dst_quad_kernel<V, L, Hh, kDDst, 3>;                                      \
      if (smem > 48 * 1024 && !set_smem(kern, smem)) return false;                                    \
      kern<<<nb, kQThreads, smem, stream>>>(p, rpw, nchunks, quad_sort());                                         \
    }                                                                                                 \
  }
    SGB_QUAD_COMBOS(X)
#undef X
    quad_colsum_kernel<<<static_cast<unsigned>(ceil_div(2 * F, 32)), dim3(32, 8), 0, stream>>>(p.partial, nb, 2 * F, F,
                                                                                               grad_att, grad_bias);
  }
  if (p.n_src > 0 && p.t_rowptr != nullptr) {
    const int rpw = pick_rpw(p.n_src, G);
    const int64_t nchunks = ceil_div(p.n_src, rpw);
    const unsigned blocks = static_cast<unsigned>(ceil_div(nchunks, kQW));
#define X(V, L, Hh)                                                                                   \
  if (qs.v == V && qs.lpr == L && p.H == Hh) {                                                        \
    const size_t smem = static_cast<size_t>(kQW) * kDSrc * (V + 1) * 512;                         \
    auto kern = gatv2_bwd_src_quad_kernel<V, L, Hh, kDSrc>;                                           \
    if (smem > 48 * 1024 && !set_smem(kern, smem)) return false;                                      \
    kern<<<blocks, kQThreads, smem, stream>>>(p, rpw, nchunks, quad_sort());                                       \
  }
    SGB_QUAD_COMBOS(X)
#undef X
  }
  return true;
}

}  // namespace sgb
