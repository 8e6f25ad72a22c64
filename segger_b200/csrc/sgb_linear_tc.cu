// Error-compensated 3xTF32 GEMM on Blackwell 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
//
//   C[m,n] = sum_k A(m,k) * B(n,k)            fp32 in, fp32 out, ~2^-21 relative product error
//
// Each fp32 operand x is split on the fly into x_hi = trunc_tf32(x) and x_lo = x - x_hi and
//   A*B ~= A_hi*B_hi + A_lo*B_hi + A_hi*B_lo            (three kind::tf32 MMAs per k-step)
// which keeps the projections inside the 1e-4 fp32 parity bar that a single TF32 pass (2^-11) misses.
//
// Persistent, warp-specialised CTA (one per SM, 416 threads):
//   warps 0-7   producers: 128-bit coalesced global loads of the A / B k-slab, hi/lo split in
//               registers, swizzle-128B stores into the UMMA canonical shared-memory layout
//               (K-major or MN-major, so x W^T, dy W and dy^T x all run without a transpose),
//               fence.proxy.async + mbarrier arrive.
//   warp  8     MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BN, K=8), tcgen05.commit
//               releases the smem stage / publishes the TMEM accumulator; owns TMEM alloc/dealloc.
//   warps 9-12  epilogue: tcgen05.ld (32 lanes x 32 columns) -> bias / activation / act' / accumulate
//               -> 128-bit global stores; the accumulator is double-buffered in TMEM so the epilogue
//               of tile i overlaps the MMAs of tile i+1.
// Split-K over the reduction (wgrad: reduction = number of rows) writes per-split partials that a
// fixed-order reduce kernel sums -> deterministic.
//
// Replaces cuBLAS SGEMM reached from PyG Linear / torch.nn.Linear
// (/root/reference/src/segger/models/ist_encoder.py:43-47,111-131,261,282-286).
#include "sgb_api_internal.cuh"
#include "sgb_linear.cuh"

namespace sgb {
namespace {

constexpr int BM = 128;            // UMMA M (cta_group::1)
constexpr int BK = 32;             // fp32 per 128-byte swizzle row = reduction elements per stage
constexpr int kProducerWarps = 8;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kMmaWarp = kProducerWarps;
constexpr int kEpiWarp0 = kProducerWarps + 1;
constexpr int kThreads = (kProducerWarps + 1 + 4) * 32;   // 416
constexpr int kMaxStages = 4;

struct TcParams {
  const float *A, *B;
  int64_t lda, ldb;
  int64_t M, N, K;          // MN extent of A, MN extent of B, reduction extent
  int splits;
  int64_t k_per_split;      // multiple of BK
  float* C;                 // output, or split-K partials [splits][M][N] (ldc = N)
  int64_t ldc;
  const float* bias;
  int act;
  float* C_act;
  int64_t ldca;
  int accumulate;
  const float* act_pre;
  int64_t ld_pre;
  int stages;
  int terms;                // 3: hi*hi + lo*hi + hi*lo (~2^-22);  4: + lo*lo (fp32-exact products)
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- descriptors (cute/arch/mma_sm100_desc.hpp bit layout) -----------------------------------
// instruction descriptor: c=F32, a=b=TF32, majors, N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
}
// shared-memory matrix descriptor, version 1 (Blackwell).  layout_type: 2 = SWIZZLE_128B (K-major
// operands), 1 = SWIZZLE_128B_BASE32B (the only layout tcgen05 accepts for MN-major tf32 operands).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  const uint64_t lo = static_cast<uint64_t>((saddr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16);
  const uint64_t hi = static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) | (1ull << 14) | (static_cast<uint64_t>(layout_type) << 29);
  return lo | (hi << 32);
}

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

// Round-to-nearest to the 11 significant bits of TF32 (low 13 mantissa bits cleared, so the tensor
// core's own truncation is a no-op).  x = hi + lo with hi = rn(x), lo = rn(x - hi) represents x to
// 2^-24 relative -- i.e. the split itself loses nothing of an fp32 value.
__device__ __forceinline__ float rn_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

// Operand slab loader: EXT rows of the MN dimension x BK reduction elements starting at (mn0, k0),
// written as hi / lo tiles in the canonical swizzle-128B layout.
//   K-major  (MN == false): element (mn, k) at src[mn * ld + k]; smem row = one mn, 128 B of k.
//   MN-major (MN == true) : element (mn, k) at src[k * ld + mn]; smem row = one k, 128 B of mn;
//                           512-byte atoms (4 k-rows, Swizzle<2,5,2>: 32-byte chunk ^= k % 4)
//                           ordered [k-group of 4][mn-block of 32].
template <int EXT, bool MN>
__device__ __forceinline__ void load_slab(const float* __restrict__ src, int64_t ld, int64_t mn0, int64_t mn_end,
                                          int64_t k0, int64_t k_end, float4 (&reg)[EXT * 8 / kProducerThreads], int t) {
  constexpr int NV = EXT * 8 / kProducerThreads;
  const int chunk = t & 7;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int ri = (t >> 3) + (kProducerThreads / 8) * i;       // 0 .. EXT-1 (K-major) or 0 .. 32*(EXT/32)-1
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!MN) {
      const int64_t mn = mn0 + ri, k = k0 + chunk * 4;
      if (mn < mn_end && k < k_end) v = ldg4(src + mn * ld + k);   // K % 4 == 0 guaranteed by the dispatcher
    } else {
      const int kk = ri & 31, blk = ri >> 5;
      const int64_t k = k0 + kk, mn = mn0 + blk * 32 + chunk * 4;
      if (k < k_end && mn < mn_end) v = ldg4(src + k * ld + mn);   // MN % 4 == 0 guaranteed
    }
    reg[i] = v;
  }
}

template <int EXT, bool MN>
__device__ __forceinline__ void store_slab(uint8_t* hi_tile, uint8_t* lo_tile, const float4 (&reg)[EXT * 8 / kProducerThreads], int t) {
  constexpr int NV = EXT * 8 / kProducerThreads;
  const int chunk = t & 7;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int ri = (t >> 3) + (kProducerThreads / 8) * i;
    uint32_t off;
    if (!MN) {
      off = (ri >> 3) * 1024 + (ri & 7) * 128 + ((chunk ^ (ri & 7)) << 4);
    } else {
      const int kk = ri & 31, blk = ri >> 5;
      const int c16 = ((((chunk >> 1) ^ (kk & 3)) << 1) | (chunk & 1));
      off = ((kk >> 2) * (EXT / 32) + blk) * 512 + (kk & 3) * 128 + (c16 << 4);
    }
    const float4 v = reg[i];
    float4 h, l;
    h.x = rn_tf32(v.x); l.x = rn_tf32(v.x - h.x);
    h.y = rn_tf32(v.y); l.y = rn_tf32(v.y - h.y);
    h.z = rn_tf32(v.z); l.z = rn_tf32(v.z - h.z);
    h.w = rn_tf32(v.w); l.w = rn_tf32(v.w - h.w);
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1) gemm_tf32x3_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr uint32_t kATile = BM * 128;            // bytes of one hi (or lo) A tile
  constexpr uint32_t kBTile = BN * 128;
  constexpr uint32_t kStageBytes = 2 * kATile + 2 * kBTile;
  constexpr uint32_t kTmemCols = (2 * BN <= 64) ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512));

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_holder;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stages = p.stages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), kProducerThreads);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&bar_tfull[a]), 1);
      mbar_init(smem_u32(&bar_tempty[a]), 128);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_holder)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_holder;

  const int64_t num_m = (p.M + BM - 1) / BM, num_n = (p.N + BN - 1) / BN;
  const int64_t tiles = num_m * num_n * p.splits;

  if (warp < kProducerWarps) {
    // ===================================== producers =====================================
    // Two register sets: the global loads of k-slab i+1 are in flight while slab i is converted and
    // stored, so a slab costs max(load latency / 2, store) instead of latency + store.
    const int t = threadIdx.x;
    int stage = 0;
    uint32_t phase = 0;
    constexpr int NA = BM * 8 / kProducerThreads, NB = BN * 8 / kProducerThreads;
    float4 ra0[NA], rb0[NB], ra1[NA], rb1[NB];
    int64_t tile = blockIdx.x, k0 = 0, ke = 0;
    auto tile_range = [&](int64_t tl, int64_t& kb_, int64_t& ke_) {
      const int64_t sp = tl / (num_n * num_m);
      kb_ = sp * p.k_per_split;
      ke_ = min(p.K, kb_ + p.k_per_split);
    };
    auto issue = [&](int64_t tl, int64_t kk, int64_t kend, float4 (&ra)[NA], float4 (&rb)[NB]) {
      const int64_t nb = tl % num_n, mb = (tl / num_n) % num_m;
      load_slab<BM, A_MN>(p.A, p.lda, mb * BM, p.M, kk, kend, ra, t);
      load_slab<BN, B_MN>(p.B, p.ldb, nb * BN, p.N, kk, kend, rb, t);
    };
    auto commit = [&](const float4 (&ra)[NA], const float4 (&rb)[NB]) {
      mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
      uint8_t* st = smem + static_cast<size_t>(stage) * kStageBytes;
      store_slab<BM, A_MN>(st, st + kATile, ra, t);
      store_slab<BN, B_MN>(st + 2 * kATile, st + 2 * kATile + kBTile, rb, t);
      fence_proxy_async();
      mbar_arrive(smem_u32(&bar_full[stage]));
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    };
    // advance (tile, k0) to the next k-slab of this CTA's work list; returns false when exhausted
    auto advance = [&]() -> bool {
      k0 += BK;
      if (k0 >= ke) {
        tile += gridDim.x;
        if (tile >= tiles) return false;
        tile_range(tile, k0, ke);
      }
      return true;
    };
    bool live = tile < tiles;
    if (live) {
      tile_range(tile, k0, ke);
      issue(tile, k0, ke, ra0, rb0);
    }
    while (live) {
      bool more = advance();
      if (more) issue(tile, k0, ke, ra1, rb1);
      commit(ra0, rb0);
      if (!more) break;
      more = advance();
      if (more) issue(tile, k0, ke, ra0, rb0);
      commit(ra1, rb1);
      live = more;
    }
  } else if (warp == kMmaWarp) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN);
      // K-major: LBO unused (1 -> encoded 16 B), SBO = 1024 B between 8-row groups, +32 B per k-step.
      // MN-major: LBO = 512 B between 32-wide MN blocks, SBO = (EXT/32)*512 B between 4-deep k groups;
      //           one K=8 MMA spans two k groups -> +2*SBO per k-step.
      constexpr uint32_t a_lbo = A_MN ? 512u : 16u, a_sbo = A_MN ? (BM / 32) * 512u : 1024u, a_step = A_MN ? 2 * a_sbo : 32u;
      constexpr uint32_t b_lbo = B_MN ? 512u : 16u, b_sbo = B_MN ? (BN / 32) * 512u : 1024u, b_step = B_MN ? 2 * b_sbo : 32u;
      constexpr uint32_t a_lt = A_MN ? 1u : 2u, b_lt = B_MN ? 1u : 2u;
      int stage = 0;
      uint32_t phase = 0;
      int64_t it = 0;
      for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int acc = static_cast<int>(it & 1);
        const uint32_t acc_phase = static_cast<uint32_t>((it >> 1) & 1);
        mbar_wait(smem_u32(&bar_tempty[acc]), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        const int64_t sp = tile / (num_n * num_m);
        const int64_t kb = sp * p.k_per_split, ke = min(p.K, kb + p.k_per_split);
        uint32_t accum = 0;
        for (int64_t k0 = kb; k0 < ke; k0 += BK) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + static_cast<size_t>(stage) * kStageBytes);
          const uint32_t a_hi = st, a_lo = st + kATile, b_hi = st + 2 * kATile, b_lo = b_hi + kBTile;
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            const uint64_t dah = make_sdesc(a_hi + j * a_step, a_lbo, a_sbo, a_lt), dal = make_sdesc(a_lo + j * a_step, a_lbo, a_sbo, a_lt);
            const uint64_t dbh = make_sdesc(b_hi + j * b_step, b_lbo, b_sbo, b_lt), dbl = make_sdesc(b_lo + j * b_step, b_lbo, b_sbo, b_lt);
            if (p.terms >= 4) { tc_mma_tf32(d_tmem, dal, dbl, idesc, accum); accum = 1u; }
            tc_mma_tf32(d_tmem, dal, dbh, idesc, accum);   // small terms first
            tc_mma_tf32(d_tmem, dah, dbl, idesc, 1u);
            tc_mma_tf32(d_tmem, dah, dbh, idesc, 1u);
            accum = 1u;
          }
          tc_commit(smem_u32(&bar_empty[stage]));           // frees this smem stage when the MMAs retire
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        if (kb >= ke) {
          // empty reduction range (K == 0): nothing was issued; the epilogue must still see zeros
          // -> handled by the dispatcher (K >= 1 and every split non-empty).
        }
        tc_commit(smem_u32(&bar_tfull[acc]));               // accumulator complete
      }
    }
  } else {
    // ===================================== epilogue =====================================
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
      const int acc = static_cast<int>(it & 1);
      const uint32_t acc_phase = static_cast<uint32_t>((it >> 1) & 1);
      const int64_t nb = tile % num_n, mb = (tile / num_n) % num_m, sp = tile / (num_n * num_m);
      const int64_t row = mb * BM + q * 32 + lane;
      const int64_t n0 = nb * BN;
      mbar_wait(smem_u32(&bar_tfull[acc]), acc_phase);
      tc_fence_after();
      float* C = p.C + (p.splits > 1 ? sp * p.M * p.N : 0);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tc_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN + c0), r);
        if (row < p.M && n0 + c0 < p.N) {
          float* dst = C + row * p.ldc + n0 + c0;
          const bool full = (n0 + c0 + 32 <= p.N) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float v[4] = {__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])};
            const int64_t n = n0 + c0 + j;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (full || n + e < p.N) {
                if (p.bias) v[e] += __ldg(p.bias + n + e);
                if (p.accumulate) v[e] += dst[j + e];
                if (p.act_pre) v[e] *= act_grad(__ldg(p.act_pre + row * p.ld_pre + n + e), p.act);
              }
            }
            if (full) {
              st4(dst + j, make_float4(v[0], v[1], v[2], v[3]));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) if (n + e < p.N) dst[j + e] = v[e];
            }
            if (p.C_act) {
              float* da = p.C_act + row * p.ldca + n;
#pragma unroll
              for (int e = 0; e < 4; ++e) if (full || n + e < p.N) da[e] = act_apply(v[e], p.act);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[acc]));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

__global__ void tc_splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t MN, int64_t N,
                                        float* __restrict__ out, int64_t ldo, int accumulate) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[static_cast<int64_t>(z) * MN + i];
  float* dst = out + (i / N) * ldo + (i % N);
  *dst = accumulate ? *dst + s : s;
}

template <int BN>
constexpr size_t stage_bytes() { return static_cast<size_t>(2 * BM * 128 + 2 * BN * 128); }

template <int BN, bool A_MN, bool B_MN>
int launch_tc(TcParams p, cudaStream_t stream) {
  constexpr size_t kBudget = 220 * 1024;
  int stages = static_cast<int>((kBudget - 1024) / stage_bytes<BN>());
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const size_t smem = stages * stage_bytes<BN>() + 1024;
  static bool configured = false;   // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return set_error(SGB_ERR_CUDA, "tc gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int64_t tiles = ceil_div(p.M, BM) * ceil_div(p.N, BN) * p.splits;
  const int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  gemm_tf32x3_kernel<BN, A_MN, B_MN><<<grid, kThreads, smem, stream>>>(p);
  return check_launch("gemm_tf32x3");
}

template <bool A_MN, bool B_MN>
int dispatch_bn(const TcParams& p, cudaStream_t stream) {
  const int64_t N = p.N;
  if (N <= 64) return launch_tc<64, A_MN, B_MN>(p, stream);
  if (N % 256 == 0) return launch_tc<256, A_MN, B_MN>(p, stream);
  if (N % 192 == 0) return launch_tc<192, A_MN, B_MN>(p, stream);
  if (N <= 128 || N % 128 == 0) return launch_tc<128, A_MN, B_MN>(p, stream);
  if (N <= 192) return launch_tc<192, A_MN, B_MN>(p, stream);
  return launch_tc<256, A_MN, B_MN>(p, stream);
}

bool is_blackwell() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
      return false;
    cached = (major == 10) ? 1 : 0;
  }
  return cached == 1;
}

bool ok_ptr(const float* p, int64_t ld) { return aligned16(p) && ld % 4 == 0; }

}  // namespace

bool tc_enabled() {
  static int flag = -1;
  if (flag < 0) {
    const char* e = getenv("SEGGER_B200_GEMM");
    flag = (e && (e[0] == 's' || e[0] == 'S')) ? 0 : 1;    // SEGGER_B200_GEMM=simt forces the exact-fp32 path
  }
  return flag == 1 && is_blackwell();
}

// Number of TF32 products per fp32 product.  Forward projections feed the attention logits, whose
// softmax gradient amplifies feature error by |x| / |x_j - o_i|: they get the 4-term (fp32-exact)
// scheme; gradients propagate error linearly and use 3 terms.  SEGGER_B200_TF32_TERMS overrides.
static int tc_terms(bool forward) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("SEGGER_B200_TF32_TERMS");
    env = (e && (e[0] == '3' || e[0] == '4')) ? (e[0] - '0') : 0;
  }
  if (env) return env;
  return forward ? 4 : 3;
}

bool tc_linear_fwd_ok(const float* x, int64_t ldx, const float* w, int64_t ldw, int64_t M, int64_t N, int64_t K) {
  return tc_enabled() && M >= 1 && N >= 8 && K >= 8 && K % 4 == 0 && ok_ptr(x, ldx) && ok_ptr(w, ldw);
}
int tc_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* b, int64_t M, int64_t N, int64_t K,
                  float* y, int64_t ldy, int act, float* y_act, int64_t ldya, cudaStream_t stream) {
  TcParams p{};
  p.A = x; p.lda = ldx; p.B = w; p.ldb = ldw; p.M = M; p.N = N; p.K = K; p.splits = 1;
  p.k_per_split = ceil_div(K, BK) * BK;
  p.C = y; p.ldc = ldy; p.bias = b; p.act = act; p.C_act = y_act; p.ldca = ldya;
  p.terms = tc_terms(true);
  return dispatch_bn<false, false>(p, stream);
}

bool tc_linear_dgrad_ok(const float* dy, int64_t ldy, const float* w, int64_t ldw, int64_t M, int64_t N, int64_t K) {
  // dx[M,K] = dy[M,N] w[N,K]: A = dy (K-major over n), B(k', n) = w[n, k'] (MN-major)
  return tc_enabled() && M >= 1 && K >= 8 && N >= 8 && N % 4 == 0 && K % 4 == 0 && ok_ptr(dy, ldy) && ok_ptr(w, ldw);
}
int tc_linear_dgrad(const float* dy, int64_t ldy, const float* w, int64_t ldw, int64_t M, int64_t N, int64_t K, float* dx,
                    int64_t ldx, int accumulate, int act, const float* act_pre, int64_t ld_pre, cudaStream_t stream) {
  TcParams p{};
  p.A = dy; p.lda = ldy; p.B = w; p.ldb = ldw; p.M = M; p.N = K; p.K = N; p.splits = 1;
  p.k_per_split = ceil_div(N, BK) * BK;
  p.C = dx; p.ldc = ldx; p.accumulate = accumulate; p.act = act; p.act_pre = act_pre; p.ld_pre = ld_pre;
  p.terms = tc_terms(false);
  return dispatch_bn<false, true>(p, stream);
}

static int tc_wgrad_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t bn = K <= 64 ? 64 : (K % 256 == 0 ? 256 : (K % 192 == 0 ? 192 : (K <= 128 || K % 128 == 0 ? 128 : (K <= 192 ? 192 : 256))));
  const int64_t tiles = ceil_div(N, BM) * ceil_div(K, bn);
  int64_t s = sm_count() / tiles;
  const int64_t max_by_rows = ceil_div(M, 8 * BK);     // at least 8 k-stages per split
  if (s > max_by_rows) s = max_by_rows;
  if (s < 1) s = 1;
  return static_cast<int>(s);
}
bool tc_linear_wgrad_ok(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K) {
  // dw[N,K] = dy^T x: A(n, m) = dy[m, n] (MN-major), B(k, m) = x[m, k] (MN-major), reduction over m
  return tc_enabled() && M >= 256 && N >= 8 && K >= 8 && N % 4 == 0 && K % 4 == 0 && ok_ptr(dy, ldy) && ok_ptr(x, ldx);
}
size_t tc_linear_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  return align_up(static_cast<size_t>(tc_wgrad_splits(M, N, K)) * N * K * sizeof(float));
}
int tc_linear_wgrad(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K, float* dw,
                    int64_t lddw, int accumulate, void* ws, cudaStream_t stream) {
  TcParams p{};
  const int splits = tc_wgrad_splits(M, N, K);
  p.A = dy; p.lda = ldy; p.B = x; p.ldb = ldx; p.M = N; p.N = K; p.K = M; p.splits = splits;
  p.k_per_split = ceil_div(ceil_div(M, splits), BK) * BK;
  p.terms = tc_terms(false);
  // every split must own at least one k-stage
  while (p.splits > 1 && static_cast<int64_t>(p.splits - 1) * p.k_per_split >= M) --p.splits;
  if (p.splits == 1) {
    p.C = dw; p.ldc = lddw; p.accumulate = accumulate;
    return dispatch_bn<true, true>(p, stream);
  }
  p.C = static_cast<float*>(ws); p.ldc = K;
  int rc = dispatch_bn<true, true>(p, stream);
  if (rc != SGB_OK) return rc;
  const int64_t MN = N * K;
  tc_splitk_reduce_kernel<<<static_cast<unsigned>(ceil_div(MN, 256)), 256, 0, stream>>>(static_cast<const float*>(ws), p.splits, MN, K,
                                                                                       dw, lddw, accumulate);
  return check_launch("tc_splitk_reduce");
}

}  // namespace sgb
