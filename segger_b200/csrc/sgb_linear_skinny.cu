// Tall-and-skinny projections (reduction depth K <= 16) on the fp32 pipe.
//
// After the positional front end is put in its low-rank form (sgb_poscheb_fwd: 12 Chebyshev basis columns stand
// in for the 256 sinusoid columns), the first Linear of the positional MLP
// (/root/reference/src/segger/models/ist_encoder.py:43-47,76-79) is y[2N, dim] = T[2N, 12] W_eff^T + b and its
// weight gradient dW_eff[dim, 12] = dy^T T.  Both move ~1 GB and do 12 FMAs per output: pure HBM streaming, for
// which a 128 x 64 x 32 tensor-core tile pipeline (zero-filled to K = 32) is the wrong tool (measured 0.61 / 0.63 ms
// against 0.17 / 0.10 ms of compulsory traffic).  Here: weights in shared memory, one thread per 128-bit output
// chunk (forward) or per output column with K accumulators in registers (wgrad), exact round-to-nearest FMAs,
// fixed-order two-stage reduction (bit-reproducible).
#include "sgb_api_internal.cuh"
#include "sgb_linear.cuh"
#include <algorithm>

namespace sgb {
namespace {

constexpr int kMaxK = 16;
constexpr int kMaxN = 256;

__device__ __forceinline__ float sk_act(float x, int act) {
  if (act == SGB_ACT_GELU) return gelu_erf(x);
  if (act == SGB_ACT_SILU) return x / (1.0f + __expf(-x));
  return x;
}

// y[r, 4q..4q+3] = sum_k x[r,k] w[4q+j,k] + b;   thread per (row, 128-bit column chunk), K4 = K/4 compile-time
template <int K4>
__global__ void __launch_bounds__(256)
skinny_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, int64_t ldw,
                  const float* __restrict__ b, int64_t M, int N, float* __restrict__ y, int64_t ldy, int act,
                  float* __restrict__ y_act, int64_t ldya) {
  constexpr int K = 4 * K4;
  __shared__ __align__(16) float wT[K * kMaxN];   // [k][n]
  __shared__ __align__(16) float bs[kMaxN];
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i - n * K;
    wT[k * N + n] = __ldg(w + static_cast<int64_t>(n) * ldw + k);
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) bs[i] = b ? __ldg(b + i) : 0.f;
  __syncthreads();
  const int Q = N / 4;
  const int64_t total = M * Q;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / Q;
    const int c = static_cast<int>(i - r * Q) * 4;
    float4 acc = *reinterpret_cast<const float4*>(bs + c);
    const float* xr = x + r * ldx;
#pragma unroll
    for (int k4 = 0; k4 < K4; ++k4) {
      const float4 xv = ldg4(xr + 4 * k4);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(wT + (4 * k4 + j) * N + c);
        acc.x = fmaf(xs[j], wv.x, acc.x); acc.y = fmaf(xs[j], wv.y, acc.y);
        acc.z = fmaf(xs[j], wv.z, acc.z); acc.w = fmaf(xs[j], wv.w, acc.w);
      }
    }
    st4(y + r * ldy + c, acc);
    if (y_act)
      st4(y_act + r * ldya + c, make_float4(sk_act(acc.x, act), sk_act(acc.y, act), sk_act(acc.z, act), sk_act(acc.w, act)));
  }
}

// partial[blk][c][0..K) = sum over the CTA's rows of dy[r,c] x[r,k];  partial[blk][c][K] = sum dy[r,c]
// thread (rg, c): rows r0 + rg, r0 + rg + RG, ...  (RG = blockDim / N row groups, fixed order within a thread)
template <int K4>
__global__ void __launch_bounds__(256)
skinny_wgrad_kernel(const float* __restrict__ dy, int64_t ldy, const float* __restrict__ x, int64_t ldx, int64_t M, int N,
                    int64_t rows_per_block, float* __restrict__ partial) {
  constexpr int K = 4 * K4, KP = K + 1;
  extern __shared__ float red[];                 // [RG][N][KP]
  const int RG = blockDim.x / N;
  const int c = threadIdx.x % N, rg = threadIdx.x / N;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  float bsum = 0.f;
  if (rg < RG) {
    constexpr int U = 4;                         // rows in flight per thread
    int64_t r = r0 + rg;
    for (; r + static_cast<int64_t>(U - 1) * RG < r1; r += static_cast<int64_t>(U) * RG) {
      float a[U];
      float4 xv[U][K4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = r + static_cast<int64_t>(u) * RG;
        a[u] = __ldg(dy + rr * ldy + c);
#pragma unroll
        for (int k4 = 0; k4 < K4; ++k4) xv[u][k4] = ldg4(x + rr * ldx + 4 * k4);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        bsum += a[u];
#pragma unroll
        for (int k4 = 0; k4 < K4; ++k4) {
          acc[4 * k4 + 0] = fmaf(a[u], xv[u][k4].x, acc[4 * k4 + 0]);
          acc[4 * k4 + 1] = fmaf(a[u], xv[u][k4].y, acc[4 * k4 + 1]);
          acc[4 * k4 + 2] = fmaf(a[u], xv[u][k4].z, acc[4 * k4 + 2]);
          acc[4 * k4 + 3] = fmaf(a[u], xv[u][k4].w, acc[4 * k4 + 3]);
        }
      }
    }
    for (; r < r1; r += RG) {
      const float a = __ldg(dy + r * ldy + c);
      bsum += a;
#pragma unroll
      for (int k4 = 0; k4 < K4; ++k4) {
        const float4 v = ldg4(x + r * ldx + 4 * k4);
        acc[4 * k4 + 0] = fmaf(a, v.x, acc[4 * k4 + 0]); acc[4 * k4 + 1] = fmaf(a, v.y, acc[4 * k4 + 1]);
        acc[4 * k4 + 2] = fmaf(a, v.z, acc[4 * k4 + 2]); acc[4 * k4 + 3] = fmaf(a, v.w, acc[4 * k4 + 3]);
      }
    }
    float* mine = red + (static_cast<size_t>(rg) * N + c) * KP;
#pragma unroll
    for (int k = 0; k < K; ++k) mine[k] = acc[k];
    mine[K] = bsum;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N * KP; i += blockDim.x) {
    float s = red[i];
    for (int g = 1; g < RG; ++g) s += red[static_cast<size_t>(g) * N * KP + i];   // fixed order
    partial[static_cast<int64_t>(blockIdx.x) * N * KP + i] = s;
  }
}

// 128-bit variant (N % 4 == 0, N <= 128, dy 16-byte aligned with ldy % 4 == 0): thread (rg, q) owns the four
// columns 4q..4q+3 -> one 128-bit dy load per row instead of four 32-bit ones, 4K accumulators in registers.
// The N/4 threads of a row group sit in one warp together with 32/(N/4) - 1 other row groups: those are folded with
// shuffles (fixed order), the per-warp sums go through shared memory.  Same partial layout as the scalar kernel.
template <int K4>
__global__ void __launch_bounds__(256)
skinny_wgrad_vec_kernel(const float* __restrict__ dy, int64_t ldy, const float* __restrict__ x, int64_t ldx, int64_t M, int N,
                        int64_t rows_per_block, float* __restrict__ partial) {
  constexpr int K = 4 * K4, KP = K + 1;
  extern __shared__ float red[];                 // [8 warps][N][KP]
  const int Q = N / 4;                           // threads per row: 1, 2, 4, 8, 16 or 32 (N = 4 .. 128, power of two)
  const int RG = blockDim.x / Q;
  const int q = threadIdx.x % Q, rg = threadIdx.x / Q;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  float acc[4][K];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[j][k] = 0.f;
  constexpr int U = 2;
  int64_t r = r0 + rg;
  auto accumulate = [&](const float4 a, const float4 (&xv)[K4]) {
    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bsum[j] += av[j];
#pragma unroll
      for (int k4 = 0; k4 < K4; ++k4) {
        acc[j][4 * k4 + 0] = fmaf(av[j], xv[k4].x, acc[j][4 * k4 + 0]);
        acc[j][4 * k4 + 1] = fmaf(av[j], xv[k4].y, acc[j][4 * k4 + 1]);
        acc[j][4 * k4 + 2] = fmaf(av[j], xv[k4].z, acc[j][4 * k4 + 2]);
        acc[j][4 * k4 + 3] = fmaf(av[j], xv[k4].w, acc[j][4 * k4 + 3]);
      }
    }
  };
  for (; r + static_cast<int64_t>(U - 1) * RG < r1; r += static_cast<int64_t>(U) * RG) {
    float4 a[U], xv[U][K4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + static_cast<int64_t>(u) * RG;
      a[u] = ldg4(dy + rr * ldy + 4 * q);
#pragma unroll
      for (int k4 = 0; k4 < K4; ++k4) xv[u][k4] = ldg4(x + rr * ldx + 4 * k4);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) accumulate(a[u], xv[u]);
  }
  for (; r < r1; r += RG) {
    float4 xv[K4];
    const float4 a = ldg4(dy + r * ldy + 4 * q);
#pragma unroll
    for (int k4 = 0; k4 < K4; ++k4) xv[k4] = ldg4(x + r * ldx + 4 * k4);
    accumulate(a, xv);
  }
  // fold the row groups that share a warp (lanes q, q + Q, q + 2Q, ...): xor butterfly over the group bits
  for (int o = Q; o < 32; o <<= 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bsum[j] += __shfl_xor_sync(kFull, bsum[j], o);
#pragma unroll
      for (int k = 0; k < K; ++k) acc[j][k] += __shfl_xor_sync(kFull, acc[j][k], o);
    }
  }
  if (lane < Q) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float* mine = red + (static_cast<size_t>(warp) * N + 4 * lane + j) * KP;
#pragma unroll
      for (int k = 0; k < K; ++k) mine[k] = acc[j][k];
      mine[K] = bsum[j];
    }
  }
  __syncthreads();
  const int nw = blockDim.x / 32;
  for (int i = threadIdx.x; i < N * KP; i += blockDim.x) {
    float s_ = red[i];
    for (int w = 1; w < nw; ++w) s_ += red[static_cast<size_t>(w) * N * KP + i];   // fixed order
    partial[static_cast<int64_t>(blockIdx.x) * N * KP + i] = s_;
  }
}

// dw[c, k] (+)= sum_blk partial[blk][c][k];  db[c] (+)= sum_blk partial[blk][c][K]   (fixed order)
__global__ void skinny_wgrad_reduce_kernel(const float* __restrict__ partial, int nblk, int N, int K, float* __restrict__ dw,
                                           int64_t lddw, float* __restrict__ db, int accumulate) {
  const int KP = K + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * KP) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[static_cast<int64_t>(b) * N * KP + i];
  const int c = i / KP, k = i - c * KP;
  if (k < K) {
    float* d = dw + static_cast<int64_t>(c) * lddw + k;
    *d = accumulate ? *d + s : s;
  } else if (db) {
    db[c] = accumulate ? db[c] + s : s;
  }
}

int wgrad_blocks(int64_t M) {
  const int64_t want = ceil_div(M, 1024);
  const int64_t cap = static_cast<int64_t>(sm_count()) * 4;
  return static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

bool skinny_linear_fwd_ok(const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K, const float* y, int64_t ldy,
                          const float* y_act, int64_t ldya) {
  return M >= 4096 && K >= 4 && K <= kMaxK && K % 4 == 0 && N >= 4 && N <= kMaxN && N % 4 == 0 && aligned16(x) &&
         ldx % 4 == 0 && aligned16(y) && ldy % 4 == 0 && (!y_act || (aligned16(y_act) && ldya % 4 == 0));
}

int skinny_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* b, int64_t M, int64_t N,
                      int64_t K, float* y, int64_t ldy, int act, float* y_act, int64_t ldya, cudaStream_t stream) {
  const int64_t total = M * (N / 4);
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(total, 256), static_cast<int64_t>(sm_count()) * 16));
  const int n = static_cast<int>(N);
#define SK_FWD(K4) skinny_fwd_kernel<K4><<<blocks, 256, 0, stream>>>(x, ldx, w, ldw, b, M, n, y, ldy, act, y_act, ldya)
  switch (K / 4) {
    case 1: SK_FWD(1); break;
    case 2: SK_FWD(2); break;
    case 3: SK_FWD(3); break;
    default: SK_FWD(4); break;
  }
#undef SK_FWD
  return check_launch("skinny_linear_fwd");
}

bool skinny_linear_wgrad_ok(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K) {
  (void)dy; (void)ldy;
  return M >= 4096 && K >= 4 && K <= kMaxK && K % 4 == 0 && N >= 1 && N <= kMaxN && aligned16(x) && ldx % 4 == 0;
}

size_t skinny_linear_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  return align_up(static_cast<size_t>(wgrad_blocks(M)) * N * (K + 1) * sizeof(float));
}

int skinny_linear_wgrad(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K,
                        float* dw, int64_t lddw, float* db, int accumulate, void* ws, cudaStream_t stream) {
  const int nblk = wgrad_blocks(M);
  const int64_t rpb = ceil_div(M, nblk);
  const int n = static_cast<int>(N), k = static_cast<int>(K);
  float* partial = static_cast<float*>(ws);
  const bool pow2 = (n & (n - 1)) == 0;
  if (pow2 && n >= 4 && n <= 128 && aligned16(dy) && ldy % 4 == 0) {
    const size_t smem = static_cast<size_t>(8) * n * (k + 1) * sizeof(float);      // <= 8 * 128 * 17 * 4 = 68 KB
#define SK_WGV(K4)                                                                                                   \
  do {                                                                                                               \
    auto kern = skinny_wgrad_vec_kernel<K4>;                                                                         \
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024) != cudaSuccess) \
      return set_error(SGB_ERR_CUDA, "skinny wgrad: cudaFuncSetAttribute failed");                                  \
    kern<<<nblk, 256, smem, stream>>>(dy, ldy, x, ldx, M, n, rpb, partial);                                          \
  } while (0)
    switch (K / 4) {
      case 1: SK_WGV(1); break;
      case 2: SK_WGV(2); break;
      case 3: SK_WGV(3); break;
      default: SK_WGV(4); break;
    }
#undef SK_WGV
  } else {
    const int RG = 256 / n;
    const size_t smem = static_cast<size_t>(RG) * n * (k + 1) * sizeof(float);
#define SK_WG(K4) skinny_wgrad_kernel<K4><<<nblk, 256, smem, stream>>>(dy, ldy, x, ldx, M, n, rpb, partial)
    switch (K / 4) {
      case 1: SK_WG(1); break;
      case 2: SK_WG(2); break;
      case 3: SK_WG(3); break;
      default: SK_WG(4); break;
    }
#undef SK_WG
  }
  skinny_wgrad_reduce_kernel<<<static_cast<unsigned>(ceil_div(N * (K + 1), 128)), 128, 0, stream>>>(partial, nblk, n, k, dw, lddw, db,
                                                                                                accumulate);
  return check_launch("skinny_linear_wgrad");
}

}  // namespace sgb
