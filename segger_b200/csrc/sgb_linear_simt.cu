// fp32 SIMT GEMM family (register-tiled, double-buffered shared memory) for the dense projections.
// This is the exact-fp32 path: it serves shapes the tcgen05 3xTF32 kernel does not take (ragged K/N,
// tiny N such as the 64-wide output projection's odd cases, unaligned leading dimensions) and is
// the parity anchor for it.  One kernel template covers forward (x W^T), dgrad (dy W) and wgrad
// (dy^T x) through two operand-layout flags; wgrad uses deterministic split-K.
//
// Replaces cuBLAS SGEMM reached from PyG Linear / torch.nn.Linear
// (/root/reference/src/segger/models/ist_encoder.py:43-47,111-131,261,282-286).
#include "sgb_api_internal.cuh"
#include "sgb_linear.cuh"

namespace sgb {
namespace {

constexpr int BM = 128, BK = 16, THREADS = 256, PAD = 4;

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

struct GemmParams {
  // C[m,n] = sum_k A(m,k) * B(k,n)
  const float *A, *B;
  int64_t lda, ldb;
  int64_t M, N, K;
  int64_t k_chunk;       // reduction range per grid.z slice (multiple of BK)
  float* C;              // output (or split-K partials [z][M][N] when gridDim.z > 1, ldc = N)
  int64_t ldc;
  const float* bias;     // [N] or null
  int act;               // epilogue activation (forward) / activation derivative selector (dgrad)
  float* C_act;          // forward: act(C) or null
  int64_t ldca;
  int accumulate;        // C += old C
  const float* act_pre;  // dgrad: multiply by act'(act_pre[m,n])
  int64_t ld_pre;
  int vec_a, vec_b;      // 128-bit global loads legal for A / B
};

// A_KC: A(m,k) at A[m*lda + k] (k contiguous) else A[k*lda + m] (m contiguous)
// B_KC: B(k,n) at B[n*ldb + k] (k contiguous) else B[k*ldb + n] (n contiguous)
template <int BN, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS) sgemm_kernel(const GemmParams p) {
  constexpr int TN = BN / 16;  // columns per thread (8 or 4)
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int t = threadIdx.x;
  const int tx = t % 16, ty = t / 16;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * BM;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * BN;
  const int64_t kbeg = static_cast<int64_t>(blockIdx.z) * p.k_chunk;
  const int64_t kend = min(p.K, kbeg + p.k_chunk);

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[BN / 64];

  auto load_a = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A_KC) {
        const int64_t m = m0 + t / 4 + 64 * i;
        const int64_t k = k0 + (t % 4) * 4;
        if (m < p.M) {
          const float* src = p.A + m * p.lda + k;
          if (p.vec_a && k + 3 < kend) v = ldg4(src);
          else {
            if (k + 0 < kend) v.x = __ldg(src + 0);
            if (k + 1 < kend) v.y = __ldg(src + 1);
            if (k + 2 < kend) v.z = __ldg(src + 2);
            if (k + 3 < kend) v.w = __ldg(src + 3);
          }
        }
      } else {
        const int64_t k = k0 + t / 32 + 8 * i;
        const int64_t m = m0 + (t % 32) * 4;
        if (k < kend) {
          const float* src = p.A + k * p.lda + m;
          if (p.vec_a && m + 3 < p.M) v = ldg4(src);
          else {
            if (m + 0 < p.M) v.x = __ldg(src + 0);
            if (m + 1 < p.M) v.y = __ldg(src + 1);
            if (m + 2 < p.M) v.z = __ldg(src + 2);
            if (m + 3 < p.M) v.w = __ldg(src + 3);
          }
        }
      }
      ra[i] = v;
    }
  };
  auto load_b = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < BN / 64; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (B_KC) {
        const int64_t n = n0 + t / 4 + 64 * i;
        const int64_t k = k0 + (t % 4) * 4;
        if (n < p.N) {
          const float* src = p.B + n * p.ldb + k;
          if (p.vec_b && k + 3 < kend) v = ldg4(src);
          else {
            if (k + 0 < kend) v.x = __ldg(src + 0);
            if (k + 1 < kend) v.y = __ldg(src + 1);
            if (k + 2 < kend) v.z = __ldg(src + 2);
            if (k + 3 < kend) v.w = __ldg(src + 3);
          }
        }
      } else {
        // BN/4 float4 per k-row; 256 threads cover 256/(BN/4) k-rows per step
        constexpr int PER_ROW = BN / 4;
        constexpr int ROWS = THREADS / PER_ROW;     // 8 (BN=128) or 16 (BN=64)
        const int64_t k = k0 + t / PER_ROW + ROWS * i;
        const int64_t n = n0 + (t % PER_ROW) * 4;
        if (k < kend && (BN == 128 || i == 0)) {
          const float* src = p.B + k * p.ldb + n;
          if (p.vec_b && n + 3 < p.N) v = ldg4(src);
          else {
            if (n + 0 < p.N) v.x = __ldg(src + 0);
            if (n + 1 < p.N) v.y = __ldg(src + 1);
            if (n + 2 < p.N) v.z = __ldg(src + 2);
            if (n + 3 < p.N) v.w = __ldg(src + 3);
          }
        }
      }
      rb[i] = v;
    }
  };
  auto store_a = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (A_KC) {
        const int m = t / 4 + 64 * i, k = (t % 4) * 4;
        As[buf][k + 0][m] = ra[i].x; As[buf][k + 1][m] = ra[i].y;
        As[buf][k + 2][m] = ra[i].z; As[buf][k + 3][m] = ra[i].w;
      } else {
        const int k = t / 32 + 8 * i, m = (t % 32) * 4;
        *reinterpret_cast<float4*>(&As[buf][k][m]) = ra[i];
      }
    }
  };
  auto store_b = [&](int buf) {
#pragma unroll
    for (int i = 0; i < BN / 64; ++i) {
      if (B_KC) {
        const int n = t / 4 + 64 * i, k = (t % 4) * 4;
        Bs[buf][k + 0][n] = rb[i].x; Bs[buf][k + 1][n] = rb[i].y;
        Bs[buf][k + 2][n] = rb[i].z; Bs[buf][k + 3][n] = rb[i].w;
      } else {
        constexpr int PER_ROW = BN / 4;
        constexpr int ROWS = THREADS / PER_ROW;
        const int k = t / PER_ROW + ROWS * i, n = (t % PER_ROW) * 4;
        if (BN == 128 || i == 0) *reinterpret_cast<float4*>(&Bs[buf][k][n]) = rb[i];
      }
    }
  };

  const int64_t nk = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;
  if (nk > 0) {
    load_a(kbeg);
    load_b(kbeg);
    store_a(0);
    store_b(0);
  }
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int cur = static_cast<int>(kt & 1);
    if (kt + 1 < nk) {
      load_a(kbeg + (kt + 1) * BK);
      load_b(kbeg + (kt + 1) * BK);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
      if constexpr (TN == 8) {
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
        bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_a(cur ^ 1);
      store_b(cur ^ 1);
    }
    __syncthreads();
  }

  // epilogue
  float* C = p.C + (gridDim.z > 1 ? static_cast<int64_t>(blockIdx.z) * p.M * p.N : 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
#pragma unroll
    for (int jh = 0; jh < TN / 4; ++jh) {
      const int64_t n = n0 + jh * 64 + tx * 4;
      if (n >= p.N) continue;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      float* dst = C + m * p.ldc + n;
      const bool full = (n + 3 < p.N);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (n + q < p.N) {
          if (p.bias) v[q] += __ldg(p.bias + n + q);
          if (p.accumulate) v[q] += dst[q];
          if (p.act_pre) v[q] *= act_grad(__ldg(p.act_pre + m * p.ld_pre + n + q), p.act);
        }
      }
      const bool vec_c = full && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
      if (vec_c) st4(dst, make_float4(v[0], v[1], v[2], v[3]));
      else {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (n + q < p.N) dst[q] = v[q];
      }
      if (p.C_act) {
        float* da = p.C_act + m * p.ldca + n;
#pragma unroll
        for (int q = 0; q < 4; ++q) if (n + q < p.N) da[q] = act_apply(v[q], p.act);
      }
    }
  }
}

// out[m,n] (+)= sum_z part[z][m][n], fixed order
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int64_t MN, int64_t N,
                                     float* __restrict__ out, int64_t ldo, int accumulate) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= MN) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[static_cast<int64_t>(z) * MN + i];
  float* dst = out + (i / N) * ldo + (i % N);
  *dst = accumulate ? *dst + s : s;
}

// column sums of dy [M,N]: stage 1 per-CTA partial over a row chunk, stage 2 ordered reduce
__global__ void __launch_bounds__(256) colsum_stage1_kernel(const float* __restrict__ dy, int64_t ldy, int64_t M,
                                                            int64_t N, int64_t rows_per_block, float* __restrict__ part) {
  // block (x = column tile of 64, y = row chunk); 256 threads = 4 row lanes x 64 columns
  __shared__ float red[4][64];
  const int c = threadIdx.x % 64, rl = threadIdx.x / 64;
  const int64_t n = static_cast<int64_t>(blockIdx.x) * 64 + c;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int64_t r = r0 + rl; r < r1; r += 4) s += __ldg(dy + r * ldy + n);
  red[rl][c] = s;
  __syncthreads();
  if (rl == 0 && n < N) part[static_cast<int64_t>(blockIdx.y) * N + n] = (red[0][c] + red[1][c]) + (red[2][c] + red[3][c]);
}
__global__ void colsum_stage2_kernel(const float* __restrict__ part, int64_t nchunks, int64_t N, float* __restrict__ out,
                                     int accumulate) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int64_t b = 0; b < nchunks; ++b) s += part[b * N + n];
  out[n] = accumulate ? out[n] + s : s;
}

template <bool A_KC, bool B_KC>
int launch(const GemmParams& p, int splits, cudaStream_t stream) {
  if (p.N > 64) {
    dim3 grid(static_cast<unsigned>(ceil_div(p.N, 128)), static_cast<unsigned>(ceil_div(p.M, BM)), splits);
    sgemm_kernel<128, A_KC, B_KC><<<grid, THREADS, 0, stream>>>(p);
  } else {
    dim3 grid(static_cast<unsigned>(ceil_div(p.N, 64)), static_cast<unsigned>(ceil_div(p.M, BM)), splits);
    sgemm_kernel<64, A_KC, B_KC><<<grid, THREADS, 0, stream>>>(p);
  }
  return check_launch("sgemm");
}

bool vec_ok(const float* p, int64_t ld, int64_t inner) { return aligned16(p) && ld % 4 == 0 && inner >= 0; }

}  // namespace

int simt_linear_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* b, int64_t M, int64_t N,
                    int64_t K, float* y, int64_t ldy, int act, float* y_act, int64_t ldya, cudaStream_t stream) {
  GemmParams p{};
  p.A = x; p.lda = ldx; p.B = w; p.ldb = ldw; p.M = M; p.N = N; p.K = K; p.k_chunk = ceil_div(K > 0 ? K : 1, BK) * BK;
  p.C = y; p.ldc = ldy; p.bias = b; p.act = act; p.C_act = y_act; p.ldca = ldya;
  p.vec_a = vec_ok(x, ldx, K); p.vec_b = vec_ok(w, ldw, K);
  return launch<true, true>(p, 1, stream);
}

int simt_linear_dgrad(const float* dy, int64_t ldy, const float* w, int64_t ldw, int64_t M, int64_t N, int64_t K,
                      float* dx, int64_t ldx, int accumulate, int act, const float* act_pre, int64_t ld_pre,
                      cudaStream_t stream) {
  // dx[M,K] = dy[M,N] * w[N,K]: reduction over N
  GemmParams p{};
  p.A = dy; p.lda = ldy; p.B = w; p.ldb = ldw; p.M = M; p.N = K; p.K = N; p.k_chunk = ceil_div(N > 0 ? N : 1, BK) * BK;
  p.C = dx; p.ldc = ldx; p.accumulate = accumulate; p.act = act; p.act_pre = act_pre; p.ld_pre = ld_pre;
  p.vec_a = vec_ok(dy, ldy, N); p.vec_b = vec_ok(w, ldw, K);
  return launch<true, false>(p, 1, stream);
}

static int wgrad_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ceil_div(N, BM) * ceil_div(K, K > 64 ? 128 : 64);
  int64_t s = (static_cast<int64_t>(sm_count()) * 2 + tiles - 1) / tiles;
  const int64_t max_by_rows = ceil_div(M > 0 ? M : 1, 4 * BK);
  if (s > max_by_rows) s = max_by_rows;
  if (s < 1) s = 1;
  if (s > 512) s = 512;
  return static_cast<int>(s);
}
static int64_t colsum_chunks(int64_t M) {
  int64_t c = ceil_div(M > 0 ? M : 1, 1024);
  if (c > 1024) c = 1024;
  return c;
}

size_t linear_colsum_workspace_bytes(int64_t M, int64_t N) {
  return align_up(static_cast<size_t>(colsum_chunks(M)) * (N > 0 ? N : 1) * sizeof(float));
}

// db[n] (+)= sum_m dy[m, n]; two deterministic stages
int linear_colsum(const float* dy, int64_t ldy, int64_t M, int64_t N, float* db, int accumulate, void* ws,
                  cudaStream_t stream) {
  float* cs = static_cast<float*>(ws);
  const int64_t chunks = colsum_chunks(M);
  const int64_t rpb = ceil_div(M > 0 ? M : 1, chunks);
  dim3 grid(static_cast<unsigned>(ceil_div(N, 64)), static_cast<unsigned>(chunks));
  colsum_stage1_kernel<<<grid, 256, 0, stream>>>(dy, ldy, M, N, rpb, cs);
  colsum_stage2_kernel<<<static_cast<unsigned>(ceil_div(N, 128)), 128, 0, stream>>>(cs, chunks, N, db, accumulate);
  return check_launch("colsum");
}

size_t simt_linear_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  return align_up(static_cast<size_t>(wgrad_splits(M, N, K)) * N * (K > 0 ? K : 1) * sizeof(float));
}

int simt_linear_wgrad(const float* dy, int64_t ldy, const float* x, int64_t ldx, int64_t M, int64_t N, int64_t K,
                      float* dw, int64_t lddw, int accumulate, void* ws, cudaStream_t stream) {
  // dw[N,K] = dy^T[N,M] * x[M,K]: reduction over M
  if (K == 0) return SGB_OK;
  const int splits = wgrad_splits(M, N, K);
  float* part = static_cast<float*>(ws);
  GemmParams p{};
  p.A = dy; p.lda = ldy; p.B = x; p.ldb = ldx; p.M = N; p.N = K; p.K = M;
  p.k_chunk = ceil_div(ceil_div(M > 0 ? M : 1, splits), BK) * BK;
  p.vec_a = vec_ok(dy, ldy, N); p.vec_b = vec_ok(x, ldx, K);
  if (splits == 1) {
    p.C = dw; p.ldc = lddw; p.accumulate = accumulate;
    return launch<false, false>(p, 1, stream);
  }
  p.C = part; p.ldc = K;
  int rc = launch<false, false>(p, splits, stream);
  if (rc != SGB_OK) return rc;
  const int64_t MN = N * K;
  splitk_reduce_kernel<<<static_cast<unsigned>(ceil_div(MN, 256)), 256, 0, stream>>>(part, splits, MN, K, dw, lddw, accumulate);
  return check_launch("splitk_reduce");
}

}  // namespace sgb
