// Shared device/host helpers for libsegger_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#define SGB_OK 0
#define SGB_ERR_ARG -1       // invalid argument (null pointer, bad dtype, unsupported shape)
#define SGB_ERR_ALIGN -2     // pointer / leading dimension not aligned for 128-bit access
#define SGB_ERR_RANGE -3     // size exceeds the 31-bit index range of one call
#define SGB_ERR_WORKSPACE -4 // workspace too small
#define SGB_ERR_CUDA -5      // CUDA runtime error at launch

namespace sgb {

// thread-local error message; no exceptions cross the ABI
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline __host__ __device__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// number of SMs of the current device (cached per process)
int sm_count();

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming (evict-first) 128-bit store: outputs that are not re-read by this kernel
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  // d/dx [x * Phi(x)] = Phi(x) + x * phi(x)
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Counter-based dropout RNG: keep(seed, original edge id, head).  32-bit integer hash (murmur3
// finaliser of the edge id mixed with the low seed word, then one multiplicative round per head
// mixed with the high seed word): ~8 integer instructions per edge + 4 per head, regenerated
// bit-identically in the backward.
__host__ __device__ __forceinline__ uint32_t rng_edge(uint64_t seed, uint32_t eid) {
  uint32_t x = eid * 0x9E3779B1u + static_cast<uint32_t>(seed);
  x ^= x >> 16; x *= 0x85EBCA6Bu;
  x ^= x >> 13; x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t rng_head(uint32_t edge_hash, uint64_t seed, int h) {
  uint32_t y = (edge_hash ^ (static_cast<uint32_t>(seed >> 32) + static_cast<uint32_t>(h) * 0x632BE5ABu)) * 0x9E3779B1u;
  y ^= y >> 15; y *= 0x2C1B3C6Du;
  y ^= y >> 12;
  return y;
}
__host__ __device__ __forceinline__ uint32_t rng_u32(uint64_t seed, uint32_t eid, int h) {
  return rng_head(rng_edge(seed, eid), seed, h);
}
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  double t = static_cast<double>(p) * 4294967296.0;
  if (t <= 0.0) return 0u;
  if (t >= 4294967295.0) return 4294967295u;
  return static_cast<uint32_t>(t);
}
// keep iff u32 >= threshold  (P(drop) = p)
__device__ __forceinline__ bool rng_keep(uint64_t seed, int64_t eid, int H, int h, uint32_t thr) {
  (void)H;
  return rng_u32(seed, static_cast<uint32_t>(eid), h) >= thr;
}

}  // namespace sgb

#define SGB_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) return sgb::set_error((code), __VA_ARGS__); \
  } while (0)
