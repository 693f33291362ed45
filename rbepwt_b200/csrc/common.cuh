// Shared helpers for the rbepwt_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FULL_MASK 0xffffffffu

namespace rbepwt {

// Offset of level `lev` (1-based) inside a per-image buffer that concatenates the level signals:
// sum_{l<lev} n >> (l-1)  =  2n - (n >> (lev-2))   (n a power of two, n >> (lev-1) >= 1).
__host__ __device__ __forceinline__ size_t level_off(size_t n, int lev) {
  return lev <= 1 ? 0 : 2 * n - (n >> (lev - 2));
}

// Bit intrinsics callable from host code too: the per-lane logic of the path walker (walk.cuh) is __host__ __device__
// so that tests can run it on the CPU against the oracle (tests/host_walk).
__host__ __device__ __forceinline__ int rb_clz(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}
__host__ __device__ __forceinline__ int rb_ffs(uint32_t x) {  // 1-based position of the lowest set bit, 0 if none
#ifdef __CUDA_ARCH__
  return __ffs((int)x);
#else
  return __builtin_ffs((int)x);
#endif
}
__host__ __device__ __forceinline__ int rb_popc(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}
__host__ __device__ __forceinline__ uint32_t rb_funnel_r(uint32_t lo, uint32_t hi, int sh) {  // (hi:lo) >> (sh & 31)
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, sh);
#else
  return (uint32_t)(((((unsigned long long)hi) << 32) | lo) >> (sh & 31));
#endif
}
__host__ __device__ __forceinline__ double rb_dmul(double a, double b) {  // never contracted into an FMA
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double p = a * b;
  return p;
#endif
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Block-wide exclusive scan of one int per thread (blockDim.x multiple of 32, <= 1024).
// `scratch` is 33 ints of shared memory.  Returns the exclusive prefix; *total gets the block sum.
// Contains two __syncthreads(); must be called by every thread of the block.
__device__ __forceinline__ int block_exclusive_scan(int v, int *scratch, int *total) {
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int y = __shfl_up_sync(FULL_MASK, inc, d);
    if (lane >= (unsigned)d) inc += y;
  }
  __syncthreads();  // protect scratch from the previous call
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nwarps ? scratch[lane] : 0;
    int winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int y = __shfl_up_sync(FULL_MASK, winc, d);
      if (lane >= (unsigned)d) winc += y;
    }
    scratch[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) scratch[32] = winc;
  }
  __syncthreads();
  *total = scratch[32];
  return scratch[warp] + inc - v;
}

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, T *scratch, Op op, T identity) {
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = op(v, __shfl_xor_sync(FULL_MASK, v, d));
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T w = lane < nwarps ? scratch[lane] : identity;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) w = op(w, __shfl_xor_sync(FULL_MASK, w, d));
    if (lane == 0) scratch[0] = w;
  }
  __syncthreads();
  return scratch[0];
}

struct OpMin { template <typename T> __device__ T operator()(T a, T b) const { return a < b ? a : b; } };
struct OpMax { template <typename T> __device__ T operator()(T a, T b) const { return a > b ? a : b; } };
struct OpSum { template <typename T> __device__ T operator()(T a, T b) const { return a + b; } };

// Monotone map double -> uint64 (a < b  <=>  key(a) < key(b)); -0.0 must be canonicalised by the caller.
__device__ __forceinline__ unsigned long long orderable(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

}  // namespace rbepwt
