// Element-type conversions at the edges of the pipeline and the compact output of the encoder.
//
// The reference's Image.read yields uint8 for grayscale files (/root/reference/rbepwt.py:200-206) and every array is
// cast to float64 before the transform (pywt.dwt); label maps of a 512x512 image hold ~10^3 values.  Shipping
// float64 pixels and int32 labels over PCIe (12 bytes per pixel) is what bounds the end-to-end rate, so the C ABI
// (rbepwt_transcode_ex) also takes uint8 / float32 pixels and uint16 labels, copies those, and widens them here --
// exactly: every uint8 and float32 value is a float64 value.  Narrow OUTPUT (float32, or uint8 = rint + the clip
// already applied) is a convenience of this library, not something the reference computes (its rint line is
// commented out, rbepwt.py:312).
#pragma once
#include "common.cuh"
#include "select.cuh"

namespace rbepwt {

__global__ void __launch_bounds__(256) k_widen_labels(const uint16_t *__restrict__ src, int32_t *__restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (int32_t)src[i];
}

template <typename T>
__global__ void __launch_bounds__(256) k_widen_pixels(const T *__restrict__ src, double *__restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = (double)src[i];
}

template <typename T>
__global__ void __launch_bounds__(256) k_narrow_pixels(const double *__restrict__ src, T *__restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = src[i];
    if (sizeof(T) == 1) dst[i] = (T)__double2int_rn(fmin(fmax(v, 0.0), 255.0));  // decoded images are clipped already
    else dst[i] = (T)v;
  }
}

// The encoder's compact output: the kept coefficients of one image as (flat index, value) pairs, index ascending,
// `kk` slots per image (unused slots: index -1, value 0).  A coefficient is kept when its magnitude bits are >= the
// image's pending threshold (k4_select), or, with none pending, when it is non-zero (k4_threshold zeroed in place).
// One CTA per image; a thread takes 8 consecutive coefficients of a tile, tiles are ranked by a block scan.
__global__ void __launch_bounds__(1024) k_kept_pairs(const double *coefs_all, int N, const ThrRec *rec_all, long long kk,
                                                      int32_t *idx_all, double *val_all) {
  __shared__ int s_scan[33];
  __shared__ int s_running;
  const int img = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const unsigned long long *c = reinterpret_cast<const unsigned long long *>(coefs_all + (size_t)img * N);
  const unsigned long long MAG = 0x7fffffffffffffffull;
  const ThrRec rec = rec_all[img];
  const unsigned long long tau = rec.active ? rec.tau : 1ull;  // no threshold pending: magnitude bits >= 1 <=> non-zero
  int32_t *idx = idx_all + (size_t)img * kk;
  double *val = val_all + (size_t)img * kk;
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < N; base += 8 * nt) {
    const int i0 = base + 8 * tid;
    unsigned long long v[8];
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      v[u] = i0 + u < N ? c[i0 + u] : 0ull;
      cnt += (v[u] & MAG) >= tau;
    }
    int total;
    int at = s_running + block_exclusive_scan(cnt, s_scan, &total);
#pragma unroll
    for (int u = 0; u < 8; u++)
      if ((v[u] & MAG) >= tau) {
        if (at < kk) { idx[at] = i0 + u; val[at] = __longlong_as_double((long long)v[u]); }
        at++;
      }
    __syncthreads();
    if (tid == 0) s_running += total;
    __syncthreads();
  }
  for (long long j = s_running + tid; j < kk; j += nt) { idx[j] = -1; val[j] = 0.0; }
}

}  // namespace rbepwt
