// The tensor-product baseline the reference compares against: class Dwt (/root/reference/rbepwt.py:2249-2298),
// pywt.wavedec2 / waverec2 with mode='periodization' -- a separable multi-level 2-D DWT, no paths involved.
//
// Layout: the coefficients of an image live in ONE H x W array in the usual pyramid arrangement.  Level l works on
// the top-left s x s block (s = side >> (l-1)): the transform along axis 0 leaves its low-pass half in the rows
// [0, s/2) and its high-pass half in [s/2, s), the transform along axis 1 does the same with the columns, so the block
// becomes  [aa | ad ; da | dd]  (pywt: cA = aa, cH = da, cV = ad, cD = dd) and the next level transforms aa.  Top-k
// thresholding (K4) runs on the array as it is: which coefficients survive does not depend on how they are laid out.
// Arithmetic: the 1-D periodized transform of dwt.cuh / oracle/pywt_port.py along each line, axis 0 first, then
// axis 1 (pywt.dwtn visits the axes in order); the inverse undoes axis 1, then axis 0.
#pragma once
#include "common.cuh"
#include "dwt.cuh"

namespace rbepwt {

struct Dwt2Params {
  const double *src;  // [B][H][W]
  double *dst;        // [B][H][W]
  const double *filt; // dec_lo[FMAX] dec_hi[FMAX] rec_lo[FMAX] rec_hi[FMAX]
  int W, N, s, flen;  // row stride, pixels per image, side of the block this level works on
  int clip;           // inverse, last pass: clip to [0,255] (Image.decode_dwt, rbepwt.py:326-328)
};

// forward along AXIS (0: down the columns, 1: along the rows) of the s x s block; thread = one (cA, cD) pair.
// The fastest thread index runs along the rows of the image in both cases, so accesses coalesce.
template <int AXIS>
__global__ void __launch_bounds__(256) k_dwt2_fwd(Dwt2Params P) {
  const int half = P.s >> 1, F = P.flen;
  const size_t img = blockIdx.z;
  const double *src = P.src + img * P.N;
  double *dst = P.dst + img * P.N;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  // AXIS 0: x = column c in [0, s), y = output o in [0, s/2);  AXIS 1: x = output o in [0, s/2), y = row r in [0, s)
  if (x >= (AXIS == 0 ? P.s : half)) return;
  const int o = AXIS == 0 ? y : x, line = AXIS == 0 ? x : y;
  double a = 0.0, d = 0.0;
  for (int j = 0; j < F; j++) {
    const int t = (2 * o + F / 2 - j) & (P.s - 1);
    const double v = AXIS == 0 ? src[(size_t)t * P.W + line] : src[(size_t)line * P.W + t];
    a = __dadd_rn(a, __dmul_rn(P.filt[j], v));
    d = __dadd_rn(d, __dmul_rn(P.filt[FMAX + j], v));
  }
  if (AXIS == 0) { dst[(size_t)o * P.W + line] = a; dst[(size_t)(half + o) * P.W + line] = d; }
  else { dst[(size_t)line * P.W + o] = a; dst[(size_t)line * P.W + half + o] = d; }
}

// inverse along AXIS of the s x s block; thread = one reconstructed sample
template <int AXIS>
__global__ void __launch_bounds__(256) k_dwt2_inv(Dwt2Params P) {
  const int half = P.s >> 1, F = P.flen;
  const size_t img = blockIdx.z;
  const double *src = P.src + img * P.N;
  double *dst = P.dst + img * P.N;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= P.s) return;
  const int t = AXIS == 0 ? y : x, line = AXIS == 0 ? x : y;  // AXIS 0: x = column, y = sample;  AXIS 1: x = sample, y = row
  const int base = t + F / 2 - 1;
  double slo = 0.0, shi = 0.0;
  for (int m = base & 1; m < F; m += 2) {
    const int o = ((base - m) >> 1) & (half - 1);
    const double ca = AXIS == 0 ? src[(size_t)o * P.W + line] : src[(size_t)line * P.W + o];
    const double cd = AXIS == 0 ? src[(size_t)(half + o) * P.W + line] : src[(size_t)line * P.W + half + o];
    slo = __dadd_rn(slo, __dmul_rn(P.filt[2 * FMAX + m], ca));
    shi = __dadd_rn(shi, __dmul_rn(P.filt[3 * FMAX + m], cd));
  }
  double v = __dadd_rn(slo, shi);
  if (P.clip) v = v > 255.0 ? 255.0 : (v < 0.0 ? 0.0 : v);
  if (AXIS == 0) dst[(size_t)t * P.W + line] = v;
  else dst[(size_t)line * P.W + t] = v;
}

// dst block <- src block (s x s), the rest of dst untouched
__global__ void __launch_bounds__(256) k_dwt2_copy_block(const double *src_all, double *dst_all, int W, int N, int s) {
  const size_t img = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < s) dst_all[img * N + (size_t)y * W + x] = src_all[img * N + (size_t)y * W + x];
}

}  // namespace rbepwt
