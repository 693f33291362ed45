// K3 / K5: one level of the periodized 1-D DWT / IDWT along the concatenated paths.
//
// Replaces, per level, the concatenation of the regions' path-ordered values
// (RegionCollection.add_region, /root/reference/rbepwt.py:1550, 2036), the ONE global
// pywt.dwt(values, wavelet, 'periodization') of the level (rbepwt.py:2041) and the value part of
// RegionCollection.reduce (1578); inverse: pywt.idwt (2067) + RegionCollection.expand (1586-1613)
// + the final scatter and clip of Image.decode_rbepwt (307-317).
//
// Layout.  x^l (length n_l = N >> (l-1)) is the level's input in its INCOMING order: x^1 = the pixels,
// x^(l+1) = cA of level l (reduce keeps the even positions of the path-ordered signal, so cA[o] is the value
// of the point at path position 2o and the next level's incoming order is exactly the order of cA).
// The level signal is s[t] = x^l[P_l[t]] where P_l is what the path kernel wrote: pixel ids at level 1
// (Q), positions in the incoming order at levels >= 2 (Pm = the reference's generating permutation, region
// offset included).  Forward: gather through P_l, dense coalesced cA / cD out.  Inverse: dense coalesced in,
// x^l[P_l[t]] = s[t] scattered back (RegionCollection.expand's argsort, rbepwt.py:1600-1608, is this scatter).
// The planes holding x^l are dense: plane[l & 1] holds x^(l+1), n_l / 2 doubles per image.
//
// Arithmetic (PyWavelets' periodization mode, restated -- see oracle/pywt_port.py):
//   cA[o] = sum_{j=0}^{F-1} dec_lo[j] * s[(2o + F/2 - j) mod n]   ascending j, one multiply and one
//   add per tap (explicit _rn intrinsics, never contracted into FMA), fp64;
//   s[t]  = (sum_m rec_lo[m] cA[o]) + (sum_m rec_hi[m] cD[o]),  m ascending over the taps with
//           t + F/2 - 1 - m even, o = (t + F/2 - 1 - m)/2 mod n/2.
#pragma once
#include "common.cuh"

namespace rbepwt {

constexpr int FMAX = 128;       // longest supported filter
constexpr int DWT_THREADS = 256;
constexpr int FWD_TILE = 512;   // low-pass outputs per CTA
constexpr int INV_TILE = 1024;  // reconstructed samples per CTA

constexpr int FT_MAX = 10;  // filter lengths up to this have kernels with the taps unrolled (bior4.4 = 10)
#ifndef RB_TAIL_MAX_POINTS
#define RB_TAIL_MAX_POINTS 2048
#endif
constexpr int TAIL_MAX_POINTS = RB_TAIL_MAX_POINTS;  // levels with at most this many points per image run in the tail kernels

struct DwtParams {
  const double *vin;   // level 1 of the forward transform: the images, image stride vin_stride
  size_t vin_stride;
  double *plane[2];    // dense planes, image stride N/2: plane[l & 1] holds x^(l+1) (= cA of level l)
  const int32_t *Q;    // [chunk][2N] paths as pixel ids (level 1 is used here)
  const int32_t *Pm;   // [chunk][2N] paths as positions in the incoming order (levels >= 2)
  double *coefs;       // [chunk][N] flat coefficients: details[1] | ... | details[L] | approx
  const double *filt;  // dec_lo[FMAX] dec_hi[FMAX] rec_lo[FMAX] rec_hi[FMAX]
  double tap_lo[FT_MAX], tap_hi[FT_MAX];  // the direction's two filters by value (constant bank) when flen <= FT_MAX
  double *out_img;     // decode, level 1: clipped image
  int flen, N, lev, levels;
  int clip;            // decode, level 1: clip to [0,255] (Image.decode_rbepwt); 0 = Rbepwt.decode's own values
  const unsigned long long *thr;  // decode: pending thresholds, one record of 2 words per image ({tau, flags}: select.cuh
                                  // ThrRec) or null -- coefficients with magnitude bits below tau are read as zero
};

struct FwdSmem {
  double e[FWD_TILE + FMAX / 2 + 2], o[FWD_TILE + FMAX / 2 + 2];
  double lo[FMAX], hi[FMAX];
};

// One tile (FWD_TILE low-pass outputs) of level `lev` of image `img`.  The caller loads sm.lo / sm.hi once.
// SAME_CTA: the input plane was written by this CTA (tail kernel) -> read it past L1.
// FT > 0: compile-time filter length, taps read from the kernel parameters; FT = 0: any length, taps in shared memory.
template <bool SAME_CTA, int FT>
__device__ __forceinline__ void dwt_tile(const DwtParams &P, FwdSmem &sm, int lev, int tile, size_t img) {
  const int tid = threadIdx.x, nt = blockDim.x, F = FT ? FT : P.flen, N = P.N;
  const int n = N >> (lev - 1), half = n >> 1, mask = n - 1;
  const int32_t *Pl = (lev == 1 ? P.Q : P.Pm) + img * 2 * (size_t)N + level_off((size_t)N, lev);
  // level 1 reads the image; level l >= 2 reads the plane level l-1 wrote (ping-pong on the level's parity)
  const double *vin = lev == 1 ? P.vin + img * P.vin_stride : P.plane[(lev - 1) & 1] + img * (size_t)(N >> 1);
  double *vout = P.plane[lev & 1] + img * (size_t)(N >> 1);
  double *coefs = P.coefs + img * (size_t)N;
  const int o0 = tile * FWD_TILE, nout = min(FWD_TILE, half - o0);
  const int tstart = 2 * o0 - F / 2 + 1, cnt = 2 * (nout - 1) + F;
  // gather, four independent index -> value chains in flight per thread
  for (int i0 = tid; i0 < cnt; i0 += 4 * nt) {
    int src[4];
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) src[u] = i0 + u * nt < cnt ? Pl[(tstart + i0 + u * nt) & mask] : 0;
#pragma unroll
    for (int u = 0; u < 4; u++) v[u] = i0 + u * nt < cnt ? (SAME_CTA ? __ldcg(vin + src[u]) : vin[src[u]]) : 0.0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + u * nt;
      if (i < cnt) { if (i & 1) sm.o[i >> 1] = v[u]; else sm.e[i >> 1] = v[u]; }
    }
  }
  __syncthreads();
  const bool last = lev == P.levels;
  const size_t det_off = (size_t)N - (size_t)n;            // sum_{l<lev} N >> l
  const size_t app_off = (size_t)N - (size_t)(N >> P.levels);
  for (int ol = tid; ol < nout; ol += nt) {
    double a = 0.0, d = 0.0;
    // local sample index of tap j: 2*ol + F-1-j  (odd for even j)
    if (FT) {
#pragma unroll
      for (int j = 0; j < FT; j += 2) {
        const int q = ol + ((FT - 2 - j) >> 1);
        const double x1 = sm.o[q], x2 = sm.e[q];
        a = __dadd_rn(a, __dmul_rn(P.tap_lo[j], x1));
        d = __dadd_rn(d, __dmul_rn(P.tap_hi[j], x1));
        a = __dadd_rn(a, __dmul_rn(P.tap_lo[j + 1], x2));
        d = __dadd_rn(d, __dmul_rn(P.tap_hi[j + 1], x2));
      }
    } else {
      for (int j = 0; j < F; j += 2) {
        const int q = ol + ((F - 2 - j) >> 1);
        const double x1 = sm.o[q], x2 = sm.e[q];
        a = __dadd_rn(a, __dmul_rn(sm.lo[j], x1));
        d = __dadd_rn(d, __dmul_rn(sm.hi[j], x1));
        a = __dadd_rn(a, __dmul_rn(sm.lo[j + 1], x2));
        d = __dadd_rn(d, __dmul_rn(sm.hi[j + 1], x2));
      }
    }
    coefs[det_off + o0 + ol] = d;
    if (last) coefs[app_off + o0 + ol] = a;
    else vout[o0 + ol] = a;
  }
}

// One level, one tile per CTA: the levels with many tiles per image.
template <int FT>
__global__ void __launch_bounds__(DWT_THREADS) k3_dwt_level(DwtParams P) {
  __shared__ FwdSmem sm;
  if (!FT) for (int i = threadIdx.x; i < P.flen; i += blockDim.x) { sm.lo[i] = P.filt[i]; sm.hi[i] = P.filt[FMAX + i]; }
  dwt_tile<false, FT>(P, sm, P.lev, blockIdx.x, blockIdx.y);
}

// Levels P.lev .. P.levels of one image in ONE CTA: the deep levels are a chain of tiny dependent passes
// (a 512^2 image has 2048 + 1024 + ... + 4 points from level 8 on) -- one launch instead of nine.
__global__ void __launch_bounds__(DWT_THREADS) k3_dwt_tail(DwtParams P) {
  __shared__ FwdSmem sm;
  for (int i = threadIdx.x; i < P.flen; i += blockDim.x) { sm.lo[i] = P.filt[i]; sm.hi[i] = P.filt[FMAX + i]; }
  for (int lev = P.lev; lev <= P.levels; lev++) {
    const int half = (P.N >> (lev - 1)) >> 1;
    for (int tile = 0; tile * FWD_TILE < half; tile++) {
      __syncthreads();  // shared tile reuse; also publishes the previous level's plane writes to the CTA
      dwt_tile<true, 0>(P, sm, lev, tile, blockIdx.x);
    }
  }
}

struct InvSmem {
  double a[INV_TILE / 2 + FMAX / 2 + 2], d[INV_TILE / 2 + FMAX / 2 + 2];
  double lo[FMAX], hi[FMAX];
};

// One tile (INV_TILE reconstructed samples) of level `lev` of image `img`.
template <bool SAME_CTA, int FT>
__device__ __forceinline__ void idwt_tile(const DwtParams &P, InvSmem &sm, int lev, int tile, size_t img) {
  const int tid = threadIdx.x, nt = blockDim.x, F = FT ? FT : P.flen, N = P.N;
  const int n = N >> (lev - 1), half = n >> 1, hmask = half - 1;
  const int32_t *Pl = (lev == 1 ? P.Q : P.Pm) + img * 2 * (size_t)N + level_off((size_t)N, lev);
  const double *vin = P.plane[lev & 1] + img * (size_t)(N >> 1);    // x^(lev+1), reconstructed by level lev+1
  double *vout = P.plane[(lev - 1) & 1] + img * (size_t)(N >> 1);  // x^lev
  const double *coefs = P.coefs + img * (size_t)N;
  const int t0 = tile * INV_TILE, nout = min(INV_TILE, n - t0);
  const int omin = (t0 - F / 2) >> 1;  // floor
  const int omax = (t0 + nout - 1 + F / 2 - 1) >> 1;
  const int cnt = omax - omin + 1;
  const bool deepest = lev == P.levels;
  const size_t det_off = (size_t)N - (size_t)n;
  const size_t app_off = (size_t)N - (size_t)(N >> P.levels);
  // a pending threshold (k4_select) is applied here, on load: tau = 0 keeps everything
  const unsigned long long tau = (P.thr && (int)P.thr[2 * img + 1]) ? P.thr[2 * img] : 0ull;
  auto kept = [&](double v) {
    return ((unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull) >= tau ? v : 0.0;
  };
  for (int i = tid; i < cnt; i += nt) {
    const int ow = (omin + i) & hmask;
    sm.a[i] = deepest ? kept(coefs[app_off + ow]) : (SAME_CTA ? __ldcg(vin + ow) : vin[ow]);
    sm.d[i] = kept(coefs[det_off + ow]);
  }
  __syncthreads();
  auto emit = [&](int t, double x) {
    const int dst = Pl[t];
    if (lev == 1) {  // Image.decode_rbepwt: clip, no rounding (rbepwt.py:312-314)
      if (P.clip) x = x > 255.0 ? 255.0 : (x < 0.0 ? 0.0 : x);
      P.out_img[img * (size_t)N + dst] = x;
    } else {
      vout[dst] = x;
    }
  };
  if (FT) {
    // a thread reconstructs the pair (t, t+1), t even: the two outputs use the taps of opposite parity on
    // (nearly) the same approximation / detail samples, all tap indices are compile-time constants
    constexpr int hh = (FT ? FT : 2) / 2 - 1, par0 = hh & 1, par1 = par0 ^ 1;
    for (int pl = tid; 2 * pl < nout; pl += nt) {
      const int t = t0 + 2 * pl, base = t + hh;
      double slo0 = 0.0, shi0 = 0.0, slo1 = 0.0, shi1 = 0.0;
#pragma unroll
      for (int mm = 0; mm < FT; mm += 2) {
        const int m0 = mm + par0, m1 = mm + par1;
        const int oi0 = ((base - m0) >> 1) - omin, oi1 = ((base + 1 - m1) >> 1) - omin;
        slo0 = __dadd_rn(slo0, __dmul_rn(P.tap_lo[m0], sm.a[oi0]));
        shi0 = __dadd_rn(shi0, __dmul_rn(P.tap_hi[m0], sm.d[oi0]));
        slo1 = __dadd_rn(slo1, __dmul_rn(P.tap_lo[m1], sm.a[oi1]));
        shi1 = __dadd_rn(shi1, __dmul_rn(P.tap_hi[m1], sm.d[oi1]));
      }
      emit(t, __dadd_rn(slo0, shi0));
      emit(t + 1, __dadd_rn(slo1, shi1));
    }
  } else {
    for (int tl = tid; tl < nout; tl += nt) {
      const int t = t0 + tl, base = t + F / 2 - 1;
      double slo = 0.0, shi = 0.0;
      for (int m = base & 1; m < F; m += 2) {
        const int oi = ((base - m) >> 1) - omin;
        slo = __dadd_rn(slo, __dmul_rn(sm.lo[m], sm.a[oi]));
        shi = __dadd_rn(shi, __dmul_rn(sm.hi[m], sm.d[oi]));
      }
      emit(t, __dadd_rn(slo, shi));
    }
  }
}

template <int FT>
__global__ void __launch_bounds__(DWT_THREADS) k5_idwt_level(DwtParams P) {
  __shared__ InvSmem sm;
  if (!FT) for (int i = threadIdx.x; i < P.flen; i += blockDim.x) { sm.lo[i] = P.filt[2 * FMAX + i]; sm.hi[i] = P.filt[3 * FMAX + i]; }
  idwt_tile<false, FT>(P, sm, P.lev, blockIdx.x, blockIdx.y);
}

// Levels P.levels down to P.lev of one image in ONE CTA (the deep levels, see k3_dwt_tail).
__global__ void __launch_bounds__(DWT_THREADS) k5_idwt_tail(DwtParams P) {
  __shared__ InvSmem sm;
  for (int i = threadIdx.x; i < P.flen; i += blockDim.x) { sm.lo[i] = P.filt[2 * FMAX + i]; sm.hi[i] = P.filt[3 * FMAX + i]; }
  for (int lev = P.levels; lev >= P.lev; lev--) {
    const int n = P.N >> (lev - 1);
    for (int tile = 0; tile * INV_TILE < n; tile++) {
      __syncthreads();
      idwt_tile<true, 0>(P, sm, lev, tile, blockIdx.x);
    }
  }
}

// EPWT only: the path kernel of level l+1 compares VALUES by pixel, so the dense cA of level l is also laid
// out by pixel: vpix[Q_l[2o]] = cA[o].
__global__ void k_plane_to_pixels(const double *__restrict__ plane, const int32_t *__restrict__ Q, int N, int lev,
                                  double *vpix) {
  const size_t img = blockIdx.y;
  const int half = (N >> (lev - 1)) >> 1;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= half) return;
  const int32_t *Ql = Q + img * 2 * (size_t)N + level_off((size_t)N, lev);
  vpix[img * (size_t)N + Ql[2 * o]] = plane[img * (size_t)(N >> 1) + o];
}

}  // namespace rbepwt
