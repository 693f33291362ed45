// K4: global top-k by |coefficient| per image (radix select on the fp64 magnitude bits) + zeroing,
// K6: PSNR, and the non-zero count.
//
// Replaces Rbepwt.threshold_coefs (/root/reference/rbepwt.py:2081-2112): the reference argsorts
// |flat| (details[1] | ... | details[L] | approx) and copies the `ncoefs` largest into a zero array,
// then writes back in place.  Quirks kept: ncoefs <= 0 or >= N keeps everything (the
// `counter == ncoefs` test never fires).  Ties at the k-th magnitude are broken by numpy's unstable
// argsort in the reference (unpinned); here the highest flat index survives.
// psnr: rbepwt.py:156-162.  nonzero count: Image.nonzero_rbepwt_coefs, rbepwt.py:427-432.
#pragma once
#include "common.cuh"

namespace rbepwt {

constexpr int SEL_THREADS = 1024;
constexpr int SEL_MAXBINS = 8192;

__global__ void __launch_bounds__(SEL_THREADS) k4_threshold(double *coefs_all, int N, long long k) {
  __shared__ int s_hist[SEL_MAXBINS];
  __shared__ int s_scan[33];
  __shared__ int s_digit, s_above, s_ceq, s_seen;
  if (k <= 0 || k >= (long long)N) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  unsigned long long *c = reinterpret_cast<unsigned long long *>(coefs_all + (size_t)blockIdx.x * N);
  const unsigned long long MAG = 0x7fffffffffffffffull;
  const int nbits[5] = {13, 13, 13, 12, 12};  // 63 magnitude bits, most significant first
  unsigned long long prefix = 0;
  int done_bits = 0;
  int krem = (int)k;
  for (int pass = 0; pass < 5; pass++) {
    const int nb = nbits[pass], nbins = 1 << nb, shift = 63 - done_bits - nb;
    for (int i = tid; i < nbins; i += nt) s_hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < N; i += nt) {
      const unsigned long long key = c[i] & MAG;
      if (pass == 0 || (key >> (shift + nb)) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & (nbins - 1))], 1);
    }
    __syncthreads();
    // thread `tid` owns the bins [nbins - (tid+1)*per, nbins - tid*per): thread 0 the largest digits
    const int per = nbins / SEL_THREADS;  // 8 or 4
    int local = 0;
    const int hi = nbins - tid * per;
    for (int d = hi - per; d < hi; d++) local += s_hist[d];
    int total;
    int above = block_exclusive_scan(local, s_scan, &total);  // elements in strictly larger digits
    for (int d = hi - 1; d >= hi - per; d--) {
      const int hcount = s_hist[d];
      if (above < krem && above + hcount >= krem) { s_digit = d; s_above = above; s_ceq = hcount; }
      above += hcount;
    }
    __syncthreads();
    prefix = (prefix << nb) | (unsigned long long)s_digit;
    krem -= s_above;
    done_bits += nb;
    __syncthreads();
  }
  const unsigned long long thr = prefix;  // magnitude bits of the k-th largest
  const int ceq = s_ceq;                  // how many coefficients have exactly that magnitude
  if (krem == ceq) {                      // every tie survives (always, for continuous data)
    for (int i = tid; i < N; i += nt)
      if ((c[i] & MAG) < thr) c[i] = 0ull;
    return;
  }
  // keep only the `krem` ties with the highest flat index
  if (tid == 0) s_seen = 0;
  __syncthreads();
  for (int base = ((N - 1) / nt) * nt; base >= 0; base -= nt) {
    const int i = base + (nt - 1 - tid);  // tid order = descending index
    const bool valid = i < N;
    const unsigned long long key = valid ? (c[i] & MAG) : 0ull;
    const bool tie = valid && key == thr;
    int total;
    const int ex = block_exclusive_scan(tie ? 1 : 0, s_scan, &total);
    if (valid && (key < thr || (tie && s_seen + ex >= krem))) c[i] = 0ull;
    __syncthreads();
    if (tid == 0) s_seen += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_nonzero(const double *coefs_all, int N, long long *out) {
  __shared__ int s_red[33];
  const double *c = coefs_all + (size_t)blockIdx.x * N;
  int cnt = 0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) cnt += c[i] != 0.0;
  cnt = block_reduce(cnt, s_red, OpSum(), 0);
  if (threadIdx.x == 0) out[blockIdx.x] = cnt;
}

// out[b] = 20 log10(255 / sqrt(sum((a-b)^2) / n)), -1 when the sum is exactly 0.
__global__ void __launch_bounds__(1024) k6_psnr(const double *a_all, const double *b_all, long long n, double *out) {
  __shared__ double s_red[33];
  const double *a = a_all + (size_t)blockIdx.x * n, *b = b_all + (size_t)blockIdx.x * n;
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = a[i] - b[i];
    acc += d * d;
  }
  acc = block_reduce(acc, s_red, OpSum(), 0.0);
  if (threadIdx.x == 0) out[blockIdx.x] = acc == 0.0 ? -1.0 : 20.0 * log10(255.0 / sqrt(acc / (double)n));
}

}  // namespace rbepwt
