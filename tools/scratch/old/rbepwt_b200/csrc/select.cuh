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
#include <cooperative_groups.h>

#include "common.cuh"

namespace rbepwt {

namespace cg = cooperative_groups;

constexpr int SEL_THREADS = 1024;
constexpr int SEL_MAXBINS = 8192;
constexpr int SEL_CLUSTER = 8;  // CTAs per image: one thread-block cluster, histograms combined through DSMEM

// One CLUSTER of SEL_CLUSTER CTAs per image.  Each CTA owns a contiguous slice of the coefficients and, in the
// reduction, a contiguous slice of the digit bins.  Per radix pass: local histogram of the slice -> cluster
// sync -> every CTA sums its bin slice over the cluster's histograms (distributed shared memory) -> cluster
// sync -> all CTAs locate the digit of the k-th largest from the 8 slice sums and the owning CTA's totals.
__global__ void __cluster_dims__(SEL_CLUSTER, 1, 1) __launch_bounds__(SEL_THREADS)
    k4_threshold(double *coefs_all, int N, long long k) {
  __shared__ int s_hist[SEL_MAXBINS];
  __shared__ int s_tot[SEL_MAXBINS / SEL_CLUSTER];
  __shared__ int s_slice, s_ties;
  __shared__ int s_scan[33];
  __shared__ int s_digit, s_above, s_ceq, s_seen;
  if (k <= 0 || k >= (long long)N) return;  // uniform over the grid: no cluster barrier is skipped by a subset
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, nt = blockDim.x;
  unsigned long long *c = reinterpret_cast<unsigned long long *>(coefs_all + (size_t)(blockIdx.x / SEL_CLUSTER) * N);
  const int per_cta = (N + SEL_CLUSTER - 1) / SEL_CLUSTER;
  const int lo = min(N, rank * per_cta), hi = min(N, lo + per_cta);
  const unsigned long long MAG = 0x7fffffffffffffffull;
  const int nbits[5] = {13, 13, 13, 12, 12};  // 63 magnitude bits, most significant first
  unsigned long long prefix = 0;
  int done_bits = 0;
  int krem = (int)k;
  for (int pass = 0; pass < 5; pass++) {
    const int nb = nbits[pass], nbins = 1 << nb, shift = 63 - done_bits - nb;
    const int bpc = nbins / SEL_CLUSTER;  // bins per CTA in the reduction: 1024 or 512
    for (int i = tid; i < nbins; i += nt) s_hist[i] = 0;
    __syncthreads();
    for (int i = lo + tid; i < hi; i += nt) {
      const unsigned long long key = c[i] & MAG;
      if (pass == 0 || (key >> (shift + nb)) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & (nbins - 1))], 1);
    }
    cluster.sync();
    int part = 0;
    for (int b = tid; b < bpc; b += nt) {
      int sum = 0;
#pragma unroll
      for (int r = 0; r < SEL_CLUSTER; r++) sum += cluster.map_shared_rank(s_hist, r)[rank * bpc + b];
      s_tot[b] = sum;
      part += sum;
    }
    part = block_reduce(part, s_scan, OpSum(), 0);
    if (tid == 0) s_slice = part;
    cluster.sync();
    // larger digits live in higher ranks: walk the slice sums from the top
    int above = 0, owner = 0;
    for (int r = SEL_CLUSTER - 1; r >= 0; r--) {
      const int sl = *cluster.map_shared_rank(&s_slice, r);
      if (above + sl >= krem) { owner = r; break; }
      above += sl;
    }
    // thread t looks at the owner's bin bpc-1-t (descending digits)
    const int *otot = cluster.map_shared_rank(s_tot, owner);
    const int local = tid < bpc ? otot[bpc - 1 - tid] : 0;
    int total;
    const int ex = block_exclusive_scan(local, s_scan, &total);
    if (tid < bpc && above + ex < krem && above + ex + local >= krem) {
      s_digit = owner * bpc + (bpc - 1 - tid); s_above = above + ex; s_ceq = local;
    }
    __syncthreads();
    prefix = (prefix << nb) | (unsigned long long)s_digit;
    krem -= s_above;
    done_bits += nb;
    cluster.sync();  // nobody still reads this CTA's s_hist / s_tot / s_slice when the next pass clears them
  }
  const unsigned long long thr = prefix;  // magnitude bits of the k-th largest
  const int ceq = s_ceq;                  // how many coefficients have exactly that magnitude
  if (krem == ceq) {                      // every tie survives (always, for continuous data)
    for (int i = lo + tid; i < hi; i += nt)
      if ((c[i] & MAG) < thr) c[i] = 0ull;
    return;
  }
  // keep only the `krem` ties with the highest flat index: slices of higher rank come first
  int ties = 0;
  for (int i = lo + tid; i < hi; i += nt) ties += (c[i] & MAG) == thr;
  ties = block_reduce(ties, s_scan, OpSum(), 0);
  if (tid == 0) s_ties = ties;
  cluster.sync();
  if (tid == 0) {
    int seen = 0;
    for (int r = rank + 1; r < SEL_CLUSTER; r++) seen += *cluster.map_shared_rank(&s_ties, r);
    s_seen = seen;
  }
  __syncthreads();
  const int len = hi - lo;
  for (int base = len > 0 ? ((len - 1) / nt) * nt : -1; base >= 0; base -= nt) {
    const int off = base + (nt - 1 - tid);  // tid order = descending index
    const bool valid = off < len;
    const int i = lo + off;
    const unsigned long long key = valid ? (c[i] & MAG) : 0ull;
    const bool tie = valid && key == thr;
    int total;
    const int ex = block_exclusive_scan(tie ? 1 : 0, s_scan, &total);
    if (valid && (key < thr || (tie && s_seen + ex >= krem))) c[i] = 0ull;
    __syncthreads();
    if (tid == 0) s_seen += total;
    __syncthreads();
  }
  cluster.sync();  // s_ties stays readable until every CTA has summed it
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
