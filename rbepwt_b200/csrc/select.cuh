// K4: global top-k by |coefficient| per image (radix select on the fp64 magnitude bits): k4_select (one pass, the
// threshold is recorded and applied by the inverse transform) with k4_threshold (exact multi-pass, zeroes in place)
// behind it for the unusual cases,
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
#ifndef SEL_CAND_N
#define SEL_CAND_N 6144
#endif
constexpr int SEL_CAND = SEL_CAND_N;  // candidate keys a CTA keeps in shared memory after the second pass
constexpr size_t SEL_SMEM = (size_t)SEL_CAND * 8;  // dynamic shared memory of k4_threshold

// The pending threshold of an image: the coefficients whose magnitude bits are >= tau survive.  k4_select leaves the
// coefficients in memory untouched and records tau; the inverse transform applies it while it loads the
// coefficients, and k_apply_threshold writes the zeros out when somebody wants to look at the coefficients.
struct ThrRec {
  unsigned long long tau;  // magnitude bits (sign cleared) of the k-th largest coefficient
  int active;              // 0: nothing pending for this image
  int pad;
};

constexpr int SEL_SAMPLE_LOG = 4;   // k4_select samples one 32-byte sector (4 coefficients) out of 16
constexpr int SEL2_BINS = 2048;      // 11-bit digits
constexpr int SEL2_CAND = 8192;      // candidate keys (64-bit) = sample keys (32-bit) x 2: the two share one buffer
constexpr size_t SEL2_SMEM = (size_t)SEL2_CAND * 8;

// r-th largest (r >= 1, r <= n) of n keys in shared memory by radix select, 11 bits per pass from bit `bits` - 1 down;
// *ceq = how many keys equal it, *krem = which of them (in descending order) the r-th is: 1..ceq.  Every thread of
// the block calls it.
template <typename KEY, typename KEYFN>
__device__ __forceinline__ KEY block_select_desc_fn(KEYFN key_at, int n, int r, int bits, int *s_hist, int *s_scan,
                                                    int *s_pick, int *krem, int *ceq) {
  const int tid = threadIdx.x, nt = blockDim.x;
  KEY prefix = 0;
  int done = 0;
  while (done < bits) {
    const int nb = min(11, bits - done), nbins = 1 << nb, shift = bits - done - nb;
    for (int i = tid; i < nbins; i += nt) s_hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
      const KEY key = key_at(i);
      if (done == 0 || (key >> (shift + nb)) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & (KEY)(nbins - 1))], 1);
    }
    __syncthreads();
    // thread t looks at the bins nbins-1-pt*t .. (descending digits), pt bins each: all 2048 bins whatever the block size
    const int pt = (SEL2_BINS + nt - 1) / nt;
    int cnt[8];
    int mine = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int b = nbins - 1 - (pt * tid + u);
      cnt[u] = (u < pt && b >= 0) ? s_hist[b] : 0;
      mine += cnt[u];
    }
    int total;
    int ex = block_exclusive_scan(mine, s_scan, &total);
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (u < pt && ex < r && ex + cnt[u] >= r) { s_pick[0] = nbins - 1 - (pt * tid + u); s_pick[1] = ex; s_pick[2] = cnt[u]; }
      ex += cnt[u];
    }
    __syncthreads();
    prefix = (prefix << nb) | (KEY)s_pick[0];
    r -= s_pick[1];
    *ceq = s_pick[2];
    done += nb;
    __syncthreads();  // s_pick and s_hist are rewritten by the next pass
  }
  *krem = r;
  return prefix;
}

template <typename KEY>
__device__ __forceinline__ KEY block_select_desc(const KEY *keys, int n, int r, int bits, int *s_hist, int *s_scan,
                                                 int *s_pick, int *krem, int *ceq) {
  return block_select_desc_fn<KEY>([keys](int i) { return keys[i]; }, n, r, bits, s_hist, s_scan, s_pick, krem, ceq);
}

// K4, common case: ONE pass over the coefficients, one CTA per image.  A sample (1/16 of the 32-byte sectors, top 32
// magnitude bits of each key) is held in shared memory; its order statistics of rank k/16 -+ a safety margin
// bracket the true k-th largest with overwhelming probability.  The full pass then only counts the keys above the
// bracket and keeps the keys inside it (a few thousand of the image's 2^18) in shared memory, where a radix select
// over all 63 magnitude bits finds the k-th largest exactly.  Checked, not assumed: if the bracket does not contain
// rank k, if the candidates overflow their buffer, or if ties at the k-th magnitude would have to be cut (only some
// of them survive: the highest flat indices), the image is flagged in `need_exact` and k4_threshold -- the exact
// multi-pass kernel below, which zeroes in place -- does it instead.
__global__ void __launch_bounds__(SEL_THREADS) k4_select(const double *coefs_all, int N, long long k, ThrRec *rec_all,
                                                         int *need_exact) {
  __shared__ int s_hist[SEL2_BINS];
  __shared__ int s_scan[33];
  __shared__ int s_pick[3];
  __shared__ int s_ncand;
  extern __shared__ unsigned long long s_cand[];  // SEL2_CAND 64-bit keys; first: up to 2 * SEL2_CAND 32-bit sample keys
  if (k <= 0 || k >= (long long)N) return;  // keeps everything (reference quirk)
  const int tid = threadIdx.x, nt = blockDim.x;
  const int img = blockIdx.x;
  const unsigned long long *c = reinterpret_cast<const unsigned long long *>(coefs_all + (size_t)img * N);
  const unsigned long long MAG = 0x7fffffffffffffffull;
  uint32_t *s_samp = reinterpret_cast<uint32_t *>(s_cand);

  // ---- A: the sample -- sector s (4 keys) of the image is sampled when s % 16 == 0 -- and the bracket [v_lo, v_hi]
  // (top 32 magnitude bits) from its order statistics
  uint32_t v_hi = 0xffffffffu, v_lo = 0u;
  const int nsample = 4 * ((N >> 2) >> SEL_SAMPLE_LOG);
  if (nsample >= 64 && nsample <= 2 * SEL2_CAND && N > SEL2_CAND) {  // (small images: everything is a candidate)
    for (int j = tid; j < nsample; j += nt) {
      const int i = ((j >> 2) << (2 + SEL_SAMPLE_LOG)) | (j & 3);
      s_samp[j] = (uint32_t)((c[i] & MAG) >> 31);
    }
    __syncthreads();
    // sample ranks bracketing k / 16 by ~5 standard deviations of the sampling error
    const double ks = (double)k / (double)(1 << SEL_SAMPLE_LOG);
    const int margin = (int)(5.0 * sqrt(ks + 1.0)) + 8;
    const int r_hi = (int)ks - margin, r_lo = (int)ks + margin + 1;
    int kr, ce;
    if (r_hi >= 1) v_hi = block_select_desc<uint32_t>(s_samp, nsample, r_hi, 32, s_hist, s_scan, s_pick, &kr, &ce);
    if (r_lo <= nsample) v_lo = block_select_desc<uint32_t>(s_samp, nsample, r_lo, 32, s_hist, s_scan, s_pick, &kr, &ce);
  }
  if (tid == 0) s_ncand = 0;
  __syncthreads();  // the sample is dead: its buffer now takes the candidates

  // ---- B: the one full pass: count the keys above the bracket, keep the keys inside it (eight loads in flight)
  int cnt_hi = 0;
  for (int i0 = 0; i0 < N; i0 += 8 * nt) {
    unsigned long long key[8];
    bool valid[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const int i = i0 + u * nt + tid;
      valid[u] = i < N;
      key[u] = valid[u] ? (c[i] & MAG) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const uint32_t top = (uint32_t)(key[u] >> 31);
      cnt_hi += valid[u] && top > v_hi;
      const bool hit = valid[u] && top <= v_hi && top >= v_lo;
      const unsigned bal = __ballot_sync(FULL_MASK, hit);
      if (bal) {
        int at = 0;
        if (lane_id() == 0) at = atomicAdd(&s_ncand, __popc(bal));
        at = __shfl_sync(FULL_MASK, at, 0) + __popc(bal & ((1u << lane_id()) - 1u));
        if (hit && at < SEL2_CAND) s_cand[at] = key[u];
      }
    }
  }
  cnt_hi = block_reduce(cnt_hi, s_scan, OpSum(), 0);
  const int nc = s_ncand;  // (block_reduce ended with a barrier)
  if (nc > SEL2_CAND || (long long)cnt_hi >= k || (long long)cnt_hi + nc < k) {  // uniform over the block
    if (tid == 0) need_exact[img] = 1;
    return;
  }

  // ---- C: the (k - cnt_hi)-th largest of the candidates, exactly
  int krem, ceq;
  const unsigned long long tau = block_select_desc<unsigned long long>(s_cand, nc, (int)(k - cnt_hi), 63, s_hist, s_scan,
                                                                      s_pick, &krem, &ceq);
  if (tid == 0) {
    if (krem == ceq) {  // every key equal to tau survives (always, for continuous data): keep key >= tau
      ThrRec r;
      r.tau = tau; r.active = 1; r.pad = 0;
      rec_all[img] = r;
    } else {
      need_exact[img] = 1;
    }
  }
}

// The pending thresholds written out: coefficients below tau become zero, the records are cleared.
__global__ void __launch_bounds__(256) k_apply_threshold(double *coefs_all, int N, ThrRec *rec_all) {
  const int img = blockIdx.y;
  ThrRec *rec = rec_all + img;
  const unsigned long long tau = rec->tau;
  const int active = rec->active;
  unsigned long long *c = reinterpret_cast<unsigned long long *>(coefs_all + (size_t)img * N);
  if (active)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x)
      if ((c[i] & 0x7fffffffffffffffull) < tau) c[i] = 0ull;
}

__global__ void k_clear_thresholds(ThrRec *rec_all, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) { rec_all[i].tau = 0ull; rec_all[i].active = 0; rec_all[i].pad = 0; }
}

// One CLUSTER of SEL_CLUSTER CTAs per image.  Each CTA owns a contiguous slice of the coefficients and, in the
// reduction, a contiguous slice of the digit bins.  Per radix pass: local histogram of the slice -> cluster
// sync -> every CTA sums its bin slice over the cluster's histograms (distributed shared memory) -> cluster
// sync -> all CTAs locate the digit of the k-th largest from the 8 slice sums and the owning CTA's totals.
// Passes 0 and 1 read the coefficients from global memory; pass 1 also copies the keys that match the first
// digit (one binade and a quarter: typically 5-15 % of the slice) into shared memory, and passes 2-4 read only
// those -- unless some CTA of the cluster has more candidates than fit, in which case the cluster keeps reading
// global memory.  (Electing one lane per bin with __match_any_sync in pass 0 was measured: slower, 1.75 against
// 1.25 ms per 512 images.)  Then one read + write pass zeroes what is below the threshold: 3 reads + 1 write per coefficient.
__global__ void __cluster_dims__(SEL_CLUSTER, 1, 1) __launch_bounds__(SEL_THREADS)
    k4_threshold(double *coefs_all, int N, long long k, const int *need_exact) {
  __shared__ int s_hist[SEL_MAXBINS];
  __shared__ int s_tot[SEL_MAXBINS / SEL_CLUSTER];
  __shared__ int s_slice, s_ties;
  __shared__ int s_scan[33];
  __shared__ int s_digit, s_above, s_ceq, s_seen;
  __shared__ int s_ncand, s_over;
  extern __shared__ unsigned long long s_cand[];  // SEL_CAND keys
  if (k <= 0 || k >= (long long)N) return;  // uniform over the grid: no cluster barrier is skipped by a subset
  if (need_exact && !need_exact[blockIdx.x / SEL_CLUSTER]) return;  // k4_select settled this image (uniform over the cluster)
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
  bool from_smem = false;  // passes 2-4 read the candidates kept in shared memory (uniform over the cluster)
  if (tid == 0) { s_ncand = 0; s_over = 0; }
  for (int pass = 0; pass < 5; pass++) {
    const int nb = nbits[pass], nbins = 1 << nb, shift = 63 - done_bits - nb;
    const int bpc = nbins / SEL_CLUSTER;  // bins per CTA in the reduction: 1024 or 512
    for (int i = tid; i < nbins; i += nt) s_hist[i] = 0;
    __syncthreads();
    if (pass == 1) {
      // global read; the matching keys are also appended to s_cand (one shared-memory atomic per warp)
      const int len = hi - lo;
      for (int base = 0; base < len; base += nt) {
        const int i = lo + base + tid;
        const unsigned long long key = base + tid < len ? (c[i] & MAG) : 0ull;
        const bool hit = base + tid < len && (key >> (shift + nb)) == prefix;
        if (hit) atomicAdd(&s_hist[(int)((key >> shift) & (nbins - 1))], 1);
        const unsigned bal = __ballot_sync(FULL_MASK, hit);
        if (bal) {
          int at = 0;
          if (lane_id() == 0) at = atomicAdd(&s_ncand, __popc(bal));
          at = __shfl_sync(FULL_MASK, at, 0) + __popc(bal & ((1u << lane_id()) - 1u));
          if (hit && at < SEL_CAND) s_cand[at] = key;
        }
      }
      __syncthreads();
      if (tid == 0) s_over = s_ncand > SEL_CAND;
    } else if (from_smem) {
      const int nc = s_ncand;
      for (int i = tid; i < nc; i += nt) {
        const unsigned long long key = s_cand[i];
        if ((key >> (shift + nb)) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & (nbins - 1))], 1);
      }
    } else {
      for (int i = lo + tid; i < hi; i += nt) {
        const unsigned long long key = c[i] & MAG;
        if (pass == 0 || (key >> (shift + nb)) == prefix) atomicAdd(&s_hist[(int)((key >> shift) & (nbins - 1))], 1);
      }
    }
    cluster.sync();
    if (pass == 1) {
      int over = 0;
#pragma unroll
      for (int r = 0; r < SEL_CLUSTER; r++) over |= *cluster.map_shared_rank(&s_over, r);
      from_smem = !over;
    }
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

// Rbepwt.threshold_by_percentage(perc) (rbepwt.py:2120-2192), one CTA per region: the region owns, of every level
// l = 1..L, the detail coefficients at the positions of its level-(l+1) segment, and the approximation coefficients of
// its level-(L+1) segment; of these n values the int(min(floor(perc n + 0.5), n)) largest in magnitude are kept, the
// others zeroed -- in the details only: the reference's thresholded approximation is lost again when
// RegionCollection.update() rebuilds the collection from its sub-regions (2187, 1529-1531), so approximation entries
// take part in the ranking but always survive.  Ties at the cut: numpy's unstable argsort in the reference (unpinned);
// here the entries later in the region's list (levels ascending, approximation last) survive.
// off1 / size: the region's level-1 offset and size; level-m segment = [ceil(off1 / 2^(m-1)), ceil((off1+size) / 2^(m-1))).
constexpr int PERC_THREADS = 256;
constexpr int PERC_MAXLEV = 32;

__global__ void __launch_bounds__(PERC_THREADS) k4_percentage(double *coefs_all, int N, int levels, const int32_t *reg_img,
                                                              const int32_t *reg_off, const int32_t *reg_size, int g0,
                                                              double perc) {
  __shared__ int s_hist[SEL2_BINS];
  __shared__ int s_scan[33];
  __shared__ int s_pick[3];
  __shared__ int s_pre[PERC_MAXLEV + 2];          // s_pre[q] = entries of the list before its q-th segment
  __shared__ long long s_base[PERC_MAXLEV + 1];   // flat index of the q-th segment's first entry
  __shared__ int s_seen;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int g = g0 + blockIdx.x;
  const unsigned long long MAG = 0x7fffffffffffffffull;
  unsigned long long *c = reinterpret_cast<unsigned long long *>(coefs_all + (size_t)reg_img[g] * N);
  const long long off1 = reg_off[g], end1 = off1 + reg_size[g];
  if (tid == 0) {
    int acc = 0;
    for (int q = 0; q <= levels; q++) {  // q < levels: details of level q+1; q == levels: approximation -- both on the level-(q+2) ... segment
      const int m = q < levels ? q + 2 : levels + 1;  // segment level
      const long long add = (1ll << (m - 1)) - 1;
      const long long s0 = (off1 + add) >> (m - 1), s1 = (end1 + add) >> (m - 1);
      s_pre[q] = acc;
      s_base[q] = (q < levels ? (long long)N - ((long long)N >> q) : (long long)N - ((long long)N >> levels)) + s0;
      acc += (int)(s1 - s0);
    }
    s_pre[levels + 1] = acc;
  }
  __syncthreads();
  const int n = s_pre[levels + 1], n_det = s_pre[levels];
  if (n == 0) return;
  long long keep = (long long)floor(perc * (double)n + 0.5);
  keep = keep > n ? n : (keep < 0 ? 0 : keep);
  if (keep >= n) return;  // everything survives
  auto flat_of = [&](int p) -> long long {
    int q = 0;
    while (p >= s_pre[q + 1]) q++;
    return s_base[q] + (p - s_pre[q]);
  };
  if (keep == 0) {
    for (int p = tid; p < n_det; p += nt) c[flat_of(p)] = 0ull;
    return;
  }
  int krem, ceq;
  const unsigned long long tau = block_select_desc_fn<unsigned long long>([&](int p) { return c[flat_of(p)] & MAG; }, n, (int)keep, 63,
                                                                         s_hist, s_scan, s_pick, &krem, &ceq);
  if (krem == ceq) {  // every entry equal to tau survives
    for (int p = tid; p < n_det; p += nt) {
      const long long f = flat_of(p);
      if ((c[f] & MAG) < tau) c[f] = 0ull;
    }
    return;
  }
  // only `krem` of the `ceq` entries equal to tau survive: those latest in the list -- walk it from the end
  if (tid == 0) s_seen = 0;
  __syncthreads();
  for (int base = ((n - 1) / nt) * nt; base >= 0; base -= nt) {
    const int p = base + (nt - 1 - tid);  // tid order = descending list position
    const bool valid = p < n;
    const long long f = valid ? flat_of(p) : 0;
    const unsigned long long key = valid ? (c[f] & MAG) : 0ull;
    const bool tie = valid && key == tau;
    int total;
    const int ex = block_exclusive_scan(tie ? 1 : 0, s_scan, &total);
    if (valid && p < n_det && (key < tau || (tie && s_seen + ex >= krem))) c[f] = 0ull;
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
