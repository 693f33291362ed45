// On-demand views for the reference's attribute surface (not on the hot path):
//   - level-1 INCOMING order of every region (row-major inside the region,
//     Segmentation.compute_label_dict, /root/reference/rbepwt.py:840-848);
//   - Region.permutation at any level (rbepwt.py:1285, 1333): index, in the region's incoming
//     order, of its t-th path point.  Incoming order at level l >= 2 is the previous level's path
//     order subsampled at the even global positions (RegionCollection.reduce, 1563-1584).
#pragma once
#include "common.cuh"
#include "regions.cuh"

namespace rbepwt {

// One warp per region of image `img`: inc[off_r + rank] = pixel id, rank = row-major rank inside the region.
__global__ void k_level1_incoming(const int32_t *__restrict__ lab, int logW, RegionArrays reg, int g0, int R,
                                  int32_t *inc) {
  const int lane = (int)lane_id();
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  const int g = g0 + r;
  const int label = reg.label[g], r0 = reg.first[g] >> logW, c0 = reg.cmin[g];
  const int h = reg.rmax[g] - r0 + 1, w = reg.cmax[g] - c0 + 1, ws = (w + 31) >> 5;
  int pos = reg.off[g];
  for (int i = 0; i < h; i++)
    for (int wd = 0; wd < ws; wd++) {
      const int col = c0 + (wd << 5) + lane;
      const int pix = ((r0 + i) << logW) + col;
      const bool in = col < c0 + w && lab[pix] == label;
      const unsigned bits = __ballot_sync(FULL_MASK, in);
      if (in) inc[pos + __popc(bits & lanemask_lt())] = pix;
      pos += __popc(bits);
    }
}

// inv[src[i * stride]] = i for i < n  (pixel -> index in the incoming order)
__global__ void k_inv_scatter(const int32_t *__restrict__ src, int stride, int n, int32_t *inv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[src[(size_t)i * stride]] = i;
}

// perm[t] = inv[Ql[t]] - off_l[r(t)],  off_l[r] = ceil(off1[r] / 2^(lev-1)), r(t) = last r with off_l[r] <= t
__global__ void k_perm_gather(const int32_t *__restrict__ Ql, int n, const int32_t *__restrict__ inv,
                              const int32_t *__restrict__ off1, int R, int lev, int32_t *perm) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int sh = lev - 1, add = (1 << sh) - 1;
  int lo = 0, hi = R - 1;  // off_l[0] = 0 <= t
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (((off1[mid] + add) >> sh) <= t) lo = mid; else hi = mid - 1;
  }
  perm[t] = inv[Ql[t]] - ((off1[lo] + add) >> sh);
}

// paths_first_level (Region.same_path, rbepwt.py:1183-1188): at every level >= 2 the path is the incoming
// order itself -- the points at the even positions of the level above, identity permutation.
// One CTA per image, level after level.
__global__ void __launch_bounds__(1024) k_same_paths(int32_t *Q_all, int32_t *Pm_all, int N, int levels) {
  int32_t *Q = Q_all + (size_t)blockIdx.x * 2 * (size_t)N, *Pm = Pm_all + (size_t)blockIdx.x * 2 * (size_t)N;
  for (int lev = 2; lev <= levels; lev++) {
    const int n = N >> (lev - 1);
    const int32_t *prev = Q + level_off((size_t)N, lev - 1);
    int32_t *cur = Q + level_off((size_t)N, lev), *pos = Pm + level_off((size_t)N, lev);
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      cur[t] = prev[2 * t];
      pos[t] = t;
    }
    __syncthreads();  // the next level reads what this one wrote (same CTA)
  }
}

// out[i] = V[src[i * stride]]
__global__ void k_gather_values(const double *__restrict__ V, const int32_t *__restrict__ src, int stride, int n,
                                double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = V[src[(size_t)i * stride]];
}

}  // namespace rbepwt
