// K0: label map -> region records, and the size-ordered work queue of the path kernel.
//
// Replaces Segmentation.compute_label_dict (/root/reference/rbepwt.py:840-848): the region
// index is the rank of the label's FIRST APPEARANCE in a row-major scan (not the label value),
// pixels inside a region are in row-major order.  The reference builds per-region Python
// tuples; here a region is a record {label, first pixel, size, level-1 offset, bounding box}
// and its pixel set is implied by the label map (the path kernel rebuilds it as a bitmap).
//
// One CTA per image.  Labels are arbitrary int32 values: a per-image table maps label -> first
// pixel, addressed directly (label - min) when the label range fits the table, else by open
// addressing.  Only the used part of the table is cleared.
#pragma once
#include "common.cuh"

namespace rbepwt {

constexpr unsigned long long TBL_EMPTY = ~0ull;
constexpr int K0_THREADS = 1024;

struct RegionArrays {
  int32_t *label;  // label value
  int32_t *first;  // first pixel (row-major id) == lexicographic-min point == level-1 start point
  int32_t *size;   // number of pixels
  int32_t *off;    // offset of the region in the image's level-1 signal
  int32_t *rmax;   // bounding box (rmin = first / W)
  int32_t *cmin;
  int32_t *cmax;
  int32_t *img;    // image index inside the batch
};

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// Slot of `label` (must have been inserted).
__device__ __forceinline__ uint32_t tbl_lookup(const unsigned long long *tbl, uint32_t tmask, int direct,
                                               int32_t labmin, int32_t label) {
  if (direct) return (uint32_t)(label - labmin);
  uint32_t h = hash32((uint32_t)label) & tmask;
  while (true) {
    unsigned long long e = tbl[h];
    if (e != TBL_EMPTY && (uint32_t)(e >> 32) == (uint32_t)label) return h;
    h = (h + 1) & tmask;
  }
}

// Pass 1: per image, label -> first pixel table and the region count R.
__global__ void __launch_bounds__(K0_THREADS) k0_count(const int32_t *__restrict__ labels, int img0, int N,
                                                       unsigned long long *tbl_all, int T, int32_t *img_R,
                                                       int32_t *img_labmin, int32_t *img_direct) {
  __shared__ int s_red[33];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int img = img0 + blockIdx.x;
  const int32_t *lab = labels + (size_t)img * N;
  unsigned long long *tbl = tbl_all + (size_t)blockIdx.x * T;
  const uint32_t tmask = (uint32_t)T - 1u;

  int lmin = INT32_MAX, lmax = INT32_MIN;
  for (int p = tid; p < N; p += nt) {
    int v = lab[p];
    lmin = min(lmin, v);
    lmax = max(lmax, v);
  }
  lmin = block_reduce(lmin, s_red, OpMin(), INT32_MAX);
  lmax = block_reduce(lmax, s_red, OpMax(), INT32_MIN);
  const long long range = (long long)lmax - (long long)lmin + 1;
  const int direct = range <= (long long)T;
  const int used = direct ? (int)range : T;
  for (int s = tid; s < used; s += nt) tbl[s] = TBL_EMPTY;
  __syncthreads();

  for (int base = 0; base < N; base += nt) {
    const int p = base + tid;
    const bool valid = p < N;
    const int v = valid ? lab[p] : 0;
    const int prev = __shfl_up_sync(FULL_MASK, v, 1);
    // run heads only: a lane whose left neighbour has the same label can never be the first pixel
    if (valid && (lane_id() == 0 || prev != v)) {
      const unsigned long long entry = ((unsigned long long)(uint32_t)v << 32) | (uint32_t)p;
      if (direct) {
        atomicMin(&tbl[v - lmin], entry);
      } else {
        uint32_t h = hash32((uint32_t)v) & tmask;
        while (true) {
          unsigned long long old = *((volatile unsigned long long *)&tbl[h]);
          if (old == TBL_EMPTY) {
            old = atomicCAS(&tbl[h], TBL_EMPTY, entry);
            if (old == TBL_EMPTY) break;
          }
          if ((uint32_t)(old >> 32) == (uint32_t)v) {
            atomicMin(&tbl[h], entry);
            break;
          }
          h = (h + 1) & tmask;
        }
      }
    }
  }
  __syncthreads();
  int cnt = 0;
  for (int s = tid; s < used; s += nt) cnt += tbl[s] != TBL_EMPTY;
  cnt = block_reduce(cnt, s_red, OpSum(), 0);
  if (tid == 0) {
    img_R[img] = cnt;
    img_labmin[img] = lmin;
    img_direct[img] = direct;
  }
}

// Pass 2: region records.  rbase[img] = index of the image's region 0 in the global region arrays.
__global__ void __launch_bounds__(K0_THREADS) k0_regions(const int32_t *__restrict__ labels, int img0, int N,
                                                         int logW, const unsigned long long *tbl_all,
                                                         int32_t *slot_rid_all, int T, const int32_t *img_R,
                                                         const int32_t *img_labmin, const int32_t *img_direct,
                                                         const int32_t *img_rbase, RegionArrays reg) {
  __shared__ int s_scan[33];
  __shared__ int s_running;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int img = img0 + blockIdx.x;
  const int32_t *lab = labels + (size_t)img * N;
  const unsigned long long *tbl = tbl_all + (size_t)blockIdx.x * T;
  int32_t *slot_rid = slot_rid_all + (size_t)blockIdx.x * T;
  const uint32_t tmask = (uint32_t)T - 1u;
  const int W = 1 << logW;
  const int labmin = img_labmin[img], direct = img_direct[img], R = img_R[img], rb = img_rbase[img];

  // A: rank the first-appearance pixels in row-major order -> region ids
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < N; base += nt) {
    const int p = base + tid;
    const bool valid = p < N;
    const int v = valid ? lab[p] : 0;
    uint32_t slot = 0;
    bool flag = false;
    if (valid) {
      slot = tbl_lookup(tbl, tmask, direct, labmin, v);
      flag = (uint32_t)tbl[slot] == (uint32_t)p;
    }
    int total;
    const int rank = s_running + block_exclusive_scan(flag ? 1 : 0, s_scan, &total);
    if (flag) {
      const int g = rb + rank;
      reg.label[g] = v;
      reg.first[g] = p;
      reg.img[g] = img;
      reg.size[g] = 0;
      reg.rmax[g] = 0;
      reg.cmin[g] = W;
      reg.cmax[g] = 0;
      slot_rid[slot] = rank;
    }
    __syncthreads();
    if (tid == 0) s_running += total;
    __syncthreads();
  }
  __threadfence_block();
  __syncthreads();

  // B: sizes and bounding boxes (warp-aggregated atomics)
  for (int base = 0; base < N; base += nt) {
    const int p = base + tid;
    const bool valid = p < N;
    int rid = -1;
    if (valid) rid = slot_rid[tbl_lookup(tbl, tmask, direct, labmin, lab[p])];
    const unsigned grp = __match_any_sync(FULL_MASK, rid);
    if (valid) {
      const int row = p >> logW, col = p & (W - 1);
      const int cmn = __reduce_min_sync(grp, col), cmx = __reduce_max_sync(grp, col);
      const int rmx = __reduce_max_sync(grp, row);
      if ((int)lane_id() == __ffs(grp) - 1) {
        const int g = rb + rid;
        atomicAdd(&reg.size[g], __popc(grp));
        atomicMax(&reg.rmax[g], rmx);
        atomicMin(&reg.cmin[g], cmn);
        atomicMax(&reg.cmax[g], cmx);
      }
    }
  }
  __threadfence_block();
  __syncthreads();

  // C: level-1 offsets = exclusive scan of the sizes in region order
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < R; base += nt) {
    const int r = base + tid;
    const int sz = r < R ? reg.size[rb + r] : 0;
    int total;
    const int ex = s_running + block_exclusive_scan(sz, s_scan, &total);
    if (r < R) reg.off[rb + r] = ex;
    __syncthreads();
    if (tid == 0) s_running += total;
    __syncthreads();
  }
}

// EPWT: one region per image holding every pixel (rbepwt.py:2004-2006).
__global__ void k0_single_region(int img0, int nimg, int H, int W, RegionArrays reg, int32_t *img_R,
                                 int32_t *img_rbase) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nimg) return;
  const int img = img0 + i;
  reg.label[img] = 0; reg.first[img] = 0; reg.size[img] = H * W; reg.off[img] = 0;
  reg.rmax[img] = H - 1; reg.cmin[img] = 0; reg.cmax[img] = W - 1; reg.img[img] = img;
  img_R[img] = 1;
  img_rbase[img] = img;
}

// ---------------------------------------------------------------- work queue -------------
// Regions are processed largest first (longest sequential chain first).  128 bins:
// class (0 = bitmap too large for a shared-memory slot, 1 = fits) x 64 size bins (descending).

constexpr int Q_BINS = 128;

__device__ __forceinline__ int region_bitmap_words(const RegionArrays &reg, int g, int logW) {
  const int h = reg.rmax[g] - (reg.first[g] >> logW) + 1;
  const int w = reg.cmax[g] - reg.cmin[g] + 1;
  return h * ((w + 31) >> 5);
}

__device__ __forceinline__ int queue_bin(int size, int words, int slot_words) {
  const int lg = 31 - __clz(size);                                   // size >= 1
  const int key = size >= 2 ? 2 * lg + ((size >> (lg - 1)) & 1) : 0;  // <= 61
  return (words > slot_words ? 0 : 64) + 63 - key;
}

__global__ void kq_hist(RegionArrays reg, int g0, int nreg, int logW, int slot_words, int *qhist) {
  __shared__ int s_h[Q_BINS];
  for (int i = threadIdx.x; i < Q_BINS; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nreg; i += gridDim.x * blockDim.x) {
    const int g = g0 + i;
    atomicAdd(&s_h[queue_bin(reg.size[g], region_bitmap_words(reg, g, logW), slot_words)], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Q_BINS; i += blockDim.x)
    if (s_h[i]) atomicAdd(&qhist[i], s_h[i]);
}

// qmeta: [0..127] bin write cursors, [128] = number of class-0 (big) regions, [129] = nreg,
//        [130] big-queue consumer cursor, [131] small-queue consumer cursor, [132] error flag
constexpr int QM_NBIG = 128, QM_NREG = 129, QM_CUR_BIG = 130, QM_CUR_SMALL = 131, QM_ERR = 132, QM_SIZE = 136;

__global__ void kq_scan(int *qhist, int *qmeta, int nreg) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < Q_BINS; i++) {
      if (i == 64) qmeta[QM_NBIG] = acc;
      qmeta[i] = acc;
      acc += qhist[i];
      qhist[i] = 0;  // ready for the next chunk
    }
    qmeta[QM_NREG] = nreg;
    qmeta[QM_CUR_BIG] = 0;
    qmeta[QM_CUR_SMALL] = 0;
  }
}

__global__ void kq_scatter(RegionArrays reg, int g0, int nreg, int logW, int slot_words, int *qmeta,
                           int32_t *queue) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nreg; i += gridDim.x * blockDim.x) {
    const int g = g0 + i;
    const int bin = queue_bin(reg.size[g], region_bitmap_words(reg, g, logW), slot_words);
    queue[atomicAdd(&qmeta[bin], 1)] = g;
  }
}

}  // namespace rbepwt
