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
#include <cooperative_groups.h>

#include "common.cuh"

namespace rbepwt {

namespace cg = cooperative_groups;

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
// One thread-block CLUSTER of K0_CLUSTER CTAs per image: every CTA owns a contiguous slice of the pixels, the
// phases (min/max -> clear the table -> insert run heads -> count) are separated by cluster barriers, and the
// per-CTA partial results are combined through distributed shared memory.
constexpr int K0_CLUSTER = 8;

__device__ __forceinline__ void k0_slice(int N, int rank, int &lo, int &hi) {
  const int per = (((N + K0_CLUSTER - 1) / K0_CLUSTER) + 31) & ~31;  // warp-aligned slices
  lo = min(N, rank * per);
  hi = min(N, lo + per);
}

__global__ void __cluster_dims__(K0_CLUSTER, 1, 1) __launch_bounds__(K0_THREADS)
    k0_count(const int32_t *__restrict__ labels, int img0, int N, unsigned long long *tbl_all, int T, int32_t *img_R,
             int32_t *img_labmin, int32_t *img_direct) {
  __shared__ int s_red[33];
  __shared__ int s_mm[2];
  __shared__ int s_cnt;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, nt = blockDim.x;
  const int li = blockIdx.x / K0_CLUSTER, img = img0 + li;
  const int32_t *lab = labels + (size_t)img * N;
  unsigned long long *tbl = tbl_all + (size_t)li * T;
  const uint32_t tmask = (uint32_t)T - 1u;
  int lo, hi;
  k0_slice(N, rank, lo, hi);

  int lmin = INT32_MAX, lmax = INT32_MIN;
  for (int p = lo + tid; p < hi; p += nt) {
    const int v = lab[p];
    lmin = min(lmin, v);
    lmax = max(lmax, v);
  }
  lmin = block_reduce(lmin, s_red, OpMin(), INT32_MAX);
  lmax = block_reduce(lmax, s_red, OpMax(), INT32_MIN);
  if (tid == 0) { s_mm[0] = lmin; s_mm[1] = lmax; }
  cluster.sync();
  for (int r = 0; r < K0_CLUSTER; r++) {
    const int *mm = cluster.map_shared_rank(s_mm, r);
    lmin = min(lmin, mm[0]);
    lmax = max(lmax, mm[1]);
  }
  const long long range = (long long)lmax - (long long)lmin + 1;
  const int direct = range <= (long long)T;
  const int used = direct ? (int)range : T;
  for (int s = rank * nt + tid; s < used; s += K0_CLUSTER * nt) tbl[s] = TBL_EMPTY;
  __threadfence();
  cluster.sync();

  for (int base = lo; base < hi; base += nt) {
    const int p = base + tid;
    const bool valid = p < hi;
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
  __threadfence();
  cluster.sync();

  int cnt = 0;
  for (int s = rank * nt + tid; s < used; s += K0_CLUSTER * nt) cnt += __ldcg(&tbl[s]) != TBL_EMPTY;
  cnt = block_reduce(cnt, s_red, OpSum(), 0);
  if (tid == 0) s_cnt = cnt;
  cluster.sync();
  if (rank == 0 && tid == 0) {
    int total = 0;
    for (int r = 0; r < K0_CLUSTER; r++) total += *cluster.map_shared_rank(&s_cnt, r);
    img_R[img] = total;
    img_labmin[img] = lmin;
    img_direct[img] = direct ? used : 0;  // slots of the direct table (>= 1), 0 = open addressing
  }
  cluster.sync();  // s_cnt stays readable until rank 0 has summed it
}

// Pass 2, common case (dense label values, at most K0F_MAXR regions, rows a multiple of 8 pixels): the label
// table and the per-region statistics live in shared memory, a thread handles 8 consecutive pixels, ranks and
// statistics are produced in ONE sweep over the labels (a label's first pixel precedes all its other pixels in
// row-major order, so its region id is known by the time they are counted).
constexpr int K0F_MAXT = 8192, K0F_MAXR = 4096, K0F_PPT = 8;
constexpr size_t K0F_SMEM = (size_t)(K0F_MAXT + 4 * K0F_MAXR) * sizeof(int);

__host__ __device__ __forceinline__ bool k0_fast_eligible(int direct_slots, int R, int logW) {
  return direct_slots > 0 && direct_slots <= K0F_MAXT && R <= K0F_MAXR && logW >= 3;
}

__global__ void __launch_bounds__(K0_THREADS) k0_regions_fast(const int32_t *__restrict__ labels, int img0, int N,
                                                              int logW, const unsigned long long *tbl_all, int T,
                                                              const int32_t *img_R, const int32_t *img_labmin,
                                                              const int32_t *img_direct, const int32_t *img_rbase,
                                                              RegionArrays reg) {
  extern __shared__ int s_dyn[];
  int *s_rid = s_dyn;                // slot -> first pixel | TAG, then -> region id
  int *s_size = s_rid + K0F_MAXT, *s_rmax = s_size + K0F_MAXR, *s_cmin = s_rmax + K0F_MAXR, *s_cmax = s_cmin + K0F_MAXR;
  __shared__ int s_scan[33];
  __shared__ int s_running;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int img = img0 + blockIdx.x;
  const int slots = img_direct[img], R = img_R[img];
  if (!k0_fast_eligible(slots, R, logW)) return;  // k0_regions handles this image
  const int32_t *lab = labels + (size_t)img * N;
  const unsigned long long *tbl = tbl_all + (size_t)blockIdx.x * T;
  const int W = 1 << logW, labmin = img_labmin[img], rb = img_rbase[img];
  const int TAG = (int)0x80000000u;

  for (int s = tid; s < slots; s += nt) {
    const unsigned long long e = tbl[s];
    s_rid[s] = e == TBL_EMPTY ? -1 : ((int)(uint32_t)e | TAG);
  }
  for (int r = tid; r < R; r += nt) { s_size[r] = 0; s_rmax[r] = 0; s_cmin[r] = W; s_cmax[r] = 0; }
  if (tid == 0) s_running = 0;
  __syncthreads();

  for (int base = 0; base < N; base += nt * K0F_PPT) {
    const int p0 = base + tid * K0F_PPT;
    const bool valid = p0 < N;  // N is a multiple of 8 here: all of a thread's pixels are valid or none
    int v[K0F_PPT];
    int prevlab = 0;
    if (valid) {
      const int4 x = *reinterpret_cast<const int4 *>(lab + p0), y = *reinterpret_cast<const int4 *>(lab + p0 + 4);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
      prevlab = p0 > 0 ? lab[p0 - 1] : ~v[0];
    }
    // first appearances among the run heads
    unsigned fmask = 0;
    if (valid) {
#pragma unroll
      for (int i = 0; i < K0F_PPT; i++) {
        const bool head = v[i] != (i == 0 ? prevlab : v[i - 1]);
        if (head && s_rid[v[i] - labmin] == ((p0 + i) | TAG)) fmask |= 1u << i;
      }
    }
    int total;
    int rank = s_running + block_exclusive_scan(__popc(fmask), s_scan, &total);
    while (fmask) {
      const int i = __ffs(fmask) - 1;
      fmask &= fmask - 1;
      const int g = rb + rank;
      s_rid[v[i] - labmin] = rank++;
      reg.label[g] = v[i];
      reg.first[g] = p0 + i;
      reg.img[g] = img;
    }
    __syncthreads();
    if (tid == 0) s_running += total;  // read again at the top of the next tile: the barrier below orders it
    // statistics, one update per run of equal labels inside the thread's 8 pixels (same row: 8 | W)
    if (valid) {
      const int row = p0 >> logW, col0 = p0 & (W - 1);
      int i = 0;
      while (i < K0F_PPT) {
        int j = i;
        while (j + 1 < K0F_PPT && v[j + 1] == v[i]) j++;
        const int rid = s_rid[v[i] - labmin];
        atomicAdd(&s_size[rid], j - i + 1);
        atomicMax(&s_rmax[rid], row);
        atomicMin(&s_cmin[rid], col0 + i);
        atomicMax(&s_cmax[rid], col0 + j);
        i = j + 1;
      }
    }
    __syncthreads();  // s_running of this tile is visible before the next tile's scan reads it
  }
  // level-1 offsets = exclusive scan of the sizes in region order
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < R; base += nt) {
    const int r = base + tid;
    const int sz = r < R ? s_size[r] : 0;
    int total;
    const int ex = s_running + block_exclusive_scan(sz, s_scan, &total);
    if (r < R) {
      const int g = rb + r;
      reg.size[g] = sz; reg.off[g] = ex;
      reg.rmax[g] = s_rmax[r]; reg.cmin[g] = s_cmin[r]; reg.cmax[g] = s_cmax[r];
    }
    __syncthreads();
    if (tid == 0) s_running += total;
    __syncthreads();
  }
}

// Pass 2, general case (any int32 labels, any region count), one cluster of K0_CLUSTER CTAs per image like
// k0_count.  rbase[img] = index of the image's region 0 in the global region arrays.
//   A1 every CTA counts the first-appearance pixels of its slice; A2 ranks them after the cluster prefix;
//   B  sizes and bounding boxes (warp-aggregated atomics);  C  level-1 offsets (rank 0).
__device__ __forceinline__ bool k0_is_first(const unsigned long long *tbl, uint32_t tmask, int direct, int labmin,
                                            const int32_t *lab, int p, int v, int prev_in_warp, uint32_t &slot) {
  const int prev = lane_id() == 0 ? (p > 0 ? lab[p - 1] : ~v) : prev_in_warp;
  if (prev == v) return false;  // not a run head: cannot be the label's first pixel
  slot = tbl_lookup(tbl, tmask, direct, labmin, v);
  return (uint32_t)__ldcg(&tbl[slot]) == (uint32_t)p;
}

__global__ void __cluster_dims__(K0_CLUSTER, 1, 1) __launch_bounds__(K0_THREADS)
    k0_regions(const int32_t *__restrict__ labels, int img0, int N, int logW, const unsigned long long *tbl_all,
               int32_t *slot_rid_all, int T, const int32_t *img_R, const int32_t *img_labmin, const int32_t *img_direct,
               const int32_t *img_rbase, RegionArrays reg) {
  __shared__ int s_scan[33];
  __shared__ int s_running, s_cnt;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, nt = blockDim.x;
  const int li = blockIdx.x / K0_CLUSTER, img = img0 + li;
  const int32_t *lab = labels + (size_t)img * N;
  const unsigned long long *tbl = tbl_all + (size_t)li * T;
  int32_t *slot_rid = slot_rid_all + (size_t)li * T;
  const uint32_t tmask = (uint32_t)T - 1u;
  const int W = 1 << logW;
  const int labmin = img_labmin[img], direct = img_direct[img], R = img_R[img], rb = img_rbase[img];
  if (k0_fast_eligible(direct, R, logW)) return;  // k0_regions_fast handles this image (uniform over the cluster)
  int lo, hi;
  k0_slice(N, rank, lo, hi);

  // A1: first-appearance pixels in this slice
  int cnt = 0;
  for (int base = lo; base < hi; base += nt) {
    const int p = base + tid;
    const bool valid = p < hi;
    const int v = valid ? lab[p] : 0;
    const int prev = __shfl_up_sync(FULL_MASK, v, 1);
    uint32_t slot;
    cnt += valid && k0_is_first(tbl, tmask, direct, labmin, lab, p, v, prev, slot);
  }
  cnt = block_reduce(cnt, s_scan, OpSum(), 0);
  if (tid == 0) s_cnt = cnt;
  cluster.sync();
  if (tid == 0) {
    int before = 0;
    for (int r = 0; r < rank; r++) before += *cluster.map_shared_rank(&s_cnt, r);
    s_running = before;
  }
  __syncthreads();

  // A2: rank them in row-major order -> region ids
  for (int base = lo; base < hi; base += nt) {
    const int p = base + tid;
    const bool valid = p < hi;
    const int v = valid ? lab[p] : 0;
    const int prev = __shfl_up_sync(FULL_MASK, v, 1);
    uint32_t slot = 0;
    const bool flag = valid && k0_is_first(tbl, tmask, direct, labmin, lab, p, v, prev, slot);
    int total;
    const int rnk = s_running + block_exclusive_scan(flag ? 1 : 0, s_scan, &total);
    if (flag) {
      const int g = rb + rnk;
      reg.label[g] = v;
      reg.first[g] = p;
      reg.img[g] = img;
      reg.size[g] = 0;
      reg.rmax[g] = 0;
      reg.cmin[g] = W;
      reg.cmax[g] = 0;
      slot_rid[slot] = rnk;
    }
    __syncthreads();
    if (tid == 0) s_running += total;
    __syncthreads();
  }
  __threadfence();
  cluster.sync();

  // B: sizes and bounding boxes (warp-aggregated atomics)
  for (int base = lo; base < hi; base += nt) {
    const int p = base + tid;
    const bool valid = p < hi;
    int rid = -1;
    if (valid) rid = __ldcg(&slot_rid[tbl_lookup(tbl, tmask, direct, labmin, lab[p])]);
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
  __threadfence();
  cluster.sync();
  if (rank != 0) return;

  // C: level-1 offsets = exclusive scan of the sizes in region order
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < R; base += nt) {
    const int r = base + tid;
    const int sz = r < R ? __ldcg(&reg.size[rb + r]) : 0;
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
// The path kernels pull regions from a queue ordered so that (a) regions whose bounding-box bitmap
// does not fit a warp's shared-memory arena come first (class 0, one warp per region, paths.cuh),
// (b) the others are grouped by bitmap size class, largest first: class c >= 1 holds the bitmaps that fit
// its kernel's arena class_chunk_size(c) at a time (1; 4, 6, 8, 12 in the windowed kernel's; 8, ..., 32 in the bulk kernel's),
// (c) inside a class regions are sorted by pixel count, descending, in eighth-octave bins: the lanes
// of a warp walk chains of similar length, and the longest chains start first.  Chunks are cut from the class's
// queue range regardless of the bins (only the last chunk of a class is partial).
// A chunk = the regions one warp walks together (thread per region, walk.cuh).

#ifndef TPR_ARENA_WORDS_N
#define TPR_ARENA_WORDS_N 2144
#endif
constexpr int TPR_ARENA_WORDS = TPR_ARENA_WORDS_N;  // shared-memory words per warp of k1_walk
// class 0 = big; classes 1..12 = the bitmaps that fit the arena class_chunk_size(c) at a time
constexpr int Q_NCLS = 13;
constexpr int Q_SIZE_BINS = 256;
constexpr int Q_BINS = Q_NCLS * Q_SIZE_BINS;

// Whole-warp walking (k1_coop_all) of regions that would fit a lane's slot: the largest regions of a group, at least
// TPR_COOP_MIN pixels each and at most COOP_PER_SM per SM of them (kq_scan), and every region of a small group.
// Measured (tools/coop_sweep.py, 256 images of 512^2): EVERY region of >= 2048 pixels on a warp of its own is 1.6x slower on
// heavy-tailed maps (40 such regions per image) -- a warp on one chain spends ~10x the issue slots per step.
#ifndef TPR_COOP_MIN_N
#define TPR_COOP_MIN_N 2048
#endif
constexpr int TPR_COOP_MIN = TPR_COOP_MIN_N;
#ifndef COOP_PER_SM_N
#define COOP_PER_SM_N 4
#endif
constexpr int COOP_PER_SM = COOP_PER_SM_N;
constexpr int TPR_COOP_ALL_BELOW = 4096;      // every region gets a warp when the whole group has at most this many regions
constexpr int TPR_MAX_SIDE = 1024;     // bounding-box side limit of k1_walk (its packed candidate key)

__device__ __forceinline__ int region_bitmap_words(const RegionArrays &reg, int g, int logW) {
  const int h = reg.rmax[g] - (reg.first[g] >> logW) + 1;
  const int w = reg.cmax[g] - reg.cmin[g] + 1;
  return h * ((w + 31) >> 5);
}

// k1_walk (walk.cuh) keeps a margin of WK_PAD = 2 empty rows / columns around the bounding box, so that the 5x5
// neighbourhood of any point of the region lies inside the bitmap (no bounds checks in its step), and TWO planes per
// region: the level's unvisited points and the survivors collected for the next level.
constexpr int WK_PAD = 2;
#ifndef WK_LIST_MAX_N
#define WK_LIST_MAX_N 32
#endif
constexpr int WK_LIST_MAX = WK_LIST_MAX_N;  // list mode from the first level with at most this many points (<= 32)

// Arena words of a region's slot: two bitmap planes of P words + 1 (the window fetch of a multi-word row reads one
// word past the row's last).  P >= WK_LIST_MAX: the plane not in use holds the list of the first list-mode level.
// The sum is odd, so the slots of a warp's lanes start in different shared-memory banks.
__host__ __device__ __forceinline__ int wk_plane_words(int h, int ws) { return max(h * ws, WK_LIST_MAX); }
__host__ __device__ __forceinline__ int wk_slot_words(int h, int ws) { return 2 * wk_plane_words(h, ws) + 1; }

// Slot words used for the queue class: regions with a side above TPR_MAX_SIDE count as oversized.
__device__ __forceinline__ int region_class_words(const RegionArrays &reg, int g, int logW) {
  const int h = reg.rmax[g] - (reg.first[g] >> logW) + 1 + 2 * WK_PAD;
  const int w = reg.cmax[g] - reg.cmin[g] + 1 + 2 * WK_PAD;
  return (h > TPR_MAX_SIDE || w > TPR_MAX_SIDE) ? INT32_MAX : wk_slot_words(h, (w + 31) >> 5);
}

// Regions per chunk of a class.  Classes 6..12 (the bulk instantiation of k1_walk, arenas of TPR_ARENA_WORDS): the slots
// of at most TPR_ARENA_WORDS / size words; two planes make slots twice as large as a single bitmap, so the steps between
// the classes are fine: a region a little too large for 32 per warp is walked 28 at a time, not 16.  Classes 2..5 (the
// windowed instantiation, fewer warps with arenas of TPR_WIDE_ARENA_WORDS): the slots of at most TPR_WIDE_ARENA_WORDS /
// size words.  Class 1: the larger slots that still fit TPR_ARENA_WORDS, one per chunk (walked by a whole warp where the
// path mode has such a kernel).
#ifndef TPR_WIDE_ARENA_WORDS_N
#define TPR_WIDE_ARENA_WORDS_N 4096
#endif
constexpr int TPR_WIDE_ARENA_WORDS = TPR_WIDE_ARENA_WORDS_N;
__host__ __device__ __forceinline__ int class_chunk_size(int cls) {
  constexpr int cs[Q_NCLS] = {1, 1, 4, 6, 8, 12, 8, 12, 16, 20, 24, 28, 32};
  return cs[cls];
}
__host__ __device__ __forceinline__ int class_arena_words(int cls) { return cls < 6 ? TPR_WIDE_ARENA_WORDS : TPR_ARENA_WORDS; }
static_assert(TPR_WIDE_ARENA_WORDS / 12 >= TPR_ARENA_WORDS / 8, "class 5 must take over where class 6 ends");

// eighth-octave size key (monotone in size, <= 8*30+7) and the smallest size of a key
__device__ __forceinline__ int size_key(int size) {
  const int lg = 31 - __clz(size);  // size >= 1
  return size >= 8 ? 8 * lg + ((size >> (lg - 3)) & 7) : size;
}
__device__ __forceinline__ int key_min_size(int key) { return key >= 24 ? (8 + (key & 7)) << ((key >> 3) - 3) : key; }

// Is a region of `words` slot words and `size` pixels walked by a whole warp?  (coop: the path mode has a whole-warp
// kernel for regions that fit an arena: euclid and gradpath)  Oversized bitmaps always are (k1_paths_big); of the
// others, with coop, class 1: the slots that fit the arena only one at a time, and the regions of >= coop_size pixels.
__device__ __forceinline__ bool walked_by_warp(int size, int words, int coop_size, bool coop) {
  return words > TPR_ARENA_WORDS || (coop && (words * class_chunk_size(2) > TPR_WIDE_ARENA_WORDS || size >= coop_size));
}

__device__ __forceinline__ int queue_bin(int size, int words, int coop_min) {
  const int key = size_key(size);
  int cls = 0;
  if (words <= TPR_ARENA_WORDS) {
    cls = Q_NCLS - 1;
    while (cls > 1 && words * class_chunk_size(cls) > class_arena_words(cls)) cls--;
    if (size >= coop_min) cls = 1;  // a long chain gets a warp of its own (one region per chunk)
  }
  return cls * Q_SIZE_BINS + (Q_SIZE_BINS - 1 - key);
}

__global__ void kq_hist(RegionArrays reg, int g0, int nreg, int logW, int coop_min, int *qhist) {
  __shared__ int s_h[Q_BINS];
  for (int i = threadIdx.x; i < Q_BINS; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nreg; i += gridDim.x * blockDim.x) {
    const int g = g0 + i;
    atomicAdd(&s_h[queue_bin(reg.size[g], region_class_words(reg, g, logW), coop_min)], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Q_BINS; i += blockDim.x)
    if (s_h[i]) atomicAdd(&qhist[i], s_h[i]);
}

// qmeta: [0, Q_BINS) bin write cursors (scatter), then the named slots below.
// qbins: per class: [0, Q_NCLS) queue start, [Q_NCLS, 2 Q_NCLS) count, [2 Q_NCLS, 3 Q_NCLS) first chunk.
constexpr int QM_NBIG = Q_BINS, QM_NREG = Q_BINS + 1, QM_CUR_BIG = Q_BINS + 2, QM_CUR_SMALL = Q_BINS + 3,
              QM_ERR = Q_BINS + 4, QM_NCHUNKS = Q_BINS + 5, QM_CHUNK_SPLIT = Q_BINS + 6, QM_CUR_WIDE = Q_BINS + 7,
              QM_COOP_SIZE = Q_BINS + 8,  // regions of at least this many pixels are walked by a whole warp (kq_scan decides)
              QM_CLS1_CHUNKS = Q_BINS + 9,  // chunks (= regions) of class 1: the first chunks of the table
              QM_CUR_COOP = Q_BINS + 10,    // k1_coop_all's chunk cursor
              QM_SIZE = Q_BINS + 11;
constexpr int Q_FIRST_NARROW_CLS = 6;  // classes 1..5 (slots of more than TPR_ARENA_WORDS / 8 words): the windowed variant of k1_walk

// One warp.  First the whole-warp threshold: kq_hist filed every region by its bitmap size alone when coop_min > 1; here
// the LARGEST regions -- at most coop_limit of them, none smaller than coop_min -- move to class 1 (one region per chunk),
// which k1_coop_all walks with a whole warp each (regwin.cuh).  A whole warp steps a chain about twice as fast as a lane
// does, and the longest chain of a group is what its path stage lasts; but a warp spends ~10x the issue slots per step,
// so this pays for the few longest chains only, not for thousands of them.  The threshold lies on a size-key boundary,
// so moving whole bins is exact.  coop_min <= 1: the host asked for a warp per region (gradpath, small groups), kq_hist
// filed them so already.
// Then the exclusive scan of the bin counts (queue offsets); per class: queue start, count, first chunk.
__global__ void kq_scan(int *qhist, int *qmeta, int *qbins, int nreg, int coop_min, int coop_limit) {
  const int lane = threadIdx.x;
  int coop_size = coop_min;
  if (coop_min > 1) {
    constexpr int PER = Q_SIZE_BINS / 32;
    int tot[PER], lsum = 0;
#pragma unroll
    for (int u = 0; u < PER; u++) {
      int t = 0;
      for (int cls = 2; cls < Q_NCLS; cls++) t += qhist[cls * Q_SIZE_BINS + lane * PER + u];
      tot[u] = t;
      lsum += t;
    }
    int inc = lsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(FULL_MASK, inc, d);
      if (lane >= d) inc += y;
    }
    const int kmin = size_key(coop_min);
    int cum = inc - lsum, last = -1;
#pragma unroll
    for (int u = 0; u < PER; u++) {  // bin j holds the size key Q_SIZE_BINS - 1 - j: bins in descending size
      const int j = lane * PER + u;
      cum += tot[u];
      if (cum <= coop_limit && Q_SIZE_BINS - 1 - j >= kmin) last = j;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) last = max(last, __shfl_xor_sync(FULL_MASK, last, d));
    coop_size = last >= 0 ? key_min_size(Q_SIZE_BINS - 1 - last) : INT32_MAX;
#pragma unroll
    for (int u = 0; u < PER; u++) {
      const int j = lane * PER + u;
      if (j <= last && tot[u]) {
        for (int cls = 2; cls < Q_NCLS; cls++) qhist[cls * Q_SIZE_BINS + j] = 0;
        qhist[Q_SIZE_BINS + j] += tot[u];
      }
    }
    __syncwarp();
  }
  if (lane == 0) qmeta[QM_COOP_SIZE] = coop_size;
  int acc = 0, cacc = 0;
  for (int cls = 0; cls < Q_NCLS; cls++) {
    const int cstart = acc;
    for (int base = cls * Q_SIZE_BINS; base < (cls + 1) * Q_SIZE_BINS; base += 32) {
      const int b = base + lane;
      const int cnt = qhist[b];
      int inc = cnt;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= d) inc += y;
      }
      qmeta[b] = acc + inc - cnt;
      qhist[b] = 0;  // ready for the next group of images
      acc += __shfl_sync(FULL_MASK, inc, 31);
    }
    const int ccount = acc - cstart, cs = class_chunk_size(cls);
    if (lane == 0) {
      qbins[cls] = cstart;
      qbins[Q_NCLS + cls] = ccount;
      qbins[2 * Q_NCLS + cls] = cacc;
      if (cls == 1) { qmeta[QM_NBIG] = cstart; qmeta[QM_CLS1_CHUNKS] = ccount; }
      if (cls == Q_FIRST_NARROW_CLS) qmeta[QM_CHUNK_SPLIT] = cacc;
    }
    if (cls >= 1) cacc += (ccount + cs - 1) / cs;
  }
  if (lane == 0) {
    qmeta[QM_NREG] = nreg;
    qmeta[QM_CUR_BIG] = 0;
    qmeta[QM_CUR_SMALL] = 0;
    qmeta[QM_CUR_WIDE] = 0;
    qmeta[QM_CUR_COOP] = 0;
    qmeta[QM_NCHUNKS] = cacc;
  }
}

__global__ void kq_scatter(RegionArrays reg, int g0, int nreg, int logW, int *qmeta, int32_t *queue) {
  const int coop_min = qmeta[QM_COOP_SIZE];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nreg; i += gridDim.x * blockDim.x) {
    const int g = g0 + i;
    const int bin = queue_bin(reg.size[g], region_class_words(reg, g, logW), coop_min);
    queue[atomicAdd(&qmeta[bin], 1)] = g;
  }
}

// One CTA per class >= 1: chunk table (queue offset, regions in the chunk).
__global__ void kq_chunks(const int *qbins, int32_t *chunk_start, int32_t *chunk_cnt) {
  const int cls = 1 + blockIdx.x;
  const int start = qbins[cls], cnt = qbins[Q_NCLS + cls], c0 = qbins[2 * Q_NCLS + cls];
  const int cs = class_chunk_size(cls);
  const int nch = (cnt + cs - 1) / cs;
  for (int k = threadIdx.x; k < nch; k += blockDim.x) {
    chunk_start[c0 + k] = start + k * cs;
    chunk_cnt[c0 + k] = min(cs, cnt - k * cs);
  }
}

}  // namespace rbepwt
