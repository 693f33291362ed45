// K1, thread-per-region variant: every LANE of a warp owns one region and walks its greedy path
// alone, 32 regions in lock step per warp.
//
// Same step rule as paths.cuh (Region.easy_path, /root/reference/rbepwt.py:1273-1347): smallest
// square window of half-width 1,2,4,... holding an unvisited point, then the lexicographic key
// (-dist, sp1, sp2).  What changes is the execution shape.  A path step is a short dependent chain
// (a few bitmap words, a handful of candidates); spreading ONE step over 32 lanes leaves most lanes
// idle and pays warp reductions per step, while a batch holds ~10^5..10^6 independent regions.  So
// the parallel axis is the region, not the window row.
//
// Integer tie-break (euclid mode).  Candidates compared by sp1 always have the same d2 = di^2+dj^2,
// hence the same norm n, and sp1 = fl(fl(dj/n)*p1 + fl(fl(di/n)*p0)) orders them like the integer
// dot product di*p0 + dj*p1 whenever the dot products differ: the true values differ by >= 1/n while
// the accumulated rounding error is < 2^-20/n for coordinates below 2^15.  Equal dot products mean
// the two candidates are mirror images about pref; then the reference's fp64 expression either ties
// exactly (-> sp2, i.e. the integer cross product, decides) or differs in the last bit.  It ties
// exactly when every product is exact: pref on an axis, or |p0| == |p1| a power of two (all unit
// steps).  Only for the remaining prefs (after jumps) is the fp64 expression evaluated, bit for bit
// as in paths.cuh.  Chebyshev mode compares candidates of different norms and always uses fp64 sp1.
//
// Shared memory: one arena of TPR_ARENA_WORDS words per warp holds the bounding-box bitmaps of the
// regions being walked; a warp claims 32 consecutive queue entries (the queue is sorted by size, so
// the 32 chains have similar length) and walks them in as many rounds as the arena requires.
#pragma once
#include "paths.cuh"

namespace rbepwt {

constexpr int TPR_WARPS = 4;
constexpr int TPR_ARENA_WORDS = K1_SLOT_WORDS;  // any region of the "small" queue class fits alone

__device__ __forceinline__ bool pref_ties_exactly(int p0, int p1) {
  if (p0 == 0 || p1 == 0) return true;
  const int a = abs(p0), b = abs(p1);
  return a == b && (a & (a - 1)) == 0;
}

struct LaneBest {
  int dist;            // d2 (euclid) or Chebyshev distance
  int dot, d2, di, dj;
  int adi, adj;        // mirror partner with the same (dist, dot), euclid only
  double sp1;          // chebyshev: lazily computed
  bool have, alt, has_sp1;
};

template <int MODE>
__device__ __forceinline__ void lane_consider(LaneBest &b, int di, int dj, int p0, int p1) {
  const int d2 = di * di + dj * dj;
  if (MODE == MODE_EUCLID) {
    if (b.have && d2 > b.dist) return;
    const int dot = di * p0 + dj * p1;
    if (!b.have || d2 < b.dist || dot > b.dot) {
      b.have = true; b.alt = false; b.dist = d2; b.dot = dot; b.di = di; b.dj = dj;
    } else if (dot == b.dot) {
      b.alt = true; b.adi = di; b.adj = dj;
    }
  } else {
    const int dist = max(abs(di), abs(dj));
    if (b.have && dist > b.dist) return;
    if (!b.have || dist < b.dist) {
      b.have = true; b.has_sp1 = false; b.dist = dist; b.d2 = d2; b.di = di; b.dj = dj;
      return;
    }
    if (!b.has_sp1) { b.sp1 = tie_sp1(b.di, b.dj, b.d2, p0, p1); b.has_sp1 = true; }
    const double sp1 = tie_sp1(di, dj, d2, p0, p1);
    bool better;
    if (sp1 != b.sp1) better = sp1 > b.sp1;
    else better = (di * p1 - dj * p0) > (b.di * p1 - b.dj * p0);
    if (better) { b.sp1 = sp1; b.d2 = d2; b.di = di; b.dj = dj; }
  }
}

// One lane, one step: the next point of the path.  bm: h rows of ws words, bit set <=> unvisited.
template <int MODE>
__device__ __forceinline__ bool lane_find_next(const uint32_t *bm, int h, int w, int ws, int ci, int cj, int p0,
                                               int p1, int &bi, int &bj) {
  LaneBest b;
  b.have = false; b.alt = false; b.has_sp1 = false;
  b.dist = 0; b.dot = 0; b.d2 = 0; b.di = 0; b.dj = 0; b.adi = 0; b.adj = 0; b.sp1 = 0.0;
  for (int rad = 1;; rad <<= 1) {
    const int i0 = max(ci - rad, 0), i1 = min(ci + rad, h - 1);
    const int j0 = max(cj - rad, 0), j1 = min(cj + rad, w - 1);
    const int w0 = j0 >> 5, w1 = j1 >> 5;
    for (int i = i0; i <= i1; i++) {
      const uint32_t *row = bm + i * ws;
      for (int wd = w0; wd <= w1; wd++) {
        uint32_t bits = row[wd];
        const int lo = wd << 5;
        if (lo < j0) bits &= 0xffffffffu << (j0 - lo);
        if (lo + 31 > j1) bits &= 0xffffffffu >> (lo + 31 - j1);
        while (bits) {
          const int j = lo + __ffs(bits) - 1;
          bits &= bits - 1;
          lane_consider<MODE>(b, i - ci, j - cj, p0, p1);
        }
      }
    }
    if (b.have) break;
    if (i0 == 0 && j0 == 0 && i1 == h - 1 && j1 == w - 1) return false;
  }
  if (MODE == MODE_EUCLID && b.alt) {  // mirror pair: sp1 ties mathematically
    const int cb = b.di * p1 - b.dj * p0, ca = b.adi * p1 - b.adj * p0;
    bool alt_better;
    if (pref_ties_exactly(p0, p1)) {
      alt_better = ca > cb;
    } else {
      const double sb = tie_sp1(b.di, b.dj, b.dist, p0, p1), sa = tie_sp1(b.adi, b.adj, b.dist, p0, p1);
      alt_better = sa != sb ? sa > sb : ca > cb;
    }
    if (alt_better) { b.di = b.adi; b.dj = b.adj; }
  }
  bi = ci + b.di;
  bj = cj + b.dj;
  return true;
}

template <int MODE>
__global__ void __launch_bounds__(TPR_WARPS * 32) k1_paths_tpr(PathParams P) {
  __shared__ uint32_t s_arena[TPR_WARPS][TPR_ARENA_WORDS];
  const int lane = (int)lane_id(), warp = threadIdx.x >> 5;
  uint32_t *arena = s_arena[warp];
  const int nbig = P.qmeta[QM_NBIG], nsmall = P.qmeta[QM_NREG] - nbig;
  const int nchunks = (nsmall + 31) >> 5;
  const int logW = P.logW, W = P.W, N = P.N, L = P.levels;
  const int Wm = W - 1;

  while (true) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(&P.qmeta[QM_CUR_SMALL], 1);
    chunk = __shfl_sync(FULL_MASK, chunk, 0);
    if (chunk >= nchunks) break;
    const int idx = (chunk << 5) + lane;
    const bool valid = idx < nsmall;
    int img = 0, label = 0, first = 0, size = 0, off = 0, r0 = 0, c0 = 0, h = 0, w = 0, ws = 0;
    if (valid) {
      const int g = P.queue[nbig + idx];
      img = P.reg.img[g]; label = P.reg.label[g]; first = P.reg.first[g];
      size = P.reg.size[g]; off = P.reg.off[g];
      r0 = first >> logW; c0 = P.reg.cmin[g];
      h = P.reg.rmax[g] - r0 + 1; w = P.reg.cmax[g] - c0 + 1; ws = (w + 31) >> 5;
    }
    const int words = h * ws;  // 0 for invalid lanes

    for (int start = 0; start < 32;) {
      // lanes [start, start+cnt): the longest run whose bitmaps fit the arena together
      int inc = lane >= start ? words : 0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= d) inc += y;
      }
      const bool mine = valid && lane >= start && inc <= TPR_ARENA_WORDS;
      const unsigned round = __ballot_sync(FULL_MASK, mine);
      const int cnt = __popc(round);
      if (cnt == 0) break;
      uint32_t *bm = arena + (inc - words);

      // cooperative bitmap build: one ballot per bitmap word, lanes = columns
      for (int r = start; r < start + cnt; r++) {
        const int img_r = __shfl_sync(FULL_MASK, img, r), label_r = __shfl_sync(FULL_MASK, label, r);
        const int r0_r = __shfl_sync(FULL_MASK, r0, r), c0_r = __shfl_sync(FULL_MASK, c0, r);
        const int h_r = __shfl_sync(FULL_MASK, h, r), w_r = __shfl_sync(FULL_MASK, w, r);
        const int ws_r = __shfl_sync(FULL_MASK, ws, r), base_r = __shfl_sync(FULL_MASK, inc - words, r);
        const int32_t *lab = P.labels + (size_t)img_r * N;
        for (int i = 0; i < h_r; i++)
          for (int wd = 0; wd < ws_r; wd++) {
            const int col = (wd << 5) + lane;
            const bool in = col < w_r && lab[((r0_r + i) << logW) + c0_r + col] == label_r;
            const unsigned bits = __ballot_sync(FULL_MASK, in);
            if (lane == 0) arena[base_r + i * ws_r + wd] = bits;
          }
      }
      __syncwarp();

      // every lane walks its own region through all levels; the warp advances level by level
      bool live = mine;
      int n = size, a = off;
      int si = 0, sj = (first & Wm) - c0;
      int32_t *Qimg = P.Q + (size_t)img * 2 * (size_t)N;
      for (int lev = 1; lev <= L; lev++) {
        int32_t *Ql = Qimg + level_off((size_t)N, lev) + a;
        int t = n, ci = si, cj = sj, p0 = 0, p1 = 1;  // prefered_direc = (0,1)   rbepwt.py:1290
        if (live) {
          bm[si * ws + (sj >> 5)] &= ~(1u << (sj & 31));
          Ql[0] = ((r0 + si) << logW) + c0 + sj;
          t = 1;
        }
        while (__any_sync(FULL_MASK, t < n)) {
          if (t < n) {
            int bi, bj;
            if (lane_find_next<MODE>(bm, h, w, ws, ci, cj, p0, p1, bi, bj)) {
              bm[bi * ws + (bj >> 5)] &= ~(1u << (bj & 31));
              Ql[t] = ((r0 + bi) << logW) + c0 + bj;
              p0 = bi - ci; p1 = bj - cj;  // rbepwt.py:1331
              ci = bi; cj = bj;
              t++;
            } else {
              atomicExch(&P.qmeta[QM_ERR], 1);
              t = n; live = false;
            }
          }
        }
        if (lev == L) break;
        // RegionCollection.reduce: the points at even GLOBAL position a+t survive (rbepwt.py:1563-1584);
        // the bitmap is all-zero here, re-mark them; the smallest surviving pixel id is the next start
        // point (lexicographic min, rbepwt.py:1035-1036)
        int minpix = INT32_MAX;
        if (live) {
          for (int tt = a & 1; tt < n; tt += 2) {
            const int pix = __ldcg(Ql + tt);
            const int i = (pix >> logW) - r0, j = (pix & Wm) - c0;
            bm[i * ws + (j >> 5)] |= 1u << (j & 31);
            minpix = min(minpix, pix);
          }
        }
        const int na = (a + 1) >> 1, nb = (a + n + 1) >> 1;
        a = na; n = nb - na;
        live = live && n > 0;
        if (!live) n = 0;
        if (live) { si = (minpix >> logW) - r0; sj = (minpix & Wm) - c0; }
        if (!__any_sync(FULL_MASK, live)) break;
      }
      __syncwarp();
      start += cnt;
    }
  }
}

}  // namespace rbepwt
