// K1, thread-per-region form: every LANE of a warp owns one region and walks its greedy path
// alone; a warp walks up to 32 regions at once.  This is the kernel the benchmark configurations
// spend their path time in (paths.cuh keeps the warp-per-region form for huge regions and EPWT).
//
// Same step rule as paths.cuh (Region.easy_path, /root/reference/rbepwt.py:1273-1347): smallest
// square window of half-width 1,2,4,... holding an unvisited point, then the lexicographic key
// (-dist, sp1, sp2).  What changes is the execution shape.  A path step is a short dependent chain
// (a few bitmap words, a handful of candidates); spreading ONE step over 32 lanes leaves most lanes
// idle and pays warp reductions per step (measured: 326 warp instructions per step, 11.5 of 32
// lanes active), while a batch holds ~10^5..10^6 independent regions.  So the parallel axis is the
// region, not the window row.
//
// Lane state machine.  Lanes need windows of different size at the same time; if each lane ran its
// whole search before the warp moved on, the warp would wait for the widest window at every step.
// Instead one trip of the warp loop lets every lane examine ONE bitmap word of its current window;
// a lane that exhausts its window commits the step (or doubles the window) and starts the next
// search on the following trip, independently of its neighbours.
//
// Window guess.  The reference probes half-widths 1,2,4,... in turn.  Scanning the window of
// half-width R once and ranking candidates by (k, dist, ...) with k = ceil(log2(Chebyshev distance))
// -- the index of the first probe that would have contained the candidate -- gives the same answer
// as the sequence of probes up to R.  Each search starts at the R that resolved the previous step
// and doubles only if that window is empty.
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
// Branch-free candidate selection (euclid mode).  In one bitmap row only the nearest unvisited point
// on each side of the current column can win (both k and d2 grow with |dj|), so a word contributes at
// most two candidates, found with clz/ffs, and a candidate is one 64-bit key
// (k << 44 | d2 << 22 | 2^21 - dot) -- valid because the kernel only takes regions whose bounding box
// has sides <= TPR_MAX_SIDE = 1024 (d2, |dot| < 2^21).  All lanes execute the same instructions.
//
// Unit-step fast path.  When the window half-width is 1 and pref is one of the 8 unit steps (the
// common case at level 1: 88 % of the steps), the 3x3 neighbourhood is gathered into a 9-bit mask and
// the answer is read from a 9 x 512 table in shared memory.  The table is filled at kernel start by
// the same candidate code the generic path runs, so it cannot disagree with it.
//
// Shared memory: one arena of TPR_ARENA_WORDS words per warp holds the bounding-box bitmaps of the
// chunk's regions (chunk table: regions.cuh; a chunk always fits).
#pragma once
#include "paths.cuh"

namespace rbepwt {

constexpr int TPR_WARPS = 4;

__device__ __forceinline__ bool pref_ties_exactly(int p0, int p1) {
  if (p0 == 0 || p1 == 0) return true;
  const int a = abs(p0), b = abs(p1);
  return a == b && (a & (a - 1)) == 0;
}

__device__ __forceinline__ int probe_index(int c) { return 32 - __clz(max(c - 1, 0)); }  // ceil(log2(c)), c >= 1

template <int MODE>
struct Search;

// ---- euclid: packed integer keys, fp64 only for mirror pairs under a non-exact pref ----------------
template <>
struct Search<MODE_EUCLID> {
  unsigned long long best;
  int di, dj, adi, adj;
  bool alt;

  __device__ __forceinline__ void reset() { best = ~0ull; alt = false; di = dj = adi = adj = 0; }
  __device__ __forceinline__ bool have() const { return best != ~0ull; }

  __device__ __forceinline__ void consider(bool valid, int cdi, int cdj, int p0, int p1) {
    const int k = probe_index(max(abs(cdi), abs(cdj)));
    const int d2 = cdi * cdi + cdj * cdj, dot = cdi * p0 + cdj * p1;
    const unsigned long long key =
        valid ? ((unsigned long long)k << 44) | ((unsigned long long)d2 << 22) | (unsigned)((1 << 21) - dot) : ~0ull;
    // selects, not branches: every lane executes the same instructions
    const bool lt = key < best;
    const bool eq = valid && key == best;  // mirror image of the incumbent about pref
    best = lt ? key : best;
    di = lt ? cdi : di;
    dj = lt ? cdj : dj;
    alt = lt ? false : (alt || eq);
    adi = eq ? cdi : adi;
    adj = eq ? cdj : adj;
  }

  // one bitmap word of row ci+rdi: columns lo..lo+31, already masked to the window
  __device__ __forceinline__ void scan_word(uint32_t bits, int lo, int rdi, int cj, int p0, int p1) {
    const int rel = min(cj - lo, 31);
    const uint32_t lmask = rel < 0 ? 0u : (2u << rel) - 1u;  // columns <= cj
    const uint32_t left = bits & lmask, right = bits & ~lmask;
    consider(left != 0u, rdi, lo + 31 - __clz(left) - cj, p0, p1);
    consider(right != 0u, rdi, lo + __ffs(right) - 1 - cj, p0, p1);
  }

  __device__ __forceinline__ void finish(int p0, int p1, int &odi, int &odj, int &k) {
    if (alt) {
      const int cb = di * p1 - dj * p0, ca = adi * p1 - adj * p0;
      bool alt_better;
      if (pref_ties_exactly(p0, p1)) {
        alt_better = ca > cb;
      } else {
        const int d2 = (int)((best >> 22) & 0x3fffffu);
        const double sb = tie_sp1(di, dj, d2, p0, p1), sa = tie_sp1(adi, adj, d2, p0, p1);
        alt_better = sa != sb ? sa > sb : ca > cb;
      }
      if (alt_better) { di = adi; dj = adj; }
    }
    odi = di; odj = dj; k = (int)(best >> 44);
  }
};

// ---- chebyshev: every point of the nearest ring competes through the fp64 sp1 ----------------------
template <>
struct Search<MODE_CHEB> {
  int c, d2, di, dj;
  double sp1;
  bool found, has_sp1;

  __device__ __forceinline__ void reset() { found = false; has_sp1 = false; c = d2 = di = dj = 0; sp1 = 0.0; }
  __device__ __forceinline__ bool have() const { return found; }

  __device__ __forceinline__ void consider(bool valid, int cdi, int cdj, int p0, int p1) {
    if (!valid) return;
    const int cc = max(abs(cdi), abs(cdj)), cd2 = cdi * cdi + cdj * cdj;
    if (found && cc > c) return;
    if (!found || cc < c) {
      found = true; has_sp1 = false; c = cc; d2 = cd2; di = cdi; dj = cdj;
      return;
    }
    if (!has_sp1) { sp1 = tie_sp1(di, dj, d2, p0, p1); has_sp1 = true; }
    const double s = tie_sp1(cdi, cdj, cd2, p0, p1);
    const bool better = s != sp1 ? s > sp1 : (cdi * p1 - cdj * p0) > (di * p1 - dj * p0);
    if (better) { sp1 = s; d2 = cd2; di = cdi; dj = cdj; }
  }

  __device__ __forceinline__ void scan_word(uint32_t bits, int lo, int rdi, int cj, int p0, int p1) {
    while (bits) {
      const int j = lo + __ffs(bits) - 1;
      bits &= bits - 1;
      consider(true, rdi, j - cj, p0, p1);
    }
  }

  __device__ __forceinline__ void finish(int, int, int &odi, int &odj, int &k) {
    odi = di; odj = dj; k = probe_index(c);
  }
};

constexpr int TPR_LUT_ROWS = 9, TPR_LUT_COLS = 512;

template <int MODE>
__global__ void __launch_bounds__(TPR_WARPS * 32) k1_paths_tpr(PathParams P) {
  __shared__ uint32_t s_arena[TPR_WARPS][TPR_ARENA_WORDS];
  __shared__ uint8_t s_lut[TPR_LUT_ROWS * TPR_LUT_COLS];
  const int lane = (int)lane_id(), warp = threadIdx.x >> 5;
  uint32_t *arena = s_arena[warp];
  const int nbig = P.qmeta[QM_NBIG], nchunks = P.qmeta[QM_NCHUNKS];
  const int logW = P.logW, W = P.W, N = P.N, L = P.levels;
  const int Wm = W - 1;
  (void)nbig;

  // unit-step table: s_lut[q * 512 + m] = index (di+1)*3 + (dj+1) of the winner among the neighbours
  // present in the 9-bit mask m, for pref = (q/3 - 1, q%3 - 1)
  for (int e = threadIdx.x; e < TPR_LUT_ROWS * TPR_LUT_COLS; e += blockDim.x) {
    const int q = e / TPR_LUT_COLS, m = e % TPR_LUT_COLS;
    const int p0 = q / 3 - 1, p1 = q % 3 - 1;
    uint8_t v = 0xff;
    if (q != 4 && !(m & 16) && m) {
      Search<MODE> S;
      S.reset();
      for (int bpos = 0; bpos < 9; bpos++)
        if (m & (1 << bpos)) S.consider(true, bpos / 3 - 1, bpos % 3 - 1, p0, p1);
      int odi, odj, k;
      S.finish(p0, p1, odi, odj, k);
      v = (uint8_t)((odi + 1) * 3 + (odj + 1));
    }
    s_lut[e] = v;
  }
  __syncthreads();

  while (true) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(&P.qmeta[QM_CUR_SMALL], 1);
    chunk = __shfl_sync(FULL_MASK, chunk, 0);
    if (chunk >= nchunks) break;
    const int qstart = P.chunk_start[chunk], cnt = P.chunk_cnt[chunk];
    const bool mine = lane < cnt;
    int img = 0, label = 0, first = 0, size = 0, off = 0, r0 = 0, c0 = 0, h = 0, w = 0, ws = 0;
    if (mine) {
      const int g = P.queue[qstart + lane];
      img = P.reg.img[g]; label = P.reg.label[g]; first = P.reg.first[g];
      size = P.reg.size[g]; off = P.reg.off[g];
      r0 = first >> logW; c0 = P.reg.cmin[g];
      h = P.reg.rmax[g] - r0 + 1; w = P.reg.cmax[g] - c0 + 1; ws = (w + 31) >> 5;
    }
    const int words = h * ws;  // 0 for idle lanes
    int inc = words;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(FULL_MASK, inc, d);
      if (lane >= d) inc += y;
    }
    const int base = inc - words;
    uint32_t *bm = arena + base;

    // cooperative bitmap build: one ballot per bitmap word, lanes = columns
    for (int r = 0; r < cnt; r++) {
      const int img_r = __shfl_sync(FULL_MASK, img, r), label_r = __shfl_sync(FULL_MASK, label, r);
      const int r0_r = __shfl_sync(FULL_MASK, r0, r), c0_r = __shfl_sync(FULL_MASK, c0, r);
      const int h_r = __shfl_sync(FULL_MASK, h, r), w_r = __shfl_sync(FULL_MASK, w, r);
      const int ws_r = __shfl_sync(FULL_MASK, ws, r), base_r = __shfl_sync(FULL_MASK, base, r);
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

    // every lane walks its own region; the warp advances level by level
    bool live = mine;
    int n = mine ? size : 0, a = off;
    int si = 0, sj = (first & Wm) - c0;
    int32_t *Qimg = P.Q + (size_t)img * 2 * (size_t)N;
    for (int lev = 1; lev <= L; lev++) {
      int32_t *Ql = Qimg + level_off((size_t)N, lev) + a;
      int t = n, ci = si, cj = sj, p0 = 0, p1 = 1;  // prefered_direc = (0,1)   rbepwt.py:1290
      int rad = 1, i = 0, wd = 0, i1 = 0, j0 = 0, j1 = 0, w0 = 0, w1 = 0;
      bool fresh = true;  // at the first word of a window
      Search<MODE> S;
      S.reset();
#define TPR_SET_WINDOW()                                            \
  do {                                                              \
    i = max(ci - rad, 0); i1 = min(ci + rad, h - 1);                \
    j0 = max(cj - rad, 0); j1 = min(cj + rad, w - 1);               \
    w0 = j0 >> 5; w1 = j1 >> 5; wd = w0; fresh = true;              \
  } while (0)
      if (live) {
        bm[si * ws + (sj >> 5)] &= ~(1u << (sj & 31));
        Ql[0] = ((r0 + si) << logW) + c0 + sj;
        t = 1;
        TPR_SET_WINDOW();
      }
      while (__any_sync(FULL_MASK, t < n)) {
        if (t < n) {
          bool commit = false, expand = false;
          int fdi = 0, fdj = 0, fk = 0;
          if (fresh && rad == 1 && (unsigned)(p0 + 1) <= 2u && (unsigned)(p1 + 1) <= 2u) {
            // unit-step fast path: 3x3 neighbourhood -> 9-bit mask -> table
            const int wq = cj >> 5, bq = cj & 31;
            unsigned m = 0;
#pragma unroll
            for (int rr = 0; rr < 3; rr++) {
              const int ri = ci + rr - 1;
              unsigned three = 0;
              if (ri >= 0 && ri < h) {
                const uint32_t *row = bm + ri * ws;
                const uint32_t x = row[wq];
                if (bq == 0) three = ((x << 1) | (wq > 0 ? row[wq - 1] >> 31 : 0u)) & 7u;
                else if (bq == 31) three = ((x >> 30) | ((wq + 1 < ws ? row[wq + 1] : 0u) << 2)) & 7u;
                else three = (x >> (bq - 1)) & 7u;
              }
              m |= three << (3 * rr);
            }
            if (m) {
              const int idx = s_lut[((p0 + 1) * 3 + (p1 + 1)) * TPR_LUT_COLS + m];
              fdi = idx / 3 - 1; fdj = idx % 3 - 1; fk = 0;
              commit = true;
            } else {
              expand = true;
            }
          } else {
            fresh = false;
            uint32_t bits = bm[i * ws + wd];
            const int lo = wd << 5;
            if (lo < j0) bits &= 0xffffffffu << (j0 - lo);
            if (lo + 31 > j1) bits &= 0xffffffffu >> (lo + 31 - j1);
            S.scan_word(bits, lo, i - ci, cj, p0, p1);
            const bool more_w = wd < w1, more_i = i < i1;
            wd = more_w ? wd + 1 : w0;
            i += (!more_w && more_i) ? 1 : 0;
            if (!more_w && !more_i) {  // window exhausted
              if (S.have()) {
                S.finish(p0, p1, fdi, fdj, fk);
                commit = true;
              } else {
                expand = true;
              }
            }
          }
          if (commit || expand) {
            if (commit) {
              const int bi = ci + fdi, bj = cj + fdj;
              bm[bi * ws + (bj >> 5)] &= ~(1u << (bj & 31));
              Ql[t] = ((r0 + bi) << logW) + c0 + bj;
              p0 = fdi; p1 = fdj;  // rbepwt.py:1331
              ci = bi; cj = bj;
              t++;
              rad = 1 << fk;
              S.reset();
            } else if (ci - rad <= 0 && cj - rad <= 0 && ci + rad >= h - 1 && cj + rad >= w - 1) {
              atomicExch(&P.qmeta[QM_ERR], 1);  // nothing unvisited in the whole box: corrupt state
              t = n; live = false;
            } else {
              rad <<= 1;
            }
            TPR_SET_WINDOW();
          }
        }
      }
#undef TPR_SET_WINDOW
      if (lev == L) break;
      // RegionCollection.reduce: the points at even GLOBAL position a+t survive (rbepwt.py:1563-1584);
      // the bitmap is all-zero here, re-mark them; the smallest surviving pixel id is the next start
      // point (lexicographic min, rbepwt.py:1035-1036)
      int minpix = INT32_MAX;
      if (live) {
        for (int tt = a & 1; tt < n; tt += 2) {
          const int pix = __ldcg(Ql + tt);
          const int pi = (pix >> logW) - r0, pj = (pix & Wm) - c0;
          bm[pi * ws + (pj >> 5)] |= 1u << (pj & 31);
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
  }
}

}  // namespace rbepwt
