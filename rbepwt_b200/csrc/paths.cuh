// K1: easy-path construction.
//
// Replaces Region.easy_path (/root/reference/rbepwt.py:1273-1347, helpers neighborhood 84-104 and
// rotate 78-82), the start-point rule of Region.__init_dict_and_extreme_values__ (1020-1036) and
// the point bookkeeping of RegionCollection.reduce / Region.reduce_points (1563-1584, 1349-1375).
//
// This file: the warp-per-region searches -- one warp owns one region, the lanes share the rows of the search window
// (find_next_geo: euclid, integer keys; find_next: chebyshev and EPWT, the reference's fp64 expressions;
// find_next_grad: gradpath) -- and the walker built on them (run_path / region_pyramid), which now serves gradpath and
// the chebyshev mode's oversized regions.  The euclidean mode's whole-warp walker (oversized regions, the largest
// regions of a group, every region of a small group) and EPWT keep the bitmap window in registers (regwin.cuh) and
// come here only for the searches beyond half-width 8 (euclid) / 2 (EPWT).  The bulk -- hundreds of thousands of
// small regions per batch -- is walked thread per region by walk.cuh.  The unvisited points are a bitmap over the
// region's bounding box (shared memory; global scratch for boxes too large).
//
// Step rule (exactly the reference's, restated order-independently):
//   candidates = unvisited points of the region inside the smallest square of half-width
//                r = 1,2,4,8,... around the current point that contains any;
//   choose the candidate maximising the lexicographic key (-dist, sp1, sp2) where
//     dist = di^2+dj^2 (euclid; the reference compares sqrt of it), max(|di|,|dj|) (chebyshev) or
//            |value[cur]-value[cand]| (EPWT),
//     sp1  = fma(v1, p1, v0*p0)  with v = (di,dj)/sqrt(di^2+dj^2) in IEEE fp64 -- this is what
//            np.dot(v, prefered_direc) evaluates to (OpenBLAS ddot tail),
//     sp2  = v . rotate(pref, -pi/2), consulted only when sp1 ties exactly, in which case the two
//            candidates are mirror images about pref and sp2 orders like the integer cross
//            product di*p1 - dj*p0;
//   pref = (0,1) at the start of a region, afterwards chosen - current (integer, not normalised).
//   A complete tie (possible only in EPWT mode: collinear candidates with bit-identical |dv|)
//   depends on CPython set order in the reference (unpinned); here the nearer point wins.
//
// For the geometric modes paths never read pixel values, and the region offsets of every level
// follow from the level-1 sizes alone (a region occupying [a, a+n) keeps its even global
// positions: [ceil(a/2), ceil((a+n)/2)) at the next level), so one warp builds the region's
// whole path pyramid, levels 1..L, without any inter-region synchronisation.
//
// Output: Q[level][a + t] = pixel id (row*W+col) of the t-th path point, and for levels >= 2
// Pm[level][a + t] = its place in the level's incoming order (= the reference's generating permutation
// plus the region offset), which is what the transform kernels gather / scatter through (dwt.cuh).
// `posmap` (pixel -> place in the next level's incoming order) links a level to the next.
#pragma once
#include "common.cuh"
#include "regions.cuh"

namespace rbepwt {

constexpr int MODE_EUCLID = 0, MODE_CHEB = 1, MODE_EPWT = 2;
constexpr int MODE_GRAD_EUCLID = 3, MODE_GRAD_CHEB = 4;  // path_type='gradpath' (Region.grad_path, rbepwt.py:1190-1271)
__host__ __device__ constexpr bool mode_is_grad(int mode) { return mode == MODE_GRAD_EUCLID || mode == MODE_GRAD_CHEB; }

struct PathParams {
  const int32_t *labels;  // [B][N]
  const double *img;      // [B][N] pixel values: gradpath only (the regions' average gradients)
  int H, W, logW, N, levels;
  RegionArrays reg;
  const int32_t *queue;
  const int32_t *chunk_start, *chunk_cnt;  // chunk table of the thread-per-region kernel (regions.cuh)
  int *qmeta;
  const uint8_t *unit_lut;  // 9 x 512 unit-step table of the path mode (global memory, built once per context)
  const uint8_t *t2_tab;    // 5x5 step table of the euclid mode (walk.cuh, T2_BYTES)
  int32_t *slot_of;         // [nreg] word offset of a region's slot in gbm, -1: its chunk builds its own (kq_slots)
  int g0, nreg;             // the path group's regions: [g0, g0 + nreg)
  int coop;                 // the path mode has a whole-warp kernel for class 1 (k1_coop_all): euclid, gradpath
  uint32_t *gbm;            // [gbm_chunks][TPR_ARENA_WORDS] chunk arena images built by k1_bitmaps (walk.cuh)
  int gbm_chunks;           // chunks beyond it build their bitmaps inside the path kernel
  int32_t *Q;  // [B][2N]
  int32_t *Pm;      // [B][2N] level l >= 2: position of the path point in the level's incoming order (= index into cA of level l-1)
  int32_t *posmap;  // [B][N] scratch: pixel -> position in the next level's incoming order
  // big-region kernel
  uint32_t *gscratch;          // global bitmap scratch, one slab per CTA (used when smem is too small)
  size_t gscratch_words;       // words per slab
  int big_smem_words;          // dynamic shared-memory words available per CTA in the big kernel
};

constexpr int TPR_LUT_ROWS = 9, TPR_LUT_COLS = 512;  // unit-step table: 9 prefs x 512 neighbourhood masks (walk.cuh)
constexpr int COOP_MAX_SIDE = 16384;                 // find_next_geo: integer dot products stay below 2^30

__host__ __device__ __forceinline__ bool pref_ties_exactly(int p0, int p1) {
  if (p0 == 0 || p1 == 0) return true;
  const int a = abs(p0), b = abs(p1);
  return a == b && (a & (a - 1)) == 0;
}

__host__ __device__ __forceinline__ int probe_index(int c) { return 32 - rb_clz((uint32_t)max(c - 1, 0)); }  // ceil(log2(c)), c >= 1

// Window row as one word: bit 15 + dj  <->  column cj + dj, dj in [-15, 16]; columns outside the
// bitmap read as 0.
__host__ __device__ __forceinline__ uint32_t row_window(const uint32_t *row, int ws, int cj) {
  const int s = cj - 15;
  const int wlo = s >> 5;  // -1 when s < 0
  const uint32_t lo = (wlo >= 0 && wlo < ws) ? row[wlo] : 0u;
  const uint32_t hi = (wlo + 1 < ws) ? row[wlo + 1] : 0u;
  return rb_funnel_r(lo, hi, s & 31);
}

struct Best {
  double dist;  // integer-valued in the geometric modes
  double val;   // EPWT: the candidate's value (becomes the current value of the next step)
  double sp1;
  int cross, d2, i, j;
  bool have, has_sp1;
};

// sp1 of the reference's tie-break, bit for bit.
__host__ __device__ __forceinline__ double tie_sp1(int di, int dj, int d2, int p0, int p1) {
  const double nrm = sqrt((double)d2);  // IEEE-correct in fp64
  const double v0 = (double)di / nrm, v1 = (double)dj / nrm;
  double s = fma(v1, (double)p1, rb_dmul(v0, (double)p0));
  if (s == 0.0) s = 0.0;  // -0.0 -> +0.0 (compares equal in the reference)
  return s;
}

// EPWT value loads: the values (the image at level 1, the previous level's approximation laid out by pixel after) are
// read-only for the whole launch and the walk is local -- the next point is a neighbour -- so they go through L1.
#ifndef RB_EPWT_LOAD
#define RB_EPWT_LOAD(p) __ldg(p)
#endif

template <int MODE>
__device__ __forceinline__ void consider(Best &b, int i, int j, int ci, int cj, int p0, int p1, double curval,
                                         const double *__restrict__ vals, int pix, bool u8wrap) {
  const int di = i - ci, dj = j - cj;
  const int d2 = di * di + dj * dj;
  double dist, val = 0.0;
  if (MODE == MODE_EUCLID) {
    dist = (double)d2;
  } else if (MODE == MODE_CHEB) {
    dist = (double)max(abs(di), abs(dj));
  } else {
    val = RB_EPWT_LOAD(vals + pix);
    const double dv = curval - val;
    dist = u8wrap ? (dv < 0.0 ? dv + 256.0 : dv) : fabs(dv);
  }
  if (b.have && dist > b.dist) return;
  const int cross = di * p1 - dj * p0;
  if (!b.have || dist < b.dist) {
    b.have = true; b.has_sp1 = false;
    b.dist = dist; b.val = val; b.cross = cross; b.d2 = d2; b.i = i; b.j = j;
    return;
  }
  // equal dist: direction tie-break
  if (!b.has_sp1) {
    b.sp1 = tie_sp1(b.i - ci, b.j - cj, b.d2, p0, p1);
    b.has_sp1 = true;
  }
  const double sp1 = tie_sp1(di, dj, d2, p0, p1);
  bool better;
  if (sp1 != b.sp1) better = sp1 > b.sp1;
  else if (cross != b.cross) better = cross > b.cross;
  else better = d2 < b.d2;
  if (better) { b.sp1 = sp1; b.val = val; b.cross = cross; b.d2 = d2; b.i = i; b.j = j; }
}

// Warp-cooperative search for the next path point.  bm: h x ws words, bit (i,j) set <=> unvisited.
// Returns false if no unvisited point exists in the whole bounding box (corrupt state).
// curval (EPWT): in = value at the current point, out = value at the chosen point.
template <int MODE>
__device__ __forceinline__ void warp_arg_best(Best &b, int ci, int cj, int p0, int p1, double &curval, int &bi, int &bj);

// start_rad > 0: the probes below that half-width are known to be empty (regwin.cuh)
template <int MODE>
__device__ __forceinline__ bool find_next(const uint32_t *bm, int h, int w, int ws, int ci, int cj, int p0, int p1,
                                          const double *__restrict__ vals, int r0, int c0, int logW, bool u8wrap,
                                          double &curval, int &bi, int &bj, int start_rad = 0) {
  const int lane = (int)lane_id();
  Best b;
  b.have = false; b.has_sp1 = false; b.dist = 0.0; b.val = 0.0; b.sp1 = 0.0; b.cross = 0; b.d2 = 0; b.i = 0; b.j = 0;
  int rad0 = start_rad > 0 ? start_rad : 1;
  if (MODE == MODE_EPWT && start_rad <= 0) {
    // half-width 1, the common case of the one-region walk: one LANE per neighbour, so the eight value
    // loads (L2 latency each) are in flight together instead of one after the other inside a lane
    const int i = ci + lane / 3 - 1, j = cj + lane % 3 - 1;
    const bool cand = lane < 9 && lane != 4 && i >= 0 && i < h && j >= 0 && j < w && ((bm[i * ws + (j >> 5)] >> (j & 31)) & 1u);
    if (__any_sync(FULL_MASK, cand)) {
      if (cand) consider<MODE>(b, i, j, ci, cj, p0, p1, curval, vals, ((r0 + i) << logW) + c0 + j, u8wrap);
      rad0 = 0;  // found: skip the window loop
    } else {
      rad0 = 2;
    }
  }
  for (int rad = rad0; rad > 0; rad <<= 1) {  // half-width 2^(k-1), k = 1,2,...   rbepwt.py:1296-1299, 90-92
    const int i0 = max(ci - rad, 0), i1 = min(ci + rad, h - 1);
    const int j0 = max(cj - rad, 0), j1 = min(cj + rad, w - 1);
    const int w0 = j0 >> 5, w1 = j1 >> 5;
    for (int i = i0 + lane; i <= i1; i += 32) {
      for (int wd = w0; wd <= w1; wd++) {
        uint32_t bits = bm[i * ws + wd];
        const int lo = wd << 5;
        if (lo < j0) bits &= 0xffffffffu << (j0 - lo);
        if (lo + 31 > j1) bits &= 0xffffffffu >> (lo + 31 - j1);
        while (bits) {
          const int j = lo + __ffs(bits) - 1;
          bits &= bits - 1;
          consider<MODE>(b, i, j, ci, cj, p0, p1, curval, vals, ((r0 + i) << logW) + c0 + j, u8wrap);
        }
      }
    }
    if (__any_sync(FULL_MASK, b.have)) break;
    if (i0 == 0 && j0 == 0 && i1 == h - 1 && j1 == w - 1) return false;
  }
  warp_arg_best<MODE>(b, ci, cj, p0, p1, curval, bi, bj);
  return true;
}

// The warp's winner among the lanes' best candidates (at least one lane has one):
// cross-lane arg-best: (dist asc, sp1 desc, cross desc, d2 asc)
template <int MODE>
__device__ __forceinline__ void warp_arg_best(Best &b, int ci, int cj, int p0, int p1, double &curval, int &bi, int &bj) {
  const int lane = (int)lane_id();
  unsigned tied;
  if (MODE == MODE_EPWT) {
    const unsigned long long k = b.have ? (unsigned long long)__double_as_longlong(b.dist) : ~0ull;  // dist >= 0
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_min_sync(FULL_MASK, hi);
    const unsigned ml = __reduce_min_sync(FULL_MASK, hi == mh ? lo : 0xffffffffu);
    tied = __ballot_sync(FULL_MASK, b.have && hi == mh && lo == ml);
  } else {
    const int d = b.have ? (int)b.dist : INT32_MAX;
    const int dm = __reduce_min_sync(FULL_MASK, d);
    tied = __ballot_sync(FULL_MASK, b.have && d == dm);
  }
  if (__popc(tied) > 1) {
    const bool in = (tied >> lane) & 1u;
    if (in && !b.has_sp1) b.sp1 = tie_sp1(b.i - ci, b.j - cj, b.d2, p0, p1);
    const unsigned long long k = in ? orderable(b.sp1) : 0ull;
    const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
    const unsigned mh = __reduce_max_sync(FULL_MASK, hi);
    const unsigned ml = __reduce_max_sync(FULL_MASK, (in && hi == mh) ? lo : 0u);
    tied = __ballot_sync(FULL_MASK, in && hi == mh && lo == ml);
    if (__popc(tied) > 1) {
      const bool in2 = (tied >> lane) & 1u;
      const int mc = __reduce_max_sync(FULL_MASK, in2 ? b.cross : INT32_MIN);
      tied = __ballot_sync(FULL_MASK, in2 && b.cross == mc);
      if (__popc(tied) > 1) {
        const bool in3 = (tied >> lane) & 1u;
        const int md = __reduce_min_sync(FULL_MASK, in3 ? b.d2 : INT32_MAX);
        tied = __ballot_sync(FULL_MASK, in3 && b.d2 == md);
      }
    }
  }
  const int src = __ffs(tied) - 1;
  bi = __shfl_sync(FULL_MASK, b.i, src);
  bj = __shfl_sync(FULL_MASK, b.j, src);
  if (MODE == MODE_EPWT) curval = __shfl_sync(FULL_MASK, b.val, src);
}

// Warp-cooperative search, euclid mode, integer keys (the derivation is in walk.cuh): the lanes take the
// rows of the window, every lane keeps the best candidate of its rows as (k, d2, dot) with k = probe index
// ceil(log2(Chebyshev distance)), three warp reductions pick the winner, and a mirror pair (equal d2 and equal dot
// product) is settled by the cross product or -- only for a pref where the reference's fp64 expression need
// not tie exactly -- by that expression itself.  `rad` carries the half-width that resolved the previous step
// (the window guess); with it == 1 and a unit pref the 3x3 table `lut` answers directly.
// Requires box sides <= COOP_MAX_SIDE.
struct GeoCand {
  int k, d2, dot, di, dj;
  bool have;
  __device__ __forceinline__ void take(int cdi, int cdj, int p0, int p1, int &adi, int &adj, bool &alt) {
    const int ck = probe_index(max(abs(cdi), abs(cdj))), cd2 = cdi * cdi + cdj * cdj, cdot = cdi * p0 + cdj * p1;
    const bool same = have && ck == k && cd2 == d2;
    const bool lt = !have || ck < k || (ck == k && cd2 < d2) || (same && cdot > dot);
    if (lt) { have = true; k = ck; d2 = cd2; dot = cdot; di = cdi; dj = cdj; alt = false; }
    else if (same && cdot == dot) { alt = true; adi = cdi; adj = cdj; }
  }
};

// The warp's winner among the lanes' best candidates (at least one lane has one): smallest (k, d2), then the largest
// dot product; a mirror pair -- a second lane's best, or the first lane's own second candidate -- by the cross
// product or, for a pref where the reference's fp64 expression need not tie exactly, by that expression.
__device__ __forceinline__ void geo_pick(const GeoCand &b, int adi, int adj, bool alt, int p0, int p1, int &di, int &dj,
                                         int &kmin) {
  kmin = __reduce_min_sync(FULL_MASK, b.have ? b.k : INT32_MAX);
  const bool s1 = b.have && b.k == kmin;
  const int dmin = __reduce_min_sync(FULL_MASK, s1 ? b.d2 : INT32_MAX);
  const bool s2 = s1 && b.d2 == dmin;
  const int dotmax = __reduce_max_sync(FULL_MASK, s2 ? b.dot : INT32_MIN);
  const unsigned tied = __ballot_sync(FULL_MASK, s2 && b.dot == dotmax);
  const int la = __ffs(tied) - 1;
  di = __shfl_sync(FULL_MASK, b.di, la); dj = __shfl_sync(FULL_MASK, b.dj, la);
  const unsigned rest = tied & (tied - 1);
  const bool alt_a = __shfl_sync(FULL_MASK, (int)alt, la) != 0;
  if (rest || alt_a) {
    const int lb = rest ? __ffs(rest) - 1 : la;
    const int odi = __shfl_sync(FULL_MASK, rest ? b.di : adi, lb), odj = __shfl_sync(FULL_MASK, rest ? b.dj : adj, lb);
    const int cb = di * p1 - dj * p0, ca = odi * p1 - odj * p0;
    bool other_better;
    if (pref_ties_exactly(p0, p1)) {
      other_better = ca > cb;
    } else {
      const double sb = tie_sp1(di, dj, dmin, p0, p1), sa = tie_sp1(odi, odj, dmin, p0, p1);
      other_better = sa != sb ? sa > sb : ca > cb;
    }
    if (other_better) { di = odi; dj = odj; }
  }
}

__device__ __forceinline__ bool find_next_geo(const uint32_t *bm, int h, int w, int ws, int ci, int cj, int p0, int p1,
                                              int &rad, const uint8_t *lut, int &bi, int &bj) {
  const int lane = (int)lane_id();
  if (lut && rad == 1 && (unsigned)(p0 + 1) <= 2u && (unsigned)(p1 + 1) <= 2u) {  // uniform over the warp
    unsigned m = 0;
#pragma unroll
    for (int rr = 0; rr < 3; rr++) {
      const int ri = ci + rr - 1;
      const uint32_t x = (ri >= 0 && ri < h) ? row_window(bm + ri * ws, ws, cj) : 0u;
      m |= ((x >> 14) & 7u) << (3 * rr);
    }
    if (m) {
      const int idx = lut[((p0 + 1) * 3 + (p1 + 1)) * TPR_LUT_COLS + m];
      bi = ci + idx / 3 - 1; bj = cj + idx % 3 - 1;
      return true;
    }
    rad = 2;
  }
  GeoCand b;
  b.have = false; b.k = 0; b.d2 = 0; b.dot = 0; b.di = 0; b.dj = 0;
  int adi = 0, adj = 0;
  bool alt = false;
  for (;; rad <<= 1) {
    const int i0 = max(ci - rad, 0), i1 = min(ci + rad, h - 1);
    if (rad <= 15) {  // one aligned word per row
      const uint32_t wmask = ((2u << (2 * rad)) - 1u) << (15 - rad);
      const int i = i0 + lane;
      if (i <= i1) {
        const uint32_t x = row_window(bm + i * ws, ws, cj) & wmask;
        const uint32_t left = x & 0x7fffu, right = x >> 15;
        const int dl = __clz(left) - 16, dr = __ffs(right) - 1;  // nearest unvisited column on each side
        if (left && (!right || dl <= dr)) b.take(i - ci, -dl, p0, p1, adi, adj, alt);
        if (right && (!left || dr <= dl)) b.take(i - ci, dr, p0, p1, adi, adj, alt);
      }
    } else {
      const int j0 = max(cj - rad, 0), j1 = min(cj + rad, w - 1);
      const int w0 = j0 >> 5, w1 = j1 >> 5;
      for (int i = i0 + lane; i <= i1; i += 32)
        for (int wd = w0; wd <= w1; wd++) {
          uint32_t bits = bm[i * ws + wd];
          const int lo = wd << 5;
          if (lo < j0) bits &= 0xffffffffu << (j0 - lo);
          if (lo + 31 > j1) bits &= 0xffffffffu >> (lo + 31 - j1);
          if (!bits) continue;
          const int rel = min(cj - lo, 31);
          const uint32_t lmask = rel < 0 ? 0u : (2u << rel) - 1u;  // columns <= cj
          const uint32_t left = bits & lmask, right = bits & ~lmask;
          if (left) b.take(i - ci, lo + 31 - __clz(left) - cj, p0, p1, adi, adj, alt);
          if (right) b.take(i - ci, lo + __ffs(right) - 1 - cj, p0, p1, adi, adj, alt);
        }
    }
    if (__any_sync(FULL_MASK, b.have)) break;
    if (ci - rad <= 0 && cj - rad <= 0 && ci + rad >= h - 1 && cj + rad >= w - 1) return false;
  }
  int di, dj, kmin;
  geo_pick(b, adi, adj, alt, p0, p1, di, dj, kmin);
  bi = ci + di; bj = cj + dj;
  rad = 1 << kmin;
  return true;
}

// ---- gradpath (Region.grad_path, rbepwt.py:1190-1271) ------------------------------------------------------
// Same probes as the easy path; among the candidates of the first non-empty probe the smallest distance wins
// (euclid: d2; chebyshev), equal distances are settled by the region's average gradient: with
// perp = rotate(avg_gradient, -pi/2), tmp = rotate(perp, -pi/2) and v = offset / ||offset||, the larger
// A = |np.dot(v, perp)| wins, then the larger B = |np.dot(v, tmp)| (1231-1247; np.dot of 2-vectors =
// fma(x1, y1, x0 * y0)).  The reference flips the sign of perp as it goes (1256-1259), which cannot change a
// choice: only absolute values are compared.  A complete tie (A and B bit-identical -- always for two opposite
// offsets -- or a NaN direction from a zero gradient) keeps whichever candidate the reference's loop over a Python
// set met first: unpinned.  Rule here (and in the fixtures, produced by the reference with its set iteration pinned
// to sorted order): the first candidate in row-major order.
struct GradDir { double perp0, perp1, tmp0, tmp1; };

struct GradCand {
  int dist, i, j;  // dist < 0: none
  double A, B;
};

// is `c` preferred to `b`?  (total order: dist asc, A desc, B desc, (i, j) asc; NaN compares as a tie)
__device__ __forceinline__ bool grad_better(const GradCand &c, const GradCand &b) {
  if (c.dist < 0) return false;
  if (b.dist < 0) return true;
  if (c.dist != b.dist) return c.dist < b.dist;
  if (c.A > b.A) return true;
  if (c.A < b.A) return false;
  if (c.A == b.A) {
    if (c.B > b.B) return true;
    if (c.B < b.B) return false;
  }
  return c.i != b.i ? c.i < b.i : c.j < b.j;
}

template <bool CHEB>
__device__ __forceinline__ bool find_next_grad(const uint32_t *bm, int h, int w, int ws, int ci, int cj, const GradDir &G,
                                               int &bi, int &bj) {
  const int lane = (int)lane_id();
  GradCand b;
  b.dist = -1; b.i = b.j = 0; b.A = b.B = 0.0;
  for (int rad = 1;; rad <<= 1) {  // half-width 2^(k-1), k = 1,2,...   rbepwt.py:1218-1222
    const int i0 = max(ci - rad, 0), i1 = min(ci + rad, h - 1);
    const int j0 = max(cj - rad, 0), j1 = min(cj + rad, w - 1);
    const int w0 = j0 >> 5, w1 = j1 >> 5;
    for (int i = i0 + lane; i <= i1; i += 32)
      for (int wd = w0; wd <= w1; wd++) {
        uint32_t bits = bm[i * ws + wd];
        const int lo = wd << 5;
        if (lo < j0) bits &= 0xffffffffu << (j0 - lo);
        if (lo + 31 > j1) bits &= 0xffffffffu >> (lo + 31 - j1);
        while (bits) {
          const int j = lo + __ffs(bits) - 1;
          bits &= bits - 1;
          const int di = i - ci, dj = j - cj, d2 = di * di + dj * dj;
          GradCand c;
          c.dist = CHEB ? max(abs(di), abs(dj)) : d2;
          c.i = i; c.j = j;
          if (b.dist >= 0 && c.dist > b.dist) continue;
          const double nrm = sqrt((double)d2);
          const double v0 = (double)di / nrm, v1 = (double)dj / nrm;
          c.A = fabs(fma(v1, G.perp1, rb_dmul(v0, G.perp0)));
          c.B = fabs(fma(v1, G.tmp1, rb_dmul(v0, G.tmp0)));
          if (grad_better(c, b)) b = c;
        }
      }
    if (__any_sync(FULL_MASK, b.dist >= 0)) break;
    if (i0 == 0 && j0 == 0 && i1 == h - 1 && j1 == w - 1) return false;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    GradCand o;
    o.dist = __shfl_xor_sync(FULL_MASK, b.dist, d);
    o.i = __shfl_xor_sync(FULL_MASK, b.i, d);
    o.j = __shfl_xor_sync(FULL_MASK, b.j, d);
    o.A = __shfl_xor_sync(FULL_MASK, b.A, d);
    o.B = __shfl_xor_sync(FULL_MASK, b.B, d);
    if (grad_better(o, b)) b = o;
  }
  bi = b.i; bj = b.j;  // identical in every lane: the order is total
  return true;
}

// rotate(v, -pi/2) as the reference computes it (rbepwt.py:78-82): np.dot of [[c, -s], [s, c]] with v
__device__ __forceinline__ void rotate_mhalfpi(double v0, double v1, double &o0, double &o1) {
  const double c = 6.123233995736766e-17, s = -1.0;  // np.cos(-np.pi/2), np.sin(-np.pi/2)
  o0 = fma(-s, v1, rb_dmul(c, v0));
  o1 = fma(c, v1, rb_dmul(s, v0));
}

// Region.compute_avg_gradient (rbepwt.py:1102-1113) over np.gradient(img) (2010-2013) for the region whose level-1
// bitmap is in `bm`: the sums run over the region's points in their level-1 order (row-major), so ONE lane adds them
// up, in that order -- fp64 addition is not associative and the direction must come out bit for bit.
__device__ __forceinline__ GradDir region_grad_dir(const double *__restrict__ img, int H, int W, const uint32_t *bm, int h, int w,
                                                   int ws, int r0, int c0, int npoints) {
  double s0 = 0.0, s1 = 0.0;
  if (lane_id() == 0) {
    for (int bi = 0; bi < h; bi++)
      for (int wd = 0; wd < ws; wd++) {
        uint32_t bits = bm[bi * ws + wd];
        while (bits) {
          const int bj = (wd << 5) + __ffs(bits) - 1;
          bits &= bits - 1;
          const int i = r0 + bi, j = c0 + bj;
          const double *p = img + (size_t)i * W + j;
          double g0, g1;  // np.gradient: central differences inside, one-sided at the border
          if (i == 0) g0 = __dsub_rn(p[W], p[0]);
          else if (i == H - 1) g0 = __dsub_rn(p[0], p[-W]);
          else g0 = __ddiv_rn(__dsub_rn(p[W], p[-W]), 2.0);
          if (j == 0) g1 = __dsub_rn(p[1], p[0]);
          else if (j == W - 1) g1 = __dsub_rn(p[0], p[-1]);
          else g1 = __ddiv_rn(__dsub_rn(p[1], p[-1]), 2.0);
          s0 = __dadd_rn(s0, g0); s1 = __dadd_rn(s1, g1);
        }
      }
    s0 = __ddiv_rn(s0, (double)npoints); s1 = __ddiv_rn(s1, (double)npoints);
    const double nrm = sqrt(fma(s1, s1, rb_dmul(s0, s0)));  // np.linalg.norm
    s0 = __ddiv_rn(s0, nrm); s1 = __ddiv_rn(s1, nrm);
  }
  s0 = __shfl_sync(FULL_MASK, s0, 0); s1 = __shfl_sync(FULL_MASK, s1, 0);
  GradDir G;
  rotate_mhalfpi(s0, s1, G.perp0, G.perp1);
  rotate_mhalfpi(G.perp0, G.perp1, G.tmp0, G.tmp1);
  return G;
}

// Walk one region's path at one level.  (ci,cj) = start point (bitmap coordinates, bit still set).
// Ql[t], t = 0..n-1, receives the pixel ids in path order.  The bitmap is all-zero afterwards.
template <int MODE>
__device__ __forceinline__ bool run_path(uint32_t *bm, int h, int w, int ws, int ci, int cj, int n, int r0, int c0,
                                         int logW, const double *__restrict__ vals, bool u8wrap,
                                         int32_t *__restrict__ Ql, int32_t *__restrict__ Pl, const int32_t *posmap,
                                         const uint8_t *lut = nullptr, const GradDir *grad = nullptr) {
  const int lane = (int)lane_id();
  const bool geo = MODE == MODE_EUCLID && h <= COOP_MAX_SIDE && w <= COOP_MAX_SIDE;
  int rad = 1;
  int myq = 0;
  if (lane == 0) {
    myq = ((r0 + ci) << logW) + c0 + cj;
    bm[ci * ws + (cj >> 5)] &= ~(1u << (cj & 31));
  }
  __syncwarp();
  int p0 = 0, p1 = 1;  // prefered_direc = (0,1)   rbepwt.py:1290
  double curval = 0.0;
  if (MODE == MODE_EPWT) curval = RB_EPWT_LOAD(vals + (((r0 + ci) << logW) + c0 + cj));
  for (int t = 1; t < n; t++) {
    int bi, bj;
    if (geo) {
      if (!find_next_geo(bm, h, w, ws, ci, cj, p0, p1, rad, lut, bi, bj)) return false;
    } else if (mode_is_grad(MODE)) {
      if (!find_next_grad<MODE == MODE_GRAD_CHEB>(bm, h, w, ws, ci, cj, *grad, bi, bj)) return false;
    } else if (!find_next<MODE>(bm, h, w, ws, ci, cj, p0, p1, vals, r0, c0, logW, u8wrap, curval, bi, bj)) {
      return false;
    }
    if (lane == 0) bm[bi * ws + (bj >> 5)] &= ~(1u << (bj & 31));
    __syncwarp();
    if ((t & 31) == lane) myq = ((r0 + bi) << logW) + c0 + bj;
    if ((t & 31) == 31) {  // coalesced flush of 32 path points (+ their positions in the incoming order)
      Ql[t - 31 + lane] = myq;
      if (Pl) Pl[t - 31 + lane] = __ldcg(posmap + myq);
    }
    p0 = bi - ci; p1 = bj - cj;  // rbepwt.py:1331
    ci = bi; cj = bj;
  }
  if (lane < (n & 31)) {
    Ql[(n & ~31) + lane] = myq;
    if (Pl) Pl[(n & ~31) + lane] = __ldcg(posmap + myq);
  }
  return true;
}

// After a level: the points at even GLOBAL position a+t survive (RegionCollection.reduce,
// rbepwt.py:1563-1584).  Re-marks them in the (all-zero) bitmap, returns the smallest surviving
// pixel id = next level's start point (lexicographic min (row,col), rbepwt.py:1035-1036).
__device__ __forceinline__ int reduce_points(uint32_t *bm, int ws, int a, int n, int r0, int c0, int logW,
                                             const int32_t *Ql, int32_t *posmap) {
  const int lane = (int)lane_id();
  const int Wm = (1 << logW) - 1;
  __syncwarp();
  int minpix = INT32_MAX;
  for (int t = lane; t < n; t += 32) {
    if (((a + t) & 1) == 0) {
      const int pix = __ldcg(Ql + t);
      const int i = (pix >> logW) - r0, j = (pix & Wm) - c0;
      atomicOr(&bm[i * ws + (j >> 5)], 1u << (j & 31));
      if (posmap) posmap[pix] = (a + t) >> 1;  // EPWT: the survivor's place in the next level's incoming order
      minpix = min(minpix, pix);
    }
  }
  minpix = __reduce_min_sync(FULL_MASK, minpix);
  __syncwarp();
  return minpix;
}

// Geometric modes: the whole pyramid of one region.
template <int MODE>
__device__ void region_pyramid(const PathParams &P, int g, uint32_t *bm, const uint8_t *lut = nullptr) {
  const int lane = (int)lane_id();
  const int logW = P.logW, W = P.W, N = P.N;
  const int img = P.reg.img[g], label = P.reg.label[g], first = P.reg.first[g];
  int n = P.reg.size[g], a = P.reg.off[g];
  const int r0 = first >> logW, c0 = P.reg.cmin[g];
  const int h = P.reg.rmax[g] - r0 + 1, w = P.reg.cmax[g] - c0 + 1, ws = (w + 31) >> 5;
  const int32_t *lab = P.labels + (size_t)img * N;
  int32_t *Q = P.Q + (size_t)img * 2 * (size_t)N;

  const int words = h * ws;
  for (int wi = 0; wi < words; wi += 4) {  // four independent label loads in flight per lane
    int lv[4];
    bool inb[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int w_ = wi + u;
      const int i = ws == 1 ? w_ : w_ / ws;
      const int col = ((w_ - i * ws) << 5) + lane;
      inb[u] = w_ < words && col < w;
      lv[u] = inb[u] ? lab[((r0 + i) << logW) + c0 + col] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const unsigned bits = __ballot_sync(FULL_MASK, inb[u] && lv[u] == label);
      if (lane == 0 && wi + u < words) bm[wi + u] = bits;
    }
  }
  __syncwarp();
  GradDir G = {0.0, 0.0, 0.0, 0.0};
  if (mode_is_grad(MODE)) G = region_grad_dir(P.img + (size_t)img * N, P.H, W, bm, h, w, ws, r0, c0, n);
  int si = 0, sj = (first & (W - 1)) - c0;
  for (int lev = 1; lev <= P.levels && n > 0; lev++) {
    int32_t *Ql = Q + level_off((size_t)N, lev) + a;
    // paths only: the positions in the incoming order (Pm) of every region are computed by k2_perm (walk.cuh)
    if (!run_path<MODE>(bm, h, w, ws, si, sj, n, r0, c0, logW, nullptr, false, Ql, nullptr, nullptr, lut, &G)) {
      if (lane == 0) atomicExch(&P.qmeta[QM_ERR], 1);
      return;
    }
    if (lev == P.levels) break;
    const int minpix = reduce_points(bm, ws, a, n, r0, c0, logW, Ql, nullptr);
    const int na = (a + 1) >> 1, nb = (a + n + 1) >> 1;
    a = na; n = nb - na;
    if (n > 0) { si = (minpix >> logW) - r0; sj = (minpix & (W - 1)) - c0; }
  }
}

// EPWT: one region per image, paths depend on the level's values, so one launch per level
// (rbepwt.py:2004-2006, 2031).  vals = pixel-addressed values of this level ([B][N]).
struct EpwtParams {
  int H, W, logW, N, lev, img0;
  const double *vals;
  int32_t *Q;        // [B][2N]
  int32_t *Pm;       // [B][2N], see PathParams
  int32_t *posmap;   // [B][N]
  uint32_t *gscratch;
  size_t gscratch_words;
  int smem_words;
  int u8wrap;
  int *qmeta;
};

// the walker with the bitmap window in registers (regwin.cuh)
__device__ bool rw_run_path_epwt(uint32_t *bm, int h, int w, int ws, int ci, int cj, int n, int logW,
                                 const double *__restrict__ vals, bool u8wrap, int32_t *__restrict__ Ql,
                                 int32_t *__restrict__ Pl, const int32_t *posmap);

__global__ void __launch_bounds__(32) k1_epwt_level(EpwtParams P) {
  extern __shared__ uint32_t s_big[];
  const int lane = (int)lane_id();
  const int img = P.img0 + blockIdx.x;
  const int H = P.H, W = P.W, N = P.N, logW = P.logW, ws = (W + 31) >> 5, words = H * ws;
  uint32_t *bm = words <= P.smem_words ? s_big : P.gscratch + (size_t)blockIdx.x * P.gscratch_words;
  int32_t *Q = P.Q + (size_t)img * 2 * (size_t)N;
  int32_t *posmap = P.posmap + (size_t)img * N;
  int32_t *Pl = P.lev >= 2 ? P.Pm + (size_t)img * 2 * (size_t)N + level_off((size_t)N, P.lev) : nullptr;
  const double *vals = P.vals + (size_t)img * N;
  const int n = N >> (P.lev - 1);
  int32_t *Ql = Q + level_off((size_t)N, P.lev);
  int start;
  if (P.lev == 1) {
    const uint32_t full = W >= 32 ? 0xffffffffu : ((1u << W) - 1u);
    for (int i = lane; i < words; i += 32) bm[i] = full;
    __syncwarp();
    start = 0;
  } else {
    for (int i = lane; i < words; i += 32) bm[i] = 0u;
    __syncwarp();
    start = reduce_points(bm, ws, 0, N >> (P.lev - 2), 0, 0, logW, Q + level_off((size_t)N, P.lev - 1), posmap);
  }
#ifdef RB_NO_REGWIN
  const bool ok = run_path<MODE_EPWT>(bm, H, W, ws, start >> logW, start & (W - 1), n, 0, 0, logW, vals,
                                      P.u8wrap && P.lev == 1, Ql, Pl, posmap);
#else
  const bool ok = rw_run_path_epwt(bm, H, W, ws, start >> logW, start & (W - 1), n, logW, vals, P.u8wrap && P.lev == 1, Ql, Pl,
                                   posmap);
#endif
  if (!ok && lane == 0) atomicExch(&P.qmeta[QM_ERR], 1);
}

}  // namespace rbepwt
