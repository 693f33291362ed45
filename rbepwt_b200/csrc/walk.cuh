// K1, thread-per-region form: every LANE of a warp owns one region and walks its greedy path pyramid
// (all levels) alone; a warp walks up to 32 regions at once.  This is the kernel the benchmark
// configurations spend their path time in (paths.cuh keeps the warp-per-region form for huge regions
// and EPWT).
//
// Step rule = Region.easy_path (/root/reference/rbepwt.py:1273-1347): smallest square window of
// half-width 1,2,4,... holding an unvisited point, then the lexicographic key (-dist, sp1, sp2).
//
// Execution shape.  A path step is a short dependent chain, while a batch holds 10^5..10^6 independent
// regions, so the parallel axis is the region, not the window.  What bounds the kernel is the number of
// instructions a warp issues per step, i.e. how many of its 32 lanes do useful work in an instruction.
// Three decisions follow from that:
//
//  * One uniform step for (almost) every step.  The probes of half-width 1 and 2 see the 5x5 window
//    around the current point.  Every bitmap carries a margin of WK_PAD = 2 empty rows and columns, so
//    five shared-memory row words and shifts (no bounds checks) give the window as a 25-bit register
//    N25, and the step is resolved from it: the classes d2 = 1, 2, 4, 5, 8 in that order are exactly
//    "probe 1, then probe 2, nearest first" (euclid; chebyshev: ring 1, then ring 2), the first
//    non-empty class holds the candidates, and the winner among them for the current preferred
//    direction is ONE byte of the table T2 ([25 prefs that are 5x5 offsets][320]: the class's subset,
//    4 or 8 bits, is turned into a table slot by a multiply-shift perfect hash of N25 & class mask --
//    no bit gathering).  The table is filled once per context by the same candidate code the other
//    steps use (wk_t2_fill), so its entries are the reference's tie-break by construction.  On the
//    benchmark 89 % of all steps end here (tools/tie_stats.c).
//  * Lanes are not synchronised per level.  A lane that finishes a level starts the next one
//    immediately (the region offsets of every level follow from the level-1 sizes alone), so a warp
//    never waits for its slowest lane at each of the 16 levels, only once per chunk.
//  * Nothing but the walk.  The survivors of a level (even GLOBAL positions, RegionCollection.reduce,
//    rbepwt.py:1563-1584) are marked in a second bitmap plane -- or appended to a list -- at the moment
//    they are visited; the two planes swap at the end of the level.  The kernel writes the paths as
//    pixel ids (Q) and never reads them back; the positions in the incoming order (Pm) the transform
//    kernels gather through are computed afterwards by k2_perm, a fully parallel kernel.
//
// Beyond the 5x5 window (FAR): the lane only flags the request; the WARP serves the requests one after the
// other (wk_far_search: lanes = rows of the requester's bitmap plane, a row's candidate = its nearest unvisited
// column on either side, ranked by the packed key (k = index of the first probe that would contain the candidate,
// d2) and the dot product, two warp reductions) -- scanning the whole plane once equals the reference's probes
// 4, 8, ... in turn.  The lane-local form of this search (kept in tests/host_walk as a second restatement) was
// measured slower: it runs with ~3 of 32 lanes.  After the searches every lane that moved commits in one place.
//
// List mode: from the first level with at most WK_LIST_MAX points the lane drops the bitmaps and holds
// the points as a list of packed (row, col); a step scans the unvisited ones (a 32-bit mask).
//
// Integer tie-break (euclid): candidates compared by sp1 have the same d2, hence the same norm n, and
// sp1 = fl(fl(dj/n)*p1 + fl(fl(di/n)*p0)) orders them like the integer dot product di*p0 + dj*p1 whenever
// the dot products differ (they differ by >= 1/n, the rounding error is < 2^-20/n for coordinates below
// 2^15).  Equal dot products = mirror images about pref: the fp64 expression ties exactly when pref is
// on an axis or |p0| == |p1| a power of two (-> sp2, the integer cross product, decides), else it is
// evaluated bit for bit.  Chebyshev compares candidates of different norms: always fp64.
//
// The per-lane logic (struct Walker) is __host__ __device__: tests/host_walk compiles it for the CPU and
// checks whole pyramids against the oracle without a GPU; the kernel below only adds the warp loop.
#pragma once
#include "paths.cuh"

namespace rbepwt {

// Bulk instantiation: two large CTAs per SM, so that the step table (8 KB per CTA) costs little of the shared memory and
// the arenas get the rest; 24 warps per SM is what the registers allow (measured: 20 warps with larger arenas 7.1 ms,
// 24 warps 6.8 ms, 26 warps with spills 8.7 ms per 512 images).
#ifndef WK_WARPS_N
#define WK_WARPS_N 12
#endif
constexpr int WK_WARPS = WK_WARPS_N;
#ifndef WK_MIN_CTAS
#define WK_MIN_CTAS 2
#endif
#ifndef WK_WIDE_WARPS_N
#define WK_WIDE_WARPS_N 4
#endif
constexpr int WK_WIDE_WARPS = WK_WIDE_WARPS_N;  // windowed instantiation: small CTAs (few chunks, the longest chains)
constexpr int COOP_WARPS = 4;                   // k1_coop_all: warps (= regions in flight) per CTA
#ifndef WK_WIDE_MIN_CTAS_N
#define WK_WIDE_MIN_CTAS_N 3
#endif
constexpr int WK_WIDE_MIN_CTAS = WK_WIDE_MIN_CTAS_N;  // measured on heavy-tailed maps: 4 x 4 warps 17.9k, 5 x 4 17.6k, 2 x 12 15.0k images/s
#ifndef WK_NEAR_REPS_N
#define WK_NEAR_REPS_N 4
#endif
constexpr int WK_NEAR_REPS = WK_NEAR_REPS_N;  // 5x5 steps a lane may take per trip of the warp loop
#ifndef WK_LIST_BATCH_N
#define WK_LIST_BATCH_N 12
#endif
constexpr int WK_LIST_BATCH = WK_LIST_BATCH_N;  // list-mode lanes wait for this many of their kind (or for the others to finish)

// what a lane does next: a step from the 5x5 window; wait for the warp's search beyond it; commit what that search
// found; a list-mode step; start the next level (t == n)
enum : int { WK_DONE = 0, WK_NEAR = 1, WK_FAR = 2, WK_LIST = 3, WK_LEVEL = 4, WK_COMMIT = 5, WK_ERROR = 6 };

template <int MODE>
struct Search;

// ---- euclid: packed integer keys, fp64 only for mirror pairs under a non-exact pref ----------------
template <>
struct Search<MODE_EUCLID> {
  // incumbent: key = k << 21 | d2 (sides <= TPR_MAX_SIDE: d2 < 2^21, k <= 11), then the larger dot product;
  // offsets packed (di << 16) | (dj & 0xffff)
  unsigned key;
  int dot, off, aoff;  // aoff: mirror partner with the same (key, dot)
  int tag, atag;       // caller's payload of the incumbent / its mirror partner (list mode: list index)
  bool alt;

  __host__ __device__ __forceinline__ void reset() { key = 0xffffffffu; dot = 0; off = aoff = 0; tag = atag = 0; alt = false; }
  __host__ __device__ __forceinline__ bool have() const { return key != 0xffffffffu; }

  __host__ __device__ __forceinline__ void consider(bool valid, int cdi, int cdj, int p0, int p1, int ctag = 0) {
    const int k = probe_index(max(abs(cdi), abs(cdj)));
    const unsigned ckey = valid ? ((unsigned)k << 21) | (unsigned)(cdi * cdi + cdj * cdj) : 0xffffffffu;
    const int cdot = cdi * p0 + cdj * p1;
    const int coff = (int)(((unsigned)cdi << 16) | ((unsigned)cdj & 0xffffu));
    // selects, not branches: every lane executes the same instructions
    const bool same = ckey == key;
    const bool lt = ckey < key || (same && cdot > dot);
    const bool eq = valid && same && cdot == dot;  // mirror image of the incumbent about pref
    key = lt ? ckey : key;
    dot = lt ? cdot : dot;
    off = lt ? coff : off;
    tag = lt ? ctag : tag;
    alt = lt ? false : (alt || eq);
    aoff = eq ? coff : aoff;
    atag = eq ? ctag : atag;
  }

  // Candidates of one row: the nearest unvisited column on the left (distance dl >= 1) and on the right
  // (dr >= 0); the nearer one dominates the other in (k, d2), both compete only when dl == dr.
  __host__ __device__ __forceinline__ void row_candidates(bool hl, int dl, bool hr, int dr, int rdi, int p0, int p1) {
    if (!(hl || hr)) return;
    const bool left_first = hl && (!hr || dl <= dr);
    consider(true, rdi, left_first ? -dl : dr, p0, p1);
    if (hl && hr && dl == dr) consider(true, rdi, dr, p0, p1);
  }

  // one bitmap word of row ci+rdi: columns lo..lo+31, already masked to the window
  __host__ __device__ __forceinline__ void scan_word(uint32_t bits, int lo, int rdi, int cj, int p0, int p1) {
    const int rel = min(cj - lo, 31);
    const uint32_t lmask = rel < 0 ? 0u : (2u << rel) - 1u;  // columns <= cj
    const uint32_t left = bits & lmask, right = bits & ~lmask;
    row_candidates(left != 0u, cj - (lo + 31 - rb_clz(left)), right != 0u, lo + rb_ffs(right) - 1 - cj, rdi, p0, p1);
  }

  // one window row as an aligned word: bit 15 + dj <-> column cj + dj
  __host__ __device__ __forceinline__ void scan_row(uint32_t x, int rdi, int p0, int p1) {
    const uint32_t left = x & 0x7fffu, right = x >> 15;
    row_candidates(left != 0u, rb_clz(left) - 16, right != 0u, rb_ffs(right) - 1, rdi, p0, p1);  // hb = 31 - clz -> dl = 15 - hb
  }

  __host__ __device__ __forceinline__ void finish(int p0, int p1, int &odi, int &odj, int &k) {
    int di = off >> 16, dj = (int)(short)(off & 0xffff);
    if (alt) {
      const int adi = aoff >> 16, adj = (int)(short)(aoff & 0xffff);
      const int cb = di * p1 - dj * p0, ca = adi * p1 - adj * p0;
      bool alt_better;
      if (pref_ties_exactly(p0, p1)) {
        alt_better = ca > cb;
      } else {
        const int d2 = (int)(key & 0x1fffffu);
        const double sb = tie_sp1(di, dj, d2, p0, p1), sa = tie_sp1(adi, adj, d2, p0, p1);
        alt_better = sa != sb ? sa > sb : ca > cb;
      }
      if (alt_better) { di = adi; dj = adj; tag = atag; }
    }
    odi = di; odj = dj; k = (int)(key >> 21);
  }
};

// ---- chebyshev: every point of the nearest ring competes through the fp64 sp1 ----------------------
template <>
struct Search<MODE_CHEB> {
  int c, d2, di, dj, tag;
  double sp1;
  bool found, has_sp1;

  __host__ __device__ __forceinline__ void reset() { found = false; has_sp1 = false; c = d2 = di = dj = tag = 0; sp1 = 0.0; }
  __host__ __device__ __forceinline__ bool have() const { return found; }

  __host__ __device__ __forceinline__ void consider(bool valid, int cdi, int cdj, int p0, int p1, int ctag = 0) {
    if (!valid) return;
    const int cc = max(abs(cdi), abs(cdj)), cd2 = cdi * cdi + cdj * cdj;
    if (found && cc > c) return;
    if (!found || cc < c) {
      found = true; has_sp1 = false; c = cc; d2 = cd2; di = cdi; dj = cdj; tag = ctag;
      return;
    }
    if (!has_sp1) { sp1 = tie_sp1(di, dj, d2, p0, p1); has_sp1 = true; }
    const double s = tie_sp1(cdi, cdj, cd2, p0, p1);
    const bool better = s != sp1 ? s > sp1 : (cdi * p1 - cdj * p0) > (di * p1 - dj * p0);
    if (better) { sp1 = s; d2 = cd2; di = cdi; dj = cdj; tag = ctag; }
  }

  __host__ __device__ __forceinline__ void scan_word(uint32_t bits, int lo, int rdi, int cj, int p0, int p1) {
    while (bits) {
      const int j = lo + rb_ffs(bits) - 1;
      bits &= bits - 1;
      consider(true, rdi, j - cj, p0, p1);
    }
  }

  __host__ __device__ __forceinline__ void scan_row(uint32_t x, int rdi, int p0, int p1) {
    while (x) {
      const int b = rb_ffs(x) - 1;
      x &= x - 1;
      consider(true, rdi, b - 15, p0, p1);
    }
  }

  __host__ __device__ __forceinline__ void finish(int, int, int &odi, int &odj, int &k) {
    odi = di; odj = dj; k = probe_index(c);
  }
};

// Unit-step table: lut[q * 512 + m] = index (di+1)*3 + (dj+1) of the winner among the neighbours present in the
// 9-bit mask m (bit 3*(di+1) + (dj+1)), for pref = (q/3 - 1, q%3 - 1).  Filled by the candidate code every other
// path takes, once per context (global memory, one table per path mode); the path kernels copy theirs into shared
// memory without the always-empty centre bit (9 x 256 bytes).
template <int MODE>
__host__ __device__ __forceinline__ uint8_t unit_lut_entry(int q, int m) {
  const int p0 = q / 3 - 1, p1 = q % 3 - 1;
  if (q == 4 || (m & 16) || !m) return 0xff;
  Search<MODE> S;
  S.reset();
  for (int bpos = 0; bpos < 9; bpos++)
    if (m & (1 << bpos)) S.consider(true, bpos / 3 - 1, bpos % 3 - 1, p0, p1);
  int odi, odj, k;
  S.finish(p0, p1, odi, odj, k);
  return (uint8_t)((odi + 1) * 3 + (odj + 1));
}

constexpr int WK_LUT_BYTES = TPR_LUT_ROWS * 256;  // compact table in shared memory
__host__ __device__ __forceinline__ int wk_lut_index(int p0, int p1, unsigned m9) {
  return ((p0 + 1) * 3 + (p1 + 1)) * 256 + (int)((m9 & 15u) | ((m9 >> 5) << 4));
}
// compact index e -> index into the 9 x 512 table
__host__ __device__ __forceinline__ int wk_lut_source(int e) {
  const int q = e >> 8, m8 = e & 255;
  return q * TPR_LUT_COLS + ((m8 & 15) | ((m8 >> 4) << 5));
}

// ---- 5x5 window classes: bit 5*(di+2) + (dj+2) of N25 <-> offset (di, dj) ----------------------------
constexpr uint32_t N25_A = (1u << 7) | (1u << 11) | (1u << 13) | (1u << 17);                                      // d2 = 1
constexpr uint32_t N25_B = (1u << 6) | (1u << 8) | (1u << 16) | (1u << 18);                                       // d2 = 2
constexpr uint32_t N25_C = (1u << 2) | (1u << 10) | (1u << 14) | (1u << 22);                                      // d2 = 4
constexpr uint32_t N25_D = (1u << 1) | (1u << 3) | (1u << 5) | (1u << 9) | (1u << 15) | (1u << 19) | (1u << 21) | (1u << 23);  // d2 = 5
constexpr uint32_t N25_E = (1u << 0) | (1u << 4) | (1u << 20) | (1u << 24);                                       // d2 = 8
constexpr uint32_t N25_RING1 = N25_A | N25_B;

// 5x5 table step (euclid).  The candidates of a step are the cells of ONE class (same d2), and which of them wins
// depends only on pref: t2[pref cell][class base + field] = winning cell, for the 24 prefs that are themselves 5x5
// offsets (the previous step came from the 5x5 window, or the level just started: all but a few percent of the
// steps).  `field` = the class's bits gathered into a dense index by a multiply-shift perfect hash (constants found
// by search; wk_t2_field is checked to be injective by tests/test_host_walk.py).  The table is filled once per
// context by the candidate code every other path takes (Search<MODE_EUCLID>, fp64 tie-break included).
constexpr int T2_ROW = 320;              // A 16 + B 16 + C 16 + D 256 + E 16 entries per pref
constexpr int T2_BYTES = 25 * T2_ROW;
constexpr int T2_NCLS = 5;
__host__ __device__ __forceinline__ uint32_t wk_t2_mask(int c) {
  return c == 0 ? N25_A : c == 1 ? N25_B : c == 2 ? N25_C : c == 3 ? N25_D : N25_E;
}
__host__ __device__ __forceinline__ uint32_t wk_t2_magic(int c) {
  return c == 0 ? 0x1121002u : c == 1 ? 0x24442045u : c == 2 ? 0x4208100u : c == 3 ? 0x24010280u : 0xc8400212u;
}
__host__ __device__ __forceinline__ int wk_t2_shift(int c) { return c == 3 ? 24 : 28; }
__host__ __device__ __forceinline__ int wk_t2_base(int c) { return c == 0 ? 0 : c == 1 ? 16 : c == 2 ? 32 : c == 3 ? 48 : 304; }
__host__ __device__ __forceinline__ int wk_t2_field(int c, uint32_t n25) {
  return (int)(((n25 & wk_t2_mask(c)) * wk_t2_magic(c)) >> wk_t2_shift(c));
}
// entry for pref cell `pc` (5 * (p0 + 2) + p1 + 2) and the cells `bits` (a non-empty subset of class c's mask)
__host__ __device__ __forceinline__ uint8_t wk_t2_entry(int pc, uint32_t bits) {
  const int p0 = pc / 5 - 2, p1 = pc % 5 - 2;
  if (pc == 12 || !bits) return 0xff;
  Search<MODE_EUCLID> S;
  S.reset();
  for (uint32_t u = bits; u; u &= u - 1u) {
    const int b = rb_ffs(u) - 1;
    S.consider(true, b / 5 - 2, b % 5 - 2, p0, p1);
  }
  int odi, odj, k;
  S.finish(p0, p1, odi, odj, k);
  return (uint8_t)((odi + 2) * 5 + (odj + 2));
}
// the i-th subset of class c's cells, as a 25-bit window value
__host__ __device__ __forceinline__ uint32_t wk_t2_subset(int c, int i) {
  uint32_t m = wk_t2_mask(c), v = 0u;
  for (int bit = 0; m; bit++, m &= m - 1u)
    if ((i >> bit) & 1) v |= m & (0u - m);
  return v;
}
// fills t2[T2_BYTES]: `job` runs over every (pref cell, class, subset)
constexpr int T2_JOBS = 25 * (16 * 4 + 256);
__host__ __device__ __forceinline__ void wk_t2_fill(uint8_t *t2, int job) {
  const int pc = job / (16 * 4 + 256);
  int r = job % (16 * 4 + 256), c;
  if (r < 48) { c = r >> 4; r &= 15; }
  else if (r < 304) { c = 3; r -= 48; }
  else { c = 4; r -= 304; }
  const uint32_t bits = wk_t2_subset(c, r);
  t2[pc * T2_ROW + wk_t2_base(c) + wk_t2_field(c, bits)] = wk_t2_entry(pc, bits);
}

// Same-distance candidates (one class of the 5x5 window, euclid): the larger integer dot product wins, a mirror
// pair (equal dot products) is settled like Search<MODE_EUCLID>::finish does.
__host__ __device__ __forceinline__ bool mirror_second_wins(int di, int dj, int adi, int adj, int d2, int p0, int p1) {
  const int cb = di * p1 - dj * p0, ca = adi * p1 - adj * p0;
  if (pref_ties_exactly(p0, p1)) return ca > cb;
  const double sb = tie_sp1(di, dj, d2, p0, p1), sa = tie_sp1(adi, adj, d2, p0, p1);
  return sa != sb ? sa > sb : ca > cb;
}

// The walker of one region: all levels, one bounded unit of work per call.
template <int MODE>
struct Walker {
  // the region
  uint32_t *bm;        // its arena slot: plane at `cur` = unvisited points of the level, plane at `nxt` = survivors
  int32_t *Qall;       // paths of the batch: [image][2N], levels concatenated
  ptrdiff_t pdelta;    // positions in the incoming order live at the same offsets of another array: Pm = Q + pdelta
                       // (written here for the list-mode levels only)
  int img;             // the region's image
  const uint8_t *lut;  // chebyshev: compact unit-step table (or null)
  const uint8_t *t2;   // euclid: 5x5 step table (or null)
  int N, W, L;
  int abase;           // word offset of the slot in the warp's arena
  int h, ws;           // bitmap rows (margins included), words per row
  int pixbase;         // pixel id of bitmap cell (0, 0): r0 * W + c0 (r0, c0 may be negative: the margin)
  bool narrow;         // every lane of the warp has ws == 1
  // the level
  int kind, lev, a, n, t;
  int ci, cj, p0, p1;
  int cur, nxt;
  int32_t *Qp;         // where the next path point goes
  bool keep, tolist;   // a next level exists / its points are collected as a list
  int ncnt, sminidx;   // list of survivors: length, index of the smallest
  uint32_t smin;       // smallest survivor (i << 16 | j) = next start point (lexicographic min, rbepwt.py:1035-1036)
  int pdi, pdj;        // WK_COMMIT: the step chosen (by the 5x5 window, or by the search beyond it)
  int rq0;             // this level's plane for the warp's search beyond the window: word offset | h << 13 | ws << 24
  uint32_t U;          // list mode: unvisited mask

  __host__ __device__ __forceinline__ bool done() const { return kind == WK_DONE || kind == WK_ERROR; }

  // region [a0, a0 + n0) of the level-1 signal; the level-1 bitmap is in plane 0, plane 1 is all zero;
  // (si, sj) = its lexicographically smallest point
  __host__ __device__ __forceinline__ void start(int a0, int n0, int si, int sj) {
    cur = 0; nxt = wk_plane_words(h, ws);
    lev = 1; a = a0; n = n0;
    smin = ((uint32_t)si << 16) | (uint32_t)sj; sminidx = 0;
    pdi = pdj = 0; U = 0u;
    if (n <= 0) { kind = WK_DONE; return; }
    begin_level(false);
  }

  __host__ __device__ __forceinline__ void emit() {  // the point (ci, cj) is the path's t-th
    *Qp++ = pixbase + ci * W + cj;
    if (keep && ((a + t) & 1) == 0) {  // even global position: survives (rbepwt.py:1570-1575)
      const uint32_t e = ((uint32_t)ci << 16) | (uint32_t)cj;
      if (tolist) {
        bm[nxt + ncnt] = e;
        if (e < smin) { smin = e; sminidx = ncnt; }
        ncnt++;
      } else {
        bm[nxt + ci * ws + (cj >> 5)] |= 1u << (cj & 31);
        smin = min(smin, e);
      }
    }
  }

  // lev, a, n are set; the level's points are in plane `cur` (bitmap, or list when from_list) and its start is smin
  __host__ __device__ __forceinline__ void begin_level(bool from_list) {
    keep = lev < L;
    const int nnext = ((a + n + 1) >> 1) - ((a + 1) >> 1);
    tolist = nnext <= WK_LIST_MAX;
    Qp = Qall + (size_t)img * 2 * (size_t)N + level_off((size_t)N, lev) + a;
    uint32_t e;
    if (from_list) {
      // a list holds the level's points in their incoming order: the list index IS the position k2_perm would compute
      e = bm[cur + sminidx];
      U = (n >= 32 ? 0xffffffffu : (1u << n) - 1u) & ~(1u << sminidx);
      Qp[pdelta] = a + sminidx;
    } else {
      e = smin;
    }
    ci = (int)(e >> 16); cj = (int)(e & 0xffffu);
    rq0 = (abase + cur) | (h << 13) | (ws << 24);
    if (!from_list) bm[cur + ci * ws + (cj >> 5)] &= ~(1u << (cj & 31));
    ncnt = 0; sminidx = 0; smin = 0xffffffffu;
    t = 0;
    emit();
    t = 1;
    p0 = 0; p1 = 1;  // prefered_direc = (0,1)   rbepwt.py:1290
    kind = t == n ? WK_LEVEL : (from_list ? WK_LIST : WK_NEAR);
  }

  // kind == WK_LEVEL (t == n): RegionCollection.reduce -- the region occupies [ceil(a/2), ceil((a+n)/2)) of the next level
  __host__ __device__ __forceinline__ void next_level() {
    const bool was_tolist = tolist;
    const int na = (a + 1) >> 1, nb = (a + n + 1) >> 1;
    if (!keep || nb - na <= 0) { kind = WK_DONE; return; }
    a = na; n = nb - na; lev++;
    const int x = cur; cur = nxt; nxt = x;  // the old plane is all zero again: every point was visited
    begin_level(was_tolist);
  }

  // the step (di, dj) was chosen
  __host__ __device__ __forceinline__ void commit(int di, int dj) {
    p0 = di; p1 = dj;  // rbepwt.py:1331
    ci += di; cj += dj;
    bm[cur + ci * ws + (cj >> 5)] &= ~(1u << (cj & 31));
    emit();
    t++;
    kind = t == n ? WK_LEVEL : WK_NEAR;
  }

  // the search beyond the 5x5 window (done for this lane by the warp, or by the host harness) found (di, dj)
  __host__ __device__ __forceinline__ void far_found(int di, int dj) { pdi = di; pdj = dj; kind = WK_COMMIT; }

  // ---- kind == WK_COMMIT: take the chosen step (lanes that chose from the window and lanes the warp searched for
  //      commit together)
  __host__ __device__ __forceinline__ void commit_step() { commit(pdi, pdj); }

  // ---- kind == WK_NEAR: choose the step from the 5x5 window (-> WK_COMMIT), or ask for a wider search (-> WK_FAR)
  __host__ __device__ __forceinline__ void near_select() {
    int di = 0, dj = 0;
    {
      const uint32_t *rp = bm + cur + (ci - WK_PAD) * ws;
      uint32_t n25;
      if (narrow) {
        const int sh = cj - WK_PAD;
        n25 = ((rp[0] >> sh) & 31u) | (((rp[1] >> sh) & 31u) << 5) | (((rp[2] >> sh) & 31u) << 10) |
              (((rp[3] >> sh) & 31u) << 15) | (((rp[4] >> sh) & 31u) << 20);
      } else {
        const int sc = cj - WK_PAD, sh = sc & 31;
        rp += sc >> 5;
        n25 = (rb_funnel_r(rp[0], rp[1], sh) & 31u) | ((rb_funnel_r(rp[ws], rp[ws + 1], sh) & 31u) << 5) |
              ((rb_funnel_r(rp[2 * ws], rp[2 * ws + 1], sh) & 31u) << 10) |
              ((rb_funnel_r(rp[3 * ws], rp[3 * ws + 1], sh) & 31u) << 15) |
              ((rb_funnel_r(rp[4 * ws], rp[4 * ws + 1], sh) & 31u) << 20);
      }
      if (n25 == 0u) {  // nothing within half-width 2: probes 4, 8, ...
        kind = WK_FAR;
        return;
      }
      const uint32_t ring1 = n25 & N25_RING1;
      if (MODE == MODE_EUCLID && t2 && (unsigned)(p0 + 2) <= 4u && (unsigned)(p1 + 2) <= 4u) {
        // the nearest class present -> its cells as a dense field -> the winner for this pref: no loop, no branch
        const int c = (n25 & N25_A) ? 0 : (n25 & N25_B) ? 1 : (n25 & N25_C) ? 2 : (n25 & N25_D) ? 3 : 4;
        const int cell = t2[((p0 + 2) * 5 + p1 + 2) * T2_ROW + wk_t2_base(c) + wk_t2_field(c, n25)];
        di = (cell * 13) >> 6;  // cell / 5 for cell < 25
        dj = cell - 5 * di - 2;
        di -= 2;
      } else if (MODE != MODE_EUCLID && lut && ring1 && (unsigned)(p0 + 1) <= 2u && (unsigned)(p1 + 1) <= 2u) {
        // unit pref, something at distance 1: the 3x3 table
        const unsigned m9 = ((n25 >> 6) & 7u) | (((n25 >> 11) & 7u) << 3) | (((n25 >> 16) & 7u) << 6);
        const int idx = lut[wk_lut_index(p0, p1, m9)];
        di = (idx * 11) >> 5;  // idx / 3 for idx < 9
        dj = idx - 3 * di - 1;
        di -= 1;
      } else {
        uint32_t cls;
        int d2 = 1;
        if (MODE == MODE_EUCLID) {
          cls = n25 & N25_A;
          if (!cls) { cls = n25 & N25_B; d2 = 2; }
          if (!cls) { cls = n25 & N25_C; d2 = 4; }
          if (!cls) { cls = n25 & N25_D; d2 = 5; }
          if (!cls) { cls = n25 & N25_E; d2 = 8; }
        } else {
          cls = ring1 ? ring1 : n25;
        }
        if ((cls & (cls - 1u)) == 0u) {  // one candidate
          const int b = rb_ffs(cls) - 1;
          di = (b * 13) >> 6;  // b / 5 for b < 25
          dj = b - 5 * di - 2;
          di -= 2;
        } else if (MODE == MODE_EUCLID) {  // same distance: the larger dot product with pref, then the mirror rule
          int bdot = INT32_MIN, adi = 0, adj = 0;
          bool alt = false;
          for (uint32_t u = cls; u; u &= u - 1u) {
            const int b = rb_ffs(u) - 1;
            const int qi = ((b * 13) >> 6) - 2, qj = b - 5 * (qi + 2) - 2;
            const int dot = qi * p0 + qj * p1;
            if (dot > bdot) { bdot = dot; di = qi; dj = qj; alt = false; }
            else if (dot == bdot) { alt = true; adi = qi; adj = qj; }
          }
          if (alt && mirror_second_wins(di, dj, adi, adj, d2, p0, p1)) { di = adi; dj = adj; }
        } else {
          Search<MODE> T;
          T.reset();
          for (uint32_t u = cls; u; u &= u - 1u) {
            const int b = rb_ffs(u) - 1;
            const int qi = (b * 13) >> 6;
            T.consider(true, qi - 2, b - 5 * qi - 2, p0, p1);
          }
          int fk;
          T.finish(p0, p1, di, dj, fk);
        }
      }
    }
    pdi = di; pdj = dj;
    kind = WK_COMMIT;
  }

  // ---- kind == WK_LIST: one whole step, candidates = the points still unvisited; then, at the end of the level,
  //      straight on to the next one (list levels are short: 16, 8, 4, ... points)
  __host__ __device__ __forceinline__ void list_step() {
    Search<MODE> S;
    S.reset();
    for (uint32_t u = U; u; u &= u - 1u) {
      const int idx = rb_ffs(u) - 1;
      const uint32_t e = bm[cur + idx];
      S.consider(true, (int)(e >> 16) - ci, (int)(e & 0xffffu) - cj, p0, p1, idx);
    }
    int di, dj, k;
    S.finish(p0, p1, di, dj, k);
    U &= ~(1u << S.tag);
    p0 = di; p1 = dj;
    ci += di; cj += dj;
    Qp[pdelta] = a + S.tag;
    emit();
    t++;
    if (t == n) {
      kind = WK_LEVEL;
      while (kind == WK_LEVEL) next_level();
    }
  }
};

#ifdef __CUDACC__

template <int MODE>
__global__ void k_build_unit_lut(uint8_t *lut) {  // any block size; once per context
  for (int e = threadIdx.x; e < TPR_LUT_ROWS * TPR_LUT_COLS; e += blockDim.x)
    lut[e] = unit_lut_entry<MODE>(e / TPR_LUT_COLS, e % TPR_LUT_COLS);
}

__global__ void k_build_t2(uint8_t *t2) {  // once per context; every slot of the table is written (the hash is a bijection)
  for (int job = blockIdx.x * blockDim.x + threadIdx.x; job < T2_JOBS; job += gridDim.x * blockDim.x) wk_t2_fill(t2, job);
}

__device__ __forceinline__ void load_unit_lut(uint8_t *s_lut, const uint8_t *g_lut) {  // the full 9 x 512 table
  const uint32_t *src = reinterpret_cast<const uint32_t *>(g_lut);
  uint32_t *dst = reinterpret_cast<uint32_t *>(s_lut);
  for (int e = threadIdx.x; e < TPR_LUT_ROWS * TPR_LUT_COLS / 4; e += blockDim.x) dst[e] = src[e];
}

// Bitmap geometry of region g in k1_walk: the bounding box with a margin of WK_PAD on every side
// (r0, c0 may be negative), ws words per row.
__device__ __forceinline__ void wk_geometry(const PathParams &P, int g, int &r0, int &c0, int &h, int &w, int &ws) {
  r0 = (P.reg.first[g] >> P.logW) - WK_PAD; c0 = P.reg.cmin[g] - WK_PAD;
  h = P.reg.rmax[g] - r0 + 1 + WK_PAD; w = P.reg.cmax[g] - c0 + 1 + WK_PAD; ws = (w + 31) >> 5;
}

// Cooperative build of ONE region's slot of an arena image (all arguments warp-uniform): the level-1 bitmap in plane 0
// and zeros in the rest of the slot.  dst = the slot's first word.  Word column by word column, 32 bitmap rows at a
// time: lane = column for the loads and ballots, lane = row for the result, which leaves as one store per 32 rows; per
// row an add, a load, a compare, a ballot and a select.  The rows and columns of the bounding box lie inside the image;
// the margin rows / columns stay zero.
__device__ __forceinline__ void wk_build_one(const PathParams &P, uint32_t *dst, int img_r, int label_r, int r0_r, int c0_r,
                                             int h_r, int w_r, int ws_r, int slot_r) {
  const int lane = (int)lane_id(), logW = P.logW;
  const int32_t *lab = P.labels + (size_t)img_r * P.N + ((r0_r + WK_PAD) << logW) + c0_r;  // first row of the bounding box
  const int nrows = h_r - 2 * WK_PAD;
  for (int wd = 0; wd < ws_r; wd++) {
    const int col = (wd << 5) + lane;
    const bool cok = col >= WK_PAD && col < w_r - WK_PAD;
    const int32_t *p = lab + col;
    for (int rb = 0; rb < h_r; rb += 32) {  // bitmap rows rb .. rb + 31 = bounding-box rows rb - WK_PAD ..
      const int i_lo = max(rb - WK_PAD, 0), i_hi = min(rb + 32 - WK_PAD, nrows);
      uint32_t mine = 0u;
      for (int i0 = i_lo; i0 < i_hi; i0 += 8) {  // eight independent label loads in flight per lane
        int lv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) lv[u] = (cok && i0 + u < i_hi) ? p[(size_t)(i0 + u) << logW] : ~label_r;
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const unsigned bits = __ballot_sync(FULL_MASK, lv[u] == label_r);
          mine = lane == i0 + u + WK_PAD - rb ? bits : mine;
        }
      }
      if (rb + lane < h_r) dst[(rb + lane) * ws_r + wd] = mine;
    }
  }
  for (int q = h_r * ws_r + lane; q < slot_r; q += 32) dst[q] = 0u;
}

// One chunk's arena image: lane r holds region r's geometry and its word offset `base` in `dst`; `slot` = its slot words.
__device__ __forceinline__ void wk_build_bitmaps(const PathParams &P, uint32_t *dst, int cnt, int img, int label, int r0,
                                                 int c0, int h, int w, int ws, int base, int slot) {
  for (int r = 0; r < cnt; r++) {
    const int img_r = __shfl_sync(FULL_MASK, img, r), label_r = __shfl_sync(FULL_MASK, label, r);
    const int r0_r = __shfl_sync(FULL_MASK, r0, r), c0_r = __shfl_sync(FULL_MASK, c0, r);
    const int h_r = __shfl_sync(FULL_MASK, h, r), w_r = __shfl_sync(FULL_MASK, w, r);
    const int ws_r = __shfl_sync(FULL_MASK, ws, r), base_r = __shfl_sync(FULL_MASK, base, r);
    const int slot_r = __shfl_sync(FULL_MASK, slot, r);
    wk_build_one(P, dst + base_r, img_r, label_r, r0_r, c0_r, h_r, w_r, ws_r, slot_r);
  }
}

// Chunk header shared by the bitmap builder and the walker: lane r < cnt holds region r of the chunk.
struct WkChunkLane {
  int g, img, label, first, size, off, r0, c0, h, w, ws, slot, base;
};

__device__ __forceinline__ WkChunkLane wk_chunk_lane(const PathParams &P, int qstart, int cnt) {
  const int lane = (int)lane_id();
  WkChunkLane c = {};
  if (lane < cnt) {
    c.g = P.queue[qstart + lane];
    c.img = P.reg.img[c.g]; c.label = P.reg.label[c.g]; c.first = P.reg.first[c.g];
    c.size = P.reg.size[c.g]; c.off = P.reg.off[c.g];
    wk_geometry(P, c.g, c.r0, c.c0, c.h, c.w, c.ws);
    c.slot = wk_slot_words(c.h, c.ws);
  }
  int inc = c.slot;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int y = __shfl_up_sync(FULL_MASK, inc, d);
    if (lane >= d) inc += y;
  }
  c.base = inc - c.slot;
  return c;
}

// The arena images of the small-bitmap chunks (the bulk kernel's: from QM_CHUNK_SPLIT on, the first P.gbm_chunks of
// them), built ahead of the walk by warps that do nothing else (the label reads are pure memory latency; inside the
// path kernel they would hold a walking warp's registers and arena).  gbm[chunk - split][TPR_ARENA_WORDS].  The
// large-bitmap chunks of the windowed instantiation build theirs in the kernel.
// Two kernels.  kq_slots: a warp per chunk writes, for every region of the chunk, where its slot lies in gbm (-1 for
// the regions of other chunks: the buffer is preset).  k1_bitmaps: a warp per region IN REGION ORDER, i.e. image by
// image and, inside an image, in first-appearance (row-major) order -- the warps running at the same time read the
// label rows of a handful of images, which stay in L2: a chunk's regions come from 32 different images, and building
// chunk by chunk fetched every bounding-box row from DRAM (3.3x the label bytes).
__global__ void __launch_bounds__(256) kq_slots(PathParams P) {
  const int split = P.qmeta[QM_CHUNK_SPLIT];
  const int nchunks = min(P.qmeta[QM_NCHUNKS], split + P.gbm_chunks);
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int chunk = split + wid; chunk < nchunks; chunk += nw) {
    const int cnt = P.chunk_cnt[chunk];
    const WkChunkLane c = wk_chunk_lane(P, P.chunk_start[chunk], cnt);
    if ((int)lane_id() < cnt) P.slot_of[c.g - P.g0] = (chunk - split) * TPR_ARENA_WORDS + c.base;
  }
}

__global__ void __launch_bounds__(256) k1_bitmaps(PathParams P) {
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int i = wid; i < P.nreg; i += nw) {
    const int so = P.slot_of[i];
    if (so < 0) continue;
    const int g = P.g0 + i;
    int r0, c0, h, w, ws;
    wk_geometry(P, g, r0, c0, h, w, ws);
    wk_build_one(P, P.gbm + so, P.reg.img[g], P.reg.label[g], r0, c0, h, w, ws, wk_slot_words(h, ws));
  }
}

#ifdef WK_STATS  // debug build only: warp trips and lane units by kind
__device__ unsigned long long g_wk_stats[16];  // [0] warp trips, [1 + kind] lanes of that kind at the start of a trip, [10] regions of the chunk, [11]/[12] searches beyond the window served (planes of <= 2 / more words per row), [13] bitmap-mode steps, [14] trips of the windowed instantiation
#endif

// The search beyond the 5x5 window of ONE lane's region, done by the whole warp (all arguments warp-uniform): the
// lanes take the rows of the bitmap plane.  Euclid: a row's candidate is its nearest unvisited column on either
// side (both when equidistant), ranked by the packed key (k << 21 | d2) and then the dot product; two warp reductions
// pick the winner.  Rows of one or two words are scanned whole in ONE pass -- ranking by (k, d2, dot) makes that
// equal to the reference's probes 4, 8, ... in turn; wider planes go window by window (paths.cuh, find_next_geo).
// Chebyshev: the reference's fp64 expressions (find_next).  Returns false if the plane holds no unvisited point;
// otherwise `step` = (di << 16) | (dj & 0xffff).
constexpr int WK_NO_PARTNER = (int)0x80000000;
#ifndef WK_FAR_RAD0
#define WK_FAR_RAD0 8
#endif

// The candidate of one bitmap row (x != 0, one or two words; row offset rdi from the current point): its nearest
// unvisited column on either side of cj -- the nearer one dominates the other in (k, d2); equidistant ones share the
// key and compete on the dot product (`partner` = the other one when the dot products tie as well).
template <typename WORD>
__device__ __forceinline__ void wk_row_candidate(WORD x, int cj, int rdi, int p0, int p1, unsigned &key, int &dot, int &off,
                                                 int &partner) {
  const WORD below = ((WORD)1 << cj) - (WORD)1;  // columns < cj
  const WORD left = x & below, right = x & ~below;
  int dl, dr;
  if (sizeof(WORD) == 4) { dl = cj - (31 - __clz((int)left)); dr = __ffs((int)right) - 1 - cj; }
  else { dl = cj - (63 - __clzll((long long)left)); dr = __ffsll((long long)right) - 1 - cj; }
  const bool both = left && right;
  const bool use_left = left && (!right || dl <= dr);
  int rdj = use_left ? -dl : dr;
  key = ((unsigned)probe_index(max(abs(rdi), abs(rdj))) << 21) | (unsigned)(rdi * rdi + rdj * rdj);
  dot = rdi * p0 + rdj * p1;
  partner = WK_NO_PARTNER;
  if (both && dl == dr) {  // (rdi, -dl) and (rdi, +dl): the other one's dot product is dot + 2 dl p1
    const int dot2 = dot + 2 * dr * p1;
    if (dot2 == dot) partner = (rdi << 16) | (dr & 0xffff);
    if (dot2 > dot) { dot = dot2; rdj = dr; }
  }
  off = (rdi << 16) | (rdj & 0xffff);
}

template <int MODE>
__device__ __forceinline__ bool wk_far_search(const uint32_t *plane, int h, int ws, int ci, int cj, int p0, int p1, int &step) {
  const int lane = (int)lane_id();
  if (MODE == MODE_EUCLID && ws <= 2) {
    unsigned bkey = 0xffffffffu;
    int bdot = 0, boff = 0, aoff = WK_NO_PARTNER;  // best candidate of this lane's rows, its mirror partner if any
    if (ws == 1 && h <= 32) {
      // the common case, straight-line: one row of one word per lane
      const uint32_t x = lane < h ? plane[lane] : 0u;
      if (x) wk_row_candidate<uint32_t>(x, cj, lane - ci, p0, p1, bkey, bdot, boff, aoff);
    } else {
      for (int i = lane; i < h; i += 32) {
        unsigned key;
        int dot, off, partner;
        if (ws == 1) {
          const uint32_t x = plane[i];
          if (!x) continue;
          wk_row_candidate<uint32_t>(x, cj, i - ci, p0, p1, key, dot, off, partner);
        } else {
          const unsigned long long x = ((unsigned long long)plane[2 * i + 1] << 32) | plane[2 * i];
          if (!x) continue;
          wk_row_candidate<unsigned long long>(x, cj, i - ci, p0, p1, key, dot, off, partner);
        }
        if (key < bkey || (key == bkey && dot > bdot)) { bkey = key; bdot = dot; boff = off; aoff = partner; }
        else if (key == bkey && dot == bdot) aoff = off;
      }
    }
    const unsigned kmin = __reduce_min_sync(FULL_MASK, bkey);
    if (kmin == 0xffffffffu) return false;
    const bool sel = bkey == kmin;
    const int dotmax = __reduce_max_sync(FULL_MASK, sel ? bdot : INT32_MIN);
    const unsigned tied = __ballot_sync(FULL_MASK, sel && bdot == dotmax);
    const int la = __ffs(tied) - 1;
    step = __shfl_sync(FULL_MASK, boff, la);
    int other = __shfl_sync(FULL_MASK, aoff, la);
    const unsigned rest = tied & (tied - 1);
    if (rest) other = __shfl_sync(FULL_MASK, boff, __ffs(rest) - 1);  // a mirror pair held by two lanes
    if (other != WK_NO_PARTNER &&
        mirror_second_wins(step >> 16, (int)(short)(step & 0xffff), other >> 16, (int)(short)(other & 0xffff),
                           (int)(kmin & 0x1fffffu), p0, p1))
      step = other;
    return true;
  }
  int bi, bj;
  if (MODE == MODE_EUCLID) {
    // half-width 8 at once: ranked by (k, d2, dot) one pass over the 17 rows equals the probes 4 and 8 in turn
    int rad = WK_FAR_RAD0;
    if (!find_next_geo(plane, h, ws << 5, ws, ci, cj, p0, p1, rad, nullptr, bi, bj)) return false;
  } else {
    double curval = 0.0;
    if (!find_next<MODE>(plane, h, ws << 5, ws, ci, cj, p0, p1, nullptr, 0, 0, 0, false, curval, bi, bj)) return false;
  }
  step = ((bi - ci) << 16) | ((bj - cj) & 0xffff);
  return true;
}

// the whole-warp walker with the bitmap window in registers (regwin.cuh, included at the end of this file)
__device__ void region_pyramid_rw(const PathParams &P, int g, uint32_t *bm, const uint8_t *t2);

// WIDEWIN = true: the instantiation for the chunks of large bitmaps (queue classes below Q_FIRST_NARROW_CLS; at most six
// regions per warp), which builds its bitmaps itself.
constexpr int wk_arena_words(bool widewin) { return widewin ? TPR_WIDE_ARENA_WORDS : TPR_ARENA_WORDS; }
constexpr size_t wk_arena_bytes(bool widewin) {
  return ((size_t)(widewin ? WK_WIDE_WARPS : WK_WARPS) * wk_arena_words(widewin) + 4) * sizeof(uint32_t);
}
static_assert(TPR_WIDE_ARENA_WORDS <= 8192 && TPR_MAX_SIDE + 2 * WK_PAD < 2048, "Walker::rq0 packs the plane offset in 13 bits, h in 11");

template <int MODE, bool WIDEWIN>
__global__ void __launch_bounds__((WIDEWIN ? WK_WIDE_WARPS : WK_WARPS) * 32, WIDEWIN ? WK_WIDE_MIN_CTAS : WK_MIN_CTAS)
    k1_walk(PathParams P) {
  constexpr int NWARPS = WIDEWIN ? WK_WIDE_WARPS : WK_WARPS;
  extern __shared__ __align__(16) uint32_t s_arena[];  // NWARPS * wk_arena_words(WIDEWIN) + 4 words (wk_arena_bytes)
  // the walker's step table: euclid the 5x5 table, chebyshev the compact unit-step table
  __shared__ __align__(16) uint8_t s_tab[MODE == MODE_EUCLID ? T2_BYTES : WK_LUT_BYTES];
  const int lane = (int)lane_id(), warp = threadIdx.x >> 5;
  constexpr int ARENA = WIDEWIN ? TPR_WIDE_ARENA_WORDS : TPR_ARENA_WORDS;
  uint32_t *arena = s_arena + warp * ARENA;
  // chunk table: class 1 (k1_coop_all's when the path mode has one: P.coop), classes 2..5 (windowed), 6.. (bulk)
  const int chunk_lo = WIDEWIN ? (P.coop ? P.qmeta[QM_CLS1_CHUNKS] : 0) : P.qmeta[QM_CHUNK_SPLIT];
  const int nchunks = (WIDEWIN ? P.qmeta[QM_CHUNK_SPLIT] : P.qmeta[QM_NCHUNKS]) - chunk_lo;
  if (nchunks <= 0) return;
  __shared__ __align__(8) unsigned long long s_mbar[NWARPS];  // one mbarrier per warp: arrival of its chunk's arena image
  uint32_t tma_phase = 0u;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_mbar[warp])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (MODE == MODE_EUCLID) {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(P.t2_tab);
    uint32_t *dst = reinterpret_cast<uint32_t *>(s_tab);
    for (int e = threadIdx.x; e < T2_BYTES / 4; e += blockDim.x) dst[e] = src[e];
  } else {
    for (int e = threadIdx.x; e < WK_LUT_BYTES; e += blockDim.x) s_tab[e] = P.unit_lut[wk_lut_source(e)];
  }
  __syncthreads();

  // Each warp takes its share of the chunks and retires: the grid is several waves of CTAs, so SM slots keep
  // freeing up for the (higher-priority) transform kernels of other units instead of being held to the end.
  const int share = max(1, (nchunks + (int)gridDim.x * NWARPS - 1) / ((int)gridDim.x * NWARPS));
  for (int taken = 0; taken < share; taken++) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(&P.qmeta[WIDEWIN ? QM_CUR_WIDE : QM_CUR_SMALL], 1);
    chunk = __shfl_sync(FULL_MASK, chunk, 0);
    if (chunk >= nchunks) break;
    chunk += chunk_lo;
    const int qstart = P.chunk_start[chunk], cnt = P.chunk_cnt[chunk];
    const WkChunkLane c = wk_chunk_lane(P, qstart, cnt);
    const bool mine = lane < cnt;
    if (!WIDEWIN && chunk - chunk_lo < P.gbm_chunks) {
      // the arena image was built by k1_bitmaps: one bulk copy (TMA, 1-D) global -> shared, completion on the warp's
      // mbarrier -- no registers, no issue slots of the walking warp spent on moving the 4..8 KB
      const int total = __shfl_sync(FULL_MASK, c.base + c.slot, 31);
      const uint32_t *src = P.gbm + (size_t)(chunk - chunk_lo) * TPR_ARENA_WORDS;
#ifndef WK_NO_TMA
      const uint32_t bytes = (uint32_t)((total + 3) >> 2) * 16u;
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_mbar[warp]);
      if (lane == 0) {
        // what the previous chunk's walk wrote to the arena (generic proxy) is ordered before the async-proxy writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(arena)), "l"(src), "r"(bytes), "r"(bar) : "memory");
      }
      uint32_t ready = 0;
      while (!ready) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ready) : "r"(bar), "r"(tma_phase) : "memory");
      }
      tma_phase ^= 1u;
#else
      const int nvec = (total + 3) >> 2;
      const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
      uint4 *dst = reinterpret_cast<uint4 *>(arena);
      for (int e0 = 0; e0 < nvec; e0 += 8 * 32) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int e = e0 + u * 32 + lane;
          v[u] = e < nvec ? __ldcs(src4 + e) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int e = e0 + u * 32 + lane;
          if (e < nvec) dst[e] = v[u];
        }
      }
#endif
    } else {
      wk_build_bitmaps(P, arena, cnt, c.img, c.label, c.r0, c.c0, c.h, c.w, c.ws, c.base, c.slot);
    }
    __syncwarp();

    Walker<MODE> wk;
    wk.bm = arena + c.base;
    wk.Qall = P.Q; wk.pdelta = P.Pm - P.Q; wk.img = c.img;
    wk.lut = MODE == MODE_EUCLID ? nullptr : s_tab;
    wk.t2 = MODE == MODE_EUCLID ? s_tab : nullptr;
    wk.N = P.N; wk.W = P.W; wk.L = P.levels;
    wk.abase = c.base;
    wk.h = c.h; wk.ws = max(c.ws, 1);
    wk.pixbase = c.r0 * P.W + c.c0;
    wk.narrow = __all_sync(FULL_MASK, c.ws <= 1);
    wk.kind = WK_DONE;
    if (mine) wk.start(c.off, c.size, WK_PAD, (c.first & (P.W - 1)) - c.c0);
    // Every lane walks its own region through all its levels.  One round of the warp loop = one step for every lane
    // that can take one: lanes choose their step from the 5x5 window; the warp searches beyond the window for each
    // lane that found it empty; all of them commit together.  Lanes at the end of a level start the next; list-mode
    // lanes take a step when enough of them are waiting (or nothing else in the warp can move).
    while (true) {
      const unsigned live = __ballot_sync(FULL_MASK, !wk.done());
      if (!live) break;
#ifdef WK_STATS
      if (lane == 0) atomicAdd(&g_wk_stats[0], 1ull);
      atomicAdd(&g_wk_stats[1 + wk.kind], 1ull);
      if (lane == 0) atomicAdd(&g_wk_stats[10], (unsigned long long)cnt);
      if (lane == 0 && WIDEWIN) atomicAdd(&g_wk_stats[14], 1ull);
#endif
      if (__any_sync(FULL_MASK, wk.kind == WK_LEVEL)) {
        if (wk.kind == WK_LEVEL) wk.next_level();
      }
#pragma unroll 1
      for (int rep = 0; rep < WK_NEAR_REPS; rep++) {
        const bool near = wk.kind == WK_NEAR;
        if (!__any_sync(FULL_MASK, near)) break;
        if (near) wk.near_select();
        unsigned farm = __ballot_sync(FULL_MASK, wk.kind == WK_FAR);
        while (farm) {
          const int src = __ffs(farm) - 1;
          farm &= farm - 1;
          // the requester's plane (word offset in the warp's arena), geometry, current point and pref, warp-uniform
          const int pa = __shfl_sync(FULL_MASK, wk.rq0, src);
          const int pb = __shfl_sync(FULL_MASK, wk.ci | (wk.cj << 16), src);
          const int pc = __shfl_sync(FULL_MASK, (wk.p0 & 0xffff) | (wk.p1 << 16), src);
#ifdef WK_STATS
          if (lane == 0) atomicAdd(&g_wk_stats[(pa >> 24) <= 2 ? 11 : 12], 1ull);
#endif
          int step = 0;
          const bool ok = wk_far_search<MODE>(arena + (pa & 0x1fff), (pa >> 13) & 0x7ff, pa >> 24, pb & 0xffff, pb >> 16,
                                              (int)(short)(pc & 0xffff), pc >> 16, step);
          if (lane == src) {
            if (ok) wk.far_found(step >> 16, (int)(short)(step & 0xffff));
            else wk.kind = WK_ERROR;
          }
        }
#ifdef WK_STATS
        if (wk.kind == WK_COMMIT) atomicAdd(&g_wk_stats[13], 1ull);
#endif
        if (wk.kind == WK_COMMIT) wk.commit_step();
      }
      const unsigned listm = __ballot_sync(FULL_MASK, wk.kind == WK_LIST);
      if (listm && (__popc(listm) >= WK_LIST_BATCH || !__ballot_sync(FULL_MASK, wk.kind == WK_NEAR || wk.kind == WK_LEVEL))) {
        if (wk.kind == WK_LIST) wk.list_step();
      }
    }
    if (wk.kind == WK_ERROR) atomicExch(&P.qmeta[QM_ERR], 1);
    __syncwarp();
  }
}

// Every region of a group walked by a whole warp, one region per chunk.  Euclid: small groups (single images, small
// batches -- latency is all that matters there), through the register-window walker (regwin.cuh).  gradpath (paths.cuh,
// region_pyramid / find_next_grad): always -- the path type exists for parity with the reference, not for throughput.
template <int MODE>
__global__ void __launch_bounds__(COOP_WARPS * 32) k1_coop_all(PathParams P) {
  __shared__ __align__(16) uint32_t s_arena[COOP_WARPS * TPR_ARENA_WORDS];
  __shared__ __align__(16) uint8_t s_t2[MODE == MODE_EUCLID ? T2_BYTES : 16];
  const int lane = (int)lane_id(), warp = threadIdx.x >> 5;
  const int nchunks = P.qmeta[QM_CLS1_CHUNKS];  // class 1 = the regions walked by a warp (all of them in a small group)
  if (nchunks == 0) return;
  if (MODE == MODE_EUCLID) {
    for (int e = threadIdx.x; e < T2_BYTES / 4; e += blockDim.x)
      reinterpret_cast<uint32_t *>(s_t2)[e] = reinterpret_cast<const uint32_t *>(P.t2_tab)[e];
    __syncthreads();
  }
  while (true) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(&P.qmeta[QM_CUR_COOP], 1);
    chunk = __shfl_sync(FULL_MASK, chunk, 0);
    if (chunk >= nchunks) break;
    const int g = P.queue[P.chunk_start[chunk]];
    if (MODE == MODE_EUCLID) region_pyramid_rw(P, g, s_arena + warp * TPR_ARENA_WORDS, s_t2);
    else region_pyramid<MODE>(P, g, s_arena + warp * TPR_ARENA_WORDS, nullptr);
    __syncwarp();
  }
}

// Big regions: one warp per CTA, bitmap in dynamic shared memory if it fits, else global scratch.
template <int MODE>
__global__ void __launch_bounds__(32) k1_paths_big(PathParams P) {
  extern __shared__ uint32_t s_big[];
  __shared__ __align__(16) uint8_t s_lut[MODE == MODE_EUCLID ? T2_BYTES : TPR_LUT_ROWS * TPR_LUT_COLS];
  const int lane = (int)lane_id();
  if (P.qmeta[QM_NBIG] == 0) return;  // the common case: nothing oversized in this group
  if (MODE == MODE_EUCLID) {
    for (int e = threadIdx.x; e < T2_BYTES / 4; e += blockDim.x)
      reinterpret_cast<uint32_t *>(s_lut)[e] = reinterpret_cast<const uint32_t *>(P.t2_tab)[e];
  } else {
    load_unit_lut(s_lut, P.unit_lut);
  }
  __syncthreads();
  const int nbig = P.qmeta[QM_NBIG];
  uint32_t *gs = P.gscratch + (size_t)blockIdx.x * P.gscratch_words;
  while (true) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(&P.qmeta[QM_CUR_BIG], 1);
    idx = __shfl_sync(FULL_MASK, idx, 0);
    if (idx >= nbig) break;
    const int g = P.queue[idx];
    const int words = region_bitmap_words(P.reg, g, P.logW);
    if (MODE == MODE_EUCLID) region_pyramid_rw(P, g, words <= P.big_smem_words ? s_big : gs, s_lut);
    else region_pyramid<MODE>(P, g, words <= P.big_smem_words ? s_big : gs, s_lut);
    __syncwarp();
  }
}

// K2: positions in the incoming order.  Pm[level][a + t] = place, in the level's incoming order, of the t-th point of
// the level's path (= the reference's generating permutation + the region offset, rbepwt.py:1285, 1333), which is what
// the transform kernels gather / scatter through.  The incoming order of level l >= 2 is the path order of level
// l-1 subsampled at the even global positions, so with pos[pixel] = (a' + t') >> 1 for the even a' + t' of level l-1:
// Pm_l[a + t] = pos[Q_l[a + t]].  One warp per region, level after level; pos lives in shared memory, relative to the
// region (16 bits per cell of the bounding box), or -- bounding boxes of more than K2_CELLS cells -- in the image's
// `posmap` in global memory.  The thread-per-region walker already wrote the positions of its list-mode levels (at
// most WK_LIST_MAX points: the list index is the position), so for its regions only the levels above are done here;
// regions walked by a whole warp (paths.cuh) get every level.
constexpr int K2_WARPS = 8;
constexpr int K2_CELLS = 1024;  // 16-bit cells per warp (2 KB): 64 warps per SM stay resident
constexpr int K2_VPL = 8;       // path points per lane held in registers: a level of up to 256 points is loaded once

// the first K2_VPL * 32 points of a level's path, lane-strided; -1 beyond n
__device__ __forceinline__ void k2_load(const int32_t *Ql, int n, int lane, int (&v)[K2_VPL]) {
#pragma unroll
  for (int u = 0; u < K2_VPL; u++) {
    if (u * 32 >= n) break;
    const int t = u * 32 + lane;
    v[u] = t < n ? __ldcs(Ql + t) : -1;
  }
}

__global__ void __launch_bounds__(K2_WARPS * 32) k2_perm(PathParams P, int nreg, int coop) {
  __shared__ uint16_t s_pos[K2_WARPS][K2_CELLS];
  const int lane = (int)lane_id(), warp = threadIdx.x >> 5;
  const int nw = gridDim.x * K2_WARPS;
  const int N = P.N, logW = P.logW, Wm = P.W - 1, L = P.levels;
  const int coop_size = P.qmeta[QM_COOP_SIZE];
  for (int q = blockIdx.x * K2_WARPS + warp; q < nreg; q += nw) {
    const int g = P.queue[q];  // queue order: the largest regions first
    const int img = P.reg.img[g], a1 = P.reg.off[g], n1 = P.reg.size[g];
    const int r0 = P.reg.first[g] >> logW, c0 = P.reg.cmin[g];
    const int hb = P.reg.rmax[g] - r0 + 1, wb = P.reg.cmax[g] - c0 + 1;
    const bool in_smem = hb * wb <= K2_CELLS && n1 < 65536;
    // levels of at most `skip_below` points were done by the walker (never level 1); 0 = this region was walked by a warp
    // (coop: the path mode hands regions of at least qmeta[QM_COOP_SIZE] pixels to the whole-warp walker)
    const bool by_warp = walked_by_warp(n1, region_class_words(P.reg, g, logW), coop_size, coop != 0);
    const int skip_below = by_warp ? 0 : WK_LIST_MAX;
    const int32_t *Qimg = P.Q + (size_t)img * 2 * (size_t)N;
    int32_t *Pimg = P.Pm + (size_t)img * 2 * (size_t)N;
    int32_t *posmap = P.posmap + (size_t)img * N;
    uint16_t *pos = s_pos[warp];
    auto cell = [&](int pix) { return ((pix >> logW) - r0) * wb + (pix & Wm) - c0; };
    int cur[K2_VPL], nxt[K2_VPL];
    int a = a1, n = n1;
    if (((a + n + 1) >> 1) - ((a + 1) >> 1) <= skip_below || L < 2) continue;  // no level >= 2 to do
    k2_load(Qimg + a, n, lane, cur);
    for (int lev = 1; lev <= L; lev++) {
      const int32_t *Ql = Qimg + level_off((size_t)N, lev) + a;
      const int an = (a + 1) >> 1;
      int nn = lev < L ? ((a + n + 1) >> 1) - an : 0;
      if (nn <= skip_below) nn = 0;  // the next level is the walker's (or does not exist): this is the last one here
      // the next level's path is requested before this level is processed: one memory latency per region, not per level
      if (nn > 0) k2_load(Qimg + level_off((size_t)N, lev + 1) + an, nn, lane, nxt);
      if (lev >= 2) {
        int32_t *Pl = Pimg + level_off((size_t)N, lev) + a;
#pragma unroll
        for (int u = 0; u < K2_VPL; u++) {
          if (u * 32 >= n) break;
          const int t = u * 32 + lane;
          if (t < n) Pl[t] = in_smem ? a + (int)pos[cell(cur[u])] : __ldcg(posmap + cur[u]);
        }
        for (int t = K2_VPL * 32 + lane; t < n; t += 32) {
          const int pix = __ldcs(Ql + t);
          Pl[t] = in_smem ? a + (int)pos[cell(pix)] : __ldcg(posmap + pix);
        }
      }
      if (nn <= 0) break;
      __syncwarp();
      // the even global positions a + t survive: their place in the next level's incoming order
#pragma unroll
      for (int u = 0; u < K2_VPL; u++) {
        if (u * 32 >= n) break;
        const int t = u * 32 + lane;
        if (t < n && ((a + t) & 1) == 0) {
          const int place = (a + t) >> 1;
          if (in_smem) pos[cell(cur[u])] = (uint16_t)(place - an);
          else posmap[cur[u]] = place;
        }
      }
      for (int t = K2_VPL * 32 + lane; t < n; t += 32)
        if (((a + t) & 1) == 0) {
          const int pix = __ldcs(Ql + t), place = (a + t) >> 1;
          if (in_smem) pos[cell(pix)] = (uint16_t)(place - an);
          else posmap[pix] = place;
        }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < K2_VPL; u++) cur[u] = nxt[u];
      a = an; n = nn;
    }
    __syncwarp();
  }
}

#endif  // __CUDACC__

}  // namespace rbepwt

#ifdef __CUDACC__
#include "regwin.cuh"
#endif
