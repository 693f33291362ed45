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
// Instead one trip of the warp loop lets every lane do one bounded unit of its current search (a few
// table steps, a few window rows or bitmap words, one list step); a lane that exhausts its window commits
// the step (or doubles the window) and starts the next search on the following trip, independently of
// its neighbours.
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
// most two candidates (usually one: the nearer side), found with clz/ffs, and a candidate is a 32-bit key
// (k << 21 | d2) plus its dot product -- valid because the kernel only takes regions whose bounding box
// has sides <= TPR_MAX_SIDE = 1024 (d2 < 2^21).  Updates are selects: all lanes execute the same instructions.
//
// Unit-step fast path.  When the window half-width is 1 and pref is one of the 8 unit steps (the
// common case at level 1: 88 % of the steps), the 3x3 neighbourhood is gathered into a mask and the
// answer is read from a table in shared memory (9 prefs x 256 masks: the centre bit is always empty).  The
// table is filled once per context by the same candidate code the generic path runs, so it cannot disagree
// with it.  Every bitmap carries a margin of TPR_PAD = 1 empty row and column around the bounding box, so
// the fetch is three shared loads and shifts without bounds checks.
//
// Five rows per trip.  For half-widths <= 15 a window row is one 32-bit word after a funnel shift
// that puts column cj at bit 15, so a trip examines up to five rows (a whole half-width-2 window); wider
// windows (scattered regions) fall back to eight bitmap words per trip.
//
// List mode.  From the first level at which a region keeps <= 32 points (spacing large, windows wide
// and mostly empty) the lane drops the bitmap and holds the points as a list of packed (row, col) in
// its arena slot; a step scans the remaining points (a 32-bit unvisited mask) with the same candidate
// code, so the cost per step is the number of points left instead of the window area.
//
// Memory latency kept off the walk.  The chunk's bitmaps are built from the labels by k1_bitmaps, a kernel
// of its own (global memory, arena layout), and copied into the arena with 16-byte loads; at the end of a level
// ONE pass over the level's path writes the incoming-order positions (levels >= 2) and re-marks the survivors,
// 16 points per trip with every load in flight.
//
// 5x5 table step (-DTPR_TABLE5, off: measured slower).  The probes of half-width 1 and 2
// resolved by three table reads with the same instructions for every lane.
//
// Shared memory: one arena of TPR_ARENA_WORDS words per warp holds the bounding-box bitmaps of the
// chunk's regions (chunk table: regions.cuh; a chunk always fits).
#pragma once
#include "paths.cuh"

namespace rbepwt {

constexpr int TPR_WARPS = 4;
#ifndef TPR_MIN_CTAS
#define TPR_MIN_CTAS 5  // measured: capping registers for a 6th CTA per SM spills and is slower
#endif

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

  __device__ __forceinline__ void reset() { key = 0xffffffffu; dot = 0; off = aoff = 0; tag = atag = 0; alt = false; }
  __device__ __forceinline__ bool have() const { return key != 0xffffffffu; }

  __device__ __forceinline__ void consider(bool valid, int cdi, int cdj, int p0, int p1, int ctag = 0) {
    const int k = probe_index(max(abs(cdi), abs(cdj)));
    const unsigned ckey = valid ? ((unsigned)k << 21) | (unsigned)(cdi * cdi + cdj * cdj) : 0xffffffffu;
    const int cdot = cdi * p0 + cdj * p1;
    const int coff = (cdi << 16) | (cdj & 0xffff);
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
  __device__ __forceinline__ void row_candidates(bool hl, int dl, bool hr, int dr, int rdi, int p0, int p1) {
    if (!(hl || hr)) return;
    const bool left_first = hl && (!hr || dl <= dr);
    consider(true, rdi, left_first ? -dl : dr, p0, p1);
    if (hl && hr && dl == dr) consider(true, rdi, dr, p0, p1);
  }

  // one bitmap word of row ci+rdi: columns lo..lo+31, already masked to the window
  __device__ __forceinline__ void scan_word(uint32_t bits, int lo, int rdi, int cj, int p0, int p1) {
    const int rel = min(cj - lo, 31);
    const uint32_t lmask = rel < 0 ? 0u : (2u << rel) - 1u;  // columns <= cj
    const uint32_t left = bits & lmask, right = bits & ~lmask;
    row_candidates(left != 0u, cj - (lo + 31 - __clz(left)), right != 0u, lo + __ffs(right) - 1 - cj, rdi, p0, p1);
  }

  // one window row as an aligned word: bit 15 + dj <-> column cj + dj
  __device__ __forceinline__ void scan_row(uint32_t x, int rdi, int p0, int p1) {
    const uint32_t left = x & 0x7fffu, right = x >> 15;
    row_candidates(left != 0u, __clz(left) - 16, right != 0u, __ffs(right) - 1, rdi, p0, p1);  // hb = 31 - clz -> dl = 15 - hb
  }

  __device__ __forceinline__ void finish(int p0, int p1, int &odi, int &odj, int &k) {
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

  __device__ __forceinline__ void reset() { found = false; has_sp1 = false; c = d2 = di = dj = tag = 0; sp1 = 0.0; }
  __device__ __forceinline__ bool have() const { return found; }

  __device__ __forceinline__ void consider(bool valid, int cdi, int cdj, int p0, int p1, int ctag = 0) {
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

  __device__ __forceinline__ void scan_word(uint32_t bits, int lo, int rdi, int cj, int p0, int p1) {
    while (bits) {
      const int j = lo + __ffs(bits) - 1;
      bits &= bits - 1;
      consider(true, rdi, j - cj, p0, p1);
    }
  }

  __device__ __forceinline__ void scan_row(uint32_t x, int rdi, int p0, int p1) {
    while (x) {
      const int b = __ffs(x) - 1;
      x &= x - 1;
      consider(true, rdi, b - 15, p0, p1);
    }
  }

  __device__ __forceinline__ void finish(int, int, int &odi, int &odj, int &k) {
    odi = di; odj = dj; k = probe_index(c);
  }
};


#ifdef TPR_STATS  // debug build only: trips and lanes served per level and trip kind
__device__ unsigned long long g_tpr_stats[16 * 4 * 2 + 32];
#endif
constexpr int TPR_LIST_MAX = 32;   // list mode from the first level with at most this many points
constexpr int TPR_SLOT_MIN = 48;   // arena words per lane: list buffers A = [0,32), B = [32,48) ping-pong
#ifndef TPR_ROWS_N
#define TPR_ROWS_N 5
#endif
constexpr int TPR_ROWS_PER_TRIP = TPR_ROWS_N;
#ifndef TPR_UNIT_STEPS_N
#define TPR_UNIT_STEPS_N 4
#endif
constexpr int TPR_UNIT_STEPS = TPR_UNIT_STEPS_N;  // unit steps a lane may take per trip
#ifndef TPR_S5_STEPS_N
#define TPR_S5_STEPS_N 2
#endif
constexpr int TPR_S5_STEPS = TPR_S5_STEPS_N;      // 5x5 table steps a lane may take per trip
#ifndef TPR_P2_MIN_N
#define TPR_P2_MIN_N 1
#endif
constexpr int TPR_P2_MIN = TPR_P2_MIN_N;          // lanes that must be waiting before the window / list code runs
// The large-bitmap instantiation walks 1..8 long chains per warp: trip overhead dominates, so a trip does more.
#ifndef TPR_WIDE_UNIT_N
#define TPR_WIDE_UNIT_N 8
#endif
constexpr int TPR_UNIT_STEPS_WIDE = TPR_WIDE_UNIT_N;
constexpr int TPR_MAX_RAD = 8;     // widest aligned-row window; beyond it the whole bitmap is scanned
#ifndef TPR_SCAN_WORDS_N
#define TPR_SCAN_WORDS_N 8
#endif
constexpr int TPR_SCAN_WORDS = TPR_SCAN_WORDS_N;  // ... that many words per trip

// Unit-step table: lut[q * 512 + m] = index (di+1)*3 + (dj+1) of the winner among the neighbours present in the
// 9-bit mask m, for pref = (q/3 - 1, q%3 - 1).  Filled by the candidate code every other path takes.
template <int MODE>
__device__ __forceinline__ void build_unit_lut(uint8_t *lut) {  // any block size; used once per context
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
    lut[e] = v;
  }
}

// ---- 5x5 table step (euclid mode, small bitmaps) ---------------------------------------------------
// The probes of half-width 1 and 2 see at most the 24 neighbours of the 5x5 window, and the key orders them by
// d2 first: the classes d2 = 1, 2, 4, 5, 8 (4, 4, 4, 8, 4 cells).  The step is then three table reads, the
// same instructions for every lane whatever its neighbourhood looks like:
//   T1[row][5 window bits]  -> the row's cells as bits of a class-grouped code (A: bits 0-3, B: 4-7, C: 8-11,
//                              D: 12-19, E: 20-23); the five rows are OR-ed;
//   lowest set bit          -> the nearest class present and its field of the code;
//   T2[pref][class, field]  -> the winning cell, for the 24 prefs that are themselves 5x5 offsets (the previous
//                              step was such a step, or the level starts: 84 % of all steps of the benchmark).
// Both tables are filled by the candidate code every other path runs (Search<MODE_EUCLID>), fp64 tie-break
// included, so they cannot disagree with it.  The bitmap carries a margin of TPR_PAD = 2 empty rows and
// columns, so the window never leaves it.
constexpr int S5_T2_ROW = 320;                   // A 16 + B 16 + C 16 + D 256 + E 16 entries per pref
constexpr int S5_T1_WORDS = 5 * 32;              // u32 [row][bits]
constexpr int S5_CELL_WORDS = 8;                 // u8 [code bit] -> cell index (di+2)*5 + (dj+2), 24 used
constexpr int S5_T2_WORDS = 25 * S5_T2_ROW / 4;  // u8 [pref cell][S5_T2_ROW]
constexpr int S5_WORDS = S5_T1_WORDS + S5_CELL_WORDS + S5_T2_WORDS;

// code bit of the cell (di, dj), -1 for the centre: classes by d2, cells of a class in row-major order
__host__ __device__ constexpr int s5_code_bit(int di, int dj) {
  const int d2 = di * di + dj * dj;
  if (d2 == 0) return -1;
  const int base = d2 == 1 ? 0 : d2 == 2 ? 4 : d2 == 4 ? 8 : d2 == 5 ? 12 : 20;
  int before = 0;
  for (int q = 0; q < (di + 2) * 5 + (dj + 2); q++) {
    const int qi = q / 5 - 2, qj = q % 5 - 2;
    if (qi * qi + qj * qj == d2) before++;
  }
  return base + before;
}

__device__ __forceinline__ void s5_field(uint32_t code, int &shift, uint32_t &fld, int &base) {
  const int f = __ffs(code) - 1, g4 = f & ~3;
  const bool isD = (unsigned)(f - 12) < 8u;
  shift = isD ? 12 : g4;
  fld = (code >> shift) & (isD ? 255u : 15u);
  base = isD ? 48 : (f >= 20 ? 304 : g4 * 4);
}

__global__ void k_build_s5_tables(uint32_t *tab) {
  uint8_t *cell_of_bit = reinterpret_cast<uint8_t *>(tab + S5_T1_WORDS);
  uint8_t *t2 = reinterpret_cast<uint8_t *>(tab + S5_T1_WORDS + S5_CELL_WORDS);
  for (int e = threadIdx.x; e < S5_T1_WORDS; e += blockDim.x) {
    const int r = e >> 5, b = e & 31;
    uint32_t code = 0;
    for (int c = 0; c < 5; c++)
      if ((b >> c) & 1) {
        const int bit = s5_code_bit(r - 2, c - 2);
        if (bit >= 0) code |= 1u << bit;
      }
    tab[e] = code;
  }
  for (int q = threadIdx.x; q < 32; q += blockDim.x) {
    if (q < 25) {
      const int bit = s5_code_bit(q / 5 - 2, q % 5 - 2);
      if (bit >= 0) cell_of_bit[bit] = (uint8_t)q;
    }
    if (q >= 24) cell_of_bit[q] = 0xff;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 25 * S5_T2_ROW; e += blockDim.x) {
    const int pid = e / S5_T2_ROW, idx = e % S5_T2_ROW;
    const int p0 = pid / 5 - 2, p1 = pid % 5 - 2;
    int shift, nb;
    uint32_t fld;
    if (idx < 48) { shift = (idx >> 4) * 4; fld = idx & 15; nb = 4; }
    else if (idx < 304) { shift = 12; fld = idx - 48; nb = 8; }
    else { shift = 20; fld = idx - 304; nb = 4; }
    uint8_t v = 0xff;
    if (pid != 12 && fld) {
      Search<MODE_EUCLID> S;
      S.reset();
      for (int b = 0; b < nb; b++)
        if ((fld >> b) & 1) {
          const int q = cell_of_bit[shift + b];
          S.consider(true, q / 5 - 2, q % 5 - 2, p0, p1);
        }
      int odi, odj, k;
      S.finish(p0, p1, odi, odj, k);
      v = (uint8_t)((odi + 2) * 5 + (odj + 2));
    }
    t2[e] = v;
  }
}

// The tables are computed once per context (global memory, one per path mode); the path kernels copy theirs
// into shared memory.
template <int MODE>
__global__ void k_build_unit_lut(uint8_t *lut) { build_unit_lut<MODE>(lut); }

__device__ __forceinline__ void load_unit_lut(uint8_t *s_lut, const uint8_t *g_lut) {
  const uint32_t *src = reinterpret_cast<const uint32_t *>(g_lut);
  uint32_t *dst = reinterpret_cast<uint32_t *>(s_lut);
  for (int e = threadIdx.x; e < TPR_LUT_ROWS * TPR_LUT_COLS / 4; e += blockDim.x) dst[e] = src[e];
}

// Arena words of a region's slot: its bitmap, at least the two list buffers, and an ODD number, so that the slots
// of a warp's 32 lanes start in different shared-memory banks (a stride of 48 would leave them two banks).
__device__ __forceinline__ int tpr_slot_words(int words) { return max(words, TPR_SLOT_MIN) | 1; }

// Bitmap geometry of region g in k1_paths_tpr: the bounding box with a margin of TPR_PAD on every side
// (r0, c0 may be negative), ws words per row.
__device__ __forceinline__ void tpr_geometry(const PathParams &P, int g, int &r0, int &c0, int &h, int &w, int &ws) {
  r0 = (P.reg.first[g] >> P.logW) - TPR_PAD; c0 = P.reg.cmin[g] - TPR_PAD;
  h = P.reg.rmax[g] - r0 + 1 + TPR_PAD; w = P.reg.cmax[g] - c0 + 1 + TPR_PAD; ws = (w + 31) >> 5;
}

// Cooperative bitmap build of one chunk: one ballot per bitmap word, lanes = columns; lane r holds region r's
// geometry and its word offset `base` in `dst`.
__device__ __forceinline__ void tpr_build_bitmaps(const PathParams &P, uint32_t *dst, int cnt, int img, int label, int r0,
                                                  int c0, int h, int w, int ws, int base) {
  const int lane = (int)lane_id(), logW = P.logW;
  for (int r = 0; r < cnt; r++) {
    const int img_r = __shfl_sync(FULL_MASK, img, r), label_r = __shfl_sync(FULL_MASK, label, r);
    const int r0_r = __shfl_sync(FULL_MASK, r0, r), c0_r = __shfl_sync(FULL_MASK, c0, r);
    const int h_r = __shfl_sync(FULL_MASK, h, r), w_r = __shfl_sync(FULL_MASK, w, r);
    const int ws_r = __shfl_sync(FULL_MASK, ws, r), base_r = __shfl_sync(FULL_MASK, base, r);
    const int32_t *lab = P.labels + (size_t)img_r * P.N;
    const int words_r = h_r * ws_r;
    if (ws_r == 1) {
      // one word per row (the usual case): lane = column, the margin rows are known to be empty
      const bool colok = lane < w_r && (unsigned)(c0_r + lane) < (unsigned)P.W;
      const int32_t *pp = lab + c0_r + lane;
      if (lane < 2 * TPR_PAD) dst[base_r + (lane < TPR_PAD ? lane : h_r - 2 * TPR_PAD + lane)] = 0u;
      for (int wi = TPR_PAD; wi < h_r - TPR_PAD; wi += 8) {  // eight independent label loads in flight per lane
        int lv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int row = r0_r + wi + u;
          const bool ok = colok && wi + u < h_r - TPR_PAD && (unsigned)row < (unsigned)P.H;
          lv[u] = ok ? pp[row << logW] : ~label_r;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const unsigned bits = __ballot_sync(FULL_MASK, lv[u] == label_r);
          if (lane == 0 && wi + u < h_r - TPR_PAD) dst[base_r + wi + u] = bits;
        }
      }
      continue;
    }
    for (int wi = 0; wi < words_r; wi += 8) {  // eight independent label loads in flight per lane
      int lv[8];
      bool inb[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int w_ = wi + u;
        const int i = w_ / ws_r;
        const int col = ((w_ - i * ws_r) << 5) + lane;
        inb[u] = w_ < words_r && col < w_r && (unsigned)(r0_r + i) < (unsigned)P.H && (unsigned)(c0_r + col) < (unsigned)P.W;
        lv[u] = inb[u] ? lab[((r0_r + i) << logW) + c0_r + col] : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const unsigned bits = __ballot_sync(FULL_MASK, inb[u] && lv[u] == label_r);
        if (lane == 0 && wi + u < words_r) dst[base_r + wi + u] = bits;
      }
    }
  }
}

// The bitmaps of the small-bitmap chunks (the bulk kernel's: from QM_CHUNK_SPLIT on, the first P.gbm_chunks of them),
// built ahead of the walk by warps that do nothing else (the label reads are pure memory latency; inside the path
// kernel they would hold a walking warp's registers and arena).  gbm[chunk - split][TPR_ARENA_WORDS]: the arena
// image k1_paths_tpr copies.  The large-bitmap chunks of the windowed instantiation build theirs in the kernel:
// one warp needs hundreds of dependent rounds for a 2000-word bitmap, and every path kernel would wait for it.
__global__ void __launch_bounds__(256) k1_bitmaps(PathParams P) {
  const int lane = (int)lane_id();
  const int split = P.qmeta[QM_CHUNK_SPLIT];
  const int nchunks = min(P.qmeta[QM_NCHUNKS], split + P.gbm_chunks);
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int chunk = split + wid; chunk < nchunks; chunk += nw) {
    const int qstart = P.chunk_start[chunk], cnt = P.chunk_cnt[chunk];
    int img = 0, label = 0, r0 = 0, c0 = 0, h = 0, w = 0, ws = 0;
    if (lane < cnt) {
      const int g = P.queue[qstart + lane];
      img = P.reg.img[g]; label = P.reg.label[g];
      tpr_geometry(P, g, r0, c0, h, w, ws);
    }
    const int slot = lane < cnt ? tpr_slot_words(h * ws) : 0;
    int inc = slot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(FULL_MASK, inc, d);
      if (lane >= d) inc += y;
    }
    tpr_build_bitmaps(P, P.gbm + (size_t)(chunk - split) * TPR_ARENA_WORDS, cnt, img, label, r0, c0, h, w, ws, inc - slot);
  }
}

// WIDEWIN = true: the instantiation for the chunks of large bitmaps (queue classes below Q_FIRST_NARROW_CLS):
// beyond half-width TPR_MAX_RAD it scans windows of half-width 16, 32, ... word by word instead of the whole
// bitmap, which for a 10^4-pixel region is the difference between tens and thousands of words per far jump.
// The common instantiation (small bitmaps) keeps the flat whole-bitmap scan and stays compact.
template <int MODE, bool WIDEWIN>
__global__ void __launch_bounds__(TPR_WARPS * 32, TPR_MIN_CTAS) k1_paths_tpr(PathParams P) {
#ifdef TPR_TABLE5  // measured slower on the benchmark (its tables cost shared memory, hence L1): off by default
  constexpr bool S5 = MODE == MODE_EUCLID && !WIDEWIN;  // 5x5 table steps beside the unit-step table
#else
  constexpr bool S5 = false;
#endif
  constexpr bool COMPACT = !WIDEWIN;  // unit-step table without the always-empty centre bit, margin-based row fetch
  __shared__ __align__(16) uint32_t s_arena[TPR_WARPS * TPR_ARENA_WORDS + 4];  // + 1: the table step reads one word past a row
  // unit-step table (table-step variant: compact, the always-empty centre bit dropped from the mask), 5x5 tables
  constexpr int ULUT_BYTES = COMPACT ? TPR_LUT_ROWS * 256 : TPR_LUT_ROWS * TPR_LUT_COLS;
  __shared__ __align__(16) uint8_t s_lut[ULUT_BYTES + (S5 ? S5_WORDS * 4 : 0)];
  const int lane = (int)lane_id(), warp = threadIdx.x >> 5;
  uint32_t *arena = s_arena + warp * TPR_ARENA_WORDS;
  const uint8_t *s5_base = s_lut + ULUT_BYTES;
  const uint32_t *s5_t1 = reinterpret_cast<const uint32_t *>(s5_base);
  const uint8_t *s5_cell = s5_base + S5_T1_WORDS * 4;
  const uint8_t *s5_t2 = s5_base + (S5_T1_WORDS + S5_CELL_WORDS) * 4;
  const int chunk_lo = WIDEWIN ? 0 : P.qmeta[QM_CHUNK_SPLIT];
  const int nchunks = (WIDEWIN ? P.qmeta[QM_CHUNK_SPLIT] : P.qmeta[QM_NCHUNKS]) - chunk_lo;
  if (nchunks <= 0) return;
  const int logW = P.logW, W = P.W, N = P.N, L = P.levels;
  const int Wm = W - 1;

  if (COMPACT) {
    for (int e = threadIdx.x; e < TPR_LUT_ROWS * 256; e += blockDim.x) {
      const int q = e >> 8, m8 = e & 255;
      s_lut[e] = P.unit_lut[q * TPR_LUT_COLS + ((m8 & 15) | ((m8 >> 4) << 5))];
    }
  } else {
    load_unit_lut(s_lut, P.unit_lut);
  }
  if (S5) {
    uint32_t *dst = reinterpret_cast<uint32_t *>(s_lut + ULUT_BYTES);
    for (int e = threadIdx.x; e < S5_WORDS; e += blockDim.x) dst[e] = P.s5_tab[e];
  }
  __syncthreads();

  // Each warp takes its share of the chunks and retires: the grid is several waves of CTAs, so SM slots keep
  // freeing up for the (higher-priority) transform kernels of other units instead of being held to the end.
  const int share = max(1, (nchunks + (int)gridDim.x * TPR_WARPS - 1) / ((int)gridDim.x * TPR_WARPS));
  for (int taken = 0; taken < share; taken++) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(&P.qmeta[WIDEWIN ? QM_CUR_WIDE : QM_CUR_SMALL], 1);
    chunk = __shfl_sync(FULL_MASK, chunk, 0);
    if (chunk >= nchunks) break;
    chunk += chunk_lo;
    const int qstart = P.chunk_start[chunk], cnt = P.chunk_cnt[chunk];
    if (WIDEWIN && MODE == MODE_EUCLID && cnt == 1) {
      // one long chain: the whole warp walks it together (paths.cuh, find_next_geo)
      const int g = P.queue[qstart];
      if (P.reg.size[g] >= P.coop_min) {
        region_pyramid<MODE>(P, g, arena, s_lut);
        __syncwarp();
        continue;
      }
    }
    const bool mine = lane < cnt;
    int img = 0, label = 0, first = 0, size = 0, off = 0, r0 = 0, c0 = 0, h = 0, w = 0, ws = 0;
    if (mine) {
      const int g = P.queue[qstart + lane];
      img = P.reg.img[g]; label = P.reg.label[g]; first = P.reg.first[g];
      size = P.reg.size[g]; off = P.reg.off[g];
      tpr_geometry(P, g, r0, c0, h, w, ws);
    }
    const bool narrow = __all_sync(FULL_MASK, ws <= 1);     // every bitmap of the chunk has one word per row
    const int slot = mine ? tpr_slot_words(h * ws) : 0;  // the chunk table guarantees the sum fits
    int inc = slot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int y = __shfl_up_sync(FULL_MASK, inc, d);
      if (lane >= d) inc += y;
    }
    const int base = inc - slot;
    uint32_t *bm = arena + base;
    const float inv_ws = 1.0f / (float)max(ws, 1);

    if (!WIDEWIN && chunk - chunk_lo < P.gbm_chunks) {
      // the bitmaps were built by k1_bitmaps (the same layout, in global memory): copy the chunk's words,
      // 16 bytes per lane and load, every load in flight at once
      const int nvec = (__shfl_sync(FULL_MASK, inc, 31) + 3) >> 2;
      const uint4 *src = reinterpret_cast<const uint4 *>(P.gbm + (size_t)(chunk - chunk_lo) * TPR_ARENA_WORDS);
      uint4 *dst = reinterpret_cast<uint4 *>(arena);
      for (int e0 = 0; e0 < nvec; e0 += 8 * 32) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int e = e0 + u * 32 + lane;
          v[u] = e < nvec ? __ldcs(src + e) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int e = e0 + u * 32 + lane;
          if (e < nvec) dst[e] = v[u];
        }
      }
    } else {
      tpr_build_bitmaps(P, arena, cnt, img, label, r0, c0, h, w, ws, base);
    }
    __syncwarp();

    // every lane walks its own region; the warp advances level by level
    bool live = mine, list = false;
    int n = mine ? size : 0, a = off;
    int si = TPR_PAD, sj = (first & Wm) - c0;  // bitmap mode: start point
    const int pixbase = (r0 << logW) + c0;
    int lb = 0, sidx = 0;                // list mode: buffer offset of the level's list, index of its start point
    int32_t *Qimg = P.Q + (size_t)img * 2 * (size_t)N;
    int32_t *Pimg = P.Pm + (size_t)img * 2 * (size_t)N;
    int32_t *posmap = P.posmap + (size_t)img * N;  // pixel -> place in the next level's incoming order
    for (int lev = 1; lev <= L; lev++) {
      int32_t *Ql = Qimg + level_off((size_t)N, lev) + a;
      int32_t *Pl = Pimg + level_off((size_t)N, lev) + a;

      const bool keep = lev < L;  // the next level exists: collect the survivors
      int t = n, ci = si, cj = sj, p0 = 0, p1 = 1;  // prefered_direc = (0,1)   rbepwt.py:1290
      int rad = 1, i = 0, wd = 0, i1 = 0;
      bool fresh = true;     // at the first row of a window
      bool near = true;      // table-step variant: the step starts with the 5x5 window
      bool ring = false;     // ... and its 3x3 neighbourhood is known to be empty (the unit-step table found nothing)
      unsigned U = 0;        // list mode: unvisited mask
      int ncnt = 0;          // list mode: survivors appended to the other buffer
      uint32_t nmin = 0xffffffffu;
      int nminidx = 0;
      Search<MODE> S;
      S.reset();
#define TPR_SET_WINDOW()                                            \
  do {                                                              \
    if (!WIDEWIN) rad = min(rad, 2 * TPR_MAX_RAD);                  \
    i = max(ci - rad, 0); i1 = min(ci + rad, h - 1);                \
    wd = 0; fresh = true;                                           \
  } while (0)
#define TPR_LIST_KEEP(entry)                                        \
  do {                                                              \
    bm[(lb ^ 32) + ncnt] = (entry);                                 \
    if ((entry) < nmin) { nmin = (entry); nminidx = ncnt; }         \
    ncnt++;                                                         \
  } while (0)
      if (live) {
        if (list) {
          const uint32_t e = bm[lb + sidx];
          ci = (int)(e >> 16); cj = (int)(e & 0xffffu);
          U = (n >= 32 ? 0xffffffffu : (1u << n) - 1u) & ~(1u << sidx);
          if (keep && (a & 1) == 0) TPR_LIST_KEEP(e);
        } else {
          bm[si * ws + (sj >> 5)] &= ~(1u << (sj & 31));
          TPR_SET_WINDOW();
        }
        const int pix0 = ((r0 + ci) << logW) + c0 + cj;
        Ql[0] = pix0;
        t = 1;
      }
      while (__any_sync(FULL_MASK, t < n)) {
        if constexpr (S5) {
        // ---- phase 1a, unit steps: 3x3 neighbourhood -> 9-bit mask -> table, up to TPR_UNIT_STEPS per trip (unit pref:
        // the state every dense stretch of a path is in).  The margin keeps the window inside the bitmap.
#pragma unroll 1
        for (int rep = 0; rep < TPR_UNIT_STEPS; rep++) {
          const bool unit = t < n && !list && near && !ring && (unsigned)(p0 + 1) <= 2u && (unsigned)(p1 + 1) <= 2u;
          if (!__any_sync(FULL_MASK, unit)) break;
          if (unit) {
#ifdef TPR_STATS
            atomicAdd(&g_tpr_stats[((min(lev, 16) - 1) * 4 + 0) * 2 + 1], 1ull);
#endif
            const int sc = cj - 1;
            const uint32_t *rp = bm + (ci - 1) * ws + (sc >> 5);
            const int sh = sc & 31;
            unsigned m;
            // three bits straddle a word boundary only when sh > 29 (never with one word per row)
            if (narrow || !__any_sync(__activemask(), sh > 29)) {
              m = ((rp[0] >> sh) & 7u) | (((rp[ws] >> sh) & 7u) << 3) | (((rp[2 * ws] >> sh) & 7u) << 6);
            } else {
              m = (__funnelshift_r(rp[0], rp[1], sh) & 7u) | ((__funnelshift_r(rp[ws], rp[ws + 1], sh) & 7u) << 3) |
                  ((__funnelshift_r(rp[2 * ws], rp[2 * ws + 1], sh) & 7u) << 6);
            }
            if (m) {
              const int idx = s_lut[((p0 + 1) * 3 + (p1 + 1)) * 256 + ((m & 15u) | ((m >> 5) << 4))];
              p0 = (idx * 11) >> 5;  // idx / 3 for idx < 9
              p1 = idx - 3 * p0 - 1;
              p0 -= 1;  // rbepwt.py:1331
              ci += p0; cj += p1;
              bm[ci * ws + (cj >> 5)] &= ~(1u << (cj & 31));
              Ql[t] = pixbase + (ci << logW) + cj;
              t++;
            } else {
              ring = true;  // nothing at distance 1: the 5x5 table step below
            }
          }
        }
        // ---- phase 1b, 5x5 table steps (pref not a unit step, or nothing at distance 1): neighbourhood -> class-grouped
        // code -> winner, up to TPR_S5_STEPS per trip; every lane that takes one runs the same instructions
#pragma unroll 1
        for (int rep = 0; rep < TPR_S5_STEPS; rep++) {
          const bool go = t < n && !list && near && (ring || (unsigned)(p0 + 1) > 2u || (unsigned)(p1 + 1) > 2u);
          if (!__any_sync(FULL_MASK, go)) break;
          if (go) {
#ifdef TPR_STATS
            atomicAdd(&g_tpr_stats[((min(lev, 16) - 1) * 4 + 1) * 2 + 1], 1ull);
#endif
            const int sc = cj - TPR_PAD;  // >= 0: the margin
            const uint32_t *rp = bm + (ci - TPR_PAD) * ws + (sc >> 5);
            const int sh = sc & 31;
            uint32_t code = 0;
            // the five bits straddle a word boundary only when sh > 27 (never with one word per row: cj + 2 < w <= 32)
            if (narrow || !__any_sync(__activemask(), sh > 27)) {
#pragma unroll
              for (int rr = 0; rr < 5; rr++) code |= s5_t1[rr * 32 + ((rp[rr * ws] >> sh) & 31u)];
            } else {
#pragma unroll
              for (int rr = 0; rr < 5; rr++)
                code |= s5_t1[rr * 32 + (__funnelshift_r(rp[rr * ws], rp[rr * ws + 1], sh) & 31u)];
            }
            ring = false;
            if (code) {
              int shift, base, di, dj;
              uint32_t fld;
              s5_field(code, shift, fld, base);
              if ((unsigned)(p0 + 2) <= 4u && (unsigned)(p1 + 2) <= 4u) {
                const int cell = s5_t2[((p0 + 2) * 5 + p1 + 2) * S5_T2_ROW + base + fld];
                di = (cell * 13) >> 6;  // cell / 5 for cell < 25
                dj = cell - 5 * di - 2;
                di -= 2;
              } else {  // pref is a longer vector (the step after a jump): the class's cells through the candidate code
                Search<MODE> T;
                T.reset();
                for (uint32_t u = fld; u; u &= u - 1) {
                  const int q = s5_cell[shift + __ffs(u) - 1];
                  const int qi = (q * 13) >> 6;
                  T.consider(true, qi - 2, q - 5 * qi - 2, p0, p1);
                }
                int fk;
                T.finish(p0, p1, di, dj, fk);
              }
              p0 = di; p1 = dj;  // rbepwt.py:1331
              ci += di; cj += dj;
              bm[ci * ws + (cj >> 5)] &= ~(1u << (cj & 31));
              Ql[t] = pixbase + (ci << logW) + cj;
              t++;
              rad = 4;  // where a miss of the next step starts
            } else {  // nothing within half-width 2: the window search below, from the half-width of the last jump
              near = false;
              rad = max(rad, 4);
              TPR_SET_WINDOW();
            }
          }
        }
        } else {
        // ---- phase 1, unit steps: 3x3 neighbourhood -> 9-bit mask -> table, up to TPR_UNIT_STEPS per trip
        // (half-width 1, unit pref: the state every dense stretch of a path is in)
#pragma unroll 1
        for (int rep = 0; rep < (WIDEWIN ? TPR_UNIT_STEPS_WIDE : TPR_UNIT_STEPS); rep++) {
          const bool unit = t < n && !list && fresh && rad == 1 && (unsigned)(p0 + 1) <= 2u && (unsigned)(p1 + 1) <= 2u;
          if (!__any_sync(FULL_MASK, unit)) break;
          if (unit) {
#ifdef TPR_STATS
            atomicAdd(&g_tpr_stats[((min(lev, 16) - 1) * 4 + 0) * 2 + 1], 1ull);
#endif
            unsigned m = 0;
            if constexpr (COMPACT) {
              // the margin keeps the 3x3 window inside the bitmap; its three bits straddle a word boundary only
              // when sh > 29 (never with one word per row)
              const int sc = cj - 1;
              const uint32_t *rp = bm + (ci - 1) * ws + (sc >> 5);
              const int sh = sc & 31;
              if (narrow || !__any_sync(__activemask(), sh > 29)) {
                m = ((rp[0] >> sh) & 7u) | (((rp[ws] >> sh) & 7u) << 3) | (((rp[2 * ws] >> sh) & 7u) << 6);
              } else {
                m = (__funnelshift_r(rp[0], rp[1], sh) & 7u) | ((__funnelshift_r(rp[ws], rp[ws + 1], sh) & 7u) << 3) |
                    ((__funnelshift_r(rp[2 * ws], rp[2 * ws + 1], sh) & 7u) << 6);
              }
            } else {
#pragma unroll
              for (int rr = 0; rr < 3; rr++) {
                const int ri = ci + rr - 1;
                const uint32_t x = (ri >= 0 && ri < h) ? row_window(bm + ri * ws, ws, cj) : 0u;
                m |= ((x >> 14) & 7u) << (3 * rr);
              }
            }
            if (m) {
              const int idx = COMPACT ? s_lut[((p0 + 1) * 3 + (p1 + 1)) * 256 + ((m & 15u) | ((m >> 5) << 4))]
                                      : s_lut[((p0 + 1) * 3 + (p1 + 1)) * TPR_LUT_COLS + m];
              p0 = idx / 3 - 1; p1 = idx % 3 - 1;  // rbepwt.py:1331
              ci += p0; cj += p1;
              bm[ci * ws + (cj >> 5)] &= ~(1u << (cj & 31));
              Ql[t] = ((r0 + ci) << logW) + c0 + cj;
              t++;  // still half-width 1, still at the start of a window
            } else {
              rad = 2;
              TPR_SET_WINDOW();
            }
          }
        }
        }
        // ---- phase 2: one unit of the other kinds
        const bool unit_now = S5 ? (!list && near)
                                 : (!list && fresh && rad == 1 && (unsigned)(p0 + 1) <= 2u && (unsigned)(p1 + 1) <= 2u);
        bool want2 = t < n && !unit_now;
        if (S5 && TPR_P2_MIN > 1) {
          // the other kinds wait until TPR_P2_MIN lanes want them (or no lane has a table step left): their code is long
          // and divergent, so it is run for several lanes at once rather than in every trip for one or two
          const int waiting = __popc(__ballot_sync(FULL_MASK, want2));
          const bool table_left = __any_sync(FULL_MASK, t < n && unit_now);
          want2 = want2 && (waiting >= TPR_P2_MIN || !table_left);
        }
        if (want2) {
          bool commit = false, expand = false;
          int fdi = 0, fdj = 0, fk = 0;
#ifdef TPR_STATS
          {
            const int kind = list ? 3 : (rad <= TPR_MAX_RAD ? 1 : 2);
            atomicAdd(&g_tpr_stats[((min(lev, 16) - 1) * 4 + kind) * 2 + 1], 1ull);
            if (lane == __ffs(__activemask()) - 1) atomicAdd(&g_tpr_stats[128 + min(lev, 16) - 1], 1ull);
          }
#endif
          if (list) {
            // ---- list mode: one whole step, candidates = the points still unvisited
            for (unsigned u = U; u; u &= u - 1) {
              const int idx = __ffs(u) - 1;
              const uint32_t e = bm[lb + idx];
              S.consider(true, (int)(e >> 16) - ci, (int)(e & 0xffffu) - cj, p0, p1, idx);
            }
            S.finish(p0, p1, fdi, fdj, fk);
            const int idx = S.tag;
            U &= ~(1u << idx);
            if (keep && ((a + t) & 1) == 0) TPR_LIST_KEEP(bm[lb + idx]);
            ci += fdi; cj += fdj;
            const int pix = ((r0 + ci) << logW) + c0 + cj;
            Ql[t] = pix;
            p0 = fdi; p1 = fdj;
            t++;
            S.reset();
          } else {
            if (rad <= TPR_MAX_RAD) {
              // ---- up to TPR_ROWS_PER_TRIP window rows, each one aligned word
              fresh = false;
              const uint32_t wmask = ((2u << (2 * rad)) - 1u) << (15 - rad);
              if (WIDEWIN) {  // few long chains per warp: the whole window in one trip
#pragma unroll 1
                for (; i <= i1; i++) {
                  const uint32_t x = row_window(bm + i * ws, ws, cj) & wmask;
                  if (x) S.scan_row(x, i - ci, p0, p1);
                }
              } else {
#pragma unroll 1
                for (int u = 0; u < TPR_ROWS_PER_TRIP && i <= i1; u++, i++) {  // only the rows the window has
                  const uint32_t x = row_window(bm + i * ws, ws, cj) & wmask;
                  if (x) S.scan_row(x, i - ci, p0, p1);
                }
              }
              if (i > i1) {
                if (S.have()) { S.finish(p0, p1, fdi, fdj, fk); commit = true; }
                else expand = true;
              }
            } else {
              fresh = false;
              if (WIDEWIN) {
                // ---- nothing within TPR_MAX_RAD: windows of half-width 16, 32, ... clipped to the box, rows
                // i..i1, words w0..w1 of each, four words per trip.  Ranking by (k, d2, dot) lets each window
                // start from scratch.
                const int j0 = max(cj - rad, 0), j1 = min(cj + rad, w - 1);
                const int w0 = j0 >> 5, w1 = j1 >> 5;
                if (wd < w0) wd = w0;  // first trip of the window
#pragma unroll
                for (int u = 0; u < 4; u++) {
                  if (i <= i1) {
                    uint32_t bits = bm[i * ws + wd];
                    const int lo = wd << 5;
                    if (lo < j0) bits &= 0xffffffffu << (j0 - lo);
                    if (lo + 31 > j1) bits &= 0xffffffffu >> (lo + 31 - j1);
                    if (bits) S.scan_word(bits, lo, i - ci, cj, p0, p1);
                    if (wd < w1) wd++;
                    else { wd = w0; i++; }
                  }
                }
                if (i > i1) {
                  if (S.have()) { S.finish(p0, p1, fdi, fdj, fk); commit = true; }
                  else expand = true;
                }
              } else {
                // ---- nothing within TPR_MAX_RAD: scan the region's whole (small) bitmap, four words per trip.
                // Ranking by (k, d2, dot) makes this equal to the remaining probes 2*TPR_MAX_RAD, ... in turn.
                const int nwords = h * ws;
                uint32_t b4[TPR_SCAN_WORDS];
#pragma unroll
                for (int u = 0; u < TPR_SCAN_WORDS; u++) b4[u] = wd + u < nwords ? bm[wd + u] : 0u;
#pragma unroll
                for (int u = 0; u < TPR_SCAN_WORDS; u++)
                  if (b4[u]) {
                    const int wi = wd + u;
                    const int ri = ws == 1 ? wi : (int)(((float)wi + 0.5f) * inv_ws);  // wi / ws (wi < 2^11: exact)
                    S.scan_word(b4[u], (wi - ri * ws) << 5, ri - ci, cj, p0, p1);
                  }
                wd += TPR_SCAN_WORDS;
                if (wd >= nwords) {
                  if (S.have()) { S.finish(p0, p1, fdi, fdj, fk); commit = true; }
                  else expand = true;  // the box is covered: reported as corrupt state below
                }
              }
            }
            if (commit || expand) {
              if (commit) {
                const int bi = ci + fdi, bj = cj + fdj;
                bm[bi * ws + (bj >> 5)] &= ~(1u << (bj & 31));
                const int pix = ((r0 + bi) << logW) + c0 + bj;
                Ql[t] = pix;
                p0 = fdi; p1 = fdj;  // rbepwt.py:1331
                ci = bi; cj = bj;
                t++;
                rad = 1 << fk;
                near = true;
                S.reset();
              } else if (rad > TPR_MAX_RAD &&
                         (!WIDEWIN || (ci - rad <= 0 && cj - rad <= 0 && ci + rad >= h - 1 && cj + rad >= w - 1))) {
                atomicExch(&P.qmeta[QM_ERR], 1);  // nothing unvisited in the whole box: corrupt state
                t = n; live = false;
              } else {
                rad <<= 1;
              }
              TPR_SET_WINDOW();
            }
          }
        }
      }
      // ---- end of the level, one pass over the level's path Ql[0, n) (16 points per trip: four 16-byte loads, then
      // their sixteen posmap reads, all in flight together -- this part of the kernel is pure memory latency):
      //  * levels >= 2: Pl[t] = place of the t-th path point in this level's incoming order (posmap, filled by the
      //    previous level's pass);
      //  * RegionCollection.reduce: the points at even GLOBAL position a+t survive (rbepwt.py:1563-1584); they get
      //    their place in the next level's incoming order (posmap) and are re-marked in the (all-zero) bitmap, or
      //    become the list when at most TPR_LIST_MAX are left; the next start point is the lexicographically
      //    smallest survivor (rbepwt.py:1035-1036).  (List mode collected its survivors during the walk.)
      const int na = (a + 1) >> 1, nb = (a + n + 1) >> 1;
      const int nnext = lev < L ? nb - na : 0;
      {
        const bool doP = lev >= 2;
        const int smode = (!live || nnext <= 0 || list) ? 0 : (nnext <= TPR_LIST_MAX ? 2 : 1);  // 1 re-mark, 2 -> list
        int minpix = INT32_MAX, lk = 0;
        uint32_t mn = 0xffffffffu;
        auto survivor = [&](int tt, int pix) {
          posmap[pix] = (a + tt) >> 1;
          const int pi = (pix >> logW) - r0, pj = (pix & Wm) - c0;
          if (smode == 1) {
            bm[pi * ws + (pj >> 5)] |= 1u << (pj & 31);
            minpix = min(minpix, pix);
          } else {
            const uint32_t e = (uint32_t)(pi << 16) | (uint32_t)pj;
            bm[lk] = e;
            if (e < mn) { mn = e; sidx = lk; }
            lk++;
          }
        };
        const int nn = live ? n : 0;
        // points per trip: 16 in the bulk instantiation; 8 in the windowed one, whose walk (with the warp-per-region
        // code inlined) has no registers to spare -- 16 there spills in the walk (40 bytes of stack instead of 16)
        constexpr int TB = WIDEWIN ? 8 : 16;
        // a ragged block (the head up to the 16-byte boundary, the tail): scalar loads, still all in flight together
        auto ragged = [&](int t0, int t1) {
          int q[TB], v[TB];
#pragma unroll
          for (int u = 0; u < TB; u++) q[u] = t0 + u < t1 ? __ldcg(Ql + t0 + u) : 0;
          if (doP) {
#pragma unroll
            for (int u = 0; u < TB; u++) v[u] = t0 + u < t1 ? __ldcg(posmap + q[u]) : 0;
#pragma unroll
            for (int u = 0; u < TB; u++)
              if (t0 + u < t1) Pl[t0 + u] = v[u];
          }
          if (smode) {
#pragma unroll
            for (int u = 0; u < TB; u++)
              if (t0 + u < t1 && ((a + t0 + u) & 1) == 0) survivor(t0 + u, q[u]);
          }
        };
        int tt = min(nn, (int)((4u - (unsigned)(((uintptr_t)Ql >> 2) & 3u)) & 3u));  // up to the 16-byte boundary
        if (tt > 0) ragged(0, tt);
        for (; tt + TB <= nn; tt += TB) {
          int q[TB], v[TB];
#pragma unroll
          for (int u = 0; u < TB / 4; u++) {
            const int4 x = __ldcg(reinterpret_cast<const int4 *>(Ql + tt) + u);
            q[4 * u] = x.x; q[4 * u + 1] = x.y; q[4 * u + 2] = x.z; q[4 * u + 3] = x.w;
          }
          if (doP) {
#pragma unroll
            for (int u = 0; u < TB; u++) v[u] = __ldcg(posmap + q[u]);
#pragma unroll
            for (int u = 0; u < TB / 4; u++)
              reinterpret_cast<int4 *>(Pl + tt)[u] = make_int4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
          }
          if (smode) {
            const int par = (a + tt) & 1;  // survivors: tt + par, tt + par + 2, ...
#pragma unroll
            for (int u = 0; u < TB / 2; u++) survivor(tt + par + 2 * u, par ? q[2 * u + 1] : q[2 * u]);
          }
        }
        if (tt < nn) ragged(tt, nn);  // the tail: fewer than TB points
        if (smode == 1) { si = (minpix >> logW) - r0; sj = (minpix & Wm) - c0; }
        if (smode == 2) { list = true; lb = 0; }
      }
      if (lev == L) break;
      if (live && nnext > 0 && list && ncnt > 0) {
        // list mode: survivors were appended to the other buffer during the walk, in order: places na, na+1, ...
        lb ^= 32; sidx = nminidx;
        for (int k = 0; k < ncnt; k++) {
          const uint32_t e = bm[lb + k];
          posmap[((r0 + (int)(e >> 16)) << logW) + c0 + (int)(e & 0xffffu)] = na + k;
        }
      }
      a = na; n = nnext;
      live = live && n > 0;
      if (!live) n = 0;
      if (!__any_sync(FULL_MASK, live)) break;
    }
#undef TPR_SET_WINDOW
#undef TPR_LIST_KEEP
    __syncwarp();
  }
}

// Big regions: one warp per CTA, bitmap in dynamic shared memory if it fits, else global scratch.
template <int MODE>
__global__ void __launch_bounds__(32) k1_paths_big(PathParams P) {
  extern __shared__ uint32_t s_big[];
  __shared__ __align__(16) uint8_t s_lut[TPR_LUT_ROWS * TPR_LUT_COLS];
  const int lane = (int)lane_id();
  if (P.qmeta[QM_NBIG] == 0) return;  // the common case: nothing oversized in this group
  load_unit_lut(s_lut, P.unit_lut);
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
    region_pyramid<MODE>(P, g, words <= P.big_smem_words ? s_big : gs, s_lut);
    __syncwarp();
  }
}


}  // namespace rbepwt
