// K1, whole-warp form with the bitmap window in REGISTERS (Euclidean easy path): the walker of the long chains --
// regions too large for a lane's arena slot (k1_paths_big), and every region of a small group (single images,
// small batches: latency is all that matters there).
//
// A path step is a dependent chain, and one warp alone on a scheduler issues one instruction every ~4-6 cycles, so
// the time of a chain is the NUMBER OF DEPENDENT INSTRUCTIONS per step (region_pyramid / find_next_geo, paths.cuh:
// ~100, of which three shared-memory round trips).  Here lane r of the warp keeps 32 bits of bitmap row R0 + r, the
// columns [C0, C0 + 32): a 32 x 32 window around the current point.  As long as the point stays in the window's
// central 16 x 16 cells,
//   * the 5x5 neighbourhood (probes 1 and 2) is five shuffles of a 5-bit slice, resolved by the same class order and
//     table (T2) as the thread-per-region walker's step (walk.cuh);
//   * the probes of half-width 4 and 8 are one pass over the lanes' registers (a row's nearest unvisited column on
//     either side, ranked by (probe index, d2, dot product): wk_row_candidate) and two warp reductions;
//   * visiting a point clears one bit of one lane's register; the bitmap in shared memory is kept in step by a
//     fire-and-forget atomic of one lane (nothing waits for it), so that the window can be re-read at any time:
//     when the point leaves the central cells, and by the search beyond half-width 8 (find_next_geo, from 16 on).
// Step rule = Region.easy_path (/root/reference/rbepwt.py:1273-1347); the candidate code is shared with walk.cuh.
#pragma once
#include "walk.cuh"

namespace rbepwt {

struct RegWin {
  uint32_t W;  // lane r: unvisited bits of bitmap row R0 + r, columns C0 .. C0 + 31 (cells outside the bitmap read 0)
  int R0, C0;
};

__device__ __forceinline__ void rw_load(RegWin &rw, const uint32_t *bm, int h, int ws, int ci, int cj) {
  rw.R0 = ci - 16; rw.C0 = cj - 16;
  const int i = rw.R0 + (int)lane_id();
  uint32_t x = 0u;
  if ((unsigned)i < (unsigned)h) {
    const int wq = rw.C0 >> 5, sh = rw.C0 & 31;  // arithmetic shift: floor, C0 may be negative
    const uint32_t lo = (unsigned)wq < (unsigned)ws ? bm[i * ws + wq] : 0u;
    const uint32_t hi = (unsigned)(wq + 1) < (unsigned)ws ? bm[i * ws + wq + 1] : 0u;
    x = __funnelshift_r(lo, hi, sh);
  }
  rw.W = x;
}

// The window's bits back into the bitmap (bits are only ever cleared, so this is an AND; lanes own distinct rows).
__device__ __forceinline__ void rw_store(const RegWin &rw, uint32_t *bm, int h, int ws) {
  const int i = rw.R0 + (int)lane_id();
  if ((unsigned)i < (unsigned)h) {
    const int wq = rw.C0 >> 5, sh = rw.C0 & 31;
    const uint32_t keep_lo = (1u << sh) - 1u;  // columns of word wq to the left of the window
    if ((unsigned)wq < (unsigned)ws) bm[i * ws + wq] &= (rw.W << sh) | keep_lo;
    if (sh && (unsigned)(wq + 1) < (unsigned)ws) bm[i * ws + wq + 1] &= (rw.W >> (32 - sh)) | ~keep_lo;
  }
}

// The step from a non-empty 5x5 window (euclid): the nearest class present, its winner for pref (p0, p1).
// `pc` = the preferred direction as a cell of the 5x5 window ((p0 + 2) * 5 + p1 + 2), or -1 when it is a jump; returns the
// chosen cell (which is the next step's pc).
__device__ __forceinline__ int rw_near_pick(uint32_t n25, int pc, int p0, int p1, const uint8_t *t2) {
  if (pc >= 0) {
    // every lane holds the same n25: the class tests are uniform branches, the hash constants immediates
    int slot;
    if (n25 & N25_A) slot = wk_t2_base(0) + wk_t2_field(0, n25);
    else if (n25 & N25_B) slot = wk_t2_base(1) + wk_t2_field(1, n25);
    else if (n25 & N25_C) slot = wk_t2_base(2) + wk_t2_field(2, n25);
    else if (n25 & N25_D) slot = wk_t2_base(3) + wk_t2_field(3, n25);
    else slot = wk_t2_base(4) + wk_t2_field(4, n25);
    return t2[pc * T2_ROW + slot];
  }
  // pref is a jump: the class's cells through the integer dot product, a mirror pair by the reference's rule
  uint32_t cls = n25 & N25_A;
  int d2 = 1;
  if (!cls) { cls = n25 & N25_B; d2 = 2; }
  if (!cls) { cls = n25 & N25_C; d2 = 4; }
  if (!cls) { cls = n25 & N25_D; d2 = 5; }
  if (!cls) { cls = n25 & N25_E; d2 = 8; }
  int bdot = INT32_MIN, di = 0, dj = 0, adi = 0, adj = 0;
  bool alt = false;
  for (uint32_t u = cls; u; u &= u - 1u) {
    const int b = __ffs(u) - 1;
    const int qi = ((b * 13) >> 6) - 2, qj = b - 5 * (qi + 2) - 2;
    const int dot = qi * p0 + qj * p1;
    if (dot > bdot) { bdot = dot; di = qi; dj = qj; alt = false; }
    else if (dot == bdot) { alt = true; adi = qi; adj = qj; }
  }
  if (alt && mirror_second_wins(di, dj, adi, adj, d2, p0, p1)) { di = adi; dj = adj; }
  return (di + 2) * 5 + dj + 2;
}

// Probes of half-width 4 and 8 from the register window ((wi, wj) = the point's window coordinates, 8 <= wi, wj <= 23):
// one pass, candidates ranked by (k, d2) and the dot product.  false: nothing within half-width 8.
__device__ __forceinline__ bool rw_probe8(const RegWin &rw, int wi, int wj, int p0, int p1, int &di, int &dj) {
  const int lane = (int)lane_id();
  const int rdi = lane - wi;
  const uint32_t x = abs(rdi) <= 8 ? (rw.W & (0x1ffffu << (wj - 8))) : 0u;
  unsigned bkey = 0xffffffffu;
  int bdot = 0, boff = 0, aoff = WK_NO_PARTNER;
  if (x) wk_row_candidate<uint32_t>(x, wj, rdi, p0, p1, bkey, bdot, boff, aoff);
  const unsigned kmin = __reduce_min_sync(FULL_MASK, bkey);
  if (kmin == 0xffffffffu) return false;
  const bool sel = bkey == kmin;
  const int dotmax = __reduce_max_sync(FULL_MASK, sel ? bdot : INT32_MIN);
  const unsigned tied = __ballot_sync(FULL_MASK, sel && bdot == dotmax);
  const int la = __ffs(tied) - 1;
  int step = __shfl_sync(FULL_MASK, boff, la);
  int other = __shfl_sync(FULL_MASK, aoff, la);
  const unsigned rest = tied & (tied - 1);
  if (rest) other = __shfl_sync(FULL_MASK, boff, __ffs(rest) - 1);  // a mirror pair held by two lanes
  if (other != WK_NO_PARTNER &&
      mirror_second_wins(step >> 16, (int)(short)(step & 0xffff), other >> 16, (int)(short)(other & 0xffff),
                         (int)(kmin & 0x1fffffu), p0, p1))
    step = other;
  di = step >> 16; dj = (int)(short)(step & 0xffff);
  return true;
}

// One level's path of one region.  (ci, cj) = start point (bitmap coordinates, bit still set); Ql[t], t = 0..n-1,
// receives the pixel ids in path order.  The bitmap is all-zero afterwards.  Loop-carried state: the point's window
// coordinates (wi, wj), its pixel id, the preferred direction; lane t mod 32 stores path point t.
__device__ __forceinline__ bool rw_run_path(uint32_t *bm, int h, int w, int ws, int ci, int cj, int n, int r0, int c0, int logW,
                                            int32_t *__restrict__ Ql, const uint8_t *t2) {
  const int lane = (int)lane_id();
  int pix = ((r0 + ci) << logW) + c0 + cj;
  if (lane == 0) {
    Ql[0] = pix;
    bm[ci * ws + (cj >> 5)] &= ~(1u << (cj & 31));
  }
  __syncwarp();
  RegWin rw;
  rw_load(rw, bm, h, ws, ci, cj);
  int wi = 16, wj = 16;
  int p0 = 0, p1 = 1, pc = 2 * 5 + 3;  // prefered_direc = (0,1)   rbepwt.py:1290
  for (int t = 1; t < n; t++) {
    if ((unsigned)(wi - 8) > 15u || (unsigned)(wj - 8) > 15u) {  // left the central cells: a fresh window
      rw_store(rw, bm, h, ws);
      __syncwarp();
      rw_load(rw, bm, h, ws, rw.R0 + wi, rw.C0 + wj);
      wi = wj = 16;
    }
    const uint32_t s5 = (rw.W >> (wj - 2)) & 31u;
    const uint32_t n25 = __shfl_sync(FULL_MASK, s5, wi - 2) | (__shfl_sync(FULL_MASK, s5, wi - 1) << 5) |
                         (__shfl_sync(FULL_MASK, s5, wi) << 10) | (__shfl_sync(FULL_MASK, s5, wi + 1) << 15) |
                         (__shfl_sync(FULL_MASK, s5, wi + 2) << 20);
    int di, dj;
    if (n25) {
      pc = rw_near_pick(n25, pc, p0, p1, t2);
      di = (pc * 13) >> 6;  // pc / 5 for pc < 25
      dj = pc - 5 * di - 2;
      di -= 2;
      wi += di; wj += dj;
      rw.W &= ~((lane == wi ? 1u : 0u) << wj);
    } else {
      if (!rw_probe8(rw, wi, wj, p0, p1, di, dj)) {
        const int ci_ = rw.R0 + wi, cj_ = rw.C0 + wj;
        int rad = 16, bi, bj;
        rw_store(rw, bm, h, ws);
        __syncwarp();
        if (!find_next_geo(bm, h, w, ws, ci_, cj_, p0, p1, rad, nullptr, bi, bj)) return false;
        di = bi - ci_; dj = bj - cj_;
      }
      pc = -1;
      wi += di; wj += dj;
      if ((unsigned)wi < 32u && (unsigned)wj < 32u) {
        rw.W &= ~((lane == wi ? 1u : 0u) << wj);
      } else {  // a jump out of the window: the bit in the bitmap itself, then a fresh window (the check above)
        const int ni = rw.R0 + wi, nj = rw.C0 + wj;
        if (lane == 0) bm[ni * ws + (nj >> 5)] &= ~(1u << (nj & 31));
        __syncwarp();
      }
    }
    pix += (di << logW) + dj;
    if (lane == (t & 31)) Ql[t] = pix;
    p0 = di; p1 = dj;  // rbepwt.py:1331
  }
  rw_store(rw, bm, h, ws);
  __syncwarp();
  return true;
}

// EPWT (rbepwt.py:1296-1306 with the value distance): the same window, the candidates of probe 1 (the 3x3 ring) or, when
// that is empty, probe 2 (the rest of the 5x5 window) one per lane -- their values are loaded together (L1: the walk is
// local) -- and the warp's arg-best of paths.cuh; beyond half-width 2 the search of paths.cuh from half-width 4 on.
// The whole image is one region (r0 = c0 = 0).  Pl[t] = posmap[Ql[t]] (levels >= 2), as run_path does.
__device__ bool rw_run_path_epwt(uint32_t *bm, int h, int w, int ws, int ci, int cj, int n, int logW,
                                 const double *__restrict__ vals, bool u8wrap, int32_t *__restrict__ Ql,
                                 int32_t *__restrict__ Pl, const int32_t *posmap) {
  const int lane = (int)lane_id();
  int pix = (ci << logW) + cj;
  int myq = pix;  // lane t mod 32 holds path point t until the warp flushes 32 of them
  if (lane == 0) bm[ci * ws + (cj >> 5)] &= ~(1u << (cj & 31));
  __syncwarp();
  RegWin rw;
  rw_load(rw, bm, h, ws, ci, cj);
  int wi = 16, wj = 16;
  int p0 = 0, p1 = 1;
  double curval = RB_EPWT_LOAD(vals + pix);
  const int ldi = ((lane * 13) >> 6) - 2, ldj = lane - 5 * (ldi + 2) - 2;  // this lane's cell of the 5x5 window (lane < 25)
  const int lpix = (ldi << logW) + ldj;
  for (int t = 1; t < n; t++) {
    if ((unsigned)(wi - 2) > 27u || (unsigned)(wj - 2) > 27u) {
      rw_store(rw, bm, h, ws);
      __syncwarp();
      rw_load(rw, bm, h, ws, rw.R0 + wi, rw.C0 + wj);
      wi = wj = 16;
    }
    const uint32_t s5 = (rw.W >> (wj - 2)) & 31u;
    const uint32_t n25 = __shfl_sync(FULL_MASK, s5, wi - 2) | (__shfl_sync(FULL_MASK, s5, wi - 1) << 5) |
                         (__shfl_sync(FULL_MASK, s5, wi) << 10) | (__shfl_sync(FULL_MASK, s5, wi + 1) << 15) |
                         (__shfl_sync(FULL_MASK, s5, wi + 2) << 20);
    const int ci_ = rw.R0 + wi, cj_ = rw.C0 + wj;
    int bi, bj;
    if (n25) {
      const uint32_t cm = (n25 & N25_RING1) ? (n25 & N25_RING1) : n25;
      const bool has = (cm >> lane) & 1u;  // this lane's cell is a candidate (cm < 2^25: lanes 25..31 never are)
      // the value distance as an orderable 64-bit key (dist >= 0), smallest over the warp; one winner is the common case
      double val = 0.0;
      unsigned hi = 0xffffffffu, lo = 0xffffffffu;
      if (has) {
        val = RB_EPWT_LOAD(vals + pix + lpix);
        const double dv = curval - val;
        const double dist = u8wrap ? (dv < 0.0 ? dv + 256.0 : dv) : fabs(dv);  // rbepwt.py:1302 (uint8 arithmetic wraps)
        const unsigned long long k = (unsigned long long)__double_as_longlong(dist);
        hi = (unsigned)(k >> 32); lo = (unsigned)k;
      }
      const unsigned mh = __reduce_min_sync(FULL_MASK, hi);
      const unsigned ml = __reduce_min_sync(FULL_MASK, hi == mh ? lo : 0xffffffffu);
      const unsigned tied = __ballot_sync(FULL_MASK, has && hi == mh && lo == ml);
      if ((tied & (tied - 1u)) == 0u) {
        const int cell = __ffs(tied) - 1;
        const int wdi = ((cell * 13) >> 6) - 2;
        bi = ci_ + wdi; bj = cj_ + cell - 5 * (wdi + 2) - 2;
        curval = __shfl_sync(FULL_MASK, val, cell);
      } else {  // equal value distances: the reference's direction tie-break (paths.cuh)
        Best b;
        b.have = false; b.has_sp1 = false; b.dist = 0.0; b.val = 0.0; b.sp1 = 0.0; b.cross = 0; b.d2 = 0; b.i = 0; b.j = 0;
        if (has) consider<MODE_EPWT>(b, ci_ + ldi, cj_ + ldj, ci_, cj_, p0, p1, curval, vals, pix + lpix, u8wrap);
        warp_arg_best<MODE_EPWT>(b, ci_, cj_, p0, p1, curval, bi, bj);
      }
    } else {
      rw_store(rw, bm, h, ws);
      __syncwarp();
      if (!find_next<MODE_EPWT>(bm, h, w, ws, ci_, cj_, p0, p1, vals, 0, 0, logW, u8wrap, curval, bi, bj, 4)) return false;
    }
    const int di = bi - ci_, dj = bj - cj_;
    wi += di; wj += dj;
    if ((unsigned)wi < 32u && (unsigned)wj < 32u) {
      rw.W &= ~((lane == wi ? 1u : 0u) << wj);
    } else {
      if (lane == 0) bm[bi * ws + (bj >> 5)] &= ~(1u << (bj & 31));
      __syncwarp();
    }
    pix = (bi << logW) + bj;
    if ((t & 31) == lane) myq = pix;
    if ((t & 31) == 31) {  // coalesced flush of 32 path points (+ their positions in the incoming order: a load the
      Ql[t - 31 + lane] = myq;  // store waits for, so not once per step)
      if (Pl) Pl[t - 31 + lane] = __ldcg(posmap + myq);
    }
    p0 = di; p1 = dj;
  }
  if (lane < (n & 31)) {
    Ql[(n & ~31) + lane] = myq;
    if (Pl) Pl[(n & ~31) + lane] = __ldcg(posmap + myq);
  }
  rw_store(rw, bm, h, ws);
  __syncwarp();
  return true;
}

// The whole pyramid of one region (paths only: the positions in the incoming order are k2_perm's).
__device__ void region_pyramid_rw(const PathParams &P, int g, uint32_t *bm, const uint8_t *t2) {
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
  int si = 0, sj = (first & (W - 1)) - c0;
  for (int lev = 1; lev <= P.levels && n > 0; lev++) {
    int32_t *Ql = Q + level_off((size_t)N, lev) + a;
    if (!rw_run_path(bm, h, w, ws, si, sj, n, r0, c0, logW, Ql, t2)) {
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

}  // namespace rbepwt
