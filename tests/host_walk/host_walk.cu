// TEST INFRASTRUCTURE: runs the per-lane logic of the CUDA path walker (struct Walker, rbepwt_b200/csrc/walk.cuh --
// __host__ __device__ code, the very source the kernel k1_walk compiles) on the CPU, one region after the other, so
// that whole path pyramids can be compared with the oracle without a GPU.  Only tests/ builds and loads this.
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../../rbepwt_b200/csrc/walk.cuh"

using namespace rbepwt;

namespace {

// A second, row-ordered restatement of the search beyond the 5x5 window (the kernel does it with the whole warp, lanes =
// rows, same candidate rule: wk_far_search), done by one thread for its own region: `plane` = h rows of ws words, (ci, cj) the
// current point (its own bit is clear).  Returns false if the plane holds no unvisited point.
//   euclid: rows outward from ci; a row's candidate is its nearest unvisited column on either side (both when
//   equidistant), ranked by the packed key (k << 21 | d2) and then the dot product -- the ranking that makes one scan
//   equal to the reference's probes 4, 8, ... in turn.  Rows at distance r only hold candidates with
//   k >= ceil(log2 r) and d2 >= r^2, which ends the scan as soon as those cannot beat the incumbent.
//   chebyshev: the smallest Chebyshev distance c of an unvisited point first (rows outward, same early end), then every
//   point of the ring c through the reference's fp64 tie-break.
__host__ __device__ __forceinline__ void wk_row_nearest(const uint32_t *row, int ws, int cj, int &dl, int &dr) {
  const int wj = cj >> 5;
  const uint32_t below = (1u << (cj & 31)) - 1u;  // columns < cj inside word wj
  dl = dr = -1;
  uint32_t x = row[wj] & below;
  int w = wj;
  while (!x && w > 0) x = row[--w];
  if (x) dl = cj - ((w << 5) + 31 - rb_clz(x));
  x = row[wj] & ~below;
  w = wj;
  while (!x && w < ws - 1) x = row[++w];
  if (x) dr = (w << 5) + rb_ffs(x) - 1 - cj;
}

template <int MODE>
__host__ __device__ __forceinline__ bool wk_far_lane(const uint32_t *plane, int h, int ws, int ci, int cj, int p0, int p1,
                                                     int &di, int &dj) {
  const int maxr = max(ci, h - 1 - ci);
  if (MODE == MODE_EUCLID) {
    unsigned bkey = 0xffffffffu;
    int bdot = 0, bdi = 0, bdj = 0, adi = 0, adj = 0;
    bool alt = false;
    for (int r = 0; r <= maxr; r++) {
      if (bkey != 0xffffffffu) {
        const unsigned kr = (unsigned)probe_index(max(r, 1)), bk = bkey >> 21;
        if (kr > bk || (kr == bk && (unsigned)(r * r) > (bkey & 0x1fffffu))) break;
      }
      for (int s = r ? 0 : 1; s < 2; s++) {
        const int i = s ? ci + r : ci - r;
        if ((unsigned)i >= (unsigned)h) continue;
        int dl, dr;
        wk_row_nearest(plane + i * ws, ws, cj, dl, dr);
        if ((dl & dr) < 0) continue;  // both -1: the row is empty
        const bool use_left = dl >= 0 && (dr < 0 || dl <= dr);
        const int rdi = i - ci;
        int rdj = use_left ? -dl : dr;
        const unsigned key = ((unsigned)probe_index(max(abs(rdi), abs(rdj))) << 21) | (unsigned)(rdi * rdi + rdj * rdj);
        int dot = rdi * p0 + rdj * p1;
        bool ralt = false;
        if (dl == dr) {  // (rdi, -dl) and (rdi, +dl): same key
          const int dot2 = rdi * p0 + dr * p1;
          ralt = dot2 == dot;
          if (dot2 > dot) { dot = dot2; rdj = dr; }
        }
        if (key < bkey || (key == bkey && dot > bdot)) {
          bkey = key; bdot = dot; bdi = rdi; bdj = rdj; alt = ralt; adi = rdi; adj = -rdj;
        } else if (key == bkey && dot == bdot) {
          alt = true; adi = rdi; adj = rdj;
        }
      }
    }
    if (bkey == 0xffffffffu) return false;
    di = bdi; dj = bdj;
    if (alt && mirror_second_wins(bdi, bdj, adi, adj, (int)(bkey & 0x1fffffu), p0, p1)) { di = adi; dj = adj; }
    return true;
  }
  int cmin = INT32_MAX;
  for (int r = 0; r <= maxr && r < cmin; r++)
    for (int s = r ? 0 : 1; s < 2; s++) {
      const int i = s ? ci + r : ci - r;
      if ((unsigned)i >= (unsigned)h) continue;
      int dl, dr;
      wk_row_nearest(plane + i * ws, ws, cj, dl, dr);
      if ((dl & dr) < 0) continue;
      const int d = dl < 0 ? dr : (dr < 0 ? dl : min(dl, dr));
      cmin = min(cmin, max(r, d));
    }
  if (cmin == INT32_MAX) return false;
  Search<MODE> S;
  S.reset();
  const int wcols = ws << 5;
  for (int rdi = -cmin; rdi <= cmin; rdi++) {
    const int i = ci + rdi;
    if ((unsigned)i >= (unsigned)h) continue;
    const uint32_t *row = plane + i * ws;
    const int step = abs(rdi) == cmin ? 1 : 2 * cmin;  // the ring's top / bottom row: every column; else its two ends
    for (int rdj = -cmin; rdj <= cmin; rdj += step) {
      const int j = cj + rdj;
      if ((unsigned)j < (unsigned)wcols && ((row[j >> 5] >> (j & 31)) & 1u)) S.consider(true, rdi, rdj, p0, p1);
    }
  }
  int k;
  S.finish(p0, p1, di, dj, k);
  return true;
}


// The search beyond the 5x5 window, which the kernel does with the whole warp (wk_far_search): here the reference's rule
// taken literally -- probes of half-width 4, 8, ... until one holds an unvisited point, every point of that window
// through the candidate code.
template <int MODE>
bool host_far(Walker<MODE> &wk) {
  const uint32_t *plane = wk.bm + wk.cur;
  const int w = wk.ws * 32;
  for (int rad = 4;; rad <<= 1) {
    const int i0 = std::max(wk.ci - rad, 0), i1 = std::min(wk.ci + rad, wk.h - 1);
    const int j0 = std::max(wk.cj - rad, 0), j1 = std::min(wk.cj + rad, w - 1);
    Search<MODE> S;
    S.reset();
    for (int i = i0; i <= i1; i++)
      for (int j = j0; j <= j1; j++)
        if ((plane[i * wk.ws + (j >> 5)] >> (j & 31)) & 1u) S.consider(true, i - wk.ci, j - wk.cj, wk.p0, wk.p1);
    if (S.have()) {
      int di, dj, k;
      S.finish(wk.p0, wk.p1, di, dj, k);
      wk.far_found(di, dj);
      return true;
    }
    if (i0 == 0 && j0 == 0 && i1 == wk.h - 1 && j1 == w - 1) return false;
  }
}

template <int MODE>
int walk_image(const int32_t *lab, int H, int W, int levels, bool use_lut, bool literal_far, int32_t *Q, int32_t *Pm, long long *steps_by_kind) {
  const int N = H * W;
  // regions in order of first appearance, row-major (Segmentation.compute_label_dict, rbepwt.py:840-848)
  std::unordered_map<int32_t, int> rid;
  std::vector<int> first, size, rmax, cmin, cmax;
  for (int p = 0; p < N; p++) {
    auto it = rid.find(lab[p]);
    int r;
    if (it == rid.end()) {
      r = (int)first.size();
      rid.emplace(lab[p], r);
      first.push_back(p); size.push_back(0); rmax.push_back(0); cmin.push_back(W); cmax.push_back(0);
    } else {
      r = it->second;
    }
    size[r]++;
    rmax[r] = std::max(rmax[r], p / W);
    cmin[r] = std::min(cmin[r], p % W);
    cmax[r] = std::max(cmax[r], p % W);
  }
  std::vector<uint8_t> lut(WK_LUT_BYTES);
  for (int e = 0; e < WK_LUT_BYTES; e++) {
    const int src = wk_lut_source(e);
    lut[e] = unit_lut_entry<MODE>(src / TPR_LUT_COLS, src % TPR_LUT_COLS);
  }
  std::vector<uint8_t> t2(T2_BYTES, 0xfe);
  for (int job = 0; job < T2_JOBS; job++) wk_t2_fill(t2.data(), job);
  int off = 0;
  for (size_t r = 0; r < first.size(); r++) {
    const int r0 = first[r] / W - WK_PAD, c0 = cmin[r] - WK_PAD;
    const int h = rmax[r] - r0 + 1 + WK_PAD, w = cmax[r] - c0 + 1 + WK_PAD, ws = (w + 31) >> 5;
    std::vector<uint32_t> slot(wk_slot_words(h, ws), 0u);
    const int32_t label = lab[first[r]];
    for (int i = WK_PAD; i < h - WK_PAD; i++)
      for (int j = WK_PAD; j < w - WK_PAD; j++)
        if (lab[(r0 + i) * W + c0 + j] == label) slot[i * ws + (j >> 5)] |= 1u << (j & 31);
    Walker<MODE> wk;
    wk.bm = slot.data();
    wk.Qall = Q; wk.pdelta = Pm - Q; wk.img = 0;
    wk.lut = use_lut && MODE != MODE_EUCLID ? lut.data() : nullptr;
    wk.t2 = use_lut && MODE == MODE_EUCLID ? t2.data() : nullptr;
    wk.N = N; wk.W = W; wk.L = levels;
    wk.abase = 0;
    wk.h = h; wk.ws = ws;
    wk.pixbase = r0 * W + c0;
    wk.narrow = ws == 1;
    wk.kind = WK_DONE;
    wk.start(off, size[r], WK_PAD, first[r] % W - c0);
    while (!wk.done()) {
      steps_by_kind[wk.kind]++;
      if (wk.kind == WK_LEVEL) wk.next_level();
      else if (wk.kind == WK_NEAR) wk.near_select();
      else if (wk.kind == WK_COMMIT) wk.commit_step();
      else if (wk.kind == WK_FAR) {
        if (literal_far) {
          if (!host_far(wk)) return -1;
        } else {  // the row-ordered restatement
          int di = 0, dj = 0;
          if (!wk_far_lane<MODE>(wk.bm + wk.cur, wk.h, wk.ws, wk.ci, wk.cj, wk.p0, wk.p1, di, dj)) return -1;
          wk.far_found(di, dj);
        }
      }
      else wk.list_step();
    }
    if (wk.kind == WK_ERROR) return -1;
    off += size[r];
  }
  return (int)first.size();
}

}  // namespace

// 1 if the multiply-shift hash of every class is a bijection of the class's subsets onto its table slots
extern "C" int hw_t2_hash_is_perfect(void) {
  for (int c = 0; c < T2_NCLS; c++) {
    const int nsub = c == 3 ? 256 : 16;
    std::vector<int> seen(nsub, 0);
    for (int i = 0; i < nsub; i++) {
      const int f = wk_t2_field(c, wk_t2_subset(c, i) | ~wk_t2_mask(c));  // other classes' bits must not matter
      if (f < 0 || f >= nsub || seen[f]++) return 0;
    }
  }
  return 1;
}

extern "C" int hw_walk_image(const int32_t *lab, int H, int W, int levels, int mode, int widewin, int use_lut, int32_t *Q,
                             int32_t *Pm, long long *steps_by_kind) {
  std::memset(steps_by_kind, 0, 8 * sizeof(long long));
  // `widewin` selects the search beyond the 5x5 window: 1 = the reference's probes taken literally, 0 = wk_far_lane
  if (mode == MODE_EUCLID) return walk_image<MODE_EUCLID>(lab, H, W, levels, use_lut, widewin != 0, Q, Pm, steps_by_kind);
  return walk_image<MODE_CHEB>(lab, H, W, levels, use_lut, widewin != 0, Q, Pm, steps_by_kind);
}
