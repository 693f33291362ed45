/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the RBEPWT encode -> threshold -> decode path.
 *
 * Plain-C restatement of the reference's algorithm (nareto/rbepwt, rbepwt.py) on flat
 * arrays.  It is the checker for the CUDA product in rbepwt_b200/csrc and the "port" CPU
 * baseline of bench.py; the product never links, imports or calls it.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.
 *
 * Pinned against the reference itself: the .npz files under tests/golden/ are produced by executing
 * the unmodified /root/reference/rbepwt.py (oracle/ref_harness.py, tests/golden/make_golden.py)
 * and tests/test_oracle_golden.py requires this file to reproduce them -- paths,
 * permutations and kept indices bit for bit, coefficients/pixels to 1e-9 relative.
 * The wavelet arithmetic (PyWavelets, absent from /root/reference and from this image)
 * is restated from its published algorithm: PARITY UNPINNED for that part, see
 * oracle/pywt_port.py.
 *
 * Reference map (file:line in /root/reference/rbepwt.py):
 *   build_regions      Segmentation.compute_label_dict            840-848
 *   easy_path          Region.easy_path, neighborhood, rotate     1273-1347, 84-104, 78-82
 *                      start point = lexicographic min (row,col)  1020-1036
 *   reduce (in encode) RegionCollection.reduce/Region.reduce_points 1563-1584, 1349-1375
 *   dwt_per/idwt_per   pywt.dwt/idwt 'periodization' call sites   2041, 2067
 *   rbo_encode         Rbepwt.encode                              1996-2053
 *   rbo_threshold      Rbepwt.threshold_coefs                     2081-2112
 *   rbo_decode         Rbepwt.decode, RegionCollection.expand,    2055-2079, 1586-1613,
 *                      Image.decode_rbepwt                        307-317
 *   rbo_psnr           psnr                                       156-162
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: the reference's tie-break uses one
 * explicit FMA -- numpy's ddot tail -- and nothing else may be contracted).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RBO_MODE_EUCLID 0 /* path_type='easypath', euclidean_distance=True  (rbepwt.py:1304) */
#define RBO_MODE_CHEB 1   /* path_type='easypath', euclidean_distance=False (rbepwt.py:1306) */
#define RBO_MODE_EPWT 2   /* path_type='epwt-easypath'                      (rbepwt.py:1302) */
#define RBO_MODE_GRAD_EUCLID 3 /* path_type='gradpath', euclidean_distance=True  (rbepwt.py:1190-1271) */
#define RBO_MODE_GRAD_CHEB 4   /* path_type='gradpath', euclidean_distance=False */

/* ------------------------------------------------------------------ label dict ---- */

/* Region id = rank of first appearance of the label in a row-major scan; pixels inside a
 * region in row-major order (rbepwt.py:840-848).  rid[] gets the region id of every pixel.
 * Returns R. */
static int build_region_ids(const int32_t *labels, int n, int32_t *rid) {
  int cap = 1;
  while (cap < 2 * n + 2) cap <<= 1;
  int32_t *keys = (int32_t *)malloc(sizeof(int32_t) * cap);
  int32_t *vals = (int32_t *)malloc(sizeof(int32_t) * cap);
  memset(vals, 0xff, sizeof(int32_t) * cap); /* -1 = empty */
  int R = 0;
  for (int p = 0; p < n; p++) {
    uint32_t h = ((uint32_t)labels[p] * 2654435761u) & (uint32_t)(cap - 1);
    while (vals[h] != -1 && keys[h] != labels[p]) h = (h + 1) & (uint32_t)(cap - 1);
    if (vals[h] == -1) {
      keys[h] = labels[p];
      vals[h] = R++;
    }
    rid[p] = vals[h];
  }
  free(keys);
  free(vals);
  return R;
}

int rbo_count_regions(const int32_t *labels, int H, int W) {
  int n = H * W;
  int32_t *rid = (int32_t *)malloc(sizeof(int32_t) * n);
  int R = build_region_ids(labels, n, rid);
  free(rid);
  return R;
}

/* ------------------------------------------------------------------ wavelet ------- */

/* cA[o] = sum_j dec_lo[j] x[(2o + F/2 - j) mod n], ascending j, mul then add, from 0. */
static void dwt_per(const double *x, int n, int flen, const double *dec_lo,
                    const double *dec_hi, double *ca, double *cd) {
  int half = n / 2;
  for (int o = 0; o < half; o++) {
    double a = 0.0, d = 0.0;
    for (int j = 0; j < flen; j++) {
      long idx = ((long)2 * o + flen / 2 - j) % n;
      if (idx < 0) idx += n;
      double xv = x[idx];
      a = a + dec_lo[j] * xv;
      d = d + dec_hi[j] * xv;
    }
    ca[o] = a;
    cd[o] = d;
  }
}

/* x[t] = S_lo + S_hi; taps m ascending with (t + F/2 - 1 - m) even, o = that / 2 mod n/2. */
static void idwt_per(const double *ca, const double *cd, int n, int flen,
                     const double *rec_lo, const double *rec_hi, double *x) {
  int half = n / 2;
  for (int t = 0; t < n; t++) {
    double slo = 0.0, shi = 0.0;
    int base = t + flen / 2 - 1;
    for (int m = base & 1; m < flen; m += 2) {
      long o = ((long)(base - m) / 2) % half; /* base-m even; may be negative */
      if (o < 0) o += half;
      slo = slo + rec_lo[m] * ca[o];
      shi = shi + rec_hi[m] * cd[o];
    }
    x[t] = slo + shi;
  }
}

/* ------------------------------------------------------------------ easy path ----- */

typedef struct {
  int H, W;
  int32_t *owner;  /* per pixel: region id if the pixel is an unvisited point of the level, else -1 */
  int32_t *idxmap; /* per pixel: index of the point in its region's incoming order */
  const double *valmap; /* per pixel value at this level (EPWT mode only) */
} path_ctx;

/* One region's path.  pix[0..n) are the region's points (pixel ids) in incoming order.
 * Writes perm[t] (index into the incoming order) and path_pix[t]. */
static void easy_path(path_ctx *c, int r, const int32_t *pix, int n, int mode, int u8wrap,
                      int32_t *perm, int32_t *path_pix) {
  if (n == 0) return;
  if (n == 1) { /* rbepwt.py:1275-1277 */
    perm[0] = 0;
    path_pix[0] = pix[0];
    return;
  }
  const int W = c->W, H = c->H;
  int start = 0, rmin = H, rmax = -1, cmin = W, cmax = -1;
  for (int i = 0; i < n; i++) {
    c->owner[pix[i]] = r;
    c->idxmap[pix[i]] = i;
    if (pix[i] < pix[start]) start = i; /* row-major id order == (row, col) lexicographic order */
    int rr = pix[i] / W, cc = pix[i] % W;
    if (rr < rmin) rmin = rr;
    if (rr > rmax) rmax = rr;
    if (cc < cmin) cmin = cc;
    if (cc > cmax) cmax = cc;
  }
  int ci = pix[start] / W, cj = pix[start] % W;
  c->owner[pix[start]] = -1;
  perm[0] = start;
  path_pix[0] = pix[start];
  long p0 = 0, p1 = 1; /* prefered_direc = (0,1), rbepwt.py:1290 */
  for (int t = 1; t < n; t++) {
    int found = 0, bi = 0, bj = 0;
    double bdist = 0.0, bsp1 = 0.0;
    long bcross = 0, bd2 = 0;
    double curval = (mode == RBO_MODE_EPWT) ? c->valmap[ci * W + cj] : 0.0;
    for (int rad = 1; !found; rad <<= 1) { /* half-width 2^(k-1), k = 1,2,..  rbepwt.py:1296-1299, 90-92 */
      int i0 = ci - rad < rmin ? rmin : ci - rad, i1 = ci + rad > rmax ? rmax : ci + rad;
      int j0 = cj - rad < cmin ? cmin : cj - rad, j1 = cj + rad > cmax ? cmax : cj + rad;
      for (int i = i0; i <= i1; i++) {
        for (int j = j0; j <= j1; j++) {
          if (c->owner[i * W + j] != r) continue;
          long di = i - ci, dj = j - cj;
          long d2 = di * di + dj * dj;
          double dist;
          if (mode == RBO_MODE_EPWT) {
            double dv = curval - c->valmap[i * W + j];
            if (u8wrap) dist = dv < 0 ? dv + 256.0 : dv; /* uint8 wrap, rbepwt.py:1302 on a uint8 image */
            else dist = fabs(dv);
          } else if (mode == RBO_MODE_EUCLID) {
            dist = (double)d2; /* norm compares like d2 */
          } else {
            long a = di < 0 ? -di : di, b = dj < 0 ? -dj : dj;
            dist = (double)(a > b ? a : b);
          }
          if (found && dist > bdist) continue;
          /* v = off/||off||; sp1 = np.dot(v, pref) = fma(v1, p1, v0*p0) (OpenBLAS ddot tail) */
          double nrm = sqrt((double)d2);
          double v0 = (double)di / nrm, v1 = (double)dj / nrm;
          double sp1 = fma(v1, (double)p1, v0 * (double)p0);
          long cross = di * p1 - dj * p0; /* orders like v . rotate(pref,-pi/2), rbepwt.py:1320-1322 */
          int better;
          if (!found || dist < bdist) better = 1;
          else if (sp1 != bsp1) better = sp1 > bsp1;
          else if (cross != bcross) better = cross > bcross;
          else better = d2 < bd2; /* complete tie (EPWT, collinear, equal |dv|): unpinned in the
                                     reference (CPython set order); our rule: nearer first */
          if (better) {
            found = 1; bi = i; bj = j; bdist = dist; bsp1 = sp1; bcross = cross; bd2 = d2;
          }
        }
      }
    }
    int pid = bi * W + bj;
    c->owner[pid] = -1;
    perm[t] = c->idxmap[pid];
    path_pix[t] = pid;
    p0 = bi - ci; p1 = bj - cj; /* rbepwt.py:1331: integer vector, not normalised */
    ci = bi; cj = bj;
  }
}

/* ------------------------------------------------------------------ grad path ----- */

/* Region.grad_path (rbepwt.py:1190-1271).  Same probes as easy_path; among the candidates of the first non-empty
 * probe the smallest distance wins (euclid: np.linalg.norm of the integer offset, i.e. d2 compares; chebyshev), and
 * equal distances are settled by the region's average gradient direction: with perp = rotate(avg_gradient, -pi/2)
 * and v = offset / ||offset||, the larger |np.dot(v, perp)| wins, then the larger |np.dot(v, rotate(perp, -pi/2))|
 * (rbepwt.py:1231-1247).  The sign flips of perp (1256-1259) cannot change a choice: only absolute values are compared.
 * A COMPLETE tie (both absolute dot products bit-identical: always for two opposite offsets v and -v) keeps whichever
 * candidate the reference's `for candidate in candidate_points` met first, i.e. CPython set order: unpinned.  Rule
 * here: the first in row-major order wins; such events are counted so that fixtures can be chosen without any.
 * np.dot of two 2-vectors is fma(x1, y1, x0*y0) (OpenBLAS ddot tail; checked against numpy in this container). */
static long g_grad_tie_events = 0;
long rbo_grad_tie_events(void) { return g_grad_tie_events; }

static void rotate_mhalfpi(const double v[2], double out[2]) { /* rbepwt.py:78-82 with theta = -pi/2 */
  const double c = 6.123233995736766e-17, s = -1.0; /* np.cos(-np.pi/2), np.sin(-np.pi/2) */
  out[0] = fma(-s, v[1], c * v[0]);
  out[1] = fma(c, v[1], s * v[0]);
}

static void grad_path(path_ctx *c, int r, const int32_t *pix, int n, int cheb, const double avg_grad[2],
                      int32_t *perm, int32_t *path_pix) {
  if (n == 0) return;
  if (n == 1) { perm[0] = 0; path_pix[0] = pix[0]; return; }
  const int W = c->W, H = c->H;
  int start = 0, rmin = H, rmax = -1, cmin = W, cmax = -1;
  for (int i = 0; i < n; i++) {
    c->owner[pix[i]] = r;
    c->idxmap[pix[i]] = i;
    if (pix[i] < pix[start]) start = i;
    int rr = pix[i] / W, cc = pix[i] % W;
    if (rr < rmin) rmin = rr;
    if (rr > rmax) rmax = rr;
    if (cc < cmin) cmin = cc;
    if (cc > cmax) cmax = cc;
  }
  int ci = pix[start] / W, cj = pix[start] % W;
  c->owner[pix[start]] = -1;
  perm[0] = start;
  path_pix[0] = pix[start];
  double perp[2], tmp[2];
  rotate_mhalfpi(avg_grad, perp); /* rbepwt.py:1207 */
  rotate_mhalfpi(perp, tmp);      /* rbepwt.py:1243 */
  for (int t = 1; t < n; t++) {
    int found = 0, bi = 0, bj = 0;
    long bdist = 0;
    double bA = 0.0, bB = 0.0;
    for (int rad = 1; !found; rad <<= 1) {
      int i0 = ci - rad < rmin ? rmin : ci - rad, i1 = ci + rad > rmax ? rmax : ci + rad;
      int j0 = cj - rad < cmin ? cmin : cj - rad, j1 = cj + rad > cmax ? cmax : cj + rad;
      for (int i = i0; i <= i1; i++)
        for (int j = j0; j <= j1; j++) {
          if (c->owner[i * W + j] != r) continue;
          long di = i - ci, dj = j - cj, d2 = di * di + dj * dj;
          long a = di < 0 ? -di : di, b = dj < 0 ? -dj : dj;
          long dist = cheb ? (a > b ? a : b) : d2;
          if (found && dist > bdist) continue;
          double nrm = sqrt((double)d2);
          double v0 = (double)di / nrm, v1 = (double)dj / nrm;
          double A = fabs(fma(v1, perp[1], v0 * perp[0])), B = fabs(fma(v1, tmp[1], v0 * tmp[0]));
          int better;
          if (!found || dist < bdist) better = 1;
          else if (A > bA) better = 1;
          else if (A == bA && B > bB) better = 1;
          else {
            better = 0;
            if (!(A < bA) && !(A == bA && B < bB)) g_grad_tie_events++; /* complete tie (or NaN direction): unpinned */
          }
          if (better) { found = 1; bi = i; bj = j; bdist = dist; bA = A; bB = B; }
        }
    }
    int pid = bi * W + bj;
    c->owner[pid] = -1;
    perm[t] = c->idxmap[pid];
    path_pix[t] = pid;
    ci = bi; cj = bj;
  }
}

/* Region.compute_avg_gradient (rbepwt.py:1102-1113) over np.gradient(img) (2010-2013): sums in the region's level-1
 * point order (row-major), divided by the count, normalised by np.linalg.norm = sqrt(fma(g1, g1, g0*g0)). */
static void region_avg_gradients(const double *img, int H, int W, const int32_t *inc_pix, const int32_t *off, int R,
                                 double *avg /* [R][2] */) {
  for (int r = 0; r < R; r++) {
    double s0 = 0.0, s1 = 0.0;
    for (int q = off[r]; q < off[r + 1]; q++) {
      int i = inc_pix[q] / W, j = inc_pix[q] % W;
      double g0, g1;
      if (i == 0) g0 = img[W + j] - img[j];
      else if (i == H - 1) g0 = img[i * W + j] - img[(i - 1) * W + j];
      else g0 = (img[(i + 1) * W + j] - img[(i - 1) * W + j]) / 2.0;
      if (j == 0) g1 = img[i * W + 1] - img[i * W];
      else if (j == W - 1) g1 = img[i * W + j] - img[i * W + j - 1];
      else g1 = (img[i * W + j + 1] - img[i * W + j - 1]) / 2.0;
      s0 += g0; s1 += g1;
    }
    const double np_ = (double)(off[r + 1] - off[r]);
    s0 /= np_; s1 /= np_;
    const double nrm = sqrt(fma(s1, s1, s0 * s0));
    avg[2 * r] = s0 / nrm; avg[2 * r + 1] = s1 / nrm;
  }
}

/* ------------------------------------------------------------------ encode -------- */

static long level_len(long n, int lev) { return n >> (lev - 1); } /* lev = 1.. */

/* Offsets of level `lev` (1-based) inside the per-level concatenated buffers. */
long rbo_level_offset(long n, int lev) {
  long off = 0;
  for (int l = 1; l < lev; l++) off += level_len(n, l);
  return off;
}

/*
 * Outputs (caller-allocated), with N = H*W, N_l = N >> (l-1), lo(l) = rbo_level_offset(N, l):
 *   roff    [(levels+1) * (R+1)]   region offsets at level l = 1..levels+1 (row l-1)
 *   inc_pix [lo(levels+2)]         pixel ids in INCOMING order at level l = 1..levels+1
 *   path_pix[lo(levels+1)]         pixel ids in PATH order at level l = 1..levels
 *   perm    [lo(levels+1)]         per-region local permutation at level l (Region.permutation)
 *   coefs   [N]                    details[1] | ... | details[levels] | approx
 * labels == NULL  <=>  one region holding every pixel (EPWT).
 */
int rbo_encode(const double *img, const int32_t *labels, int H, int W, int levels, int flen,
               const double *dec_lo, const double *dec_hi, int mode, int u8wrap, int paths_first_level, int R,
               int32_t *roff, int32_t *inc_pix, int32_t *path_pix, int32_t *perm,
               double *coefs) {
  const int N = H * W;
  if (N <= 0 || (N & (N - 1))) return -1;       /* rbepwt.py:301-302 */
  if (levels < 1 || levels > 30 || ((long)1 << levels) > N) return -2; /* rbepwt.py:1977-1978 */
  int32_t *rid = (int32_t *)malloc(sizeof(int32_t) * N);
  if (labels) {
    int R2 = build_region_ids(labels, N, rid);
    if (R2 != R) { free(rid); return -3; }
  } else {
    if (R != 1) { free(rid); return -3; }
    memset(rid, 0, sizeof(int32_t) * N);
  }
  /* level-1 incoming order: regions by first appearance, pixels row-major (counting sort) */
  int32_t *off = roff;
  memset(off, 0, sizeof(int32_t) * (R + 1));
  for (int p = 0; p < N; p++) off[rid[p] + 1]++;
  for (int r = 0; r < R; r++) off[r + 1] += off[r];
  int32_t *fill = (int32_t *)malloc(sizeof(int32_t) * (R + 1));
  memcpy(fill, off, sizeof(int32_t) * (R + 1));
  for (int p = 0; p < N; p++) inc_pix[fill[rid[p]]++] = p;
  free(fill);
  free(rid);

  path_ctx c;
  c.H = H; c.W = W;
  c.owner = (int32_t *)malloc(sizeof(int32_t) * N);
  c.idxmap = (int32_t *)malloc(sizeof(int32_t) * N);
  double *valmap = (double *)malloc(sizeof(double) * N);
  c.valmap = valmap;
  memset(c.owner, 0xff, sizeof(int32_t) * N);
  double *val = (double *)malloc(sizeof(double) * N); /* values in incoming order */
  double *sig = (double *)malloc(sizeof(double) * N); /* values in path order */
  double *ca = (double *)malloc(sizeof(double) * N);
  for (int i = 0; i < N; i++) val[i] = img[inc_pix[i]];

  double *avg_grad = NULL;
  if (mode == RBO_MODE_GRAD_EUCLID || mode == RBO_MODE_GRAD_CHEB) {
    if (H < 2 || W < 2) { free(c.owner); free(c.idxmap); free(valmap); free(val); free(sig); free(ca); return -4; } /* np.gradient needs 2 samples */
    avg_grad = (double *)malloc(sizeof(double) * 2 * R);
    region_avg_gradients(img, H, W, inc_pix, roff, R, avg_grad);
    g_grad_tie_events = 0;
  }
  long coef_off = 0;
  for (int lev = 1; lev <= levels; lev++) {
    const long nl = level_len(N, lev), lo = rbo_level_offset(N, lev);
    const int32_t *cur_off = roff + (long)(lev - 1) * (R + 1);
    int32_t *nxt_off = roff + (long)lev * (R + 1);
    const int32_t *ipix = inc_pix + lo;
    int32_t *ppix = path_pix + lo, *pm = perm + lo;
    if (mode == RBO_MODE_EPWT)
      for (long i = 0; i < nl; i++) valmap[ipix[i]] = val[i];
    for (int r = 0; r < R; r++) {
      int a = cur_off[r], n = cur_off[r + 1] - a;
      if (lev > 1 && paths_first_level) { /* Region.same_path: identity permutation (rbepwt.py:1183-1188, 2024-2025) */
        for (int t = 0; t < n; t++) { pm[a + t] = t; ppix[a + t] = ipix[a + t]; }
      } else if (avg_grad) {
        grad_path(&c, r, ipix + a, n, mode == RBO_MODE_GRAD_CHEB, avg_grad + 2 * r, pm + a, ppix + a);
      } else {
        easy_path(&c, r, ipix + a, n, mode, u8wrap && lev == 1, pm + a, ppix + a);
      }
      for (int t = 0; t < n; t++) sig[a + t] = val[a + pm[a + t]];
    }
    dwt_per(sig, (int)nl, flen, dec_lo, dec_hi, ca, coefs + coef_off);
    coef_off += nl / 2;
    /* reduce: keep the points at even global position g; they carry cA[g/2] (rbepwt.py:1563-1584) */
    int32_t *npix = inc_pix + rbo_level_offset(N, lev + 1);
    for (int r = 0; r <= R; r++) nxt_off[r] = (cur_off[r] + 1) / 2;
    for (long g = 0; g < nl; g += 2) {
      npix[g / 2] = ppix[g];
      val[g / 2] = ca[g / 2];
    }
  }
  memcpy(coefs + coef_off, val, sizeof(double) * level_len(N, levels + 1));
  free(c.owner); free(c.idxmap); free(valmap); free(val); free(sig); free(ca); free(avg_grad);
  return 0;
}

/* ------------------------------------------------------------------ threshold ----- */

typedef struct { double mag; int64_t idx; } mag_idx;
static int cmp_mag_desc(const void *a, const void *b) {
  const mag_idx *x = (const mag_idx *)a, *y = (const mag_idx *)b;
  if (x->mag != y->mag) return x->mag > y->mag ? -1 : 1;
  return x->idx > y->idx ? -1 : (x->idx < y->idx ? 1 : 0); /* ties: highest flat index first */
}

/* Keep the k largest |coef|, zero the rest, in place (rbepwt.py:2081-2112).  k <= 0 or k >= n
 * keeps everything (the reference's `counter == ncoefs` test never fires).  Ties at the k-th
 * magnitude are unpinned in the reference (unstable argsort); rule here: highest index first. */
int rbo_threshold(double *coefs, int64_t n, int64_t k) {
  if (k <= 0 || k >= n) return 0;
  mag_idx *v = (mag_idx *)malloc(sizeof(mag_idx) * n);
  for (int64_t i = 0; i < n; i++) { v[i].mag = fabs(coefs[i]); v[i].idx = i; }
  qsort(v, n, sizeof(mag_idx), cmp_mag_desc);
  for (int64_t i = k; i < n; i++) coefs[v[i].idx] = 0.0;
  free(v);
  return 0;
}

/* Rbepwt.threshold_by_percentage(perc) (rbepwt.py:2120-2192): "keeps only perc proportion of coefficients for each
 * region".  Region r owns, of every level l = 1..L, the detail coefficients at the positions of its level-(l+1) segment,
 * and the approximation coefficients of its level-(L+1) segment (2124-2148); of these n_r values the
 * int(min(floor(perc * n_r + 0.5), n_r)) largest in magnitude are kept (2155: myround), the others zeroed -- in the DETAILS
 * only: the thresholded approximation is assigned to the level-(L+1) collection and then lost again, because
 * RegionCollection.update() (2187, 1529-1531) rebuilds the collection's values from its sub-regions, which still hold the
 * original ones.  So approximation coefficients take part in the ranking but are never zeroed (checked by running the
 * reference: the fixtures).  Ties at the cut are broken by numpy's unstable argsort (unpinned); here the entry later in
 * the region's list (levels ascending, approximation last) survives.
 * roff: [(levels+1)][R+1] as produced by rbo_encode. */
typedef struct { double mag; long pos; } mag_pos;
static int cmp_mag_pos_desc(const void *a, const void *b) {
  const mag_pos *x = (const mag_pos *)a, *y = (const mag_pos *)b;
  if (x->mag != y->mag) return x->mag > y->mag ? -1 : 1;
  return x->pos > y->pos ? -1 : (x->pos < y->pos ? 1 : 0);
}

int rbo_threshold_percentage(double *coefs, int H, int W, int levels, int R, const int32_t *roff, double perc) {
  const long N = (long)H * W;
  mag_pos *buf = (mag_pos *)malloc(sizeof(mag_pos) * (N + 1));
  long *flat = (long *)malloc(sizeof(long) * (N + 1));
  for (int r = 0; r < R; r++) {
    long n = 0;
    long det_off = 0;
    for (int lev = 1; lev <= levels; lev++) { /* details[lev] has N >> lev entries, segment of the level-(lev+1) offsets */
      const int32_t *o = roff + (long)lev * (R + 1);
      for (long q = o[r]; q < o[r + 1]; q++) { flat[n] = det_off + q; n++; }
      det_off += N >> lev;
    }
    const long n_det = n;
    const int32_t *o = roff + (long)levels * (R + 1);
    for (long q = o[r]; q < o[r + 1]; q++) { flat[n] = det_off + q; n++; }
    if (n == 0) continue;
    for (long i = 0; i < n; i++) { buf[i].mag = fabs(coefs[flat[i]]); buf[i].pos = i; }
    long keep = (long)floor(perc * (double)n + 0.5);
    if (keep > n) keep = n;
    if (keep < 0) keep = 0;
    qsort(buf, n, sizeof(mag_pos), cmp_mag_pos_desc);
    for (long i = keep; i < n; i++)
      if (buf[i].pos < n_det) coefs[flat[buf[i].pos]] = 0.0; /* approximation entries are never zeroed (see above) */
  }
  free(buf); free(flat);
  return 0;
}

/* ------------------------------------------------------------------ decode -------- */

int rbo_decode(const double *coefs, int H, int W, int levels, int flen, const double *rec_lo,
               const double *rec_hi, int R, const int32_t *roff, const int32_t *inc_pix1,
               const int32_t *perm, double *out_img) {
  const int N = H * W;
  double *x = (double *)malloc(sizeof(double) * N);
  double *s = (double *)malloc(sizeof(double) * N);
  long coef_off = N - level_len(N, levels + 1);
  memcpy(x, coefs + coef_off, sizeof(double) * level_len(N, levels + 1));
  for (int lev = levels; lev >= 1; lev--) {
    const long nl = level_len(N, lev), lo = rbo_level_offset(N, lev);
    coef_off -= nl / 2;
    idwt_per(x, coefs + coef_off, (int)nl, flen, rec_lo, rec_hi, s);
    const int32_t *cur_off = roff + (long)(lev - 1) * (R + 1);
    for (int r = 0; r < R; r++) { /* expand: undo the region's permutation (rbepwt.py:1600-1608) */
      int a = cur_off[r], n = cur_off[r + 1] - a;
      for (int t = 0; t < n; t++) x[a + perm[lo + a + t]] = s[a + t];
    }
  }
  for (int i = 0; i < N; i++) { /* rbepwt.py:309-314: no rounding, clip to [0,255] */
    double v = x[i];
    if (v > 255.0) v = 255.0;
    if (v < 0.0) v = 0.0;
    out_img[inc_pix1[i]] = v;
  }
  free(x); free(s);
  return 0;
}

/* psnr (rbepwt.py:156-162): -1 when the images are identical. */
double rbo_psnr(const double *a, const double *b, int64_t n) {
  double mse = 0.0;
  for (int64_t i = 0; i < n; i++) { double d = a[i] - b[i]; mse += d * d; }
  if (mse == 0.0) return -1.0;
  mse /= (double)n;
  return 20.0 * log10(255.0 / sqrt(mse));
}
