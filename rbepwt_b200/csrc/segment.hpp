// Felzenszwalb-Huttenlocher graph segmentation of a grayscale image -- what the reference's Image.segment calls
// (skimage.segmentation.felzenszwalb, /root/reference/rbepwt.py:779-785, defaults scale=200 sigma=2 min_size=10 at 224).
// HOST code, like the reference's (scikit-image's is Cython): it runs once per image, before the path this library
// accelerates, and the greedy merge over the sorted edges is sequential by nature.
//
// scikit-image is a third-party dependency that is neither under /root/reference nor installed here: this restates the
// published algorithm (Felzenszwalb & Huttenlocher, IJCV 2004) in the form scikit-image gives it --
//   1. image as float64 (the caller passes what img_as_float64 would: uint8 / 255), scale /= 255;
//   2. Gaussian smoothing like scipy.ndimage.gaussian_filter: separable, kernel radius int(4 sigma + 0.5), weights
//      exp(-x^2 / (2 sigma^2)) normalised, borders reflected (d c b a | a b c d | d c b a), rows axis first;
//   3. 8-connectivity edges (right, down, down-right, up-right), weight = |difference| of the smoothed values;
//   4. edges in ascending weight (equal weights: scikit-image's order comes from an unstable argsort -- unpinned; here
//      the order of step 3's list, i.e. a stable sort); an edge joins two components when its weight is below
//      min(int(C0) + scale / |C0|, int(C1) + scale / |C1|), the new component's int() is that weight;
//   5. a second pass over the edges joins every component of fewer than min_size pixels to its neighbour;
//   6. labels = rank of a component's root in ascending order (np.unique(..., return_inverse=True)); the root of a
//      component is its smallest pixel index, so labels are numbered in order of first appearance, row-major.
// PARITY UNPINNED against scikit-image itself (nothing to run it against here); tests/ pin it to a numpy restatement
// (oracle/fh_port.py) and to the algorithm's invariants.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace rbepwt {

inline int fh_reflect(int i, int n) {  // scipy's 'reflect': ... 1 0 | 0 1 2 ... n-1 | n-1 n-2 ...
  if (n == 1) return 0;
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

// one axis of scipy.ndimage.gaussian_filter (correlate1d with a symmetric kernel: centre first, then the pairs from the
// farthest to the nearest)
inline void fh_gauss_axis(const std::vector<double> &in, std::vector<double> &out, int H, int W, int axis,
                          const std::vector<double> &w, int radius) {
  const int n = axis == 0 ? H : W, m = axis == 0 ? W : H;
  const size_t step = axis == 0 ? (size_t)W : 1, other = axis == 0 ? 1 : (size_t)W;
  std::vector<double> line(n + 2 * radius);
  for (int q = 0; q < m; q++) {
    const double *src = in.data() + q * other;
    for (int i = -radius; i < n + radius; i++) line[i + radius] = src[(size_t)fh_reflect(i, n) * step];
    double *dst = out.data() + q * other;
    for (int i = 0; i < n; i++) {
      const double *c = line.data() + i + radius;
      double t = c[0] * w[radius];
      for (int k = -radius; k < 0; k++) t += (c[k] + c[-k]) * w[k + radius];
      dst[(size_t)i * step] = t;
    }
  }
}

inline int fh_find(std::vector<int32_t> &f, int i) {
  int r = i;
  while (f[r] != r) r = f[r];
  while (f[i] != r) { const int nx = f[i]; f[i] = r; i = nx; }  // path compression
  return r;
}

// img: H*W float64 (already in the range scikit-image would see), labels: H*W int32 out.  Returns the number of segments.
inline int felzenszwalb(const double *img, int H, int W, double scale, double sigma, int min_size, int32_t *labels) {
  const size_t N = (size_t)H * W;
  scale /= 255.0;
  std::vector<double> a(img, img + N), b(N);
  if (sigma > 0.0) {
    const int radius = (int)(4.0 * sigma + 0.5);
    std::vector<double> w(2 * radius + 1);
    double sum = 0.0;
    for (int x = -radius; x <= radius; x++) { w[x + radius] = std::exp(-0.5 / (sigma * sigma) * (double)(x * x)); sum += w[x + radius]; }
    for (double &v : w) v /= sum;
    fh_gauss_axis(a, b, H, W, 0, w, radius);
    fh_gauss_axis(b, a, H, W, 1, w, radius);
  }
  // edges in scikit-image's order: right, down, down-right, up-right; (first, second) as it lists them
  struct Edge { double cost; int32_t p, q; };
  std::vector<Edge> e;
  e.reserve(4 * N);
  auto cost = [&](size_t p, size_t q) { return std::sqrt((a[p] - a[q]) * (a[p] - a[q])); };
  for (int i = 0; i < H; i++)
    for (int j = 1; j < W; j++) e.push_back({cost((size_t)i * W + j, (size_t)i * W + j - 1), i * W + j, i * W + j - 1});
  for (int i = 1; i < H; i++)
    for (int j = 0; j < W; j++) e.push_back({cost((size_t)i * W + j, (size_t)(i - 1) * W + j), i * W + j, (i - 1) * W + j});
  for (int i = 1; i < H; i++)
    for (int j = 1; j < W; j++)
      e.push_back({cost((size_t)i * W + j, (size_t)(i - 1) * W + j - 1), i * W + j, (i - 1) * W + j - 1});
  for (int i = 1; i < H; i++)
    for (int j = 0; j + 1 < W; j++)
      e.push_back({cost((size_t)i * W + j, (size_t)(i - 1) * W + j + 1), (i - 1) * W + j + 1, i * W + j});
  std::stable_sort(e.begin(), e.end(), [](const Edge &x, const Edge &y) { return x.cost < y.cost; });
  std::vector<int32_t> forest(N), size(N, 1);
  std::iota(forest.begin(), forest.end(), 0);
  std::vector<double> cint(N, 0.0);
  auto join = [&](int r0, int r1) {  // the smaller index becomes the root
    const int root = std::min(r0, r1), child = std::max(r0, r1);
    forest[child] = root;
    size[root] = size[r0] + size[r1];
    return root;
  };
  for (const Edge &ed : e) {
    const int r0 = fh_find(forest, ed.p), r1 = fh_find(forest, ed.q);
    if (r0 == r1) continue;
    const double in0 = cint[r0] + scale / size[r0], in1 = cint[r1] + scale / size[r1];
    if (ed.cost < std::min(in0, in1)) cint[join(r0, r1)] = ed.cost;
  }
  for (const Edge &ed : e) {
    const int r0 = fh_find(forest, ed.p), r1 = fh_find(forest, ed.q);
    if (r0 == r1) continue;
    if (size[r0] < min_size || size[r1] < min_size) join(r0, r1);
  }
  // roots ascending = first appearance, row-major
  std::vector<int32_t> rank(N, -1);
  int nseg = 0;
  for (size_t p = 0; p < N; p++)
    if (fh_find(forest, (int)p) == (int)p) rank[p] = nseg++;
  for (size_t p = 0; p < N; p++) labels[p] = rank[fh_find(forest, (int)p)];
  return nseg;
}

}  // namespace rbepwt
