/*
 * rbepwt_b200 -- C ABI of the B200-native RBEPWT encode -> threshold -> decode path.
 *
 * The reference (nareto/rbepwt) is a single Python module with no FFI of its own
 * (SURVEY.md section 8b); this is the boundary a maintainer binds with ctypes underneath
 * rbepwt.Image.  Each entry point names the reference code it replaces
 * (file:line in rbepwt.py).  Plain pointers and sizes only; no torch / numpy types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative RBEPWT_E_* code on failure;
 *     rbepwt_last_error() returns a thread-local message for the last failure.
 *   - one context = one GPU + one CUDA stream + device-resident state of ONE encoded batch
 *     (B images of the same H x W, levels, wavelet, path mode).  Not thread-safe per context;
 *     use one context per host thread / per GPU.
 *   - `flags & RBEPWT_DEVICE_PTRS`: img/labels/out pointers are DEVICE pointers on the
 *     context's GPU (no copy is made; inputs must stay valid until the call returns);
 *     otherwise they are HOST pointers and the call copies through the context's stream.
 *   - images are float64 [B][H][W] row-major; labels int32 [B][H][W]; H*W must be a power
 *     of two and 2^levels <= H*W (rbepwt.py:301-302, 1977-1978).
 *   - calls are asynchronous on the context's stream when given device pointers; host-pointer
 *     calls return after their copies completed.  rbepwt_sync() waits for everything.
 */
#ifndef RBEPWT_B200_H
#define RBEPWT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rbepwt_ctx rbepwt_ctx;

/* error codes */
#define RBEPWT_OK 0
#define RBEPWT_E_CUDA (-1)        /* CUDA runtime error, see rbepwt_last_error() */
#define RBEPWT_E_NOT_POW2 (-2)    /* "Image size must be a power of 2"            rbepwt.py:301-302 */
#define RBEPWT_E_LEVELS (-3)      /* "2^levels must be smaller or equal ..."      rbepwt.py:1977-1978 */
#define RBEPWT_E_NO_ENCODING (-4) /* "There is no saved encoding to decode"       rbepwt.py:2057-2058 */
#define RBEPWT_E_ARG (-5)         /* bad argument */
#define RBEPWT_E_NO_WAVELET (-6)  /* rbepwt_set_wavelet not called */
#define RBEPWT_E_NO_GPU (-7)      /* no usable CUDA device: there is NO CPU fallback */

/* path modes: Region.easy_path distance rule                                      rbepwt.py:1301-1306 */
#define RBEPWT_PATH_EUCLID 0 /* path_type='easypath', euclidean_distance=True  */
#define RBEPWT_PATH_CHEB 1   /* path_type='easypath', euclidean_distance=False */
#define RBEPWT_PATH_EPWT 2   /* path_type='epwt-easypath' (labels ignored, one region) */
#define RBEPWT_PATH_GRAD 3      /* path_type='gradpath', euclidean_distance=True: Region.grad_path  rbepwt.py:1190-1271 */
#define RBEPWT_PATH_GRAD_CHEB 4 /* path_type='gradpath', euclidean_distance=False.  Complete ties of the gradient
                                   preference depend on CPython set order in the reference (unpinned); here the first
                                   candidate in row-major order wins.  Every region is walked by one warp. */

/* flags */
#define RBEPWT_DEVICE_PTRS 1u /* pointer arguments are device pointers */
#define RBEPWT_U8_WRAP 2u     /* EPWT level 1: |a-b| wraps modulo 256 like numpy uint8 scalars (rbepwt.py:1302) */
#define RBEPWT_PATHS_FIRST_LEVEL 4u /* paths_first_level=True: Region.same_path (identity permutation) at every
                                       level >= 2, paths are searched at level 1 only (rbepwt.py:1183-1188, 2024-2025) */

#define RBEPWT_NO_CLIP 8u /* rbepwt_decode: the values Rbepwt.decode returns (rbepwt.py:2055-2079), without the clip to
                            [0,255] that Image.decode_rbepwt applies afterwards (rbepwt.py:313-314) */

/* Create a context on CUDA device `device`.  `stream` is a cudaStream_t to run on (e.g.
 * torch's current stream) or NULL to create a private one.  Fails with RBEPWT_E_NO_GPU when
 * there is no CUDA device -- the library has no CPU path. */
int rbepwt_create(int device, void *stream, rbepwt_ctx **out);
void rbepwt_destroy(rbepwt_ctx *ctx);
const char *rbepwt_last_error(void);
int rbepwt_sync(rbepwt_ctx *ctx);

/* Filter bank of the wavelet, PyWavelets convention (pywt.Wavelet(name).filter_bank):
 * replaces the `wavelet` argument of pywt.dwt / pywt.idwt                     rbepwt.py:2041, 2067 */
int rbepwt_set_wavelet(rbepwt_ctx *ctx, int filter_len, const double *dec_lo, const double *dec_hi,
                       const double *rec_lo, const double *rec_hi);

/* Image.encode_rbepwt / Rbepwt.encode for a batch                     rbepwt.py:298-305, 1996-2053
 * (+ Segmentation.compute_label_dict 840-848, Region.easy_path 1273-1347,
 *    RegionCollection.reduce 1563-1584, pywt.dwt 2041).
 * labels may be NULL iff path_mode == RBEPWT_PATH_EPWT. */
int rbepwt_encode(rbepwt_ctx *ctx, const double *img, const int32_t *labels, int B, int H, int W,
                  int levels, int path_mode, unsigned flags);

/* Execution options (set between calls).
 *   RBEPWT_OPT_STREAMS : 1 or 2 (default 2) units of each kind in flight: the batch is cut into path groups and
 *                        transform sub-batches whose copies, path kernels and transform kernels overlap on
 *                        internal streams; every call is still ordered on the context's stream.
 *                        1 = all kernels on one stream (per-kernel timing).
 *   RBEPWT_OPT_SUBBATCH: images per transform sub-batch (0 = auto: about 2^24 pixels).
 *   RBEPWT_OPT_PATHGROUP: images per path group -- label scan + path pyramid (0 = auto: about 2^26 pixels;
 *                        rounded to a multiple of the sub-batch).
 *   RBEPWT_OPT_COOP_LIMIT: how many regions of a path group (the largest ones, of >= 2048 pixels) are walked by a whole
 *                        warp instead of one lane of the path kernel (-1 = auto: 4 per SM, and every region of a group
 *                        of at most 4096 regions; 0 = none).  A tuning knob: results do not depend on it. */
#define RBEPWT_OPT_STREAMS 1
#define RBEPWT_OPT_SUBBATCH 2
#define RBEPWT_OPT_PATHGROUP 3
#define RBEPWT_OPT_COOP_LIMIT 4
int rbepwt_set_option(rbepwt_ctx *ctx, int option, int64_t value);

/* Image.segment(method='felzenszwalb', scale, sigma, min_size)                rbepwt.py:220-245, 779-785
 * The label map the path starts from, when the caller has none: Felzenszwalb-Huttenlocher graph segmentation as
 * scikit-image's felzenszwalb() computes it for a 2-D image.  HOST code and host pointers (no context, no GPU): like the
 * reference's, it runs once per image before the accelerated path.  img: float64 [H][W] as scikit-image's
 * img_as_float64 would deliver it (a uint8 image divided by 255; a float image as it is); labels: int32 [H][W] out,
 * numbered in order of first appearance; *nlabels (may be NULL): their count.  Parity with scikit-image is unpinned
 * (it is not available to test against): csrc/segment.hpp states what is restated. */
int rbepwt_felzenszwalb(const double *img, int H, int W, double scale, double sigma, int min_size, int32_t *labels,
                        int32_t *nlabels);

/* Rbepwt.threshold_coefs(ncoefs), per image                                 rbepwt.py:2081-2112
 * k <= 0 or k >= H*W keeps everything (reference quirk).  Ties at the k-th magnitude are
 * unpinned in the reference; here the highest flat index survives. */
int rbepwt_threshold(rbepwt_ctx *ctx, int64_t k);

/* The tensor-product baseline the reference compares against: Dwt.encode = pywt.wavedec2(img, wavelet, level=levels,
 * mode='periodization')                                                            rbepwt.py:2249-2263, 318-324
 * for a batch of square images.  The context then holds THIS encoding: rbepwt_threshold (Dwt.threshold_coefs, 2265-2298:
 * same top-k rule and quirks), rbepwt_decode (Dwt.decode = pywt.waverec2 + the clip of Image.decode_dwt, 326-333),
 * rbepwt_get_coefs / set_coefs / nonzero_coefs / psnr work on it; everything about paths and regions does not.
 * Coefficient layout of get/set_coefs: one H x W pyramid per image -- level l's sub-bands are the quadrants of the
 * top-left block of side W >> (l-1): [cA | cV ; cH | cD] in pywt's naming (see csrc/dwt2.cuh). */
int rbepwt_dwt2_encode(rbepwt_ctx *ctx, const double *img, int B, int H, int W, int levels, unsigned flags);

/* Rbepwt.threshold_by_percentage(perc), per image and region                  rbepwt.py:2120-2192
 * Of the n coefficients a region owns (its segment of every detail level and of the approximation) the
 * int(min(floor(perc*n + 0.5), n)) largest in magnitude are kept, the other DETAIL coefficients zeroed; approximation
 * coefficients take part in the ranking but always survive, which is what the reference does (its
 * RegionCollection.update() at 2187 discards the thresholded approximation).  Ties: unpinned in the reference; here
 * the entries later in the region's list (levels ascending, approximation last) survive. */
int rbepwt_threshold_percentage(rbepwt_ctx *ctx, double perc);

/* Rbepwt.decode + RegionCollection.expand + Image.decode_rbepwt (clip to [0,255], no rounding)
 *                                                           rbepwt.py:2055-2079, 1586-1613, 307-317
 * out: float64 [B][H][W].  flags: RBEPWT_DEVICE_PTRS, RBEPWT_NO_CLIP. */
int rbepwt_decode(rbepwt_ctx *ctx, double *out_img, unsigned flags);

/* encode -> threshold(k) -> decode in ONE call: the three reference calls above back to back
 * (rbepwt.py:298, 441, 307), with the sub-batches pipelined so that, for host pointers, the input copy of
 * one sub-batch, the kernels of the next and the output copy of the previous overlap.  Leaves the same
 * state as the three calls (thresholded coefficients, paths).  out_img: float64 [B][H][W]. */
int rbepwt_transcode(rbepwt_ctx *ctx, const double *img, const int32_t *labels, int B, int H, int W,
                     int levels, int path_mode, int64_t k, double *out_img, unsigned flags);

/* element types of rbepwt_transcode_ex */
#define RBEPWT_F64 0 /* pixels: float64 (what every other entry point takes) */
#define RBEPWT_F32 1 /* pixels: float32 */
#define RBEPWT_U8 2  /* pixels: uint8 -- what the reference's Image.read yields for grayscale files (rbepwt.py:200-206) */
#define RBEPWT_I32 0 /* labels: int32 */
#define RBEPWT_U16 1 /* labels: uint16 */

/* rbepwt_transcode with narrow element types and the outputs a codec needs.  The same three reference calls
 * (rbepwt.py:298, 441, 307); what changes is what crosses the boundary:
 *   img / labels : pixels as float64, float32 or uint8, labels as int32 or uint16 -- copied as they are and widened on
 *                  the GPU, which is exact (the reference casts to float64 itself before pywt.dwt).  A uint8 image in
 *                  EPWT mode implies RBEPWT_U8_WRAP, as a uint8 array does in the reference (rbepwt.py:1302).
 *   out_img      : decoded images as float64 (bit-identical to rbepwt_transcode), float32 or uint8 (rint of the
 *                  clipped value: what the reference's commented-out line rbepwt.py:312 would store), or NULL.
 *   psnr_out     : HOST float64 [B], psnr(img, decoded) per image (rbepwt.py:361-368), or NULL.
 *   kept_idx/val : HOST int32 / float64 [B][k]: the k surviving coefficients of every image as (flat index ascending,
 *                  value) -- with the label map, the compressed representation -- or both NULL.  Needs 1 <= k < H*W.
 * With out_img and psnr_out both NULL nothing is decoded.  Leaves the same state as rbepwt_transcode.  Synchronous
 * whenever anything is written to host memory. */
int rbepwt_transcode_ex(rbepwt_ctx *ctx, const void *img, int img_dtype, const void *labels, int label_dtype, int B,
                        int H, int W, int levels, int path_mode, int64_t k, void *out_img, int out_dtype,
                        double *psnr_out, int32_t *kept_idx, double *kept_val, unsigned flags);

/* Decoder-side path regeneration (full_decode): build all paths from label maps alone, then
 * decode caller-supplied coefficients (flat layout below, [B][H*W]).          rbepwt.py:106-130
 * Only for the geometric path modes (paths do not depend on pixel values). */
int rbepwt_full_decode(rbepwt_ctx *ctx, const double *coefs, const int32_t *labels, int B, int H,
                       int W, int levels, int path_mode, double *out_img, unsigned flags);

/* psnr(img1, img2) per image: 20 log10(255 / sqrt(mean sq err)), -1 if identical  rbepwt.py:156-162
 * a, b: float64 [B][H][W]; out: HOST float64 [B]. */
int rbepwt_psnr(rbepwt_ctx *ctx, const double *a, const double *b, int B, int64_t n, double *out,
                unsigned flags);

/* Image.nonzero_coefs per image; out: HOST int64 [B]                         rbepwt.py:421-432 */
int rbepwt_nonzero_coefs(rbepwt_ctx *ctx, int64_t *out);

/* Coefficients of image b in the reference's flat layout (Rbepwt.flat_wavelet, rbepwt.py:2195-2204):
 * details[1] | details[2] | ... | details[L] | approximation, H*W doubles.  HOST pointers.
 * set_coefs is what scripts/compute_basis_elements.py:58-81 needs (edit, then decode). */
int rbepwt_get_coefs(rbepwt_ctx *ctx, int b, double *flat);
int rbepwt_set_coefs(rbepwt_ctx *ctx, int b, const double *flat);

/* Region bookkeeping of image b (HOST outputs).
 *   region_count : number of regions R (label classes, in first-appearance order 840-848)
 *   region_offsets: int32 [R+1] offsets at LEVEL 1; level l offsets are ceil(off / 2^(l-1))
 *                  (RegionCollection.reduce keeps the even global positions, 1563-1584)
 *   region_labels: int32 [R] label value of each region */
int rbepwt_region_count(rbepwt_ctx *ctx, int b, int32_t *R);
int rbepwt_region_offsets(rbepwt_ctx *ctx, int b, int32_t *off);
int rbepwt_region_labels(rbepwt_ctx *ctx, int b, int32_t *labels);

/* Paths of image b at `level` (1..L): pixel id (row*W+col) of every point of the level's
 * concatenated signal, regions in order, each region in PATH order (Region.base_points after
 * easy_path, 1338-1342).  level == L+1: the approximation's points in their (incoming) order.
 * level == 0: level-1 INCOMING order (row-major inside each region).  HOST int32 [N >> (level-1)].
 * Level 0 (and rbepwt_get_perm at level 1) re-reads the LABELS given to the encoding call: with
 * RBEPWT_DEVICE_PTRS the caller's label buffer must still be valid and unchanged. */
int rbepwt_get_paths(rbepwt_ctx *ctx, int b, int level, int32_t *pix);
/* Region.permutation of every region at `level` (1..L), concatenated: perm[off_r + t] = index in
 * the region's incoming order of its t-th path point                          rbepwt.py:1285, 1333 */
int rbepwt_get_perm(rbepwt_ctx *ctx, int b, int level, int32_t *perm);

/* Values of image b in the INCOMING order of `level` (1..L): level 1 = pixel values in region order,
 * level l >= 2 = the low-pass output of level l-1 (RegionCollection.values after reduce, 1578).
 * Recomputed from the image given to rbepwt_encode, which must still be valid when device
 * pointers were used.  HOST float64 [N >> (level-1)]. */
int rbepwt_get_level_values(rbepwt_ctx *ctx, int b, int level, double *vals);

/* Per-stage device timings (CUDA events on the context's stream) of the most recent
 * encode / threshold / decode when enabled.  ms[RBEPWT_T_*]; returns the number of stages. */
#define RBEPWT_T_H2D 0
#define RBEPWT_T_REGIONS 1 /* K0: label map -> region records + work queue */
#define RBEPWT_T_PATHS 2   /* K1: easy-path pyramid, regions whose bitmap fits a per-warp shared-memory slot */
#define RBEPWT_T_DWT 3     /* K3: gather + analysis filter bank, all levels */
#define RBEPWT_T_SELECT 4  /* K4: top-k radix select + zeroing */
#define RBEPWT_T_IDWT 5    /* K5: synthesis filter bank + scatter, all levels */
#define RBEPWT_T_D2H 6
#define RBEPWT_T_PATHS_BIG 7 /* K1: regions with a large bounding box (one warp per CTA, whole-image bitmap) */
#define RBEPWT_T_PERM 8      /* K2: positions of the path points in each level's incoming order (levels >= 2) */
#define RBEPWT_T_COUNT 9
int rbepwt_enable_timing(rbepwt_ctx *ctx, int on);
/* Sums the stage events recorded since the previous call (they accumulate across encode /
 * threshold / decode calls), synchronises the stream, returns RBEPWT_T_COUNT. */
int rbepwt_get_timings(rbepwt_ctx *ctx, float *ms, int n);
/* Kernel launches per stage covered by the last rbepwt_get_timings call. */
int rbepwt_get_stage_launches(rbepwt_ctx *ctx, int64_t *launches, int n);
/* Number of kernel launches issued by this context since creation. */
int64_t rbepwt_launch_count(rbepwt_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* RBEPWT_B200_H */
