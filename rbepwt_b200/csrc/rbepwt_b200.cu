// rbepwt_b200: context, launch sequencing and the C ABI declared in include/rbepwt_b200.h.
// There is no CPU implementation in this library: without a CUDA device rbepwt_create fails.
#include "../../include/rbepwt_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "dwt.cuh"
#include "dwt2.cuh"
#include "io.cuh"
#include "paths.cuh"
#include "perm.cuh"
#include "regions.cuh"
#include "segment.hpp"
#include "select.cuh"
#include "walk.cuh"

using namespace rbepwt;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return fail(RBEPWT_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {  // contents not preserved; cudaFree waits for work in flight on the old buffer
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  cudaError_t ensure_slack(size_t bytes) { return bytes <= cap ? cudaSuccess : ensure(bytes + bytes / 4); }
  cudaError_t ensure_keep(size_t bytes, size_t used, cudaStream_t s) {  // first `used` bytes preserved
    if (bytes <= cap) return cudaSuccess;
    size_t ncap = std::max(bytes, cap * 2);
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, ncap);
    if (e != cudaSuccess) return e;
    if (p && used) {
      e = cudaMemcpyAsync(q, p, used, cudaMemcpyDeviceToDevice, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    if (p) cudaFree(p);
    p = q; cap = ncap;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct StageEv { int stage; int launches; cudaEvent_t a, b; };

constexpr int NSLOT = 2;   // path groups in flight
#ifndef RBEPWT_NXSLOT
#define RBEPWT_NXSLOT 2
#endif
constexpr int NXSLOT = RBEPWT_NXSLOT;  // transform sub-batches in flight
constexpr int NSLOTS = NSLOT + NXSLOT;
#ifndef TPR_WIDE_CTAS_PER_SM
#define TPR_WIDE_CTAS_PER_SM 3
#endif
constexpr int TPR_WAVES = 8;  // k1_walk grid = this many waves of resident CTAs (see walk.cuh)

// Workspace + stream of one unit of work in flight.  The batch is cut twice: into PATH GROUPS (label scan,
// region records, path pyramid -- large, the path kernel has a long tail and wants many regions per launch)
// and into TRANSFORM SUB-BATCHES (DWT, top-k, IDWT, copies -- small, so that copies and kernels overlap).
// Units alternate between NSLOT slots per kind, so the issue-bound path kernel of one group overlaps the
// memory-bound transform kernels and the PCIe copies of other units.
struct Slot {
  cudaStream_t s = nullptr;
  cudaStream_t aux = nullptr;     // path slots: the large-bitmap path kernel runs beside the bulk one
  cudaStream_t aux2 = nullptr;    // ... and so does the kernel of the oversized regions (k1_paths_big)
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
  DevBuf VA, VB, Vpix, queue, qhist, qmeta, qbins, chunk_start, chunk_cnt, gscratch, gbm, slot_of;
  // a path group between its two phases (build_regions_and_paths): kernel parameters and launch shapes
  PathParams P;
  int p_nreg = 0, p_big_ctas = 0;
  size_t p_smem_bytes = 0;
  bool p_coop_all = false;
};

struct rbepwt_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;  // the caller-visible stream: every call is ordered on it
  bool own_stream = false;
  int sm_count = 0;
  size_t smem_optin = 0;
  // internal streams: input copies + label scan, output copies, one per slot
  cudaStream_t s_in = nullptr, s_out = nullptr;
  Slot slot[NSLOTS];      // [0, NSLOT): path groups, [NSLOT, NSLOTS): transform sub-batches
  int nslot = NSLOT;      // RBEPWT_OPT_STREAMS
  int opt_sub = 0;        // RBEPWT_OPT_SUBBATCH (0 = auto)
  int opt_group = 0;      // RBEPWT_OPT_PATHGROUP (0 = auto)
  int opt_coop_limit = -1;  // RBEPWT_OPT_COOP_LIMIT (-1 = auto)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaEvent_t> ev_lab, ev_path, ev_img, ev_done;
  int32_t *pin_R = nullptr, *pin_rbase = nullptr, *pin_direct = nullptr;  // pinned staging, capacity pin_cap images
  int *pin_err = nullptr;
  int pin_cap = 0;
  // wavelet
  bool has_wavelet = false;
  int flen = 0;
  DevBuf filt, unit_lut, t2_tab;  // unit_lut: two 9 x 512 tables (euclid, chebyshev); t2_tab: the 5x5 step table (euclid)
  double h_filt[4][FT_MAX] = {};  // host copy (flen <= FT_MAX): passed to the transform kernels by value
  // state of the encoded batch
  bool has_encoding = false, has_paths = false;
  bool is_dwt2 = false;  // the encoding held is the tensor-product baseline (rbepwt_dwt2_encode), not an RBEPWT one
  DevBuf dwt2_tmp;
  int B = 0, H = 0, W = 0, N = 0, logW = 0, levels = 0, mode = 0;
  unsigned enc_flags = 0;
  bool decode_noclip = false;  // RBEPWT_NO_CLIP of the decode in progress
  const int32_t *labels_dev = nullptr;  // ours (labels_own) or the caller's device pointer
  const double *img_dev = nullptr;
  DevBuf labels_own, img_own, out_own, coef_up;
  DevBuf Q, Pm, posmap, coefs;
  // element types and extra outputs of the call in progress (rbepwt_transcode_ex; everything else: f64 / i32, none)
  struct IoSpec {
    int img_dtype = RBEPWT_F64, lab_dtype = RBEPWT_I32, out_dtype = RBEPWT_F64;
    bool src_on_device = false;                  // inputs to stage are device pointers (narrow types only)
    void *out_dst = nullptr; bool out_on_device = false;  // decoded images in out_dtype (NULL: not wanted)
    double *psnr_host = nullptr;                 // [B]
    int32_t *kept_idx_host = nullptr; double *kept_val_host = nullptr; long long kept_k = 0;  // [B][k]
  } io;
  DevBuf img_stage, lab_stage, out_stage, psnr_dev, kept_idx_dev, kept_val_dev;
  DevBuf thr, need_exact;       // per image: pending threshold (select.cuh ThrRec), "k4_select gave up" flag
  bool thr_pending = false;     // some image may have a threshold recorded but not yet written into its coefficients
  DevBuf reg[8];
  int totalR = 0;
  DevBuf img_R, img_rbase, img_labmin, img_direct;
  std::vector<int32_t> h_R, h_rbase;
  // workspace shared by the sub-batches of a chunk / by the getters
  DevBuf tbl, slot_rid, scratch_i32, scratch_i32b, psnr_out, nz_out;
  // timing
  bool timing = false;
  std::vector<StageEv> evs;
  std::vector<cudaEvent_t> ev_pool;
  long long launches = 0;
  long long stage_launches[RBEPWT_T_COUNT] = {};

  RegionArrays regs() const {
    RegionArrays r;
    r.label = reg[0].as<int32_t>(); r.first = reg[1].as<int32_t>(); r.size = reg[2].as<int32_t>();
    r.off = reg[3].as<int32_t>(); r.rmax = reg[4].as<int32_t>(); r.cmin = reg[5].as<int32_t>();
    r.cmax = reg[6].as<int32_t>(); r.img = reg[7].as<int32_t>();
    return r;
  }
};

namespace {

enum { DO_PATHS = 1, DO_DWT = 2, DO_THRESH = 4, DO_DECODE = 8 };

void set_taps(const rbepwt_ctx *c, DwtParams &D, bool inverse) {
  for (int i = 0; i < FT_MAX; i++) { D.tap_lo[i] = c->h_filt[inverse ? 2 : 0][i]; D.tap_hi[i] = c->h_filt[inverse ? 3 : 1][i]; }
}

void launch_dwt_level(const DwtParams &D, dim3 grid, cudaStream_t s) {
  switch (D.flen) {
    case 2: k3_dwt_level<2><<<grid, DWT_THREADS, 0, s>>>(D); break;
    case 4: k3_dwt_level<4><<<grid, DWT_THREADS, 0, s>>>(D); break;
    case 6: k3_dwt_level<6><<<grid, DWT_THREADS, 0, s>>>(D); break;
    case 8: k3_dwt_level<8><<<grid, DWT_THREADS, 0, s>>>(D); break;
    case 10: k3_dwt_level<10><<<grid, DWT_THREADS, 0, s>>>(D); break;
    default: k3_dwt_level<0><<<grid, DWT_THREADS, 0, s>>>(D); break;
  }
}

void launch_idwt_level(const DwtParams &D, dim3 grid, cudaStream_t s) {
  switch (D.flen) {
    case 2: k5_idwt_level<2><<<grid, DWT_THREADS, 0, s>>>(D); break;
    case 4: k5_idwt_level<4><<<grid, DWT_THREADS, 0, s>>>(D); break;
    case 6: k5_idwt_level<6><<<grid, DWT_THREADS, 0, s>>>(D); break;
    case 8: k5_idwt_level<8><<<grid, DWT_THREADS, 0, s>>>(D); break;
    case 10: k5_idwt_level<10><<<grid, DWT_THREADS, 0, s>>>(D); break;
    default: k5_idwt_level<0><<<grid, DWT_THREADS, 0, s>>>(D); break;
  }
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

cudaEvent_t get_event(rbepwt_ctx *c) {
  if (!c->ev_pool.empty()) { cudaEvent_t e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}

struct StageTimer {  // records a pair of events around a stage, on the stream the stage runs on
  rbepwt_ctx *c; cudaStream_t s; StageEv ev; bool on; long long l0;
  StageTimer(rbepwt_ctx *c_, int stage, cudaStream_t s_) : c(c_), s(s_), on(c_->timing), l0(c_->launches) {
    if (on) { ev.stage = stage; ev.a = get_event(c); ev.b = get_event(c); cudaEventRecord(ev.a, s); }
  }
  ~StageTimer() {
    if (on) { cudaEventRecord(ev.b, s); ev.launches = (int)(c->launches - l0); c->evs.push_back(ev); }
  }
};

void clear_events(rbepwt_ctx *c) {
  for (auto &e : c->evs) { c->ev_pool.push_back(e.a); c->ev_pool.push_back(e.b); }
  c->evs.clear();
}

int ilog2(int x) { int l = 0; while ((1 << l) < x) l++; return l; }

// images per chunk: the label table + slot ids (24 N bytes/image) live for a whole chunk
int chunk_images(const rbepwt_ctx *c, int B, int N) {
  const size_t per = (size_t)N * 24;
  size_t budget = (size_t)6 << 30;
  int m = (int)std::max<size_t>(1, budget / per);
  return std::min(std::min(m, 1024), B);
}

// images per transform sub-batch.  Host inputs: about 2^24 pixels -- enough CTAs per launch to fill the GPU, few enough
// that copies and kernels of neighbouring sub-batches overlap.  Device-resident inputs: nothing to overlap with but the
// next path group, and every launch of a small level costs its few microseconds whatever the size: about 2^26 pixels
// (measured, 512 images of 512^2: forward + select + inverse 4.2 ms in sub-batches of 64, 2.6 ms in sub-batches of 256).
int sub_images(const rbepwt_ctx *c, int nb, int N, bool host_inputs) {
  if (c->opt_sub > 0) return std::min(c->opt_sub, nb);
  const int m = (int)std::max<long long>(1, (1ll << (host_inputs ? 24 : 26)) / N);
  return std::min(m, nb);
}

// images per path group: about 2^26 pixels (256 images of 512^2), a multiple of the sub-batch
int group_images(const rbepwt_ctx *c, int nb, int N, int Bs) {
  long long m = c->opt_group > 0 ? c->opt_group : std::max<long long>(1, (1ll << 26) / N);
  m = std::max<long long>(Bs, m / Bs * Bs);
  return (int)std::min<long long>(m, (nb + Bs - 1) / Bs * Bs);
}

int sync_internal(rbepwt_ctx *c) {
  CK(cudaStreamSynchronize(c->s_in));
  CK(cudaStreamSynchronize(c->s_out));
  for (int i = 0; i < NSLOTS; i++) CK(cudaStreamSynchronize(c->slot[i].s));
  return RBEPWT_OK;
}

// every internal stream starts after what the caller already queued on ctx->stream ...
int fork_streams(rbepwt_ctx *c) {
  CK(cudaEventRecord(c->ev_fork, c->stream));
  CK(cudaStreamWaitEvent(c->s_in, c->ev_fork, 0));
  CK(cudaStreamWaitEvent(c->s_out, c->ev_fork, 0));
  for (int i = 0; i < NSLOTS; i++) CK(cudaStreamWaitEvent(c->slot[i].s, c->ev_fork, 0));
  return RBEPWT_OK;
}

// ... and ctx->stream continues after all of them
int join_streams(rbepwt_ctx *c) {
  cudaStream_t all[NSLOTS + 2] = {c->s_in, c->s_out};
  for (int i = 0; i < NSLOTS; i++) all[2 + i] = c->slot[i].s;
  for (cudaStream_t s : all) {
    CK(cudaEventRecord(c->ev_join, s));
    CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  }
  return RBEPWT_OK;
}

int check_path_error(rbepwt_ctx *c) {
  CK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < NSLOTS; i++) {
    if (!c->slot[i].qmeta.p) continue;
    CK(cudaMemcpyAsync(c->pin_err, c->slot[i].qmeta.as<int>() + QM_ERR, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (*c->pin_err) return fail(RBEPWT_E_CUDA, "path kernel found no unvisited point (corrupt region state)");
  }
  return RBEPWT_OK;
}

int validate_shape(int B, int H, int W, int levels, int path_mode) {
  if (B < 1 || H < 1 || W < 1) return fail(RBEPWT_E_ARG, "B, H, W must be positive");
  const long long n = (long long)H * W;
  if (n & (n - 1)) return fail(RBEPWT_E_NOT_POW2, "Image size must be a power of 2");
  // per-image tables hold 2 * H*W entries indexed with int
  if (n > (1ll << 29) || W > 32768 || H > 32768) return fail(RBEPWT_E_ARG, "image too large (H*W <= 2^29, sides <= 32768)");
  if (levels < 1 || levels > 30 || (1ll << levels) > n)
    return fail(RBEPWT_E_LEVELS, "2^levels must be smaller or equal to the number of pixels in the image");
  if (path_mode < 0 || path_mode > RBEPWT_PATH_GRAD_CHEB) return fail(RBEPWT_E_ARG, "unknown path mode %d", path_mode);
  if ((path_mode == RBEPWT_PATH_GRAD || path_mode == RBEPWT_PATH_GRAD_CHEB) && (H < 2 || W < 2))
    return fail(RBEPWT_E_ARG, "Shape of array too small to calculate a numerical gradient, at least 2 elements are required.");
  return RBEPWT_OK;
}

// Region arrays are global over the batch and grow with the region count, which is only known sub-batch
// by sub-batch; growing them means waiting for everything in flight (rare: the first guess is generous).
int grow_regs(rbepwt_ctx *c, size_t need_entries, size_t used_entries) {
  if (need_entries * 4 <= c->reg[0].cap) return RBEPWT_OK;
  int rc = sync_internal(c);
  if (rc) return rc;
  for (int i = 0; i < 8; i++) CK(c->reg[i].ensure_keep(need_entries * 4, used_entries * 4, c->slot[0].s));
  return RBEPWT_OK;
}

int ensure_slot_workspace(rbepwt_ctx *c, Slot &sl, int nb) {  // nb = 0: no value planes (path slots)
  const size_t N = c->N;
  CK(sl.VA.ensure((size_t)nb * (N / 2) * 8));  // dense planes: x^(l+1) has at most N/2 entries
  CK(sl.VB.ensure((size_t)nb * (N / 2) * 8));
  if (c->mode == RBEPWT_PATH_EPWT) CK(sl.Vpix.ensure((size_t)nb * N * 8));
  if (!sl.qhist.p) {
    CK(sl.qhist.ensure(Q_BINS * 4));
    CK(cudaMemsetAsync(sl.qhist.p, 0, Q_BINS * 4, c->stream));  // before the fork: ordered ahead of every stream
  }
  if (!sl.qmeta.p) {
    CK(sl.qmeta.ensure(QM_SIZE * 4));
    CK(cudaMemsetAsync(sl.qmeta.p, 0, QM_SIZE * 4, c->stream));
  }
  if (!sl.qbins.p) CK(sl.qbins.ensure(3 * Q_BINS * 4));
  return RBEPWT_OK;
}

// Input stream, path group [a, a+nb): (host mode) copy its labels in, then scan them -- label table and
// region count per image; the counts are read back for the host's bookkeeping.
size_t dtype_size(int dt_img_or_out) { return dt_img_or_out == RBEPWT_F64 ? 8 : dt_img_or_out == RBEPWT_F32 ? 4 : 1; }
size_t label_size(int dt) { return dt == RBEPWT_I32 ? 4 : 2; }

int stage_labels(rbepwt_ctx *c, int chunk0, int a, int nb, const void *lab_src, cudaEvent_t ready) {
  const size_t N = c->N;
  cudaStream_t s = c->s_in;
  if (c->mode != RBEPWT_PATH_EPWT) {
    if (lab_src) {
      StageTimer t(c, RBEPWT_T_H2D, s);
      const size_t esz = label_size(c->io.lab_dtype), cnt = (size_t)nb * N;
      const char *src = static_cast<const char *>(lab_src) + (size_t)a * N * esz;
      int32_t *dst = c->labels_own.as<int32_t>() + (size_t)a * N;
      if (c->io.lab_dtype == RBEPWT_I32) {
        CK(cudaMemcpyAsync(dst, src, cnt * 4, cudaMemcpyHostToDevice, s));
      } else {  // narrow labels: 2 bytes per pixel over PCIe, widened on the device
        const uint16_t *dev_src = reinterpret_cast<const uint16_t *>(src);
        if (!c->io.src_on_device) {
          uint16_t *stage = c->lab_stage.as<uint16_t>() + (size_t)a * N;
          CK(cudaMemcpyAsync(stage, src, cnt * esz, cudaMemcpyHostToDevice, s));
          dev_src = stage;
        }
        k_widen_labels<<<(unsigned)std::min<size_t>((cnt + 1023) / 1024, 4096), 256, 0, s>>>(dev_src, dst, cnt);
        c->launches++;
      }
    }
    StageTimer t(c, RBEPWT_T_REGIONS, s);
    const int T = 2 * c->N;
    k0_count<<<nb * K0_CLUSTER, K0_THREADS, 0, s>>>(c->labels_dev, a, c->N, c->tbl.as<unsigned long long>() + (size_t)(a - chunk0) * T, T,
                                       c->img_R.as<int32_t>(), c->img_labmin.as<int32_t>(), c->img_direct.as<int32_t>());
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c->pin_R + a, c->img_R.as<int32_t>() + a, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(c->pin_direct + a, c->img_direct.as<int32_t>() + a, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaEventRecord(ready, s));
  return RBEPWT_OK;
}

// Input stream, transform sub-batch [a, a+nb): (host mode) copy its images in.
int stage_images(rbepwt_ctx *c, int a, int nb, const void *img_src, cudaEvent_t ready) {
  const size_t N = c->N;
  cudaStream_t s = c->s_in;
  StageTimer t(c, RBEPWT_T_H2D, s);
  const size_t esz = dtype_size(c->io.img_dtype), cnt = (size_t)nb * N;
  const char *src = static_cast<const char *>(img_src) + (size_t)a * N * esz;
  double *dst = c->img_own.as<double>() + (size_t)a * N;
  if (c->io.img_dtype == RBEPWT_F64) {
    CK(cudaMemcpyAsync(dst, src, cnt * 8, cudaMemcpyHostToDevice, s));
  } else {  // uint8 / float32 pixels: 1 or 4 bytes per pixel over PCIe, widened (exactly) on the device
    const void *dev_src = src;
    if (!c->io.src_on_device) {
      char *stage = c->img_stage.as<char>() + (size_t)a * N * esz;
      CK(cudaMemcpyAsync(stage, src, cnt * esz, cudaMemcpyHostToDevice, s));
      dev_src = stage;
    }
    const unsigned grid = (unsigned)std::min<size_t>((cnt + 1023) / 1024, 4096);
    if (c->io.img_dtype == RBEPWT_U8) k_widen_pixels<uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t *>(dev_src), dst, cnt);
    else k_widen_pixels<float><<<grid, 256, 0, s>>>(static_cast<const float *>(dev_src), dst, cnt);
    c->launches++;
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(ready, s));
  return RBEPWT_OK;
}

// K0 (region records, work queue) + K1 (path pyramid) of one path group on stream s (workspace: sl), in two phases:
//   PHASE_REGIONS: label scan results -> region records, the size-sorted queue;
//   PHASE_WALK   : the chunk arena images, the path kernels and the positions (k2_perm).
// run_pipeline enqueues the first phase of EVERY group of a chunk before any second phase: the first phase is latency /
// atomics bound and its groups overlap well with each other, but its 1024-thread CTAs cannot get onto an SM while walk
// CTAs fill it -- enqueued after walk(g) it waited ~3 ms for the walk to drain (tools/pipe_stages.py) and delayed
// walk(g+1) by its whole length.
// EPWT: one region per image; its paths depend on the level's values and are built in transform_sub().
enum { PHASE_REGIONS = 1, PHASE_WALK = 2 };
int build_regions_and_paths(rbepwt_ctx *c, Slot &sl, cudaStream_t s, int chunk0, int a, int nb, cudaEvent_t ready, int phases) {
  const int N = c->N, T = 2 * N;
  if (c->mode == RBEPWT_PATH_EPWT) {
    if (!(phases & PHASE_REGIONS)) return RBEPWT_OK;
    int rc = grow_regs(c, (size_t)c->B, (size_t)a);
    if (rc) return rc;
    CK(cudaStreamWaitEvent(s, ready, 0));
    StageTimer t(c, RBEPWT_T_REGIONS, s);
    k0_single_region<<<(nb + 127) / 128, 128, 0, s>>>(a, nb, c->H, c->W, c->regs(), c->img_R.as<int32_t>(),
                                                     c->img_rbase.as<int32_t>());
    c->launches++;
    for (int i = 0; i < nb; i++) { c->h_R[a + i] = 1; c->h_rbase[a + i] = a + i; }
    c->totalR = a + nb;
    CK(cudaGetLastError());
    return RBEPWT_OK;
  }
  if (phases & PHASE_REGIONS) {
  CK(cudaEventSynchronize(ready));  // the host needs the sub-batch's region counts
  const int g0 = c->totalR;
  bool any_general = false;  // an image the shared-memory sweep cannot take (sparse label values, > 4096 regions, ...)
  for (int i = 0; i < nb; i++) any_general |= !k0_fast_eligible(c->pin_direct[a + i], c->pin_R[a + i], c->logW);
  for (int i = 0; i < nb; i++) {
    c->h_R[a + i] = c->pin_R[a + i];
    c->h_rbase[a + i] = c->pin_rbase[a + i] = c->totalR;
    c->totalR += c->h_R[a + i];
  }
  const int nreg = c->totalR - g0;
  bool coop_all = false;  // every region of the group walked by a whole warp (k1_coop_all)
  int rc = grow_regs(c, std::max<size_t>((size_t)c->totalR, (size_t)c->B * 2048 + 4096), (size_t)g0);
  if (rc) return rc;
  CK(sl.queue.ensure_slack((size_t)nreg * 4));
  CK(sl.chunk_start.ensure_slack(((size_t)nreg + Q_NCLS) * 4));
  CK(sl.chunk_cnt.ensure_slack(((size_t)nreg + Q_NCLS) * 4));
  // chunk arena images: room for the usual ~nreg/32 chunks and then some; more chunks than that (many large
  // bitmaps) build theirs inside the path kernel
  const size_t gbm_chunks = (size_t)nreg / 16 + 64;
  CK(sl.gbm.ensure_slack(gbm_chunks * TPR_ARENA_WORDS * 4));
  CK(sl.slot_of.ensure_slack((size_t)nreg * 4));
  CK(cudaStreamWaitEvent(s, ready, 0));
  {
    StageTimer t(c, RBEPWT_T_REGIONS, s);
    CK(cudaMemcpyAsync(c->img_rbase.as<int32_t>() + a, c->pin_rbase + a, (size_t)nb * 4, cudaMemcpyHostToDevice, s));
    CK(cudaFuncSetAttribute(k0_regions_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K0F_SMEM));
    k0_regions_fast<<<nb, K0_THREADS, K0F_SMEM, s>>>(c->labels_dev, a, N, c->logW,
                                                    c->tbl.as<unsigned long long>() + (size_t)(a - chunk0) * T, T,
                                                    c->img_R.as<int32_t>(), c->img_labmin.as<int32_t>(),
                                                    c->img_direct.as<int32_t>(), c->img_rbase.as<int32_t>(), c->regs());
    if (any_general)  // (a cluster launch of 1024-thread CTAs waits for eight free SMs: not worth an empty launch)
      k0_regions<<<nb * K0_CLUSTER, K0_THREADS, 0, s>>>(c->labels_dev, a, N, c->logW,
                                           c->tbl.as<unsigned long long>() + (size_t)(a - chunk0) * T,
                                           c->slot_rid.as<int32_t>() + (size_t)(a - chunk0) * T, T, c->img_R.as<int32_t>(),
                                           c->img_labmin.as<int32_t>(), c->img_direct.as<int32_t>(),
                                           c->img_rbase.as<int32_t>(), c->regs());
    const int qb = std::max(1, std::min((nreg + 255) / 256, c->sm_count * 8));
    // few regions in flight (single images, small batches): every region gets its own warp -- latency, not throughput;
    // gradpath: always.  Otherwise (euclid) the largest regions of the group do, kq_scan picks the threshold.
    // RBEPWT_OPT_COOP_LIMIT: how many (-1 = auto: COOP_PER_SM per SM; 0 = none, and no small-group rule either)
    const bool grad = c->mode == RBEPWT_PATH_GRAD || c->mode == RBEPWT_PATH_GRAD_CHEB;
    coop_all = grad || (c->mode == RBEPWT_PATH_EUCLID && nreg <= TPR_COOP_ALL_BELOW && c->opt_coop_limit != 0);
    const int coop_min = coop_all ? 1 : TPR_COOP_MIN;
    const int coop_limit = c->mode != RBEPWT_PATH_EUCLID ? 0 : (c->opt_coop_limit >= 0 ? c->opt_coop_limit : c->sm_count * COOP_PER_SM);
    kq_hist<<<qb, 256, 0, s>>>(c->regs(), g0, nreg, c->logW, coop_all ? 1 : INT32_MAX, sl.qhist.as<int>());
    kq_scan<<<1, 32, 0, s>>>(sl.qhist.as<int>(), sl.qmeta.as<int>(), sl.qbins.as<int>(), nreg, coop_min, coop_limit);
    kq_scatter<<<qb, 256, 0, s>>>(c->regs(), g0, nreg, c->logW, sl.qmeta.as<int>(), sl.queue.as<int32_t>());
    kq_chunks<<<Q_NCLS - 1, 1024, 0, s>>>(sl.qbins.as<int>(), sl.chunk_start.as<int32_t>(), sl.chunk_cnt.as<int32_t>());
    c->launches += 6;
    CK(cudaGetLastError());
  }
  PathParams P;
  P.labels = c->labels_dev;
  P.img = c->img_dev;
  P.H = c->H; P.W = c->W; P.logW = c->logW; P.N = N;
  P.levels = (c->enc_flags & RBEPWT_PATHS_FIRST_LEVEL) ? 1 : c->levels;
  P.reg = c->regs();
  P.queue = sl.queue.as<int32_t>();
  P.chunk_start = sl.chunk_start.as<int32_t>();
  P.chunk_cnt = sl.chunk_cnt.as<int32_t>();
  P.qmeta = sl.qmeta.as<int>();
  P.unit_lut = c->unit_lut.as<uint8_t>() + (c->mode == RBEPWT_PATH_CHEB ? TPR_LUT_ROWS * TPR_LUT_COLS : 0);
  P.t2_tab = c->t2_tab.as<uint8_t>();
  P.coop = c->mode != RBEPWT_PATH_CHEB ? 1 : 0;
  P.gbm = sl.gbm.as<uint32_t>();
  P.slot_of = sl.slot_of.as<int32_t>();
  P.g0 = g0; P.nreg = nreg;
  P.gbm_chunks = (int)gbm_chunks;
  P.Q = c->Q.as<int32_t>();
  P.Pm = c->Pm.as<int32_t>();
  P.posmap = c->posmap.as<int32_t>();
  // big-region kernel: whole-image bitmap in dynamic shared memory when it fits
  const size_t img_words = (size_t)c->H * ((c->W + 31) / 32);
  const size_t smem_cap = std::min<size_t>(c->smem_optin, (size_t)200 * 1024);
  const size_t smem_bytes = std::min(img_words * 4, smem_cap);
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, ((size_t)220 * 1024) / (smem_bytes + 1024)));
  const int big_ctas = c->sm_count * per_sm;
  P.big_smem_words = (int)(smem_bytes / 4);
  P.gscratch = nullptr; P.gscratch_words = 0;
  if (img_words * 4 > smem_bytes) {
    CK(sl.gscratch.ensure(img_words * 4 * (size_t)big_ctas));
    P.gscratch = sl.gscratch.as<uint32_t>();
    P.gscratch_words = img_words;
  }
  sl.P = P; sl.p_nreg = nreg; sl.p_big_ctas = big_ctas; sl.p_smem_bytes = smem_bytes; sl.p_coop_all = coop_all;
  }  // PHASE_REGIONS
  if (!(phases & PHASE_WALK)) return RBEPWT_OK;
  sl.P.reg = c->regs();  // another group's first phase may have grown (moved) the region arrays since this group's
  const PathParams &P = sl.P;
  const int nreg = sl.p_nreg, big_ctas = sl.p_big_ctas;
  const size_t smem_bytes = sl.p_smem_bytes;
  const bool coop_all = sl.p_coop_all;
  const bool grad_mode = c->mode == RBEPWT_PATH_GRAD || c->mode == RBEPWT_PATH_GRAD_CHEB;
  if (coop_all) {
    // regions whose bitmap exceeds a shared-memory arena in k1_paths_big (own stream), all the others one warp each
    StageTimer t(c, RBEPWT_T_PATHS, s);
    cudaStream_t sbig = c->nslot == 1 ? s : sl.aux2;
    CK(cudaEventRecord(sl.ev_a, s));
    CK(cudaStreamWaitEvent(sl.aux2, sl.ev_a, 0));
    if (c->mode == RBEPWT_PATH_GRAD) {
      CK(cudaFuncSetAttribute(k1_paths_big<MODE_GRAD_EUCLID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      k1_paths_big<MODE_GRAD_EUCLID><<<big_ctas, 32, smem_bytes, sbig>>>(P);
      k1_coop_all<MODE_GRAD_EUCLID><<<c->sm_count * 4, COOP_WARPS * 32, 0, s>>>(P);
    } else if (c->mode == RBEPWT_PATH_GRAD_CHEB) {
      CK(cudaFuncSetAttribute(k1_paths_big<MODE_GRAD_CHEB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      k1_paths_big<MODE_GRAD_CHEB><<<big_ctas, 32, smem_bytes, sbig>>>(P);
      k1_coop_all<MODE_GRAD_CHEB><<<c->sm_count * 4, COOP_WARPS * 32, 0, s>>>(P);
    } else {
      CK(cudaFuncSetAttribute(k1_paths_big<MODE_EUCLID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      k1_paths_big<MODE_EUCLID><<<big_ctas, 32, smem_bytes, sbig>>>(P);
      k1_coop_all<MODE_EUCLID><<<c->sm_count * 4, COOP_WARPS * 32, 0, s>>>(P);
    }
    CK(cudaEventRecord(sl.ev_c, sbig));
    CK(cudaStreamWaitEvent(s, sl.ev_c, 0));
    c->launches += 2;
  } else {
  int tpr_per_sm = 1;  // grid = TPR_WAVES waves of resident CTAs
  if (c->mode == RBEPWT_PATH_EUCLID)
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tpr_per_sm, k1_walk<MODE_EUCLID, false>, WK_WARPS * 32, wk_arena_bytes(false)));
  else
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tpr_per_sm, k1_walk<MODE_CHEB, false>, WK_WARPS * 32, wk_arena_bytes(false)));
  const int small_ctas = c->sm_count * std::max(tpr_per_sm, 1) * TPR_WAVES;
  // the oversized regions (one warp each, the longest chains of all) run beside everything else, on their own stream
  // (RBEPWT_OPT_STREAMS = 1, per-kernel timing: in line, so that its time is its own and not the wait for an SM)
  cudaStream_t sbig = c->nslot == 1 ? s : sl.aux2;
  CK(cudaEventRecord(sl.ev_a, s));
  CK(cudaStreamWaitEvent(sl.aux2, sl.ev_a, 0));
  {
    StageTimer tb(c, RBEPWT_T_PATHS_BIG, sbig);
    if (c->mode == RBEPWT_PATH_EUCLID) {
      CK(cudaFuncSetAttribute(k1_paths_big<MODE_EUCLID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      k1_paths_big<MODE_EUCLID><<<big_ctas, 32, smem_bytes, sbig>>>(P);
    } else {
      CK(cudaFuncSetAttribute(k1_paths_big<MODE_CHEB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      k1_paths_big<MODE_CHEB><<<big_ctas, 32, smem_bytes, sbig>>>(P);
    }
    c->launches++;
  }
  if (c->mode == RBEPWT_PATH_EUCLID) {  // the group's largest regions, a warp each (class 1), beside the lane walkers
    k1_coop_all<MODE_EUCLID><<<c->sm_count * 4, COOP_WARPS * 32, 0, sbig>>>(P);
    c->launches++;
  }
  CK(cudaEventRecord(sl.ev_c, sbig));
  {
    StageTimer t(c, RBEPWT_T_PATHS, s);
    // the large-bitmap chunks (few, the longest chains) run on the slot's auxiliary stream, beside the bulk; they
    // build their own bitmaps, so they start right away, under the bulk kernel's bitmap builder
    CK(cudaStreamWaitEvent(sl.aux, sl.ev_a, 0));
    if (c->mode == RBEPWT_PATH_EUCLID)
      k1_walk<MODE_EUCLID, true><<<c->sm_count * TPR_WIDE_CTAS_PER_SM, WK_WIDE_WARPS * 32, wk_arena_bytes(true), sl.aux>>>(P);
    else
      k1_walk<MODE_CHEB, true><<<c->sm_count * TPR_WIDE_CTAS_PER_SM, WK_WIDE_WARPS * 32, wk_arena_bytes(true), sl.aux>>>(P);
    CK(cudaMemsetAsync(P.slot_of, 0xff, (size_t)nreg * 4, s));
    kq_slots<<<c->sm_count * 4, 256, 0, s>>>(P);
    k1_bitmaps<<<c->sm_count * 8, 256, 0, s>>>(P);
    c->launches += 2;
    if (c->mode == RBEPWT_PATH_EUCLID)
      k1_walk<MODE_EUCLID, false><<<small_ctas, WK_WARPS * 32, wk_arena_bytes(false), s>>>(P);
    else
      k1_walk<MODE_CHEB, false><<<small_ctas, WK_WARPS * 32, wk_arena_bytes(false), s>>>(P);
    CK(cudaEventRecord(sl.ev_b, sl.aux));
    CK(cudaStreamWaitEvent(s, sl.ev_b, 0));
    c->launches += 2;
  }
  CK(cudaStreamWaitEvent(s, sl.ev_c, 0));
  }
  {
    StageTimer t(c, RBEPWT_T_PERM, s);
    if ((c->enc_flags & RBEPWT_PATHS_FIRST_LEVEL) && c->levels > 1) {  // identity permutations at the levels >= 2
      k_same_paths<<<nb, 1024, 0, s>>>(c->Q.as<int32_t>() + (size_t)a * 2 * N, c->Pm.as<int32_t>() + (size_t)a * 2 * N, N,
                                       c->levels);
      c->launches++;
    } else if (c->levels > 1) {  // positions in the incoming order of every level >= 2, from the paths
      const int grid = std::max(1, std::min((nreg + K2_WARPS - 1) / K2_WARPS, c->sm_count * 16));
      k2_perm<<<grid, K2_WARPS * 32, 0, s>>>(P, nreg, (c->mode == RBEPWT_PATH_EUCLID || grad_mode) ? 1 : 0);
      c->launches++;
    }
  }
  CK(cudaGetLastError());
  return RBEPWT_OK;
}

// K3 for every level of the sub-batch (EPWT: K1 level kernel before each level's DWT).
int transform_sub(rbepwt_ctx *c, Slot &sl, cudaStream_t s, int a, int nb) {
  const int N = c->N;
  DwtParams D;
  D.Q = c->Q.as<int32_t>() + (size_t)a * 2 * N;
  D.Pm = c->Pm.as<int32_t>() + (size_t)a * 2 * N;
  D.coefs = c->coefs.as<double>() + (size_t)a * N;
  D.filt = c->filt.as<double>();
  D.out_img = nullptr;
  D.flen = c->flen; D.N = N; D.levels = c->levels; D.clip = 1; D.thr = nullptr;
  set_taps(c, D, false);
  double *V[2] = {sl.VA.as<double>(), sl.VB.as<double>()};
  EpwtParams E;
  size_t epwt_smem = 0;
  if (c->mode == RBEPWT_PATH_EPWT) {
    const size_t img_words = (size_t)c->H * ((c->W + 31) / 32);
    const size_t smem_cap = std::min<size_t>(c->smem_optin, (size_t)200 * 1024);
    epwt_smem = std::min(img_words * 4, smem_cap);
    E.H = c->H; E.W = c->W; E.logW = c->logW; E.N = N; E.img0 = 0;
    E.Q = const_cast<int32_t *>(D.Q);
    E.Pm = c->Pm.as<int32_t>() + (size_t)a * 2 * N;
    E.posmap = c->posmap.as<int32_t>() + (size_t)a * N;
    E.smem_words = (int)(epwt_smem / 4);
    E.gscratch = nullptr; E.gscratch_words = 0;
    if (img_words * 4 > epwt_smem) {
      CK(sl.gscratch.ensure(img_words * 4 * (size_t)nb));
      E.gscratch = sl.gscratch.as<uint32_t>();
      E.gscratch_words = img_words;
    }
    E.u8wrap = (c->enc_flags & RBEPWT_U8_WRAP) ? 1 : 0;
    E.qmeta = sl.qmeta.as<int>();
    CK(cudaFuncSetAttribute(k1_epwt_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)epwt_smem));
  }
  D.vin = c->img_dev + (size_t)a * N; D.vin_stride = N;
  D.plane[0] = V[0]; D.plane[1] = V[1];
  for (int lev = 1; lev <= c->levels; lev++) {
    D.lev = lev;
    const int n = N >> (lev - 1);
    const bool same_paths = (c->enc_flags & RBEPWT_PATHS_FIRST_LEVEL) != 0;
    if (c->mode == RBEPWT_PATH_EPWT && !(same_paths && lev > 1)) {
      StageTimer t(c, RBEPWT_T_PATHS, s);
      E.lev = lev;
      E.vals = lev == 1 ? D.vin : sl.Vpix.as<double>();  // values BY PIXEL: the image, then cA laid out by pixel
      k1_epwt_level<<<nb, 32, epwt_smem, s>>>(E);
      c->launches++;
      if (same_paths && c->levels > 1) {
        k_same_paths<<<nb, 1024, 0, s>>>(c->Q.as<int32_t>() + (size_t)a * 2 * N, c->Pm.as<int32_t>() + (size_t)a * 2 * N, N,
                                         c->levels);
        c->launches++;
      }
    } else if (n <= TAIL_MAX_POINTS && (c->mode != RBEPWT_PATH_EPWT || same_paths)) {  // all remaining levels in one launch, one CTA per image
      StageTimer t(c, RBEPWT_T_DWT, s);
      k3_dwt_tail<<<nb, DWT_THREADS, 0, s>>>(D);
      c->launches++;
      break;
    }
    {
      StageTimer t(c, RBEPWT_T_DWT, s);
      dim3 grid(((n >> 1) + FWD_TILE - 1) / FWD_TILE, nb);
      launch_dwt_level(D, grid, s);
      c->launches++;
      if (c->mode == RBEPWT_PATH_EPWT && lev < c->levels && !same_paths) {
        k_plane_to_pixels<<<dim3(((n >> 1) + 255) / 256, nb), 256, 0, s>>>(V[lev & 1], D.Q, N, lev, sl.Vpix.as<double>());
        c->launches++;
      }
    }
  }
  CK(cudaGetLastError());
  return RBEPWT_OK;
}

// K4 for images [a, a+nb): one-pass select that records the threshold (applied by the inverse transform while it
// loads the coefficients, or written out by materialise_thresholds), then the exact in-place kernel for the images
// the first one flagged (none, for continuous data).  The images must have no threshold pending.
int threshold_sub(rbepwt_ctx *c, cudaStream_t s, int a, int nb, long long k) {
  StageTimer t(c, RBEPWT_T_SELECT, s);
  if (k <= 0 || k >= (long long)c->N) return RBEPWT_OK;  // keeps everything (reference quirk)
  CK(cudaFuncSetAttribute(k4_threshold, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEL_SMEM));
  CK(cudaFuncSetAttribute(k4_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SEL2_SMEM));
  double *coefs = c->coefs.as<double>() + (size_t)a * c->N;
  int *need = c->need_exact.as<int>() + a;
  CK(cudaMemsetAsync(need, 0, (size_t)nb * sizeof(int), s));
  k4_select<<<nb, SEL_THREADS, SEL2_SMEM, s>>>(coefs, c->N, k, c->thr.as<ThrRec>() + a, need);
  k4_threshold<<<nb * SEL_CLUSTER, SEL_THREADS, SEL_SMEM, s>>>(coefs, c->N, k, need);
  c->launches += 2;
  c->thr_pending = true;
  CK(cudaGetLastError());
  return RBEPWT_OK;
}

// Pending thresholds written into the coefficients (whole batch, on the context's stream): before anything looks at or
// replaces the coefficients, and before they are thresholded again.
int materialise_thresholds(rbepwt_ctx *c) {
  if (!c->thr_pending) return RBEPWT_OK;
  const int gx = std::max(1, std::min(c->N / 1024, 64));
  k_apply_threshold<<<dim3(gx, c->B), 256, 0, c->stream>>>(c->coefs.as<double>(), c->N, c->thr.as<ThrRec>());
  k_clear_thresholds<<<(c->B + 255) / 256, 256, 0, c->stream>>>(c->thr.as<ThrRec>(), c->B);
  c->launches += 2;
  c->thr_pending = false;
  CK(cudaGetLastError());
  return RBEPWT_OK;
}

int decode_sub(rbepwt_ctx *c, Slot &sl, cudaStream_t s, int a, int nb, double *out_dev) {
  const int N = c->N;
  double *V[2] = {sl.VA.as<double>(), sl.VB.as<double>()};
  StageTimer t(c, RBEPWT_T_IDWT, s);
  DwtParams D;
  D.Q = c->Q.as<int32_t>() + (size_t)a * 2 * N;
  D.Pm = c->Pm.as<int32_t>() + (size_t)a * 2 * N;
  D.coefs = c->coefs.as<double>() + (size_t)a * N;
  D.filt = c->filt.as<double>();
  D.out_img = out_dev + (size_t)a * N;
  D.flen = c->flen; D.N = N; D.levels = c->levels; D.clip = c->decode_noclip ? 0 : 1;
  D.thr = reinterpret_cast<const unsigned long long *>(c->thr.as<ThrRec>() + a);
  set_taps(c, D, true);
  D.vin = nullptr; D.vin_stride = N;
  D.plane[0] = V[0]; D.plane[1] = V[1];
  int top = c->levels;  // deepest level still to be inverted by a per-level launch
  int first_tail = 1;
  while (first_tail <= c->levels && (N >> (first_tail - 1)) > TAIL_MAX_POINTS) first_tail++;
  if (first_tail <= c->levels) {  // levels L .. first_tail in one launch, one CTA per image
    D.lev = first_tail;
    k5_idwt_tail<<<nb, DWT_THREADS, 0, s>>>(D);
    c->launches++;
    top = first_tail - 1;
  }
  for (int lev = top; lev >= 1; lev--) {
    D.lev = lev;
    const int n = N >> (lev - 1);
    dim3 grid((n + INV_TILE - 1) / INV_TILE, nb);
    launch_idwt_level(D, grid, s);
    c->launches++;
  }
  CK(cudaGetLastError());
  return RBEPWT_OK;
}

int alloc_state(rbepwt_ctx *c, int B, int H, int W, int levels, int path_mode, unsigned flags) {
  c->has_encoding = false; c->has_paths = false; c->is_dwt2 = false;
  c->B = B; c->H = H; c->W = W; c->N = H * W; c->logW = ilog2(W); c->levels = levels; c->mode = path_mode;
  c->enc_flags = flags;
  c->totalR = 0;
  c->h_R.assign(B, 0); c->h_rbase.assign(B, 0);
  const size_t N = c->N;
  CK(c->Q.ensure((size_t)B * 2 * N * 4));
  CK(c->Pm.ensure((size_t)B * 2 * N * 4));
  CK(c->posmap.ensure((size_t)B * N * 4));
  CK(c->coefs.ensure((size_t)B * N * 8));
  CK(c->thr.ensure((size_t)B * sizeof(ThrRec)));
  CK(c->need_exact.ensure((size_t)B * sizeof(int)));
  CK(cudaMemsetAsync(c->thr.p, 0, (size_t)B * sizeof(ThrRec), c->stream));  // before the fork: ordered ahead of every stream
  c->thr_pending = false;
  CK(c->img_R.ensure((size_t)B * 4)); CK(c->img_rbase.ensure((size_t)B * 4));
  CK(c->img_labmin.ensure((size_t)B * 4)); CK(c->img_direct.ensure((size_t)B * 4));
  if (B > c->pin_cap) {
    if (c->pin_R) cudaFreeHost(c->pin_R);
    if (c->pin_rbase) cudaFreeHost(c->pin_rbase);
    if (c->pin_direct) cudaFreeHost(c->pin_direct);
    c->pin_R = c->pin_rbase = c->pin_direct = nullptr; c->pin_cap = 0;
    CK(cudaMallocHost((void **)&c->pin_R, (size_t)B * 4));
    CK(cudaMallocHost((void **)&c->pin_rbase, (size_t)B * 4));
    CK(cudaMallocHost((void **)&c->pin_direct, (size_t)B * 4));
    c->pin_cap = B;
  }
  const int Bc = chunk_images(c, B, c->N);
  if (path_mode != RBEPWT_PATH_EPWT) {
    CK(c->tbl.ensure((size_t)Bc * 2 * N * 8));
    CK(c->slot_rid.ensure((size_t)Bc * 2 * N * 4));
  }
  for (int i = 0; i < NSLOTS; i++)
    if (c->slot[i].qmeta.p) CK(cudaMemsetAsync(c->slot[i].qmeta.p, 0, QM_SIZE * 4, c->stream));
  return RBEPWT_OK;
}

// The pipeline behind encode / decode / transcode.
//   img_host, lab_host : inputs to stage -- host buffers to copy in, or (c->io.src_on_device) narrow-typed device buffers
//                        to widen; NULL: float64 / int32 inputs already on the device
//   out_dev            : device destination of the decoded float64 images (DO_DECODE)
//   out_host           : host destination to copy them to (NULL: none)
//   c->io              : narrow output type, PSNR and kept-coefficient outputs (rbepwt_transcode_ex)
// Order on the input stream: all label groups first (the path stage needs nothing else and is the long
// pole), then the images sub-batch by sub-batch.  With RBEPWT_OPT_STREAMS = 1 every kernel runs on one
// stream (no overlap between kernels; copies still use their own streams).
int run_pipeline(rbepwt_ctx *c, int what, long long k, const void *img_host, const void *lab_host, double *out_dev,
                 double *out_host) {
  const int B = c->B, N = c->N;
  const int Bc = chunk_images(c, B, N);
  const bool serial = c->nslot == 1;
  int rc;
  auto need_events = [&](std::vector<cudaEvent_t> &v, int n) -> int {
    while ((int)v.size() < n) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); v.push_back(e); }
    return RBEPWT_OK;
  };
  for (int c0 = 0; c0 < B; c0 += Bc) {
    const int nbc = std::min(Bc, B - c0);
    const bool host_io = lab_host || img_host || out_host || (c->io.out_dst && !c->io.out_on_device);
    const int Bs = sub_images(c, nbc, N, host_io);
    const int Bp = group_images(c, nbc, N, Bs);
    const int nsub = (nbc + Bs - 1) / Bs;
    // Path groups as [start, end) in sub-batches.  With host inputs the groups ramp up (1, 1, 2, 4, ...
    // sub-batches up to the full group): the first decoded images leave for the host while most labels are
    // still arriving, which is what lets the output copy overlap the input copy.
    std::vector<int> gstart;
    {
      // device-resident inputs: groups of the full size (2^26 pixels), so that the transform of one group runs under the
      // path kernel of the next (43.6k against 43.2k images/s with one group of 512)
      const int full = Bp / Bs;
      int len = (lab_host && c->opt_group == 0 && !serial) ? 1 : full, first = 1;
      for (int s = 0; s < nsub;) {
        gstart.push_back(s);
        s += std::min(len, full);
        if (len < full) { if (!first) len *= 2; first = 0; }
      }
      gstart.push_back(nsub);
    }
    const int ngrp = (int)gstart.size() - 1;
    std::vector<int> group_of(nsub);
    for (int g = 0; g < ngrp; g++)
      for (int s = gstart[g]; s < gstart[g + 1]; s++) group_of[s] = g;
    auto grp_a = [&](int g) { return c0 + gstart[g] * Bs; };
    auto grp_nb = [&](int g) { return std::min(c0 + nbc, c0 + gstart[g + 1] * Bs) - grp_a(g); };
    if ((rc = need_events(c->ev_lab, ngrp)) || (rc = need_events(c->ev_path, ngrp)) || (rc = need_events(c->ev_img, nsub)) ||
        (rc = need_events(c->ev_done, nsub)))
      return rc;
    const int nx = serial ? 1 : NXSLOT;
    for (int i = 0; i < c->nslot; i++)
      if ((what & DO_PATHS) && (rc = ensure_slot_workspace(c, c->slot[i], 0))) return rc;
    for (int i = 0; i < nx; i++)
      if ((what & (DO_DWT | DO_DECODE)) && (rc = ensure_slot_workspace(c, c->slot[NSLOT + i], Bs))) return rc;
    if ((rc = fork_streams(c))) return rc;
    if (what & DO_PATHS)
      for (int g = 0; g < ngrp; g++)
        if ((rc = stage_labels(c, c0, grp_a(g), grp_nb(g), lab_host, c->ev_lab[g]))) return rc;
    if ((what & DO_DWT) && img_host)
      for (int s = 0; s < nsub; s++) {
        const int a = c0 + s * Bs, nb = std::min(Bs, c0 + nbc - a);
        if ((rc = stage_images(c, a, nb, img_host, c->ev_img[s]))) return rc;
      }
    if (what & DO_PATHS) {
      auto run_phase = [&](int g, int phases) -> int {
        Slot &sl = c->slot[g % c->nslot];
        cudaStream_t st = serial ? c->slot[0].s : sl.s;
        if ((phases & PHASE_REGIONS) && (c->mode == RBEPWT_PATH_GRAD || c->mode == RBEPWT_PATH_GRAD_CHEB) && (what & DO_DWT) && img_host)
          for (int sb = gstart[g]; sb < gstart[g + 1]; sb++) CK(cudaStreamWaitEvent(st, c->ev_img[sb], 0));  // gradpath reads pixel values
        int rc2 = build_regions_and_paths(c, sl, st, c0, grp_a(g), grp_nb(g), c->ev_lab[g], phases);
        if (rc2) return rc2;
        if (phases & PHASE_WALK) CK(cudaEventRecord(c->ev_path[g], st));
        return RBEPWT_OK;
      };
      // regions(0), regions(1), walk(0), regions(2), walk(1), ...: the first phase runs one group ahead (two path slots:
      // a slot's next group is enqueued on the slot's stream after its previous group's walk, so nothing is overwritten).
      // One stream (RBEPWT_OPT_STREAMS = 1): group after group.
      for (int g = 0; g < ngrp; g++) {
        if (serial) { if ((rc = run_phase(g, PHASE_REGIONS | PHASE_WALK))) return rc; continue; }
        if ((rc = run_phase(g, PHASE_REGIONS))) return rc;
        if (g >= 1 && (rc = run_phase(g - 1, PHASE_WALK))) return rc;
      }
      if (!serial && (rc = run_phase(ngrp - 1, PHASE_WALK))) return rc;
    }
    if (what & (DO_DWT | DO_THRESH | DO_DECODE))
      for (int s = 0; s < nsub; s++) {
        const int a = c0 + s * Bs, nb = std::min(Bs, c0 + nbc - a);
        Slot &sl = c->slot[NSLOT + s % nx];
        cudaStream_t st = serial ? c->slot[0].s : sl.s;
        if (what & DO_PATHS) CK(cudaStreamWaitEvent(st, c->ev_path[group_of[s]], 0));
        if ((what & DO_DWT) && img_host) CK(cudaStreamWaitEvent(st, c->ev_img[s], 0));
        if (what & DO_DWT)
          if ((rc = transform_sub(c, sl, st, a, nb))) return rc;
        if (what & DO_THRESH)
          if ((rc = threshold_sub(c, st, a, nb, k))) return rc;
        if ((what & DO_THRESH) && c->io.kept_idx_host) {  // the kept coefficients as (flat index, value) pairs
          const long long kk = c->io.kept_k;
          k_kept_pairs<<<nb, 1024, 0, st>>>(c->coefs.as<double>() + (size_t)a * N, N, c->thr.as<ThrRec>() + a, kk,
                                            c->kept_idx_dev.as<int32_t>() + (size_t)a * kk, c->kept_val_dev.as<double>() + (size_t)a * kk);
          c->launches++;
        }
        bool copy_img = false;
        const int odt = c->io.out_dtype;
        if (what & DO_DECODE) {
          if ((rc = decode_sub(c, sl, st, a, nb, out_dev))) return rc;
          if (c->io.psnr_host) {  // PSNR of every decoded image against its input (rbepwt.py:361-368)
            k6_psnr<<<nb, 1024, 0, st>>>(c->img_dev + (size_t)a * N, out_dev + (size_t)a * N, (long long)N, c->psnr_dev.as<double>() + a);
            c->launches++;
          }
          if (c->io.out_dst && odt != RBEPWT_F64) {  // decoded images in a narrow type: converted on the device
            const size_t esz = dtype_size(odt), cnt = (size_t)nb * N;
            char *dev_dst = c->io.out_on_device ? static_cast<char *>(c->io.out_dst) + (size_t)a * N * esz
                                                : c->out_stage.as<char>() + (size_t)a * N * esz;
            const unsigned grid = (unsigned)std::min<size_t>((cnt + 1023) / 1024, 4096);
            if (odt == RBEPWT_U8) k_narrow_pixels<uint8_t><<<grid, 256, 0, st>>>(out_dev + (size_t)a * N, reinterpret_cast<uint8_t *>(dev_dst), cnt);
            else k_narrow_pixels<float><<<grid, 256, 0, st>>>(out_dev + (size_t)a * N, reinterpret_cast<float *>(dev_dst), cnt);
            c->launches++;
          }
          copy_img = out_host || (c->io.out_dst && odt != RBEPWT_F64 && !c->io.out_on_device);
        }
        const bool copy_psnr = (what & DO_DECODE) && c->io.psnr_host, copy_kept = (what & DO_THRESH) && c->io.kept_idx_host;
        if (copy_img || copy_psnr || copy_kept) {  // this sub-batch's results leave on the output stream
          CK(cudaEventRecord(c->ev_done[s], st));
          CK(cudaStreamWaitEvent(c->s_out, c->ev_done[s], 0));
          StageTimer t(c, RBEPWT_T_D2H, c->s_out);
          if (copy_img && out_host)
            CK(cudaMemcpyAsync(out_host + (size_t)a * N, out_dev + (size_t)a * N, (size_t)nb * N * 8, cudaMemcpyDeviceToHost,
                               c->s_out));
          else if (copy_img) {
            const size_t esz = dtype_size(odt);
            CK(cudaMemcpyAsync(static_cast<char *>(c->io.out_dst) + (size_t)a * N * esz, c->out_stage.as<char>() + (size_t)a * N * esz,
                               (size_t)nb * N * esz, cudaMemcpyDeviceToHost, c->s_out));
          }
          if (copy_psnr)
            CK(cudaMemcpyAsync(c->io.psnr_host + a, c->psnr_dev.as<double>() + a, (size_t)nb * 8, cudaMemcpyDeviceToHost, c->s_out));
          if (copy_kept) {
            const long long kk = c->io.kept_k;
            CK(cudaMemcpyAsync(c->io.kept_idx_host + (size_t)a * kk, c->kept_idx_dev.as<int32_t>() + (size_t)a * kk,
                               (size_t)nb * kk * 4, cudaMemcpyDeviceToHost, c->s_out));
            CK(cudaMemcpyAsync(c->io.kept_val_host + (size_t)a * kk, c->kept_val_dev.as<double>() + (size_t)a * kk,
                               (size_t)nb * kk * 8, cudaMemcpyDeviceToHost, c->s_out));
          }
        }
      }
    if ((rc = join_streams(c))) return rc;
  }
  return RBEPWT_OK;
}

// shared front end of encode / full_decode / transcode: validate, size the state, attach the inputs
int begin_batch(rbepwt_ctx *c, const void *img, const void *labels, int B, int H, int W, int levels, int path_mode,
                unsigned flags, bool need_img) {
  int rc = validate_shape(B, H, W, levels, path_mode);
  if (rc) return rc;
  if (!c->has_wavelet) return fail(RBEPWT_E_NO_WAVELET, "rbepwt_set_wavelet has not been called");
  if (path_mode != RBEPWT_PATH_EPWT && !labels) return fail(RBEPWT_E_ARG, "labels are required unless path_mode is EPWT");
  if (c->evs.size() > 65536) clear_events(c);  // stage events accumulate until rbepwt_get_timings reads them
  if ((rc = alloc_state(c, B, H, W, levels, path_mode, flags))) return rc;
  const size_t n = (size_t)B * c->N;
  const bool dev = (flags & RBEPWT_DEVICE_PTRS) != 0;
  // inputs are used where they lie only when they are device-resident AND of the kernels' own types
  const bool own_lab = !dev || c->io.lab_dtype != RBEPWT_I32, own_img = !dev || c->io.img_dtype != RBEPWT_F64;
  c->io.src_on_device = dev;
  if (path_mode == RBEPWT_PATH_EPWT) c->labels_dev = nullptr;
  else if (own_lab) {
    CK(c->labels_own.ensure(n * 4)); c->labels_dev = c->labels_own.as<int32_t>();
    if (!dev && c->io.lab_dtype != RBEPWT_I32) CK(c->lab_stage.ensure(n * label_size(c->io.lab_dtype)));
  } else c->labels_dev = static_cast<const int32_t *>(labels);
  if (!need_img) c->img_dev = nullptr;
  else if (own_img) {
    CK(c->img_own.ensure(n * 8)); c->img_dev = c->img_own.as<double>();
    if (!dev && c->io.img_dtype != RBEPWT_F64) CK(c->img_stage.ensure(n * dtype_size(c->io.img_dtype)));
  } else c->img_dev = static_cast<const double *>(img);
  return RBEPWT_OK;
}

}  // namespace

// ------------------------------------------------------------------------- C ABI --------

extern "C" {

const char *rbepwt_last_error(void) { return g_err.c_str(); }

static int create_impl(rbepwt_ctx *c, int device, void *stream) {
  c->device = device;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
  else { CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  // copies and transform kernels outrank the long path kernels: when an SM slot frees up, their CTAs go first
  int prio_lo = 0, prio_hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  CK(cudaStreamCreateWithPriority(&c->s_in, cudaStreamNonBlocking, prio_hi));
  CK(cudaStreamCreateWithPriority(&c->s_out, cudaStreamNonBlocking, prio_hi));
  for (int i = 0; i < NSLOTS; i++) {
    CK(cudaStreamCreateWithPriority(&c->slot[i].s, cudaStreamNonBlocking, i < NSLOT ? prio_lo : prio_hi));
    // a path slot's auxiliary stream carries the windowed path kernel: few CTAs, the longest chains of the group.
    // It must not queue behind the bulk kernel's thousands of CTAs (measured: the stage lasts 12.5 instead of
    // 10.7 ms when it does), so it outranks it.
    CK(cudaStreamCreateWithPriority(&c->slot[i].aux, cudaStreamNonBlocking, prio_hi));
    CK(cudaStreamCreateWithPriority(&c->slot[i].aux2, cudaStreamNonBlocking, prio_hi));
    CK(cudaEventCreateWithFlags(&c->slot[i].ev_c, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->slot[i].ev_a, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->slot[i].ev_b, cudaEventDisableTiming));
  }
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  CK(cudaMallocHost((void **)&c->pin_err, sizeof(int)));
  CK(c->filt.ensure(4 * FMAX * sizeof(double)));
  CK(c->unit_lut.ensure(2 * TPR_LUT_ROWS * TPR_LUT_COLS));
  k_build_unit_lut<MODE_EUCLID><<<1, 256, 0, c->stream>>>(c->unit_lut.as<uint8_t>());
  k_build_unit_lut<MODE_CHEB><<<1, 256, 0, c->stream>>>(c->unit_lut.as<uint8_t>() + TPR_LUT_ROWS * TPR_LUT_COLS);
  CK(c->t2_tab.ensure(T2_BYTES));
  k_build_t2<<<(T2_JOBS + 255) / 256, 256, 0, c->stream>>>(c->t2_tab.as<uint8_t>());
  // the path kernels fill the SM's shared memory with their arenas
  CK(cudaFuncSetAttribute(k1_walk<MODE_EUCLID, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wk_arena_bytes(false)));
  CK(cudaFuncSetAttribute(k1_walk<MODE_EUCLID, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wk_arena_bytes(true)));
  CK(cudaFuncSetAttribute(k1_walk<MODE_CHEB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wk_arena_bytes(false)));
  CK(cudaFuncSetAttribute(k1_walk<MODE_CHEB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wk_arena_bytes(true)));
  CK(cudaFuncSetAttribute(k1_walk<MODE_EUCLID, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(k1_walk<MODE_EUCLID, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(k1_walk<MODE_CHEB, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaFuncSetAttribute(k1_walk<MODE_CHEB, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  return RBEPWT_OK;
}

int rbepwt_create(int device, void *stream, rbepwt_ctx **out) {
  if (!out) return fail(RBEPWT_E_ARG, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(RBEPWT_E_NO_GPU, "no CUDA device available (%s); rbepwt_b200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= ndev) return fail(RBEPWT_E_ARG, "device %d out of range (%d devices)", device, ndev);
  DeviceGuard g(device);  // the caller's current device is restored on every return path
  rbepwt_ctx *c = new rbepwt_ctx();
  const int rc = create_impl(c, device, stream);
  if (rc) {  // release whatever was created (destroy tolerates the members still unset)
    const std::string msg = g_err;
    rbepwt_destroy(c);
    g_err = msg;
    return rc;
  }
  *out = c;
  return RBEPWT_OK;
}

void rbepwt_destroy(rbepwt_ctx *c) {
  if (!c) return;
  DeviceGuard g(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (cudaStream_t st : {c->s_in, c->s_out}) if (st) cudaStreamSynchronize(st);
  for (int i = 0; i < NSLOTS; i++) if (c->slot[i].s) cudaStreamSynchronize(c->slot[i].s);
  clear_events(c);
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  for (auto *v : {&c->ev_lab, &c->ev_path, &c->ev_img, &c->ev_done})
    for (auto e : *v) cudaEventDestroy(e);
  DevBuf *bufs[] = {&c->filt, &c->unit_lut, &c->t2_tab, &c->labels_own, &c->img_own, &c->out_own, &c->coef_up, &c->Q, &c->Pm, &c->posmap, &c->coefs, &c->thr, &c->need_exact, &c->dwt2_tmp, &c->img_stage, &c->lab_stage, &c->out_stage, &c->psnr_dev, &c->kept_idx_dev, &c->kept_val_dev, &c->img_R,
                    &c->img_rbase, &c->img_labmin, &c->img_direct, &c->tbl, &c->slot_rid, &c->scratch_i32,
                    &c->scratch_i32b, &c->psnr_out, &c->nz_out};
  for (auto b : bufs) b->release();
  for (auto &b : c->reg) b.release();
  for (int i = 0; i < NSLOTS; i++) {
    Slot &sl = c->slot[i];
    DevBuf *sb[] = {&sl.VA, &sl.VB, &sl.Vpix, &sl.queue, &sl.qhist, &sl.qmeta, &sl.qbins, &sl.chunk_start, &sl.chunk_cnt, &sl.gscratch, &sl.gbm, &sl.slot_of};
    for (auto b : sb) b->release();
    if (sl.s) cudaStreamDestroy(sl.s);
    if (sl.aux) cudaStreamDestroy(sl.aux);
    if (sl.aux2) cudaStreamDestroy(sl.aux2);
    if (sl.ev_c) cudaEventDestroy(sl.ev_c);
    if (sl.ev_a) cudaEventDestroy(sl.ev_a);
    if (sl.ev_b) cudaEventDestroy(sl.ev_b);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  if (c->pin_R) cudaFreeHost(c->pin_R);
  if (c->pin_rbase) cudaFreeHost(c->pin_rbase);
  if (c->pin_direct) cudaFreeHost(c->pin_direct);
  if (c->pin_err) cudaFreeHost(c->pin_err);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int rbepwt_sync(rbepwt_ctx *c) {
  if (!c) return fail(RBEPWT_E_ARG, "ctx is NULL");
  DeviceGuard g(c->device);
  // device-pointer calls are asynchronous: this is where a path kernel's "corrupt region state" flag surfaces
  return check_path_error(c);
}

int rbepwt_set_option(rbepwt_ctx *c, int option, int64_t value) {
  if (!c) return fail(RBEPWT_E_ARG, "ctx is NULL");
  switch (option) {
    case RBEPWT_OPT_STREAMS:
      if (value < 1 || value > NSLOT) return fail(RBEPWT_E_ARG, "RBEPWT_OPT_STREAMS must be 1..%d", NSLOT);
      c->nslot = (int)value;
      return RBEPWT_OK;
    case RBEPWT_OPT_SUBBATCH:
      if (value < 0 || value > (1 << 20)) return fail(RBEPWT_E_ARG, "RBEPWT_OPT_SUBBATCH out of range");
      c->opt_sub = (int)value;
      return RBEPWT_OK;
    case RBEPWT_OPT_COOP_LIMIT:
      if (value < -1 || value > (1 << 30)) return fail(RBEPWT_E_ARG, "RBEPWT_OPT_COOP_LIMIT out of range");
      c->opt_coop_limit = (int)value;
      return RBEPWT_OK;
    case RBEPWT_OPT_PATHGROUP:
      if (value < 0 || value > (1 << 20)) return fail(RBEPWT_E_ARG, "RBEPWT_OPT_PATHGROUP out of range");
      c->opt_group = (int)value;
      return RBEPWT_OK;
  }
  return fail(RBEPWT_E_ARG, "unknown option %d", option);
}

int rbepwt_set_wavelet(rbepwt_ctx *c, int flen, const double *dec_lo, const double *dec_hi, const double *rec_lo,
                       const double *rec_hi) {
  if (!c) return fail(RBEPWT_E_ARG, "ctx is NULL");
  if (flen < 2 || flen > FMAX || (flen & 1)) return fail(RBEPWT_E_ARG, "filter length must be even, 2..%d", FMAX);
  DeviceGuard g(c->device);
  std::vector<double> h(4 * FMAX, 0.0);
  memcpy(&h[0], dec_lo, flen * 8); memcpy(&h[FMAX], dec_hi, flen * 8);
  memcpy(&h[2 * FMAX], rec_lo, flen * 8); memcpy(&h[3 * FMAX], rec_hi, flen * 8);
  CK(cudaMemcpyAsync(c->filt.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (flen <= FT_MAX) {
    const double *src[4] = {dec_lo, dec_hi, rec_lo, rec_hi};
    for (int f = 0; f < 4; f++)
      for (int i = 0; i < FT_MAX; i++) c->h_filt[f][i] = i < flen ? src[f][i] : 0.0;
  }
  c->flen = flen;
  c->has_wavelet = true;
  return RBEPWT_OK;
}

int rbepwt_encode(rbepwt_ctx *c, const double *img, const int32_t *labels, int B, int H, int W, int levels,
                  int path_mode, unsigned flags) {
  if (!c || !img) return fail(RBEPWT_E_ARG, "ctx / img is NULL");
  DeviceGuard g(c->device);
  c->io = rbepwt_ctx::IoSpec();
  int rc = begin_batch(c, img, labels, B, H, W, levels, path_mode, flags, true);
  if (rc) return rc;
  const bool host = !(flags & RBEPWT_DEVICE_PTRS);
  if ((rc = run_pipeline(c, DO_PATHS | DO_DWT, 0, host ? img : nullptr, host ? labels : nullptr, nullptr, nullptr))) return rc;
  c->has_paths = true;
  c->has_encoding = true;
  if (host) return check_path_error(c);  // host-pointer calls are synchronous
  return RBEPWT_OK;
}

int rbepwt_threshold(rbepwt_ctx *c, int64_t k) {
  if (!c) return fail(RBEPWT_E_ARG, "ctx is NULL");
  if (!c->has_encoding) return fail(RBEPWT_E_NO_ENCODING, "There is no saved encoding to decode");
  DeviceGuard g(c->device);
  int rc = materialise_thresholds(c);  // thresholding twice: the second one sees the first one's zeros
  if (rc) return rc;
  return threshold_sub(c, c->stream, 0, c->B, (long long)k);
}

int rbepwt_threshold_percentage(rbepwt_ctx *c, double perc) {
  if (!c) return fail(RBEPWT_E_ARG, "ctx is NULL");
  if (!c->has_encoding) return fail(RBEPWT_E_NO_ENCODING, "There is no saved encoding to decode");
  if (!c->has_paths || c->totalR <= 0) return fail(RBEPWT_E_ARG, "threshold_by_percentage needs the regions of an encoding made by this context");
  if (!(perc >= 0.0)) return fail(RBEPWT_E_ARG, "perc must be >= 0");
  if (c->levels > PERC_MAXLEV) return fail(RBEPWT_E_ARG, "too many levels");
  DeviceGuard g(c->device);
  int rc = materialise_thresholds(c);
  if (rc) return rc;
  k4_percentage<<<c->totalR, PERC_THREADS, 0, c->stream>>>(c->coefs.as<double>(), c->N, c->levels, c->reg[7].as<int32_t>(),
                                                         c->reg[3].as<int32_t>(), c->reg[2].as<int32_t>(), 0, perc);
  c->launches++;
  CK(cudaGetLastError());
  return RBEPWT_OK;
}

// ---- the tensor-product baseline (class Dwt, rbepwt.py:2249-2298): see dwt2.cuh ------------------------------------
static int dwt2_decode(rbepwt_ctx *c, double *out_img, unsigned flags) {
  const bool host = !(flags & RBEPWT_DEVICE_PTRS);
  const size_t bytes = (size_t)c->B * c->N * 8;
  int rc = materialise_thresholds(c);  // (the pending-threshold shortcut belongs to the path transform's loader)
  if (rc) return rc;
  double *D = out_img;
  if (host) { CK(c->out_own.ensure(bytes)); D = c->out_own.as<double>(); }
  CK(c->dwt2_tmp.ensure(bytes));
  cudaStream_t s = c->stream;
  CK(cudaMemcpyAsync(D, c->coefs.p, bytes, cudaMemcpyDeviceToDevice, s));
  Dwt2Params P;
  P.filt = c->filt.as<double>(); P.W = c->W; P.N = c->N; P.flen = c->flen;
  for (int lev = c->levels; lev >= 1; lev--) {
    P.s = c->W >> (lev - 1);
    const dim3 grid((P.s + 255) / 256, P.s, c->B);
    P.src = D; P.dst = c->dwt2_tmp.as<double>(); P.clip = 0;
    k_dwt2_inv<1><<<grid, 256, 0, s>>>(P);
    P.src = c->dwt2_tmp.as<double>(); P.dst = D; P.clip = (lev == 1 && !(flags & RBEPWT_NO_CLIP)) ? 1 : 0;
    k_dwt2_inv<0><<<grid, 256, 0, s>>>(P);
    c->launches += 2;
  }
  CK(cudaGetLastError());
  if (host) {
    CK(cudaMemcpyAsync(out_img, D, bytes, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return RBEPWT_OK;
}

int rbepwt_dwt2_encode(rbepwt_ctx *c, const double *img, int B, int H, int W, int levels, unsigned flags) {
  if (!c || !img) return fail(RBEPWT_E_ARG, "ctx / img is NULL");
  int rc = validate_shape(B, H, W, levels, RBEPWT_PATH_EUCLID);
  if (rc) return rc;
  if (H != W) return fail(RBEPWT_E_ARG, "the 2-D DWT baseline needs a square image (the reference reshapes its sub-bands as squares, rbepwt.py:2283-2296)");
  if ((1 << levels) > W) return fail(RBEPWT_E_LEVELS, "the 2-D DWT baseline needs 2^levels <= side");
  if (!c->has_wavelet) return fail(RBEPWT_E_NO_WAVELET, "rbepwt_set_wavelet has not been called");
  DeviceGuard g(c->device);
  c->io = rbepwt_ctx::IoSpec();
  c->has_encoding = false; c->has_paths = false; c->is_dwt2 = true;
  c->B = B; c->H = H; c->W = W; c->N = H * W; c->logW = ilog2(W); c->levels = levels; c->mode = RBEPWT_PATH_EUCLID;
  c->enc_flags = 0; c->totalR = 0;
  c->h_R.assign(B, 0); c->h_rbase.assign(B, 0);
  const size_t bytes = (size_t)B * c->N * 8;
  CK(c->coefs.ensure(bytes));
  CK(c->dwt2_tmp.ensure(bytes));
  CK(c->thr.ensure((size_t)B * sizeof(ThrRec)));
  CK(c->need_exact.ensure((size_t)B * sizeof(int)));
  cudaStream_t s = c->stream;
  CK(cudaMemsetAsync(c->thr.p, 0, (size_t)B * sizeof(ThrRec), s));
  c->thr_pending = false;
  CK(cudaMemcpyAsync(c->coefs.p, img, bytes, (flags & RBEPWT_DEVICE_PTRS) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  Dwt2Params P;
  P.filt = c->filt.as<double>(); P.W = W; P.N = c->N; P.flen = c->flen; P.clip = 0;
  for (int lev = 1; lev <= levels; lev++) {
    P.s = W >> (lev - 1);
    P.src = c->coefs.as<double>(); P.dst = c->dwt2_tmp.as<double>();
    k_dwt2_fwd<0><<<dim3((P.s + 255) / 256, P.s / 2, B), 256, 0, s>>>(P);
    P.src = c->dwt2_tmp.as<double>(); P.dst = c->coefs.as<double>();
    k_dwt2_fwd<1><<<dim3((P.s / 2 + 255) / 256, P.s, B), 256, 0, s>>>(P);
    c->launches += 2;
  }
  CK(cudaGetLastError());
  c->has_encoding = true;
  if (!(flags & RBEPWT_DEVICE_PTRS)) CK(cudaStreamSynchronize(s));
  return RBEPWT_OK;
}

int rbepwt_decode(rbepwt_ctx *c, double *out_img, unsigned flags) {
  if (!c || !out_img) return fail(RBEPWT_E_ARG, "ctx / out is NULL");
  if (!c->has_encoding) return fail(RBEPWT_E_NO_ENCODING, "There is no saved encoding to decode");
  DeviceGuard g(c->device);
  c->io = rbepwt_ctx::IoSpec();
  if (c->is_dwt2) return dwt2_decode(c, out_img, flags);
  const bool host = !(flags & RBEPWT_DEVICE_PTRS);
  double *out_dev = out_img;
  if (host) {
    CK(c->out_own.ensure((size_t)c->B * c->N * 8));
    out_dev = c->out_own.as<double>();
  }
  c->decode_noclip = (flags & RBEPWT_NO_CLIP) != 0;
  int rc = run_pipeline(c, DO_DECODE, 0, nullptr, nullptr, out_dev, host ? out_img : nullptr);
  c->decode_noclip = false;
  if (rc) return rc;
  if (host) return check_path_error(c);
  return RBEPWT_OK;
}

int rbepwt_transcode(rbepwt_ctx *c, const double *img, const int32_t *labels, int B, int H, int W, int levels,
                     int path_mode, int64_t k, double *out_img, unsigned flags) {
  if (!c || !img || !out_img) return fail(RBEPWT_E_ARG, "ctx / img / out is NULL");
  DeviceGuard g(c->device);
  c->io = rbepwt_ctx::IoSpec();
  int rc = begin_batch(c, img, labels, B, H, W, levels, path_mode, flags, true);
  if (rc) return rc;
  const bool host = !(flags & RBEPWT_DEVICE_PTRS);
  double *out_dev = out_img;
  if (host) {
    CK(c->out_own.ensure((size_t)B * c->N * 8));
    out_dev = c->out_own.as<double>();
  }
  if ((rc = run_pipeline(c, DO_PATHS | DO_DWT | DO_THRESH | DO_DECODE, (long long)k, host ? img : nullptr,
                         host ? labels : nullptr, out_dev, host ? out_img : nullptr)))
    return rc;
  c->has_paths = true;
  c->has_encoding = true;
  if (host) return check_path_error(c);
  return RBEPWT_OK;
}

int rbepwt_transcode_ex(rbepwt_ctx *c, const void *img, int img_dtype, const void *labels, int label_dtype, int B, int H,
                        int W, int levels, int path_mode, int64_t k, void *out_img, int out_dtype, double *psnr_out,
                        int32_t *kept_idx, double *kept_val, unsigned flags) {
  if (!c || !img) return fail(RBEPWT_E_ARG, "ctx / img is NULL");
  if (img_dtype != RBEPWT_F64 && img_dtype != RBEPWT_F32 && img_dtype != RBEPWT_U8) return fail(RBEPWT_E_ARG, "unknown image element type %d", img_dtype);
  if (out_dtype != RBEPWT_F64 && out_dtype != RBEPWT_F32 && out_dtype != RBEPWT_U8) return fail(RBEPWT_E_ARG, "unknown output element type %d", out_dtype);
  if (label_dtype != RBEPWT_I32 && label_dtype != RBEPWT_U16) return fail(RBEPWT_E_ARG, "unknown label element type %d", label_dtype);
  if ((kept_idx == nullptr) != (kept_val == nullptr)) return fail(RBEPWT_E_ARG, "kept_idx and kept_val go together");
  if (kept_idx && (k < 1 || k >= (int64_t)H * W)) return fail(RBEPWT_E_ARG, "the kept-coefficient output needs 1 <= k < H*W");
  DeviceGuard g(c->device);
  const bool host = !(flags & RBEPWT_DEVICE_PTRS);
  c->io = rbepwt_ctx::IoSpec();
  c->io.img_dtype = img_dtype; c->io.lab_dtype = label_dtype; c->io.out_dtype = out_dtype;
  // a uint8 image makes the EPWT walk compare values the way numpy uint8 scalars subtract (rbepwt.py:1302)
  if (img_dtype == RBEPWT_U8 && path_mode == RBEPWT_PATH_EPWT) flags |= RBEPWT_U8_WRAP;
  int rc = begin_batch(c, img, labels, B, H, W, levels, path_mode, flags, true);
  if (rc) return rc;
  const size_t n = (size_t)B * c->N;
  const bool want_decode = out_img != nullptr || psnr_out != nullptr;
  double *out_dev = nullptr;
  double *out_host = nullptr;
  if (want_decode) {
    if (out_img && out_dtype == RBEPWT_F64 && !host) out_dev = static_cast<double *>(out_img);
    else { CK(c->out_own.ensure(n * 8)); out_dev = c->out_own.as<double>(); }
    if (out_img && out_dtype == RBEPWT_F64 && host) out_host = static_cast<double *>(out_img);
    if (out_img && out_dtype != RBEPWT_F64) {
      c->io.out_dst = out_img; c->io.out_on_device = !host;
      if (host) CK(c->out_stage.ensure(n * dtype_size(out_dtype)));
    }
  }
  if (psnr_out) { CK(c->psnr_dev.ensure((size_t)B * 8)); c->io.psnr_host = psnr_out; }
  if (kept_idx) {
    CK(c->kept_idx_dev.ensure((size_t)B * k * 4)); CK(c->kept_val_dev.ensure((size_t)B * k * 8));
    c->io.kept_idx_host = kept_idx; c->io.kept_val_host = kept_val; c->io.kept_k = k;
  }
  const bool stage_img = host || img_dtype != RBEPWT_F64, stage_lab = host || label_dtype != RBEPWT_I32;
  rc = run_pipeline(c, DO_PATHS | DO_DWT | DO_THRESH | (want_decode ? DO_DECODE : 0), (long long)k, stage_img ? img : nullptr,
                    (stage_lab && path_mode != RBEPWT_PATH_EPWT) ? labels : nullptr, out_dev, out_host);
  c->io = rbepwt_ctx::IoSpec();
  if (rc) return rc;
  c->has_paths = true;
  c->has_encoding = true;
  if (host || psnr_out || kept_idx) return check_path_error(c);  // anything written to host memory: synchronous
  return RBEPWT_OK;
}

int rbepwt_full_decode(rbepwt_ctx *c, const double *coefs, const int32_t *labels, int B, int H, int W, int levels,
                       int path_mode, double *out_img, unsigned flags) {
  if (!c || !coefs || !out_img) return fail(RBEPWT_E_ARG, "ctx / coefs / out is NULL");
  if (path_mode == RBEPWT_PATH_EPWT || path_mode == RBEPWT_PATH_GRAD || path_mode == RBEPWT_PATH_GRAD_CHEB) {
    int rc = validate_shape(B, H, W, levels, path_mode);
    if (rc) return rc;
    return fail(RBEPWT_E_ARG, "full_decode needs value-independent paths (EPWT and gradpath paths depend on the image)");
  }
  DeviceGuard g(c->device);
  c->io = rbepwt_ctx::IoSpec();
  int rc = begin_batch(c, nullptr, labels, B, H, W, levels, path_mode, flags, false);
  if (rc) return rc;
  const bool host = !(flags & RBEPWT_DEVICE_PTRS);
  if ((rc = run_pipeline(c, DO_PATHS, 0, nullptr, host ? labels : nullptr, nullptr, nullptr))) return rc;
  c->has_paths = true;
  CK(cudaMemcpyAsync(c->coefs.p, coefs, (size_t)B * c->N * 8, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                     c->stream));
  c->has_encoding = true;
  return rbepwt_decode(c, out_img, flags);
}

int rbepwt_psnr(rbepwt_ctx *c, const double *a, const double *b, int B, int64_t n, double *out, unsigned flags) {
  if (!c || !a || !b || !out || B < 1 || n < 1) return fail(RBEPWT_E_ARG, "bad psnr arguments");
  DeviceGuard g(c->device);
  const size_t bytes = (size_t)B * n * 8;
  const double *da = a, *db = b;
  if (!(flags & RBEPWT_DEVICE_PTRS)) {
    CK(c->scratch_i32.ensure(bytes)); CK(c->scratch_i32b.ensure(bytes));
    CK(cudaMemcpyAsync(c->scratch_i32.p, a, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->scratch_i32b.p, b, bytes, cudaMemcpyHostToDevice, c->stream));
    da = c->scratch_i32.as<double>(); db = c->scratch_i32b.as<double>();
  }
  CK(c->psnr_out.ensure((size_t)B * 8));
  k6_psnr<<<B, 1024, 0, c->stream>>>(da, db, (long long)n, c->psnr_out.as<double>());
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, c->psnr_out.p, (size_t)B * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return RBEPWT_OK;
}

int rbepwt_nonzero_coefs(rbepwt_ctx *c, int64_t *out) {
  if (!c || !out) return fail(RBEPWT_E_ARG, "ctx / out is NULL");
  if (!c->has_encoding) return fail(RBEPWT_E_NO_ENCODING, "There is no saved encoding to decode");
  DeviceGuard g(c->device);
  int rc = materialise_thresholds(c);
  if (rc) return rc;
  CK(c->nz_out.ensure((size_t)c->B * 8));
  k_nonzero<<<c->B, 256, 0, c->stream>>>(c->coefs.as<double>(), c->N, c->nz_out.as<long long>());
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, c->nz_out.p, (size_t)c->B * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return RBEPWT_OK;
}

static int check_img(rbepwt_ctx *c, int b, bool need_enc) {
  if (!c) return fail(RBEPWT_E_ARG, "ctx is NULL");
  if (need_enc ? !c->has_encoding : !c->has_paths) return fail(RBEPWT_E_NO_ENCODING, "There is no saved encoding to decode");
  if (b < 0 || b >= c->B) return fail(RBEPWT_E_ARG, "image index %d out of range", b);
  return RBEPWT_OK;
}

int rbepwt_get_coefs(rbepwt_ctx *c, int b, double *flat) {
  int rc = check_img(c, b, true);
  if (rc) return rc;
  DeviceGuard g(c->device);
  if ((rc = materialise_thresholds(c))) return rc;
  CK(cudaMemcpyAsync(flat, c->coefs.as<double>() + (size_t)b * c->N, (size_t)c->N * 8, cudaMemcpyDeviceToHost, c->stream));
  return check_path_error(c);
}

int rbepwt_set_coefs(rbepwt_ctx *c, int b, const double *flat) {
  int rc = check_img(c, b, true);
  if (rc) return rc;
  DeviceGuard g(c->device);
  if ((rc = materialise_thresholds(c))) return rc;
  CK(cudaMemcpyAsync(c->coefs.as<double>() + (size_t)b * c->N, flat, (size_t)c->N * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return RBEPWT_OK;
}

int rbepwt_region_count(rbepwt_ctx *c, int b, int32_t *R) {
  int rc = check_img(c, b, false);
  if (rc) return rc;
  *R = c->h_R[b];
  return RBEPWT_OK;
}

static int copy_region_array(rbepwt_ctx *c, int b, int which, int32_t *out) {
  DeviceGuard g(c->device);
  CK(cudaMemcpyAsync(out, c->reg[which].as<int32_t>() + c->h_rbase[b], (size_t)c->h_R[b] * 4, cudaMemcpyDeviceToHost,
                     c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return RBEPWT_OK;
}

int rbepwt_region_offsets(rbepwt_ctx *c, int b, int32_t *off) {
  int rc = check_img(c, b, false);
  if (rc) return rc;
  if ((rc = copy_region_array(c, b, 3, off))) return rc;
  off[c->h_R[b]] = c->N;
  return RBEPWT_OK;
}

int rbepwt_region_labels(rbepwt_ctx *c, int b, int32_t *labels) {
  int rc = check_img(c, b, false);
  if (rc) return rc;
  return copy_region_array(c, b, 0, labels);
}

// level-1 incoming order of image b into scratch_i32 (device)
static int level1_incoming(rbepwt_ctx *c, int b) {
  CK(c->scratch_i32.ensure((size_t)c->N * 8));
  int32_t *inc = c->scratch_i32.as<int32_t>();
  if (c->mode == RBEPWT_PATH_EPWT) {  // one region, row-major
    std::vector<int32_t> id(c->N);
    for (int i = 0; i < c->N; i++) id[i] = i;
    CK(cudaMemcpyAsync(inc, id.data(), (size_t)c->N * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return RBEPWT_OK;
  }
  const int R = c->h_R[b];
  k_level1_incoming<<<(R * 32 + 255) / 256, 256, 0, c->stream>>>(c->labels_dev + (size_t)b * c->N, c->logW, c->regs(),
                                                                 c->h_rbase[b], R, inc);
  c->launches++;
  CK(cudaGetLastError());
  return RBEPWT_OK;
}

int rbepwt_get_paths(rbepwt_ctx *c, int b, int level, int32_t *pix) {
  int rc = check_img(c, b, c && c->mode == RBEPWT_PATH_EPWT);
  if (rc) return rc;
  if (level < 0 || level > c->levels + 1) return fail(RBEPWT_E_ARG, "level %d out of range", level);
  DeviceGuard g(c->device);
  const size_t N = c->N;
  const int32_t *Q = c->Q.as<int32_t>() + (size_t)b * 2 * N;
  if (level == 0) {
    if ((rc = level1_incoming(c, b))) return rc;
    CK(cudaMemcpyAsync(pix, c->scratch_i32.p, N * 4, cudaMemcpyDeviceToHost, c->stream));
  } else if (level <= c->levels) {
    CK(cudaMemcpyAsync(pix, Q + level_off(N, level), (N >> (level - 1)) * 4, cudaMemcpyDeviceToHost, c->stream));
  } else {  // approximation's points: even positions of the last level's paths
    const size_t n = N >> (level - 1);
    CK(cudaMemcpy2DAsync(pix, 4, Q + level_off(N, c->levels), 8, 4, n, cudaMemcpyDeviceToHost, c->stream));
  }
  return check_path_error(c);
}

int rbepwt_get_perm(rbepwt_ctx *c, int b, int level, int32_t *perm) {
  int rc = check_img(c, b, c && c->mode == RBEPWT_PATH_EPWT);
  if (rc) return rc;
  if (level < 1 || level > c->levels) return fail(RBEPWT_E_ARG, "level %d out of range", level);
  DeviceGuard g(c->device);
  const size_t N = c->N;
  const int n = (int)(N >> (level - 1));
  const int32_t *Q = c->Q.as<int32_t>() + (size_t)b * 2 * N;
  CK(c->scratch_i32b.ensure(N * 8));
  int32_t *inv = c->scratch_i32b.as<int32_t>(), *out = inv + N;
  if (level == 1) {
    if ((rc = level1_incoming(c, b))) return rc;
    k_inv_scatter<<<(n + 255) / 256, 256, 0, c->stream>>>(c->scratch_i32.as<int32_t>(), 1, n, inv);
  } else {
    k_inv_scatter<<<(n + 255) / 256, 256, 0, c->stream>>>(Q + level_off(N, level - 1), 2, n, inv);
  }
  k_perm_gather<<<(n + 255) / 256, 256, 0, c->stream>>>(Q + level_off(N, level), n, inv,
                                                        c->reg[3].as<int32_t>() + c->h_rbase[b], c->h_R[b], level, out);
  c->launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(perm, out, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  return check_path_error(c);
}

int rbepwt_get_level_values(rbepwt_ctx *c, int b, int level, double *vals) {
  int rc = check_img(c, b, true);
  if (rc) return rc;
  if (level < 1 || level > c->levels) return fail(RBEPWT_E_ARG, "level %d out of range", level);
  if (!c->img_dev) return fail(RBEPWT_E_ARG, "no image is attached to this encoding");
  DeviceGuard g(c->device);
  const size_t N = c->N;
  const int n = (int)(N >> (level - 1));
  cudaStream_t s = c->stream;
  CK(c->coef_up.ensure(N * 8 * 2));
  double *scratch = c->coef_up.as<double>(), *out = scratch + N;
  if ((rc = ensure_slot_workspace(c, c->slot[NSLOT], 1))) return rc;
  double *V[2] = {c->slot[NSLOT].VA.as<double>(), c->slot[NSLOT].VB.as<double>()};
  DwtParams D;
  D.Q = c->Q.as<int32_t>() + (size_t)b * 2 * N;
  D.Pm = c->Pm.as<int32_t>() + (size_t)b * 2 * N;
  D.coefs = scratch;
  D.filt = c->filt.as<double>();
  D.out_img = nullptr;
  D.flen = c->flen; D.N = c->N; D.levels = 31;  // never the "last" level: low-pass always goes to vout
  D.clip = 1; D.thr = nullptr;
  set_taps(c, D, false);
  for (int lev = 1; lev < level; lev++) {
    D.lev = lev;
    D.vin = c->img_dev + (size_t)b * N;
    D.vin_stride = N;
    D.plane[0] = V[0]; D.plane[1] = V[1];
    const int half = (int)((N >> (lev - 1)) >> 1);
    launch_dwt_level(D, dim3((half + FWD_TILE - 1) / FWD_TILE, 1), s);
    c->launches++;
  }
  if (level == 1) {
    if ((rc = level1_incoming(c, b))) return rc;
    k_gather_values<<<(n + 255) / 256, 256, 0, s>>>(c->img_dev + (size_t)b * N, c->scratch_i32.as<int32_t>(), 1, n, out);
  } else {  // x^level = cA of level-1, already in the incoming order
    CK(cudaMemcpyAsync(out, V[(level - 1) & 1], (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
  }
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(vals, out, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  return check_path_error(c);
}

int rbepwt_enable_timing(rbepwt_ctx *c, int on) {
  if (!c) return fail(RBEPWT_E_ARG, "ctx is NULL");
  c->timing = on != 0;
  if (!on) { DeviceGuard g(c->device); cudaStreamSynchronize(c->stream); clear_events(c); }
  return RBEPWT_OK;
}

int rbepwt_get_timings(rbepwt_ctx *c, float *ms, int n) {
  if (!c || !ms) return fail(RBEPWT_E_ARG, "ctx / ms is NULL");
  DeviceGuard g(c->device);
  CK(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; i++) ms[i] = 0.f;
  for (int i = 0; i < RBEPWT_T_COUNT; i++) c->stage_launches[i] = 0;
  for (auto &e : c->evs) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, e.a, e.b));
    if (e.stage < n) ms[e.stage] += t;
    c->stage_launches[e.stage] += e.launches;
  }
  clear_events(c);
  return RBEPWT_T_COUNT;
}

int rbepwt_get_stage_launches(rbepwt_ctx *c, int64_t *launches, int n) {
  if (!c || !launches) return fail(RBEPWT_E_ARG, "ctx / launches is NULL");
  for (int i = 0; i < n; i++) launches[i] = i < RBEPWT_T_COUNT ? c->stage_launches[i] : 0;
  return RBEPWT_T_COUNT;
}

int64_t rbepwt_launch_count(rbepwt_ctx *c) { return c ? c->launches : 0; }

int rbepwt_felzenszwalb(const double *img, int H, int W, double scale, double sigma, int min_size, int32_t *labels,
                        int32_t *nlabels) {
  if (!img || !labels) return fail(RBEPWT_E_ARG, "img and labels must not be NULL");
  if (H < 1 || W < 1 || (long long)H * W > (1ll << 29)) return fail(RBEPWT_E_ARG, "image shape out of range");
  if (!(scale >= 0.0) || !(sigma >= 0.0) || min_size < 0) return fail(RBEPWT_E_ARG, "scale, sigma and min_size must be non-negative");
  try {
    const int n = felzenszwalb(img, H, W, scale, sigma, min_size, labels);
    if (nlabels) *nlabels = n;
  } catch (const std::exception &e) {
    return fail(RBEPWT_E_CUDA, "felzenszwalb: %s", e.what());
  }
  return RBEPWT_OK;
}

#ifdef WK_STATS  // debug build only (tools/wk_stats.py): warp trips and lane units by kind of k1_walk
int rbepwt_debug_wk_stats(rbepwt_ctx *c, unsigned long long *out, int reset) {
  DeviceGuard g(c->device);
  cudaStreamSynchronize(c->stream);
  cudaMemcpyFromSymbol(out, g_wk_stats, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(g_wk_stats, z, sizeof z); }
  return 0;
}
// the named queue slots (qmeta[Q_BINS ..]) and the per-class table (qbins) of path slot `slot` after the last group it ran
int rbepwt_debug_queue(rbepwt_ctx *c, int slot, int *out) {
  DeviceGuard g(c->device);
  sync_internal(c);
  cudaStreamSynchronize(c->stream);
  cudaMemcpy(out, c->slot[slot].qmeta.as<int>() + Q_BINS, (QM_SIZE - Q_BINS) * sizeof(int), cudaMemcpyDeviceToHost);
  cudaMemcpy(out + 16, c->slot[slot].qbins.as<int>(), 3 * Q_NCLS * sizeof(int), cudaMemcpyDeviceToHost);
  return 0;
}
#endif

}  // extern "C"
