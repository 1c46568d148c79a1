// ofri_api.cu -- handle, workspace, the coarse-to-fine driver (genericPyramidalOpticalFlow restated for the
// GPU: GenericPyramidalOpticalFlow.py:238-416) and the extern "C" boundary declared in include/ofri.h.
// Host-side only; every numeric stage is a kernel in ofri_stages.cu / ofri_hs.cu / ofri_ls.cu.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "ofri_internal.h"
#include "ofri_pixel.cuh"
#include "ofri_tables.h"

using namespace ofri;

extern "C" int ofri_gaussian_taps(double sigma, int n_taps, float* taps_out);

namespace {

std::mutex g_err_mutex;
std::string g_last_error;   // errors without a handle (ofri_create)

struct DevResizeTaps { int* xmin = nullptr; int* cnt = nullptr; double* w = nullptr; int kmax = 0; };
struct DevSplineSys { double* lo = nullptr; double* cp = nullptr; double* den = nullptr; int conv = 0; double den_c = 0, cp_c = 0; };

struct StageTime { std::string name; cudaEvent_t e0, e1; };

}  // namespace

struct ofri_ctx {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr, s_in = nullptr, s_out = nullptr, s_aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // U / V spline up-samples run concurrently on stream + s_aux
  cudaStream_t s_comm = nullptr;                      // row-band mode: ghost-row exchanges overlapping interior tiles
  cudaEvent_t ev_c0 = nullptr, ev_c1 = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  std::string err;
  // bump arena for per-call workspace (stream-ordered reuse)
  char* arena = nullptr;
  size_t arena_cap = 0, arena_off = 0;
  // staging buffers of the host-pointer entry points (double buffered)
  char* stage = nullptr;
  size_t stage_cap = 0;
  // pinned bounce ring of the host-pointer path for PAGEABLE caller buffers (same slot layout as `stage`)
  char* hstage = nullptr;
  size_t hstage_cap = 0;
  const ofri_params* ext_params = nullptr;
  int ext_rc = 0;                     // error code of a failed external compute() (run_adapter returns -1)
  ofri_adapter_fn ext_fn = nullptr;   // set for the duration of ofri_pyramidal_flow_external
  void* ext_user = nullptr;
  std::vector<float> ext_buf;         // dense host copies handed to the callback: im1, im2, U, V
  int last_host_path = 0;        // read-only: 1 = direct copies (pinned caller buffers / one chunk), 2 = pinned bounce ring
  int last_hs_fuse_fine = 0, last_hs_fuse_coarse = 0, last_ls_fuse = 0;   // read-only: fuse factors the last call used
  std::map<std::pair<int, int>, DevResizeTaps> taps, taps_bilinear;
  ofri_farneback_params fb = {};      // parameters of adapters of kind OFRI_ALGO_FB (ofri_set_farneback)
  bool fb_set = false;
  ofri_lk_params lk = {};             // parameters of adapters of kind OFRI_ALGO_LK (ofri_set_lk)
  bool lk_set = false;
  std::map<int, DevSplineSys> splines;
  LaunchCounter lc;
  // options
  int hs_fuse = 4, hs_variant = 24, ls_fuse = 4, ls_variant = 8, chunk_pairs = 0, timing = 0;
  int hs_precise = 1;   // 0 = fast f32/FMA everywhere, 1 = reference arithmetic where a coarse level's HS result reaches the warp unrefined (hs_needs_precise), 2 = everywhere
  // row-band (domain-decomposed) mode
  ofri::Comm* comm = nullptr;
  int hs_fuse_fast = 0;     // fuse factor of the fast-arithmetic Horn-Schunck launches only (0 = hs_fuse)
  int last_chunk_pairs = 0; // read-only: pairs per chunk the last batched call used
  int hs_fuse_precise = 0;  // fuse factor of the reference-arithmetic launches only (0 = hs_fuse); not used in row-band mode
  int auto_fuse = 1;        // deeper fusion for launches that cannot fill the GPU (see eff_hs_fuse)
  int band_exchange = 32;   // Horn-Schunck sweeps between two ghost-row exchanges (rounded up to a multiple of hs_fuse)
  int band_reach = 8;       // rows the warp may reach beyond a band's ghost frame (>= max |v| / 2 + 2)
  int band_reserve_sms = 8; // SMs an interior Horn-Schunck launch leaves free for the ghost-row exchange beside it (NCCL kernels)
  int host_bounce = 1;      // host-pointer path: 1 = pageable caller buffers go through the pinned bounce ring, 0 = always direct copies
  int spline_variant = 1;   // 1 = chunk-parallel windowed solves + fused row kernel, 0 = sequential line solves (A/B)
  // timings of the last call
  std::vector<StageTime> times;
  std::vector<std::pair<std::string, float>> times_ms;
};

namespace {

int fail(ofri_handle h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  else {
    std::lock_guard<std::mutex> g(g_err_mutex);
    g_last_error = buf;
  }
  return code;
}

#define OFRI_CUDA(h, call)                                                                             \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess)                                                                             \
      return fail(h, e_ == cudaErrorMemoryAllocation ? OFRI_ERR_OOM : OFRI_ERR_CUDA, "%s failed: %s", #call, \
                  cudaGetErrorString(e_));                                                             \
  } while (0)

inline long round_up(long v, long m) { return (v + m - 1) / m * m; }

// ---- workspace ---------------------------------------------------------------------------------------------------
int arena_reserve(ofri_handle h, size_t bytes) {
  if (bytes <= h->arena_cap) return OFRI_OK;
  OFRI_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->arena) cudaFree(h->arena);
  h->arena = nullptr;
  h->arena_cap = 0;
  size_t want = bytes + (bytes >> 3);
  cudaError_t e = cudaMalloc(&h->arena, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&h->arena, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(h, OFRI_ERR_OOM, "cudaMalloc of %zu workspace bytes failed: %s", want, cudaGetErrorString(e));
  }
  h->arena_cap = want;
  return OFRI_OK;
}
struct Bump {
  char* base;
  size_t off = 0, cap;
  bool dry;
  Bump(char* b, size_t c, bool d) : base(b), cap(c), dry(d) {}
  void* take(size_t bytes) {
    size_t o = (off + 255) & ~(size_t)255;
    off = o + bytes;
    return dry ? nullptr : (void*)(base + o);
  }
  Img plane(int batch, int H, int W) {
    Img m;
    m.H = H; m.W = W; m.batch = batch;
    m.pitch = round_up(W, 4);
    m.stride = m.pitch * H;
    m.p = (float*)take(sizeof(float) * (size_t)m.stride * batch);
    return m;
  }
  ImgD planed(int batch, int H, int W) {
    ImgD m;
    m.H = H; m.W = W; m.batch = batch;
    m.pitch = W;
    m.stride = (long)W * H;
    m.p = (double*)take(sizeof(double) * (size_t)m.stride * batch);
    return m;
  }
};
// view of a plane buffer (allocated for the finest level) at a coarser level's size
Img view(const Img& base, int H, int W) {
  Img m = base;
  m.H = H; m.W = W;
  m.pitch = round_up(W, 4);
  m.stride = m.pitch * H;
  return m;
}
ImgD viewd(const ImgD& base, int H, int W) {
  ImgD m = base;
  m.H = H; m.W = W; m.pitch = W; m.stride = (long)W * H;
  return m;
}
Img dense(const float* p, int batch, int H, int W) {
  Img m;
  m.p = const_cast<float*>(p);
  m.H = H; m.W = W; m.batch = batch; m.pitch = W; m.stride = (long)W * H;
  return m;
}

// ---- cached per-size tables ----------------------------------------------------------------------------------------
int get_resize_taps(ofri_handle h, int in_size, int out_size, ResizeTaps* out, bool bilinear = false) {
  auto key = std::make_pair(in_size, out_size);
  auto& cache = bilinear ? h->taps_bilinear : h->taps;
  auto it = cache.find(key);
  if (it == cache.end()) {
    HostResizeTaps t = build_resize_taps(in_size, out_size, bilinear);
    DevResizeTaps d;
    d.kmax = t.kmax;
    OFRI_CUDA(h, cudaMalloc(&d.xmin, sizeof(int) * out_size));
    OFRI_CUDA(h, cudaMalloc(&d.cnt, sizeof(int) * out_size));
    OFRI_CUDA(h, cudaMalloc(&d.w, sizeof(double) * t.w.size()));
    OFRI_CUDA(h, cudaMemcpy(d.xmin, t.xmin.data(), sizeof(int) * out_size, cudaMemcpyHostToDevice));
    OFRI_CUDA(h, cudaMemcpy(d.cnt, t.cnt.data(), sizeof(int) * out_size, cudaMemcpyHostToDevice));
    OFRI_CUDA(h, cudaMemcpy(d.w, t.w.data(), sizeof(double) * t.w.size(), cudaMemcpyHostToDevice));
    it = cache.emplace(key, d).first;
  }
  out->xmin = it->second.xmin;
  out->cnt = it->second.cnt;
  out->w = it->second.w;
  out->kmax = it->second.kmax;
  out->in_size = in_size;
  out->out_size = out_size;
  return OFRI_OK;
}
int get_spline_sys(ofri_handle h, int n, SplineSys* out) {
  auto it = h->splines.find(n);
  if (it == h->splines.end()) {
    HostSplineSys t = build_spline_sys(n);
    const int m = n - 2;
    DevSplineSys d;
    OFRI_CUDA(h, cudaMalloc(&d.lo, sizeof(double) * m));
    OFRI_CUDA(h, cudaMalloc(&d.cp, sizeof(double) * m));
    OFRI_CUDA(h, cudaMalloc(&d.den, sizeof(double) * m));
    OFRI_CUDA(h, cudaMemcpy(d.lo, t.lo.data(), sizeof(double) * m, cudaMemcpyHostToDevice));
    OFRI_CUDA(h, cudaMemcpy(d.cp, t.cp.data(), sizeof(double) * m, cudaMemcpyHostToDevice));
    OFRI_CUDA(h, cudaMemcpy(d.den, t.den.data(), sizeof(double) * m, cudaMemcpyHostToDevice));
    d.conv = t.conv;
    d.den_c = t.den[t.conv < m ? t.conv : m - 1];
    d.cp_c = t.cp[t.conv < m ? t.conv : m - 1];
    it = h->splines.emplace(n, d).first;
  }
  out->lo = it->second.lo;
  out->cp = it->second.cp;
  out->den = it->second.den;
  out->n = n;
  out->conv = it->second.conv;
  out->den_c = it->second.den_c;
  out->cp_c = it->second.cp_c;
  out->rcp_c = 1.0 / it->second.den_c;
  return OFRI_OK;
}

int level_size(int n, double scale) { return level_size_half_even(n, scale); }

// ---- timing -----------------------------------------------------------------------------------------------------------
struct Timed {
  ofri_handle h;
  bool on;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  std::string name;
  Timed(ofri_handle hh, const char* n) : h(hh), on(hh->timing != 0), name(n) {
    if (on) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, h->stream);
    }
  }
  ~Timed() {
    if (on) {
      cudaEventRecord(e1, h->stream);
      h->times.push_back({name, e0, e1});
    }
  }
};
void collect_times(ofri_handle h) {
  h->times_ms.clear();
  if (h->times.empty()) return;
  cudaStreamSynchronize(h->stream);
  std::map<std::string, float> agg;
  std::vector<std::string> order;
  for (auto& t : h->times) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t.e0, t.e1);
    if (!agg.count(t.name)) order.push_back(t.name);
    agg[t.name] += ms;
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  for (auto& n : order) h->times_ms.emplace_back(n, agg[n]);
  h->times.clear();
}

// ---- parameter validation ------------------------------------------------------------------------------------------------
int check_params(ofri_handle h, const ofri_params* p, int H, int W) {
  if (!p) return fail(h, OFRI_ERR_INVALID, "params is NULL");
  if (p->size != sizeof(ofri_params))
    return fail(h, OFRI_ERR_INVALID, "ofri_params.size = %u, library expects %zu (ABI mismatch)", p->size,
                sizeof(ofri_params));
  if (p->pyramid_levels < 1 || p->pyramid_levels > 16) return fail(h, OFRI_ERR_INVALID, "Invalid scale level");
  if (p->k_levels < 0) return fail(h, OFRI_ERR_INVALID, "k_levels < 0");
  if (p->n_taps_lsw < 0 || p->n_taps_lsw > OFRI_MAX_GAUSS_TAPS || (p->n_taps_lsw && !(p->n_taps_lsw & 1)))
    return fail(h, OFRI_ERR_INVALID, "bad Gaussian tap count");
  if (p->n_taps_main < 0 || p->n_taps_main > OFRI_MAX_GAUSS_TAPS || p->n_taps_opt < 0 ||
      p->n_taps_opt > OFRI_MAX_GAUSS_TAPS || (p->n_taps_main && !(p->n_taps_main & 1)) ||
      (p->n_taps_opt && !(p->n_taps_opt & 1)))
    return fail(h, OFRI_ERR_INVALID, "bad Gaussian tap count");
  const ofri_algo* algos[2] = {&p->main_algo, &p->opt_algo};
  for (int a = 0; a < 2; ++a) {
    const ofri_algo* g = algos[a];
    if (g->kind == OFRI_ALGO_FB) {
      if (!h || !h->fb_set) return fail(h, OFRI_ERR_INVALID, "Farneback adapter without ofri_set_farneback");
      continue;
    }
    if (g->kind == OFRI_ALGO_LK) {
      if (!h || !h->lk_set) return fail(h, OFRI_ERR_INVALID, "Lucas-Kanade adapter without ofri_set_lk");
      continue;
    }
    const bool ext_ok = g->kind == OFRI_ALGO_EXTERNAL && h && h->ext_fn;
    if (a == 0 && g->kind != OFRI_ALGO_HS && g->kind != OFRI_ALGO_LS && !ext_ok)
      return fail(h, OFRI_ERR_INVALID, "main adapter must be HS or LS (external adapters: ofri_pyramidal_flow_external)");
    if (ext_ok) continue;
    if (g->kind == OFRI_ALGO_HS) {
      if (g->n_alphas < 0 || g->n_alphas > OFRI_MAX_ALPHAS) return fail(h, OFRI_ERR_INVALID, "bad n_alphas");
      if (g->n_alphas < p->pyramid_levels * p->k_levels)
        return fail(h, OFRI_ERR_ALPHAS, "pop from empty list");   // IndexError at HornSchunck.py:36
      if (g->hs_niter < 0) return fail(h, OFRI_ERR_INVALID, "Niter < 0");
    } else if (g->kind == OFRI_ALGO_LS) {
      if (g->ls_maxiter < 1 || g->ls_maxiter > 100000) return fail(h, OFRI_ERR_INVALID, "bad ls_maxiter");
    } else if (g->kind != OFRI_ALGO_NONE) {
      return fail(h, OFRI_ERR_INVALID, "unknown adapter kind %d", g->kind);
    }
  }
  // every level must have >= 4 samples per axis for the cubic spline (scipy raises otherwise) and the Gaussian
  // padding needs n >= half kernel
  double scale = 1.0 / std::pow(2.0, p->pyramid_levels - 1);
  for (int l = 1; l <= p->pyramid_levels; ++l) {
    int hl = l == p->pyramid_levels ? H : level_size(H, scale);
    int wl = l == p->pyramid_levels ? W : level_size(W, scale);
    if (hl < 1 || wl < 1) return fail(h, OFRI_ERR_TOO_SMALL, "pyramid level %d is empty (%d x %d)", l, hl, wl);
    if (l < p->pyramid_levels && (hl < 4 || wl < 4))
      return fail(h, OFRI_ERR_TOO_SMALL, "pyramid level %d is %d x %d; the cubic spline needs >= 4 samples", l, hl, wl);
    int hk = (p->n_taps_main > p->n_taps_opt ? p->n_taps_main : p->n_taps_opt) / 2;
    if (hl < hk || wl < hk) return fail(h, OFRI_ERR_TOO_SMALL, "level %d smaller than the Gaussian half-width", l);
    if (p->warping && !p->bilinear && (l > 1 || p->k_levels > 1) && (hl < 36 || wl < 36))
      return fail(h, OFRI_ERR_TOO_SMALL, "level %d smaller than the 73-tap Gaussian of the Liu-Shen warp", l);
    scale *= 2.0;
  }
  return OFRI_OK;
}

// ---- the driver ------------------------------------------------------------------------------------------------------------
struct Workspace {
  Img lvl1, lvl2, warp1, warp2, work1, work2, opt1, opt2, tmp, fx, fy, ft, U[2], V[2], U0, V0, Uacc, Vacc, us, vs;
  Img lsw_sc, lsw_out;      // biLinear = False: scattered frame, warped copy of the k > 0 branch
  LsPlanes ls;
  ImgD M1, T1, M2, M1b, T1b, M2b;
  double* hs_acc = nullptr;
  double* ls_errs = nullptr;
  int* ls_state = nullptr;
  unsigned* ls_max = nullptr;
  int* lsw_flag = nullptr;
  FbWorkspace fb;
};

void plan_fb_workspace(Bump& b, int batch, int H, int W, FbWorkspace* f) {
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) f->flow[i][j] = b.plane(batch, H, W);
  f->blur = b.plane(batch, H, W);
  f->level = b.plane(batch, H, W);
  f->tmp = b.plane(batch, H, W);
  for (int i = 0; i < 3; ++i) f->poly[i] = b.plane(batch, H, W);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 5; ++j) f->R[i][j] = b.plane(batch, H, W);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 5; ++j) f->M[i][j] = b.plane(batch, H, W);
  f->d_imgs = (Img*)b.take(sizeof(Img) * 16);
}

void plan_workspace(Bump& b, int batch, int H, int W, const ofri_params* p, Workspace* ws) {
  const bool multi = p->pyramid_levels > 1;
  const bool has_opt = p->opt_algo.kind != OFRI_ALGO_NONE;
  const bool has_ls = p->main_algo.kind == OFRI_ALGO_LS || p->opt_algo.kind == OFRI_ALGO_LS;
  const bool has_hs = p->main_algo.kind == OFRI_ALGO_HS || p->opt_algo.kind == OFRI_ALGO_HS;
  if (multi) { ws->lvl1 = b.plane(batch, H, W); ws->lvl2 = b.plane(batch, H, W); }
  if (multi || p->k_levels > 1) { ws->warp1 = b.plane(batch, H, W); ws->warp2 = b.plane(batch, H, W); }
  ws->work1 = b.plane(batch, H, W);
  ws->work2 = b.plane(batch, H, W);
  if (has_opt) { ws->opt1 = b.plane(batch, H, W); ws->opt2 = b.plane(batch, H, W); }
  ws->tmp = b.plane(batch, H, W);
  if (has_hs) { ws->fx = b.plane(batch, H, W); ws->fy = b.plane(batch, H, W); ws->ft = b.plane(batch, H, W); }
  for (int i = 0; i < 2; ++i) { ws->U[i] = b.plane(batch, H, W); ws->V[i] = b.plane(batch, H, W); }
  ws->U0 = b.plane(batch, H, W);
  ws->V0 = b.plane(batch, H, W);
  ws->Uacc = b.plane(batch, H, W);
  ws->Vacc = b.plane(batch, H, W);
  if (multi) {
    ws->us = b.plane(batch, H, W);
    ws->vs = b.plane(batch, H, W);
    int hc = (H + 1) / 2 + 1, wc = (W + 1) / 2 + 1;      // the coarser level is at most this big
    ws->M1 = b.planed(batch, hc, wc);
    ws->T1 = b.planed(batch, H, wc);
    ws->M2 = b.planed(batch, H, wc);
    ws->M1b = b.planed(batch, hc, wc);
    ws->T1b = b.planed(batch, H, wc);
    ws->M2b = b.planed(batch, H, wc);
  }
  if (has_ls) {
    for (int c = 0; c < 8; ++c) ws->ls.c[c] = b.plane(batch, H, W);
    int maxit = 1;
    if (p->main_algo.kind == OFRI_ALGO_LS) maxit = p->main_algo.ls_maxiter;
    if (p->opt_algo.kind == OFRI_ALGO_LS && p->opt_algo.ls_maxiter > maxit) maxit = p->opt_algo.ls_maxiter;
    ws->ls_errs = (double*)b.take(sizeof(double) * 2 * (size_t)maxit * batch);
    ws->ls_state = (int*)b.take(sizeof(int) * 4 * batch);
    ws->ls_max = (unsigned*)b.take(sizeof(unsigned) * 2 * batch);
  }
  ws->hs_acc = (double*)b.take(sizeof(double) * 2 * batch);
  if (p->main_algo.kind == OFRI_ALGO_FB || p->opt_algo.kind == OFRI_ALGO_FB) plan_fb_workspace(b, batch, H, W, &ws->fb);
  if (p->warping && !p->bilinear && (multi || p->k_levels > 1)) {
    if (!multi) { ws->lvl1 = b.plane(batch, H, W); }
    ws->lsw_sc = b.plane(batch, H, W);
    ws->lsw_out = b.plane(batch, H, W);
  }
  ws->lsw_flag = (int*)b.take(sizeof(int) * 4);
}

// Automatic fuse factors (auto_fuse = 1; results do not depend on the fuse factor: bit-identical, tested).
//  * A launch that cannot fill the GPU (fewer tiles than SMs, e.g. one 512 x 512 pair) is bound by launch latency, not
//    by throughput: fuse 8 (Horn-Schunck) / 4 (Liu-Shen) sweeps per launch.
//  * Very large frames (>= 4096 in both directions): the image border wastes < 3 % of the (smaller) T = 8 tiles, and the
//    fast Horn-Schunck kernel, HBM-bound at T = 4, becomes issue-bound at T = 8 and 12 % faster (16384^2: 200.6 ms
//    instead of 225.7 for 600 sweeps); Liu-Shen gains 6 % from T = 4.  At 1024^2 the border waste cancels the gain.
// Row-band mode ties the exchange interval to the configured factor and applies the same rules there.
bool big_frame(int H, int W) { return H >= 4096 && W >= 4096; }
// Tile-time model of the fast Horn-Schunck TMA kernel (phase profile profiles/r2_phase_*.jsonl):
// a 66 x 128 tile costs ~2.2 us of fixed work (staging -> registers, stores, ghost columns) + 0.72 us per fused sweep and
// yields (66 - 2T) x (128 - 2 HX) cells per sweep; the image border wastes a different share of the last tile row /
// column for every T.  Cost of one sweep over an H x W level, in model us per pair:
double hs_tile_cost(int H, int W, int T) {
  const int HX = T <= 4 ? 4 : 8, TW = 128 - 2 * HX, TH = 66 - 2 * T;
  return (double)((W + TW - 1) / TW) * ((H + TH - 1) / TH) * (2.2 + 0.72 * T) / T;
}
int eff_hs_fuse(ofri_handle h, int H, int W, int batch, bool precise = false, int niter = 0) {
  const int fuse = precise ? (h->hs_fuse_precise > 0 ? h->hs_fuse_precise : h->hs_fuse)
                           : (h->hs_fuse_fast > 0 ? h->hs_fuse_fast : h->hs_fuse);
  if (!h->auto_fuse || fuse < 1 || fuse >= 8) return fuse;
  if (!precise && h->hs_fuse_fast == 0 && big_frame(H, W)) return 8;
  if ((long)((W + 119) / 120) * ((H + 57) / 58) * batch < 148) return 8;
  if (!precise && h->hs_fuse_fast == 0 && fuse == 4 && niter > 0) {
    // per-level choice among the fuse factors the TMA kernel has (4, 6, 8): e.g. 512 x 512 (the coarse level of the
    // 1024 x 1024 workload) loses 15 % of its T = 4 tile area to the border but only 9 % at T = 6 (measured: 18.0 ->
    // 16.3 ms per 64 pairs); only factors that divide the sweep count (no tail launches on the slower fall-back kernel)
    int best = fuse;
    double cbest = hs_tile_cost(H, W, fuse) * 0.97;      // change only for a modelled gain of more than 3 %
    for (int T : {6, 8})
      if (niter % T == 0 && hs_tile_cost(H, W, T) < cbest) { best = T; cbest = hs_tile_cost(H, W, T); }
    return best;
  }
  return fuse;
}
int eff_ls_fuse(ofri_handle h, int H, int W, int batch) {
  if (!h->auto_fuse || h->ls_fuse < 1 || h->ls_fuse >= 4) return h->ls_fuse;
  if (big_frame(H, W)) return 4;
  return (long)((W + 119) / 120) * ((H + 29) / 30) * batch < 148 ? 4 : h->ls_fuse;
}

GaussTaps make_taps(const float* k, int n) {
  GaussTaps t;
  t.K = n;
  for (int i = 0; i < OFRI_MAX_GAUSS_TAPS; ++i) t.k[i] = i < n ? k[i] : 0.0f;
  return t;
}

// hs_precise = 1: which Horn-Schunck solves run in the reference's exact arithmetic.  A rounding difference of ~1e-6 px in
// a COARSE level's flow can flip the float32 rounding of a warp coordinate of the next level (GPOF:200-201: one ulp is
// 3e-5 .. 6e-5 px at x ~ 500 .. 1000), which a weakly regularised fine-level solve follows 1:1 -- so a coarse
// Horn-Schunck result that reaches the warp as it is must be exact.  When the Liu-Shen refinement runs after it on the
// same level (the reference's optionalOFlowAlgoAdapter), its 60 sweeps contract the difference before the warp sees it:
// measured final deviation of the all-fast path <= 1.2e-5 px for alpha = 0.5 .. 45 on the bundled pair and on synthetic
// 1024^2 pairs, against up to 2e-2 px without the refinement (tools/precise_vs_fast.py, profiles/r1_precise_vs_fast.jsonl).
bool hs_needs_precise(const ofri_params* p, bool coarse_level, bool is_main) {
  if (!coarse_level) return false;
  // refined by Liu-Shen before the warp -- with the reference's own stopping parameters (60 sweeps, 1e-8: LS:141), the
  // only ones the contraction was measured for; a shortened refinement keeps the exact arithmetic
  if (is_main && p->opt_algo.kind == OFRI_ALGO_LS && p->opt_algo.ls_maxiter >= 60 && p->opt_algo.ls_tol <= 1e-8)
    return false;
  return true;
}

// one adapter compute() on level planes.  U/V state lives in ws.U[cur] / ws.V[cur]; returns the new cur.
// uv_zero: the initial guess is identically zero (lets HS skip the copy of U0 for its error norm).
int run_adapter(ofri_handle h, const ofri_algo& a, int call_index, Workspace& ws, const Img& im1, const Img& im2,
                int Hl, int Wl, int cur, bool uv_zero, bool coarse_level, float* d_err, int err_stride,
                bool finest = true) {
  cudaStream_t s = h->stream;
  Img U[2] = {view(ws.U[0], Hl, Wl), view(ws.U[1], Hl, Wl)};
  Img V[2] = {view(ws.V[0], Hl, Wl), view(ws.V[1], Hl, Wl)};
  if (a.kind == OFRI_ALGO_EXTERNAL) {
    // a foreign adapter's compute() on the host: frames + current U, V down, result up; everything else stays put
    Timed t(h, "external_adapter");
    const size_t n = (size_t)Hl * Wl;
    h->ext_buf.resize(4 * n);
    float* hb = h->ext_buf.data();
    const Img* src[4] = {&im1, &im2, &U[cur], &V[cur]};
    for (int i = 0; i < 4; ++i)
      if (cudaMemcpy2DAsync(hb + i * n, sizeof(float) * Wl, src[i]->p, sizeof(float) * src[i]->pitch, sizeof(float) * Wl, Hl,
                            cudaMemcpyDeviceToHost, s) != cudaSuccess)
        return h->ext_rc = fail(h, OFRI_ERR_CUDA, "D2H for the external adapter failed: %s", cudaGetErrorString(cudaGetLastError())), -1;
    if (cudaStreamSynchronize(s) != cudaSuccess)
      return h->ext_rc = fail(h, OFRI_ERR_CUDA, "execution failed before the external adapter: %s", cudaGetErrorString(cudaGetLastError())), -1;
    float err = 0.0f;
    const int which = (&a == &h->ext_params->opt_algo) ? 1 : 0;
    if (h->ext_fn(h->ext_user, which, call_index, hb, hb + n, hb + 2 * n, hb + 3 * n, Hl, Wl, &err) != 0)
      return h->ext_rc = fail(h, OFRI_ERR_CALLBACK, "the external adapter's compute() failed (call %d)", call_index), -1;
    cudaMemcpy2DAsync(U[cur].p, sizeof(float) * U[cur].pitch, hb + 2 * n, sizeof(float) * Wl, sizeof(float) * Wl, Hl,
                      cudaMemcpyHostToDevice, s);
    cudaMemcpy2DAsync(V[cur].p, sizeof(float) * V[cur].pitch, hb + 3 * n, sizeof(float) * Wl, sizeof(float) * Wl, Hl,
                      cudaMemcpyHostToDevice, s);
    if (d_err) cudaMemcpyAsync(d_err, &err, sizeof(float), cudaMemcpyHostToDevice, s);
    if (cudaStreamSynchronize(s) != cudaSuccess)      // the host buffers are reused by the next compute()
      return h->ext_rc = fail(h, OFRI_ERR_CUDA, "H2D after the external adapter failed: %s", cudaGetErrorString(cudaGetLastError())), -1;
    return cur;
  }
  if (a.kind == OFRI_ALGO_FB) {                       // Farneback_PyCL.compute: (U, V) in -> (U, V) out, error 'Unknown'
    Timed t(h, "farneback");
    int rc = launch_farneback(im1, im2, U[cur], V[cur], &h->fb, ws.fb,
                              [h](int in, int out, ResizeTaps* t) { return get_resize_taps(h, in, out, t, true); }, s, h->lc);
    if (rc) return h->ext_rc = fail(h, rc, "Farneback adapter failed (level %d x %d)", Hl, Wl), -1;
    if (d_err) cudaMemsetAsync(d_err, 0, sizeof(float), s);
    return cur;
  }
  if (a.kind == OFRI_ALGO_LK) {                       // denseLucasKanade_PyCl.compute: (U, V) in -> (U, V) out, error `True`
    Timed t(h, "lucas_kanade");
    int rc = launch_lk(im1, im2, U[cur], V[cur], &h->lk, s, h->lc);
    if (rc) return h->ext_rc = fail(h, rc, "Lucas-Kanade adapter failed (level %d x %d)", Hl, Wl), -1;
    static const float one = 1.0f;                    // the reference returns `calcErr` (= True) as the error (LK:169)
    if (d_err) cudaMemcpyAsync(d_err, &one, sizeof(float), cudaMemcpyHostToDevice, s);
    return cur;
  }
  if (a.kind == OFRI_ALGO_HS) {
    Img fx = view(ws.fx, Hl, Wl), fy = view(ws.fy, Hl, Wl), ft = view(ws.ft, Hl, Wl);
    {
      Timed t(h, "hs_derivs");
      launch_hs_derivs(im1, im2, fx, fy, ft, s, h->lc);
    }
    Img U0, V0;   // null = zero initial guess
    if (!uv_zero && d_err) {
      U0 = view(ws.U0, Hl, Wl);
      V0 = view(ws.V0, Hl, Wl);
      launch_copy(U0, U[cur], s, h->lc);
      launch_copy(V0, V[cur], s, h->lc);
    }
    int res;
    {
      // rounding differences made on a coarse level are amplified by the warp + solve of the finer levels (up to
      // x50 for weakly regularised problems), so the coarse levels default to the reference's exact arithmetic
      // (coarse_level = "a coarse level whose Horn-Schunck result reaches the warp unrefined": see hs_needs_precise)
      const bool precise = h->hs_precise >= 2 || (h->hs_precise == 1 && coarse_level);
      // stage timers: the fast kernel on the finest level / on the coarser levels / the reference-arithmetic kernel
      Timed t(h, precise ? "hs_iterate_precise" : (finest ? "hs_iterate" : "hs_iterate_coarse"));
      const int fuse = eff_hs_fuse(h, Hl, Wl, U[cur].batch, precise, a.hs_niter);
      (finest ? h->last_hs_fuse_fine : h->last_hs_fuse_coarse) = fuse;
      res = launch_hs_iterate(U[cur], V[cur], U[cur ^ 1], V[cur ^ 1], fx, fy, ft, a.alphas[call_index], a.hs_niter, fuse,
                              h->hs_variant, precise, s, h->lc);
    }
    cur = res ? (cur ^ 1) : cur;
    if (d_err) {
      Timed t(h, "hs_error");
      launch_hs_error(U[cur], V[cur], U0, V0, ws.hs_acc, d_err, err_stride, s, h->lc);
    }
    return cur;
  }
  // Liu-Shen: inside the solver u = ROW component (our V), v = COLUMN component (our U)  (LS:38-39)
  LsPlanes co;
  for (int c = 0; c < 8; ++c) co.c[c] = view(ws.ls.c[c], Hl, Wl);
  {
    Timed t(h, "ls_coefficients");
    launch_ls_coefficients(im1, im2, a.ls_h, co, ws.ls_max, s, h->lc);
  }
  {
    Timed t(h, "ls_iterate");
    h->last_ls_fuse = eff_ls_fuse(h, Hl, Wl, U[cur].batch);
    launch_ls_solve(V[cur], U[cur], V[cur ^ 1], U[cur ^ 1], co, a.ls_h, a.ls_maxiter, a.ls_tol,
                    h->last_ls_fuse, h->ls_variant, ws.ls_errs,
                    ws.ls_state, V[cur], U[cur], d_err, err_stride, nullptr, s, h->lc);
  }
  return cur;
}

// Whole pyramid for `batch` pairs resident on the device.  im1/im2/uo/vo may have any pitch.
int run_pyramid(ofri_handle h, const Img& im1, const Img& im2, const ofri_params* p, const Img& uo, const Img& vo,
                float* d_err, Workspace& ws) {
  cudaStream_t s = h->stream;
  const int H = im1.H, W = im1.W, L = p->pyramid_levels, KL = p->k_levels;
  const bool has_opt = p->opt_algo.kind != OFRI_ALGO_NONE;
  const GaussTaps taps_main = make_taps(p->taps_main, p->n_taps_main);
  const GaussTaps taps_opt = make_taps(p->taps_opt, p->n_taps_opt);
  GaussTaps taps_lsw = make_taps(p->taps_lsw, p->n_taps_lsw);
  if (p->warping && !p->bilinear && p->n_taps_lsw == 0) {                   // gaussian_filter(x, 0.6*3, truncate=4/0.6*3)
    float t73[OFRI_MAX_GAUSS_TAPS];
    const double sg = 0.6 * 3, tr = 4.0 / 0.6 * 3;
    const int K = 2 * (int)(tr * sg + 0.5) + 1;
    ofri_gaussian_taps(sg, K, t73);
    taps_lsw = make_taps(t73, K);
  }
  cudaMemsetAsync(ws.lsw_flag, 0, sizeof(int) * 4, s);
  const int err_stride = L * KL * 2;
  double scale = 1.0 / std::pow(2.0, L - 1);
  int prevH = 0, prevW = 0;
  int cur = 0;
  int call_index = 0;
  Img Uacc = ws.Uacc, Vacc = ws.Vacc, us = ws.us, vs = ws.vs;   // buffers; swapped after a warp
  for (int level = 1; level <= L; ++level) {
    const bool last = level == L;
    const bool local_scaling = last ? p->final_scaling != 0 : p->intermediate_scaling != 0;
    int Hl = H, Wl = W;
    Img n1 = im1, n2 = im2;
    if (scale < 1.0 && !last) {                                           // GPOF:336-343
      Hl = level_size(H, scale);
      Wl = level_size(W, scale);
      n1 = view(ws.lvl1, Hl, Wl);
      n2 = view(ws.lvl2, Hl, Wl);
      ResizeTaps tx, ty;
      int rc = get_resize_taps(h, W, Wl, &tx);
      if (rc) return rc;
      rc = get_resize_taps(h, H, Hl, &ty);
      if (rc) return rc;
      Timed t(h, "resize");
      Img tmp = view(ws.tmp, H, Wl);
      launch_resize(im1, tmp, n1, tx, ty, s, h->lc);
      launch_resize(im2, tmp, n2, tx, ty, s, h->lc);
    }
    Img w1 = n1, w2 = n2;
    bool uv_zero = true;
    Img Ucur = view(ws.U[cur], Hl, Wl), Vcur = view(ws.V[cur], Hl, Wl);
    if (level > 1) {                                                       // GPOF:351-355 -> 118-235
      Img ua = view(Uacc, prevH, prevW), va = view(Vacc, prevH, prevW);
      Img un = view(us, Hl, Wl), vn = view(vs, Hl, Wl);
      if (prevH != Hl || prevW != Wl) {
        Timed t(h, "spline_upsample");
        SplineSys sy, sx;
        int rc = get_spline_sys(h, prevH, &sy);
        if (rc) return rc;
        rc = get_spline_sys(h, prevW, &sx);
        if (rc) return rc;
        float mx = 1.0f, my = 1.0f;
        if (local_scaling) {                                               // GPOF:167-172 (f32 division)
          mx = (float)Wl / (float)prevW;
          my = (float)Hl / (float)prevH;
        }
        ImgD M1 = viewd(ws.M1, prevH, prevW), T1 = viewd(ws.T1, Hl, prevW), M2 = viewd(ws.M2, Hl, prevW);
        ImgD M1b = viewd(ws.M1b, prevH, prevW), T1b = viewd(ws.T1b, Hl, prevW), M2b = viewd(ws.M2b, Hl, prevW);
        // the tridiagonal solves have one thread per line (latency bound): run the two components side by side
        cudaEventRecord(h->ev_fork, s);
        cudaStreamWaitEvent(h->s_aux, h->ev_fork, 0);
        bool ok = true;
        if (h->spline_variant == 0) {
          launch_spline_seq(ua, un, mx, sy, sx, M1, T1, M2, s, h->lc);
          launch_spline_seq(va, vn, my, sy, sx, M1b, T1b, M2b, h->s_aux, h->lc);
        } else {                      // T1 (H x w) doubles as the forward-elimination scratch (h x w)
          ok = launch_spline(ua, un, mx, sy, sx, M1, viewd(ws.T1, prevH, prevW), s, h->lc) &&
               launch_spline(va, vn, my, sy, sx, M1b, viewd(ws.T1b, prevH, prevW), h->s_aux, h->lc);
        }
        cudaEventRecord(h->ev_join, h->s_aux);
        cudaStreamWaitEvent(s, h->ev_join, 0);
        if (!ok) return fail(h, OFRI_ERR_CUDA, "spline up-sample launch failed");
      } else {
        launch_copy(un, ua, s, h->lc);
        launch_copy(vn, va, s, h->lc);
        if (local_scaling) { /* factor is exactly 1 */ }
      }
      if (p->warping && !p->bilinear) {                                    // "Liu-Shen warp" (GPOF:204-221)
        Timed t(h, "warp");
        // the reference warps frame 1 IN PLACE (so the optional adapter and the k-loop see the warped frame too) and
        // leaves frame 2 alone; here the warped frame replaces the level's frame 1 (the caller's array is not touched)
        Img n1m = view(ws.lvl1, Hl, Wl);
        launch_liu_shen_warp(n1, un, vn, n1m, (int*)ws.warp2.p, view(ws.work1, Hl, Wl), view(ws.work2, Hl, Wl),
                             view(ws.U0, Hl, Wl), view(ws.V0, Hl, Wl), view(ws.lsw_sc, Hl, Wl), view(ws.tmp, Hl, Wl),
                             taps_lsw, ws.lsw_flag, s, h->lc);
        n1 = n1m;
        w1 = n1;
        w2 = n2;
        std::swap(Uacc, us);
        std::swap(Vacc, vs);
      } else if (p->warping) {
        Timed t(h, "warp");
        w1 = view(ws.warp1, Hl, Wl);
        w2 = view(ws.warp2, Hl, Wl);
        launch_warp_pair(n1, n2, un, vn, w1, w2, s, h->lc);
        std::swap(Uacc, us);                                               // finalUAccum = usNew (GPOF:226-227)
        std::swap(Vacc, vs);
      } else {                                                             // GPOF:228-232
        launch_copy(Ucur, un, s, h->lc);
        launch_copy(Vcur, vn, s, h->lc);
        uv_zero = false;
      }
    }
    Img UaccL = view(Uacc, Hl, Wl), VaccL = view(Vacc, Hl, Wl);
    if (level == 1 || !p->warping) {
      // level 1: Uaccum = Vaccum = 0 (GPOF:365-366); no-warp transition: accumulators restart at 0 (GPOF:231-232)
      cudaMemsetAsync(UaccL.p, 0, sizeof(float) * (size_t)UaccL.stride * UaccL.batch, s);
      cudaMemsetAsync(VaccL.p, 0, sizeof(float) * (size_t)VaccL.stride * VaccL.batch, s);
    }
    if (uv_zero) {
      cudaMemsetAsync(Ucur.p, 0, sizeof(float) * (size_t)Ucur.stride * Ucur.batch, s);
      cudaMemsetAsync(Vcur.p, 0, sizeof(float) * (size_t)Vcur.stride * Vcur.batch, s);
    }
    // pre-filters (GPOF:368-386)
    Img work1 = w1, work2 = w2;
    if (p->n_taps_main > 0) {
      Timed t(h, "gauss");
      work1 = view(ws.work1, Hl, Wl);
      work2 = view(ws.work2, Hl, Wl);
      Img tmp = view(ws.tmp, Hl, Wl);
      launch_gauss(w1, tmp, work1, taps_main, s, h->lc);
      launch_gauss(w2, tmp, work2, taps_main, s, h->lc);
    }
    Img o1 = n1, o2 = n2;
    if (has_opt && p->n_taps_opt > 0) {
      Timed t(h, "gauss");
      o1 = view(ws.opt1, Hl, Wl);
      o2 = view(ws.opt2, Hl, Wl);
      Img tmp = view(ws.tmp, Hl, Wl);
      launch_gauss(n1, tmp, o1, taps_opt, s, h->lc);
      launch_gauss(n2, tmp, o2, taps_opt, s, h->lc);
    }
    for (int k = 0; k < KL; ++k) {
      if (k > 0) {                                                         // GPOF:392-404
        if (p->warping) {
          // same-size "transition": usNew = Uaccum (no spline, no scaling), re-warp the level images
          Timed t(h, "warp");
          if (p->bilinear) {
            w1 = view(ws.warp1, Hl, Wl);
            w2 = view(ws.warp2, Hl, Wl);
            launch_warp_pair(n1, n2, UaccL, VaccL, w1, w2, s, h->lc);
          } else {                                                         // warps a COPY of frame 1 this time (GPOF:394)
            w1 = view(ws.lsw_out, Hl, Wl);
            w2 = n2;
            launch_liu_shen_warp(n1, UaccL, VaccL, w1, (int*)ws.warp2.p, view(ws.work1, Hl, Wl), view(ws.work2, Hl, Wl),
                                 view(ws.U0, Hl, Wl), view(ws.V0, Hl, Wl), view(ws.lsw_sc, Hl, Wl), view(ws.tmp, Hl, Wl),
                                 taps_lsw, ws.lsw_flag, s, h->lc);
          }
          cudaMemsetAsync(Ucur.p, 0, sizeof(float) * (size_t)Ucur.stride * Ucur.batch, s);
          cudaMemsetAsync(Vcur.p, 0, sizeof(float) * (size_t)Vcur.stride * Vcur.batch, s);
          uv_zero = true;
          if (p->n_taps_main > 0 && p->refilter_k) {                       // FILTER > 1 (GPOF:396)
            work1 = view(ws.work1, Hl, Wl);
            work2 = view(ws.work2, Hl, Wl);
            Img tmp = view(ws.tmp, Hl, Wl);
            launch_gauss(w1, tmp, work1, taps_main, s, h->lc);
            launch_gauss(w2, tmp, work2, taps_main, s, h->lc);
          } else {
            work1 = w1;
            work2 = w2;
          }
        } else {
          // U, V = Uaccum, Vaccum; accumulators restart at zero; work images unchanged (GPOF:403-404)
          launch_copy(Ucur, UaccL, s, h->lc);
          launch_copy(Vcur, VaccL, s, h->lc);
          cudaMemsetAsync(UaccL.p, 0, sizeof(float) * (size_t)UaccL.stride * UaccL.batch, s);
          cudaMemsetAsync(VaccL.p, 0, sizeof(float) * (size_t)VaccL.stride * VaccL.batch, s);
          uv_zero = false;
        }
      }
      float* e_main = d_err ? d_err + 2 * call_index : nullptr;
      // "coarse" for the arithmetic rule = a later warp consumes this result: every level but the last, and on the last
      // level every k but the last (the k-loop re-warps by the accumulated flow, GPOF:392-404)
      const bool feeds_warp = !last || (k + 1 < KL && p->warping);
      cur = run_adapter(h, p->main_algo, call_index, ws, work1, work2, Hl, Wl, cur, uv_zero, hs_needs_precise(p, feeds_warp, true),
                        e_main, err_stride, last);
      if (cur < 0) return h->ext_rc;
      if (has_opt) {
        float* e_opt = d_err ? d_err + 2 * call_index + 1 : nullptr;
        cur = run_adapter(h, p->opt_algo, call_index, ws, o1, o2, Hl, Wl, cur, false, hs_needs_precise(p, feeds_warp, false), e_opt,
                          err_stride, last);
        if (cur < 0) return h->ext_rc;
      }
      Ucur = view(ws.U[cur], Hl, Wl);
      Vcur = view(ws.V[cur], Hl, Wl);
      {
        Timed t(h, "accumulate");
        launch_axpy(UaccL, Ucur, s, h->lc);                                // GPOF:413-414
        launch_axpy(VaccL, Vcur, s, h->lc);
      }
      ++call_index;
    }
    prevH = Hl;
    prevW = Wl;
    scale *= 2.0;
  }
  launch_copy(uo, view(Uacc, H, W), s, h->lc);
  launch_copy(vo, view(Vacc, H, W), s, h->lc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, OFRI_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return OFRI_OK;
}

// biLinear = False: did a scatter target leave the frame (the reference raises IndexError at GPOF:207)?  Synchronises.
int check_lsw_flag(ofri_handle h, const ofri_params* p, const Workspace& ws) {
  if (!(p->warping && !p->bilinear && (p->pyramid_levels > 1 || p->k_levels > 1))) return OFRI_OK;
  int flag = 0;
  OFRI_CUDA(h, cudaMemcpyAsync(&flag, ws.lsw_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  OFRI_CUDA(h, cudaStreamSynchronize(h->stream));
  if (flag) return fail(h, OFRI_ERR_INDEX, "index out of bounds: the Liu-Shen warp moved a pixel outside the frame");
  return OFRI_OK;
}

size_t workspace_bytes(int batch, int H, int W, const ofri_params* p) {
  Bump dry(nullptr, 0, true);
  Workspace ws;
  plan_workspace(dry, batch, H, W, p, &ws);
  return dry.off + 4096;
}

int ensure_stage(ofri_handle h, size_t bytes) {
  if (bytes <= h->stage_cap) return OFRI_OK;
  OFRI_CUDA(h, cudaDeviceSynchronize());
  if (h->stage) cudaFree(h->stage);
  h->stage = nullptr;
  h->stage_cap = 0;
  cudaError_t e = cudaMalloc(&h->stage, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(h, OFRI_ERR_OOM, "cudaMalloc of %zu staging bytes failed", bytes);
  }
  h->stage_cap = bytes;
  return OFRI_OK;
}

int ensure_hstage(ofri_handle h, size_t bytes) {
  if (bytes <= h->hstage_cap) return OFRI_OK;
  OFRI_CUDA(h, cudaDeviceSynchronize());
  if (h->hstage) cudaFreeHost(h->hstage);
  h->hstage = nullptr;
  h->hstage_cap = 0;
  cudaError_t e = cudaHostAlloc((void**)&h->hstage, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(h, OFRI_ERR_OOM, "cudaHostAlloc of %zu pinned staging bytes failed", bytes);
  }
  h->hstage_cap = bytes;
  return OFRI_OK;
}
// true iff p points into page-locked (or device-accessible) host memory: async copies from / to it really are async
bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// error exit of the chunked host-pointer path: earlier chunks' copies may still be in flight into the caller's buffers
// (and into staging slots the next call would reuse) -- wait for them before reporting the error
int drain_streams(ofri_handle h, int rc) {
  cudaStreamSynchronize(h->s_in);
  cudaStreamSynchronize(h->stream);
  cudaStreamSynchronize(h->s_out);
  cudaGetLastError();
  return rc;
}

int pick_chunk_impl(ofri_handle h, int batch, int H, int W, const ofri_params* p, size_t extra_per_pair, int cap);
int pick_chunk(ofri_handle h, int batch, int H, int W, const ofri_params* p, size_t extra_per_pair, int cap = 64) {
  h->last_chunk_pairs = pick_chunk_impl(h, batch, H, W, p, extra_per_pair, cap);
  return h->last_chunk_pairs;
}
int pick_chunk_impl(ofri_handle h, int batch, int H, int W, const ofri_params* p, size_t extra_per_pair, int cap) {
  if (h->chunk_pairs > 0) return h->chunk_pairs < batch ? h->chunk_pairs : batch;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  size_t budget = (free_b + h->arena_cap + h->stage_cap) / 2;     // leave half of the device to the caller
  size_t per_pair = workspace_bytes(1, H, W, p) + extra_per_pair;
  long n = (long)(budget / (per_pair ? per_pair : 1));
  if (n < 1) n = 1;
  if (n > cap) n = cap;        // enough to fill 148 SMs many times over; keeps the working set bounded
  // (a wave-fit choice of the chunk size -- tile counts just below a multiple of the SM count -- was measured: no
  // difference, 459.0 / 459.5 vs 460.6 / 459.3 pairs/s, profiles/r1_chunk_wavefit_ab.txt)
  return (int)(n < batch ? n : batch);
}

// generic helper for the stage-level host entry points: upload dense host planes into pitched device planes
struct HostCall {
  ofri_handle h;
  Bump b;
  HostCall(ofri_handle hh) : h(hh), b(hh->arena, hh->arena_cap, false) {}
};
int upload(ofri_handle h, const Img& d, const float* src) {
  OFRI_CUDA(h, cudaMemcpy2DAsync(d.p, sizeof(float) * d.pitch, src, sizeof(float) * d.W, sizeof(float) * d.W,
                                 (size_t)d.H * d.batch, cudaMemcpyHostToDevice, h->stream));
  return OFRI_OK;
}
int download(ofri_handle h, float* dst, const Img& d) {
  OFRI_CUDA(h, cudaMemcpy2DAsync(dst, sizeof(float) * d.W, d.p, sizeof(float) * d.pitch, sizeof(float) * d.W,
                                 (size_t)d.H * d.batch, cudaMemcpyDeviceToHost, h->stream));
  return OFRI_OK;
}
int finish(ofri_handle h) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, OFRI_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) return fail(h, OFRI_ERR_CUDA, "execution failed: %s", cudaGetErrorString(e));
  collect_times(h);
  return OFRI_OK;
}
// NB: planes uploaded with upload() are contiguous over (batch*H) rows only when stride == pitch*H, which Bump::plane
// guarantees.

#define OFRI_ENTER(h)                                                        \
  if (!(h)) return fail(nullptr, OFRI_ERR_INVALID, "NULL handle");           \
  OFRI_CUDA(h, cudaSetDevice((h)->device));                                  \
  (h)->err.clear();

#define OFRI_DIMS(h, batch, H, W)                                                                      \
  if ((batch) < 1 || (H) < 1 || (W) < 1) return fail(h, OFRI_ERR_INVALID, "batch, H, W must be >= 1"); \
  if ((batch) > 65535) return fail(h, OFRI_ERR_INVALID, "batch > 65535 per call");


// =====================================================================================================================
// Row-band domain decomposition: ONE very large frame pair split over the ranks of a communicator.
// =====================================================================================================================
// Rank r owns rows [r H/n, (r+1) H/n) of the frame (and the halved ranges on the coarser pyramid levels).  Every
// per-level array of a rank covers its EXTENDED band = owned rows +- G ghost rows (clipped at the image border) and is
// treated by the stage kernels as a stand-alone image: the kernels need no band logic, a boundary rule applied at an
// artificial band edge only spoils the outermost ghost rows, and every stage shrinks the valid part of the ghost frame
// by its radius.  G = E + 2 + Rw covers: E Horn-Schunck sweeps between two ghost-row exchanges, the 2x2 derivative and
// 3-tap Gaussian rows, and the rows the bilinear warp may reach (Rw, checked at run time).  What is NOT halo-local:
//   * ghost rows of (U, V) are refreshed from the neighbours' owned rows every E sweeps (Comm::exchange);
//   * Liu-Shen's image maxima (LS:96-97) and per-sweep residual sums (LS:79, 141) and Horn-Schunck's error sums
//     (HS:100) are all-reduced, so every rank takes the same decisions and reports the same scalars;
//   * the spline up-sample solves along whole columns: a rank exchanges a halo of the coarse flow with its neighbours
//     and solves only the WINDOW of every column its band needs (windowed Thomas solve, ofri_spline.cuh: bit-identical
//     to the full solve); bands thinner than the halo fall back to an all-gather + redundant full solve;
//   * Pillow's resample and the warp use GLOBAL row coordinates (tap tables of the whole image; float32 rounding of
//     y +- v/2 depends on the magnitude of y).
// Owned rows are therefore bit-identical to the single-GPU result (tested with N virtual bands on one GPU).
struct BandLevel { int Hl, Wl, own0, own1, ext0, ext1, G; };
struct BandPlanInt {
  int L = 0, E = 0, Rw = 0, G = 0, in0 = 0, in1 = 0;
  BandLevel lv[16];
};

int make_band_plan_opts(ofri_handle h, int H, int W, const ofri_params* p, int rank, int n, int hs_fuse,
                        int band_exchange, int band_reach, BandPlanInt* bp);
int make_band_plan(ofri_handle h, int H, int W, const ofri_params* p, int rank, int n, BandPlanInt* bp) {
  return make_band_plan_opts(h, H, W, p, rank, n, h->hs_fuse, h->band_exchange, h->band_reach, bp);
}
// h may be NULL (host-only planning: errors then go to the global message)
int make_band_plan_opts(ofri_handle h, int H, int W, const ofri_params* p, int rank, int n, int hs_fuse,
                        int band_exchange, int band_reach, BandPlanInt* bp) {
  int rc = check_params(h, p, H, W);
  if (rc) return rc;
  if (n < 1 || rank < 0 || rank >= n) return fail(h, OFRI_ERR_INVALID, "bad rank %d of %d", rank, n);
  const int L = p->pyramid_levels;
  if (p->k_levels != 1) return fail(h, OFRI_ERR_UNSUPPORTED, "row-band mode supports kLevels = 1 only");
  if ((p->main_algo.kind != OFRI_ALGO_HS && p->main_algo.kind != OFRI_ALGO_LS) ||
      (p->opt_algo.kind != OFRI_ALGO_NONE && p->opt_algo.kind != OFRI_ALGO_HS && p->opt_algo.kind != OFRI_ALGO_LS))
    return fail(h, OFRI_ERR_UNSUPPORTED, "row-band mode supports the HS and Liu-Shen adapters only");
  if (L > 1 && (!p->warping || !p->bilinear))
    return fail(h, OFRI_ERR_UNSUPPORTED, "row-band mode needs warping = biLinear = True");
  const int fmax = 1 << (L - 1);
  if (H % (n * fmax) != 0)
    return fail(h, OFRI_ERR_UNSUPPORTED, "row-band mode needs H (%d) divisible by ranks x 2^(levels-1) = %d", H, n * fmax);
  int T = hs_fuse > 0 ? (hs_fuse == 7 ? 6 : (hs_fuse > 8 ? 8 : hs_fuse)) : 1;
  int E = band_exchange > 0 ? band_exchange : 32;
  E = (E + T - 1) / T * T;
  bp->L = L;
  bp->E = E;
  bp->Rw = band_reach > 0 ? band_reach : 8;
  bp->G = E + 2 + bp->Rw;
  double scale = 1.0 / std::pow(2.0, L - 1);
  long in0 = H, in1 = 0;
  for (int l = 0; l < L; ++l) {
    const int f = 1 << (L - 1 - l);
    BandLevel& b = bp->lv[l];
    b.Hl = H / f;
    b.Wl = l == L - 1 ? W : level_size(W, scale);
    const int per = b.Hl / n;
    if (n > 1 && per < E)
      return fail(h, OFRI_ERR_TOO_SMALL, "row-band mode: level %d has %d rows per rank, fewer than the %d-row exchange", l + 1,
                  per, E);
    b.own0 = rank * per;
    b.own1 = b.own0 + per;
    // ghost rows of this level: E sweeps between exchanges + the derivative and Gaussian rows, + the rows the warp may
    // reach on every level that is warped (all but the coarsest): the coarsest level's smaller frame saves a tile row of
    // its bands (8 ranks, 16384^2: 1092 instead of 1108 rows = 22 instead of 23 tile rows = 11 instead of 12 waves)
    // (rows the one-off stages spoil at a band edge: main pre-filter half-width + 1 derivative row, or optional
    // pre-filter half-width + 1 Liu-Shen coefficient row)
    int spoil = p->n_taps_main / 2 + 1;
    if (p->opt_algo.kind != OFRI_ALGO_NONE && p->n_taps_opt / 2 + 1 > spoil) spoil = p->n_taps_opt / 2 + 1;
    if (p->main_algo.kind == OFRI_ALGO_LS && spoil < 2) spoil = 2;
    b.G = l > 0 ? E + (spoil > 2 ? spoil : 2) + bp->Rw : E + spoil;
    b.ext0 = b.own0 - b.G < 0 ? 0 : b.own0 - b.G;
    b.ext1 = b.own1 + b.G > b.Hl ? b.Hl : b.own1 + b.G;
    long lo = (long)f * b.ext0 - (f > 1 ? 2L * f : 0), hi = (long)f * (b.ext1 - 1) + (f > 1 ? 3L * f : 1);
    if (lo < in0) in0 = lo;
    if (hi > in1) in1 = hi;
    scale *= 2.0;
  }
  bp->in0 = in0 < 0 ? 0 : (int)in0;
  bp->in1 = in1 > H ? H : (int)in1;
  return OFRI_OK;
}

struct BandWs {
  Img lvl1, lvl2, warp1, warp2, work1, work2, opt1, opt2, tmp, fx, fy, ft, U[2], V[2], U0, V0, Uacc, Vacc, us, vs, gU, gV;
  LsPlanes ls;
  ImgD M1, T1, M2, M1b, T1b, M2b;
  double* hs_acc = nullptr;
  double* ls_errs = nullptr;
  int* ls_state = nullptr;
  unsigned* ls_max = nullptr;
  int* flag = nullptr;
};
void plan_band_ws(Bump& b, const BandPlanInt& bp, int W, const ofri_params* p, BandWs* ws) {
  const BandLevel& f = bp.lv[bp.L - 1];
  const int rows = f.ext1 - f.ext0, in_rows = bp.in1 - bp.in0;
  const bool multi = bp.L > 1;
  const bool has_opt = p->opt_algo.kind != OFRI_ALGO_NONE;
  const bool has_ls = p->main_algo.kind == OFRI_ALGO_LS || p->opt_algo.kind == OFRI_ALGO_LS;
  const bool has_hs = p->main_algo.kind == OFRI_ALGO_HS || p->opt_algo.kind == OFRI_ALGO_HS;
  if (multi) {
    ws->lvl1 = b.plane(1, rows, W); ws->lvl2 = b.plane(1, rows, W);
    ws->warp1 = b.plane(1, rows, W); ws->warp2 = b.plane(1, rows, W);
    ws->tmp = b.plane(1, in_rows, W);
  }
  ws->work1 = b.plane(1, rows, W);
  ws->work2 = b.plane(1, rows, W);
  if (has_opt) { ws->opt1 = b.plane(1, rows, W); ws->opt2 = b.plane(1, rows, W); }
  if (has_hs) { ws->fx = b.plane(1, rows, W); ws->fy = b.plane(1, rows, W); ws->ft = b.plane(1, rows, W); }
  for (int i = 0; i < 2; ++i) { ws->U[i] = b.plane(1, rows, W); ws->V[i] = b.plane(1, rows, W); }
  ws->U0 = b.plane(1, rows, W);
  ws->V0 = b.plane(1, rows, W);
  ws->Uacc = b.plane(1, rows, W);
  ws->Vacc = b.plane(1, rows, W);
  if (multi) {
    ws->us = b.plane(1, rows, W);
    ws->vs = b.plane(1, rows, W);
    const BandLevel& c = bp.lv[bp.L - 2];     // the largest coarse level
    ws->gU = b.plane(1, c.Hl, c.Wl);
    ws->gV = b.plane(1, c.Hl, c.Wl);
    ws->M1 = b.planed(1, c.Hl, c.Wl);
    const int trows = rows > c.Hl ? rows : c.Hl;   // T1: legacy T1 (rows x w) or forward-elimination scratch (strip x w)
    ws->T1 = b.planed(1, trows, c.Wl);
    ws->M2 = b.planed(1, rows, c.Wl);
    ws->M1b = b.planed(1, c.Hl, c.Wl);
    ws->T1b = b.planed(1, trows, c.Wl);
    ws->M2b = b.planed(1, rows, c.Wl);
  }
  if (has_ls) {
    for (int c = 0; c < 8; ++c) ws->ls.c[c] = b.plane(1, rows, W);
    int maxit = 1;
    if (p->main_algo.kind == OFRI_ALGO_LS) maxit = p->main_algo.ls_maxiter;
    if (p->opt_algo.kind == OFRI_ALGO_LS && p->opt_algo.ls_maxiter > maxit) maxit = p->opt_algo.ls_maxiter;
    ws->ls_errs = (double*)b.take(sizeof(double) * 2 * (size_t)maxit);
    ws->ls_state = (int*)b.take(sizeof(int) * 4);
    ws->ls_max = (unsigned*)b.take(sizeof(unsigned) * 2);
  }
  ws->hs_acc = (double*)b.take(sizeof(double) * 2);
  ws->flag = (int*)b.take(sizeof(int) * 4);
}

// ghost rows of one (U, V) pair <- the neighbours' owned rows; o0 / o1 = LOCAL row range the band owns
int band_exchange_uv(ofri_handle h, const Img& U, const Img& V, int o0, int o1, int E, cudaStream_t stream = nullptr) {
  ofri::Comm* c = h->comm;
  if (!stream) stream = h->stream;
  if (!c || c->nranks == 1) return OFRI_OK;
  const long pitch = U.pitch;
  const bool up = c->rank > 0, dn = c->rank < c->nranks - 1;
  const float* su[2] = {U.p + (long)o0 * pitch, V.p + (long)o0 * pitch};
  float* ru[2] = {up ? U.p + (long)(o0 - E) * pitch : U.p, up ? V.p + (long)(o0 - E) * pitch : V.p};
  const float* sd[2] = {U.p + (long)(o1 - E) * pitch, V.p + (long)(o1 - E) * pitch};
  float* rd[2] = {dn ? U.p + (long)o1 * pitch : U.p, dn ? V.p + (long)o1 * pitch : V.p};
  if (c->exchange(2, su, ru, sd, rd, (size_t)E * pitch, stream))
    return fail(h, OFRI_ERR_COMM, "ghost-row exchange failed: %s", c->error());
  return OFRI_OK;
}

__global__ void band_reach_kernel(Img vs, float limit, int* flag) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= vs.W || y >= vs.H) return;
  float v = vs.p[(long)y * vs.pitch + x];
  if (!(fabsf(v) * 0.5f + 2.0f <= limit)) atomicExch(flag, 1);
}

int run_adapter_banded(ofri_handle h, const ofri_algo& a, int call_index, BandWs& ws, const Img& im1, const Img& im2,
                       const BandLevel& bl, int E, int cur, bool uv_zero, bool coarse_level, float* d_err,
                       int* comm_rc, bool finest = true) {
  cudaStream_t s = h->stream;
  ofri::Comm* c = h->comm;
  const int rows = bl.ext1 - bl.ext0, Wl = bl.Wl;
  const int o0 = bl.own0 - bl.ext0, o1 = bl.own1 - bl.ext0;
  const double npix = (double)bl.Hl * (double)bl.Wl;
  Img U[2] = {view(ws.U[0], rows, Wl), view(ws.U[1], rows, Wl)};
  Img V[2] = {view(ws.V[0], rows, Wl), view(ws.V[1], rows, Wl)};
  if (a.kind == OFRI_ALGO_HS) {
    Img fx = view(ws.fx, rows, Wl), fy = view(ws.fy, rows, Wl), ft = view(ws.ft, rows, Wl);
    {
      Timed t(h, "hs_derivs");
      launch_hs_derivs(im1, im2, fx, fy, ft, s, h->lc);
    }
    Img U0, V0;
    if (!uv_zero && d_err) {
      U0 = view(ws.U0, rows, Wl);
      V0 = view(ws.V0, rows, Wl);
      launch_copy(U0, U[cur], s, h->lc);
      launch_copy(V0, V[cur], s, h->lc);
    }
    int res;
    {
      const bool precise = h->hs_precise >= 2 || (h->hs_precise == 1 && coarse_level);
      Timed t(h, precise ? "hs_iterate_precise" : (finest ? "hs_iterate" : "hs_iterate_coarse"));
      const int niter = a.hs_niter;
      const int c0 = cur;
      // every E sweeps: the tiles producing the rows the neighbours need run first, the exchange of the buffer just
      // written then proceeds on the communication stream under the interior tiles (HsSplit, ofri_internal.h)
      HsSplit split;
      split.every = E;
      split.mid_lo = o0 + E;
      split.mid_hi = o1 - E;
      split.reserve_sms = (c && c->nranks > 1 && c->uses_sms()) ? h->band_reserve_sms : 0;
      split.begin = [&](int which) {
        if (*comm_rc || !c || c->nranks == 1) return;
        const int b = which ? (c0 ^ 1) : c0;
        cudaEventRecord(h->ev_c0, s);
        cudaStreamWaitEvent(h->s_comm, h->ev_c0, 0);
        *comm_rc = band_exchange_uv(h, U[b], V[b], o0, o1, E, h->s_comm);
        cudaEventRecord(h->ev_c1, h->s_comm);
      };
      split.end = [&]() {
        if (!c || c->nranks == 1) return;
        cudaStreamWaitEvent(s, h->ev_c1, 0);
      };
      // several ranks: a band's launches are short, so the fixed cost per launch (and per exchange / all-reduce) weighs
      // more than on one GPU -- fuse 8 sweeps per fast-arithmetic launch unless configured otherwise
      int fuse = h->hs_fuse;
      if (!precise && h->hs_fuse_fast > 0 && E % h->hs_fuse_fast == 0) fuse = h->hs_fuse_fast;
      else if (!precise && h->hs_fuse_fast == 0 && h->auto_fuse && fuse >= 1 && fuse < 8 && E % 8 == 0 &&
               ((c && c->nranks > 1) || big_frame(bl.Hl, bl.Wl)))
        fuse = 8;
      res = launch_hs_iterate(U[cur], V[cur], U[cur ^ 1], V[cur ^ 1], fx, fy, ft, a.alphas[call_index], niter, fuse,
                              h->hs_variant, precise, s, h->lc, HsHook(), &split);
    }
    cur = res ? (cur ^ 1) : cur;
    if (d_err) {
      Timed t(h, "hs_error");
      launch_hs_error_sums(U[cur], V[cur], U0, V0, ws.hs_acc, o0, o1, s, h->lc);
      if (c && c->nranks > 1 && c->allreduce_sum(ws.hs_acc, 2, s)) *comm_rc = fail(h, OFRI_ERR_COMM, "%s", c->error());
      launch_hs_error_finish(ws.hs_acc, d_err, 1, 1, npix, s, h->lc);
    }
    return cur;
  }
  // Liu-Shen: u = ROW component (our V), v = COLUMN component (our U)
  LsPlanes co;
  for (int k = 0; k < 8; ++k) co.c[k] = view(ws.ls.c[k], rows, Wl);
  {
    Timed t(h, "ls_coefficients");
    Img m1 = im1, m2 = im2;                            // owned rows only: ghost rows may hold band-edge artefacts
    m1.p += (size_t)o0 * m1.pitch; m1.H = o1 - o0;
    m2.p += (size_t)o0 * m2.pitch; m2.H = o1 - o0;
    launch_ls_max(m1, m2, ws.ls_max, s, h->lc);
    if (c && c->nranks > 1 && c->allreduce_max_u32(ws.ls_max, 2, s)) *comm_rc = fail(h, OFRI_ERR_COMM, "%s", c->error());
    launch_ls_coef(im1, im2, a.ls_h, co, ws.ls_max, s, h->lc, bl.ext0, bl.Hl);
  }
  {
    Timed t(h, "ls_iterate");
    LsBand band{o0, o1, npix};
    const int c0 = cur;
    // The ghost frame holds E valid rows after an exchange and every sweep spoils one more: refresh it only when the
    // next block would run out (the residual sums, in contrast, are needed by the very next launch's stopping rule).
    int ls_fuse = h->ls_fuse;
    if (h->auto_fuse && ls_fuse >= 1 && ls_fuse < 4 && ((c && c->nranks > 1) || big_frame(bl.Hl, bl.Wl)))
      ls_fuse = 4;                                                            // half the launches / all-reduces
    const int Tl = ls_fuse > 4 ? 4 : (ls_fuse < 1 ? 1 : ls_fuse);
    int spoiled = 0;
    LsHook hook = [&](int k0, int n, int written) {
      if (*comm_rc || !c || c->nranks == 1) return;
      if (n > 0 && c->allreduce_sum(ws.ls_errs + 2 * k0, 2 * (size_t)n, s)) {
        *comm_rc = fail(h, OFRI_ERR_COMM, "%s", c->error());
        return;
      }
      spoiled += n;
      if (written != 2 && spoiled + Tl <= E) return;
      spoiled = 0;
      if (written == 0 || written == 2) *comm_rc = band_exchange_uv(h, U[c0], V[c0], o0, o1, E);
      if (!*comm_rc && (written == 1 || written == 2)) *comm_rc = band_exchange_uv(h, U[c0 ^ 1], V[c0 ^ 1], o0, o1, E);
    };
    launch_ls_solve(V[cur], U[cur], V[cur ^ 1], U[cur ^ 1], co, a.ls_h, a.ls_maxiter, a.ls_tol, ls_fuse, h->ls_variant,
                    ws.ls_errs, ws.ls_state, V[cur], U[cur], d_err, 1, nullptr, s, h->lc, &band, hook);
    if (!*comm_rc) *comm_rc = band_exchange_uv(h, U[cur], V[cur], o0, o1, E);    // ghosts of the selected final state
  }
  return cur;
}

int run_pyramid_banded(ofri_handle h, const float* d_im1, const float* d_im2, int H, int W, const ofri_params* p,
                       const BandPlanInt& bp, float* d_u, float* d_v, float* d_err, BandWs& ws) {
  cudaStream_t s = h->stream;
  ofri::Comm* c = h->comm;
  const int L = bp.L, n = c ? c->nranks : 1;
  const bool has_opt = p->opt_algo.kind != OFRI_ALGO_NONE;
  const GaussTaps taps_main = make_taps(p->taps_main, p->n_taps_main);
  const GaussTaps taps_opt = make_taps(p->taps_opt, p->n_taps_opt);
  const int in_rows = bp.in1 - bp.in0;
  Img in1 = dense(d_im1, 1, in_rows, W), in2 = dense(d_im2, 1, in_rows, W);
  Img Uacc = ws.Uacc, Vacc = ws.Vacc, us = ws.us, vs = ws.vs;
  int cur = 0, comm_rc = 0;
  cudaMemsetAsync(ws.flag, 0, sizeof(int) * 4, s);
  for (int l = 0; l < L; ++l) {
    const BandLevel& bl = bp.lv[l];
    const bool last = l == L - 1;
    const bool local_scaling = last ? p->final_scaling != 0 : p->intermediate_scaling != 0;
    const int rows = bl.ext1 - bl.ext0, Wl = bl.Wl;
    const int o0 = bl.own0 - bl.ext0;
    Img n1, n2;
    if (last) {   // the frames themselves: a window of the input band
      n1 = dense(d_im1 + (size_t)(bl.ext0 - bp.in0) * W, 1, rows, W);
      n2 = dense(d_im2 + (size_t)(bl.ext0 - bp.in0) * W, 1, rows, W);
    } else {
      n1 = view(ws.lvl1, rows, Wl);
      n2 = view(ws.lvl2, rows, Wl);
      ResizeTaps tx, ty;
      int rc = get_resize_taps(h, W, Wl, &tx);
      if (rc) return rc;
      rc = get_resize_taps(h, H, bl.Hl, &ty);
      if (rc) return rc;
      Timed t(h, "resize");
      Img tmp = view(ws.tmp, in_rows, Wl);
      launch_resize(in1, tmp, n1, tx, ty, s, h->lc, bp.in0, bl.ext0);
      launch_resize(in2, tmp, n2, tx, ty, s, h->lc, bp.in0, bl.ext0);
    }
    Img w1 = n1, w2 = n2;
    Img Ucur = view(ws.U[cur], rows, Wl), Vcur = view(ws.V[cur], rows, Wl);
    if (l > 0) {
      const BandLevel& pl = bp.lv[l - 1];
      const int prow = pl.ext1 - pl.ext0;
      Img ua = view(Uacc, prow, pl.Wl), va = view(Vacc, prow, pl.Wl);
      Img un = view(us, rows, Wl), vn = view(vs, rows, Wl);
      SplineSys sy, sx;
      int rc = get_spline_sys(h, pl.Hl, &sy);
      if (rc) return rc;
      rc = get_spline_sys(h, pl.Wl, &sx);
      if (rc) return rc;
      const int per_c = pl.own1 - pl.own0, per_f = bl.own1 - bl.own0;
      const size_t own_cnt = (size_t)per_c * ua.pitch;
      const float* su = ua.p + (size_t)(pl.own0 - pl.ext0) * ua.pitch;
      const float* sv = va.p + (size_t)(pl.own0 - pl.ext0) * va.pitch;
      // The column-direction solve needs, for this band's output rows, only a WINDOW of every column: the band's own
      // coarse rows + `halo` rows of each neighbour (ofri_spline.cuh).  halo = the largest reach over all ranks, so
      // that the (symmetric) exchange moves the same count everywhere.  Bands thinner than the halo fall back to the
      // all-gather of the whole coarse plane.
      int halo = 0;
      for (int rr = 0; rr < n; ++rr) {
        const int e0 = rr * per_f - bl.G < 0 ? 0 : rr * per_f - bl.G;
        const int e1 = (rr + 1) * per_f + bl.G > bl.Hl ? bl.Hl : (rr + 1) * per_f + bl.G;
        int lo, hi;
        spline_rows_needed(e0, e1 - e0, pl.Hl, bl.Hl, sy, &lo, &hi);
        if (rr * per_c - lo > halo) halo = rr * per_c - lo;
        if (hi - (rr + 1) * per_c > halo) halo = hi - (rr + 1) * per_c;
      }
      const bool windowed = h->spline_variant != 0 && (n == 1 || halo <= per_c);
      int S0 = 0, S1 = pl.Hl;                       // coarse rows held by the strip gU / gV
      if (windowed && n > 1) {
        S0 = pl.own0 - halo < 0 ? 0 : pl.own0 - halo;
        S1 = pl.own1 + halo > pl.Hl ? pl.Hl : pl.own1 + halo;
      }
      Img gU = view(ws.gU, S1 - S0, pl.Wl), gV = view(ws.gV, S1 - S0, pl.Wl);
      {
        Timed t(h, "gather");
        if (c && n > 1 && windowed) {
          float* ou = gU.p + (size_t)(pl.own0 - S0) * gU.pitch;
          float* ov = gV.p + (size_t)(pl.own0 - S0) * gV.pitch;
          cudaMemcpyAsync(ou, su, own_cnt * sizeof(float), cudaMemcpyDeviceToDevice, s);
          cudaMemcpyAsync(ov, sv, own_cnt * sizeof(float), cudaMemcpyDeviceToDevice, s);
          const size_t hc = (size_t)halo * gU.pitch;
          const float* s_up[2] = {su, sv};
          float* r_up[2] = {gU.p, gV.p};                                  // rows [own0 - halo, own0) (ranks > 0: S0 = own0 - halo)
          const float* s_dn[2] = {su + own_cnt - hc, sv + own_cnt - hc};
          float* r_dn[2] = {ou + own_cnt, ov + own_cnt};                  // rows [own1, own1 + halo)
          if (c->exchange(2, s_up, r_up, s_dn, r_dn, hc, s))
            return fail(h, OFRI_ERR_COMM, "halo exchange of the coarse flow failed: %s", c->error());
        } else if (c && n > 1) {
          if (c->allgather(su, gU.p, own_cnt, s) || c->allgather(sv, gV.p, own_cnt, s))
            return fail(h, OFRI_ERR_COMM, "all-gather of the coarse flow failed: %s", c->error());
        } else {
          cudaMemcpyAsync(gU.p, su, own_cnt * sizeof(float), cudaMemcpyDeviceToDevice, s);
          cudaMemcpyAsync(gV.p, sv, own_cnt * sizeof(float), cudaMemcpyDeviceToDevice, s);
        }
      }
      {
        Timed t(h, "spline_upsample");
        float mx = 1.0f, my = 1.0f;
        if (local_scaling) {
          mx = (float)Wl / (float)pl.Wl;
          my = (float)bl.Hl / (float)pl.Hl;
        }
        cudaEventRecord(h->ev_fork, s);
        cudaStreamWaitEvent(h->s_aux, h->ev_fork, 0);
        bool ok = true;
        if (h->spline_variant == 0) {
          ImgD M1 = viewd(ws.M1, pl.Hl, pl.Wl), T1 = viewd(ws.T1, rows, pl.Wl), M2 = viewd(ws.M2, rows, pl.Wl);
          ImgD M1b = viewd(ws.M1b, pl.Hl, pl.Wl), T1b = viewd(ws.T1b, rows, pl.Wl), M2b = viewd(ws.M2b, rows, pl.Wl);
          launch_spline_seq(gU, un, mx, sy, sx, M1, T1, M2, s, h->lc, bl.ext0, bl.Hl);
          launch_spline_seq(gV, vn, my, sy, sx, M1b, T1b, M2b, h->s_aux, h->lc, bl.ext0, bl.Hl);
        } else {
          ImgD M1 = viewd(ws.M1, S1 - S0, pl.Wl), D1 = viewd(ws.T1, S1 - S0, pl.Wl);
          ImgD M1b = viewd(ws.M1b, S1 - S0, pl.Wl), D1b = viewd(ws.T1b, S1 - S0, pl.Wl);
          ok = launch_spline(gU, un, mx, sy, sx, M1, D1, s, h->lc, S0, pl.Hl, bl.ext0, bl.Hl) &&
               launch_spline(gV, vn, my, sy, sx, M1b, D1b, h->s_aux, h->lc, S0, pl.Hl, bl.ext0, bl.Hl);
        }
        cudaEventRecord(h->ev_join, h->s_aux);
        cudaStreamWaitEvent(s, h->ev_join, 0);
        if (!ok) return fail(h, OFRI_ERR_INVALID, "spline up-sample: the coarse strip does not cover the band");
      }
      {
        Timed t(h, "warp");
        dim3 b(32, 8), g((Wl + 31) / 32, (rows + 7) / 8, 1);
        band_reach_kernel<<<g, b, 0, s>>>(vn, (float)bp.Rw, ws.flag);
        h->lc.n += 1;
        w1 = view(ws.warp1, rows, Wl);
        w2 = view(ws.warp2, rows, Wl);
        launch_warp_pair(n1, n2, un, vn, w1, w2, s, h->lc, bl.ext0, bl.ext0, bl.Hl);
        std::swap(Uacc, us);
        std::swap(Vacc, vs);
      }
    }
    Img UaccL = view(Uacc, rows, Wl), VaccL = view(Vacc, rows, Wl);
    if (l == 0) {
      cudaMemsetAsync(UaccL.p, 0, sizeof(float) * (size_t)UaccL.stride, s);
      cudaMemsetAsync(VaccL.p, 0, sizeof(float) * (size_t)VaccL.stride, s);
    }
    cudaMemsetAsync(Ucur.p, 0, sizeof(float) * (size_t)Ucur.stride, s);
    cudaMemsetAsync(Vcur.p, 0, sizeof(float) * (size_t)Vcur.stride, s);
    Img work1 = w1, work2 = w2;
    if (p->n_taps_main > 0) {
      Timed t(h, "gauss");
      work1 = view(ws.work1, rows, Wl);
      work2 = view(ws.work2, rows, Wl);
      Img gt = view(ws.U0, rows, Wl);      // scratch of the generic two-pass filter (free until the adapters run)
      launch_gauss(w1, gt, work1, taps_main, s, h->lc);
      launch_gauss(w2, gt, work2, taps_main, s, h->lc);
    }
    Img o1 = n1, o2 = n2;
    if (has_opt && p->n_taps_opt > 0) {
      Timed t(h, "gauss");
      o1 = view(ws.opt1, rows, Wl);
      o2 = view(ws.opt2, rows, Wl);
      Img gt = view(ws.U0, rows, Wl);
      launch_gauss(n1, gt, o1, taps_opt, s, h->lc);
      launch_gauss(n2, gt, o2, taps_opt, s, h->lc);
    }
    float* e_main = d_err ? d_err + 2 * l : nullptr;
    cur = run_adapter_banded(h, p->main_algo, l, ws, work1, work2, bl, bp.E, cur, true, hs_needs_precise(p, !last, true), e_main,
                             &comm_rc, last);
    if (comm_rc) return comm_rc;
    if (has_opt) {
      float* e_opt = d_err ? d_err + 2 * l + 1 : nullptr;
      cur = run_adapter_banded(h, p->opt_algo, l, ws, o1, o2, bl, bp.E, cur, false, hs_needs_precise(p, !last, false), e_opt,
                               &comm_rc, last);
      if (comm_rc) return comm_rc;
    }
    Ucur = view(ws.U[cur], rows, Wl);
    Vcur = view(ws.V[cur], rows, Wl);
    {
      Timed t(h, "accumulate");
      launch_axpy(UaccL, Ucur, s, h->lc);
      launch_axpy(VaccL, Vcur, s, h->lc);
    }
    if (last) {
      Img src_u = UaccL, src_v = VaccL;
      src_u.p += (size_t)o0 * src_u.pitch;
      src_v.p += (size_t)o0 * src_v.pitch;
      src_u.H = src_v.H = bl.own1 - bl.own0;
      launch_copy(dense(d_u, 1, bl.own1 - bl.own0, W), src_u, s, h->lc);
      launch_copy(dense(d_v, 1, bl.own1 - bl.own0, W), src_v, s, h->lc);
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, OFRI_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  return OFRI_OK;
}

size_t band_workspace_bytes(const BandPlanInt& bp, int W, const ofri_params* p) {
  Bump dry(nullptr, 0, true);
  BandWs ws;
  plan_band_ws(dry, bp, W, p, &ws);
  return dry.off + 4096;
}

}  // namespace

// =====================================================================================================================
// extern "C"
// =====================================================================================================================
extern "C" {

int ofri_abi_version(void) { return OFRI_ABI_VERSION; }

int ofri_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    fail(nullptr, OFRI_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return OFRI_ERR_NO_DEVICE;
  }
  return n;
}

int ofri_create(int device, ofri_handle* out) {
  if (!out) return fail(nullptr, OFRI_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int n = ofri_device_count();
  if (n <= 0) return fail(nullptr, OFRI_ERR_NO_DEVICE, "no CUDA device available (libofri has no CPU fallback)");
  if (device < 0 || device >= n) return fail(nullptr, OFRI_ERR_INVALID, "device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    cudaGetLastError();
    return fail(nullptr, OFRI_ERR_NO_DEVICE, "cannot open CUDA device %d", device);
  }
  if (prop.major != 10)
    return fail(nullptr, OFRI_ERR_NO_DEVICE, "device %d is sm_%d%d; libofri is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
  ofri_ctx* h = new ofri_ctx();
  h->device = device;
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->s_aux, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->s_comm, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return fail(nullptr, OFRI_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming | cudaEventBlockingSync);
    cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming | cudaEventBlockingSync);
  }
  cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_c0, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_c1, cudaEventDisableTiming);
  h->stream = h->own_stream;
  *out = h;
  return OFRI_OK;
}

int ofri_destroy(ofri_handle h) {
  if (!h) return OFRI_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& kv : h->taps) { cudaFree(kv.second.xmin); cudaFree(kv.second.cnt); cudaFree(kv.second.w); }
  for (auto& kv : h->taps_bilinear) { cudaFree(kv.second.xmin); cudaFree(kv.second.cnt); cudaFree(kv.second.w); }
  for (auto& kv : h->splines) { cudaFree(kv.second.lo); cudaFree(kv.second.cp); cudaFree(kv.second.den); }
  if (h->arena) cudaFree(h->arena);
  if (h->stage) cudaFree(h->stage);
  if (h->hstage) cudaFreeHost(h->hstage);
  delete h->comm;
  for (int i = 0; i < 2; ++i) {
    cudaEventDestroy(h->ev_h2d[i]);
    cudaEventDestroy(h->ev_comp[i]);
    cudaEventDestroy(h->ev_d2h[i]);
  }
  cudaStreamDestroy(h->own_stream);
  cudaStreamDestroy(h->s_in);
  cudaStreamDestroy(h->s_out);
  cudaStreamDestroy(h->s_aux);
  cudaStreamDestroy(h->s_comm);
  cudaEventDestroy(h->ev_c0);
  cudaEventDestroy(h->ev_c1);
  cudaEventDestroy(h->ev_fork);
  cudaEventDestroy(h->ev_join);
  delete h;
  return OFRI_OK;
}

const char* ofri_last_error(ofri_handle h) {
  if (h) return h->err.c_str();
  std::lock_guard<std::mutex> g(g_err_mutex);
  static thread_local std::string copy;
  copy = g_last_error;
  return copy.c_str();
}

int ofri_set_stream(ofri_handle h, void* cuda_stream) {
  OFRI_ENTER(h);
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return OFRI_OK;
}

int ofri_synchronize(ofri_handle h) {
  OFRI_ENTER(h);
  OFRI_CUDA(h, cudaStreamSynchronize(h->stream));
  OFRI_CUDA(h, cudaStreamSynchronize(h->s_in));
  OFRI_CUDA(h, cudaStreamSynchronize(h->s_out));
  collect_times(h);     // stage timings of device-pointer calls become readable after a synchronise
  return OFRI_OK;
}

static int* option_slot(ofri_handle h, const char* key) {
  if (!key) return nullptr;
  if (!strcmp(key, "hs_fuse")) return &h->hs_fuse;
  if (!strcmp(key, "hs_variant")) return &h->hs_variant;
  if (!strcmp(key, "hs_precise")) return &h->hs_precise;
  if (!strcmp(key, "ls_fuse")) return &h->ls_fuse;
  if (!strcmp(key, "ls_variant")) return &h->ls_variant;
  if (!strcmp(key, "chunk_pairs")) return &h->chunk_pairs;
  if (!strcmp(key, "timing")) return &h->timing;
  if (!strcmp(key, "auto_fuse")) return &h->auto_fuse;
  if (!strcmp(key, "hs_fuse_fast")) return &h->hs_fuse_fast;
  if (!strcmp(key, "hs_fuse_precise")) return &h->hs_fuse_precise;
  if (!strcmp(key, "last_chunk_pairs")) return &h->last_chunk_pairs;
  if (!strcmp(key, "last_host_path")) return &h->last_host_path;
  if (!strcmp(key, "last_hs_fuse_fine")) return &h->last_hs_fuse_fine;
  if (!strcmp(key, "last_hs_fuse_coarse")) return &h->last_hs_fuse_coarse;
  if (!strcmp(key, "last_ls_fuse")) return &h->last_ls_fuse;
  if (!strcmp(key, "host_bounce")) return &h->host_bounce;
  if (!strcmp(key, "band_exchange")) return &h->band_exchange;
  if (!strcmp(key, "band_reach")) return &h->band_reach;
  if (!strcmp(key, "spline_variant")) return &h->spline_variant;
  if (!strcmp(key, "band_reserve_sms")) return &h->band_reserve_sms;
  return nullptr;
}
int ofri_set_option(ofri_handle h, const char* key, int value) {
  OFRI_ENTER(h);
  int* s = option_slot(h, key);
  if (!s) return fail(h, OFRI_ERR_INVALID, "unknown option '%s'", key ? key : "(null)");
  *s = value;
  return OFRI_OK;
}
int ofri_get_option(ofri_handle h, const char* key, int* value) {
  OFRI_ENTER(h);
  if (key && value && !strcmp(key, "comm_peer_allreduce")) {      // read-only: which all-reduce the communicator uses
    *value = h->comm ? h->comm->peer_allreduce() : 0;
    return OFRI_OK;
  }
  int* s = option_slot(h, key);
  if (!s || !value) return fail(h, OFRI_ERR_INVALID, "unknown option '%s'", key ? key : "(null)");
  *value = *s;
  return OFRI_OK;
}
int64_t ofri_launch_count(ofri_handle h) { return h ? h->lc.n : 0; }

int ofri_debug_phase_read(ofri_handle h, int family, unsigned long long* out8) {
  OFRI_ENTER(h);
  if (!out8 || family < 0 || family > 1) return fail(h, OFRI_ERR_INVALID, "bad arguments");
  OFRI_CUDA(h, cudaDeviceSynchronize());
  if (family == 0) ofri::hs_tma_phase_read(out8);
  else ofri::ls_tma_phase_read(out8);
  return OFRI_OK;
}

int ofri_stage_timings(ofri_handle h, const char** names, float* ms, int max_entries) {
  if (!h) return 0;
  int n = 0;
  for (auto& kv : h->times_ms) {
    if (n >= max_entries) break;
    if (names) names[n] = kv.first.c_str();
    if (ms) ms[n] = kv.second;
    ++n;
  }
  return n;
}

// ---- whole path -------------------------------------------------------------------------------------------------------
int ofri_pyramidal_flow_dev(ofri_handle h, const float* d_im1, const float* d_im2, int batch, int H, int W,
                            const ofri_params* p, float* d_u_out, float* d_v_out, float* d_err_out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!d_im1 || !d_im2 || !d_u_out || !d_v_out) return fail(h, OFRI_ERR_INVALID, "NULL image / output pointer");
  int rc = check_params(h, p, H, W);
  if (rc) return rc;
  const int chunk = pick_chunk(h, batch, H, W, p, 0);
  rc = arena_reserve(h, workspace_bytes(chunk, H, W, p));
  if (rc) return rc;
  const size_t plane = (size_t)H * W;
  const int err_stride = p->pyramid_levels * p->k_levels * 2;
  for (int b0 = 0; b0 < batch; b0 += chunk) {
    int nb = batch - b0 < chunk ? batch - b0 : chunk;
    Bump bump(h->arena, h->arena_cap, false);
    Workspace ws;
    plan_workspace(bump, nb, H, W, p, &ws);
    rc = run_pyramid(h, dense(d_im1 + b0 * plane, nb, H, W), dense(d_im2 + b0 * plane, nb, H, W), p,
                     dense(d_u_out + b0 * plane, nb, H, W), dense(d_v_out + b0 * plane, nb, H, W),
                     d_err_out ? d_err_out + (size_t)b0 * err_stride : nullptr, ws);
    if (rc) return rc;
    if ((rc = check_lsw_flag(h, p, ws))) return rc;
  }
  return OFRI_OK;
}

int ofri_pyramidal_flow(ofri_handle h, const float* im1, const float* im2, int batch, int H, int W,
                        const ofri_params* p, float* u_out, float* v_out, float* err_out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!im1 || !im2 || !u_out || !v_out) return fail(h, OFRI_ERR_INVALID, "NULL image / output pointer");
  int rc = check_params(h, p, H, W);
  if (rc) return rc;
  const size_t plane = (size_t)H * W, plane_b = plane * sizeof(float);
  const int err_stride = p->pyramid_levels * p->k_levels * 2;
  const size_t err_b = sizeof(float) * err_stride;
  // per pair staging: 2 inputs + 2 outputs + errors, double buffered
  const size_t slot_per_pair = 4 * plane_b + ((err_b + 255) & ~(size_t)255);
  // Pageable caller buffers (what every numpy caller of the drop-in API passes): cudaMemcpyAsync on them is synchronous
  // and staged by the driver, so chunk c's copies would serialise with the enqueueing of chunk c+1.  They go through a
  // pinned bounce ring instead: one host thread fills the ring's input slots from the caller's frames, this thread
  // issues asynchronous copies ring <-> device around the compute, a second host thread empties the output slots into
  // the caller's arrays -- both memcpy streams overlap the GPU work.  Pinned caller buffers (and calls of one chunk,
  // where nothing can overlap) keep the direct copies.
  const bool pinned = is_pinned(im1) && is_pinned(im2) && is_pinned(u_out) && is_pinned(v_out) &&
                      (!err_out || is_pinned(err_out));
  // smaller chunks than the device-pointer path: the first H2D and the last D2H of a call are not overlapped
  int chunk = pick_chunk(h, batch, H, W, p, 2 * slot_per_pair, pinned ? 32 : 16);
  const bool bounce = !pinned && h->host_bounce && batch > chunk;
  if (!pinned && !bounce) chunk = pick_chunk(h, batch, H, W, p, 2 * slot_per_pair, 32);
  h->last_host_path = bounce ? 2 : 1;
  rc = arena_reserve(h, workspace_bytes(chunk, H, W, p));
  if (rc) return rc;
  const size_t slot_b = ((slot_per_pair * chunk) + 255) & ~(size_t)255;
  rc = ensure_stage(h, 2 * slot_b);
  if (rc) return rc;
  if (bounce && (rc = ensure_hstage(h, 2 * slot_b))) return rc;
  cudaStream_t s = h->stream;
  const int nchunks = (batch + chunk - 1) / chunk;   // (quarter-size first / last chunks were measured: no gain, profiles/)

  // ---- helper threads of the bounce path -----------------------------------------------------------------------------
  struct Pipe {
    std::mutex m;
    std::condition_variable cv;
    int loaded = 0, issued = 0, unloaded = 0;   // chunks whose input is in the ring / whose copies are enqueued / whose output reached the caller
    bool abort = false;
  } pipe;
  auto wait_for = [&](int Pipe::*field, int value) {   // false = aborted
    std::unique_lock<std::mutex> lk(pipe.m);
    pipe.cv.wait(lk, [&] { return pipe.abort || pipe.*field >= value; });
    return !pipe.abort;
  };
  auto bump = [&](int Pipe::*field) {
    { std::lock_guard<std::mutex> lk(pipe.m); pipe.*field += 1; }
    pipe.cv.notify_all();
  };
  auto slot_ptrs = [&](char* base, float** i1, float** i2, float** u, float** v, float** e) {
    *i1 = (float*)base;
    *i2 = *i1 + plane * chunk;
    *u = *i2 + plane * chunk;
    *v = *u + plane * chunk;
    *e = *v + plane * chunk;
  };
  std::thread loader, unloader;
  const int dev = h->device;
  if (bounce) {
    loader = std::thread([&] {
      cudaSetDevice(dev);
      for (int c = 0; c < nchunks; ++c) {
        const int b0 = c * chunk, nb = batch - b0 < chunk ? batch - b0 : chunk, slot = c & 1;
        if (c >= 2) {                        // the H2D of chunk c-2 must have drained this input slot
          if (!wait_for(&Pipe::issued, c - 1)) return;
          cudaEventSynchronize(h->ev_h2d[slot]);
        }
        float *i1, *i2, *u, *v, *e;
        slot_ptrs(h->hstage + slot * slot_b, &i1, &i2, &u, &v, &e);
        memcpy(i1, im1 + b0 * plane, plane_b * nb);
        memcpy(i2, im2 + b0 * plane, plane_b * nb);
        bump(&Pipe::loaded);
      }
    });
    unloader = std::thread([&] {
      cudaSetDevice(dev);
      for (int c = 0; c < nchunks; ++c) {
        const int b0 = c * chunk, nb = batch - b0 < chunk ? batch - b0 : chunk, slot = c & 1;
        if (!wait_for(&Pipe::issued, c + 1)) return;
        // ev_d2h[slot] was recorded for chunk c and is not re-recorded before chunk c+2, which waits for `unloaded`
        cudaEventSynchronize(h->ev_d2h[slot]);
        float *i1, *i2, *u, *v, *e;
        slot_ptrs(h->hstage + slot * slot_b, &i1, &i2, &u, &v, &e);
        memcpy(u_out + b0 * plane, u, plane_b * nb);
        memcpy(v_out + b0 * plane, v, plane_b * nb);
        if (err_out) memcpy(err_out + (size_t)b0 * err_stride, e, err_b * nb);
        bump(&Pipe::unloaded);
      }
    });
  }
  auto stop_threads = [&](bool aborting) {
    if (!bounce) return;
    if (aborting) {
      { std::lock_guard<std::mutex> lk(pipe.m); pipe.abort = true; }
      pipe.cv.notify_all();
    }
    if (loader.joinable()) loader.join();
    if (unloader.joinable()) unloader.join();
  };
  auto bail = [&](int code) {
    drain_streams(h, code);
    stop_threads(true);
    return code;
  };
#define OFRI_CUDA_BAIL(call)                                                                                 \
  do {                                                                                                       \
    cudaError_t e_ = (call);                                                                                 \
    if (e_ != cudaSuccess)                                                                                   \
      return bail(fail(h, e_ == cudaErrorMemoryAllocation ? OFRI_ERR_OOM : OFRI_ERR_CUDA, "%s failed: %s", #call, \
                       cudaGetErrorString(e_)));                                                             \
  } while (0)

  for (int c = 0; c < nchunks; ++c) {
    const int b0 = c * chunk;
    const int nb = batch - b0 < chunk ? batch - b0 : chunk;
    const int slot = c & 1;
    float *d_i1, *d_i2, *d_u, *d_v, *d_e;
    slot_ptrs(h->stage + slot * slot_b, &d_i1, &d_i2, &d_u, &d_v, &d_e);
    float *h_i1 = nullptr, *h_i2 = nullptr, *h_u = nullptr, *h_v = nullptr, *h_e = nullptr;
    if (bounce) {
      slot_ptrs(h->hstage + slot * slot_b, &h_i1, &h_i2, &h_u, &h_v, &h_e);
      if (!wait_for(&Pipe::loaded, c + 1)) return bail(fail(h, OFRI_ERR_CUDA, "host staging aborted"));
    }
    // H2D on the copy-in stream once the previous user of this slot (compute of chunk c-2) is done
    if (c >= 2) OFRI_CUDA_BAIL(cudaStreamWaitEvent(h->s_in, h->ev_comp[slot], 0));
    OFRI_CUDA_BAIL(cudaMemcpyAsync(d_i1, bounce ? h_i1 : im1 + b0 * plane, plane_b * nb, cudaMemcpyHostToDevice, h->s_in));
    OFRI_CUDA_BAIL(cudaMemcpyAsync(d_i2, bounce ? h_i2 : im2 + b0 * plane, plane_b * nb, cudaMemcpyHostToDevice, h->s_in));
    OFRI_CUDA_BAIL(cudaEventRecord(h->ev_h2d[slot], h->s_in));
    OFRI_CUDA_BAIL(cudaStreamWaitEvent(s, h->ev_h2d[slot], 0));
    if (c >= 2) OFRI_CUDA_BAIL(cudaStreamWaitEvent(s, h->ev_d2h[slot], 0));   // output slot drained
    Bump bump_alloc(h->arena, h->arena_cap, false);
    Workspace ws;
    plan_workspace(bump_alloc, nb, H, W, p, &ws);
    rc = run_pyramid(h, dense(d_i1, nb, H, W), dense(d_i2, nb, H, W), p, dense(d_u, nb, H, W), dense(d_v, nb, H, W),
                     err_out ? d_e : nullptr, ws);
    if (!rc) rc = check_lsw_flag(h, p, ws);
    if (rc) return bail(rc);
    OFRI_CUDA_BAIL(cudaEventRecord(h->ev_comp[slot], s));
    OFRI_CUDA_BAIL(cudaStreamWaitEvent(h->s_out, h->ev_comp[slot], 0));
    // the ring's output slot is free once the unloader has copied chunk c-2 out of it
    if (bounce && c >= 2 && !wait_for(&Pipe::unloaded, c - 1)) return bail(fail(h, OFRI_ERR_CUDA, "host staging aborted"));
    OFRI_CUDA_BAIL(cudaMemcpyAsync(bounce ? h_u : u_out + b0 * plane, d_u, plane_b * nb, cudaMemcpyDeviceToHost, h->s_out));
    OFRI_CUDA_BAIL(cudaMemcpyAsync(bounce ? h_v : v_out + b0 * plane, d_v, plane_b * nb, cudaMemcpyDeviceToHost, h->s_out));
    if (err_out)
      OFRI_CUDA_BAIL(cudaMemcpyAsync(bounce ? h_e : err_out + (size_t)b0 * err_stride, d_e, err_b * nb,
                                     cudaMemcpyDeviceToHost, h->s_out));
    OFRI_CUDA_BAIL(cudaEventRecord(h->ev_d2h[slot], h->s_out));
    if (bounce) bump(&Pipe::issued);
  }
#undef OFRI_CUDA_BAIL
  stop_threads(false);                       // the unloader returns after the last chunk reached the caller's arrays
  OFRI_CUDA(h, cudaStreamSynchronize(h->s_out));
  rc = finish(h);
  return rc;
}

int ofri_resize_bilinear(ofri_handle h, const float* in, int batch, int H, int W, int out_h, int out_w, float* out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!in || !out || out_h < 1 || out_w < 1) return fail(h, OFRI_ERR_INVALID, "bad resize arguments");
  const int mh = out_h > H ? out_h : H, mw = out_w > W ? out_w : W;
  int rc = arena_reserve(h, (sizeof(float) * (size_t)round_up(mw, 4) * mh * batch + 256) * 3 + 4096);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img i = b.plane(batch, H, W), t = b.plane(batch, H, out_w), o = b.plane(batch, out_h, out_w);
  ResizeTaps tx, ty;
  if ((rc = get_resize_taps(h, W, out_w, &tx, true)) || (rc = get_resize_taps(h, H, out_h, &ty, true))) return rc;
  if ((rc = upload(h, i, in))) return rc;
  launch_resize(i, t, o, tx, ty, h->stream, h->lc);
  if ((rc = download(h, out, o))) return rc;
  return finish(h);
}

static int check_lk(ofri_handle h, const ofri_lk_params* lp) {
  if (!lp) return fail(h, OFRI_ERR_INVALID, "Lucas-Kanade parameters are NULL");
  if (lp->size != sizeof(ofri_lk_params))
    return fail(h, OFRI_ERR_INVALID, "ofri_lk_params.size = %u, library expects %zu (ABI mismatch)", lp->size,
                sizeof(ofri_lk_params));
  if (lp->n_iters < 0 || lp->half_window < 0 || lp->half_window > 4096)
    return fail(h, OFRI_ERR_INVALID, "bad Lucas-Kanade iteration count / window");
  for (int i = 0; i < 4; ++i)
    if (lp->asym[i] != 0 && lp->asym[i] != 1) return fail(h, OFRI_ERR_INVALID, "asymmetric-window switches must be 0 or 1");
  return OFRI_OK;
}
int ofri_set_lk(ofri_handle h, const ofri_lk_params* lp) {
  OFRI_ENTER(h);
  int rc = check_lk(h, lp);
  if (rc) return rc;
  h->lk = *lp;
  h->lk_set = true;
  return OFRI_OK;
}
int ofri_lk_compute(ofri_handle h, const float* im1, const float* im2, const float* u0, const float* v0, int batch, int H,
                    int W, const ofri_lk_params* lp, float* u_out, float* v_out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!im1 || !im2 || !u_out || !v_out) return fail(h, OFRI_ERR_INVALID, "NULL pointer");
  int rc = check_lk(h, lp);
  if (rc) return rc;
  rc = arena_reserve(h, (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 4 + 4096);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img i1 = b.plane(batch, H, W), i2 = b.plane(batch, H, W), U = b.plane(batch, H, W), V = b.plane(batch, H, W);
  if ((rc = upload(h, i1, im1)) || (rc = upload(h, i2, im2))) return rc;
  if (!u0 || !v0) {
    cudaMemsetAsync(U.p, 0, sizeof(float) * U.stride * batch, h->stream);
    cudaMemsetAsync(V.p, 0, sizeof(float) * V.stride * batch, h->stream);
  } else if ((rc = upload(h, U, u0)) || (rc = upload(h, V, v0))) {
    return rc;
  }
  rc = launch_lk(i1, i2, U, V, lp, h->stream, h->lc);
  if (rc) return fail(h, rc, "Lucas-Kanade adapter failed");
  if ((rc = download(h, u_out, U)) || (rc = download(h, v_out, V))) return rc;
  return finish(h);
}

static int check_fb(ofri_handle h, const ofri_farneback_params* fp) {
  if (!fp) return fail(h, OFRI_ERR_INVALID, "Farneback parameters are NULL");
  if (fp->size != sizeof(ofri_farneback_params))
    return fail(h, OFRI_ERR_INVALID, "ofri_farneback_params.size = %u, library expects %zu (ABI mismatch)", fp->size,
                sizeof(ofri_farneback_params));
  if (!(fp->window_size & 1)) return fail(h, OFRI_ERR_INVALID, "windowSize must be an odd value");      // FB:97-98
  if (fp->window_size < 1 || fp->window_size / 2 > OFRI_FB_MAX_HALF) return fail(h, OFRI_ERR_INVALID, "windowSize out of range");
  if (fp->poly_n != 5 && fp->poly_n != 7) return fail(h, OFRI_ERR_INVALID, "polyN must be 5 or 7");    // FB:463
  if (!(fp->pyr_scale > 0.0f && fp->pyr_scale < 1.0f)) return fail(h, OFRI_ERR_INVALID, "pyrScale must be in (0, 1)");
  if (fp->n_iters < 0 || fp->extra_levels < 0 || fp->extra_levels >= OFRI_FB_MAX_LEVELS)
    return fail(h, OFRI_ERR_INVALID, "bad Farneback iteration / level count");
  for (int k = 0; k <= fp->extra_levels; ++k)
    if (fp->n_blur[k] < 0 || fp->n_blur[k] > OFRI_FB_MAX_HALF) return fail(h, OFRI_ERR_INVALID, "pre-blur kernel too long");
  return OFRI_OK;
}
int ofri_set_farneback(ofri_handle h, const ofri_farneback_params* fp) {
  OFRI_ENTER(h);
  int rc = check_fb(h, fp);
  if (rc) return rc;
  h->fb = *fp;
  h->fb_set = true;
  return OFRI_OK;
}
int ofri_farneback_compute(ofri_handle h, const float* im1, const float* im2, const float* u0, const float* v0, int batch,
                           int H, int W, const ofri_farneback_params* fp, float* u_out, float* v_out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!im1 || !im2 || !u_out || !v_out) return fail(h, OFRI_ERR_INVALID, "NULL pointer");
  int rc = check_fb(h, fp);
  if (rc) return rc;
  rc = arena_reserve(h, (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 40 + 8192);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img i1 = b.plane(batch, H, W), i2 = b.plane(batch, H, W), U = b.plane(batch, H, W), V = b.plane(batch, H, W);
  FbWorkspace ws;
  plan_fb_workspace(b, batch, H, W, &ws);
  if ((rc = upload(h, i1, im1)) || (rc = upload(h, i2, im2))) return rc;
  if (!u0 || !v0) {
    cudaMemsetAsync(U.p, 0, sizeof(float) * U.stride * batch, h->stream);
    cudaMemsetAsync(V.p, 0, sizeof(float) * V.stride * batch, h->stream);
  } else if ((rc = upload(h, U, u0)) || (rc = upload(h, V, v0))) {
    return rc;
  }
  rc = launch_farneback(i1, i2, U, V, fp, ws, [h](int in, int out, ResizeTaps* t) { return get_resize_taps(h, in, out, t, true); },
                        h->stream, h->lc);
  if (rc) return fail(h, rc, "Farneback adapter failed");
  if ((rc = download(h, u_out, U)) || (rc = download(h, v_out, V))) return rc;
  return finish(h);
}

int ofri_host_alloc(ofri_handle h, size_t bytes, void** out) {
  OFRI_ENTER(h);
  if (!out || bytes == 0) return fail(h, OFRI_ERR_INVALID, "bad host allocation request");
  *out = nullptr;
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(h, OFRI_ERR_OOM, "cudaHostAlloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
  }
  return OFRI_OK;
}
int ofri_host_free(ofri_handle h, void* p) {
  OFRI_ENTER(h);
  if (p) OFRI_CUDA(h, cudaFreeHost(p));
  return OFRI_OK;
}

int ofri_pyramidal_flow_external(ofri_handle h, const float* im1, const float* im2, int H, int W, const ofri_params* p,
                                 ofri_adapter_fn fn, void* user, float* u_out, float* v_out, float* err_out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, 1, H, W);
  if (!im1 || !im2 || !u_out || !v_out || !fn) return fail(h, OFRI_ERR_INVALID, "NULL image / output / callback pointer");
  h->ext_fn = fn;
  h->ext_user = user;
  h->ext_params = p;
  struct Reset { ofri_handle h; ~Reset() { h->ext_fn = nullptr; h->ext_user = nullptr; h->ext_params = nullptr; } } reset{h};
  int rc = check_params(h, p, H, W);
  if (rc) return rc;
  rc = arena_reserve(h, workspace_bytes(1, H, W, p) + 4 * (sizeof(float) * (size_t)round_up(W, 4) * H + 256) + 4096);
  if (rc) return rc;
  const int err_stride = p->pyramid_levels * p->k_levels * 2;
  Bump bump(h->arena, h->arena_cap, false);
  Workspace ws;
  plan_workspace(bump, 1, H, W, p, &ws);
  Img i1 = bump.plane(1, H, W), i2 = bump.plane(1, H, W), uo = bump.plane(1, H, W), vo = bump.plane(1, H, W);
  float* d_err = err_out ? (float*)bump.take(sizeof(float) * err_stride) : nullptr;
  if (d_err) cudaMemsetAsync(d_err, 0, sizeof(float) * err_stride, h->stream);
  if ((rc = upload(h, i1, im1)) || (rc = upload(h, i2, im2))) return rc;
  rc = run_pyramid(h, i1, i2, p, uo, vo, d_err, ws);
  if (!rc) rc = check_lsw_flag(h, p, ws);
  if (rc) {
    cudaStreamSynchronize(h->stream);
    cudaGetLastError();
    return rc;
  }
  if ((rc = download(h, u_out, uo)) || (rc = download(h, v_out, vo))) return rc;
  if (err_out) OFRI_CUDA(h, cudaMemcpyAsync(err_out, d_err, sizeof(float) * err_stride, cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

// ---- adapters stand-alone ------------------------------------------------------------------------------------------------
int ofri_hs_compute(ofri_handle h, const float* im1, const float* im2, const float* u0, const float* v0, int batch,
                    int H, int W, float alpha, int niter, float* u_out, float* v_out, float* err) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!im1 || !im2 || !u_out || !v_out) return fail(h, OFRI_ERR_INVALID, "NULL pointer");
  if (niter < 0) return fail(h, OFRI_ERR_INVALID, "Niter < 0");
  size_t need = (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 11 + sizeof(double) * 2 * batch +
                sizeof(float) * batch + 4096;
  int rc = arena_reserve(h, need);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img i1 = b.plane(batch, H, W), i2 = b.plane(batch, H, W), fx = b.plane(batch, H, W), fy = b.plane(batch, H, W),
      ft = b.plane(batch, H, W);
  Img U[2] = {b.plane(batch, H, W), b.plane(batch, H, W)}, V[2] = {b.plane(batch, H, W), b.plane(batch, H, W)};
  Img U0 = b.plane(batch, H, W), V0 = b.plane(batch, H, W);
  double* acc = (double*)b.take(sizeof(double) * 2 * batch);
  float* d_err = (float*)b.take(sizeof(float) * batch);
  if ((rc = upload(h, i1, im1)) || (rc = upload(h, i2, im2))) return rc;
  const bool zero0 = !u0 || !v0;
  if (zero0) {
    cudaMemsetAsync(U[0].p, 0, sizeof(float) * U[0].stride * batch, h->stream);
    cudaMemsetAsync(V[0].p, 0, sizeof(float) * V[0].stride * batch, h->stream);
  } else {
    if ((rc = upload(h, U[0], u0)) || (rc = upload(h, V[0], v0)) || (rc = upload(h, U0, u0)) ||
        (rc = upload(h, V0, v0)))
      return rc;
  }
  launch_hs_derivs(i1, i2, fx, fy, ft, h->stream, h->lc);
  int res = launch_hs_iterate(U[0], V[0], U[1], V[1], fx, fy, ft, alpha, niter, eff_hs_fuse(h, H, W, batch),
                              h->hs_variant, h->hs_precise >= 2, h->stream, h->lc);
  Img nu, nv;
  launch_hs_error(U[res], V[res], zero0 ? nu : U0, zero0 ? nv : V0, acc, d_err, 1, h->stream, h->lc);
  if ((rc = download(h, u_out, U[res])) || (rc = download(h, v_out, V[res]))) return rc;
  if (err) OFRI_CUDA(h, cudaMemcpyAsync(err, d_err, sizeof(float) * batch, cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

int ofri_ls_compute(ofri_handle h, const float* im1, const float* im2, const float* u0, const float* v0, int batch,
                    int H, int W, float hpar, int maxiter, double tol, float* u_out, float* v_out, float* err,
                    int32_t* iters) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!im1 || !im2 || !u_out || !v_out) return fail(h, OFRI_ERR_INVALID, "NULL pointer");
  if (maxiter < 1 || maxiter > 100000) return fail(h, OFRI_ERR_INVALID, "bad maxiter");
  size_t need = (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 14 +
                sizeof(double) * 2 * (size_t)maxiter * batch + 64 * (size_t)batch + 8192;
  int rc = arena_reserve(h, need);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img i1 = b.plane(batch, H, W), i2 = b.plane(batch, H, W);
  Img U[2] = {b.plane(batch, H, W), b.plane(batch, H, W)}, V[2] = {b.plane(batch, H, W), b.plane(batch, H, W)};
  LsPlanes co;
  for (int c = 0; c < 8; ++c) co.c[c] = b.plane(batch, H, W);
  double* errs = (double*)b.take(sizeof(double) * 2 * (size_t)maxiter * batch);
  int* state = (int*)b.take(sizeof(int) * 4 * batch);
  unsigned* mx = (unsigned*)b.take(sizeof(unsigned) * 2 * batch);
  float* d_err = (float*)b.take(sizeof(float) * batch);
  int* d_it = (int*)b.take(sizeof(int) * batch);
  if ((rc = upload(h, i1, im1)) || (rc = upload(h, i2, im2))) return rc;
  if (!u0 || !v0) {
    cudaMemsetAsync(U[0].p, 0, sizeof(float) * U[0].stride * batch, h->stream);
    cudaMemsetAsync(V[0].p, 0, sizeof(float) * V[0].stride * batch, h->stream);
  } else if ((rc = upload(h, U[0], u0)) || (rc = upload(h, V[0], v0))) {
    return rc;
  }
  launch_ls_coefficients(i1, i2, hpar, co, mx, h->stream, h->lc);
  launch_ls_solve(V[0], U[0], V[1], U[1], co, hpar, maxiter, tol, eff_ls_fuse(h, H, W, batch), h->ls_variant, errs, state, V[0], U[0],
                  d_err, 1, d_it,
                  h->stream, h->lc);
  if ((rc = download(h, u_out, U[0])) || (rc = download(h, v_out, V[0]))) return rc;
  if (err) OFRI_CUDA(h, cudaMemcpyAsync(err, d_err, sizeof(float) * batch, cudaMemcpyDeviceToHost, h->stream));
  if (iters) OFRI_CUDA(h, cudaMemcpyAsync(iters, d_it, sizeof(int) * batch, cudaMemcpyDeviceToHost, h->stream));
  return finish(h);
}

// ---- stages ------------------------------------------------------------------------------------------------------------------
int ofri_gaussian_taps(double sigma, int n_taps, float* taps_out) {
  if (!taps_out || n_taps < 1 || n_taps > OFRI_MAX_GAUSS_TAPS || !(n_taps & 1)) return OFRI_ERR_INVALID;
  // gaussian_filter.py:47-52: f64 sample -> f32 store; sum and divide in f32
  const int hk = n_taps / 2;
  float sum = 0.0f;
  for (int i = 0; i < n_taps; ++i) {
    double x = (double)(i - hk);
    double v = 1.0 / std::sqrt(2.0 * M_PI * sigma * sigma) * std::exp(-(x * x) / (2.0 * sigma * sigma));
    taps_out[i] = (float)v;
  }
  // numpy's pairwise sum equals a left-to-right sum for fewer than 8 elements; for longer kernels it differs in the
  // last bit at most -- callers that need bit-identical taps (the Python shim) pass numpy-generated coefficients.
  for (int i = 0; i < n_taps; ++i) { volatile float t = sum + taps_out[i]; sum = t; }
  for (int i = 0; i < n_taps; ++i) { volatile float t = taps_out[i] / sum; taps_out[i] = t; }
  return OFRI_OK;
}

int ofri_level_size(int n, double scale) { return level_size(n, scale); }

int ofri_gauss_px(ofri_handle h, const float* in, int batch, int H, int W, const float* taps, int n_taps, float* out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!in || !out || !taps || n_taps < 1 || n_taps > OFRI_MAX_GAUSS_TAPS || !(n_taps & 1))
    return fail(h, OFRI_ERR_INVALID, "bad gauss arguments");
  if (H < n_taps / 2 || W < n_taps / 2) return fail(h, OFRI_ERR_TOO_SMALL, "image smaller than the kernel half-width");
  int rc = arena_reserve(h, (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 3 + 4096);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img i = b.plane(batch, H, W), t = b.plane(batch, H, W), o = b.plane(batch, H, W);
  if ((rc = upload(h, i, in))) return rc;
  launch_gauss(i, t, o, make_taps(taps, n_taps), h->stream, h->lc);
  if ((rc = download(h, out, o))) return rc;
  return finish(h);
}

int ofri_resize_bicubic(ofri_handle h, const float* in, int batch, int H, int W, int out_h, int out_w, float* out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!in || !out || out_h < 1 || out_w < 1) return fail(h, OFRI_ERR_INVALID, "bad resize arguments");
  int rc = arena_reserve(h, (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 3 + 4096);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  if (out_w > W || out_h > H) return fail(h, OFRI_ERR_UNSUPPORTED, "only down-sampling is on the path");
  Img i = b.plane(batch, H, W), t = b.plane(batch, H, out_w), o = b.plane(batch, out_h, out_w);
  ResizeTaps tx, ty;
  if ((rc = get_resize_taps(h, W, out_w, &tx)) || (rc = get_resize_taps(h, H, out_h, &ty))) return rc;
  if ((rc = upload(h, i, in))) return rc;
  launch_resize(i, t, o, tx, ty, h->stream, h->lc);
  if ((rc = download(h, out, o))) return rc;
  return finish(h);
}

int ofri_spline_upsample(ofri_handle h, const float* in, int batch, int in_h, int in_w, int out_h, int out_w, float mul,
                         float* out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, in_h, in_w);
  if (!in || !out || out_h < 1 || out_w < 1) return fail(h, OFRI_ERR_INVALID, "bad spline arguments");
  if (in_h < 4 || in_w < 4) return fail(h, OFRI_ERR_TOO_SMALL, "the cubic spline needs >= 4 samples per axis");
  size_t need = sizeof(float) * ((size_t)round_up(in_w, 4) * in_h + (size_t)round_up(out_w, 4) * out_h) * batch +
                sizeof(double) * ((size_t)in_h * in_w + 2 * (size_t)(out_h > in_h ? out_h : in_h) * in_w) * batch + 8192;
  int rc = arena_reserve(h, need);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img i = b.plane(batch, in_h, in_w), o = b.plane(batch, out_h, out_w);
  const int th = out_h > in_h ? out_h : in_h;
  ImgD M1 = b.planed(batch, in_h, in_w), T1 = b.planed(batch, th, in_w), M2 = b.planed(batch, th, in_w);
  SplineSys sy, sx;
  if ((rc = get_spline_sys(h, in_h, &sy)) || (rc = get_spline_sys(h, in_w, &sx))) return rc;
  if ((rc = upload(h, i, in))) return rc;
  if (h->spline_variant == 0)
    launch_spline_seq(i, o, mul, sy, sx, M1, viewd(T1, out_h, in_w), viewd(M2, out_h, in_w), h->stream, h->lc);
  else if (!launch_spline(i, o, mul, sy, sx, M1, viewd(T1, in_h, in_w), h->stream, h->lc))
    return fail(h, OFRI_ERR_CUDA, "spline up-sample launch failed");
  if ((rc = download(h, out, o))) return rc;
  return finish(h);
}

int ofri_warp_bilinear(ofri_handle h, const float* img, const float* cy, const float* cx, int batch, int H, int W,
                       float* out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!img || !cy || !cx || !out) return fail(h, OFRI_ERR_INVALID, "NULL pointer");
  int rc = arena_reserve(h, (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 4 + 4096);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img i = b.plane(batch, H, W), y = b.plane(batch, H, W), x = b.plane(batch, H, W), o = b.plane(batch, H, W);
  if ((rc = upload(h, i, img)) || (rc = upload(h, y, cy)) || (rc = upload(h, x, cx))) return rc;
  launch_warp_coords(i, y, x, o, h->stream, h->lc);
  if ((rc = download(h, out, o))) return rc;
  return finish(h);
}

int ofri_warp_pair(ofri_handle h, const float* im1, const float* im2, const float* us, const float* vs, int batch, int H,
                   int W, float* out1, float* out2) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!im1 || !im2 || !us || !vs || !out1 || !out2) return fail(h, OFRI_ERR_INVALID, "NULL pointer");
  int rc = arena_reserve(h, (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 6 + 4096);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img a = b.plane(batch, H, W), c = b.plane(batch, H, W), u = b.plane(batch, H, W), v = b.plane(batch, H, W),
      o1 = b.plane(batch, H, W), o2 = b.plane(batch, H, W);
  if ((rc = upload(h, a, im1)) || (rc = upload(h, c, im2)) || (rc = upload(h, u, us)) || (rc = upload(h, v, vs)))
    return rc;
  launch_warp_pair(a, c, u, v, o1, o2, h->stream, h->lc);
  if ((rc = download(h, out1, o1)) || (rc = download(h, out2, o2))) return rc;
  return finish(h);
}

int ofri_liu_shen_warp(ofri_handle h, const float* im1, const float* us, const float* vs, int batch, int H, int W,
                       const float* taps, int n_taps, float* out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!im1 || !us || !vs || !out) return fail(h, OFRI_ERR_INVALID, "NULL pointer");
  float t73[OFRI_MAX_GAUSS_TAPS];
  if (!taps) {
    const double sg = 0.6 * 3, tr = 4.0 / 0.6 * 3;
    n_taps = 2 * (int)(tr * sg + 0.5) + 1;
    ofri_gaussian_taps(sg, n_taps, t73);
    taps = t73;
  }
  if (n_taps < 1 || n_taps > OFRI_MAX_GAUSS_TAPS || !(n_taps & 1)) return fail(h, OFRI_ERR_INVALID, "bad tap count");
  if (H < n_taps / 2 || W < n_taps / 2) return fail(h, OFRI_ERR_TOO_SMALL, "image smaller than the kernel half-width");
  int rc = arena_reserve(h, (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 11 + 4096);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img i1 = b.plane(batch, H, W), u = b.plane(batch, H, W), v = b.plane(batch, H, W), o = b.plane(batch, H, W),
      win = b.plane(batch, H, W), dU = b.plane(batch, H, W), dV = b.plane(batch, H, W), fU = b.plane(batch, H, W),
      fV = b.plane(batch, H, W), sc = b.plane(batch, H, W), tmp = b.plane(batch, H, W);
  int* flag = (int*)b.take(sizeof(int) * 4);
  if ((rc = upload(h, i1, im1)) || (rc = upload(h, u, us)) || (rc = upload(h, v, vs))) return rc;
  cudaMemsetAsync(flag, 0, sizeof(int) * 4, h->stream);
  launch_liu_shen_warp(i1, u, v, o, (int*)win.p, dU, dV, fU, fV, sc, tmp, make_taps(taps, n_taps), flag, h->stream, h->lc);
  if ((rc = download(h, out, o))) return rc;
  int hf = 0;
  OFRI_CUDA(h, cudaMemcpyAsync(&hf, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if ((rc = finish(h))) return rc;
  if (hf) return fail(h, OFRI_ERR_INDEX, "index out of bounds: the Liu-Shen warp moved a pixel outside the frame");
  return OFRI_OK;
}

int ofri_hs_derivatives(ofri_handle h, const float* im1, const float* im2, int batch, int H, int W, float* fx,
                        float* fy, float* ft) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!im1 || !im2 || !fx || !fy || !ft) return fail(h, OFRI_ERR_INVALID, "NULL pointer");
  int rc = arena_reserve(h, (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 5 + 4096);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img a = b.plane(batch, H, W), c = b.plane(batch, H, W), dx = b.plane(batch, H, W), dy = b.plane(batch, H, W),
      dt = b.plane(batch, H, W);
  if ((rc = upload(h, a, im1)) || (rc = upload(h, c, im2))) return rc;
  launch_hs_derivs(a, c, dx, dy, dt, h->stream, h->lc);
  if ((rc = download(h, fx, dx)) || (rc = download(h, fy, dy)) || (rc = download(h, ft, dt))) return rc;
  return finish(h);
}

int ofri_hs_iterate(ofri_handle h, const float* u0, const float* v0, const float* fx, const float* fy, const float* ft,
                    int batch, int H, int W, float alpha, int niter, float* u_out, float* v_out) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!u0 || !v0 || !fx || !fy || !ft || !u_out || !v_out) return fail(h, OFRI_ERR_INVALID, "NULL pointer");
  if (niter < 0) return fail(h, OFRI_ERR_INVALID, "Niter < 0");
  int rc = arena_reserve(h, (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 7 + 4096);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img dx = b.plane(batch, H, W), dy = b.plane(batch, H, W), dt = b.plane(batch, H, W);
  Img U[2] = {b.plane(batch, H, W), b.plane(batch, H, W)}, V[2] = {b.plane(batch, H, W), b.plane(batch, H, W)};
  if ((rc = upload(h, dx, fx)) || (rc = upload(h, dy, fy)) || (rc = upload(h, dt, ft)) || (rc = upload(h, U[0], u0)) ||
      (rc = upload(h, V[0], v0)))
    return rc;
  int res = launch_hs_iterate(U[0], V[0], U[1], V[1], dx, dy, dt, alpha, niter, h->hs_fuse, h->hs_variant,
                              h->hs_precise >= 2, h->stream, h->lc);
  if ((rc = download(h, u_out, U[res])) || (rc = download(h, v_out, V[res]))) return rc;
  return finish(h);
}

int ofri_ls_coefficients(ofri_handle h, const float* im1, const float* im2, int batch, int H, int W, float hpar,
                         float* coef) {
  OFRI_ENTER(h);
  OFRI_DIMS(h, batch, H, W);
  if (!im1 || !im2 || !coef) return fail(h, OFRI_ERR_INVALID, "NULL pointer");
  int rc = arena_reserve(h, (sizeof(float) * (size_t)round_up(W, 4) * H * batch + 256) * 10 + 64 * (size_t)batch + 4096);
  if (rc) return rc;
  Bump b(h->arena, h->arena_cap, false);
  Img a = b.plane(batch, H, W), c = b.plane(batch, H, W);
  LsPlanes co;
  for (int i = 0; i < 8; ++i) co.c[i] = b.plane(batch, H, W);
  unsigned* mx = (unsigned*)b.take(sizeof(unsigned) * 2 * batch);
  if ((rc = upload(h, a, im1)) || (rc = upload(h, c, im2))) return rc;
  launch_ls_coefficients(a, c, hpar, co, mx, h->stream, h->lc);
  for (int i = 0; i < 8; ++i)
    if ((rc = download(h, coef + (size_t)i * batch * H * W, co.c[i]))) return rc;
  return finish(h);
}


// ---- row-band mode ------------------------------------------------------------------------------------------------------
int ofri_nccl_unique_id(void* out128) {
  std::string err;
  if (!out128) return fail(nullptr, OFRI_ERR_INVALID, "NULL pointer");
  if (ofri::nccl_unique_id(out128, &err)) return fail(nullptr, OFRI_ERR_COMM, "%s", err.c_str());
  return OFRI_OK;
}
int ofri_comm_init_nccl(ofri_handle h, int rank, int nranks, const void* uid128) {
  OFRI_ENTER(h);
  if (!uid128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, OFRI_ERR_INVALID, "bad communicator arguments");
  std::string err;
  ofri::Comm* c = ofri::make_nccl_comm(rank, nranks, uid128, &err);
  if (!c) return fail(h, OFRI_ERR_COMM, "%s", err.c_str());
  delete h->comm;
  h->comm = c;
  return OFRI_OK;
}
int ofri_local_group_create(int nranks, void** group) {
  if (!group) return fail(nullptr, OFRI_ERR_INVALID, "NULL pointer");
  *group = ofri::make_local_group(nranks);
  return *group ? OFRI_OK : fail(nullptr, OFRI_ERR_INVALID, "bad group size %d", nranks);
}
int ofri_local_group_destroy(void* group) {
  ofri::free_local_group((ofri::LocalGroup*)group);
  return OFRI_OK;
}
int ofri_local_group_abort(void* group) {
  ofri::abort_local_group((ofri::LocalGroup*)group);
  return OFRI_OK;
}
int ofri_comm_init_local(ofri_handle h, void* group, int rank) {
  OFRI_ENTER(h);
  std::string err;
  ofri::Comm* c = ofri::make_local_comm((ofri::LocalGroup*)group, rank, &err);
  if (!c) return fail(h, OFRI_ERR_COMM, "%s", err.c_str());
  delete h->comm;
  h->comm = c;
  return OFRI_OK;
}
int ofri_comm_destroy(ofri_handle h) {
  OFRI_ENTER(h);
  OFRI_CUDA(h, cudaStreamSynchronize(h->stream));
  delete h->comm;
  h->comm = nullptr;
  return OFRI_OK;
}
int ofri_band_plan(ofri_handle h, int H, int W, const ofri_params* p, int rank, int nranks, ofri_band* out) {
  OFRI_ENTER(h);
  if (!out) return fail(h, OFRI_ERR_INVALID, "NULL pointer");
  BandPlanInt bp;
  int rc = make_band_plan(h, H, W, p, rank, nranks, &bp);
  if (rc) return rc;
  const BandLevel& f = bp.lv[bp.L - 1];
  out->rank = rank; out->nranks = nranks;
  out->own0 = f.own0; out->own1 = f.own1;
  out->in0 = bp.in0; out->in1 = bp.in1;
  out->ghost = f.G; out->exchange = bp.E;
  return OFRI_OK;
}
int ofri_band_plan_host(int H, int W, const ofri_params* p, int rank, int nranks, int hs_fuse, int band_exchange,
                        int band_reach, ofri_band* out) {
  if (!out) return fail(nullptr, OFRI_ERR_INVALID, "NULL pointer");
  BandPlanInt bp;
  int rc = make_band_plan_opts(nullptr, H, W, p, rank, nranks, hs_fuse, band_exchange, band_reach, &bp);
  if (rc) return rc;
  const BandLevel& f = bp.lv[bp.L - 1];
  out->rank = rank; out->nranks = nranks;
  out->own0 = f.own0; out->own1 = f.own1;
  out->in0 = bp.in0; out->in1 = bp.in1;
  out->ghost = f.G; out->exchange = bp.E;
  return OFRI_OK;
}
int ofri_pyramidal_flow_banded_dev(ofri_handle h, const float* d_im1_rows, const float* d_im2_rows, int H, int W,
                                   const ofri_params* p, float* d_u_rows, float* d_v_rows, float* d_err_out) {
  OFRI_ENTER(h);
  if (!d_im1_rows || !d_im2_rows || !d_u_rows || !d_v_rows) return fail(h, OFRI_ERR_INVALID, "NULL image / output pointer");
  const int rank = h->comm ? h->comm->rank : 0, n = h->comm ? h->comm->nranks : 1;
  BandPlanInt bp;
  int rc = make_band_plan(h, H, W, p, rank, n, &bp);
  if (rc) return rc;
  rc = arena_reserve(h, band_workspace_bytes(bp, W, p));
  if (rc) return rc;
  Bump bump(h->arena, h->arena_cap, false);
  BandWs ws;
  plan_band_ws(bump, bp, W, p, &ws);
  rc = run_pyramid_banded(h, d_im1_rows, d_im2_rows, H, W, p, bp, d_u_rows, d_v_rows, d_err_out, ws);
  if (rc) return rc;
  int flag = 0;      // one synchronisation at the end: did the warp stay inside the ghost frame?
  OFRI_CUDA(h, cudaMemcpyAsync(&flag, ws.flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  OFRI_CUDA(h, cudaStreamSynchronize(h->stream));
  collect_times(h);
  if (flag)
    return fail(h, OFRI_ERR_UNSUPPORTED, "row-band mode: the flow displaces rows by more than band_reach = %d; raise the "
                "'band_reach' option", bp.Rw);
  return OFRI_OK;
}

}  // extern "C"
