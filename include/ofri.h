/* ofri.h -- C ABI of libofri.so: B200 (sm_100a) implementation of the OpticalFlow-RI variational hot path.
 *
 * This is the drop-in boundary for ONE path of alexlib/OpticalFlow-RI: the Horn-Schunck Jacobi solver and
 * the Liu-Shen physics-based solver inside the `genericPyramidalOpticalFlow` coarse-to-fine driver, with its
 * per-level stages.  Every entry point is what a ctypes binding of the reference's Python functions would
 * bind (the reference is pure Python: its "FFI" is the call signature of the functions cited below; paths
 * are relative to the reference's src/ directory).  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - images / flow planes: float32, row-major [batch][H][W], dense (row pitch == W) unless a `_dev` entry
 *     point takes an explicit pitch.  U = x/column component, V = y/row component
 *     (GenericPyramidalOpticalFlow.py:256-268).
 *   - host entry points take HOST pointers and do the H2D / D2H copies themselves (the reference's functions
 *     take and return numpy arrays); `_dev` entry points take DEVICE pointers on the handle's device and
 *     enqueue on the handle's stream without synchronising.
 *   - every function returns 0 on success or a negative ofri_status; ofri_last_error() returns the message.
 *     There is NO CPU fallback: without a CUDA device ofri_create() fails with OFRI_ERR_NO_DEVICE.
 *   - a handle owns one device, one stream and its workspace; it is not re-entrant.  Use one handle per
 *     host thread / per GPU (one process per GPU under torchrun).
 */
#ifndef OFRI_H_
#define OFRI_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define OFRI_API __attribute__((visibility("default")))
#else
#define OFRI_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define OFRI_ABI_VERSION 1
#define OFRI_MAX_GAUSS_TAPS 129
#define OFRI_MAX_ALPHAS 64

typedef struct ofri_ctx* ofri_handle;

typedef enum {
  OFRI_OK = 0,
  OFRI_ERR_INVALID = -1,       /* bad argument (maps to ValueError / Exception('Invalid scale level')) */
  OFRI_ERR_NO_DEVICE = -2,     /* no CUDA device / driver: the library never falls back to the CPU */
  OFRI_ERR_CUDA = -3,          /* a CUDA runtime call or kernel failed; see ofri_last_error */
  OFRI_ERR_OOM = -4,           /* device or pinned-host allocation failed */
  OFRI_ERR_ALPHAS = -5,        /* HS alpha list exhausted (IndexError at HornSchunck.py:36) */
  OFRI_ERR_FILTER_OPT = -6,    /* optional adapter given without FILTER_OPT (TypeError at GPOF:380) */
  OFRI_ERR_TOO_SMALL = -7,     /* a pyramid level has < 4 samples on an axis (scipy raises in the spline) */
  OFRI_ERR_UNSUPPORTED = -8,   /* combination outside the native path (e.g. row-band mode with kLevels > 1) */
  OFRI_ERR_COMM = -9,          /* NCCL / multi-GPU error */
  OFRI_ERR_INDEX = -10,        /* biLinear=False warp: a scatter target left the frame (IndexError at GPOF:207) */
  OFRI_ERR_CALLBACK = -11      /* an external adapter's compute() callback reported failure (ofri_pyramidal_flow_external) */
} ofri_status;

/* OFRI_ALGO_EXTERNAL: a foreign adapter (any object with the reference's duck-typed compute(), GPOF:256-290) driven
 * through a callback -- only valid with ofri_pyramidal_flow_external */
typedef enum { OFRI_ALGO_NONE = -1, OFRI_ALGO_HS = 0, OFRI_ALGO_LS = 1, OFRI_ALGO_EXTERNAL = 2,
               OFRI_ALGO_FB = 3 /* Farneback: parameters from ofri_set_farneback */,
               OFRI_ALGO_LK = 4 /* dense Lucas-Kanade: parameters from ofri_set_lk */ } ofri_algo_kind;

/* One optical-flow algorithm adapter (the reference's plugin protocol: compute(im1, im2, U, V) -> (U, V, error),
 * GenericPyramidalOpticalFlow.py:256-290).
 *   HS: HSOpticalFlowAlgoAdapter(alphas, Niter)  HornSchunck.py:29-50.  `alphas` holds the values in the ORDER
 *       OF USE (the Python shim pops them from the END of the caller's list, HornSchunck.py:36): one per
 *       (level, k) compute call.
 *   LS: LiuShenOpticalFlowAlgoAdapter(alpha)     PhysicsBasedOpticalFlowLiuShen.py:33-45; ls_maxiter=60 and
 *       ls_tol=1e-8 are the reference's hard-coded maxnum / tol (PhysicsBasedOpticalFlowLiuShen.py:88-89). */
typedef struct {
  int32_t kind;                      /* ofri_algo_kind */
  int32_t hs_niter;
  int32_t n_alphas;
  int32_t ls_maxiter;
  float   alphas[OFRI_MAX_ALPHAS];
  float   ls_h;
  float   reserved_;
  double  ls_tol;
} ofri_algo;

/* Arguments of genericPyramidalOpticalFlow (GenericPyramidalOpticalFlow.py:238-239) AFTER the adapter-default
 * override of lines 304-327 has been applied by the caller.  The Gaussian taps are passed as coefficients
 * (generated on the host exactly as gaussian_filter.py:47-52 does, or by ofri_gaussian_taps) so that they are
 * bit-identical to the reference's; n_taps == 0 means "filter off" (FILTER <= 1e-3, GPOF:368 / 380). */
typedef struct {
  uint32_t size;                     /* sizeof(ofri_params): ABI versioning */
  int32_t  pyramid_levels;           /* pyramidalLevels */
  int32_t  k_levels;                 /* kLevels */
  int32_t  warping;                  /* warping */
  int32_t  bilinear;                 /* biLinear: 1 = symmetric bilinear warp, 0 = "Liu-Shen warp" of frame 1 (GPOF:204-221) */
  int32_t  intermediate_scaling;     /* pyramidalIntermediateScaling */
  int32_t  final_scaling;            /* pyramidalScaling */
  int32_t  n_taps_main;              /* FILTER taps (3 in the reference), 0 = off */
  int32_t  n_taps_opt;               /* FILTER_OPT taps (5 in the reference), 0 = off */
  int32_t  refilter_k;               /* k>0 re-warp branch filters iff FILTER > 1 (GPOF:396) */
  float    taps_main[OFRI_MAX_GAUSS_TAPS];
  float    taps_opt[OFRI_MAX_GAUSS_TAPS];
  ofri_algo main_algo;               /* mainOFlowAlgoAdapter */
  ofri_algo opt_algo;                /* optionalOFlowAlgoAdapter (kind = OFRI_ALGO_NONE if absent) */
  /* biLinear=False only: taps of gaussian_filter(x, 0.6*3, truncate=4.0/0.6*3) (GPOF:210-212; 73 taps), generated by
   * the caller like the other taps; n_taps_lsw == 0 lets the library generate them (ofri_gaussian_taps) */
  int32_t  n_taps_lsw;
  float    taps_lsw[OFRI_MAX_GAUSS_TAPS];
} ofri_params;

/* ---- lifetime ------------------------------------------------------------------------------------------ */
OFRI_API int ofri_abi_version(void);
OFRI_API int ofri_device_count(void);                                   /* <0 on error (no driver) */
OFRI_API int ofri_create(int device, ofri_handle* out);
OFRI_API int ofri_destroy(ofri_handle h);
OFRI_API const char* ofri_last_error(ofri_handle h);                    /* h may be NULL: last error of ofri_create */
/* run on a caller-owned stream (e.g. torch's current stream) instead of the handle's own; 0 restores */
OFRI_API int ofri_set_stream(ofri_handle h, void* cuda_stream);
OFRI_API int ofri_synchronize(ofri_handle h);
/* tuning / A-B switches, e.g. ("hs_fuse", 4), ("ls_fuse", 2), ("chunk_pairs", 16); the full table is in
 * INTEGRATION.md section 5; unknown key -> OFRI_ERR_INVALID */
OFRI_API int ofri_set_option(ofri_handle h, const char* key, int value);
OFRI_API int ofri_get_option(ofri_handle h, const char* key, int* value);
/* number of kernel launches issued by this handle since creation (for bench.py's gpu_launches) */
OFRI_API int64_t ofri_launch_count(ofri_handle h);
/* device-time breakdown of the last host-level call: names[i] -> ms[i]; returns the number of entries */
OFRI_API int ofri_stage_timings(ofri_handle h, const char** names, float* ms, int max_entries);
/* profiling hook (no reference counterpart): SM cycles thread 0 of every CTA of the persistent sweep kernels spent per
 * phase since the last read; family 0 = Horn-Schunck kernel, 1 = Liu-Shen kernel; out8[8].  All zero unless the
 * library was built with -DOFRI_PHASE_TIMING (tools/phase_timing.py builds that variant as libofri_phase.so). */
OFRI_API int ofri_debug_phase_read(ofri_handle h, int family, unsigned long long* out8);

/* ---- whole path: replaces genericPyramidalOpticalFlow(...) (GenericPyramidalOpticalFlow.py:238-416) and
 *      GenericPyramidalOpticalFlowWrapper.calculateFlow (GenericPyramidalOpticalFlowWrapper.py:41-64) --------
 * im1/im2: [batch][H][W]; u_out/v_out: [batch][H][W] (the returned Uaccum, Vaccum).
 * err_out: optional [batch][levels*k_levels][2] = (main error, optional error) per compute call, or NULL. */
OFRI_API int ofri_pyramidal_flow(ofri_handle h, const float* im1, const float* im2, int batch, int H, int W,
                        const ofri_params* p, float* u_out, float* v_out, float* err_out);
/* ---- Farneback adapter: replaces Farneback_PyCL (Farneback_PyCL.py:65-616) and its six OpenCL kernels
 *      (optical_flow_farneback.cl:72-429) -----------------------------------------------------------------------------
 * Constructor parameters (FB:70-71) plus the coefficient tables the reference computes on the host (the caller passes
 * them so that they are bit-identical to the reference's: FarnebackPrepareGaussian FB:124-176, setGaussianBlurKernel /
 * getGaussianKernelBitExact FB:194-207).  extra_levels = pyramidalLevels - 1 (the adapter's INTERNAL pyramid, FB:78);
 * level k of it runs at scale pyr_scale^k and pre-blurs the frames with blur_kernel[k][0 .. n_blur[k]] (centre + right
 * half).  win_kernel: centre + right half of the window kernel (use_gaussian), window_size / 2 + 1 entries. */
#define OFRI_FB_MAX_HALF 64
#define OFRI_FB_MAX_LEVELS 12
typedef struct {
  uint32_t size;                     /* sizeof(ofri_farneback_params) */
  int32_t  window_size, n_iters, poly_n, use_gaussian, extra_levels;
  float    pyr_scale;
  float    g[8], xg[8], xxg[8], ig[4];
  float    win_kernel[OFRI_FB_MAX_HALF + 1];
  int32_t  n_blur[OFRI_FB_MAX_LEVELS];
  float    blur_kernel[OFRI_FB_MAX_LEVELS][OFRI_FB_MAX_HALF + 1];
} ofri_farneback_params;
/* compute(im1, im2, U, V) -> (U, V) of the adapter (FB:462-604); u0 / v0 NULL = zero initial flow */
OFRI_API int ofri_farneback_compute(ofri_handle h, const float* im1, const float* im2, const float* u0, const float* v0,
                                    int batch, int H, int W, const ofri_farneback_params* fp, float* u_out, float* v_out);
/* imresize of Farneback_PyCL.py:61-62: Pillow BILINEAR (antialiased triangle filter), up- and down-sampling */
OFRI_API int ofri_resize_bilinear(ofri_handle h, const float* in, int batch, int H, int W, int out_h, int out_w, float* out);
/* the parameters an adapter of kind OFRI_ALGO_FB uses inside ofri_pyramidal_flow* (copied into the handle) */
OFRI_API int ofri_set_farneback(ofri_handle h, const ofri_farneback_params* fp);

/* ---- dense Lucas-Kanade adapter (reference: src/denseLucasKanade_PyCL.py LK:line + the OpenCL kernel
 * lkDense, src/pyrlkDenseLargeW.cl) ------------------------------------------------------------------------------------
 * Constructor parameters (LK:34): n_iters (Niter), half_window (window = 2 half_window + 1 on both axes; the kernel's
 * sample grid covers at most 32 x 32).  asym = {left, right, top, bottom}: the asymmetric-window switches the adapter
 * derives from the mean vorticity of the incoming flow when enableVorticityEnhancement is set (LK:75-92); all zero
 * otherwise.  The OpenCL original leaves three operations to the device (sampler filter precision, mad, division):
 * this library uses the OpenCL-specification bilinear filter in full float32, fused multiply-add and IEEE division. */
typedef struct {
  uint32_t size;                     /* sizeof(ofri_lk_params) */
  int32_t  n_iters, half_window;
  int32_t  asym[4];
} ofri_lk_params;
/* compute(im1, im2, U, V) -> (U, V) of the adapter (LK:113-169); u0 / v0 NULL = zero initial flow */
OFRI_API int ofri_lk_compute(ofri_handle h, const float* im1, const float* im2, const float* u0, const float* v0,
                             int batch, int H, int W, const ofri_lk_params* lp, float* u_out, float* v_out);
/* the parameters an adapter of kind OFRI_ALGO_LK uses inside ofri_pyramidal_flow* (copied into the handle) */
OFRI_API int ofri_set_lk(ofri_handle h, const ofri_lk_params* lp);

/* page-locked host memory for callers without a CUDA binding of their own (opticalflow_ri_b200/pipeline.py: the frame
 * ring of the file -> GPU -> file pipeline): buffers from here make the host-pointer call above fully asynchronous */
OFRI_API int ofri_host_alloc(ofri_handle h, size_t bytes, void** out);
OFRI_API int ofri_host_free(ofri_handle h, void* p);

/* The same driver with FOREIGN adapters (the reference's plugin protocol, GenericPyramidalOpticalFlow.py:256-290, call
 * sites :406 and :410; e.g. examples/LiuSE_denseLK_Fs2_0_PyrLvls2.py:68-74: a dense Lucas-Kanade main adapter refined by
 * Liu-Shen): an adapter whose kind is OFRI_ALGO_EXTERNAL is computed by `fn` -- the host-side compute(im1, im2, U, V) of
 * that adapter -- everything else (level stages, the other adapter if it is HS / LS, accumulation) stays on the device.
 * Per external compute(): one D2H of the level's (filtered, warped) frames and of the current U, V into dense host
 * buffers, the callback, one H2D of the U, V it wrote; nothing else crosses the bus until u_out / v_out at the end.
 * fn(user, which, call_index, im1, im2, U, V, H, W, err): which = 0 main / 1 optional; call_index = level * k_levels + k;
 * im1 / im2 read-only H x W; U / V in: current flow increment, out: the adapter's result; *err = its error estimate;
 * return 0, or non-zero to abort the call with OFRI_ERR_CALLBACK.  One pair per call (foreign adapters are not batched). */
typedef int (*ofri_adapter_fn)(void* user, int which, int call_index, const float* im1, const float* im2, float* U,
                               float* V, int H, int W, float* err);
OFRI_API int ofri_pyramidal_flow_external(ofri_handle h, const float* im1, const float* im2, int H, int W,
                                          const ofri_params* p, ofri_adapter_fn fn, void* user, float* u_out, float* v_out,
                                          float* err_out);
/* same with DEVICE pointers (dense, pitch == W), enqueued on the handle's stream, no synchronisation */
OFRI_API int ofri_pyramidal_flow_dev(ofri_handle h, const float* d_im1, const float* d_im2, int batch, int H, int W,
                            const ofri_params* p, float* d_u_out, float* d_v_out, float* d_err_out);

/* ---- adapter compute() stand-alone ------------------------------------------------------------------------
 * HSOpticalFlowAlgoAdapter.compute for ONE alpha (HornSchunck.py:35-37 -> HS, 73-105): exactly `niter` Jacobi
 * sweeps from (u0, v0); err[b] = (||U-u0||_F + ||V-v0||_F) / (H*W). */
OFRI_API int ofri_hs_compute(ofri_handle h, const float* im1, const float* im2, const float* u0, const float* v0,
                    int batch, int H, int W, float alpha, int niter, float* u_out, float* v_out, float* err);
/* LiuShenOpticalFlowAlgoAdapter.compute (PhysicsBasedOpticalFlowLiuShen.py:37-39 -> 82-158), including the
 * U/V swap; iters[b] = sweeps actually run (early exit when total_error <= tol). */
OFRI_API int ofri_ls_compute(ofri_handle h, const float* im1, const float* im2, const float* u0, const float* v0,
                    int batch, int H, int W, float hpar, int maxiter, double tol,
                    float* u_out, float* v_out, float* err, int32_t* iters);

/* ---- per-level stages (test hooks; each is one or two sm_100a kernels) -------------------------------------
 * gaussian_filterPx / convolveSeparableFilter with explicit taps (gaussian_filter.py:54-94) */
OFRI_API int ofri_gauss_px(ofri_handle h, const float* in, int batch, int H, int W, const float* taps, int n_taps, float* out);
/* prepareGaussianKernel (gaussian_filter.py:47-52) computed in C (double exp, float store, float normalise) */
OFRI_API int ofri_gaussian_taps(double sigma, int n_taps, float* taps_out);
/* imresize: Pillow BICUBIC antialiased resample (GenericPyramidalOpticalFlow.py:67-68) */
OFRI_API int ofri_resize_bicubic(ofri_handle h, const float* in, int batch, int H, int W, int out_h, int out_w, float* out);
/* level size int32(round(n*scale)), half-to-even (GenericPyramidalOpticalFlow.py:338-343) */
OFRI_API int ofri_level_size(int n, double scale);
/* RectBivariateSpline up-sample of one flow component + optional scale factor (GPOF:155-172); mul = 1 -> none */
OFRI_API int ofri_spline_upsample(ofri_handle h, const float* in, int batch, int in_h, int in_w, int out_h, int out_w,
                         float mul, float* out);
/* doBiLinearWarping(img, coordsY, coordsX) (GenericPyramidalOpticalFlow.py:70-116) */
OFRI_API int ofri_warp_bilinear(ofri_handle h, const float* img, const float* cy, const float* cx, int batch, int H, int W,
                       float* out);
/* the symmetric pair warp of updateNextPyramidalLevel (GPOF:200-201): im1 at (y-v/2, x-u/2), im2 at (y+v/2, x+u/2) */
OFRI_API int ofri_warp_pair(ofri_handle h, const float* im1, const float* im2, const float* us, const float* vs,
                   int batch, int H, int W, float* out1, float* out2);
/* the biLinear=False "Liu-Shen warp" of frame 1 by the (already up-sampled) flow (GPOF:190-196, 204-221); taps = the
 * 73-tap kernel of gaussian_filter(x, 0.6*3, truncate=4.0/0.6*3) or NULL (generated by the library);
 * OFRI_ERR_INDEX when a scatter target leaves the frame */
OFRI_API int ofri_liu_shen_warp(ofri_handle h, const float* im1, const float* us, const float* vs, int batch, int H, int W,
                       const float* taps, int n_taps, float* out);
/* computeDerivatives as reached from compute(im1, im2) (HornSchunck.py:84, 107-127) */
OFRI_API int ofri_hs_derivatives(ofri_handle h, const float* im1, const float* im2, int batch, int H, int W,
                        float* fx, float* fy, float* ft);
/* HS_helper: `niter` Jacobi sweeps on given derivative planes (HornSchunck.py:62-71) */
OFRI_API int ofri_hs_iterate(ofri_handle h, const float* u0, const float* v0, const float* fx, const float* fy,
                    const float* ft, int batch, int H, int W, float alpha, int niter, float* u_out, float* v_out);
/* Liu-Shen coefficient planes, coef = [8][batch][H][W]: IIx, IIy, II, Ixt, Iyt, B11, B12, B22
 * (PhysicsBasedOpticalFlowLiuShen.py:96-97, 124-128, 47-73) */
OFRI_API int ofri_ls_coefficients(ofri_handle h, const float* im1, const float* im2, int batch, int H, int W, float hpar,
                         float* coef);

/* ---- row-band domain decomposition: ONE very large frame pair over several GPUs -----------------------------------
 * New work (the reference is single-process; BASELINE.json configs[4]).  One handle per GPU; the handles of one job
 * share a communicator: NCCL (one process per GPU: rank 0 calls ofri_nccl_unique_id and ships the 128 bytes to the
 * other ranks, e.g. with torch.distributed.broadcast) or "local" (one process, one host thread per band, any number
 * of bands per GPU -- used by the band-invariance tests).  Rank r owns rows [own0, own1) of the frame and must supply
 * rows [in0, in1) of both frames (owned rows plus the halo the pyramid needs); ofri_band_plan tells both.
 * Requirements: k_levels == 1, warping == bilinear == 1 when pyramid_levels > 1, H divisible by
 * nranks * 2^(pyramid_levels-1).  Owned rows of the result are bit-identical to ofri_pyramidal_flow_dev's. */
typedef struct {
  int32_t rank, nranks;
  int32_t own0, own1;     /* rows of the frame this rank owns (and returns) */
  int32_t in0, in1;       /* rows of the input frames this rank must be given */
  int32_t ghost;          /* ghost rows per side of the per-level arrays */
  int32_t exchange;       /* Horn-Schunck sweeps between two ghost-row exchanges */
} ofri_band;
OFRI_API int ofri_nccl_unique_id(void* out128);
OFRI_API int ofri_comm_init_nccl(ofri_handle h, int rank, int nranks, const void* uid128);
OFRI_API int ofri_local_group_create(int nranks, void** group);
OFRI_API int ofri_local_group_destroy(void* group);
/* a rank of the group failed (or will not reach its collectives): wake the ranks blocked in a collective and make
 * every later collective of the group return OFRI_ERR_COMM instead of waiting forever */
OFRI_API int ofri_local_group_abort(void* group);
OFRI_API int ofri_comm_init_local(ofri_handle h, void* group, int rank);   /* call from the thread that drives `rank` */
OFRI_API int ofri_comm_destroy(ofri_handle h);
OFRI_API int ofri_band_plan(ofri_handle h, int H, int W, const ofri_params* p, int rank, int nranks, ofri_band* out);
/* the same without a handle / GPU (planning on a loader or scheduler host): options passed explicitly, 0 = default */
OFRI_API int ofri_band_plan_host(int H, int W, const ofri_params* p, int rank, int nranks, int hs_fuse, int band_exchange,
                        int band_reach, ofri_band* out);
/* d_im*_rows: DEVICE, rows [in0, in1) of the frames, dense (pitch W); d_u_rows / d_v_rows: DEVICE, rows [own0, own1);
 * d_err_out: optional DEVICE [levels][2].  Collective: every rank of the communicator must call it.  Synchronises the
 * handle's stream before returning. */
OFRI_API int ofri_pyramidal_flow_banded_dev(ofri_handle h, const float* d_im1_rows, const float* d_im2_rows, int H, int W,
                                   const ofri_params* p, float* d_u_rows, float* d_v_rows, float* d_err_out);

#ifdef __cplusplus
}
#endif
#endif /* OFRI_H_ */
