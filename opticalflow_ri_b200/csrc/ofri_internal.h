// ofri_internal.h -- launcher interface between the C-ABI / driver (ofri_api.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <functional>
#include <string>
#include "../../include/ofri.h"
#include "ofri_spline.cuh"

namespace ofri {

// A stack of `batch` float32 planes in device memory: element (b, y, x) at p[b*stride + y*pitch + x].
// Internal planes use pitch = round_up(W, 4) so that rows start 16-byte aligned (float4 / TMA).
struct Img {
  float* p = nullptr;
  int H = 0, W = 0;
  long pitch = 0;
  long stride = 0;
  int batch = 0;
  __host__ __device__ float* at(int b) const { return p + (long)b * stride; }
};
struct ImgD {  // same for float64 scratch (spline)
  double* p = nullptr;
  int H = 0, W = 0;
  long pitch = 0;
  long stride = 0;
  int batch = 0;
};

struct GaussTaps {
  int K;
  float k[OFRI_MAX_GAUSS_TAPS];
  __host__ __device__ float operator[](int i) const { return k[i]; }
};

// Pillow tap table for one axis (device pointers)
struct ResizeTaps {
  const int* xmin = nullptr;
  const int* cnt = nullptr;
  const double* w = nullptr;   // [out][kmax]
  int kmax = 0;
  int in_size = 0, out_size = 0;
};
// Thomas-algorithm constants of the not-a-knot system for n samples (device pointers, m = n-2 entries)
typedef SplineSysView SplineSys;

struct LaunchCounter { int64_t n = 0; };

// Row-band mode of the Liu-Shen sweeps: the residual sums cover local rows [own_lo, own_hi) only (the rows this band
// owns) and total_error is normalised by the pixel count of the WHOLE image.
struct LsBand { int own_lo; int own_hi; double npix; };
// Called (host side, between launches) after sweeps [k0, k0 + n) were enqueued; `written` = which ping-pong buffer
// the launch wrote (0 = a, 1 = b, 2 = possibly both).  The band driver all-reduces errs[k0 .. k0+n) and exchanges the
// ghost rows here; stream-ordered, no host synchronisation.
typedef std::function<void(int k0, int n, int written)> LsHook;
// Same for Horn-Schunck: after `done` sweeps in total, the current state is in buffer `cur` (0 = a, 1 = b).
typedef std::function<void(int done, int cur)> HsHook;
// Split launch of the TMA kernel: only the tile rows lying entirely inside output rows [row_lo, row_hi) (inside = true)
// or all the other tile rows (inside = false).
// reserve_sms: leave that many SMs without a CTA of the (persistent, one CTA per SM, whole register file) kernel, so that
// a communication kernel enqueued on another stream can start while this launch runs.
struct HsTileRows { int row_lo; int row_hi; bool inside; int reserve_sms = 0; };
// Row-band mode with overlap: the launch that completes a block of `every` sweeps (and the last launch) is issued in
// two parts -- first the tiles that produce the rows the neighbours need (everything outside [mid_lo, mid_hi)), then
// begin(cur) starts the ghost-row exchange of buffer `cur` on the communication stream, then the interior tiles run
// under it, then end() makes the compute stream wait for the exchange.
struct HsSplit {
  int every = 0;
  int mid_lo = 0, mid_hi = 0;
  int reserve_sms = 0;            // SMs the interior launch leaves to the exchange running beside it
  std::function<void(int cur)> begin;
  std::function<void()> end;
};


// ---- stages (ofri_stages.cu) ---------------------------------------------------------------------------------
void launch_gauss(const Img& in, const Img& tmp, const Img& out, const GaussTaps& taps, cudaStream_t s, LaunchCounter& lc);
// in_row0 / out_row0: global row of local row 0 of `in` / `out` when they are row bands of larger images (ty is
// always the tap table of the WHOLE image)
void launch_resize(const Img& in, const Img& tmp, const Img& out, const ResizeTaps& tx, const ResizeTaps& ty,
                   cudaStream_t s, LaunchCounter& lc, int in_row0 = 0, int out_row0 = 0);
// up-sample `in` (h x w) to `out` (H x W) and multiply by mul.
// Strip form (row bands): `in` holds rows [in_row0, in_row0 + in.H) of the hg-row coarse plane, `out` rows
// [row0, row0 + out.H) of the Hg-row result; in_row0 = row0 = 0, hg = in.H, Hg = out.H is the whole-image case.
// f64 scratch: M1 and D1 at least in.H x in.W (same strip indexing as `in`).  Chunk-parallel windowed solves
// (ofri_spline.cuh) + one fused kernel per output row (axis-0 evaluation, axis-1 solve and evaluation in shared
// memory).  Returns false if `in` does not cover the rows the requested output rows need (spline_rows_needed).
bool launch_spline(const Img& in, const Img& out, float mul, const SplineSys& sy, const SplineSys& sx, const ImgD& M1,
                   const ImgD& D1, cudaStream_t s, LaunchCounter& lc, int in_row0 = 0, int hg = 0, int row0 = 0,
                   int Hg = 0);
// coarse rows [*lo, *hi) a strip must hold for output rows [row0, row0 + rows) of Hg (coarse plane: hg rows, system sy)
void spline_rows_needed(int row0, int rows, int hg, int Hg, const SplineSys& sy, int* lo, int* hi);
// previous generation: one thread per line, sequential full-length solves (whole coarse plane only; kept as the A/B
// reference of the tests).  Scratch: M1 (h x w), T1 (H x w), M2 (H x w) f64.
void launch_spline_seq(const Img& in, const Img& out, float mul, const SplineSys& sy, const SplineSys& sx,
                       const ImgD& M1, const ImgD& T1, const ImgD& M2, cudaStream_t s, LaunchCounter& lc, int row0 = 0,
                       int Hg = 0);
// band form: us / vs / out hold rows [row0, ..) and im1 / im2 rows [img_row0, ..) of an image of Hg rows
void launch_warp_pair(const Img& im1, const Img& im2, const Img& us, const Img& vs, const Img& out1, const Img& out2,
                      cudaStream_t s, LaunchCounter& lc, int row0 = 0, int img_row0 = 0, int Hg = 0);
// biLinear = False "Liu-Shen warp" of frame 1 (GPOF:190-196, 204-221); *flag is set if a scatter target leaves the frame
void launch_liu_shen_warp(const Img& im1, const Img& us, const Img& vs, const Img& out, int* winner, const Img& dU,
                          const Img& dV, const Img& fU, const Img& fV, const Img& sc, const Img& tmp,
                          const GaussTaps& taps, int* flag, cudaStream_t s, LaunchCounter& lc);
void launch_warp_coords(const Img& img, const Img& cy, const Img& cx, const Img& out, cudaStream_t s, LaunchCounter& lc);
void launch_axpy(const Img& acc, const Img& x, cudaStream_t s, LaunchCounter& lc);       // acc += x
void launch_scale(const Img& x, float mul, cudaStream_t s, LaunchCounter& lc);            // x *= mul
void launch_copy(const Img& dst, const Img& src, cudaStream_t s, LaunchCounter& lc);      // dst = src (kernel copy)
void launch_fill(const Img& dst, float v, cudaStream_t s, LaunchCounter& lc);

// ---- Horn-Schunck (ofri_hs.cu) ----------------------------------------------------------------------------------
void launch_hs_derivs(const Img& im1, const Img& im2, const Img& fx, const Img& fy, const Img& ft, cudaStream_t s,
                      LaunchCounter& lc);
// `niter` Jacobi sweeps.  u0/v0 -> result in ua/va or ub/vb (ping-pong); returns which (0 = a, 1 = b).
// fuse = sweeps per launch (0 = simple per-pixel kernel; >= 1 = temporally blocked shared-memory kernel).
// precise = reference arithmetic bit for bit (f64-accumulated stencil, IEEE division) instead of the f32/FMA fast path.
int launch_hs_iterate(const Img& ua, const Img& va, const Img& ub, const Img& vb, const Img& fx, const Img& fy,
                      const Img& ft, float alpha, int niter, int fuse, int variant, bool precise, cudaStream_t s,
                      LaunchCounter& lc, const HsHook& hook = HsHook(), const HsSplit* split = nullptr);
// persistent TMA-fed register-resident fused sweeps (ofri_hs_tma.cu); false = not applicable, use another kernel
bool launch_hs_tma(int T, int variant, bool precise, const Img& ui, const Img& vi, const Img& uo, const Img& vo,
                   const Img& fx, const Img& fy, const Img& ft, float alpha2, cudaStream_t s,
                   const HsTileRows* sub = nullptr);
// err[b] = (sqrt(sum (u-u0)^2) + sqrt(sum (v-v0)^2)) / (H*W); u0.p == nullptr means u0 = v0 = 0.  acc: [batch][2] f64 scratch
void launch_hs_error(const Img& u, const Img& v, const Img& u0, const Img& v0, double* acc, float* err, int err_stride,
                     cudaStream_t s, LaunchCounter& lc);
// the two halves of launch_hs_error for row bands: sums over local rows [y_lo, y_hi) into acc (to be all-reduced),
// then err = (sqrt(acc0) + sqrt(acc1)) / npix
void launch_hs_error_sums(const Img& u, const Img& v, const Img& u0, const Img& v0, double* acc, int y_lo, int y_hi,
                          cudaStream_t s, LaunchCounter& lc);
void launch_hs_error_finish(const double* acc, float* err, int err_stride, int batch, double npix, cudaStream_t s,
                            LaunchCounter& lc);

// ---- Liu-Shen (ofri_ls.cu) -----------------------------------------------------------------------------------------
struct LsPlanes { Img c[8]; };   // IIx, IIy, II, Ixt, Iyt, B11, B12, B22
void launch_ls_max(const Img& im1, const Img& im2, unsigned* maxenc, cudaStream_t s, LaunchCounter& lc);
void launch_ls_coef(const Img& im1, const Img& im2, float hpar, const LsPlanes& coef, const unsigned* maxenc,
                    cudaStream_t s, LaunchCounter& lc, int row0 = 0, int Hg = 0);
// maxenc: [batch][2] uint32 scratch (ordered-int encoded maxima of im1 / im2)
void launch_ls_coefficients(const Img& im1, const Img& im2, float hpar, const LsPlanes& coef, unsigned* maxenc,
                            cudaStream_t s, LaunchCounter& lc);
// Runs up to maxiter sweeps with the reference's stopping rule per pair.  (ua,va) holds the initial guess
// (ROW component in ua!), (ub,vb) is the ping-pong partner; the final state is copied into (uo, vo).
// errs: [batch][maxiter][2] f64 scratch; err_out[b*err_stride] = last total_error; iters_out[b] = sweeps run.
void launch_ls_solve(const Img& ua, const Img& va, const Img& ub, const Img& vb, const LsPlanes& coef, float hpar,
                     int maxiter, double tol, int fuse, int variant, double* errs, int* state, const Img& uo,
                     const Img& vo, float* err_out, int err_stride, int* iters_out, cudaStream_t s, LaunchCounter& lc,
                     const LsBand* band = nullptr, const LsHook& hook = LsHook());

// ---- communication back ends of the row-band driver (ofri_comm.cu) ------------------------------------------------
// All operations are enqueued on stream s (plus host-side rendezvous for the local back end); 0 = ok, -1 = see error().
struct Comm {
  int rank = 0, nranks = 1;
  virtual ~Comm() {}
  // ghost rows: `count` floats per segment go from send_up[i] to rank-1 and from send_dn[i] to rank+1; recv_up[i] is
  // filled by rank-1's send_dn[i], recv_dn[i] by rank+1's send_up[i].  Ranks 0 / n-1 have no upper / lower neighbour.
  virtual int exchange(int nseg, const float* const* send_up, float* const* recv_up, const float* const* send_dn,
                       float* const* recv_dn, size_t count, cudaStream_t s) = 0;
  virtual int allreduce_sum(double* p, size_t n, cudaStream_t s) = 0;
  virtual int allreduce_max_u32(unsigned* p, size_t n, cudaStream_t s) = 0;
  virtual int allgather(const float* send, float* recv, size_t count, cudaStream_t s) = 0;   // recv: [nranks][count]
  virtual const char* error() const = 0;
  virtual bool uses_sms() const { return false; }   // true: exchanges run as kernels (NCCL) and need free SMs to overlap
  virtual int peer_allreduce() const { return 0; }  // 1: the scalar all-reduces run as our own kernel over NVLink peer memory
};
struct LocalGroup;
int nccl_unique_id(void* out128, std::string* err);
Comm* make_nccl_comm(int rank, int nranks, const void* uid128, std::string* err);
LocalGroup* make_local_group(int n);
void free_local_group(LocalGroup* g);
void abort_local_group(LocalGroup* g);   // a rank failed: wake the peers blocked in a collective and make them fail too
Comm* make_local_comm(LocalGroup* g, int rank, std::string* err);

// persistent TMA-fed fused Liu-Shen block (ofri_ls_tma.cu): T sweeps ui, vi -> uo, vo; false = not applicable
bool launch_ls_tma(int T, const Img& ui, const Img& vi, const Img& uo, const Img& vo, const LsPlanes& co, float hpar,
                   int k0, int maxiter, double tol, double* errs, const LsBand& band, cudaStream_t s, int variant = 8);

// cycles per phase of the persistent kernels since the last read (all zero unless built with -DOFRI_PHASE_TIMING)
void hs_tma_phase_read(unsigned long long* out8);
void ls_tma_phase_read(unsigned long long* out8);

// ---- Farneback adapter (ofri_farneback.cu) ------------------------------------------------------------------------------
struct FbWorkspace {
  Img flow[2][2];      // current flow of the internal level (ping-pong between levels)
  Img blur, level, tmp, poly[3];
  Img R[2][5], M[3][5];
  Img* d_imgs = nullptr;   // device scratch for 15 Img structs (plane tables of the filter kernels)
};
int launch_farneback(const Img& im1, const Img& im2, const Img& u_io, const Img& v_io, const ofri_farneback_params* fp,
                     const FbWorkspace& ws, const std::function<int(int, int, ResizeTaps*)>& resize_taps, cudaStream_t s,
                     LaunchCounter& lc);
// dense Lucas-Kanade adapter (ofri_lk.cu): u_io / v_io initial flow in, refined flow out
int launch_lk(const Img& im1, const Img& im2, const Img& u_io, const Img& v_io, const ofri_lk_params* lp, cudaStream_t s,
              LaunchCounter& lc);

const char* kernel_build_info();

}  // namespace ofri
