// ofri_internal.h -- launcher interface between the C-ABI / driver (ofri_api.cu) and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ofri.h"

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
struct SplineSys {
  const double* lo = nullptr;
  const double* cp = nullptr;
  const double* den = nullptr;
  int n = 0;
};

struct LaunchCounter { int64_t n = 0; };

// ---- stages (ofri_stages.cu) ---------------------------------------------------------------------------------
void launch_gauss(const Img& in, const Img& tmp, const Img& out, const GaussTaps& taps, cudaStream_t s, LaunchCounter& lc);
void launch_resize(const Img& in, const Img& tmp, const Img& out, const ResizeTaps& tx, const ResizeTaps& ty,
                   cudaStream_t s, LaunchCounter& lc);
// up-sample `in` (h x w) to `out` (H x W) and multiply by mul; scratch: M1 (h x w), T1 (H x w), M2 (H x w) f64
void launch_spline(const Img& in, const Img& out, float mul, const SplineSys& sy, const SplineSys& sx,
                   const ImgD& M1, const ImgD& T1, const ImgD& M2, cudaStream_t s, LaunchCounter& lc);
void launch_warp_pair(const Img& im1, const Img& im2, const Img& us, const Img& vs, const Img& out1, const Img& out2,
                      cudaStream_t s, LaunchCounter& lc);
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
                      LaunchCounter& lc);
// packed-f32x2 register-resident fused sweeps (ofri_hs_pk.cu): T sweeps ui,vi -> uo,vo on prepared (a, b, c) planes
void launch_hs_packed(int T, int variant, const Img& ui, const Img& vi, const Img& uo, const Img& vo, const Img& fx,
                      const Img& fy, const Img& ft, cudaStream_t s);
// persistent TMA-fed register-resident fused sweeps (ofri_hs_tma.cu); false = not applicable, use another kernel
bool launch_hs_tma(int T, int variant, bool precise, const Img& ui, const Img& vi, const Img& uo, const Img& vo,
                   const Img& fx, const Img& fy, const Img& ft, float alpha2, cudaStream_t s);
// err[b] = (sqrt(sum (u-u0)^2) + sqrt(sum (v-v0)^2)) / (H*W); u0.p == nullptr means u0 = v0 = 0.  acc: [batch][2] f64 scratch
void launch_hs_error(const Img& u, const Img& v, const Img& u0, const Img& v0, double* acc, float* err, int err_stride,
                     cudaStream_t s, LaunchCounter& lc);

// ---- Liu-Shen (ofri_ls.cu) -----------------------------------------------------------------------------------------
struct LsPlanes { Img c[8]; };   // IIx, IIy, II, Ixt, Iyt, B11, B12, B22
// maxenc: [batch][2] uint32 scratch (ordered-int encoded maxima of im1 / im2)
void launch_ls_coefficients(const Img& im1, const Img& im2, float hpar, const LsPlanes& coef, unsigned* maxenc,
                            cudaStream_t s, LaunchCounter& lc);
// Runs up to maxiter sweeps with the reference's stopping rule per pair.  (ua,va) holds the initial guess
// (ROW component in ua!), (ub,vb) is the ping-pong partner; the final state is copied into (uo, vo).
// errs: [batch][maxiter][2] f64 scratch; err_out[b*err_stride] = last total_error; iters_out[b] = sweeps run.
void launch_ls_solve(const Img& ua, const Img& va, const Img& ub, const Img& vb, const LsPlanes& coef, float hpar,
                     int maxiter, double tol, int fuse, int variant, double* errs, int* state, const Img& uo,
                     const Img& vo, float* err_out, int err_stride, int* iters_out, cudaStream_t s, LaunchCounter& lc);

const char* kernel_build_info();

}  // namespace ofri
