// ofri_farneback.cu -- the Farneback adapter of the reference as sm_100a kernels.
//
// Replaces Farneback_PyCL.compute (Farneback_PyCL.py:462-604) and the six OpenCL kernels it drives
// (optical_flow_farneback.cl:72-429): Gaussian pre-blur of the frames, Pillow-BILINEAR level resample, polynomial
// expansion, the G / h matrices from the displaced expansions, their Gaussian (or box) window average, and the 2 x 2
// solve per pixel.  Every arithmetic expression is written in the order of the reference's kernels with separately
// rounded f32 multiplies and adds (ofri_pixel.cuh: fmul / fadd / fsub), so the results are bit-identical to the CPU
// restatement oracle/ofri_farneback_oracle.py (the reference's own OpenCL path cannot run in this image, see there).
// The separable filters run as a vertical pass into a float32 plane followed by a horizontal pass -- the same two
// roundings as the reference's shared-memory row cache.  All kernels take plane stacks ([batch][H][pitch]); the five
// coefficient planes of an expansion / matrix are five Img structs instead of one 5H x W array.
#include "ofri_internal.h"
#include "ofri_pixel.cuh"

namespace ofri {

namespace {

struct FbTaps { float k[OFRI_FB_MAX_HALF + 1]; int kh; };
struct FbPoly { float g[8], xg[8], xxg[8], ig[4]; int n; };
struct Img5 { Img p[5]; };
struct Img3 { Img p[3]; };

// border rules of the reference's kernels (optical_flow_farneback.cl:134-157): reflect-101 with a modulo guard
__device__ __forceinline__ int idx_low(int i, int last) { return abs(i) % (last + 1); }
__device__ __forceinline__ int idx_high(int i, int last) { return abs(last - abs(last - i)) % (last + 1); }
__device__ __forceinline__ int idx_refl(int i, int last) { return idx_low(idx_high(i, last), last); }

inline dim3 grid3(int W, int H, int Z, dim3 b) { return dim3((W + b.x - 1) / b.x, (H + b.y - 1) / b.y, Z); }

// ---- separable window filters: MODE 0 = Gaussian taps + reflect-101 (gaussianBlur, gaussianBlur5), 1 = box + clamp ----------
// vertical pass: out(y, x) = src(y, x) k0 + sum_j (src(lo_j, x) + src(hi_j, x)) k_j      (CL:173-181, 218-232, 370-384)
template <int MODE>
__global__ void fb_filter_v_kernel(const Img* __restrict__ src, const Img* __restrict__ dst, int nplanes, FbTaps t) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const int pl = blockIdx.z % nplanes, b = blockIdx.z / nplanes;
  const Img s = src[pl], d = dst[pl];
  if (x >= s.W || y >= s.H) return;
  const float* p = s.p + (long)b * s.stride + x;
  float acc;
  if (MODE == 0) {
    acc = fmul(p[(long)y * s.pitch], t.k[0]);
    if (y - t.kh >= 0 && y + t.kh <= s.H - 1) {        // interior rows: no border mapping (its modulo guards cost more than the taps)
      const float* lo = p + (long)(y - 1) * s.pitch;
      const float* hi = p + (long)(y + 1) * s.pitch;
#pragma unroll 4
      for (int j = 1; j <= t.kh; ++j, lo -= s.pitch, hi += s.pitch) acc = fadd(acc, fmul(fadd(*lo, *hi), t.k[j]));
    } else {
      for (int j = 1; j <= t.kh; ++j) {
        const float a = p[(long)idx_low(y - j, s.H - 1) * s.pitch], c = p[(long)idx_high(y + j, s.H - 1) * s.pitch];
        acc = fadd(acc, fmul(fadd(a, c), t.k[j]));
      }
    }
  } else {
    acc = p[(long)y * s.pitch];
    for (int j = 1; j <= t.kh; ++j) {
      const int lo = y - j < 0 ? 0 : y - j, hi = y + j > s.H - 1 ? s.H - 1 : y + j;
      acc = fadd(acc, fadd(p[(long)lo * s.pitch], p[(long)hi * s.pitch]));
    }
  }
  d.p[(long)b * d.stride + (long)y * d.pitch + x] = acc;
}
// horizontal pass over the vertical results; the columns beyond the image are the vertical results of the border-mapped
// columns (the reference fills its row cache from xExt = idx_col(x) / clamp(x))                     (CL:185-195, 236-253)
template <int MODE>
__global__ void fb_filter_h_kernel(const Img* __restrict__ src, const Img* __restrict__ dst, int nplanes, FbTaps t,
                                   float box_inv) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const int pl = blockIdx.z % nplanes, b = blockIdx.z / nplanes;
  const Img s = src[pl], d = dst[pl];
  if (x >= s.W || y >= s.H) return;
  const float* p = s.p + (long)b * s.stride + (long)y * s.pitch;
  float acc;
  if (MODE == 0) {
    acc = fmul(p[x], t.k[0]);
    if (x - t.kh >= 0 && x + t.kh <= s.W - 1) {        // interior columns
#pragma unroll 4
      for (int i = 1; i <= t.kh; ++i) acc = fadd(acc, fmul(fadd(p[x - i], p[x + i]), t.k[i]));
    } else {
      for (int i = 1; i <= t.kh; ++i)
        acc = fadd(acc, fmul(fadd(p[idx_refl(x - i, s.W - 1)], p[idx_refl(x + i, s.W - 1)]), t.k[i]));
    }
  } else {
    acc = p[x];
    for (int i = 1; i <= t.kh; ++i) {
      const int lo = x - i < 0 ? 0 : x - i, hi = x + i > s.W - 1 ? s.W - 1 : x + i;
      acc = fadd(acc, fadd(p[lo], p[hi]));
    }
    acc = fmul(acc, box_inv);
  }
  d.p[(long)b * d.stride + (long)y * d.pitch + x] = acc;
}

// Register sliding-window forms of the Gaussian passes for a compile-time half width KH: a thread produces RPT consecutive
// outputs along the filter axis from RPT + 2 KH inputs it loads once (5 instead of 33 loads per output at KH = 16), each
// output accumulated in exactly the order of the kernels above (-> the same bits).  Threads whose outputs need the border
// mapping take the per-output path.
template <int KH, int RPT>
__global__ void __launch_bounds__(256) fb_filter_v_sw_kernel(const Img* __restrict__ src, const Img* __restrict__ dst,
                                                             int nplanes, FbTaps t) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y0 = (blockIdx.y * blockDim.y + threadIdx.y) * RPT;
  const int pl = blockIdx.z % nplanes, b = blockIdx.z / nplanes;
  const Img s = src[pl], d = dst[pl];
  if (x >= s.W || y0 >= s.H) return;
  const float* p = s.p + (long)b * s.stride + x;
  float* o = d.p + (long)b * d.stride + x;
  if (y0 - KH >= 0 && y0 + RPT - 1 + KH <= s.H - 1) {
    float v[RPT + 2 * KH];
    const float* q = p + (long)(y0 - KH) * s.pitch;
#pragma unroll
    for (int i = 0; i < RPT + 2 * KH; ++i, q += s.pitch) v[i] = *q;
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      float acc = fmul(v[r + KH], t.k[0]);
#pragma unroll
      for (int j = 1; j <= KH; ++j) acc = fadd(acc, fmul(fadd(v[r + KH - j], v[r + KH + j]), t.k[j]));
      o[(long)(y0 + r) * d.pitch] = acc;
    }
    return;
  }
  for (int r = 0; r < RPT && y0 + r < s.H; ++r) {
    const int y = y0 + r;
    float acc = fmul(p[(long)y * s.pitch], t.k[0]);
    for (int j = 1; j <= KH; ++j) {
      const float a = p[(long)idx_low(y - j, s.H - 1) * s.pitch], c = p[(long)idx_high(y + j, s.H - 1) * s.pitch];
      acc = fadd(acc, fmul(fadd(a, c), t.k[j]));
    }
    o[(long)y * d.pitch] = acc;
  }
}
// horizontal: a warp owns 32 RPT consecutive outputs of one row; the row segment is staged in shared memory (coalesced),
// skewed by one float per 32 so that the lanes' RPT-strided windows fall into different banks
template <int KH, int RPT>
__global__ void __launch_bounds__(256) fb_filter_h_sw_kernel(const Img* __restrict__ src, const Img* __restrict__ dst,
                                                             int nplanes, FbTaps t) {
  constexpr int SEG = 32 * RPT, NIN = SEG + 2 * KH, NSK = NIN + NIN / 32 + 1;
  __shared__ float sm[8][NSK];
  const int lane = threadIdx.x, wy = threadIdx.y;
  const int xs = blockIdx.x * SEG, y = blockIdx.y * blockDim.y + wy;
  const int pl = blockIdx.z % nplanes, b = blockIdx.z / nplanes;
  const Img s = src[pl], d = dst[pl];
  if (y >= s.H || xs >= s.W) return;                 // warp-uniform
  const float* p = s.p + (long)b * s.stride + (long)y * s.pitch;
  float* o = d.p + (long)b * d.stride + (long)y * d.pitch;
  float* row = sm[wy];
  for (int e = lane; e < NIN; e += 32) row[e + (e >> 5)] = p[idx_refl(xs - KH + e, s.W - 1)];
  __syncwarp();
  const int x0 = xs + lane * RPT;
  if (x0 >= s.W) return;
  float v[RPT + 2 * KH];
#pragma unroll
  for (int i = 0; i < RPT + 2 * KH; ++i) {
    const int e = lane * RPT + i;
    v[i] = row[e + (e >> 5)];
  }
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    float acc = fmul(v[r + KH], t.k[0]);
#pragma unroll
    for (int j = 1; j <= KH; ++j) acc = fadd(acc, fmul(fadd(v[r + KH - j], v[r + KH + j]), t.k[j]));
    if (x0 + r < s.W) o[x0 + r] = acc;
  }
}

// ---- polynomial expansion (CL:72-132) ------------------------------------------------------------------------------------------
// vertical pass: the three row-cache planes row[0], row[bdx], row[2 bdx]; rows clamped (replicate)
__global__ void fb_poly_v_kernel(Img src, Img3 r, FbPoly c) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= src.W || y >= src.H) return;
  const float* p = src.p + (long)b * src.stride + x;
  float r0 = fmul(p[(long)y * src.pitch], c.g[0]), r1 = 0.0f, r2 = 0.0f;
  for (int k = 1; k <= c.n; ++k) {
    const int lo = y - k < 0 ? 0 : y - k, hi = y + k > src.H - 1 ? src.H - 1 : y + k;
    const float t0 = p[(long)lo * src.pitch], t1 = p[(long)hi * src.pitch];
    r0 = fadd(r0, fmul(c.g[k], fadd(t0, t1)));
    r1 = fadd(r1, fmul(c.xg[k], fsub(t1, t0)));
    r2 = fadd(r2, fmul(c.xxg[k], fadd(t0, t1)));
  }
  const long o = (long)b * r.p[0].stride + (long)y * r.p[0].pitch + x;
  r.p[0].p[o] = r0;
  r.p[1].p[o] = r1;
  r.p[2].p[o] = r2;
}
// horizontal pass: b1..b6 and the five coefficient planes; columns clamped (xWarped)
__global__ void fb_poly_h_kernel(Img3 r, Img5 dst, FbPoly c) {
  const int W = dst.p[0].W, H = dst.p[0].H;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= W || y >= H) return;
  const long ro = (long)b * r.p[0].stride + (long)y * r.p[0].pitch;
  const float* q0 = r.p[0].p + ro;
  const float* q1 = r.p[1].p + ro;
  const float* q2 = r.p[2].p + ro;
  float b1 = fmul(c.g[0], q0[x]), b3 = fmul(c.g[0], q1[x]), b5 = fmul(c.g[0], q2[x]);
  float b2 = 0.0f, b4 = 0.0f, b6 = 0.0f;
  for (int k = 1; k <= c.n; ++k) {
    const int xp = x + k > W - 1 ? W - 1 : x + k, xm = x - k < 0 ? 0 : x - k;
    const float s0 = fadd(q0[xp], q0[xm]), d0 = fsub(q0[xp], q0[xm]);
    const float s1 = fadd(q1[xp], q1[xm]), d1 = fsub(q1[xp], q1[xm]);
    const float s2 = fadd(q2[xp], q2[xm]);
    b1 = fadd(b1, fmul(s0, c.g[k]));
    b4 = fadd(b4, fmul(s0, c.xxg[k]));
    b2 = fadd(b2, fmul(d0, c.xg[k]));
    b3 = fadd(b3, fmul(s1, c.g[k]));
    b6 = fadd(b6, fmul(d1, c.xg[k]));
    b5 = fadd(b5, fmul(s2, c.g[k]));
  }
  const long o = (long)b * dst.p[0].stride + (long)y * dst.p[0].pitch + x;
  dst.p[0].p[o] = fmul(b3, c.ig[0]);
  dst.p[1].p[o] = fmul(b2, c.ig[0]);
  dst.p[2].p[o] = fadd(fmul(b1, c.ig[1]), fmul(b5, c.ig[2]));
  dst.p[3].p[o] = fadd(fmul(b1, c.ig[1]), fmul(b4, c.ig[2]));
  dst.p[4].p[o] = fmul(b6, c.ig[3]);
}

// ---- G / h matrices from the displaced expansions (CL:256-348) ------------------------------------------------------------------
__constant__ float c_fb_border[6] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f, 1.0f};

__global__ void fb_update_matrices_kernel(Img fxp, Img fyp, Img5 R0, Img5 R1, Img5 M) {
  const int cols = fxp.W, rows = fxp.H;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= cols || y >= rows) return;
  const float dx = fxp.p[(long)b * fxp.stride + (long)y * fxp.pitch + x];
  const float dy = fyp.p[(long)b * fyp.stride + (long)y * fyp.pitch + x];
  float fx = fadd((float)x, dx), fy = fadd((float)y, dy);
  // floor() of absurd / NaN displacements saturates instead of trapping (the reference's behaviour there is undefined)
  const float flx = fminf(fmaxf(floorf(fx), -1.0e9f), 1.0e9f), fly = fminf(fmaxf(floorf(fy), -1.0e9f), 1.0e9f);
  const int x1 = (int)flx, y1 = (int)fly;
  fx = fsub(fx, (float)x1);
  fy = fsub(fy, (float)y1);
  const long o = (long)b * R0.p[0].stride + (long)y * R0.p[0].pitch + x;
  float r2, r3, r4, r5, r6;
  if (x1 >= 0 && y1 >= 0 && x1 < cols - 1 && y1 < rows - 1) {
    const float a00 = fmul(fsub(1.0f, fx), fsub(1.0f, fy)), a01 = fmul(fx, fsub(1.0f, fy));
    const float a10 = fmul(fsub(1.0f, fx), fy), a11 = fmul(fx, fy);
    const long o1 = (long)b * R1.p[0].stride + (long)y1 * R1.p[0].pitch + x1, q = R1.p[0].pitch;
    auto samp = [&](const Img& im) {
      const float* p = im.p + o1;
      return fadd(fadd(fadd(fmul(a00, p[0]), fmul(a01, p[1])), fmul(a10, p[q])), fmul(a11, p[q + 1]));
    };
    r2 = samp(R1.p[0]);
    r3 = samp(R1.p[1]);
    r4 = samp(R1.p[2]);
    r5 = samp(R1.p[3]);
    r6 = samp(R1.p[4]);
    r4 = fmul(fadd(R0.p[2].p[o], r4), 0.5f);
    r5 = fmul(fadd(R0.p[3].p[o], r5), 0.5f);
    r6 = fmul(fadd(R0.p[4].p[o], r6), 0.25f);
  } else {
    r2 = r3 = 0.0f;
    r4 = R0.p[2].p[o];
    r5 = R0.p[3].p[o];
    r6 = fmul(R0.p[4].p[o], 0.5f);
  }
  r2 = fmul(fsub(R0.p[0].p[o], r2), 0.5f);
  r3 = fmul(fsub(R0.p[1].p[o], r3), 0.5f);
  r2 = fadd(r2, fadd(fmul(r4, dy), fmul(r6, dx)));      // r2 += r4*dy + r6*dx: the right-hand side is summed first
  r3 = fadd(r3, fadd(fmul(r6, dy), fmul(r5, dx)));
  const int bs = 5;
  const float scale = fmul(fmul(fmul(c_fb_border[x < bs ? x : bs], c_fb_border[y < bs ? y : bs]),
                                c_fb_border[cols - x - 1 < bs ? cols - x - 1 : bs]),
                           c_fb_border[rows - y - 1 < bs ? rows - y - 1 : bs]);
  r2 = fmul(r2, scale);
  r3 = fmul(r3, scale);
  r4 = fmul(r4, scale);
  r5 = fmul(r5, scale);
  r6 = fmul(r6, scale);
  const long om = (long)b * M.p[0].stride + (long)y * M.p[0].pitch + x;
  M.p[0].p[om] = fadd(fmul(r4, r4), fmul(r6, r6));
  M.p[1].p[om] = fmul(fadd(r4, r5), r6);
  M.p[2].p[om] = fadd(fmul(r5, r5), fmul(r6, r6));
  M.p[3].p[om] = fadd(fmul(r4, r2), fmul(r6, r3));
  M.p[4].p[om] = fadd(fmul(r6, r2), fmul(r5, r3));
}

// ---- 2 x 2 solve per pixel (CL:408-429) -------------------------------------------------------------------------------------------
__global__ void fb_update_flow_kernel(Img5 M, Img fxp, Img fyp) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= fxp.W || y >= fxp.H) return;
  const long o = (long)b * M.p[0].stride + (long)y * M.p[0].pitch + x;
  const float g11 = M.p[0].p[o], g12 = M.p[1].p[o], g22 = M.p[2].p[o], h1 = M.p[3].p[o], h2 = M.p[4].p[o];
  const float det_inv = fdiv(1.0f, fadd(fsub(fmul(g11, g22), fmul(g12, g12)), 1e-3f));
  fxp.p[(long)b * fxp.stride + (long)y * fxp.pitch + x] = fmul(fsub(fmul(g11, h2), fmul(g12, h1)), det_inv);
  fyp.p[(long)b * fyp.stride + (long)y * fyp.pitch + x] = fmul(fsub(fmul(g22, h1), fmul(g12, h2)), det_inv);
}

}  // namespace

// device copies of small Img arrays: the filter kernels index planes by blockIdx.z
static void upload_imgs(const Img* host, int n, Img* dev, cudaStream_t s) {
  cudaMemcpyAsync(dev, host, sizeof(Img) * n, cudaMemcpyHostToDevice, s);
}

// One window / pre-blur pass of `n` planes: src -> tmp (vertical) -> dst (horizontal).  d_imgs: device scratch for 3 n Img.
static void fb_filter(const Img* src, const Img* tmp, const Img* dst, int n, const float* taps, int kh, bool box,
                      Img* d_imgs, cudaStream_t s, LaunchCounter& lc) {
  FbTaps t;
  t.kh = kh;
  for (int i = 0; i <= OFRI_FB_MAX_HALF; ++i) t.k[i] = (!box && i <= kh) ? taps[i] : 0.0f;
  std::vector<Img> all(src, src + n);
  all.insert(all.end(), tmp, tmp + n);
  all.insert(all.end(), dst, dst + n);
  upload_imgs(all.data(), 3 * n, d_imgs, s);
  dim3 b(32, 8), g = grid3(src[0].W, src[0].H, n * src[0].batch, b);
  if (box) {
    volatile float area = (float)((1 + 2 * kh) * (1 + 2 * kh));
    volatile float inv = 1.0f / area;
    fb_filter_v_kernel<1><<<g, b, 0, s>>>(d_imgs, d_imgs + n, n, t);
    fb_filter_h_kernel<1><<<g, b, 0, s>>>(d_imgs + n, d_imgs + 2 * n, n, t, inv);
  } else if (kh == 16 && src[0].H > 2 * kh && src[0].W > 2 * kh) {   // the default 33-tap window: sliding-window forms
    constexpr int RPT = 8;
    dim3 gv((src[0].W + 31) / 32, (src[0].H + 8 * RPT - 1) / (8 * RPT), n * src[0].batch);
    dim3 gh((src[0].W + 32 * RPT - 1) / (32 * RPT), (src[0].H + 7) / 8, n * src[0].batch);
    fb_filter_v_sw_kernel<16, RPT><<<gv, b, 0, s>>>(d_imgs, d_imgs + n, n, t);
    fb_filter_h_sw_kernel<16, RPT><<<gh, b, 0, s>>>(d_imgs + n, d_imgs + 2 * n, n, t);
  } else {
    fb_filter_v_kernel<0><<<g, b, 0, s>>>(d_imgs, d_imgs + n, n, t);
    fb_filter_h_kernel<0><<<g, b, 0, s>>>(d_imgs + n, d_imgs + 2 * n, n, t, 0.0f);
  }
  lc.n += 2;
}

// Farneback_PyCL.compute for `batch` pairs resident on the device.  u_io / v_io: the initial flow in, the result out.
// ws: scratch planes of the frames' size, at least 30 of them (FbWorkspace in ofri_api.cu); resize_taps(in, out, &t)
// hands out Pillow BILINEAR tap tables.
int launch_farneback(const Img& im1, const Img& im2, const Img& u_io, const Img& v_io, const ofri_farneback_params* fp,
                     const FbWorkspace& ws, const std::function<int(int, int, ResizeTaps*)>& resize_taps, cudaStream_t s,
                     LaunchCounter& lc) {
  const int H = im1.H, W = im1.W;
  // crop unnecessary pyramid levels (FB:483-489)
  const int min_size = 32;
  double scale = 1.0;
  int levels = 0;
  while (levels < fp->extra_levels) {
    scale *= fp->pyr_scale;
    if (W * scale < min_size || H * scale < min_size) break;
    ++levels;
  }
  FbPoly pc;
  pc.n = fp->poly_n;
  for (int i = 0; i < 8; ++i) { pc.g[i] = fp->g[i]; pc.xg[i] = fp->xg[i]; pc.xxg[i] = fp->xxg[i]; }
  for (int i = 0; i < 4; ++i) pc.ig[i] = fp->ig[i];
  auto vw = [](const Img& base, int h, int w) {
    Img m = base;
    m.H = h; m.W = w;
    m.pitch = (w + 3) / 4 * 4;
    m.stride = m.pitch * h;
    return m;
  };
  Img prev_x, prev_y;
  int prev_h = 0, prev_w = 0;
  dim3 b(32, 8);
  for (int k = levels; k >= 0; --k) {
    scale = 1.0;
    for (int i = 0; i < k; ++i) scale *= fp->pyr_scale;
    const int width = (int)std::nearbyint(W * scale), height = (int)std::nearbyint(H * scale);   // Python round(): half-even
    if (width < 1 || height < 1) return OFRI_ERR_TOO_SMALL;
    const int kb = fp->n_blur[k];                        // int(smoothSize / 2) of level k
    Img cx = vw(ws.flow[k & 1][0], height, width), cy = vw(ws.flow[k & 1][1], height, width);
    // initial flow of the level: the caller's (U, V) resized and scaled at the coarsest level, else the previous level's
    {
      const Img& sx = prev_h ? prev_x : u_io;
      const Img& sy = prev_h ? prev_y : v_io;
      const int sh = prev_h ? prev_h : H, sw = prev_h ? prev_w : W;
      const float mul = prev_h ? (float)(1.0 / fp->pyr_scale) : (float)scale;
      ResizeTaps tx, ty;
      int rc = resize_taps(sw, width, &tx);
      if (rc) return rc;
      rc = resize_taps(sh, height, &ty);
      if (rc) return rc;
      launch_resize(sx, vw(ws.tmp, sh, width), cx, tx, ty, s, lc);
      launch_resize(sy, vw(ws.tmp, sh, width), cy, tx, ty, s, lc);
      if (mul != 1.0f) {
        launch_scale(cx, mul, s, lc);
        launch_scale(cy, mul, s, lc);
      }
    }
    // blurred frames -> level size -> polynomial expansions RA, RB (FB:560-585)
    Img5 R[2];
    for (int f = 0; f < 2; ++f) {
      const Img& frame = f ? im2 : im1;
      Img blurred = vw(ws.blur, H, W), vt = vw(ws.tmp, H, W);
      fb_filter(&frame, &vt, &blurred, 1, fp->blur_kernel[k], kb, false, ws.d_imgs, s, lc);
      Img lvl = vw(ws.level, height, width);
      ResizeTaps tx, ty;
      int rc = resize_taps(W, width, &tx);
      if (rc) return rc;
      rc = resize_taps(H, height, &ty);
      if (rc) return rc;
      launch_resize(blurred, vw(ws.tmp, H, width), lvl, tx, ty, s, lc);
      Img3 r3;
      for (int i = 0; i < 3; ++i) r3.p[i] = vw(ws.poly[i], height, width);
      for (int i = 0; i < 5; ++i) R[f].p[i] = vw(ws.R[f][i], height, width);
      const dim3 g = grid3(width, height, lvl.batch, b);
      fb_poly_v_kernel<<<g, b, 0, s>>>(lvl, r3, pc);
      fb_poly_h_kernel<<<g, b, 0, s>>>(r3, R[f], pc);
      lc.n += 2;
    }
    Img5 M, Mb, Mt;
    for (int i = 0; i < 5; ++i) {
      M.p[i] = vw(ws.M[0][i], height, width);
      Mb.p[i] = vw(ws.M[1][i], height, width);
      Mt.p[i] = vw(ws.M[2][i], height, width);
    }
    const dim3 g = grid3(width, height, cx.batch, b);
    fb_update_matrices_kernel<<<g, b, 0, s>>>(cx, cy, R[0], R[1], M);
    lc.n += 1;
    const int wh = fp->window_size / 2;
    for (int it = 0; it < fp->n_iters; ++it) {
      fb_filter(M.p, Mt.p, Mb.p, 5, fp->win_kernel, wh, !fp->use_gaussian, ws.d_imgs, s, lc);      // bufM = blur5(M)
      std::swap(M, Mb);
      fb_update_flow_kernel<<<g, b, 0, s>>>(M, cx, cy);
      lc.n += 1;
      if (it < fp->n_iters - 1) {
        fb_update_matrices_kernel<<<g, b, 0, s>>>(cx, cy, R[0], R[1], M);
        lc.n += 1;
      }
    }
    prev_x = cx;
    prev_y = cy;
    prev_h = height;
    prev_w = width;
  }
  // level 0 has the frames' size: the result replaces the initial flow
  launch_copy(u_io, prev_x, s, lc);
  launch_copy(v_io, prev_y, s, lc);
  return cudaPeekAtLastError() == cudaSuccess ? OFRI_OK : OFRI_ERR_CUDA;
}

}  // namespace ofri
