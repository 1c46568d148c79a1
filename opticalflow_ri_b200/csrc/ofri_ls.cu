// ofri_ls.cu -- Liu-Shen physics-based optical-flow kernels (sm_100a): max-normalisation, coefficient planes,
// the Jacobi-type sweep (simple per-pixel kernel and temporally blocked shared-memory kernel) with the
// per-sweep residual norm, and the device-side stopping rule.
//
// Reference: PhysicsBasedOpticalFlowLiuShen.py:47-158.  Stopping rule (LS:141): sweep k runs iff k < maxnum and
// total_error_{k-1} > tol, total_error = (||unew-u||_2 + ||vnew-v||_2) / (r c).  Here every sweep's
// (sum du^2, sum dv^2) is accumulated per pair in errs[pair][k][2] (f64, warp-shuffle + one atomicAdd per CTA);
// a launch first looks at the previous sweep's sums and does nothing for pairs that have stopped, so no host
// synchronisation is needed.  When T sweeps are fused per launch and the rule trips inside a block, the block
// has overshot by < T sweeps: ls_finalize_kernel then schedules a replay of the exact count from the block's
// input buffer (still intact thanks to the ping-pong), executed by up to T-1 conditional single-sweep launches.
//
// Algorithmic HBM traffic of a sweep launch: 48 B per pixel (read u, v and 8 coefficient planes; write u, v).
#include "ofri_internal.h"
#include "ofri_pixel.cuh"
#include "ofri_ls_common.cuh"

namespace ofri {

// ---------------------------------------------------------------------------------------------------------------
// max(im1), max(im2) per pair  (LS:96-97)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned enc_ordered(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(unsigned e) {
  unsigned u = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
  return __uint_as_float(u);
}
__global__ void ls_max_kernel(Img a, Img c, unsigned* maxenc) {
  const int b = blockIdx.z;
  float ma = -INFINITY, mc = -INFINITY;
  for (int y = blockIdx.y; y < a.H; y += gridDim.y)
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < a.W; x += gridDim.x * blockDim.x) {
      ma = fmaxf(ma, a.p[(long)b * a.stride + (long)y * a.pitch + x]);
      mc = fmaxf(mc, c.p[(long)b * c.stride + (long)y * c.pitch + x]);
    }
  for (int o = 16; o > 0; o >>= 1) {
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
    mc = fmaxf(mc, __shfl_xor_sync(0xffffffffu, mc, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(maxenc + 2 * b, enc_ordered(ma));
    atomicMax(maxenc + 2 * b + 1, enc_ordered(mc));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// coefficient planes (LS:96-97, 124-128, generate_invmatrix LS:47-73)
// ---------------------------------------------------------------------------------------------------------------
// row0 / Hg: the planes hold rows [row0, row0 + H) of an image of Hg rows (row bands; 0 / H otherwise): the count of
// in-bounds neighbours (8 / 5 / 3) refers to the IMAGE border, not the band's
__global__ void __launch_bounds__(256)
ls_coef_kernel(Img im1, Img im2, float hpar, LsPlanes co, const unsigned* maxenc, int row0, int Hg) {
  // the block's 32 x 8 pixels + a clamped 1-pixel frame are normalised ONCE into shared memory (a division per
  // element instead of one per use: 18 per pixel), then every thread reads its 3 x 3 neighbourhoods from there
  __shared__ float sa[10][34], sd[10][34];
  const int b = blockIdx.z;
  const int W = im1.W, H = im1.H;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  const float m1 = dec_ordered(maxenc[2 * b]), m2 = dec_ordered(maxenc[2 * b + 1]);
  const float* A = im1.p + (long)b * im1.stride;
  const float* B = im2.p + (long)b * im2.stride;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  for (int i = tid; i < 340; i += 256) {
    const int sy = i / 34, sx = i - sy * 34;
    const int yy = clampi(y0 + sy - 1, 0, H - 1), xx = clampi(x0 + sx - 1, 0, W - 1);
    const float i1 = fdiv(A[(long)yy * im1.pitch + xx], m1);
    const float i2 = fdiv(B[(long)yy * im2.pitch + xx], m2);
    sa[sy][sx] = i1;
    sd[sy][sx] = fsub(i2, i1);
  }
  __syncthreads();
  const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
  if (x >= W || y >= H) return;
  float a[3][3], d[3][3];
  int cnt = 0;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      int yy = y + r - 1, xx = x + q - 1;
      bool in = (yy + row0 >= 0) && (yy + row0 < Hg) && (xx >= 0) && (xx < W);
      if (in && !(r == 1 && q == 1)) ++cnt;
      a[r][q] = sa[threadIdx.y + r][threadIdx.x + q];
      d[r][q] = sd[threadIdx.y + r][threadIdx.x + q];
    }
  LsCoef c = ls_coef_point(a, d, hpar, (float)cnt);
  long o = (long)b * co.c[0].stride + (long)y * co.c[0].pitch + x;
  co.c[0].p[o] = c.IIx;
  co.c[1].p[o] = c.IIy;
  co.c[2].p[o] = c.II;
  co.c[3].p[o] = c.Ixt;
  co.c[4].p[o] = c.Iyt;
  co.c[5].p[o] = c.B11;
  co.c[6].p[o] = c.B12;
  co.c[7].p[o] = c.B22;
}
void launch_ls_max(const Img& im1, const Img& im2, unsigned* maxenc, cudaStream_t s, LaunchCounter& lc) {
  cudaMemsetAsync(maxenc, 0, sizeof(unsigned) * 2 * im1.batch, s);   // 0 encodes below every float
  int gy = im1.H < 64 ? im1.H : 64;
  int gx = (im1.W + 255) / 256;
  if (gx > 4) gx = 4;
  ls_max_kernel<<<dim3(gx, gy, im1.batch), 256, 0, s>>>(im1, im2, maxenc);
  lc.n += 1;
}
void launch_ls_coef(const Img& im1, const Img& im2, float hpar, const LsPlanes& coef, const unsigned* maxenc,
                    cudaStream_t s, LaunchCounter& lc, int row0, int Hg) {
  dim3 b(32, 8), g((im1.W + 31) / 32, (im1.H + 7) / 8, im1.batch);
  ls_coef_kernel<<<g, b, 0, s>>>(im1, im2, hpar, coef, maxenc, row0, Hg > 0 ? Hg : im1.H);
  lc.n += 1;
}
void launch_ls_coefficients(const Img& im1, const Img& im2, float hpar, const LsPlanes& coef, unsigned* maxenc,
                            cudaStream_t s, LaunchCounter& lc) {
  launch_ls_max(im1, im2, maxenc, s, lc);
  launch_ls_coef(im1, im2, hpar, coef, maxenc, s, lc, 0, 0);
}

// ---------------------------------------------------------------------------------------------------------------
// stopping rule helpers.  state[pair*4 + {0: replay source buffer, 1: replay count, 2: final buffer, 3: iters}]
// ---------------------------------------------------------------------------------------------------------------
// block-level f64 sum of two values, one atomicAdd per CTA
__device__ __forceinline__ void block_atomic_add2(double su, double sv, double* dst, double* sh /* >= 64 */) {
  for (int o = 16; o > 0; o >>= 1) {
    su += __shfl_xor_sync(0xffffffffu, su, o);
    sv += __shfl_xor_sync(0xffffffffu, sv, o);
  }
  const int nthreads = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int w = tid >> 5, l = tid & 31, nw = (nthreads + 31) >> 5;
  if (l == 0) { sh[w] = su; sh[32 + w] = sv; }
  __syncthreads();
  if (w == 0) {
    su = l < nw ? sh[l] : 0.0;
    sv = l < nw ? sh[32 + l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      su += __shfl_xor_sync(0xffffffffu, su, o);
      sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if (l == 0) {
      atomicAdd(dst, su);
      atomicAdd(dst + 1, sv);
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------------
// simple sweep: one thread per pixel, one sweep per launch.
// mode 0: regular sweep k (skipped for stopped pairs).  mode 1: replay step j (runs iff j < state.replay_count;
// buffers chosen from state.replay_source; no error accumulation).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ls_sweep_simple_kernel(Img u0, Img v0, Img u1, Img v1, LsPlanes co, float hpar, int k, int maxiter, double tol,
                       double* errs, const int* state, int mode, int lookback, LsBand band) {
  __shared__ double sh[64];
  __shared__ int s_stop;
  const int b = blockIdx.z;
  const int W = u0.W, H = u0.H;
  const double npix = band.npix;
  Img ui, vi, uo, vo;
  if (mode == 0) {
    if (k > 0) {
      if (threadIdx.x == 0 && threadIdx.y == 0)
        s_stop = ls_stopped_before(errs + (long)b * maxiter * 2, k, tol, npix, lookback) ? 1 : 0;
      __syncthreads();
      if (s_stop) return;
    }
    if (k & 1) { ui = u1; vi = v1; uo = u0; vo = v0; } else { ui = u0; vi = v0; uo = u1; vo = v1; }
  } else {
    const int* st = state + 4 * b;
    if (k >= st[1]) return;
    int src = (st[0] + k) & 1;
    if (src) { ui = u1; vi = v1; uo = u0; vo = v0; } else { ui = u0; vi = v0; uo = u1; vo = v1; }
  }
  // mode 0 is launched with one block per 8 rows (the loop runs once); the replay launches, which almost always return
  // above, use a few blocks per column strip and pair (262 144 empty blocks cost 88 us per launch at 64 pairs of 1024^2)
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  double du2 = 0.0, dv2 = 0.0;
  for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < H; y += gridDim.y * blockDim.y) {
    if (x >= W) break;
    const float* U = ui.p + (long)b * ui.stride;
    const float* V = vi.p + (long)b * vi.stride;
    float uc[3][3], vc[3][3];
    unsigned inb = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        int yy = y + r - 1, xx = x + q - 1;
        if ((yy >= 0) && (yy < H) && (xx >= 0) && (xx < W)) inb |= 1u << (3 * r + q);
        yy = clampi(yy, 0, H - 1);
        xx = clampi(xx, 0, W - 1);
        uc[r][q] = U[(long)yy * ui.pitch + xx];
        vc[r][q] = V[(long)yy * vi.pitch + xx];
      }
    long o = (long)b * co.c[0].stride + (long)y * co.c[0].pitch + x;
    LsCoef c;
    c.IIx = co.c[0].p[o]; c.IIy = co.c[1].p[o]; c.II = co.c[2].p[o]; c.Ixt = co.c[3].p[o];
    c.Iyt = co.c[4].p[o]; c.B11 = co.c[5].p[o]; c.B12 = co.c[6].p[o]; c.B22 = co.c[7].p[o];
    float un, vn;
    ls_update(uc, vc, inb, c, hpar, &un, &vn);
    uo.p[(long)b * uo.stride + (long)y * uo.pitch + x] = un;
    vo.p[(long)b * vo.stride + (long)y * vo.pitch + x] = vn;
    if (y >= band.own_lo && y < band.own_hi) {     // residual over the rows this band owns (all rows normally)
      float eu = fsub(un, uc[1][1]), ev = fsub(vn, vc[1][1]);
      du2 += (double)eu * (double)eu;
      dv2 += (double)ev * (double)ev;
    }
  }
  if (mode == 0) block_atomic_add2(du2, dv2, errs + ((long)b * maxiter + k) * 2, sh);
}

// ---------------------------------------------------------------------------------------------------------------
// fused (temporally blocked) sweep kernel: T sweeps per launch
// ---------------------------------------------------------------------------------------------------------------
// Same geometry as the Horn-Schunck kernel (ofri_hs.cu): shared tile SH x SW, SW = 4 NG, SH = R NRG + 2; thread
// (cg, rg) owns the 4 x R strip at columns 4cg.., rows 1 + rg R .., the same cells in every sweep; NG is 16 or 32 so
// the halo columns come from the neighbouring lanes by shuffle.  Liu-Shen has 8 coefficient planes per pixel -- too
// many for registers -- so they are staged in shared memory next to the two ping-pong buffers of u and v (12 planes,
// 16-byte cp.async).  The residual of every sweep is accumulated in f32 per thread over the CTA's own output cells,
// reduced in f64 (warp shuffle) and added to errs[pair][k] with one atomicAdd per CTA.  The last sweep stores its
// interior results straight to HBM.  EDGE instantiation (tiles touching the image border): 'nearest' clamp for the
// D / F / M stencils and zero padding for the 8-neighbour sum, re-applied every sweep.
__device__ __forceinline__ void cp_async16_ls(void* smem_dst, const void* gsrc, bool valid) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

template <int T, int R, int NRG, int NG>
struct LsCfg {
  static constexpr int HX = 4;
  static constexpr int SW = 4 * NG;
  static constexpr int SH = R * NRG + 2;
  static constexpr int NT = NG * NRG;
  static constexpr int TW = SW - 2 * HX;
  static constexpr int TH = SH - 2 * T;
  static constexpr int PLANE = SH * SW;
  static constexpr int SMEM_BYTES = 12 * PLANE * 4;    // u[2], v[2], 8 coefficient planes
  static_assert((NG == 16 || NG == 32) && NT % 32 == 0 && T <= HX && TW > 0 && TH > 0 && NT <= 1024, "bad tile");
};

template <int T, int R, int NRG, int NG, bool EDGE, bool LAST>
__device__ __forceinline__ void ls_sweep(const float* __restrict__ cu, const float* __restrict__ cv,
                                         float* __restrict__ nu, float* __restrict__ nv, const float* __restrict__ sC,
                                         int r0, int sx, const LsEdge& eg, float hpar, float* __restrict__ gU,
                                         float* __restrict__ gV, long gpitch, int gy0, int gx, int H, int W,
                                         int own_lo, int own_hi, float& du2, float& dv2) {
  using C = LsCfg<T, R, NRG, NG>;
  constexpr int SW = C::SW;
  float wu[3][6], wv[3][6];
  const float* pu = cu + (r0 - 1) * SW + sx;
  const float* pv = cv + (r0 - 1) * SW + sx;
  ls_row6<EDGE>(pu, pv, eg, wu[0], wv[0]);
  ls_row6<EDGE>(pu + SW, pv + SW, eg, wu[1], wv[1]);
  const bool in_cols = (sx >= C::HX) && (sx < SW - C::HX);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int A = j % 3, B = (j + 1) % 3, Cc = (j + 2) % 3;
    ls_row6<EDGE>(pu + (j + 2) * SW, pv + (j + 2) * SW, eg, wu[Cc], wv[Cc]);
    float ou[4], ov[4];
    const int so = (r0 + j) * SW + sx;
    if (EDGE && j == eg.top_j)        // global row 0: 'nearest' -> the row above is the row itself; H8: zero
      ls_row_update<EDGE>(wu[B], wu[B], wu[Cc], wv[B], wv[B], wv[Cc], sC, C::PLANE, so, hpar, true, false, eg, ou, ov);
    else if (EDGE && j == eg.bot_j)
      ls_row_update<EDGE>(wu[A], wu[B], wu[B], wv[A], wv[B], wv[B], sC, C::PLANE, so, hpar, false, true, eg, ou, ov);
    else
      ls_row_update<EDGE>(wu[A], wu[B], wu[Cc], wv[A], wv[B], wv[Cc], sC, C::PLANE, so, hpar, false, false, eg, ou, ov);
    // residual over the CTA's own output cells only (each pixel counted by exactly one CTA)
    const int sy = r0 + j, gy = gy0 + j;
    const bool own = in_cols && (sy >= T) && (sy < C::SH - T) && (gy < H);
    if (own && gy >= own_lo && gy < own_hi) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!EDGE || gx + q < W) {
          float eu = fsub(ou[q], wu[B][q + 1]), ev = fsub(ov[q], wv[B][q + 1]);
          du2 = fmaf(eu, eu, du2);
          dv2 = fmaf(ev, ev, dv2);
        }
      }
    }
    if (!LAST) {
      *reinterpret_cast<float4*>(nu + so) = make_float4(ou[0], ou[1], ou[2], ou[3]);
      *reinterpret_cast<float4*>(nv + so) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    } else if (own && gx < W) {
      const long go = (long)gy * gpitch + gx;
      *reinterpret_cast<float4*>(gU + go) = make_float4(ou[0], ou[1], ou[2], ou[3]);
      *reinterpret_cast<float4*>(gV + go) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    }
  }
}

template <int T, int R, int NRG, int NG, bool EDGE>
__device__ __forceinline__ void ls_fused_body(const Img& ui, const Img& vi, const Img& uo, const Img& vo,
                                              const LsPlanes& co, float hpar, int k0, double* errs_pair, float* smem,
                                              double* sh, int own_lo, int own_hi) {
  using C = LsCfg<T, R, NRG, NG>;
  constexpr int SW = C::SW, SH = C::SH, HX = C::HX;
  const int b = blockIdx.z;
  const int W = ui.W, H = ui.H;
  const int x0 = blockIdx.x * C::TW - HX;
  const int y0 = blockIdx.y * C::TH - T;
  const int tid = threadIdx.x;
  float* sC = smem + 4 * C::PLANE;
  {
    const float* gU = ui.p + (long)b * ui.stride;
    const float* gV = vi.p + (long)b * vi.stride;
    const long cb = (long)b * co.c[0].stride;
    for (int i = tid; i < SH * NG; i += C::NT) {
      int sy = i / NG, sg = i - sy * NG;
      int gy = y0 + sy, gx = x0 + 4 * sg;
      bool ok = (gy >= 0) && (gy < H) && (gx >= 0) && (gx < (int)ui.pitch);
      long go = ok ? (long)gy * ui.pitch + gx : 0;
      int so = sy * SW + 4 * sg;
      cp_async16_ls(smem + so, gU + go, ok);
      cp_async16_ls(smem + 2 * C::PLANE + so, gV + go, ok);
#pragma unroll
      for (int c = 0; c < 8; ++c) cp_async16_ls(sC + c * C::PLANE + so, co.c[c].p + cb + go, ok);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  const int cg = tid % NG, rg = tid / NG;
  const int sx = 4 * cg;
  const int r0 = 1 + rg * R;
  const int gx = x0 + sx;
  LsEdge eg;
  eg.left_edge = EDGE && (gx == 0);
  eg.right_j = EDGE ? (W - 1) - gx : -1;
  eg.top_j = EDGE ? -(y0 + r0) : -1000;
  eg.bot_j = EDGE ? (H - 1) - (y0 + r0) : -1000;
  float* gU = uo.p + (long)b * uo.stride;
  float* gV = vo.p + (long)b * vo.stride;
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();

#pragma unroll 1
  for (int s = 0; s < T - 1; ++s) {
    const float* cu = smem + (s & 1) * C::PLANE;
    const float* cv = smem + (2 + (s & 1)) * C::PLANE;
    float* nu = smem + ((s + 1) & 1) * C::PLANE;
    float* nv = smem + (2 + ((s + 1) & 1)) * C::PLANE;
    float du2 = 0.0f, dv2 = 0.0f;
    ls_sweep<T, R, NRG, NG, EDGE, false>(cu, cv, nu, nv, sC, r0, sx, eg, hpar, gU, gV, uo.pitch, y0 + r0, gx, H, W,
                                         own_lo, own_hi, du2, dv2);
    block_atomic_add2((double)du2, (double)dv2, errs_pair + 2 * (k0 + s), sh);   // also the sweep barrier
  }
  {
    constexpr int s = T - 1;
    const float* cu = smem + (s & 1) * C::PLANE;
    const float* cv = smem + (2 + (s & 1)) * C::PLANE;
    float du2 = 0.0f, dv2 = 0.0f;
    ls_sweep<T, R, NRG, NG, EDGE, true>(cu, cv, nullptr, nullptr, sC, r0, sx, eg, hpar, gU, gV, uo.pitch, y0 + r0, gx, H,
                                        W, own_lo, own_hi, du2, dv2);
    block_atomic_add2((double)du2, (double)dv2, errs_pair + 2 * (k0 + s), sh);
  }
}

template <int T, int R, int NRG, int NG, int MINB>
__global__ void __launch_bounds__(LsCfg<T, R, NRG, NG>::NT, MINB)
ls_fused_kernel(Img u0, Img v0, Img u1, Img v1, LsPlanes co, float hpar, int k0, int maxiter, double tol, double* errs,
                LsBand band) {
  using C = LsCfg<T, R, NRG, NG>;
  extern __shared__ __align__(16) float smem[];
  __shared__ double sh[64];
  __shared__ int s_stop;
  const int b = blockIdx.z;
  const double npix = band.npix;
  double* errs_pair = errs + (long)b * maxiter * 2;
  if (k0 > 0) {   // stopping rule evaluated by one thread (f64 square roots), CTA-uniform result
    if (threadIdx.x == 0) s_stop = ls_stopped_before(errs_pair, k0, tol, npix, T) ? 1 : 0;
    __syncthreads();
    if (s_stop) return;
  }
  // launch index parity selects the ping-pong direction: launches alternate u0->u1, u1->u0
  const bool odd = ((k0 / T) & 1) != 0;   // every earlier launch fused exactly T sweeps
  const Img& ui = odd ? u1 : u0;
  const Img& vi = odd ? v1 : v0;
  const Img& uo = odd ? u0 : u1;
  const Img& vo = odd ? v0 : v1;
  const int x0 = blockIdx.x * C::TW - C::HX, y0 = blockIdx.y * C::TH - T;
  const bool edge = (x0 < 0) || (x0 + C::SW > u0.W) || (y0 < 0) || (y0 + C::SH > u0.H);
  if (edge)
    ls_fused_body<T, R, NRG, NG, true>(ui, vi, uo, vo, co, hpar, k0, errs_pair, smem, sh, band.own_lo, band.own_hi);
  else
    ls_fused_body<T, R, NRG, NG, false>(ui, vi, uo, vo, co, hpar, k0, errs_pair, smem, sh, band.own_lo, band.own_hi);
}

template <int T, int R, int NRG, int NG, int MINB>
static void launch_ls_fused_cfg(const Img& u0, const Img& v0, const Img& u1, const Img& v1, const LsPlanes& co,
                                float hpar, int k0, int maxiter, double tol, double* errs, const LsBand& band,
                                cudaStream_t s) {
  using C = LsCfg<T, R, NRG, NG>;
  auto kern = ls_fused_kernel<T, R, NRG, NG, MINB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  dim3 g((u0.W + C::TW - 1) / C::TW, (u0.H + C::TH - 1) / C::TH, u0.batch);
  kern<<<g, C::NT, C::SMEM_BYTES, s>>>(u0, v0, u1, v1, co, hpar, k0, maxiter, tol, errs, band);
}

// ---------------------------------------------------------------------------------------------------------------
// finalize: per pair, find the number of sweeps the reference would have run and where the state lives.
// launches were: nfull launches of T sweeps starting at 0, then (maxiter - nfull*T) single sweeps.
// ---------------------------------------------------------------------------------------------------------------
__global__ void ls_finalize_kernel(const double* errs, int* state, int batch, int maxiter, double tol, double npix,
                                   int T, int nfull) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const double* e = errs + (long)b * maxiter * 2;
  // reference: sweeps run = smallest n >= 1 with total_error_{n-1} <= tol (or NaN), capped at maxiter
  int n = maxiter;
  for (int k = 0; k < maxiter; ++k) {
    double te = ls_total_error(e + 2 * k, npix);
    if (!(te > tol)) { n = k + 1; break; }
  }
  // which launch contained sweep n-1, how many sweeps that launch executed, and the buffer parity before it
  int src = 0, replay = 0, fin = 0;
  int last = n - 1;
  if (last < nfull * T) {
    int li = last / T;                 // fused launch index; it executed T sweeps, we want (last - li*T + 1)
    int want = last - li * T + 1;
    if (want == T) { fin = (li + 1) & 1; }
    else { src = li & 1; replay = want; fin = (li + want) & 1; }
  } else {
    int li = nfull + (last - nfull * T);   // single-sweep launches after the fused ones
    fin = (li + 1) & 1;
  }
  state[4 * b + 0] = src;
  state[4 * b + 1] = replay;
  state[4 * b + 2] = fin;
  state[4 * b + 3] = n;
}
// copy the final state out (u = ROW component -> vo is written from the v-buffer etc. is handled by the caller's
// choice of uo/vo), and report error / iteration count
__global__ void ls_select_kernel(Img u0, Img v0, Img u1, Img v1, Img uo, Img vo, const int* state, const double* errs,
                                 int maxiter, double npix, float* err_out, int err_stride, int* iters_out) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  const int fin = state[4 * b + 2];
  if (x == 0 && y == 0) {
    int n = state[4 * b + 3];
    if (err_out) err_out[(long)b * err_stride] = (float)ls_total_error(errs + ((long)b * maxiter + (n - 1)) * 2, npix);
    if (iters_out) iters_out[b] = n;
  }
  if (x >= uo.W || y >= uo.H) return;
  const Img& su = fin ? u1 : u0;
  const Img& sv = fin ? v1 : v0;
  uo.p[(long)b * uo.stride + (long)y * uo.pitch + x] = su.p[(long)b * su.stride + (long)y * su.pitch + x];
  vo.p[(long)b * vo.stride + (long)y * vo.pitch + x] = sv.p[(long)b * sv.stride + (long)y * sv.pitch + x];
}

template <int T>
static void launch_ls_fused_T(const Img& u0, const Img& v0, const Img& u1, const Img& v1, const LsPlanes& co, float hpar,
                              int k0, int maxiter, double tol, double* errs, const LsBand& band, cudaStream_t s) {
  launch_ls_fused_cfg<T, 4, 8, 16, 2>(u0, v0, u1, v1, co, hpar, k0, maxiter, tol, errs, band, s);   // 34 x 64 tile
}
static void launch_ls_fused(int T, const Img& u0, const Img& v0, const Img& u1, const Img& v1,
                            const LsPlanes& co, float hpar, int k0, int maxiter, double tol, double* errs,
                            const LsBand& band, cudaStream_t s) {
  switch (T) {
    case 1: launch_ls_fused_T<1>(u0, v0, u1, v1, co, hpar, k0, maxiter, tol, errs, band, s); break;
    case 2: launch_ls_fused_T<2>(u0, v0, u1, v1, co, hpar, k0, maxiter, tol, errs, band, s); break;
    case 3: launch_ls_fused_T<3>(u0, v0, u1, v1, co, hpar, k0, maxiter, tol, errs, band, s); break;
    default: launch_ls_fused_T<4>(u0, v0, u1, v1, co, hpar, k0, maxiter, tol, errs, band, s); break;
  }
}

void launch_ls_solve(const Img& ua, const Img& va, const Img& ub, const Img& vb, const LsPlanes& coef, float hpar,
                     int maxiter, double tol, int fuse, int variant, double* errs, int* state, const Img& uo,
                     const Img& vo,
                     float* err_out, int err_stride, int* iters_out, cudaStream_t s, LaunchCounter& lc,
                     const LsBand* band_in, const LsHook& hook) {
  const int batch = ua.batch;
  LsBand band;
  band.own_lo = 0; band.own_hi = ua.H; band.npix = (double)ua.H * (double)ua.W;
  if (band_in) band = *band_in;
  const double npix = band.npix;
  cudaMemsetAsync(errs, 0, sizeof(double) * 2 * (size_t)maxiter * batch, s);
  bool can_fuse = fuse >= 1 && ua.W >= 2 && ua.H >= 2 && (ua.pitch % 4 == 0) && ua.pitch == va.pitch &&
                  ua.pitch == ub.pitch && ua.pitch == vb.pitch && ((uintptr_t)ua.p % 16 == 0) &&
                  ((uintptr_t)ub.p % 16 == 0) && ((uintptr_t)va.p % 16 == 0) && ((uintptr_t)vb.p % 16 == 0) &&
                  (ua.stride % 4 == 0);
  for (int c = 0; c < 8; ++c)
    can_fuse = can_fuse && coef.c[c].pitch == ua.pitch && coef.c[c].stride == coef.c[0].stride &&
               ((uintptr_t)coef.c[c].p % 16 == 0);
  int T = fuse > 4 ? 4 : fuse;
  int nfull = 0;
  dim3 bs(32, 8), gs((ua.W + 31) / 32, (ua.H + 7) / 8, batch);
  if (can_fuse && T >= 1) {
    nfull = maxiter / T;
    for (int i = 0; i < nfull; ++i) {
      bool done = false;
      if (variant >= 8)      // persistent TMA-fed register-resident kernel (ofri_ls_tma.cu); launch i reads buffer (i & 1)
        done = (i & 1) ? launch_ls_tma(T, ub, vb, ua, va, coef, hpar, i * T, maxiter, tol, errs, band, s, variant)
                       : launch_ls_tma(T, ua, va, ub, vb, coef, hpar, i * T, maxiter, tol, errs, band, s, variant);
      if (!done) launch_ls_fused(T, ua, va, ub, vb, coef, hpar, i * T, maxiter, tol, errs, band, s);
      lc.n += 1;
      // launch i wrote buffer b (odd launches write a); band mode: sum the block's residuals over all bands and refresh
      // the ghost rows of the buffer just written
      if (hook) hook(i * T, T, (i & 1) ? 0 : 1);
    }
  } else {
    T = 1;
  }
  // remaining single sweeps: launch index li = nfull + j processes sweep k = nfull*T + j, reading buffer (li & 1)
  for (int k = nfull * T, li = nfull; k < maxiter; ++k, ++li) {
    // the simple kernel derives the direction from k's parity, so feed it buffers swapped when (li - k) is odd
    if (((li - k) & 1) == 0)
      ls_sweep_simple_kernel<<<gs, bs, 0, s>>>(ua, va, ub, vb, coef, hpar, k, maxiter, tol, errs, state, 0, T, band);
    else
      ls_sweep_simple_kernel<<<gs, bs, 0, s>>>(ub, vb, ua, va, coef, hpar, k, maxiter, tol, errs, state, 0, T, band);
    lc.n += 1;
    if (hook) hook(k, 1, (li & 1) ? 0 : 1);
  }
  ls_finalize_kernel<<<(batch + 127) / 128, 128, 0, s>>>(errs, state, batch, maxiter, tol, npix, T, nfull);
  lc.n += 1;
  const dim3 gr(gs.x, gs.y < 8 ? gs.y : 8, gs.z);   // replay launches: their blocks loop over the rows (see the kernel)
  for (int j = 0; j < T - 1 && nfull > 0; ++j) {   // conditional replay of an overshot fused block
    ls_sweep_simple_kernel<<<gr, bs, 0, s>>>(ua, va, ub, vb, coef, hpar, j, maxiter, tol, errs, state, 1, T, band);
    lc.n += 1;
    if (hook) hook(-1, 0, 2);    // replay step: both buffers may have been written; refresh the ghost rows of both
  }
  ls_select_kernel<<<gs, bs, 0, s>>>(ua, va, ub, vb, uo, vo, state, errs, maxiter, npix, err_out, err_stride,
                                     iters_out);
  lc.n += 1;
}

}  // namespace ofri
