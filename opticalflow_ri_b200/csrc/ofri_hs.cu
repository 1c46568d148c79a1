// ofri_hs.cu -- Horn-Schunck kernels (sm_100a): derivative stencils, the Jacobi sweep (simple per-pixel kernel
// and the temporally blocked shared-memory kernel), and the error norm.
//
// Reference: HornSchunck.py:52-127.  Per sweep and pixel the reference computes
//   uAvg = K (*) U, vAvg = K (*) V (3x3 weighted average, 'mirror' boundary), der = (fx uAvg + fy vAvg + ft) /
//   (alpha^2 + fx^2 + fy^2), U = uAvg - fx der, V = vAvg - fy der, for exactly Niter sweeps.
//
// Fused kernel (hs_fused_kernel<T,...>, see the comment above it): one launch advances T sweeps on shared-memory
// tiles of U, V with the per-pixel coefficients held in registers.  Algorithmic HBM traffic: 28 B per pixel per
// launch (read U, V, fx, fy, ft; write U, V), i.e. 28/T B per pixel-sweep.
#include "ofri_internal.h"
#include "ofri_pixel.cuh"
#include "ofri_hs_common.cuh"

namespace ofri {

// ---------------------------------------------------------------------------------------------------------------
// derivatives
// ---------------------------------------------------------------------------------------------------------------
__global__ void hs_derivs_kernel(Img im1, Img im2, Img fx, Img fy, Img ft) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= im1.W || y >= im1.H) return;
  const float* A = im1.p + (long)b * im1.stride;
  const float* B = im2.p + (long)b * im2.stride;
  int x1 = mirror1(x + 1, im1.W), y1 = mirror1(y + 1, im1.H);
  float a00 = A[(long)y * im1.pitch + x], a01 = A[(long)y * im1.pitch + x1];
  float a10 = A[(long)y1 * im1.pitch + x], a11 = A[(long)y1 * im1.pitch + x1];
  float b00 = B[(long)y * im2.pitch + x], b01 = B[(long)y * im2.pitch + x1];
  float b10 = B[(long)y1 * im2.pitch + x], b11 = B[(long)y1 * im2.pitch + x1];
  float dx, dy, dt;
  hs_deriv_point(a00, a01, a10, a11, b00, b01, b10, b11, &dx, &dy, &dt);
  fx.p[(long)b * fx.stride + (long)y * fx.pitch + x] = dx;
  fy.p[(long)b * fy.stride + (long)y * fy.pitch + x] = dy;
  ft.p[(long)b * ft.stride + (long)y * ft.pitch + x] = dt;
}
void launch_hs_derivs(const Img& im1, const Img& im2, const Img& fx, const Img& fy, const Img& ft, cudaStream_t s,
                      LaunchCounter& lc) {
  dim3 b(32, 8), g((im1.W + 31) / 32, (im1.H + 7) / 8, im1.batch);
  hs_derivs_kernel<<<g, b, 0, s>>>(im1, im2, fx, fy, ft);
  lc.n += 1;
}

// ---------------------------------------------------------------------------------------------------------------
// coefficient preparation for the fast path: (fx, fy, ft) -> (a, b, c) in place, once per level
// ---------------------------------------------------------------------------------------------------------------
__global__ void hs_prepare_kernel(Img fx, Img fy, Img ft, float alpha2) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= fx.W || y >= fx.H) return;
  long o = (long)b * fx.stride + (long)y * fx.pitch + x;
  float a, c, d;
  hs_normalise(fx.p[o], fy.p[o], ft.p[o], alpha2, &a, &c, &d);
  fx.p[o] = a;
  fy.p[o] = c;
  ft.p[o] = d;
}

// ---------------------------------------------------------------------------------------------------------------
// simple Jacobi sweep: one thread per pixel, one sweep per launch (cross-check for the fused kernel; also the
// fallback for images narrower than 2 pixels).  Fast path: planes hold the prepared (a, b, c); precise: raw.
// ---------------------------------------------------------------------------------------------------------------
template <bool PRECISE>
__global__ void hs_sweep_simple_kernel(Img ui, Img vi, Img uo, Img vo, Img fx, Img fy, Img ft, float alpha2) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  const int W = ui.W, H = ui.H;
  if (x >= W || y >= H) return;
  const float* U = ui.p + (long)b * ui.stride;
  const float* V = vi.p + (long)b * vi.stride;
  int xl = mirror1(x - 1, W), xr = mirror1(x + 1, W), yu = mirror1(y - 1, H), yd = mirror1(y + 1, H);
  long ru = (long)yu * ui.pitch, rm = (long)y * ui.pitch, rd = (long)yd * ui.pitch;
  long su = (long)yu * vi.pitch, sm = (long)y * vi.pitch, sd = (long)yd * vi.pitch;
  float dx = fx.p[(long)b * fx.stride + (long)y * fx.pitch + x];
  float dy = fy.p[(long)b * fy.stride + (long)y * fy.pitch + x];
  float dt = ft.p[(long)b * ft.stride + (long)y * ft.pitch + x];
  float un, vn;
  if (!PRECISE) {
    float ua = hs_avg_cols(fadd(U[ru + xl], U[rd + xl]), fadd(U[ru + x], U[rd + x]), fadd(U[ru + xr], U[rd + xr]),
                           U[rm + xl], U[rm + xr]);
    float va = hs_avg_cols(fadd(V[su + xl], V[sd + xl]), fadd(V[su + x], V[sd + x]), fadd(V[su + xr], V[sd + xr]),
                           V[sm + xl], V[sm + xr]);
    hs_update_n(ua, va, dx, dy, dt, &un, &vn);
  } else {
    float ua = hs_avg_cols_precise(dadd((double)U[ru + xl], (double)U[rd + xl]), dadd((double)U[ru + x], (double)U[rd + x]),
                                   dadd((double)U[ru + xr], (double)U[rd + xr]), (double)U[rm + xl], (double)U[rm + xr]);
    float va = hs_avg_cols_precise(dadd((double)V[su + xl], (double)V[sd + xl]), dadd((double)V[su + x], (double)V[sd + x]),
                                   dadd((double)V[su + xr], (double)V[sd + xr]), (double)V[sm + xl], (double)V[sm + xr]);
    float den = hs_den(dx, dy, alpha2);
    hs_update_precise(ua, va, dx, dy, dt, den, rcp_rn(den), &un, &vn);
  }
  uo.p[(long)b * uo.stride + (long)y * uo.pitch + x] = un;
  vo.p[(long)b * vo.stride + (long)y * vo.pitch + x] = vn;
}

// ---------------------------------------------------------------------------------------------------------------
// fused (temporally blocked) Jacobi kernel, register-resident coefficients
// ---------------------------------------------------------------------------------------------------------------
// Geometry: a CTA owns a shared tile of SH x SW cells, SW = 4 NG, SH = R NRG + 2.  Thread (cg, rg) owns the 4 x R
// cell strip at columns 4cg..4cg+3, rows 1 + rg R .. rg R + R -- the SAME cells in every sweep, so its coefficients
// live in registers for the whole launch (loaded straight from HBM with LDG.128, never staged).  Only U and V go
// through shared memory (two ping-pong buffers each, staged with 16-byte cp.async).  Per sweep a thread slides a
// 3-row register window down its strip: one LDS.128 per plane per row; the two halo columns come from the
// neighbouring lanes by warp shuffle.  With NG a multiple or divisor of 32 a warp never straddles a strip row, so
// the lanes whose neighbour is in another warp are exactly the tile-border columns, whose cells are stale anyway:
// no shared-memory fallback is needed.  After sweep s the cells at distance <= s from the tile border are stale;
// the last sweep stores the TH x TW interior (distance >= T rows / HX columns) straight to HBM with float4 stores.
// Tiles that touch the image border run the EDGE instantiation, which re-applies scipy's 'mirror' rule at every
// sweep; interior tiles carry no boundary code at all.
// PRECISE instantiation: stencil accumulated in float64 and rounded once, update with separately rounded f32
// operations and a correctly rounded division -- the reference's arithmetic (scipy correlate + numba), bit for bit.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;   // src-size 0 -> 16 bytes of zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

template <int T, int R, int NRG, int NG>
struct HsCfg {
  static constexpr int HX = (T <= 4) ? 4 : 8;
  static constexpr int SW = 4 * NG;
  static constexpr int SH = R * NRG + 2;
  static constexpr int NACT = NG * NRG;                 // threads that own a strip
  static constexpr int NT = ((NACT + 31) / 32) * 32;    // whole warps (the halo exchange shuffles)
  static constexpr bool ALIGNED = (NG % 32 == 0) || (32 % NG == 0);
  static constexpr int TW = SW - 2 * HX;
  static constexpr int TH = SH - 2 * T;
  static constexpr int PLANE = SH * SW;
  static constexpr int SMEM_BYTES = 4 * PLANE * 4;      // U[2], V[2]
  static_assert(TW > 0 && TH > 0 && HX >= T, "bad tile");
};


// one shared row of this thread's strip: columns sx-1 .. sx+4 of U and V
template <int SW, bool EDGE, bool ALIGNED>
__device__ __forceinline__ void hs_row6(const float* __restrict__ pu, const float* __restrict__ pv, int offl, int offr,
                                        bool lane_lo, bool lane_hi, const HsEdge& eg, float (&du)[6], float (&dv)[6]) {
  float4 qu = *reinterpret_cast<const float4*>(pu);
  float4 qv = *reinterpret_cast<const float4*>(pv);
  du[1] = qu.x; du[2] = qu.y; du[3] = qu.z; du[4] = qu.w;
  dv[1] = qv.x; dv[2] = qv.y; dv[3] = qv.z; dv[4] = qv.w;
  du[0] = __shfl_up_sync(0xffffffffu, qu.w, 1);
  du[5] = __shfl_down_sync(0xffffffffu, qu.x, 1);
  dv[0] = __shfl_up_sync(0xffffffffu, qv.w, 1);
  dv[5] = __shfl_down_sync(0xffffffffu, qv.x, 1);
  if (!ALIGNED) {   // lanes 0 / 31 may have their neighbour in another warp: fetch it from shared memory
    float ul = pu[offl], ur = pu[offr], vl = pv[offl], vr = pv[offr];
    du[0] = lane_lo ? ul : du[0];
    du[5] = lane_hi ? ur : du[5];
    dv[0] = lane_lo ? vl : dv[0];
    dv[5] = lane_hi ? vr : dv[5];
  }
  if (EDGE) {   // 'mirror' in x: left neighbour of column 0 is column 1, right neighbour of column W-1 is column W-2
    if (eg.left_edge) { du[0] = du[2]; dv[0] = dv[2]; }
    if (eg.right_j == 0) { du[2] = du[0]; dv[2] = dv[0]; }
    if (eg.right_j == 1) { du[3] = du[1]; dv[3] = dv[1]; }
    if (eg.right_j == 2) { du[4] = du[2]; dv[4] = dv[2]; }
    if (eg.right_j == 3) { du[5] = du[3]; dv[5] = dv[3]; }
  }
}

// one sweep over this thread's strip.  LAST: results go to HBM (interior cells only) instead of the next buffer.
template <int T, int R, int NRG, int NG, bool EDGE, bool PRECISE, bool LAST>
__device__ __forceinline__ void hs_sweep(const float* __restrict__ cu, const float* __restrict__ cv,
                                         float* __restrict__ nu, float* __restrict__ nv, int r0, int sx, int offl,
                                         int offr, bool lane_lo, bool lane_hi, const HsEdge& eg, bool active,
                                         const HsCoef<PRECISE> (&k)[R], float* __restrict__ gU, float* __restrict__ gV,
                                         long gpitch, int gy0, int gx, int H, int W) {
  using C = HsCfg<T, R, NRG, NG>;
  constexpr int SW = C::SW;
  float wu[3][6], wv[3][6];
  const float* pu = cu + (r0 - 1) * SW + sx;
  const float* pv = cv + (r0 - 1) * SW + sx;
  hs_row6<SW, EDGE, C::ALIGNED>(pu, pv, offl, offr, lane_lo, lane_hi, eg, wu[0], wv[0]);
  hs_row6<SW, EDGE, C::ALIGNED>(pu + SW, pv + SW, offl, offr, lane_lo, lane_hi, eg, wu[1], wv[1]);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int A = j % 3, B = (j + 1) % 3, Cc = (j + 2) % 3;     // window roles rotate statically
    hs_row6<SW, EDGE, C::ALIGNED>(pu + (j + 2) * SW, pv + (j + 2) * SW, offl, offr, lane_lo, lane_hi, eg, wu[Cc],
                                  wv[Cc]);
    float ou[4], ov[4];
    if (EDGE && j == eg.top_j)        // global row 0: the row above is row 1 (= down)
      hs_row_update<PRECISE>(wu[Cc], wu[B], wu[Cc], wv[Cc], wv[B], wv[Cc], k[j], ou, ov);
    else if (EDGE && j == eg.bot_j)   // global row H-1: the row below is row H-2 (= up)
      hs_row_update<PRECISE>(wu[A], wu[B], wu[A], wv[A], wv[B], wv[A], k[j], ou, ov);
    else
      hs_row_update<PRECISE>(wu[A], wu[B], wu[Cc], wv[A], wv[B], wv[Cc], k[j], ou, ov);
    if (!LAST) {
      if (active) {
        const int so = (r0 + j) * SW + sx;
        *reinterpret_cast<float4*>(nu + so) = make_float4(ou[0], ou[1], ou[2], ou[3]);
        *reinterpret_cast<float4*>(nv + so) = make_float4(ov[0], ov[1], ov[2], ov[3]);
      }
    } else {
      const int sy = r0 + j, gy = gy0 + j;
      const bool in_rows = (sy >= T) && (sy < C::SH - T) && (gy < H);
      const bool in_cols = (sx >= C::HX) && (sx < SW - C::HX) && (gx < W);
      if (active && in_rows && in_cols) {   // the last group of a row may spill into the pitch padding
        const long go = (long)gy * gpitch + gx;
        *reinterpret_cast<float4*>(gU + go) = make_float4(ou[0], ou[1], ou[2], ou[3]);
        *reinterpret_cast<float4*>(gV + go) = make_float4(ov[0], ov[1], ov[2], ov[3]);
      }
    }
  }
}

template <int T, int R, int NRG, int NG, bool EDGE, bool PRECISE>
__device__ __forceinline__ void hs_fused_body(const Img& ui, const Img& vi, const Img& uo, const Img& vo,
                                              const Img& fx, const Img& fy, const Img& ft, float alpha2, float* smem) {
  using C = HsCfg<T, R, NRG, NG>;
  constexpr int SW = C::SW, SH = C::SH, HX = C::HX;
  const int b = blockIdx.z;
  const int W = ui.W, H = ui.H;
  const int x0 = blockIdx.x * C::TW - HX;      // global x of shared column 0 (multiple of 4)
  const int y0 = blockIdx.y * C::TH - T;       // global y of shared row 0
  const int tid = threadIdx.x;
  const int lane = tid & 31;

  // ---- stage U, V (cp.async, zero fill outside the allocation) -------------------------------------------------
  {
    const float* gU = ui.p + (long)b * ui.stride;
    const float* gV = vi.p + (long)b * vi.stride;
    for (int i = tid; i < SH * NG; i += C::NT) {
      int sy = i / NG, sg = i - sy * NG;
      int gy = y0 + sy, gx = x0 + 4 * sg;
      bool ok = (gy >= 0) && (gy < H) && (gx >= 0) && (gx < (int)ui.pitch);
      long go = ok ? (long)gy * ui.pitch + gx : 0;
      int so = sy * SW + 4 * sg;
      cp_async16(smem + so, gU + go, ok);
      cp_async16(smem + 2 * C::PLANE + so, gV + go, ok);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  // ---- this thread's strip and its coefficients (HBM -> registers) ---------------------------------------------
  const bool active = (C::NT == C::NACT) || (tid < C::NACT);
  const int t2 = active ? tid : C::NACT - 1;     // surplus threads shadow the last strip (they never store)
  const int cg = t2 % NG, rg = t2 / NG;
  const int sx = 4 * cg;
  const int r0 = 1 + rg * R;
  const int gx = x0 + sx;
  HsCoef<PRECISE> k[R];
  {
    const float* g0 = fx.p + (long)b * fx.stride;
    const float* g1 = fy.p + (long)b * fy.stride;
    const float* g2 = ft.p + (long)b * ft.stride;
    const bool okx = (gx >= 0) && (gx < (int)fx.pitch);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      int gy = y0 + r0 + j;
      bool ok = okx && (gy >= 0) && (gy < H);
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a, d = a;
      if (!EDGE || ok) {      // interior tiles lie entirely inside the image
        long go = (long)gy * fx.pitch + gx;
        a = ldg_f4(g0 + go);
        c = ldg_f4(g1 + go);
        d = ldg_f4(g2 + go);
      }
      k[j].c0[0] = a.x; k[j].c0[1] = a.y; k[j].c0[2] = a.z; k[j].c0[3] = a.w;
      k[j].c1[0] = c.x; k[j].c1[1] = c.y; k[j].c1[2] = c.z; k[j].c1[3] = c.w;
      k[j].c2[0] = d.x; k[j].c2[1] = d.y; k[j].c2[2] = d.z; k[j].c2[3] = d.w;
    }
    if constexpr (PRECISE) {
#pragma unroll
      for (int j = 0; j < R; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          k[j].c3[q] = hs_den(k[j].c0[q], k[j].c1[q], alpha2);
          k[j].c4[q] = rcp_rn(k[j].c3[q]);
        }
    }
  }
  const int offl = sx > 0 ? -1 : 0;                 // clamped halo offsets (only used when !ALIGNED)
  const int offr = sx + 4 < SW ? 4 : 3;
  const bool lane_lo = (lane == 0), lane_hi = (lane == 31);
  HsEdge eg;
  eg.left_edge = EDGE && (gx == 0);
  eg.right_j = EDGE ? (W - 1) - gx : -1;
  eg.top_j = EDGE ? -(y0 + r0) : -1000;
  eg.bot_j = EDGE ? (H - 1) - (y0 + r0) : -1000;
  float* gU = uo.p + (long)b * uo.stride;
  float* gV = vo.p + (long)b * vo.stride;

  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();

  // ---- T sweeps: T-1 through shared memory, the last one straight to HBM ----------------------------------------
#pragma unroll 1
  for (int s = 0; s < T - 1; ++s) {
    const float* cu = smem + (s & 1) * C::PLANE;
    const float* cv = smem + (2 + (s & 1)) * C::PLANE;
    float* nu = smem + ((s + 1) & 1) * C::PLANE;
    float* nv = smem + (2 + ((s + 1) & 1)) * C::PLANE;
    hs_sweep<T, R, NRG, NG, EDGE, PRECISE, false>(cu, cv, nu, nv, r0, sx, offl, offr, lane_lo, lane_hi, eg, active, k,
                                                  gU, gV, uo.pitch, y0 + r0, gx, H, W);
    __syncthreads();
  }
  {
    constexpr int s = T - 1;
    const float* cu = smem + (s & 1) * C::PLANE;
    const float* cv = smem + (2 + (s & 1)) * C::PLANE;
    hs_sweep<T, R, NRG, NG, EDGE, PRECISE, true>(cu, cv, nullptr, nullptr, r0, sx, offl, offr, lane_lo, lane_hi, eg,
                                                 active, k, gU, gV, uo.pitch, y0 + r0, gx, H, W);
  }
}

template <int T, int R, int NRG, int NG, bool PRECISE, int MINB>
__global__ void __launch_bounds__(HsCfg<T, R, NRG, NG>::NT, MINB)
hs_fused_kernel(Img ui, Img vi, Img uo, Img vo, Img fx, Img fy, Img ft, float alpha2) {
  using C = HsCfg<T, R, NRG, NG>;
  extern __shared__ __align__(16) float smem[];
  const int x0 = blockIdx.x * C::TW - C::HX, y0 = blockIdx.y * C::TH - T;
  const bool edge = (x0 < 0) || (x0 + C::SW > ui.W) || (y0 < 0) || (y0 + C::SH > ui.H);   // CTA-uniform
  if (edge)
    hs_fused_body<T, R, NRG, NG, true, PRECISE>(ui, vi, uo, vo, fx, fy, ft, alpha2, smem);
  else
    hs_fused_body<T, R, NRG, NG, false, PRECISE>(ui, vi, uo, vo, fx, fy, ft, alpha2, smem);
}

template <int T, int R, int NRG, int NG, bool PRECISE, int MINB>
static void launch_hs_fused_cfg(const Img& ui, const Img& vi, const Img& uo, const Img& vo, const Img& fx,
                                const Img& fy, const Img& ft, float alpha2, cudaStream_t s) {
  using C = HsCfg<T, R, NRG, NG>;
  auto kern = hs_fused_kernel<T, R, NRG, NG, PRECISE, MINB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  dim3 g((ui.W + C::TW - 1) / C::TW, (ui.H + C::TH - 1) / C::TH, ui.batch);
  kern<<<g, C::NT, C::SMEM_BYTES, s>>>(ui, vi, uo, vo, fx, fy, ft, alpha2);
}

// one tile shape: 34 x 128 cells (R = 4 rows per thread, 8 row groups, 32 float4 column groups)
template <int T, bool PRECISE>
static void launch_hs_fused_T(const Img& ui, const Img& vi, const Img& uo, const Img& vo, const Img& fx, const Img& fy,
                              const Img& ft, float alpha2, cudaStream_t s) {
  launch_hs_fused_cfg<T, 4, 8, 32, PRECISE, PRECISE ? 1 : 2>(ui, vi, uo, vo, fx, fy, ft, alpha2, s);
}

// returns true if the launch honoured `sub` (only the TMA kernel can run a subset of the tile rows)
static bool launch_hs_fused(int T, int variant, bool precise, const Img& ui, const Img& vi, const Img& uo,
                            const Img& vo, const Img& fx, const Img& fy, const Img& ft, float alpha2, cudaStream_t s,
                            const HsTileRows* sub = nullptr) {
  // hs_variant >= 24: the persistent TMA-fed register-resident kernel (ofri_hs_tma.cu); it covers T in {4, 6, 8}
  // (reference arithmetic: also 2, 3) on images larger than its ghost frame.  Everything else -- other T (the tail of a
  // sweep count that is not a multiple of the fuse factor), tiny images, hs_variant = 0 -- runs the shared-memory kernel.
  if (variant >= 24 && launch_hs_tma(T, variant, precise, ui, vi, uo, vo, fx, fy, ft, alpha2, s, sub)) return true;
  if (sub && !sub->inside) return false;      // the shared-memory kernel always runs whole launches: do it in the second part
#define OFRI_HS_T(TT)                                                                     \
  case TT:                                                                                \
    if (precise) launch_hs_fused_T<TT, true>(ui, vi, uo, vo, fx, fy, ft, alpha2, s);      \
    else launch_hs_fused_T<TT, false>(ui, vi, uo, vo, fx, fy, ft, alpha2, s);             \
    break;
  switch (T) {
    OFRI_HS_T(1)
    OFRI_HS_T(2)
    OFRI_HS_T(3)
    OFRI_HS_T(4)
    OFRI_HS_T(5)
    OFRI_HS_T(6)
    default:
      if (precise) launch_hs_fused_T<8, true>(ui, vi, uo, vo, fx, fy, ft, alpha2, s);
      else launch_hs_fused_T<8, false>(ui, vi, uo, vo, fx, fy, ft, alpha2, s);
      break;
  }
#undef OFRI_HS_T
  return true;
}

int launch_hs_iterate(const Img& ua, const Img& va, const Img& ub, const Img& vb, const Img& fx, const Img& fy,
                      const Img& ft, float alpha, int niter, int fuse, int variant, bool precise, cudaStream_t s,
                      LaunchCounter& lc, const HsHook& hook, const HsSplit* split) {
  const float alpha2 = alpha * alpha;   // f32 alpha**2 as in the numba signature (HornSchunck.py:52-55)
  int cur = 0;
  const Img* U[2] = {&ua, &ub};
  const Img* V[2] = {&va, &vb};
  if (fuse == 7) fuse = 6;
  if (fuse > 8) fuse = 8;
  const bool can_fuse = fuse >= 1 && ua.W >= 2 && ua.H >= 2 && (ua.pitch % 4 == 0) && ua.pitch == va.pitch &&
                        ua.pitch == ub.pitch && ua.pitch == vb.pitch && ua.pitch == fx.pitch &&
                        ua.pitch == fy.pitch && ua.pitch == ft.pitch && ((uintptr_t)ua.p % 16 == 0) &&
                        ((uintptr_t)va.p % 16 == 0) && ((uintptr_t)ub.p % 16 == 0) && ((uintptr_t)vb.p % 16 == 0) &&
                        ((uintptr_t)fx.p % 16 == 0) && ((uintptr_t)fy.p % 16 == 0) && ((uintptr_t)ft.p % 16 == 0) &&
                        (ua.stride % 4 == 0) && (fx.stride % 4 == 0);
  if (niter > 0 && !precise) {   // fast path: (fx, fy, ft) -> normalised (a, b, c), in place
    dim3 b(32, 8), g((fx.W + 31) / 32, (fx.H + 7) / 8, fx.batch);
    hs_prepare_kernel<<<g, b, 0, s>>>(fx, fy, ft, alpha2);
    lc.n += 1;
  }
  int done = 0;
  while (done < niter) {
    int left = niter - done;
    if (fuse <= 0 || !can_fuse) {
      dim3 b(32, 8), g((ua.W + 31) / 32, (ua.H + 7) / 8, ua.batch);
      if (precise)
        hs_sweep_simple_kernel<true><<<g, b, 0, s>>>(*U[cur], *V[cur], *U[cur ^ 1], *V[cur ^ 1], fx, fy, ft, alpha2);
      else
        hs_sweep_simple_kernel<false><<<g, b, 0, s>>>(*U[cur], *V[cur], *U[cur ^ 1], *V[cur ^ 1], fx, fy, ft, alpha2);
      if (split && split->every > 0 && ((done + 1) % split->every == 0 || done + 1 == niter)) {
        split->begin(cur ^ 1);
        split->end();
      }
      done += 1;
    } else {
      int T = left < fuse ? left : fuse;
      if (T == 7) T = 6;
      const bool block_end = split && split->every > 0 && ((done + T) % split->every == 0 || done + T == niter);
      if (block_end) {
        // boundary tiles first, then the exchange of the freshly written buffer starts (other stream) while the
        // interior tiles run; a kernel that cannot split runs whole and the exchange follows it
        const HsTileRows outer{split->mid_lo, split->mid_hi, false, 0}, inner{split->mid_lo, split->mid_hi, true, split->reserve_sms};
        if (launch_hs_fused(T, variant, precise, *U[cur], *V[cur], *U[cur ^ 1], *V[cur ^ 1], fx, fy, ft, alpha2, s, &outer)) {
          split->begin(cur ^ 1);
          launch_hs_fused(T, variant, precise, *U[cur], *V[cur], *U[cur ^ 1], *V[cur ^ 1], fx, fy, ft, alpha2, s, &inner);
          lc.n += 1;
        } else {
          launch_hs_fused(T, variant, precise, *U[cur], *V[cur], *U[cur ^ 1], *V[cur ^ 1], fx, fy, ft, alpha2, s);
          split->begin(cur ^ 1);
        }
        split->end();
      } else {
        launch_hs_fused(T, variant, precise, *U[cur], *V[cur], *U[cur ^ 1], *V[cur ^ 1], fx, fy, ft, alpha2, s);
      }
      done += T;
    }
    cur ^= 1;
    lc.n += 1;
    if (hook) hook(done, cur);
  }
  return cur;
}

// ---------------------------------------------------------------------------------------------------------------
// error norm (HornSchunck.py:100)
// ---------------------------------------------------------------------------------------------------------------
__global__ void sqdiff_reduce_kernel(Img u, Img v, Img u0, Img v0, int has0, double* acc, int y_lo, int y_hi) {
  const int b = blockIdx.z;
  double su = 0.0, sv = 0.0;
  for (int y = y_lo + blockIdx.y; y < y_hi; y += gridDim.y) {
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < u.W; x += gridDim.x * blockDim.x) {
      float a = u.p[(long)b * u.stride + (long)y * u.pitch + x];
      float c = v.p[(long)b * v.stride + (long)y * v.pitch + x];
      if (has0) {
        a = fsub(a, u0.p[(long)b * u0.stride + (long)y * u0.pitch + x]);
        c = fsub(c, v0.p[(long)b * v0.stride + (long)y * v0.pitch + x]);
      }
      su += (double)a * (double)a;
      sv += (double)c * (double)c;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    su += __shfl_xor_sync(0xffffffffu, su, o);
    sv += __shfl_xor_sync(0xffffffffu, sv, o);
  }
  __shared__ double shu[32], shv[32];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { shu[w] = su; shv[w] = sv; }
  __syncthreads();
  if (w == 0) {
    int nw = blockDim.x >> 5;
    su = l < nw ? shu[l] : 0.0;
    sv = l < nw ? shv[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) {
      su += __shfl_xor_sync(0xffffffffu, su, o);
      sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    if (l == 0) {
      atomicAdd(acc + 2 * b, su);
      atomicAdd(acc + 2 * b + 1, sv);
    }
  }
}
__global__ void hs_error_finish_kernel(const double* acc, float* err, int err_stride, int batch, double npix) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  err[(long)b * err_stride] = (float)((sqrt(acc[2 * b]) + sqrt(acc[2 * b + 1])) / npix);
}
void launch_hs_error_sums(const Img& u, const Img& v, const Img& u0, const Img& v0, double* acc, int y_lo, int y_hi,
                          cudaStream_t s, LaunchCounter& lc) {
  cudaMemsetAsync(acc, 0, sizeof(double) * 2 * u.batch, s);
  int rows = y_hi - y_lo;
  int gy = rows < 64 ? (rows > 0 ? rows : 1) : 64;
  dim3 b(256), g((u.W + 255) / 256 > 4 ? 4 : (u.W + 255) / 256, gy, u.batch);
  sqdiff_reduce_kernel<<<g, b, 0, s>>>(u, v, u0, v0, u0.p != nullptr ? 1 : 0, acc, y_lo, y_hi);
  lc.n += 1;
}
void launch_hs_error_finish(const double* acc, float* err, int err_stride, int batch, double npix, cudaStream_t s,
                            LaunchCounter& lc) {
  hs_error_finish_kernel<<<(batch + 127) / 128, 128, 0, s>>>(acc, err, err_stride, batch, npix);
  lc.n += 1;
}
void launch_hs_error(const Img& u, const Img& v, const Img& u0, const Img& v0, double* acc, float* err, int err_stride,
                     cudaStream_t s, LaunchCounter& lc) {
  launch_hs_error_sums(u, v, u0, v0, acc, 0, u.H, s, lc);
  launch_hs_error_finish(acc, err, err_stride, u.batch, (double)u.H * (double)u.W, s, lc);
}

}  // namespace ofri
