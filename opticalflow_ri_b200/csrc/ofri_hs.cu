// ofri_hs.cu -- Horn-Schunck kernels (sm_100a): derivative stencils, the Jacobi sweep (simple per-pixel kernel
// and the temporally blocked shared-memory kernel), and the error norm.
//
// Reference: HornSchunck.py:52-127.  Per sweep and pixel the reference computes
//   uAvg = K (*) U, vAvg = K (*) V (3x3 weighted average, 'mirror' boundary), der = (fx uAvg + fy vAvg + ft) /
//   (alpha^2 + fx^2 + fy^2), U = uAvg - fx der, V = vAvg - fy der, for exactly Niter sweeps.
//
// Fused kernel (hs_fused_kernel<T,...>): one launch advances T sweeps.  A CTA stages an SH x SW tile of U, V, fx,
// fy, ft (halo T rows / HX >= T columns, HX a multiple of 4 so every row segment is 16-byte aligned) into shared
// memory with 16-byte cp.async (zero-fill outside the image), computes 1/(alpha^2+fx^2+fy^2) once, then runs T
// sweeps ping-ponging between two shared U/V buffers; after sweep s only cells at distance > s from the tile
// border are valid, and the cells at distance >= T are written back with float4 stores.  The 'mirror' rule is
// re-applied at every sweep for cells on the image border.  Algorithmic HBM traffic: 28 B per pixel per launch
// (read U, V, fx, fy, ft; write U, V), i.e. 28/T B per pixel-sweep.
#include "ofri_internal.h"
#include "ofri_pixel.cuh"

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
// simple Jacobi sweep: one thread per pixel, one sweep per launch (cross-check for the fused kernel; also the
// fallback for images narrower than 2 pixels)
// ---------------------------------------------------------------------------------------------------------------
__global__ void hs_sweep_simple_kernel(Img ui, Img vi, Img uo, Img vo, Img fx, Img fy, Img ft, float alpha2) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  const int W = ui.W, H = ui.H;
  if (x >= W || y >= H) return;
  const float* U = ui.p + (long)b * ui.stride;
  const float* V = vi.p + (long)b * vi.stride;
  int xl = mirror1(x - 1, W), xr = mirror1(x + 1, W), yu = mirror1(y - 1, H), yd = mirror1(y + 1, H);
  long ru = (long)yu * ui.pitch, rm = (long)y * ui.pitch, rd = (long)yd * ui.pitch;
  float ua = hs_avg_cols(fadd(U[ru + xl], U[rd + xl]), fadd(U[ru + x], U[rd + x]), fadd(U[ru + xr], U[rd + xr]),
                         U[rm + xl], U[rm + xr]);
  ru = (long)yu * vi.pitch, rm = (long)y * vi.pitch, rd = (long)yd * vi.pitch;
  float va = hs_avg_cols(fadd(V[ru + xl], V[rd + xl]), fadd(V[ru + x], V[rd + x]), fadd(V[ru + xr], V[rd + xr]),
                         V[rm + xl], V[rm + xr]);
  float dx = fx.p[(long)b * fx.stride + (long)y * fx.pitch + x];
  float dy = fy.p[(long)b * fy.stride + (long)y * fy.pitch + x];
  float dt = ft.p[(long)b * ft.stride + (long)y * ft.pitch + x];
  float inv = hs_inv_den(dx, dy, alpha2);
  float un, vn;
  hs_update(ua, va, dx, dy, dt, inv, &un, &vn);
  uo.p[(long)b * uo.stride + (long)y * uo.pitch + x] = un;
  vo.p[(long)b * vo.stride + (long)y * vo.pitch + x] = vn;
}

// ---------------------------------------------------------------------------------------------------------------
// fused (temporally blocked) Jacobi kernel
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;   // src-size 0 -> 16 bytes of zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

template <int T, int SW, int SH, int HX, int NRG>
struct HsFusedCfg {
  static constexpr int NG = SW / 4;           // float4 column groups per tile row
  static constexpr int NT = NG * NRG;         // threads per CTA
  static constexpr int TW = SW - 2 * HX;      // output tile
  static constexpr int TH = SH - 2 * T;
  static constexpr int PLANE = SH * SW;       // floats per shared plane
  static constexpr int SMEM_BYTES = 8 * PLANE * 4;   // U[2], V[2], fx, fy, ft, inv
  static_assert(SW % 4 == 0 && HX % 4 == 0 && HX >= T && TW > 0 && TH > 0, "bad tile");
};

// load one shared row (6 columns sx-1 .. sx+4 of U and V) and apply the 'mirror' rule in x
template <int SW>
__device__ __forceinline__ void hs_load_row(const float* __restrict__ cu, const float* __restrict__ cv, int r, int sx,
                                            int sxl, int sxr, bool left_edge, int right_j, float (&du)[6],
                                            float (&dv)[6]) {
  const float* pu = cu + r * SW;
  const float* pv = cv + r * SW;
  float4 q = *reinterpret_cast<const float4*>(pu + sx);
  du[0] = pu[sxl]; du[1] = q.x; du[2] = q.y; du[3] = q.z; du[4] = q.w; du[5] = pu[sxr];
  q = *reinterpret_cast<const float4*>(pv + sx);
  dv[0] = pv[sxl]; dv[1] = q.x; dv[2] = q.y; dv[3] = q.z; dv[4] = q.w; dv[5] = pv[sxr];
  // the left neighbour of global column 0 is column 1; the right neighbour of column W-1 is column W-2
  if (left_edge) { du[0] = du[2]; dv[0] = dv[2]; }
  if (right_j == 0) { du[2] = du[0]; dv[2] = dv[0]; }
  if (right_j == 1) { du[3] = du[1]; dv[3] = dv[1]; }
  if (right_j == 2) { du[4] = du[2]; dv[4] = dv[2]; }
  if (right_j == 3) { du[5] = du[3]; dv[5] = dv[3]; }
}
// one row of 4 pixels: window rows (up, mid, down), coefficients from shared memory, result to the next buffer
__device__ __forceinline__ void hs_row_compute(const float (&uu)[6], const float (&um)[6], const float (&ud)[6],
                                               const float (&vu)[6], const float (&vm)[6], const float (&vd)[6],
                                               const float* __restrict__ sFx, const float* __restrict__ sFy,
                                               const float* __restrict__ sFt, const float* __restrict__ sIn, int so,
                                               float* __restrict__ nu, float* __restrict__ nv) {
  float vsu[6], vsv[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    vsu[c] = fadd(uu[c], ud[c]);
    vsv[c] = fadd(vu[c], vd[c]);
  }
  float4 qfx = *reinterpret_cast<const float4*>(sFx + so);
  float4 qfy = *reinterpret_cast<const float4*>(sFy + so);
  float4 qft = *reinterpret_cast<const float4*>(sFt + so);
  float4 qin = *reinterpret_cast<const float4*>(sIn + so);
  const float afx[4] = {qfx.x, qfx.y, qfx.z, qfx.w};
  const float afy[4] = {qfy.x, qfy.y, qfy.z, qfy.w};
  const float aft[4] = {qft.x, qft.y, qft.z, qft.w};
  const float ain[4] = {qin.x, qin.y, qin.z, qin.w};
  float ou[4], ov[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float ua = hs_avg_cols(vsu[j], vsu[j + 1], vsu[j + 2], um[j], um[j + 2]);
    float va = hs_avg_cols(vsv[j], vsv[j + 1], vsv[j + 2], vm[j], vm[j + 2]);
    hs_update(ua, va, afx[j], afy[j], aft[j], ain[j], &ou[j], &ov[j]);
  }
  *reinterpret_cast<float4*>(nu + so) = make_float4(ou[0], ou[1], ou[2], ou[3]);
  *reinterpret_cast<float4*>(nv + so) = make_float4(ov[0], ov[1], ov[2], ov[3]);
}

template <int T, int SW, int SH, int HX, int NRG>
__global__ void __launch_bounds__(HsFusedCfg<T, SW, SH, HX, NRG>::NT)
hs_fused_kernel(Img ui, Img vi, Img uo, Img vo, Img fx, Img fy, Img ft, float alpha2) {
  using C = HsFusedCfg<T, SW, SH, HX, NRG>;
  extern __shared__ __align__(16) float smem[];
  // planes: U[0], U[1], V[0], V[1], fx, fy, ft, inv
  float* sFx = smem + 4 * C::PLANE;
  float* sFy = smem + 5 * C::PLANE;
  float* sFt = smem + 6 * C::PLANE;
  float* sIn = smem + 7 * C::PLANE;

  const int b = blockIdx.z;
  const int W = ui.W, H = ui.H;
  const int x0 = blockIdx.x * C::TW - HX;      // global x of shared column 0 (multiple of 4)
  const int y0 = blockIdx.y * C::TH - T;       // global y of shared row 0
  const int tid = threadIdx.x;

  // ---- stage the tile (16-byte cp.async, zero fill outside the allocation) --------------------------------------
  {
    const float* gU = ui.p + (long)b * ui.stride;
    const float* gV = vi.p + (long)b * vi.stride;
    const float* gFx = fx.p + (long)b * fx.stride;
    const float* gFy = fy.p + (long)b * fy.stride;
    const float* gFt = ft.p + (long)b * ft.stride;
    for (int i = tid; i < SH * C::NG; i += C::NT) {
      int sy = i / C::NG, sg = i - sy * C::NG;
      int gy = y0 + sy, gx = x0 + 4 * sg;
      bool ok = (gy >= 0) && (gy < H) && (gx >= 0) && (gx < (int)ui.pitch);
      int cy = ok ? gy : 0, cx = ok ? gx : 0;
      int so = sy * SW + 4 * sg;
      cp_async16(smem + so, gU + (long)cy * ui.pitch + cx, ok);
      cp_async16(smem + 2 * C::PLANE + so, gV + (long)cy * vi.pitch + cx, ok);
      cp_async16(sFx + so, gFx + (long)cy * fx.pitch + cx, ok);
      cp_async16(sFy + so, gFy + (long)cy * fy.pitch + cx, ok);
      cp_async16(sFt + so, gFt + (long)cy * ft.pitch + cx, ok);
    }
    cp_async_commit_wait_all();
    __syncthreads();
    for (int i = tid; i < SH * C::NG; i += C::NT) {
      float4 a = reinterpret_cast<const float4*>(sFx)[i];
      float4 c = reinterpret_cast<const float4*>(sFy)[i];
      float4 r;
      r.x = hs_inv_den(a.x, c.x, alpha2);
      r.y = hs_inv_den(a.y, c.y, alpha2);
      r.z = hs_inv_den(a.z, c.z, alpha2);
      r.w = hs_inv_den(a.w, c.w, alpha2);
      reinterpret_cast<float4*>(sIn)[i] = r;
    }
    __syncthreads();
  }

  // ---- T sweeps --------------------------------------------------------------------------------------------------
  const int cg = tid % C::NG, rg = tid / C::NG;
  const int sx = 4 * cg;                          // first shared column of this thread's 4-pixel group
  const int gx = x0 + sx;                         // its global x
  const int sxl = sx > 0 ? sx - 1 : 0;            // clamped halo columns (garbage only reaches invalid cells)
  const int sxr = sx + 4 < SW ? sx + 4 : SW - 1;
  const bool left_edge = (gx == 0);
  const int right_j = (W - 1) - gx;               // pixel j in [0,4) sitting on global column W-1 (else out of range)

#pragma unroll 1
  for (int s = 0; s < T; ++s) {
    const float* cu = smem + (s & 1) * C::PLANE;
    const float* cv = smem + (2 + (s & 1)) * C::PLANE;
    float* nu = smem + ((s + 1) & 1) * C::PLANE;
    float* nv = smem + (2 + ((s + 1) & 1)) * C::PLANE;
    // rows that can still become valid after this sweep: [s+1, SH-s-1), clipped to the image
    int lo = s + 1, hi = SH - s - 1;
    if (y0 + lo < 0) lo = -y0;
    if (y0 + hi > H) hi = H - y0;
    const int R = (hi - lo + NRG - 1) / NRG;
    const int r0 = lo + rg * R;
    const int r1 = (r0 + R < hi) ? r0 + R : hi;
    if (r0 < r1) {
      float wu[3][6], wv[3][6];   // sliding window of three shared rows x columns (sx-1 .. sx+4)
      // prime: row r0-1 (or its mirror, row -1 -> row 1) and row r0
      hs_load_row<SW>(cu, cv, (y0 + r0 == 0) ? r0 + 1 : r0 - 1, sx, sxl, sxr, left_edge, right_j, wu[0], wv[0]);
      hs_load_row<SW>(cu, cv, r0, sx, sxl, sxr, left_edge, right_j, wu[1], wv[1]);
      int r = r0;
      // rotate the window roles instead of moving registers (mirror: row H -> row H-2)
#define OFRI_HS_STEP(A, B, Cc)                                                                                       \
  hs_load_row<SW>(cu, cv, (y0 + r == H - 1) ? r - 1 : r + 1, sx, sxl, sxr, left_edge, right_j, wu[Cc], wv[Cc]);      \
  hs_row_compute(wu[A], wu[B], wu[Cc], wv[A], wv[B], wv[Cc], sFx, sFy, sFt, sIn, r * SW + sx, nu, nv);               \
  if (++r >= r1) break;
      while (true) {
        OFRI_HS_STEP(0, 1, 2)
        OFRI_HS_STEP(1, 2, 0)
        OFRI_HS_STEP(2, 0, 1)
      }
#undef OFRI_HS_STEP
    }
    __syncthreads();
  }

  // ---- write back the valid interior (distance >= T from the tile border) ----------------------------------------
  {
    const float* fu = smem + (T & 1) * C::PLANE;
    const float* fv = smem + (2 + (T & 1)) * C::PLANE;
    float* gU = uo.p + (long)b * uo.stride;
    float* gV = vo.p + (long)b * vo.stride;
    constexpr int OG = C::TW / 4;
    for (int i = tid; i < C::TH * OG; i += C::NT) {
      int ty = i / OG, tg = i - ty * OG;
      int sy = ty + T, sxx = HX + 4 * tg;
      int gy = y0 + sy, gxx = x0 + sxx;
      if (gy < H && gxx < W) {      // the last group of a row may spill into the pitch padding (garbage there is fine)
        float4 a = *reinterpret_cast<const float4*>(fu + sy * SW + sxx);
        float4 c = *reinterpret_cast<const float4*>(fv + sy * SW + sxx);
        *reinterpret_cast<float4*>(gU + (long)gy * uo.pitch + gxx) = a;
        *reinterpret_cast<float4*>(gV + (long)gy * vo.pitch + gxx) = c;
      }
    }
  }
}

template <int T, int SW, int SH, int HX, int NRG>
static void launch_hs_fused_cfg(const Img& ui, const Img& vi, const Img& uo, const Img& vo, const Img& fx,
                                const Img& fy, const Img& ft, float alpha2, cudaStream_t s) {
  using C = HsFusedCfg<T, SW, SH, HX, NRG>;
  auto kern = hs_fused_kernel<T, SW, SH, HX, NRG>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  dim3 g((ui.W + C::TW - 1) / C::TW, (ui.H + C::TH - 1) / C::TH, ui.batch);
  kern<<<g, C::NT, C::SMEM_BYTES, s>>>(ui, vi, uo, vo, fx, fy, ft, alpha2);
}

// variant table: (SW, SH, NRG) choices per T.  variant 0 = default.
template <int T>
static void launch_hs_fused_T(int variant, const Img& ui, const Img& vi, const Img& uo, const Img& vo, const Img& fx,
                              const Img& fy, const Img& ft, float alpha2, cudaStream_t s) {
  constexpr int HX = (T <= 4) ? 4 : 8;
  switch (variant) {
    default:
    case 0: launch_hs_fused_cfg<T, 64 + 2 * HX, 32 + 2 * T, HX, 8>(ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
    case 1: launch_hs_fused_cfg<T, 128 + 2 * HX, 16 + 2 * T, HX, 6>(ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
    case 2: launch_hs_fused_cfg<T, 128 + 2 * HX, 32 + 2 * T, HX, 8>(ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
    case 3: launch_hs_fused_cfg<T, 64 + 2 * HX, 16 + 2 * T, HX, 6>(ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
    case 4: launch_hs_fused_cfg<T, 32 + 2 * HX, 32 + 2 * T, HX, 8>(ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
    case 5: launch_hs_fused_cfg<T, 64 + 2 * HX, 32 + 2 * T, HX, 16>(ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
  }
}

static void launch_hs_fused(int T, int variant, const Img& ui, const Img& vi, const Img& uo, const Img& vo,
                            const Img& fx, const Img& fy, const Img& ft, float alpha2, cudaStream_t s) {
  switch (T) {
    case 1: launch_hs_fused_T<1>(variant, ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
    case 2: launch_hs_fused_T<2>(variant, ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
    case 3: launch_hs_fused_T<3>(variant, ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
    case 4: launch_hs_fused_T<4>(variant, ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
    case 5: launch_hs_fused_T<5>(variant, ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
    case 6: launch_hs_fused_T<6>(variant, ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
    default: launch_hs_fused_T<8>(variant, ui, vi, uo, vo, fx, fy, ft, alpha2, s); break;
  }
}

int launch_hs_iterate(const Img& ua, const Img& va, const Img& ub, const Img& vb, const Img& fx, const Img& fy,
                      const Img& ft, float alpha, int niter, int fuse, int variant, cudaStream_t s,
                      LaunchCounter& lc) {
  const float alpha2 = alpha * alpha;   // f32 alpha**2 as in the numba signature (HornSchunck.py:52-55)
  int cur = 0;
  const Img* U[2] = {&ua, &ub};
  const Img* V[2] = {&va, &vb};
  if (fuse == 7) fuse = 6;
  if (fuse > 8) fuse = 8;
  const bool can_fuse = fuse >= 1 && ua.W >= 2 && ua.H >= 2 && (ua.pitch % 4 == 0) && ua.pitch == va.pitch &&
                        ua.pitch == ub.pitch && ua.pitch == vb.pitch && ua.pitch == fx.pitch &&
                        ua.pitch == fy.pitch && ua.pitch == ft.pitch && ((uintptr_t)ua.p % 16 == 0) &&
                        ((uintptr_t)ub.p % 16 == 0) && ((uintptr_t)fx.p % 16 == 0) && (ua.stride % 4 == 0);
  int done = 0;
  while (done < niter) {
    int left = niter - done;
    if (fuse <= 0 || !can_fuse) {
      dim3 b(32, 8), g((ua.W + 31) / 32, (ua.H + 7) / 8, ua.batch);
      hs_sweep_simple_kernel<<<g, b, 0, s>>>(*U[cur], *V[cur], *U[cur ^ 1], *V[cur ^ 1], fx, fy, ft, alpha2);
      done += 1;
    } else {
      int T = left < fuse ? left : fuse;
      if (T == 7) T = 6;
      launch_hs_fused(T, variant, *U[cur], *V[cur], *U[cur ^ 1], *V[cur ^ 1], fx, fy, ft, alpha2, s);
      done += T;
    }
    cur ^= 1;
    lc.n += 1;
  }
  return cur;
}

// ---------------------------------------------------------------------------------------------------------------
// error norm (HornSchunck.py:100)
// ---------------------------------------------------------------------------------------------------------------
__global__ void sqdiff_reduce_kernel(Img u, Img v, Img u0, Img v0, int has0, double* acc) {
  const int b = blockIdx.z;
  double su = 0.0, sv = 0.0;
  for (int y = blockIdx.y; y < u.H; y += gridDim.y) {
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
void launch_hs_error(const Img& u, const Img& v, const Img& u0, const Img& v0, double* acc, float* err, int err_stride,
                     cudaStream_t s, LaunchCounter& lc) {
  cudaMemsetAsync(acc, 0, sizeof(double) * 2 * u.batch, s);
  int gy = u.H < 64 ? u.H : 64;
  dim3 b(256), g((u.W + 255) / 256 > 4 ? 4 : (u.W + 255) / 256, gy, u.batch);
  sqdiff_reduce_kernel<<<g, b, 0, s>>>(u, v, u0, v0, u0.p != nullptr ? 1 : 0, acc);
  hs_error_finish_kernel<<<(u.batch + 127) / 128, 128, 0, s>>>(acc, err, err_stride, u.batch,
                                                               (double)u.H * (double)u.W);
  lc.n += 2;
}

}  // namespace ofri
