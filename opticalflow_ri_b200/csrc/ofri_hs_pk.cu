// ofri_hs_pk.cu -- register-resident, packed-f32x2 Horn-Schunck Jacobi kernel (sm_100a only).
//
// Reference: HornSchunck.py:52-71 (HS_helper / HS_helper2), fast arithmetic (ofri_pixel.cuh hs_avg_cols + hs_update_n).
//
// Tile geometry and ownership are those of hs_fused_kernel (ofri_hs.cu): a CTA owns SH x 128 cells, SH = R NRG + 2;
// thread (lane, rg) owns the 4 x R strip at columns 4 lane .., rows 1 + rg R ..; a warp is exactly one row group.  What
// differs:
//   * the strip's state lives in REGISTERS for all T sweeps: U and V as packed pairs (one 64-bit register pair per
//     cell), the coefficients as a packed (a, b) pair plus c.  HBM is read once (LDG.128) and written once (STG.128).
//   * arithmetic on the pairs uses Blackwell's packed FP32 instructions (add/mul/fma.rn.f32x2 -> FADD2/FMUL2/FFMA2):
//     U and V go through identical stencil arithmetic, so one instruction does both components -- half the issue
//     slots for the same IEEE operations, i.e. bit-identical to the scalar kernels.
//   * the only data crossing warps is the row above / below each strip: every sweep a thread publishes its first and
//     last row in a double-buffered shared array (2 x STS.128 each) and reads its two neighbours' rows; halo columns
//     come from the neighbouring lanes by shuffle.  One __syncthreads per sweep; 2 x 2(NRG+2) KB of shared memory.
// After sweep s the cells at distance <= s from the tile border are stale; the T x HX frame is discarded at the end.
// Tiles touching the image border run the EDGE instantiation (scipy 'mirror' re-applied every sweep).
// Algorithmic HBM traffic: 28 B per pixel per launch (read U, V, a, b, c; write U, V).
#include "ofri_internal.h"
#include "ofri_pixel.cuh"

namespace ofri {

typedef unsigned long long f2;   // two packed floats: lo = U (or a), hi = V (or b)

__device__ __forceinline__ f2 pk(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo_of(f2 x) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x));
  return a;
}
__device__ __forceinline__ float hi_of(f2 x) {
  float a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x));
  return b;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

template <int T, int R, int NRG>
struct PkCfg {
  static constexpr int HX = (T <= 4) ? 4 : 8;
  static constexpr int SW = 128;
  static constexpr int SH = R * NRG + 2;
  static constexpr int NT = 32 * NRG;
  static constexpr int TW = SW - 2 * HX;
  static constexpr int TH = SH - 2 * T;
  static constexpr int XG = NRG + 2;                       // edge-row slots: g = rg + 1; g = 0 / NRG+1 = tile halos
  static constexpr int ROWF = 2 * SW;                      // floats per published row (128 packed cells)
  static constexpr int SMEM_BYTES = 2 * XG * 2 * ROWF * 4; // [buffer][g][top / bottom][row]
  static_assert(TW > 0 && TH > 0 && HX >= T && NT <= 1024, "bad tile");
};

struct PkEdge {
  bool left_edge;
  int right_j, top_j, bot_j;
};

// a published row: cells {0,1} of every lane first (16 B per lane, conflict-free LDS.128), then cells {2,3}
__device__ __forceinline__ void row_store(float* row, int lane, const f2 (&c)[4]) {
  *reinterpret_cast<float4*>(row + 4 * lane) = make_float4(lo_of(c[0]), hi_of(c[0]), lo_of(c[1]), hi_of(c[1]));
  *reinterpret_cast<float4*>(row + 128 + 4 * lane) = make_float4(lo_of(c[2]), hi_of(c[2]), lo_of(c[3]), hi_of(c[3]));
}
__device__ __forceinline__ void row_load(const float* row, int lane, f2 (&c)[4]) {
  float4 p = *reinterpret_cast<const float4*>(row + 4 * lane);
  float4 q = *reinterpret_cast<const float4*>(row + 128 + 4 * lane);
  c[0] = pk(p.x, p.y); c[1] = pk(p.z, p.w); c[2] = pk(q.x, q.y); c[3] = pk(q.z, q.w);
}
// 6-wide window row from the 4 own cells: halo columns from lane -/+ 1 (lanes 0 / 31 get their own value back: the
// tile-border columns are stale by construction); EDGE: scipy 'mirror' in x
template <bool EDGE>
__device__ __forceinline__ void row6(const f2 (&c)[4], const PkEdge& eg, f2 (&d)[6]) {
  d[1] = c[0]; d[2] = c[1]; d[3] = c[2]; d[4] = c[3];
  float lu = __shfl_up_sync(0xffffffffu, lo_of(c[3]), 1), lv = __shfl_up_sync(0xffffffffu, hi_of(c[3]), 1);
  float ru = __shfl_down_sync(0xffffffffu, lo_of(c[0]), 1), rv = __shfl_down_sync(0xffffffffu, hi_of(c[0]), 1);
  d[0] = pk(lu, lv);
  d[5] = pk(ru, rv);
  if (EDGE) {
    if (eg.left_edge) d[0] = d[2];
    if (eg.right_j == 0) d[2] = d[0];
    if (eg.right_j == 1) d[3] = d[1];
    if (eg.right_j == 2) d[4] = d[2];
    if (eg.right_j == 3) d[5] = d[3];
  }
}
// one strip row: same expression tree as hs_avg_cols + hs_update_n (ofri_pixel.cuh), both components per instruction
__device__ __forceinline__ void row_update(const f2 (&up)[6], const f2 (&mid)[6], const f2 (&dn)[6], const f2 (&ab)[4],
                                           const float (&cc)[4], f2 c6, f2 c12, f2 (&out)[4]) {
  f2 vs[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) vs[c] = add2(up[c], dn[c]);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    f2 edges = add2(vs[q + 1], add2(mid[q], mid[q + 2]));
    f2 corners = add2(vs[q], vs[q + 2]);
    f2 avg = fma2(corners, c12, mul2(edges, c6));
    // -g = -(a ua + b va + c): round-to-nearest is sign-symmetric, so this is exactly the negation of hs_update_n's g
    float ng = fmaf(-lo_of(ab[q]), lo_of(avg), fmaf(-hi_of(ab[q]), hi_of(avg), -cc[q]));
    out[q] = fma2(ab[q], pk(ng, ng), avg);      // (ua - a g, va - b g)
  }
}

template <int T, int R, int NRG, bool EDGE>
__device__ __forceinline__ void hs_pk_body(const Img& ui, const Img& vi, const Img& uo, const Img& vo, const Img& fx,
                                           const Img& fy, const Img& ft, float* smem) {
  using C = PkCfg<T, R, NRG>;
  constexpr int SW = C::SW, SH = C::SH, HX = C::HX;
  const int b = blockIdx.z;
  const int W = ui.W, H = ui.H;
  const int x0 = blockIdx.x * C::TW - HX;
  const int y0 = blockIdx.y * C::TH - T;
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int sx = 4 * lane, gx = x0 + sx;
  const int r0 = 1 + rg * R, gy0 = y0 + r0;
  const long pitch = ui.pitch;
  const bool okx = (gx >= 0) && (gx < (int)pitch);
  const float* gUi = ui.p + (long)b * ui.stride;
  const float* gVi = vi.p + (long)b * vi.stride;
  auto X = [&](int buf, int g, int which) -> float* { return smem + ((buf * C::XG + g) * 2 + which) * C::ROWF; };

  // ---- tile halo rows (tile rows 0 and SH-1), one warp each, into both buffers ------------------------------------
  if (rg == 0 || rg == NRG - 1) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      if ((e == 0 && rg != 0) || (e == 1 && rg != NRG - 1)) continue;
      const int gy = e == 0 ? y0 : y0 + SH - 1;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
      if (!EDGE || (okx && gy >= 0 && gy < H)) {
        a = ldg4(gUi + (long)gy * pitch + gx);
        c = ldg4(gVi + (long)gy * pitch + gx);
      }
      const f2 cells[4] = {pk(a.x, c.x), pk(a.y, c.y), pk(a.z, c.z), pk(a.w, c.w)};
      row_store(X(0, e == 0 ? 0 : NRG + 1, e == 0 ? 1 : 0), lane, cells);
      row_store(X(1, e == 0 ? 0 : NRG + 1, e == 0 ? 1 : 0), lane, cells);
    }
  }
  // ---- this thread's strip, HBM -> registers --------------------------------------------------------------------------
  f2 s[R][4], ab[R][4];
  float cc[R][4];
  {
    const float* g0 = fx.p + (long)b * fx.stride;
    const float* g1 = fy.p + (long)b * fy.stride;
    const float* g2 = ft.p + (long)b * ft.stride;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int gy = gy0 + j;
      float4 qu = make_float4(0.f, 0.f, 0.f, 0.f), qv = qu, a = qu, c = qu, d = qu;
      if (!EDGE || (okx && gy >= 0 && gy < H)) {
        const long go = (long)gy * pitch + gx;
        qu = ldg4(gUi + go);
        qv = ldg4(gVi + go);
        a = ldg4(g0 + go);
        c = ldg4(g1 + go);
        d = ldg4(g2 + go);
      }
      s[j][0] = pk(qu.x, qv.x); s[j][1] = pk(qu.y, qv.y); s[j][2] = pk(qu.z, qv.z); s[j][3] = pk(qu.w, qv.w);
      ab[j][0] = pk(a.x, c.x); ab[j][1] = pk(a.y, c.y); ab[j][2] = pk(a.z, c.z); ab[j][3] = pk(a.w, c.w);
      cc[j][0] = d.x; cc[j][1] = d.y; cc[j][2] = d.z; cc[j][3] = d.w;
    }
  }
  PkEdge eg;
  eg.left_edge = EDGE && (gx == 0);
  eg.right_j = EDGE ? (W - 1) - gx : -1;
  eg.top_j = EDGE ? -gy0 : -1000;
  eg.bot_j = EDGE ? (H - 1) - gy0 : -1000;
  const f2 c6 = pk(0.16666667f, 0.16666667f), c12 = pk(0.083333336f, 0.083333336f);
  row_store(X(0, rg + 1, 0), lane, s[0]);
  row_store(X(0, rg + 1, 1), lane, s[R - 1]);
  __syncthreads();

  // ---- T sweeps in registers -------------------------------------------------------------------------------------------
#pragma unroll 1
  for (int sw = 0; sw < T; ++sw) {
    const int cur = sw & 1;
    f2 w[3][6];
    {
      f2 t[4];
      row_load(X(cur, rg, 1), lane, t);            // last row of the row group above
      row6<EDGE>(t, eg, w[0]);
    }
    row6<EDGE>(s[0], eg, w[1]);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int A = j % 3, B = (j + 1) % 3, Cc = (j + 2) % 3;
      if (j + 1 < R) {
        row6<EDGE>(s[j + 1], eg, w[Cc]);           // still the previous sweep's values
      } else {
        f2 t[4];
        row_load(X(cur, rg + 2, 0), lane, t);      // first row of the row group below
        row6<EDGE>(t, eg, w[Cc]);
      }
      f2 o[4];
      if (EDGE && j == eg.top_j) row_update(w[Cc], w[B], w[Cc], ab[j], cc[j], c6, c12, o);       // row -1 mirrors to row 1
      else if (EDGE && j == eg.bot_j) row_update(w[A], w[B], w[A], ab[j], cc[j], c6, c12, o);    // row H mirrors to H-2
      else row_update(w[A], w[B], w[Cc], ab[j], cc[j], c6, c12, o);
#pragma unroll
      for (int q = 0; q < 4; ++q) s[j][q] = o[q];
    }
    if (sw + 1 < T) {
      row_store(X(cur ^ 1, rg + 1, 0), lane, s[0]);
      row_store(X(cur ^ 1, rg + 1, 1), lane, s[R - 1]);
      __syncthreads();
    }
  }
  // ---- interior cells -> HBM ----------------------------------------------------------------------------------------------
  float* gU = uo.p + (long)b * uo.stride;
  float* gV = vo.p + (long)b * vo.stride;
  const bool in_cols = (sx >= HX) && (sx < SW - HX) && (gx < W);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int sy = r0 + j, gy = gy0 + j;
    if (in_cols && (sy >= T) && (sy < SH - T) && (gy < H)) {
      const long go = (long)gy * uo.pitch + gx;
      *reinterpret_cast<float4*>(gU + go) = make_float4(lo_of(s[j][0]), lo_of(s[j][1]), lo_of(s[j][2]), lo_of(s[j][3]));
      *reinterpret_cast<float4*>(gV + go) = make_float4(hi_of(s[j][0]), hi_of(s[j][1]), hi_of(s[j][2]), hi_of(s[j][3]));
    }
  }
}

template <int T, int R, int NRG, int MINB>
__global__ void __launch_bounds__(PkCfg<T, R, NRG>::NT, MINB)
hs_pk_kernel(Img ui, Img vi, Img uo, Img vo, Img fx, Img fy, Img ft) {
  using C = PkCfg<T, R, NRG>;
  extern __shared__ __align__(16) float smem[];
  const int x0 = blockIdx.x * C::TW - C::HX, y0 = blockIdx.y * C::TH - T;
  const bool edge = (x0 < 0) || (x0 + C::SW > ui.W) || (y0 < 0) || (y0 + C::SH > ui.H);   // CTA-uniform
  if (edge)
    hs_pk_body<T, R, NRG, true>(ui, vi, uo, vo, fx, fy, ft, smem);
  else
    hs_pk_body<T, R, NRG, false>(ui, vi, uo, vo, fx, fy, ft, smem);
}

template <int T, int R, int NRG, int MINB>
static void launch_cfg(const Img& ui, const Img& vi, const Img& uo, const Img& vo, const Img& fx, const Img& fy,
                       const Img& ft, cudaStream_t s) {
  using C = PkCfg<T, R, NRG>;
  auto kern = hs_pk_kernel<T, R, NRG, MINB>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
  dim3 g((ui.W + C::TW - 1) / C::TW, (ui.H + C::TH - 1) / C::TH, ui.batch);
  kern<<<g, C::NT, C::SMEM_BYTES, s>>>(ui, vi, uo, vo, fx, fy, ft);
}

template <int T>
static void launch_T(int variant, const Img& ui, const Img& vi, const Img& uo, const Img& vo, const Img& fx,
                     const Img& fy, const Img& ft, cudaStream_t s) {
  switch (variant) {
    default:
    case 16: launch_cfg<T, 4, 8, 2>(ui, vi, uo, vo, fx, fy, ft, s); break;     // 34 x 128, 256 threads, 2 CTAs / SM
    case 18: launch_cfg<T, 8, 8, 1>(ui, vi, uo, vo, fx, fy, ft, s); break;     // 66 x 128, 256 threads
  }
}

// T in {1..6, 8}; planes must be 16-byte aligned with pitch % 4 == 0 (checked by the caller)
void launch_hs_packed(int T, int variant, const Img& ui, const Img& vi, const Img& uo, const Img& vo, const Img& fx,
                      const Img& fy, const Img& ft, cudaStream_t s) {
  switch (T) {
    case 1: launch_T<1>(variant, ui, vi, uo, vo, fx, fy, ft, s); break;
    case 2: launch_T<2>(variant, ui, vi, uo, vo, fx, fy, ft, s); break;
    case 3: launch_T<3>(variant, ui, vi, uo, vo, fx, fy, ft, s); break;
    case 4: launch_T<4>(variant, ui, vi, uo, vo, fx, fy, ft, s); break;
    case 5: launch_T<5>(variant, ui, vi, uo, vo, fx, fy, ft, s); break;
    case 6: launch_T<6>(variant, ui, vi, uo, vo, fx, fy, ft, s); break;
    default: launch_T<8>(variant, ui, vi, uo, vo, fx, fy, ft, s); break;
  }
}

}  // namespace ofri
