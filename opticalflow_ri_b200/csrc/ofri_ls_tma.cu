// ofri_ls_tma.cu -- persistent, TMA-fed, register-resident Liu-Shen sweep kernel (sm_100a).
//
// Reference: PhysicsBasedOpticalFlowLiuShen.py:141-156 (iteration loop) and 75-80 (helper); same expression tree as
// the other Liu-Shen kernels (ofri_ls_common.cuh: ls_row_update_regs -> ofri_pixel.cuh: ls_update2), hence
// bit-identical results.  Structure as in ofri_hs_tma.cu:
//   * one CTA per SM walks over the tiles (SH x 128 cells of one pair, SH = R NRG + 2) of the launch;
//   * the TEN planes of a tile (u, v and the 8 coefficient planes) arrive in a shared staging buffer by TMA (zero fill
//     outside the image) and complete on an mbarrier; every thread moves its 4 x R strip -- state AND coefficients --
//     into registers, after which one thread issues the TMA loads of the CTA's next tile, which overlap with
//   * T sweeps done in registers: 3-row sliding window, halo columns by shuffle, the rows above / below a strip
//     through a small double-buffered shared array (one __syncthreads per sweep);
//   * Liu-Shen's boundary rules are not reflections ('nearest' for the difference stencils, zero padding for the
//     8-neighbour sum), so border tiles run an EDGE instantiation that re-applies them every sweep;
//   * the residual sums of every sweep (LS:79) are accumulated per thread over the tile's own output cells, reduced by
//     f32 warp butterfly into the warp's own f64 slot in shared memory, added to errs[pair][k] with one f64 atomic per
//     CTA, pair and sweep;
//   * the stopping rule (LS:141) is evaluated per tile from the previous block's sums: tiles of stopped pairs are
//     skipped (their TMA load is still consumed so the pipeline keeps its phase).
// Algorithmic HBM traffic: 48 B per pixel per launch (read u, v + 8 planes; write u, v) for T sweeps.
#include <type_traits>

#include "ofri_ls_common.cuh"
#include "ofri_tma.cuh"

namespace ofri {

template <int T, int R, int NRG>
struct LtCfg {
  static constexpr int HX = 4;
  static constexpr int SW = 128;
  static constexpr int SH = R * NRG + 2;
  static constexpr int NT = 32 * NRG;
  static constexpr int CTAS = NT <= 128 ? 2 : 1;              // 4-warp CTAs: two per SM, phases of one under the other's arithmetic
  static constexpr int TW = SW - 2 * HX;
  static constexpr int TH = SH - 2 * T;
  static constexpr int PLANE = SH * SW;
  static constexpr int XROWS = 2 * NRG + 2;                  // [tile row 0 | first, last row of every group | tile row SH-1]
  static constexpr int XPLANE = XROWS * SW;
  static constexpr int STAGE_BYTES = 10 * PLANE * 4;
  static constexpr int X_BYTES = 2 * 2 * XPLANE * 4;
  static constexpr int SMEM_BYTES = STAGE_BYTES + X_BYTES + 1024;  // + mbarrier, stop flag, residual accumulators
  static_assert(TW > 0 && TH > 0 && T <= HX && NT <= 1024 && SH <= 256, "bad tile");
  static_assert(CTAS * (SMEM_BYTES + 1024) <= 227 * 1024, "tile does not fit in shared memory");
  static_assert((PLANE * 4) % 128 == 0, "TMA destination alignment");
};

struct LtMaps { CUtensorMap m[10]; };   // u, v, IIx, IIy, II, Ixt, Iyt, B11, B12, B22
struct LtTile { int x0, y0, b; };

// tile list position as a mixed-radix counter (see TmWalk in ofri_hs_tma.cu): no integer divisions per tile
struct LtWalk {
  int b, by, bx, sb, sy, sx, tiles_x, tiles_y;
  __device__ __forceinline__ void init(int tile, int step, int tx, int ty) {
    tiles_x = tx; tiles_y = ty;
    const int per = tx * ty;
    b = tile / per;
    int r = tile - b * per;
    by = r / tx;
    bx = r - by * tx;
    sb = step / per;
    r = step - sb * per;
    sy = r / tx;
    sx = r - sy * tx;
  }
  __device__ __forceinline__ void advance() {
    bx += sx;
    int c = bx >= tiles_x ? 1 : 0;
    bx -= c ? tiles_x : 0;
    by += sy + c;
    c = by >= tiles_y ? 1 : 0;
    by -= c ? tiles_y : 0;
    b += sb + c;
  }
};
template <int T, int R, int NRG>
__device__ __forceinline__ LtTile lt_tile(const LtWalk& w) {
  using C = LtCfg<T, R, NRG>;
  LtTile t;
  t.b = w.b;
  t.x0 = w.bx * C::TW - C::HX;
  t.y0 = w.by * C::TH - T;
  return t;
}

template <bool EDGE>
__device__ __forceinline__ void lt_clamp_x(const LsEdge& eg, float (&du)[6], float (&dv)[6]) {
  if (EDGE) {   // 'nearest': the left neighbour of column 0 is column 0; the right neighbour of column W-1 is column W-1
    if (eg.left_edge) { du[0] = du[1]; dv[0] = dv[1]; }
    if (eg.right_j == 0) { du[2] = du[1]; dv[2] = dv[1]; }
    if (eg.right_j == 1) { du[3] = du[2]; dv[3] = dv[2]; }
    if (eg.right_j == 2) { du[4] = du[3]; dv[4] = dv[3]; }
    if (eg.right_j == 3) { du[5] = du[4]; dv[5] = dv[4]; }
  }
}
template <bool EDGE>
__device__ __forceinline__ void lt_row6_vals(const float (&ru)[4], const float (&rv)[4], const LsEdge& eg,
                                             float (&du)[6], float (&dv)[6]) {
  du[1] = ru[0]; du[2] = ru[1]; du[3] = ru[2]; du[4] = ru[3];
  dv[1] = rv[0]; dv[2] = rv[1]; dv[3] = rv[2]; dv[4] = rv[3];
  du[0] = __shfl_up_sync(0xffffffffu, ru[3], 1);
  du[5] = __shfl_down_sync(0xffffffffu, ru[0], 1);
  dv[0] = __shfl_up_sync(0xffffffffu, rv[3], 1);
  dv[5] = __shfl_down_sync(0xffffffffu, rv[0], 1);
  lt_clamp_x<EDGE>(eg, du, dv);
}
template <bool EDGE>
__device__ __forceinline__ void lt_row6_smem(const float* __restrict__ pu, const float* __restrict__ pv,
                                             const LsEdge& eg, float (&du)[6], float (&dv)[6]) {
  const float4 qu = *reinterpret_cast<const float4*>(pu);
  const float4 qv = *reinterpret_cast<const float4*>(pv);
  const float ru[4] = {qu.x, qu.y, qu.z, qu.w}, rv[4] = {qv.x, qv.y, qv.z, qv.w};
  lt_row6_vals<EDGE>(ru, rv, eg, du, dv);
}

// per-CTA shared bookkeeping behind the staging + exchange buffers
constexpr int LT_STOP_WORDS = 48;   // stop flags of up to 1536 pairs per launch as a bit mask in shared memory
struct LtShared {
  unsigned long long bar;
  int stop;
  int pad;
  double acc[4][12][2];  // residual sums of the current pair, per fused sweep and WARP (own slot: no atomics)
  unsigned stopbits[LT_STOP_WORDS];
};
static_assert(sizeof(LtShared) <= 1024, "LtShared must fit the reserve behind the exchange buffers");

template <int T, int R, int NRG>
__global__ void __launch_bounds__(LtCfg<T, R, NRG>::NT, LtCfg<T, R, NRG>::CTAS)
ls_tma_kernel(const __grid_constant__ LtMaps maps, Img uo, Img vo, int W, int H, int tiles_x, int tiles_y, int ntiles,
              float hpar, int k0, int maxiter, double tol, double* errs, LsBand band) {
  using C = LtCfg<T, R, NRG>;
  constexpr int SW = C::SW, SH = C::SH, HX = C::HX;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* stage = reinterpret_cast<float*>(smem_raw);
  float* xbuf = stage + 10 * C::PLANE;
  LtShared* sh = reinterpret_cast<LtShared*>(smem_raw + C::STAGE_BYTES + C::X_BYTES);
  const unsigned bar = smem_u32(&sh->bar);
  const int tid = threadIdx.x, lane = tid & 31, rg = tid >> 5;
  const int sx = 4 * lane;
  const int r0 = 1 + rg * R;
  int tile = blockIdx.x;
  if (tile >= ntiles) return;
  auto issue = [&](const LtTile& t) {
    mbar_expect_tx(bar, (unsigned)C::STAGE_BYTES);
    const unsigned dst = smem_u32(stage);
#pragma unroll
    for (int c = 0; c < 10; ++c) tma_load_3d(dst + c * C::PLANE * 4, &maps.m[c], t.x0, t.y0, t.b, bar);
  };
  LtWalk walk;
  walk.init(tile, (int)gridDim.x, tiles_x, tiles_y);
  LtTile tl = lt_tile<T, R, NRG>(walk);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    issue(tl);
  }
  if (tid < 4 * 12 * 2) (&sh->acc[0][0][0])[tid] = 0.0;
  // Stopping rule (LS:141) of every pair of the launch, evaluated ONCE per CTA by all threads in parallel (it used to be
  // evaluated by one thread per tile -- global loads + double-precision square roots on the critical path of every
  // tile: 22 % of the kernel's time in the phase profile).  The previous launches are complete (stream order), so the
  // sums are final.  Launches with more pairs than the mask holds fall back to the per-tile evaluation.
  const int npairs = ntiles / (tiles_x * tiles_y);
  const bool use_mask = npairs <= 32 * LT_STOP_WORDS;
  if (use_mask && k0 > 0) {
    for (int w = tid; w < LT_STOP_WORDS; w += C::NT) sh->stopbits[w] = 0u;
    __syncthreads();
    for (int b = tid; b < npairs; b += C::NT)
      if (ls_stopped_before(errs + (long)b * maxiter * 2, k0, tol, band.npix, T)) atomicOr(&sh->stopbits[b >> 5], 1u << (b & 31));
  }
  __syncthreads();
  auto X = [&](int buf, int plane, int g, int which) -> float* {
    return xbuf + ((buf * 2 + plane) * C::XROWS + 2 * g + which - 1) * SW + sx;   // (g = 0: which = 1; g = NRG + 1: which = 0)
  };
  OFRI_PH_INIT;
  unsigned phase = 0;
  int acc_pair = -1;      // pair whose residual sums sit in sh->acc (CTA-uniform)
  auto flush = [&]() {    // all threads; ends with a barrier
    __syncthreads();
    if (acc_pair >= 0 && tid < 2 * T) {
      double v = 0.0;
#pragma unroll
      for (int g = 0; g < NRG; ++g) {
        v += sh->acc[tid >> 1][g][tid & 1];
        sh->acc[tid >> 1][g][tid & 1] = 0.0;
      }
      if (v != 0.0) atomicAdd(errs + ((long)acc_pair * maxiter + k0 + (tid >> 1)) * 2 + (tid & 1), v);
    }
    __syncthreads();
  };
  for (; tile < ntiles; tile += gridDim.x) {
    const bool has_next = tile + (int)gridDim.x < ntiles;
    walk.advance();
    const LtTile nxt_tile = lt_tile<T, R, NRG>(walk);
    if (tl.b != acc_pair) {
      flush();
      acc_pair = tl.b;
    }
    bool stopped;         // CTA-uniform: this pair stopped in an earlier block (LS:141)
    if (use_mask) {
      stopped = k0 > 0 && ((sh->stopbits[tl.b >> 5] >> (tl.b & 31)) & 1u);
    } else {
      if (tid == 0)
        sh->stop = (k0 > 0 && ls_stopped_before(errs + (long)tl.b * maxiter * 2, k0, tol, band.npix, T)) ? 1 : 0;
      stopped = false;    // read after the barrier below
    }
    OFRI_PH(5);
    mbar_wait(bar, phase);
    phase ^= 1;
    OFRI_PH(0);
    // ---- staging buffer -> registers ---------------------------------------------------------------------------------
    const int gx = tl.x0 + sx, gy0 = tl.y0 + r0;
    float u[R][4], v[R][4];
    LsCoefRow k[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int so = (r0 + j) * SW + sx;
      const float4 qu = *reinterpret_cast<const float4*>(stage + 0 * C::PLANE + so);
      const float4 qv = *reinterpret_cast<const float4*>(stage + 1 * C::PLANE + so);
      u[j][0] = qu.x; u[j][1] = qu.y; u[j][2] = qu.z; u[j][3] = qu.w;
      v[j][0] = qv.x; v[j][1] = qv.y; v[j][2] = qv.z; v[j][3] = qv.w;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 q = *reinterpret_cast<const float4*>(stage + (2 + c) * C::PLANE + so);
        k[j].c[c][0] = q.x; k[j].c[c][1] = q.y; k[j].c[c][2] = q.z; k[j].c[c][3] = q.w;
      }
    }
    if (rg == 0) {
      const float4 a = *reinterpret_cast<const float4*>(stage + 0 * C::PLANE + sx);
      const float4 c = *reinterpret_cast<const float4*>(stage + 1 * C::PLANE + sx);
      *reinterpret_cast<float4*>(X(0, 0, 0, 1)) = a;
      *reinterpret_cast<float4*>(X(0, 1, 0, 1)) = c;
      *reinterpret_cast<float4*>(X(1, 0, 0, 1)) = a;
      *reinterpret_cast<float4*>(X(1, 1, 0, 1)) = c;
    }
    if (rg == NRG - 1) {
      const float4 a = *reinterpret_cast<const float4*>(stage + 0 * C::PLANE + (SH - 1) * SW + sx);
      const float4 c = *reinterpret_cast<const float4*>(stage + 1 * C::PLANE + (SH - 1) * SW + sx);
      *reinterpret_cast<float4*>(X(0, 0, NRG + 1, 0)) = a;
      *reinterpret_cast<float4*>(X(0, 1, NRG + 1, 0)) = c;
      *reinterpret_cast<float4*>(X(1, 0, NRG + 1, 0)) = a;
      *reinterpret_cast<float4*>(X(1, 1, NRG + 1, 0)) = c;
    }
    *reinterpret_cast<float4*>(X(0, 0, rg + 1, 0)) = make_float4(u[0][0], u[0][1], u[0][2], u[0][3]);
    *reinterpret_cast<float4*>(X(0, 1, rg + 1, 0)) = make_float4(v[0][0], v[0][1], v[0][2], v[0][3]);
    *reinterpret_cast<float4*>(X(0, 0, rg + 1, 1)) = make_float4(u[R - 1][0], u[R - 1][1], u[R - 1][2], u[R - 1][3]);
    *reinterpret_cast<float4*>(X(0, 1, rg + 1, 1)) = make_float4(v[R - 1][0], v[R - 1][1], v[R - 1][2], v[R - 1][3]);
    __syncthreads();   // staging buffer drained, exchange buffer 0 complete, stop flag visible
    OFRI_PH(2);
    if (has_next && tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      issue(nxt_tile);
    }
    if (!use_mask) stopped = sh->stop != 0;
    if (stopped) {
      __syncthreads();
      tl = nxt_tile;
      continue;
    }
    const bool edge = (tl.x0 < 0) || (tl.x0 + SW > W) || (tl.y0 < 0) || (tl.y0 + SH > H);   // CTA-uniform
    LsEdge eg;
    eg.left_edge = edge && (gx == 0);
    eg.right_j = edge ? (W - 1) - gx : -1;
    eg.top_j = edge ? -gy0 : -1000;
    eg.bot_j = edge ? (H - 1) - gy0 : -1000;
    const bool in_cols = (sx >= HX) && (sx < SW - HX);
    const int own_lo_ = band.own_lo, own_hi_ = band.own_hi;
    float* const pU0 = uo.p + ((long)tl.b * uo.stride + (long)gy0 * uo.pitch + gx);   // + j * pitch per strip row
    float* const pV0 = vo.p + ((long)tl.b * vo.stride + (long)gy0 * vo.pitch + gx);
    const long qU = uo.pitch, qV = vo.pitch;

    // which strip rows count for the residual (the tile's own output cells of the rows this band owns: every pixel is
    // counted by exactly one tile) and which are stored -- one bit per row, so the sweep body has no per-row branches
    // (they end ptxas' scheduling regions: 6 % of the sweep's samples were branch_resolving in the round-1 profile)
    unsigned resmask = 0u, stmask = 0u;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int sy = r0 + j, gy = gy0 + j;
      const bool own = in_cols && (sy >= T) && (sy < SH - T) && (gy < H);
      if (own && gy >= own_lo_ && gy < own_hi_) resmask |= 1u << j;
      if (own && gx < W) stmask |= 1u << j;
    }
    auto sweep_all = [&](auto edge_tag) {
      constexpr bool EDGE = decltype(edge_tag)::value;
      float pu2 = 0.0f, pv2 = 0.0f;        // residual partial sums of the PREVIOUS sweep, reduced under this sweep's arithmetic
      auto reduce_pending = [&](int sweep) {
        // f32 butterfly over the warp (a balanced tree: relative error ~1e-7, below the f32 rounding of the reference's
        // own BLAS norm), then f64 accumulation in the warp's own shared slot (no atomics)
        float su = pu2, sv = pv2;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          su = fadd(su, __shfl_xor_sync(0xffffffffu, su, o));
          sv = fadd(sv, __shfl_xor_sync(0xffffffffu, sv, o));
        }
        if (lane == 0) {
          sh->acc[sweep][rg][0] += (double)su;
          sh->acc[sweep][rg][1] += (double)sv;
        }
      };
#pragma unroll 1
      for (int s = 0; s < T; ++s) {
        const int cur = s & 1;
        float du2 = 0.0f, dv2 = 0.0f;
        float wu[3][6], wv[3][6];
        lt_row6_smem<EDGE>(X(cur, 0, rg, 1), X(cur, 1, rg, 1), eg, wu[0], wv[0]);
        lt_row6_vals<EDGE>(u[0], v[0], eg, wu[1], wv[1]);
        if (s > 0) reduce_pending(s - 1);  // its 10 dependent shuffles hide under the rows below instead of before the barrier
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int A = j % 3, B = (j + 1) % 3, Cc = (j + 2) % 3;
          if (j + 1 < R)
            lt_row6_vals<EDGE>(u[j + 1], v[j + 1], eg, wu[Cc], wv[Cc]);
          else
            lt_row6_smem<EDGE>(X(cur, 0, rg + 2, 0), X(cur, 1, rg + 2, 0), eg, wu[Cc], wv[Cc]);
          float ou[4], ov[4];
          if (EDGE && j == eg.top_j)        // global row 0: 'nearest' -> the row above is the row itself; H8: zero
            ls_row_update_regs<EDGE>(wu[B], wu[B], wu[Cc], wv[B], wv[B], wv[Cc], k[j], hpar, true, false, eg, ou, ov);
          else if (EDGE && j == eg.bot_j)
            ls_row_update_regs<EDGE>(wu[A], wu[B], wu[B], wv[A], wv[B], wv[B], k[j], hpar, false, true, eg, ou, ov);
          else
            ls_row_update_regs<EDGE>(wu[A], wu[B], wu[Cc], wv[A], wv[B], wv[Cc], k[j], hpar, false, false, eg, ou, ov);
          // residual of the row (LS:79), weighted 1 / 0 by the row's bit; EDGE: cells beyond the last column do not count
          float ru = 0.0f, rv = 0.0f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float eu = fsub(ou[q], wu[B][q + 1]), ev = fsub(ov[q], wv[B][q + 1]);
            if (EDGE && gx + q >= W) { eu = 0.0f; ev = 0.0f; }
            ru = fmaf(eu, eu, ru);
            rv = fmaf(ev, ev, rv);
          }
          const float wj = (resmask >> j) & 1u ? 1.0f : 0.0f;
          du2 = fmaf(wj, ru, du2);
          dv2 = fmaf(wj, rv, dv2);
#pragma unroll
          for (int q = 0; q < 4; ++q) { u[j][q] = ou[q]; v[j][q] = ov[q]; }
        }
        pu2 = du2;
        pv2 = dv2;
        if (s + 1 < T) {
          const int nxt = cur ^ 1;
          *reinterpret_cast<float4*>(X(nxt, 0, rg + 1, 0)) = make_float4(u[0][0], u[0][1], u[0][2], u[0][3]);
          *reinterpret_cast<float4*>(X(nxt, 1, rg + 1, 0)) = make_float4(v[0][0], v[0][1], v[0][2], v[0][3]);
          *reinterpret_cast<float4*>(X(nxt, 0, rg + 1, 1)) = make_float4(u[R - 1][0], u[R - 1][1], u[R - 1][2], u[R - 1][3]);
          *reinterpret_cast<float4*>(X(nxt, 1, rg + 1, 1)) = make_float4(v[R - 1][0], v[R - 1][1], v[R - 1][2], v[R - 1][3]);
          __syncthreads();
        }
      }
      reduce_pending(T - 1);
    };
    if (edge) sweep_all(std::true_type{});
    else sweep_all(std::false_type{});
    // the tile's own cells -> HBM from the registers of the last sweep (4 rows x 2 planes per thread)
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if ((stmask >> j) & 1u) {
        *reinterpret_cast<float4*>(pU0 + j * qU) = make_float4(u[j][0], u[j][1], u[j][2], u[j][3]);
        *reinterpret_cast<float4*>(pV0 + j * qV) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
      }
    }
    OFRI_PH(3);
    __syncthreads();   // exchange buffers free for the next tile
    OFRI_PH(4);
    tl = nxt_tile;
  }
  flush();
  OFRI_PH_FLUSH;
}

template <int T, int R, int NRG>
static bool launch_cfg(const Img& ui, const Img& vi, const Img& uo, const Img& vo, const LsPlanes& co, float hpar, int k0,
                       int maxiter, double tol, double* errs, const LsBand& band, int num_sms, cudaStream_t s) {
  using C = LtCfg<T, R, NRG>;
  LtMaps maps;
  if (!make_map(&maps.m[0], ui, C::SH) || !make_map(&maps.m[1], vi, C::SH)) return false;
  for (int c = 0; c < 8; ++c)
    if (!make_map(&maps.m[2 + c], co.c[c], C::SH)) return false;
  auto kern = ls_tma_kernel<T, R, NRG>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (C::CTAS > 1)   // two CTAs per SM only fit with the largest shared-memory carve-out
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  const int tiles_x = (ui.W + C::TW - 1) / C::TW, tiles_y = (ui.H + C::TH - 1) / C::TH;
  const long ntiles = (long)tiles_x * tiles_y * ui.batch;
  if (ntiles > 0x7fffffffL) return false;
  const int slots = num_sms * C::CTAS;
  const int grid = (int)(ntiles < slots ? ntiles : slots);
  kern<<<grid, C::NT, C::SMEM_BYTES, s>>>(maps, uo, vo, ui.W, ui.H, tiles_x, tiles_y, (int)ntiles, hpar, k0, maxiter, tol,
                                          errs, band);
  return true;
}

// One fused block of T sweeps (T in 2..4) ui, vi -> uo, vo.  false = not applicable (caller uses the other kernels).
bool launch_ls_tma(int T, const Img& ui, const Img& vi, const Img& uo, const Img& vo, const LsPlanes& co, float hpar,
                   int k0, int maxiter, double tol, double* errs, const LsBand& band, cudaStream_t s, int variant) {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  (void)variant;
  switch (T) {
    case 2: return launch_cfg<2, 4, 8>(ui, vi, uo, vo, co, hpar, k0, maxiter, tol, errs, band, num_sms, s);
    case 3: return launch_cfg<3, 4, 8>(ui, vi, uo, vo, co, hpar, k0, maxiter, tol, errs, band, num_sms, s);
    case 4: return launch_cfg<4, 4, 8>(ui, vi, uo, vo, co, hpar, k0, maxiter, tol, errs, band, num_sms, s);
    default: return false;
  }
}

// phase-timing table (see ofri_hs_tma.cu): [0] mbarrier wait, [2] staging -> registers, [3] sweeps + stores, [4] final
// barrier, [5] per-tile bookkeeping (decode, residual flush, stop flag)
void ls_tma_phase_read(unsigned long long* out) { OFRI_PH_READ(out); }

}  // namespace ofri
