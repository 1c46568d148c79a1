// ofri_hs_tma.cu -- persistent, TMA-fed, register-resident Horn-Schunck Jacobi kernel (sm_100a).
//
// Reference: HornSchunck.py:52-71 (HS_helper / HS_helper2), fast arithmetic; same expression tree as every other
// Horn-Schunck kernel here (ofri_hs_common.cuh: hs_row_update), hence bit-identical results.
//
// One CTA per SM walks over the tiles of the launch (tile = SH x 128 cells of one pair, SH = R NRG + 2; static
// round-robin).  Per tile:
//   1. the five planes of the tile (U, V and the prepared coefficients a, b, c) arrive in a shared staging buffer by
//      TMA (cp.async.bulk.tensor, 3-D tensor maps [batch][H][W]; out-of-image elements are zero-filled by the copy
//      engine, so border tiles need no address logic) and complete on an mbarrier;
//      tiles that touch the image border then turn the out-of-image frame into GHOST cells: columns -1..-HX become
//      copies of columns 1..HX (W..W+HX-1 of W-2..W-1-HX), rows likewise, for all five planes.  scipy's 'mirror'
//      rule is a reflection about the border pixel and every operation of the update is symmetric under it (the
//      stencil sums commute), so a ghost cell evolves bit-identically to the cell it mirrors: the boundary condition
//      holds at every fused sweep with NO boundary code in the sweep itself;
//   2. every thread moves its 4 x R strip from the staging buffer into REGISTERS (LDS.128) -- after a CTA barrier the
//      staging buffer is free again and one thread immediately issues the TMA loads of the CTA's NEXT tile, which
//      then overlap with
//   3. T Jacobi sweeps done entirely in registers (3-row sliding window, halo columns by warp shuffle, the rows
//      above / below a strip exchanged through a small double-buffered shared array, one __syncthreads per sweep), and
//   4. float4 stores of the (SH - 2T) x (128 - 2 HX) interior straight from registers to HBM.  (Round 2 measured two
//      ways of issuing these stores row by row inside the last sweep so that they drain under its arithmetic: a peeled
//      last sweep saved 6.8 Mcycles of the 19.5 the burst costs per CTA but the sweeps got 8.6 % slower, predicated
//      stores in the one sweep body made every sweep 25 % slower (a branch per row ends the scheduling region) --
//      both were net losses, profiles/r2_phase_hs_store_variants.jsonl; the burst after the sweeps stays.)
// HBM reads of tile i+1 therefore run under the arithmetic of tile i, with a single CTA (8-12 warps, up to 255
// registers per thread) per SM and no redundant staging of state through shared memory during the sweeps.
// Algorithmic HBM traffic: 28 B per pixel per launch (read U, V, a, b, c; write U, V) for T sweeps.
//
// Tried in round 2 and removed again (profiles/r2_cluster_experiment.md): the same kernel as thread-block CLUSTERS of two
// CTAs that form one 130-row tile, the row between the halves exchanged every sweep through distributed shared memory
// (st.async + mbarrier in the partner's shared memory, split cluster barrier per tile, the lower CTA holding its half
// upside down so that both run the same sweep code).  Bit-identical and dead-lock free, but slower: the row group that
// talks to the partner carries the extra mbarrier test / arm / st.async latency on its critical path in EVERY sweep and
// all other row groups wait for it at the sweep's CTA barrier (+21 % per sweep, phase profile), which eats the 14 % more
// useful rows per tile: 56.8 ms per 64 pairs at T = 8 against 55.2 (T = 8) / 54.2 (T = 4) for this kernel.
#include "ofri_hs_common.cuh"
#include "ofri_tma.cuh"

namespace ofri {

template <int T, int R, int NRG>
struct TmCfg {
  static constexpr int HX = (T <= 4) ? 4 : 8;
  static constexpr int SW = 128;
  static constexpr int SH = R * NRG + 2;
  static constexpr int NT = 32 * NRG;
  static constexpr int CTAS = NT <= 128 ? 2 : 1;              // 4-warp CTAs: two per SM (255 registers x 128 threads each), so
                                                              // one CTA's load / store phases run under the other's arithmetic
  static constexpr int TW = SW - 2 * HX;
  static constexpr int TH = SH - 2 * T;
  static constexpr int PLANE = SH * SW;                      // floats per staged plane
  static constexpr int XG = NRG + 2;
  static constexpr int XPLANE = XG * 2 * SW;                 // [g][top / bottom][SW]
  static constexpr int STAGE_BYTES = 5 * PLANE * 4;
  static constexpr int X_BYTES = 2 * 2 * XPLANE * 4;         // [buffer][U / V]
  static constexpr int SMEM_BYTES = STAGE_BYTES + X_BYTES + 64;
  static_assert(TW > 0 && TH > 0 && HX >= T && NT <= 1024 && SH <= 256, "bad tile");
  static_assert(CTAS * (SMEM_BYTES + 1024) <= 227 * 1024, "tile does not fit in shared memory");
  static_assert((PLANE * 4) % 128 == 0, "TMA destination alignment");
};

struct TmTile { int x0, y0, b; };

// Position of a CTA in the launch's tile list as a mixed-radix counter (pair, tile row, tile column): advancing by the
// grid size is three adds with carries instead of two integer divisions per tile (which sat on the critical path of
// every tile: 5 % of the kernel's time in the phase profile, profiles/r2_phase_*.jsonl).
struct TmWalk {
  int b, by, bx;          // current tile
  int sb, sy, sx;         // gridDim.x decomposed in the same radix
  int tiles_x, tiles_y;
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
__device__ __forceinline__ TmTile tm_tile(const TmWalk& w, int4 rows) {
  using C = TmCfg<T, R, NRG>;
  TmTile t;
  t.b = w.b;
  const int by = w.by + (w.by < rows.x ? rows.y : rows.z);   // tile-row subset of a split launch (rows = {cut, off0, off1, -}); identity otherwise
  t.x0 = w.bx * C::TW - C::HX;
  t.y0 = by * C::TH - T;
  return t;
}

// Border tiles: fill the ghost frame of the staged planes (see the file comment).  x first, then y over the full tile
// width, so the corners come out right.  Needs W > HX and H > T (checked on the host).
template <int T, int R, int NRG>
__device__ __forceinline__ void hs_tma_ghosts(const TmTile& tl, int W, int H, float* stage) {
  using C = TmCfg<T, R, NRG>;
  constexpr int SW = C::SW, SH = C::SH, HX = C::HX;
  const int tid = threadIdx.x;
  const bool left = tl.x0 < 0;                       // then x0 == -HX: global column 0 is tile column HX
  const int cw = (W - 1) - tl.x0;                    // tile column of global column W-1
  const bool right = cw + 1 < SW;
  if (left || right) {
    for (int i = tid; i < 5 * SH * HX; i += C::NT) {
      const int k = 1 + i % HX, row = (i / HX) % SH, pl = i / (HX * SH);
      float* r = stage + pl * C::PLANE + row * SW;
      if (left) r[HX - k] = r[HX + k];
      if (right && cw + k < SW) r[cw + k] = r[cw - k];
    }
    __syncthreads();
  }
  const bool top = tl.y0 < 0;                        // then y0 == -T: global row 0 is tile row T
  const int rh = (H - 1) - tl.y0;                    // tile row of global row H-1
  const bool bottom = rh + 1 < SH;
  if (top || bottom) {
    for (int i = tid; i < 5 * T * (SW / 4); i += C::NT) {
      const int c4 = i % (SW / 4), k = 1 + (i / (SW / 4)) % T, pl = i / ((SW / 4) * T);
      float* p = stage + pl * C::PLANE + 4 * c4;
      if (top) *reinterpret_cast<float4*>(p + (T - k) * SW) = *reinterpret_cast<const float4*>(p + (T + k) * SW);
      if (bottom && rh + k < SH)
        *reinterpret_cast<float4*>(p + (rh + k) * SW) = *reinterpret_cast<const float4*>(p + (rh - k) * SW);
    }
    __syncthreads();
  }
}

template <int T, int R, int NRG>
__device__ __forceinline__ void hs_tma_tile(const TmTile& tl, int W, int H, const Img& uo, const Img& vo,
                                            const float* stage, float* xbuf, bool issue_next, const TmTile& nx,
                                            const CUtensorMap* mU, const CUtensorMap* mV, const CUtensorMap* mA,
                                            const CUtensorMap* mB, const CUtensorMap* mC, unsigned bar) {
  using C = TmCfg<T, R, NRG>;
  constexpr int SW = C::SW, SH = C::SH, HX = C::HX;
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int sx = 4 * lane, gx = tl.x0 + sx;
  const int r0 = 1 + rg * R, gy0 = tl.y0 + r0;
  auto X = [&](int buf, int plane, int g, int which) -> float* {
    return xbuf + ((buf * 2 + plane) * C::XG + g) * 2 * SW + which * SW + sx;
  };
  // ---- staging buffer -> registers -----------------------------------------------------------------------------------
  float u[R][4], v[R][4];
  HsCoef<false> k[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int so = (r0 + j) * SW + sx;
    const float4 qu = *reinterpret_cast<const float4*>(stage + 0 * C::PLANE + so);
    const float4 qv = *reinterpret_cast<const float4*>(stage + 1 * C::PLANE + so);
    const float4 a = *reinterpret_cast<const float4*>(stage + 2 * C::PLANE + so);
    const float4 c = *reinterpret_cast<const float4*>(stage + 3 * C::PLANE + so);
    const float4 d = *reinterpret_cast<const float4*>(stage + 4 * C::PLANE + so);
    u[j][0] = qu.x; u[j][1] = qu.y; u[j][2] = qu.z; u[j][3] = qu.w;
    v[j][0] = qv.x; v[j][1] = qv.y; v[j][2] = qv.z; v[j][3] = qv.w;
    k[j].c0[0] = a.x; k[j].c0[1] = a.y; k[j].c0[2] = a.z; k[j].c0[3] = a.w;
    k[j].c1[0] = c.x; k[j].c1[1] = c.y; k[j].c1[2] = c.z; k[j].c1[3] = c.w;
    k[j].c2[0] = d.x; k[j].c2[1] = d.y; k[j].c2[2] = d.z; k[j].c2[3] = d.w;
  }
  // tile halo rows (tile rows 0 and SH-1) -> both exchange buffers; first / last strip row -> buffer 0
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
  __syncthreads();   // every thread has drained the staging buffer; exchange buffer 0 is complete
  OFRI_PH(2);
  // ---- prefetch the CTA's next tile under this tile's arithmetic ---------------------------------------------------------
  if (issue_next && threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads above -> async-proxy writes below
    mbar_expect_tx(bar, (unsigned)C::STAGE_BYTES);
    const unsigned dst = smem_u32(stage);
    tma_load_3d(dst + 0 * C::PLANE * 4, mU, nx.x0, nx.y0, nx.b, bar);
    tma_load_3d(dst + 1 * C::PLANE * 4, mV, nx.x0, nx.y0, nx.b, bar);
    tma_load_3d(dst + 2 * C::PLANE * 4, mA, nx.x0, nx.y0, nx.b, bar);
    tma_load_3d(dst + 3 * C::PLANE * 4, mB, nx.x0, nx.y0, nx.b, bar);
    tma_load_3d(dst + 4 * C::PLANE * 4, mC, nx.x0, nx.y0, nx.b, bar);
  }
  const HsEdge eg = {false, -1, -1000, -1000};   // unused: the ghost frame carries the boundary condition

  // ---- T sweeps in registers (no boundary code: see hs_tma_ghosts) ----------------------------------------------------------------------------------------------
#pragma unroll 1
  for (int s = 0; s < T; ++s) {
    const int cur = s & 1;
    float wu[3][6], wv[3][6];
    hs_row6_smem<false>(X(cur, 0, rg, 1), X(cur, 1, rg, 1), eg, wu[0], wv[0]);         // last row of the group above
    hs_row6_vals<false>(u[0], v[0], eg, wu[1], wv[1]);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int A = j % 3, B = (j + 1) % 3, Cc = (j + 2) % 3;
      if (j + 1 < R)
        hs_row6_vals<false>(u[j + 1], v[j + 1], eg, wu[Cc], wv[Cc]);                   // still the previous sweep's values
      else
        hs_row6_smem<false>(X(cur, 0, rg + 2, 0), X(cur, 1, rg + 2, 0), eg, wu[Cc], wv[Cc]);  // first row of the group below
      float ou[4], ov[4];
      hs_row_update<false>(wu[A], wu[B], wu[Cc], wv[A], wv[B], wv[Cc], k[j], ou, ov);
#pragma unroll
      for (int q = 0; q < 4; ++q) { u[j][q] = ou[q]; v[j][q] = ov[q]; }
    }
    if (s + 1 < T) {
      const int nxt = cur ^ 1;
      *reinterpret_cast<float4*>(X(nxt, 0, rg + 1, 0)) = make_float4(u[0][0], u[0][1], u[0][2], u[0][3]);
      *reinterpret_cast<float4*>(X(nxt, 1, rg + 1, 0)) = make_float4(v[0][0], v[0][1], v[0][2], v[0][3]);
      *reinterpret_cast<float4*>(X(nxt, 0, rg + 1, 1)) = make_float4(u[R - 1][0], u[R - 1][1], u[R - 1][2], u[R - 1][3]);
      *reinterpret_cast<float4*>(X(nxt, 1, rg + 1, 1)) = make_float4(v[R - 1][0], v[R - 1][1], v[R - 1][2], v[R - 1][3]);
      __syncthreads();
    }
  }
  OFRI_PH(3);
  // ---- interior cells -> HBM ------------------------------------------------------------------------------------------------
  // strip rows j with T <= r0 + j < SH - T inside the image; one 64-bit base address per plane and tile, then + pitch
  // per row (the per-row 64-bit index arithmetic it replaces was 0.7 of a sweep's instructions per tile)
  const bool in_cols = (sx >= HX) && (sx < SW - HX) && (gx < W);
  const int j_lo = T - r0 > 0 ? T - r0 : 0;
  int j_hi = SH - T - r0;
  if (H - gy0 < j_hi) j_hi = H - gy0;
  float* pU = uo.p + ((long)tl.b * uo.stride + (long)gy0 * uo.pitch + gx);
  float* pV = vo.p + ((long)tl.b * vo.stride + (long)gy0 * vo.pitch + gx);
  const long qU = uo.pitch, qV = vo.pitch;
  if (in_cols) {
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (j >= j_lo && j < j_hi) {
        *reinterpret_cast<float4*>(pU) = make_float4(u[j][0], u[j][1], u[j][2], u[j][3]);
        *reinterpret_cast<float4*>(pV) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
      }
      pU += qU;
      pV += qV;
    }
  }
  __syncthreads();   // the exchange buffers are free for the next tile
  OFRI_PH(4);
}


// ---------------------------------------------------------------------------------------------------------------
// PRECISE tile body: the reference's arithmetic bit for bit (float64-accumulated stencil rounded once, separately
// rounded float32 update with a correctly rounded division: hs_avg_cols_precise + hs_update_precise).  The strip's
// U, V live in registers as DOUBLES (exact images of the float32 values), so a value is widened once per sweep
// instead of once per use; halo columns are exchanged as doubles by shuffle, the rows above / below as float32
// through the same shared exchange array.  Coefficient planes are the RAW fx, fy, ft; den and 1/den per strip cell
// are computed once per launch.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double shfl_up_d(double x) {
  int lo = __double2loint(x), hi = __double2hiint(x);
  lo = __shfl_up_sync(0xffffffffu, lo, 1);
  hi = __shfl_up_sync(0xffffffffu, hi, 1);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_down_d(double x) {
  int lo = __double2loint(x), hi = __double2hiint(x);
  lo = __shfl_down_sync(0xffffffffu, lo, 1);
  hi = __shfl_down_sync(0xffffffffu, hi, 1);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void row6_state_d(const double (&ru)[4], const double (&rv)[4], double (&du)[6],
                                             double (&dv)[6]) {
  du[1] = ru[0]; du[2] = ru[1]; du[3] = ru[2]; du[4] = ru[3];
  dv[1] = rv[0]; dv[2] = rv[1]; dv[3] = rv[2]; dv[4] = rv[3];
  du[0] = shfl_up_d(ru[3]);
  du[5] = shfl_down_d(ru[0]);
  dv[0] = shfl_up_d(rv[3]);
  dv[5] = shfl_down_d(rv[0]);
}
__device__ __forceinline__ void row6_smem_d(const float* __restrict__ pu, const float* __restrict__ pv, double (&du)[6],
                                            double (&dv)[6]) {
  const float4 qu = *reinterpret_cast<const float4*>(pu);
  const float4 qv = *reinterpret_cast<const float4*>(pv);
  const float ul = __shfl_up_sync(0xffffffffu, qu.w, 1), ur = __shfl_down_sync(0xffffffffu, qu.x, 1);
  const float vl = __shfl_up_sync(0xffffffffu, qv.w, 1), vr = __shfl_down_sync(0xffffffffu, qv.x, 1);
  du[0] = (double)ul; du[1] = (double)qu.x; du[2] = (double)qu.y; du[3] = (double)qu.z; du[4] = (double)qu.w; du[5] = (double)ur;
  dv[0] = (double)vl; dv[1] = (double)qv.x; dv[2] = (double)qv.y; dv[3] = (double)qv.z; dv[4] = (double)qv.w; dv[5] = (double)vr;
}
struct HsCoefP { float fx[4], fy[4], ft[4], den[4], rcp[4]; };
__device__ __forceinline__ void row_update_d(const double (&uu)[6], const double (&um)[6], const double (&ud)[6],
                                             const double (&vu)[6], const double (&vm)[6], const double (&vd)[6],
                                             const HsCoefP& k, float (&ou)[4], float (&ov)[4]) {
  double vsu[6], vsv[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    vsu[c] = dadd(uu[c], ud[c]);
    vsv[c] = dadd(vu[c], vd[c]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float ua = hs_avg_cols_precise(vsu[q], vsu[q + 1], vsu[q + 2], um[q], um[q + 2]);
    const float va = hs_avg_cols_precise(vsv[q], vsv[q + 1], vsv[q + 2], vm[q], vm[q + 2]);
    hs_update_precise(ua, va, k.fx[q], k.fy[q], k.ft[q], k.den[q], k.rcp[q], &ou[q], &ov[q]);
  }
}

template <int T, int R, int NRG>
__device__ __forceinline__ void hs_tma_tile_precise(const TmTile& tl, int W, int H, const Img& uo, const Img& vo,
                                                    const float* stage, float* xbuf, bool issue_next,
                                                    const TmTile& nx, const CUtensorMap* mU, const CUtensorMap* mV,
                                                    const CUtensorMap* mA, const CUtensorMap* mB,
                                                    const CUtensorMap* mC, unsigned bar, float alpha2) {
  using C = TmCfg<T, R, NRG>;
  constexpr int SW = C::SW, SH = C::SH, HX = C::HX;
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int sx = 4 * lane, gx = tl.x0 + sx;
  const int r0 = 1 + rg * R, gy0 = tl.y0 + r0;
  auto X = [&](int buf, int plane, int g, int which) -> float* {
    return xbuf + ((buf * 2 + plane) * C::XG + g) * 2 * SW + which * SW + sx;
  };
  double u[R][4], v[R][4];
  HsCoefP k[R];
  float4 eu0, ev0, eu1, ev1;      // float32 copies of the strip's first / last row (what gets published)
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int so = (r0 + j) * SW + sx;
    const float4 qu = *reinterpret_cast<const float4*>(stage + 0 * C::PLANE + so);
    const float4 qv = *reinterpret_cast<const float4*>(stage + 1 * C::PLANE + so);
    const float4 a = *reinterpret_cast<const float4*>(stage + 2 * C::PLANE + so);
    const float4 c = *reinterpret_cast<const float4*>(stage + 3 * C::PLANE + so);
    const float4 d = *reinterpret_cast<const float4*>(stage + 4 * C::PLANE + so);
    u[j][0] = (double)qu.x; u[j][1] = (double)qu.y; u[j][2] = (double)qu.z; u[j][3] = (double)qu.w;
    v[j][0] = (double)qv.x; v[j][1] = (double)qv.y; v[j][2] = (double)qv.z; v[j][3] = (double)qv.w;
    if (j == 0) { eu0 = qu; ev0 = qv; }
    if (j == R - 1) { eu1 = qu; ev1 = qv; }
    k[j].fx[0] = a.x; k[j].fx[1] = a.y; k[j].fx[2] = a.z; k[j].fx[3] = a.w;
    k[j].fy[0] = c.x; k[j].fy[1] = c.y; k[j].fy[2] = c.z; k[j].fy[3] = c.w;
    k[j].ft[0] = d.x; k[j].ft[1] = d.y; k[j].ft[2] = d.z; k[j].ft[3] = d.w;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      k[j].den[q] = hs_den(k[j].fx[q], k[j].fy[q], alpha2);
      k[j].rcp[q] = rcp_rn(k[j].den[q]);
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
  *reinterpret_cast<float4*>(X(0, 0, rg + 1, 0)) = eu0;
  *reinterpret_cast<float4*>(X(0, 1, rg + 1, 0)) = ev0;
  *reinterpret_cast<float4*>(X(0, 0, rg + 1, 1)) = eu1;
  *reinterpret_cast<float4*>(X(0, 1, rg + 1, 1)) = ev1;
  __syncthreads();
  OFRI_PH(2);
  if (issue_next && threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    mbar_expect_tx(bar, (unsigned)C::STAGE_BYTES);
    const unsigned dst = smem_u32(stage);
    tma_load_3d(dst + 0 * C::PLANE * 4, mU, nx.x0, nx.y0, nx.b, bar);
    tma_load_3d(dst + 1 * C::PLANE * 4, mV, nx.x0, nx.y0, nx.b, bar);
    tma_load_3d(dst + 2 * C::PLANE * 4, mA, nx.x0, nx.y0, nx.b, bar);
    tma_load_3d(dst + 3 * C::PLANE * 4, mB, nx.x0, nx.y0, nx.b, bar);
    tma_load_3d(dst + 4 * C::PLANE * 4, mC, nx.x0, nx.y0, nx.b, bar);
  }
  const bool in_cols = (sx >= HX) && (sx < SW - HX) && (gx < W);
  const int j_lo = T - r0 > 0 ? T - r0 : 0;          // strip rows stored: T <= r0 + j < SH - T, inside the image
  int j_hi = SH - T - r0;
  if (H - gy0 < j_hi) j_hi = H - gy0;
  float* const pU0 = uo.p + ((long)tl.b * uo.stride + (long)gy0 * uo.pitch + gx);
  float* const pV0 = vo.p + ((long)tl.b * vo.stride + (long)gy0 * vo.pitch + gx);
  const long qU = uo.pitch, qV = vo.pitch;
#pragma unroll 1
  for (int s = 0; s < T; ++s) {
    const int cur = s & 1;
    const bool last = s + 1 == T;
    double wu[3][6], wv[3][6];
    row6_smem_d(X(cur, 0, rg, 1), X(cur, 1, rg, 1), wu[0], wv[0]);
    row6_state_d(u[0], v[0], wu[1], wv[1]);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int A = j % 3, B = (j + 1) % 3, Cc = (j + 2) % 3;
      if (j + 1 < R)
        row6_state_d(u[j + 1], v[j + 1], wu[Cc], wv[Cc]);
      else
        row6_smem_d(X(cur, 0, rg + 2, 0), X(cur, 1, rg + 2, 0), wu[Cc], wv[Cc]);
      float ou[4], ov[4];
      row_update_d(wu[A], wu[B], wu[Cc], wv[A], wv[B], wv[Cc], k[j], ou, ov);
#pragma unroll
      for (int q = 0; q < 4; ++q) { u[j][q] = (double)ou[q]; v[j][q] = (double)ov[q]; }
      if (j == 0) { eu0 = make_float4(ou[0], ou[1], ou[2], ou[3]); ev0 = make_float4(ov[0], ov[1], ov[2], ov[3]); }
      if (j == R - 1) { eu1 = make_float4(ou[0], ou[1], ou[2], ou[3]); ev1 = make_float4(ov[0], ov[1], ov[2], ov[3]); }
      if (last && in_cols && j >= j_lo && j < j_hi) {   // interior cells -> HBM straight from the float32 results
        *reinterpret_cast<float4*>(pU0 + j * qU) = make_float4(ou[0], ou[1], ou[2], ou[3]);
        *reinterpret_cast<float4*>(pV0 + j * qV) = make_float4(ov[0], ov[1], ov[2], ov[3]);
      }
    }
    if (!last) {
      const int nxt = cur ^ 1;
      *reinterpret_cast<float4*>(X(nxt, 0, rg + 1, 0)) = eu0;
      *reinterpret_cast<float4*>(X(nxt, 1, rg + 1, 0)) = ev0;
      *reinterpret_cast<float4*>(X(nxt, 0, rg + 1, 1)) = eu1;
      *reinterpret_cast<float4*>(X(nxt, 1, rg + 1, 1)) = ev1;
      __syncthreads();
    }
  }
  OFRI_PH(3);
  __syncthreads();
  OFRI_PH(4);
}

template <int T, int R, int NRG, bool PRECISE>
__global__ void __launch_bounds__(TmCfg<T, R, NRG>::NT, TmCfg<T, R, NRG>::CTAS)
hs_tma_kernel(const __grid_constant__ CUtensorMap mU, const __grid_constant__ CUtensorMap mV,
              const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB,
              const __grid_constant__ CUtensorMap mC, Img uo, Img vo, int W, int H, int tiles_x, int tiles_y,
              int ntiles, float alpha2, int4 rows) {
  using C = TmCfg<T, R, NRG>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* stage = reinterpret_cast<float*>(smem_raw);
  float* xbuf = stage + 5 * C::PLANE;
  const unsigned bar = smem_u32(smem_raw + C::STAGE_BYTES + C::X_BYTES);
  int tile = blockIdx.x;
  if (tile >= ntiles) return;
  TmWalk walk;
  walk.init(tile, (int)gridDim.x, tiles_x, tiles_y);
  TmTile tl = tm_tile<T, R, NRG>(walk, rows);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    mbar_expect_tx(bar, (unsigned)C::STAGE_BYTES);
    const unsigned dst = smem_u32(stage);
    tma_load_3d(dst + 0 * C::PLANE * 4, &mU, tl.x0, tl.y0, tl.b, bar);
    tma_load_3d(dst + 1 * C::PLANE * 4, &mV, tl.x0, tl.y0, tl.b, bar);
    tma_load_3d(dst + 2 * C::PLANE * 4, &mA, tl.x0, tl.y0, tl.b, bar);
    tma_load_3d(dst + 3 * C::PLANE * 4, &mB, tl.x0, tl.y0, tl.b, bar);
    tma_load_3d(dst + 4 * C::PLANE * 4, &mC, tl.x0, tl.y0, tl.b, bar);
  }
  __syncthreads();
  OFRI_PH_INIT;
  unsigned phase = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    const bool has_next = tile + (int)gridDim.x < ntiles;
    walk.advance();
    const TmTile nx = tm_tile<T, R, NRG>(walk, rows);     // only used if has_next
    const bool edge = (tl.x0 < 0) || (tl.x0 + C::SW > W) || (tl.y0 < 0) || (tl.y0 + C::SH > H);   // CTA-uniform
    OFRI_PH(5);
    mbar_wait(bar, phase);
    phase ^= 1;
    OFRI_PH(0);
    if (edge) hs_tma_ghosts<T, R, NRG>(tl, W, H, stage);
    OFRI_PH(1);
    if constexpr (PRECISE)
      hs_tma_tile_precise<T, R, NRG>(tl, W, H, uo, vo, stage, xbuf, has_next, nx, &mU, &mV, &mA, &mB, &mC, bar, alpha2);
    else
      hs_tma_tile<T, R, NRG>(tl, W, H, uo, vo, stage, xbuf, has_next, nx, &mU, &mV, &mA, &mB, &mC, bar);
    tl = nx;
  }
  OFRI_PH_FLUSH;
}


template <int T, int R, int NRG, bool PRECISE>
static bool launch_cfg(const Img& ui, const Img& vi, const Img& uo, const Img& vo, const Img& fx, const Img& fy,
                       const Img& ft, float alpha2, int num_sms, cudaStream_t s, const HsTileRows* sub) {
  using C = TmCfg<T, R, NRG>;
  if (ui.W <= C::HX + 1 || ui.H <= T + 1) return false;   // the ghost frame mirrors HX columns / T rows of real cells
  CUtensorMap mU, mV, mA, mB, mC;
  if (!make_map(&mU, ui, C::SH) || !make_map(&mV, vi, C::SH) || !make_map(&mA, fx, C::SH) ||
      !make_map(&mB, fy, C::SH) || !make_map(&mC, ft, C::SH))
    return false;
  auto kern = hs_tma_kernel<T, R, NRG, PRECISE>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (C::CTAS > 1)   // two CTAs per SM only fit with the largest shared-memory carve-out
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  const int tiles_x = (ui.W + C::TW - 1) / C::TW, tiles_y_all = (ui.H + C::TH - 1) / C::TH;
  int tiles_y = tiles_y_all;
  int4 rows = make_int4(tiles_y_all, 0, 0, 0);
  if (sub) {   // split launch: the tile rows lying entirely inside output rows [row_lo, row_hi), or all the others
    int m_lo = (sub->row_lo + C::TH - 1) / C::TH, m_hi = sub->row_hi / C::TH;
    if (m_lo < 0) m_lo = 0;
    if (m_hi > tiles_y_all) m_hi = tiles_y_all;
    if (m_hi < m_lo) m_hi = m_lo;
    if (sub->inside) {
      tiles_y = m_hi - m_lo;
      rows = make_int4(tiles_y, m_lo, 0, 0);
    } else {
      tiles_y = m_lo + (tiles_y_all - m_hi);
      rows = make_int4(m_lo, 0, m_hi - m_lo, 0);
    }
    if (tiles_y == 0) return true;
  }
  const long ntiles = (long)tiles_x * tiles_y * ui.batch;
  if (ntiles > 0x7fffffffL) return false;
  int avail = (num_sms - (sub ? sub->reserve_sms : 0)) * C::CTAS;
  if (avail < 1) avail = 1;
  const int grid = (int)(ntiles < avail ? ntiles : avail);
  kern<<<grid, C::NT, C::SMEM_BYTES, s>>>(mU, mV, mA, mB, mC, uo, vo, ui.W, ui.H, tiles_x, tiles_y, (int)ntiles,
                                          alpha2, rows);
  return true;
}

template <int T>
static bool launch_T(int variant, bool precise, const Img& ui, const Img& vi, const Img& uo, const Img& vo,
                     const Img& fx, const Img& fy, const Img& ft, float alpha2, int num_sms, cudaStream_t s,
                     const HsTileRows* sub) {
  (void)variant;
  if (precise)   // doubles in registers: 4-row strips, 34 x 128 tile, 256 threads
    return launch_cfg<T, 4, 8, true>(ui, vi, uo, vo, fx, fy, ft, alpha2, num_sms, s, sub);
  return launch_cfg<T, 8, 8, false>(ui, vi, uo, vo, fx, fy, ft, alpha2, num_sms, s, sub);     // 66 x 128, 256 threads
}

// T in {4, 6, 8} (precise arithmetic: also 2, 3).  Returns false if this kernel cannot run (other T, no driver entry point, map encoding failed): the
// caller then uses the non-persistent kernels.
bool launch_hs_tma(int T, int variant, bool precise, const Img& ui, const Img& vi, const Img& uo, const Img& vo,
                   const Img& fx, const Img& fy, const Img& ft, float alpha2, cudaStream_t s, const HsTileRows* sub) {
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  switch (T) {
    // T = 2, 3: precise arithmetic only (fewer stale border cells per tile; the fast kernel would be HBM-bound there)
    case 2: return precise ? launch_cfg<2, 4, 8, true>(ui, vi, uo, vo, fx, fy, ft, alpha2, num_sms, s, sub) : false;
    case 3: return precise ? launch_cfg<3, 4, 8, true>(ui, vi, uo, vo, fx, fy, ft, alpha2, num_sms, s, sub) : false;
    case 4: return launch_T<4>(variant, precise, ui, vi, uo, vo, fx, fy, ft, alpha2, num_sms, s, sub);
    case 6: return launch_T<6>(variant, precise, ui, vi, uo, vo, fx, fy, ft, alpha2, num_sms, s, sub);
    case 8: return launch_T<8>(variant, precise, ui, vi, uo, vo, fx, fy, ft, alpha2, num_sms, s, sub);
    default: return false;
  }
}

// phase-timing table of this translation unit (all zero unless built with -DOFRI_PHASE_TIMING): cycles of thread 0 of
// every CTA in [0] mbarrier wait, [1] ghost fix-up, [2] staging -> registers, [3] sweeps, [4] stores, [5] tile decode
void hs_tma_phase_read(unsigned long long* out) { OFRI_PH_READ(out); }

}  // namespace ofri
