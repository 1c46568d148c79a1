// ofri_ls_common.cuh -- device helpers shared by the Liu-Shen sweep kernels (ofri_ls.cu, ofri_ls_tma.cu): window rows
// with the 'nearest' rule in x, and the per-row update (one expression tree for every kernel -> bit-identical results).
// Reference: PhysicsBasedOpticalFlowLiuShen.py:141-156, 75-80.
#pragma once
#include "ofri_internal.h"
#include "ofri_pixel.cuh"

namespace ofri {

// ---- stopping rule (LS:141) ---------------------------------------------------------------------------------------
__device__ __forceinline__ double ls_total_error(const double* e, double npix) {
  // two separate square roots, then the sum (LS:79); f32 rounding of each norm as numba's np.linalg.norm returns f32
  return ((double)(float)sqrt(e[0]) + (double)(float)sqrt(e[1])) / npix;
}
// true iff the pair must NOT run sweep k (k >= 1): some earlier sweep already met the tolerance.  Launches are
// issued in order and a stopped pair writes nothing (its sums stay 0 -> error 0 -> "stopped"), so it suffices to
// look back over the sweeps of the previous launch (`lookback` = the fuse factor): this also catches a trip in the
// MIDDLE of a fused block whose later sweeps went back above the tolerance.
__device__ __forceinline__ bool ls_stopped_before(const double* errs_pair, int k, double tol, double npix,
                                                  int lookback) {
  int first = k - lookback;
  if (first < 0) first = 0;
  for (int i = first; i < k; ++i) {
    double te = ls_total_error(errs_pair + 2 * i, npix);
    if (!(te > tol)) return true;
  }
  return false;
}

struct LsEdge {
  bool left_edge;
  int right_j, top_j, bot_j;
};

// one shared row of this thread's strip: columns sx-1 .. sx+4 of u and v; EDGE: 'nearest' clamp in x
template <bool EDGE>
__device__ __forceinline__ void ls_row6(const float* __restrict__ pu, const float* __restrict__ pv, const LsEdge& eg,
                                        float (&du)[6], float (&dv)[6]) {
  float4 qu = *reinterpret_cast<const float4*>(pu);
  float4 qv = *reinterpret_cast<const float4*>(pv);
  du[1] = qu.x; du[2] = qu.y; du[3] = qu.z; du[4] = qu.w;
  dv[1] = qv.x; dv[2] = qv.y; dv[3] = qv.z; dv[4] = qv.w;
  du[0] = __shfl_up_sync(0xffffffffu, qu.w, 1);
  du[5] = __shfl_down_sync(0xffffffffu, qu.x, 1);
  dv[0] = __shfl_up_sync(0xffffffffu, qv.w, 1);
  dv[5] = __shfl_down_sync(0xffffffffu, qv.x, 1);
  if (EDGE) {   // the left neighbour of column 0 is column 0; the right neighbour of column W-1 is column W-1
    if (eg.left_edge) { du[0] = du[1]; dv[0] = dv[1]; }
    if (eg.right_j == 0) { du[2] = du[1]; dv[2] = dv[1]; }
    if (eg.right_j == 1) { du[3] = du[2]; dv[3] = dv[2]; }
    if (eg.right_j == 2) { du[4] = du[3]; dv[4] = dv[3]; }
    if (eg.right_j == 3) { du[5] = du[4]; dv[5] = dv[4]; }
  }
}

// 4 pixels of one strip row.  (uu, um, ud) = window rows above / at / below, already clamped in x (and in y by the
// caller's choice of rows).  ztop / zbot: the row above / below lies outside the image (zero padding of H8).
// coefficient registers of one strip row (4 cells x 8 planes: IIx, IIy, II, Ixt, Iyt, B11, B12, B22)
struct LsCoefRow { float c[8][4]; };
template <bool EDGE>
__device__ __forceinline__ void ls_row_update_regs(const float (&uu)[6], const float (&um)[6], const float (&ud)[6],
                                                   const float (&vu)[6], const float (&vm)[6], const float (&vd)[6],
                                                   const LsCoefRow& k, float hpar, bool ztop, bool zbot,
                                                   const LsEdge& eg, float (&ou)[4], float (&ov)[4]) {
  const float (&c0)[4] = k.c[0], (&c1)[4] = k.c[1], (&c2)[4] = k.c[2], (&c3)[4] = k.c[3];
  const float (&c4)[4] = k.c[4], (&c5)[4] = k.c[5], (&c6)[4] = k.c[6], (&c7)[4] = k.c[7];
  // zero-padded column sums for H8 (interior: identical to the clamped values) and the vertical differences (clamped)
  // that 2 Dr and 4 Mx are made of -- one each per column, shared by the pixels left and right of it
  float zsu[6], zsv[6], zmu[6], zmv[6], vdu[6], vdv[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    float a = uu[c], d = ud[c], e = vu[c], f = vd[c];
    vdu[c] = fsub(d, a);
    vdv[c] = fsub(f, e);
    if (EDGE) {
      if (ztop) { a = 0.0f; e = 0.0f; }
      if (zbot) { d = 0.0f; f = 0.0f; }
    }
    zsu[c] = fadd(a, d);
    zsv[c] = fadd(e, f);
    zmu[c] = um[c];
    zmv[c] = vm[c];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float vlu = zsu[j], vru = zsu[j + 2], wu_ = zmu[j], eu_ = zmu[j + 2];
    float vlv = zsv[j], vrv = zsv[j + 2], wv_ = zmv[j], ev_ = zmv[j + 2];
    if (EDGE) {
      if (eg.left_edge && j == 0) { vlu = 0.0f; vlv = 0.0f; wu_ = 0.0f; wv_ = 0.0f; }
      if (eg.right_j == j) { vru = 0.0f; vrv = 0.0f; eu_ = 0.0f; ev_ = 0.0f; }
    }
    float h8u = ls_h8_cols(vlu, zsu[j + 1], vru, wu_, eu_);
    float h8v = ls_h8_cols(vlv, zsv[j + 1], vrv, wv_, ev_);
    LsSt t;                                   // the same expressions as ls_update2 forms from the 8 neighbours
    t.dr_u = vdu[j + 1];
    t.dr_v = vdv[j + 1];
    t.dc_u = fsub(um[j + 2], um[j]);
    t.dc_v = fsub(vm[j + 2], vm[j]);
    t.fr_u = fadd(uu[j + 1], ud[j + 1]);
    t.fc_v = fadd(vm[j], vm[j + 2]);
    t.mx_u = fsub(vdu[j + 2], vdu[j]);
    t.mx_v = fsub(vdv[j + 2], vdv[j]);
    LsCoef c;
    c.IIx = c0[j]; c.IIy = c1[j]; c.II = c2[j]; c.Ixt = c3[j]; c.Iyt = c4[j]; c.B11 = c5[j]; c.B12 = c6[j]; c.B22 = c7[j];
    ls_update3(t, h8u, h8v, c, hpar, &ou[j], &ov[j]);
  }
}

// same with the coefficients staged in shared memory (plane-major, `plane` floats apart)
template <bool EDGE>
__device__ __forceinline__ void ls_row_update(const float (&uu)[6], const float (&um)[6], const float (&ud)[6],
                                              const float (&vu)[6], const float (&vm)[6], const float (&vd)[6],
                                              const float* __restrict__ sC, int plane, int so, float hpar, bool ztop,
                                              bool zbot, const LsEdge& eg, float (&ou)[4], float (&ov)[4]) {
  LsCoefRow k;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 q = *reinterpret_cast<const float4*>(sC + c * plane + so);
    k.c[c][0] = q.x; k.c[c][1] = q.y; k.c[c][2] = q.z; k.c[c][3] = q.w;
  }
  ls_row_update_regs<EDGE>(uu, um, ud, vu, vm, vd, k, hpar, ztop, zbot, eg, ou, ov);
}

}  // namespace ofri
