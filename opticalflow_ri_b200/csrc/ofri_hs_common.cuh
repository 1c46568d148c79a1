// ofri_hs_common.cuh -- device helpers shared by the Horn-Schunck sweep kernels (ofri_hs.cu, ofri_hs_tma.cu): the
// 6-wide window rows, the per-row update (one expression tree for every kernel, so all of them are bit-identical)
// and the boundary bookkeeping of EDGE tiles.  Reference: HornSchunck.py:52-71.
#pragma once
#include "ofri_internal.h"
#include "ofri_pixel.cuh"

namespace ofri {

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

struct HsEdge {      // per-thread boundary facts (EDGE tiles only)
  bool left_edge;    // this strip starts at global column 0
  int right_j;       // strip column j sitting on global column W-1 (else out of [0,4))
  int top_j, bot_j;  // strip row j sitting on global row 0 / H-1 (else out of [0,R))
};

// coefficient registers of one strip row: fast path (a, b, c); precise (fx, fy, ft, den, 1/den)
template <bool PRECISE>
struct HsCoef {
  float c0[4], c1[4], c2[4], c3[PRECISE ? 4 : 1], c4[PRECISE ? 4 : 1];
};

template <bool PRECISE>
__device__ __forceinline__ void hs_row_update(const float (&uu)[6], const float (&um)[6], const float (&ud)[6],
                                              const float (&vu)[6], const float (&vm)[6], const float (&vd)[6],
                                              const HsCoef<PRECISE>& k, float (&ou)[4], float (&ov)[4]) {
  if constexpr (!PRECISE) {
    float vsu[6], vsv[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      vsu[c] = fadd(uu[c], ud[c]);
      vsv[c] = fadd(vu[c], vd[c]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float ua = hs_avg_cols(vsu[j], vsu[j + 1], vsu[j + 2], um[j], um[j + 2]);
      float va = hs_avg_cols(vsv[j], vsv[j + 1], vsv[j + 2], vm[j], vm[j + 2]);
      hs_update_n(ua, va, k.c0[j], k.c1[j], k.c2[j], &ou[j], &ov[j]);
    }
  } else {
    double vsu[6], vsv[6], mu[6], mv[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      vsu[c] = dadd((double)uu[c], (double)ud[c]);
      vsv[c] = dadd((double)vu[c], (double)vd[c]);
      mu[c] = (double)um[c];
      mv[c] = (double)vm[c];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float ua = hs_avg_cols_precise(vsu[j], vsu[j + 1], vsu[j + 2], mu[j], mu[j + 2]);
      float va = hs_avg_cols_precise(vsv[j], vsv[j + 1], vsv[j + 2], mv[j], mv[j + 2]);
      hs_update_precise(ua, va, k.c0[j], k.c1[j], k.c2[j], k.c3[j], k.c4[j], &ou[j], &ov[j]);
    }
  }
}

template <bool EDGE>
__device__ __forceinline__ void hs_mirror_x(const HsEdge& eg, float (&du)[6], float (&dv)[6]) {
  if (EDGE) {
    if (eg.left_edge) { du[0] = du[2]; dv[0] = dv[2]; }
    if (eg.right_j == 0) { du[2] = du[0]; dv[2] = dv[0]; }
    if (eg.right_j == 1) { du[3] = du[1]; dv[3] = dv[1]; }
    if (eg.right_j == 2) { du[4] = du[2]; dv[4] = dv[2]; }
    if (eg.right_j == 3) { du[5] = du[3]; dv[5] = dv[3]; }
  }
}
template <bool EDGE>
__device__ __forceinline__ void hs_row6_vals(const float (&ru)[4], const float (&rv)[4], const HsEdge& eg,
                                             float (&du)[6], float (&dv)[6]) {
  du[1] = ru[0]; du[2] = ru[1]; du[3] = ru[2]; du[4] = ru[3];
  dv[1] = rv[0]; dv[2] = rv[1]; dv[3] = rv[2]; dv[4] = rv[3];
  du[0] = __shfl_up_sync(0xffffffffu, ru[3], 1);
  du[5] = __shfl_down_sync(0xffffffffu, ru[0], 1);
  dv[0] = __shfl_up_sync(0xffffffffu, rv[3], 1);
  dv[5] = __shfl_down_sync(0xffffffffu, rv[0], 1);
  hs_mirror_x<EDGE>(eg, du, dv);
}
template <bool EDGE>
__device__ __forceinline__ void hs_row6_smem(const float* __restrict__ pu, const float* __restrict__ pv,
                                             const HsEdge& eg, float (&du)[6], float (&dv)[6]) {
  float4 qu = *reinterpret_cast<const float4*>(pu);
  float4 qv = *reinterpret_cast<const float4*>(pv);
  const float ru[4] = {qu.x, qu.y, qu.z, qu.w}, rv[4] = {qv.x, qv.y, qv.z, qv.w};
  hs_row6_vals<EDGE>(ru, rv, eg, du, dv);
}

}  // namespace ofri
