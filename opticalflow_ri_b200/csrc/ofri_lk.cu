// ofri_lk.cu -- the dense Lucas-Kanade adapter of the reference as an sm_100a kernel.
//
// Replaces denseLucasKanade_PyCl.compute (denseLucasKanade_PyCL.py:113-169) and its OpenCL kernel lkDense
// (pyrlkDenseLargeW.cl:305-669, the non-CPU branch).  Per pixel: the structure tensor of a (2 hw + 1)^2 window of
// frame 1 (Scharr derivatives of the clamped image), then up to n_iters Gauss-Newton steps, each sampling frame 2
// bilinearly at the displaced window, until both components of the step fall below 0.01 px.
//
// Work split (the reference's, which fixes the order of every floating-point sum): 64 work-items per pixel, work-item
// (xid, yid) owns the 4 x 4 grid of window samples (xid + 8 tx, yid + 8 ty) with 0/1 weights that cut the 32 x 32
// grid down to the window; per-work-item sums accumulate in that order by FMA, the 64 partial sums are added by the
// fixed tree s[t] += s[t + 32], + 16, + 8, + 4, (s0 + s1) + (s2 + s3).  Here: one 64-thread CTA per pixel (persistent
// grid, runs of 8 neighbouring pixels per CTA), the tree's first level through shared memory (warp 1 -> warp 0), the
// rest by warp shuffles in the same operand order.  Frame 1's patch (and its Scharr derivatives) of a run of 8 pixels
// and, per Gauss-Newton step, the 36 x 36 window of frame 2 that the step's 1024 bilinear samples fall into are staged
// in shared memory with the clamp-to-edge rule applied while loading (row-coalesced reads instead of four scattered
// cache lines per warp and tap).  The three device-defined operations of the OpenCL original are fixed as in
// oracle/ofri_lk_oracle.c (see its header): image sampler = OpenCL-specification bilinear filter in full float32 with
// clamp-to-edge addressing, `mad` = fused multiply-add, IEEE division -- so the result is bit-identical to that oracle.
#include <climits>

#include "ofri_internal.h"
#include "ofri_pixel.cuh"

namespace ofri {

namespace {

struct LkWeights { float wx[8][4], wy[8][4]; };

// read_imagef(.., unnormalised coordinates | clamp to edge | linear filter, (x, y)) (pyrlkDenseLargeW.cl:236) out of the
// CTA's staged window of frame 2: win[yy][xx] holds J at (by + yy, bx + xx) with the clamp-to-edge rule already applied.
// The filter separates per axis: texel index and weight pair depend on x (resp. y) alone, so a work-item prepares them
// once for each of its 4 sample columns and 4 sample rows (LkAxis) and combines them per sample -- the same float32
// operations on the same operands as a direct image read, hence the same bits.  The index clamp into the window is a
// memory-safety guard only (the window is one texel wider on every side than the positions can reach).
// pitch 40 (72 for the 41 columns of a run's frame-1 patch): = 8 mod 32, so the 8 x 4 sample lattice of a warp hits 32
// different banks
constexpr int LK_WIN = 36, LK_WPITCH = 40, LK_PPITCH = 72;
struct LkAxis { float a, oa; int off; };             // weight of the upper texel, of the lower texel, window offset
__device__ __forceinline__ LkAxis lk_axis(float x, int base, int scale) {
  const float fx = fsub(x, 0.5f);
  const float ix = floorf(fx);
  LkAxis r;
  r.a = fsub(fx, ix);
  r.oa = fsub(1.0f, r.a);
  r.off = min(max((int)ix - base, 0), LK_WIN - 2) * scale;
  return r;
}
__device__ __forceinline__ float lk_sample(const float* __restrict__ win, const LkAxis& X, const LkAxis& Y) {
  const float* c = win + Y.off + X.off;
  const float t00 = c[0], t10 = c[1], t01 = c[LK_WPITCH], t11 = c[LK_WPITCH + 1];
  float r = fmul(fmul(X.oa, Y.oa), t00);
  r = fadd(r, fmul(fmul(X.a, Y.oa), t10));
  r = fadd(r, fmul(fmul(X.oa, Y.a), t01));
  r = fadd(r, fmul(fmul(X.a, Y.a), t11));
  return r;
}

// Stage an N x NC window of `im` whose top-left texel is (y0, x0) into shared memory (row pitch `pitch_s`), clamp-to-edge
// applied while loading.  The two warps take alternate rows; a lane loads column `lane` and, for the first N - 32 lanes,
// column 32 + lane: one clamped column offset per lane for the whole window, one clamped row pointer per row.
template <int N, int NC = N>
__device__ __forceinline__ void lk_stage_window(const float* __restrict__ im, int H, int W, long pitch, int y0, int x0,
                                                float* __restrict__ dst, int pitch_s, int tid) {
  static_assert(NC > 32 && NC <= 64, "two column loads per lane");
  const int lane = tid & 31, half = tid >> 5;
  const bool second = lane < NC - 32;
  if (y0 >= 0 && x0 >= 0 && y0 + N <= H && x0 + NC <= W) {      // CTA-uniform: window inside the image, no clamping; the
    const float* src = im + (long)(y0 + half) * pitch + x0 + lane;   // row pointers advance by two rows per trip
    float* d = dst + half * pitch_s + lane;
    const long step = 2 * pitch;
    const int dstep = 2 * pitch_s;
#pragma unroll 3
    for (int yy = half; yy < N; yy += 2, src += step, d += dstep) {
      d[0] = __ldg(src);
      if (second) d[32] = __ldg(src + 32);
    }
    return;
  }
  const int c0 = min(max(x0 + lane, 0), W - 1), c1 = min(max(x0 + 32 + lane, 0), W - 1);
#pragma unroll 2
  for (int yy = half; yy < N; yy += 2) {
    const float* row = im + (long)min(max(y0 + yy, 0), H - 1) * pitch;
    dst[yy * pitch_s + lane] = __ldg(row + c0);
    if (second) dst[yy * pitch_s + 32 + lane] = __ldg(row + c1);
  }
}

// The reference's work-group sum of N values per work-item at once (pyrlkDenseLargeW.cl:113-155), 64 threads.
// sm: N x 33 floats.  Every thread returns the totals.
template <int N>
__device__ __forceinline__ void lk_group_sum(float (&val)[N], float* sm, int tid) {
  if (tid >= 32) {
#pragma unroll
    for (int n = 0; n < N; ++n) sm[n * 33 + tid - 32] = val[n];
  }
  __syncthreads();
  if (tid < 32) {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      float s = fadd(val[n], sm[n * 33 + tid]);
      s = fadd(s, __shfl_down_sync(0xffffffffu, s, 16));
      s = fadd(s, __shfl_down_sync(0xffffffffu, s, 8));
      s = fadd(s, __shfl_down_sync(0xffffffffu, s, 4));
      const float s1 = __shfl_sync(0xffffffffu, s, 1), s2 = __shfl_sync(0xffffffffu, s, 2),
                  s3 = __shfl_sync(0xffffffffu, s, 3);
      if (tid == 0) sm[n * 33 + 32] = fadd(fadd(s, s1), fadd(s2, s3));
    }
  }
  __syncthreads();
#pragma unroll
  for (int n = 0; n < N; ++n) val[n] = sm[n * 33 + 32];
}

__global__ void __launch_bounds__(64) lk_dense_kernel(Img I, Img J, Img U, Img V, int iters, float hw, LkWeights wt) {
  __shared__ float patch[34 * LK_PPITCH];             // frame 1 around the run: rows py-1 .. py+32, 41 columns
  __shared__ float ddx[32 * LK_WPITCH], ddy[32 * LK_WPITCH];   // its (unweighted) Scharr derivatives, 32 x 39
  __shared__ float jwin[LK_WIN * LK_WPITCH];
  __shared__ float red[3 * 33];
  const int tid = threadIdx.x, xid = tid & 7, yid = tid >> 3;
  const int H = I.H, W = I.W;
  const float* __restrict__ pI = I.at(blockIdx.y);
  const float* __restrict__ pJ = J.at(blockIdx.y);
  float* pU = U.at(blockIdx.y);
  float* pV = V.at(blockIdx.y);
  float w[4][4];
#pragma unroll
  for (int ty = 0; ty < 4; ++ty)
#pragma unroll
    for (int tx = 0; tx < 4; ++tx) w[ty][tx] = fmul(wt.wy[yid][ty], wt.wx[xid][tx]);
  const int ihw = (int)hw;
  // Pixels in RUNS of 8 neighbours of one image row per CTA (round-robin over the runs).  The windows of a run overlap in
  // all but 7 columns, so frame 1 is staged once per run (34 x 41 texels) and its Scharr derivatives are formed once per
  // run (32 x 39) into shared memory; a pixel then only reads its 16 samples per work-item and applies its 0/1 weights
  // -- the same products in the same order as a per-pixel evaluation.
  const int runs_per_row = (W + 7) / 8;
  const long nruns = (long)H * runs_per_row;
  for (long run = blockIdx.x; run < nruns; run += gridDim.x) {
    const int i = (int)(run / runs_per_row), j0 = 8 * (int)(run - (long)i * runs_per_row);
    const int py = i - ihw;
    __syncthreads();                                   // the previous run's patch / derivatives are no longer read
    lk_stage_window<34, 41>(pI, H, W, I.pitch, py - 1, j0 - ihw - 1, patch, LK_PPITCH, tid);
    __syncthreads();
    for (int e = tid; e < 32 * 39; e += 64) {
      const int y = e / 39, x = e - y * 39;
      constexpr int P = LK_PPITCH;
      const float* c = patch + (y + 1) * P + (x + 1);
      const float sx = fsub(fsub(fadd(c[-P + 1], c[P + 1]), c[-P - 1]), c[P - 1]);
      const float sy = fsub(fsub(fadd(c[P - 1], c[P + 1]), c[-P - 1]), c[-P + 1]);
      ddx[y * LK_WPITCH + x] = __fmaf_rn(sx, 3.0f, fmul(fsub(c[1], c[-1]), 10.0f));
      ddy[y * LK_WPITCH + x] = __fmaf_rn(sy, 3.0f, fmul(fsub(c[P], c[-P]), 10.0f));
    }
    __syncthreads();
   for (int kp = 0; kp < 8 && j0 + kp < W; ++kp) {
    const int j = j0 + kp;
    float pv[4][4], dxs[4][4], dys[4][4];
    float acc[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int ty = 0; ty < 4; ++ty)
#pragma unroll
      for (int tx = 0; tx < 4; ++tx) {
        const int y = ty * 8 + yid, x = tx * 8 + xid + kp;
        const float dx = fmul(ddx[y * LK_WPITCH + x], w[ty][tx]);
        const float dy = fmul(ddy[y * LK_WPITCH + x], w[ty][tx]);
        pv[ty][tx] = patch[(y + 1) * LK_PPITCH + x + 1];
        dxs[ty][tx] = dx;
        dys[ty][tx] = dy;
        acc[0] = __fmaf_rn(dx, dx, acc[0]);
        acc[1] = __fmaf_rn(dx, dy, acc[1]);
        acc[2] = __fmaf_rn(dy, dy, acc[2]);
      }
    lk_group_sum<3>(acc, red, tid);
    float A11 = acc[0], A12 = acc[1], A22 = acc[2];
    const float D = __fmaf_rn(A11, A22, -fmul(A12, A12));
    if (D < 1.192092896e-07f) continue;                // CTA-uniform: the pixel's flow stays as it came in
    A11 = __fdiv_rn(A11, D);
    A12 = __fdiv_rn(A12, D);
    A22 = __fdiv_rn(A22, D);
    float ppx = fsub(fadd((float)j, pU[(long)i * U.pitch + j]), hw);
    float ppy = fsub(fadd((float)i, pV[(long)i * V.pitch + j]), hw);
    float lx[4], ly[4];
    lx[0] = fadd(ppx, fadd((float)xid, 0.5f));
    ly[0] = fadd(ppy, fadd((float)yid, 0.5f));
#pragma unroll
    for (int t = 1; t < 4; ++t) {
      lx[t] = fadd(lx[t - 1], 8.0f);
      ly[t] = fadd(ly[t - 1], 8.0f);
    }
    int bx = INT_MIN, by = INT_MIN;                      // window of frame 2 currently staged (none)
    for (int k = 0; k < iters; ++k) {
      if (ppx < -hw || ppx >= (float)W || ppy < -hw || ppy >= (float)H) break;
      // stage the window of frame 2 this step samples (all 64 x 16 positions lie in [ppx, ppx + 32] x [ppy, ppy + 32];
      // the previous step's reads ended before the barriers of its work-group sum) -- unless the step moved the window
      // origin by less than a texel and the staged window still covers it (CTA-uniform)
      const int nbx = (int)floorf(ppx) - 1, nby = (int)floorf(ppy) - 1;
      if (nbx != bx || nby != by) {
        bx = nbx;
        by = nby;
        lk_stage_window<LK_WIN>(pJ, H, W, J.pitch, by, bx, jwin, LK_WPITCH, tid);
        __syncthreads();
      }
      float b[2] = {0.0f, 0.0f};
      LkAxis ax[4], ay[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        ax[t] = lk_axis(lx[t], bx, 1);
        ay[t] = lk_axis(ly[t], by, LK_WPITCH);
      }
#pragma unroll
      for (int ty = 0; ty < 4; ++ty)
#pragma unroll
        for (int tx = 0; tx < 4; ++tx) {
          const float diff = fmul(fsub(lk_sample(jwin, ax[tx], ay[ty]), pv[ty][tx]), w[ty][tx]);
          b[0] = __fmaf_rn(diff, dxs[ty][tx], b[0]);
          b[1] = __fmaf_rn(diff, dys[ty][tx], b[1]);
        }
      lk_group_sum<2>(b, red, tid);
      const float ddx = fmul(__fmaf_rn(A12, b[1], -fmul(A22, b[0])), 32.0f);
      const float ddy = fmul(__fmaf_rn(A12, b[0], -fmul(A11, b[1])), 32.0f);
      ppx = fadd(ppx, ddx);
      ppy = fadd(ppy, ddy);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        lx[t] = fadd(lx[t], ddx);
        ly[t] = fadd(ly[t], ddy);
      }
      if (fabsf(ddx) < 0.01f && fabsf(ddy) < 0.01f) break;
    }
    if (tid == 0) {
      pU[(long)i * U.pitch + j] = fsub(fadd(ppx, hw), (float)j);
      pV[(long)i * V.pitch + j] = fsub(fadd(ppy, hw), (float)i);
    }
   }
  }
}

// per-axis 0/1 weights of the four grid columns (rows) of work-item `id` (pyrlkDenseLargeW.cl:339-372): `win` = window
// extent, lo / hi = the asymmetric-window switches of the axis (left / right, top / bottom)
void axis_weights(int id, int win, int lo, int hi, float* w) {
  w[0] = 1.0f;
  if (win >= 16) {
    w[1] = id == 0 ? (float)(1 - lo) : 1.0f;
    w[2] = (16 + id < win - hi) ? 1.0f : 0.0f;
    w[3] = (24 + id < win - hi) ? 1.0f : 0.0f;
  } else {
    w[1] = id == 0 ? (float)(1 - lo) : ((8 + id < win - hi) ? 1.0f : 0.0f);
    w[2] = 0.0f;
    w[3] = 0.0f;
  }
}

}  // namespace

// u_io / v_io: initial flow in, refined flow out (every CTA touches only its own pixel's entry)
int launch_lk(const Img& im1, const Img& im2, const Img& u_io, const Img& v_io, const ofri_lk_params* lp, cudaStream_t s,
              LaunchCounter& lc) {
  const int win = 2 * lp->half_window + 1;
  LkWeights wt;
  for (int id = 0; id < 8; ++id) {
    axis_weights(id, win, lp->asym[0], lp->asym[1], wt.wx[id]);
    axis_weights(id, win, lp->asym[2], lp->asym[3], wt.wy[id]);
  }
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const long want = (long)num_sms * 8;                  // 8 resident 64-thread CTAs per SM (128 registers per thread; 12 CTAs at 80
                                                        // registers measured 4 % slower: the kernel is issue-bound)
  const long runs = (long)im1.H * ((im1.W + 7) / 8);
  dim3 grid((unsigned)(runs < want ? runs : want), (unsigned)im1.batch);
  lk_dense_kernel<<<grid, 64, 0, s>>>(im1, im2, u_io, v_io, lp->n_iters, (float)((win - 1) >> 1), wt);
  ++lc.n;
  return cudaGetLastError() == cudaSuccess ? OFRI_OK : OFRI_ERR_CUDA;
}

}  // namespace ofri
