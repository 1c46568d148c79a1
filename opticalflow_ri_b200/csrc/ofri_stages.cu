// ofri_stages.cu -- per-level stage kernels (sm_100a): Gaussian pre-filter, Pillow-bicubic down-sample,
// not-a-knot spline up-sample, symmetric bilinear warp, and small element-wise helpers.
// Each of these runs once per pyramid level (a few % of the bytes of the iteration kernels); they are written
// for coalesced access (threads along x) and exact reproduction of the reference arithmetic (ofri_pixel.cuh).
#include "ofri_internal.h"
#include "ofri_pixel.cuh"
#include "ofri_spline.cuh"

namespace ofri {

static inline dim3 grid2d(int W, int H, int batch, dim3 b) {
  return dim3((W + b.x - 1) / b.x, (H + b.y - 1) / b.y, batch);
}

// ---------------------------------------------------------------------------------------------------------------
// Gaussian pre-filter: rows then columns (gaussian_filter.py:54-85)
// ---------------------------------------------------------------------------------------------------------------
// Fused separable filter for small kernels: a CTA stages a (TY + 2h) x (TX + 2h) source tile -- already
// remapped through the reference's padding rule -- in shared memory, row-filters TY+2h rows into a second
// shared tile, then column-filters.  Rows that the column pass needs above/below the tile are row-filtered
// redundantly (h rows each side), so the intermediate image never goes to HBM: 4 B read + 4 B written per pixel.
template <int K, int TX, int TY>
__global__ void __launch_bounds__(256) gauss_fused_kernel(Img in, Img out, GaussTaps taps) {
  constexpr int h = K / 2;
  constexpr int SW = TX + 2 * h, SH = TY + 2 * h, NT = 256;
  __shared__ float src[SH][SW + 1];
  __shared__ float rowf[SH][TX + 1];
  const int b = blockIdx.z;
  const float* ip = in.p + (long)b * in.stride;
  float* op = out.p + (long)b * out.stride;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int tid = threadIdx.x;
  // stage: padded coordinates (py, px) of the tile origin are (y0, x0) .. ; padded index p maps to source
  // index gauss_src_index(p, n, h).  Rows/cols beyond the image are clamped (never used by valid outputs).
  for (int i = tid; i < SH * SW; i += NT) {
    int sy = i / SW, sx = i - sy * SW;
    int py = y0 + sy, px = x0 + sx;                       // padded-line positions
    int gy = gauss_src_index(py < in.H + 2 * h ? py : in.H + 2 * h - 1, in.H, h);
    int gx = gauss_src_index(px < in.W + 2 * h ? px : in.W + 2 * h - 1, in.W, h);
    src[sy][sx] = ip[(long)gy * in.pitch + gx];
  }
  __syncthreads();
  // row pass: rowf[sy][x] = sum_j P[x + 2h - j] k[j] for the SH staged rows
  for (int i = tid; i < SH * TX; i += NT) {
    int sy = i / TX, x = i - sy * TX;
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < K; ++j) acc = fadd(acc, fmul(src[sy][x + 2 * h - j], taps.k[j]));
    rowf[sy][x] = acc;
  }
  __syncthreads();
  for (int i = tid; i < TY * TX; i += NT) {
    int ty = i / TX, tx = i - ty * TX;
    const int x = x0 + tx, y = y0 + ty;
    if (x < in.W && y < in.H) {
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < K; ++j) acc = fadd(acc, fmul(rowf[ty + 2 * h - j][tx], taps.k[j]));
      op[(long)y * out.pitch + x] = acc;
    }
  }
}
// NOTE on the fused kernel's column pass: the padded column of the ROW-FILTERED image is built from row-filtered
// rows with the same padding rule; since staging applied the rule to source rows before row-filtering, and the
// row filter acts independently on each row, rowf[sy] IS the row-filtered image row gauss_src_index(y0+sy).

// generic two-pass fallback for any K (used by the truncate variant, K up to OFRI_MAX_GAUSS_TAPS)
__global__ void gauss_rows_kernel(Img in, Img out, GaussTaps taps) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= in.W || y >= in.H) return;
  const float* row = in.p + (long)b * in.stride + (long)y * in.pitch;
  out.p[(long)b * out.stride + (long)y * out.pitch + x] = gauss_point(row, 1, x, in.W, taps, taps.K);
}
__global__ void gauss_cols_kernel(Img in, Img out, GaussTaps taps) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= in.W || y >= in.H) return;
  const float* col = in.p + (long)b * in.stride + x;
  out.p[(long)b * out.stride + (long)y * out.pitch + x] = gauss_point(col, in.pitch, y, in.H, taps, taps.K);
}

void launch_gauss(const Img& in, const Img& tmp, const Img& out, const GaussTaps& taps, cudaStream_t s,
                  LaunchCounter& lc) {
  if (taps.K == 3) {      // 64 x 32 output tile per 256-thread block: 9 % halo instead of 33 %
    gauss_fused_kernel<3, 64, 32><<<grid2d(in.W, in.H, in.batch, dim3(64, 32)), 256, 0, s>>>(in, out, taps);
    lc.n += 1;
  } else if (taps.K == 5) {
    gauss_fused_kernel<5, 64, 32><<<grid2d(in.W, in.H, in.batch, dim3(64, 32)), 256, 0, s>>>(in, out, taps);
    lc.n += 1;
  } else {
    dim3 b(32, 8);
    gauss_rows_kernel<<<grid2d(in.W, in.H, in.batch, b), b, 0, s>>>(in, tmp, taps);
    gauss_cols_kernel<<<grid2d(in.W, in.H, in.batch, b), b, 0, s>>>(tmp, out, taps);
    lc.n += 2;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Pillow BICUBIC down-sample: horizontal pass (f32 intermediate) then vertical pass
// ---------------------------------------------------------------------------------------------------------------
__global__ void resize_h_kernel(Img in, Img out, ResizeTaps t) {
  int ox = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (ox >= out.W || y >= out.H) return;
  const float* row = in.p + (long)b * in.stride + (long)y * in.pitch;
  out.p[(long)b * out.stride + (long)y * out.pitch + ox] =
      resample_point(row, 1, t.xmin[ox], t.cnt[ox], t.w + (long)ox * t.kmax);
}
// in_row0 / out_row0: global row of local row 0 of `in` / `out` (row bands; 0 for whole images).  Output rows whose
// taps are not all inside the band are written as 0 (they can only be ghost rows the driver never uses).
__global__ void resize_v_kernel(Img in, Img out, ResizeTaps t, int in_row0, int out_row0) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= out.W || oy >= out.H) return;
  const int og = oy + out_row0;
  const int start = t.xmin[og] - in_row0, cnt = t.cnt[og];
  float r = 0.0f;
  if (start >= 0 && start + cnt <= in.H)
    r = resample_point(in.p + (long)b * in.stride + x, in.pitch, start, cnt, t.w + (long)og * t.kmax);
  out.p[(long)b * out.stride + (long)oy * out.pitch + x] = r;
}
void launch_resize(const Img& in, const Img& tmp, const Img& out, const ResizeTaps& tx, const ResizeTaps& ty,
                   cudaStream_t s, LaunchCounter& lc, int in_row0, int out_row0) {
  dim3 b(32, 8);
  // tmp: in.H x out.W
  Img t = tmp;
  t.H = in.H;
  t.W = out.W;
  if (out.W != in.W) {
    resize_h_kernel<<<grid2d(t.W, t.H, in.batch, b), b, 0, s>>>(in, t, tx);
    lc.n += 1;
  } else {
    t = in;
  }
  if (out.H != in.H || in_row0 != 0 || out_row0 != 0) {
    resize_v_kernel<<<grid2d(out.W, out.H, in.batch, b), b, 0, s>>>(t, out, ty, in_row0, out_row0);
    lc.n += 1;
  } else {
    launch_copy(out, t, s, lc);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Not-a-knot cubic spline up-sample (RectBivariateSpline kx=ky=3, s=0 restated).  f64 throughout.
// ---------------------------------------------------------------------------------------------------------------
// One thread per line.  Element e of line l at base[l*line_stride + e*elem_stride].  Writes the second derivatives
// M (same indexing, f64).  Forward sweep keeps dp in M[e+1]; the backward sweep overwrites it in place.
template <typename TIn>
__global__ void spline_solve_kernel(const TIn* __restrict__ y, long y_elem, long y_line, long y_batch,
                                    double* __restrict__ M, long m_elem, long m_line, long m_batch, int n, int lines,
                                    SplineSys sys) {
  int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= lines) return;
  const TIn* yy = y + (long)blockIdx.z * y_batch + (long)l * y_line;
  double* mm = M + (long)blockIdx.z * m_batch + (long)l * m_line;
  const int m = n - 2;
  // The recurrences are sequential per line (one thread each): a launch is bound by the latency of one step.  After a
  // short head the rows of the system share one (den, cp) pair (SplineSys::conv), so the steady-state step keeps its
  // constants in registers, divides by the constant with a reciprocal + two FMA corrections (correctly rounded: the
  // same quotient as the IEEE division of the head, without its special-case call that stops the compiler from
  // moving loads), and fetches its inputs CH steps ahead.
  constexpr int CH = 8;
  double y0 = (double)yy[0], y1 = (double)yy[y_elem], y2 = (double)yy[2 * y_elem];
  double dp = ddiv(dmul(6.0, dadd(dsub(y0, dmul(2.0, y1)), y2)), sys.den[0]);
  mm[m_elem] = dp;
  y0 = y1;
  y1 = y2;
  const int conv = sys.conv < 1 ? 1 : sys.conv;
  int i = 1;
  for (; i < m && i < conv; ++i) {                   // head: tabulated constants, IEEE division
    y2 = (double)yy[(long)(i + 2) * y_elem];
    double rhs = dmul(6.0, dadd(dsub(y0, dmul(2.0, y1)), y2));
    dp = ddiv(dsub(rhs, dmul(sys.lo[i], dp)), sys.den[i]);
    mm[(long)(i + 1) * m_elem] = dp;
    y0 = y1;
    y1 = y2;
  }
  {
    const double den = sys.den_c, rcp = sys.rcp_c;
    // steady state (lo == 1): CH steps per trip; the loads of trip k+1 are issued before the arithmetic of trip k
    double yb[CH], yn[CH];
    if (i + CH <= m - 1) {
#pragma unroll
      for (int q = 0; q < CH; ++q) yb[q] = (double)yy[(long)(i + q + 2) * y_elem];
    }
    for (; i + CH <= m - 1; i += CH) {
      const bool more = i + 2 * CH <= m - 1;
      if (more) {
#pragma unroll
        for (int q = 0; q < CH; ++q) yn[q] = (double)yy[(long)(i + CH + q + 2) * y_elem];
      }
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        double rhs = dmul(6.0, dadd(dsub(y0, dmul(2.0, y1)), yb[q]));
        double t = dsub(rhs, dp);
        double qq = dmul(t, rcp);
        double e = fma(-den, qq, t);
        qq = fma(e, rcp, qq);
        e = fma(-den, qq, t);
        dp = fma(e, rcp, qq);
        mm[(long)(i + q + 1) * m_elem] = dp;
        y0 = y1;
        y1 = yb[q];
      }
      if (more) {
#pragma unroll
        for (int q = 0; q < CH; ++q) yb[q] = yn[q];
      }
    }
  }
  for (; i < m; ++i) {                               // tail (and the special last row)
    y2 = (double)yy[(long)(i + 2) * y_elem];
    double rhs = dmul(6.0, dadd(dsub(y0, dmul(2.0, y1)), y2));
    dp = ddiv(dsub(rhs, dmul(sys.lo[i], dp)), sys.den[i]);
    mm[(long)(i + 1) * m_elem] = dp;
    y0 = y1;
    y1 = y2;
  }
  double next = dp;                                // M[m] = dp[m-1]
  i = m - 2;
  {
    const double cp = sys.cp_c;
    double mb[CH], mn[CH];                           // steady state of the back substitution, same pipelining
    if (i - CH + 1 >= conv) {
#pragma unroll
      for (int q = 0; q < CH; ++q) mb[q] = mm[(long)(i - q + 1) * m_elem];
    }
    for (; i - CH + 1 >= conv; i -= CH) {
      const bool more = i - 2 * CH + 1 >= conv;
      if (more) {
#pragma unroll
        for (int q = 0; q < CH; ++q) mn[q] = mm[(long)(i - CH - q + 1) * m_elem];
      }
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        double v = dsub(mb[q], dmul(cp, next));
        mm[(long)(i - q + 1) * m_elem] = v;
        next = v;
      }
      if (more) {
#pragma unroll
        for (int q = 0; q < CH; ++q) mb[q] = mn[q];
      }
    }
  }
  for (; i >= 0; --i) {
    double v = dsub(mm[(long)(i + 1) * m_elem], dmul(sys.cp[i], next));
    mm[(long)(i + 1) * m_elem] = v;
    next = v;
  }
  mm[0] = dsub(dmul(2.0, mm[m_elem]), mm[2 * m_elem]);
  mm[(long)(n - 1) * m_elem] = dsub(dmul(2.0, mm[(long)(n - 2) * m_elem]), mm[(long)(n - 3) * m_elem]);
}
// axis-0 evaluation: T1[k][x] for k < H from y[h][w] (f32) and M1[h][w]
// rows [row0, row0 + T1.H) of the Hg-row result (row bands; row0 = 0, Hg = T1.H for whole images)
__global__ void spline_eval0_kernel(Img in, ImgD M1, ImgD T1, int row0, int Hg, double rHg) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= in.W || k >= T1.H) return;
  int i;
  double sfr;
  spline_locate(k + row0, in.H, Hg, rHg, &i, &sfr);
  const float* yp = in.p + (long)b * in.stride;
  const double* mp = M1.p + (long)b * M1.stride;
  double r = spline_eval((double)yp[(long)i * in.pitch + x], (double)yp[(long)(i + 1) * in.pitch + x],
                         mp[(long)i * M1.pitch + x], mp[(long)(i + 1) * M1.pitch + x], sfr);
  T1.p[(long)b * T1.stride + (long)k * T1.pitch + x] = r;
}
// axis-1 evaluation + f32 cast + optional scale (GPOF:160, 167-172)
__global__ void spline_eval1_kernel(ImgD T1, ImgD M2, Img out, float mul, int apply_mul, double rW) {
  int l = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (l >= out.W || k >= out.H) return;
  int i;
  double sfr;
  spline_locate(l, T1.W, out.W, rW, &i, &sfr);
  const double* tp = T1.p + (long)b * T1.stride + (long)k * T1.pitch;
  const double* mp = M2.p + (long)b * M2.stride + (long)k * M2.pitch;
  float r = (float)spline_eval(tp[i], tp[i + 1], mp[i], mp[i + 1], sfr);
  if (apply_mul) r = fmul(r, mul);
  out.p[(long)b * out.stride + (long)k * out.pitch + l] = r;
}
void launch_spline_seq(const Img& in, const Img& out, float mul, const SplineSys& sy, const SplineSys& sx, const ImgD& M1,
                       const ImgD& T1, const ImgD& M2, cudaStream_t s, LaunchCounter& lc, int row0, int Hg) {
  const int h = in.H, w = in.W, H = out.H;
  if (Hg <= 0) Hg = H;
  // axis 0: one thread per column
  {
    dim3 b(128), g((w + 127) / 128, 1, in.batch);
    spline_solve_kernel<float><<<g, b, 0, s>>>(in.p, in.pitch, 1, in.stride, M1.p, M1.pitch, 1, M1.stride, h, w, sy);
    dim3 b2(32, 8);
    spline_eval0_kernel<<<grid2d(w, H, in.batch, b2), b2, 0, s>>>(in, M1, T1, row0, Hg, 1.0 / (double)Hg);
  }
  // axis 1: one thread per row of T1
  {
    dim3 b(64), g((H + 63) / 64, 1, in.batch);
    spline_solve_kernel<double><<<g, b, 0, s>>>(T1.p, 1, T1.pitch, T1.stride, M2.p, 1, M2.pitch, M2.stride, w, H, sx);
    dim3 b2(32, 8);
    spline_eval1_kernel<<<grid2d(out.W, out.H, in.batch, b2), b2, 0, s>>>(T1, M2, out, mul, mul != 1.0f ? 1 : 0,
                                                                          1.0 / (double)out.W);
  }
  lc.n += 4;
}

// ---- chunk-parallel / windowed generation (ofri_spline.cuh) ----------------------------------------------------------
// Axis 0 (across rows; one thread per column and chunk, coalesced along x).  `in`, D and M are STRIPS: local row r is
// row A + r of the hg-row coarse plane.  Pass 1 -> D (forward elimination, own rows of every chunk of [RA, RF]);
// passes 2 + 3 -> M (back substitution; chunks that hold wanted rows only).  Two launches = the grid-wide barrier
// between "every chunk has stored its dp" and "a chunk reads the dp of the chunks behind it".
__global__ void spline_cols_fwd_kernel(Img in, ImgD D, int A, int hg, SplineSys sys, SplineWindow wy, int C) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= in.W) return;
  const SplineChunk k = spline_chunk(wy, blockIdx.y, C, sys.conv, kSplineWarm);
  const float* yp = in.p + (long)blockIdx.z * in.stride + x - (long)A * in.pitch;
  double* dp = D.p + (long)blockIdx.z * D.stride + x - (long)A * D.pitch;
  const long yq = in.pitch, dq = D.pitch;
  spline_chunk_forward(k, hg - 2, sys, [&](int e) { return (double)__ldg(yp + (long)e * yq); },
                       [&](int e) -> double& { return dp[(long)e * dq]; });
}
__global__ void spline_cols_bwd_kernel(ImgD D, ImgD M, int A, int hg, SplineSys sys, SplineWindow wy, int C) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= M.W) return;
  const SplineChunk k = spline_chunk(wy, blockIdx.y, C, sys.conv, kSplineWarm);
  if (k.ra > wy.RB) return;
  const double* dp = D.p + (long)blockIdx.z * D.stride + x - (long)A * D.pitch;
  double* mp = M.p + (long)blockIdx.z * M.stride + x - (long)A * M.pitch;
  const long dq = D.pitch, mq = M.pitch;
  auto Dr = [&](int e) { return dp[(long)e * dq]; };
  const double next = spline_chunk_tail(k, wy, hg - 2, kSplineWarm, sys, Dr);
  spline_chunk_back(k, hg, sys, next, Dr, [&](int e) -> double& { return mp[(long)e * mq]; });
}

// Axis 0 evaluation + axis 1 solve + axis 1 evaluation of RPB output rows (x segment `blockIdx.x`) in ONE block:
//   phase 0  T[r][e] = axis-0 spline of column e at output row k (all threads, coalesced along e) -> shared
//   phase 1  forward elimination, one thread per (row, chunk) of the segment's window
//   phase 2  back substitution through the rows behind each chunk (read only)      } block barriers in between
//   phase 3  back substitution of each chunk's own rows, in place                   }
//   phase 4  out[k][l] for the segment's output columns (all threads, coalesced along l)
// so the f64 intermediates (T1, M2 of the previous generation: 16 B per coarse column and output row each way) never
// touch HBM.  Chunks are an odd number of doubles apart -> conflict-free shared-memory access in the solve phases.
struct SplineRowsArgs {
  int A, hg;            // coarse strip origin / coarse plane height
  int row0, Hg;         // output strip origin / output plane height
  double rHg, rW;       // RN(1 / Hg), RN(1 / out.W)
  float mul;
  int apply_mul;
  int seg_len, nseg;    // knot intervals per x segment
  int C, rpb;           // chunk size (odd), rows per block
  int ne_max;           // doubles per row and array in shared memory
  int small_ix;         // out.W * in.W < 2^31: 32-bit index arithmetic in spline_locate
};
__device__ __forceinline__ void spline_locate_dev(int k, int n, int N, double rN, int small, int* i_out, double* s_out) {
  if (small) {
    const unsigned num = (unsigned)k * (unsigned)n;
    unsigned i = num / (unsigned)N;
    double s = ddiv_const((double)(num - i * (unsigned)N), (double)N, rN);
    if ((int)i >= n - 1) { i = n - 2; s = 1.0; }
    *i_out = (int)i;
    *s_out = s;
  } else {
    spline_locate(k, n, N, rN, i_out, s_out);
  }
}
template <int NT>
__global__ void __launch_bounds__(NT) spline_rows_kernel(Img in, ImgD M1, Img out, SplineSys sx, SplineRowsArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, b = blockIdx.z;
  const int w = in.W, m = w - 2;
  const int k0 = blockIdx.y * a.rpb;                                    // first local output row of this block
  const int nrows = min(a.rpb, out.H - k0);
  // segment -> wanted unknowns [xa, xb] -> window
  const int xa = blockIdx.x * a.seg_len;
  const bool last_seg = blockIdx.x == a.nseg - 1;
  const int xb = last_seg ? w - 1 : xa + a.seg_len;
  const SplineWindow wx = spline_window(xa, xb, w, sx.conv, kSplineWarm);
  const int E0 = wx.FS, ne = wx.RF + 2 - wx.FS + 1;                     // elements E0 .. E0 + ne - 1 are staged
  double* ysm = sm;
  double* msm = sm + (size_t)a.rpb * a.ne_max;
  const float* ip = in.p + (long)b * in.stride;
  const double* m1 = M1.p + (long)b * M1.stride;
  // ---- phase 0 ------------------------------------------------------------------------------------------------------
  for (int r = 0; r < nrows; ++r) {
    int i;
    double sfr;
    spline_locate(k0 + r + a.row0, a.hg, a.Hg, a.rHg, &i, &sfr);
    const SplinePos pos = spline_pos(sfr);
    const float* y0 = ip + (long)(i - a.A) * in.pitch + E0;
    const float* y1 = y0 + in.pitch;
    const double* q0 = m1 + (long)(i - a.A) * M1.pitch + E0;
    const double* q1 = q0 + M1.pitch;
    double* dst = ysm + (size_t)r * a.ne_max;
    int e = tid;
    for (; e + NT < ne; e += 2 * NT) {        // two independent samples per trip: their loads and chains interleave
      const double ya = (double)__ldg(y0 + e), yb = (double)__ldg(y1 + e), ma = __ldg(q0 + e), mb = __ldg(q1 + e);
      const double yc = (double)__ldg(y0 + e + NT), yd = (double)__ldg(y1 + e + NT), mc = __ldg(q0 + e + NT),
                   md = __ldg(q1 + e + NT);
      dst[e] = spline_eval_at(ya, yb, ma, mb, pos);
      dst[e + NT] = spline_eval_at(yc, yd, mc, md, pos);
    }
    if (e < ne)
      dst[e] = spline_eval_at((double)__ldg(y0 + e), (double)__ldg(y1 + e), __ldg(q0 + e), __ldg(q1 + e), pos);
  }
  __syncthreads();
  // ---- phases 1-3 ---------------------------------------------------------------------------------------------------
  const int nch = spline_num_chunks(wx, a.C);
  const int r = tid / nch, c = tid - r * nch;
  const bool solver = r < nrows;
  SplineChunk k;
  double* yr = ysm + (size_t)r * a.ne_max - E0;
  double* mr = msm + (size_t)r * a.ne_max - E0;
  auto D = [&](int e) -> double& { return mr[e]; };
  if (solver) {
    k = spline_chunk(wx, c, a.C, sx.conv, kSplineWarm);
    spline_chunk_forward(k, m, sx, [&](int e) { return yr[e]; }, D);
  }
  __syncthreads();
  double next = 0.0;
  const bool wanted = solver && k.ra <= wx.RB;
  if (wanted) next = spline_chunk_tail(k, wx, m, kSplineWarm, sx, D);
  __syncthreads();
  if (wanted) spline_chunk_back(k, w, sx, next, D, D);
  __syncthreads();
  // ---- phase 4: one thread per knot interval i; its outputs l in [ceil(i W / w), ceil((i+1) W / w)) (two at 2:1) share
  // the interval's samples and the position-independent quotients M/6 ------------------------------------------------
  const long W = out.W;
  const int i_end = last_seg ? w - 1 : xb;                                 // intervals xa .. i_end - 1
  for (int rr = 0; rr < nrows; ++rr) {
    const double* tp = ysm + (size_t)rr * a.ne_max - E0;
    const double* mp = msm + (size_t)rr * a.ne_max - E0;
    float* op = out.p + (long)b * out.stride + (long)(k0 + rr) * out.pitch;
    for (int i = xa + tid; i < i_end; i += NT) {
      const double yi = tp[i], yj = tp[i + 1], Mi = mp[i], Mj = mp[i + 1];
      const SplineM6 q = spline_m6(Mi, Mj);
      const int la = (int)(((long)i * W + w - 1) / w);
      const int lb = i == w - 2 ? (int)W : (int)(((long)(i + 1) * W + w - 1) / w);
      for (int l = la; l < lb; ++l) {
        int il;
        double sfr;
        spline_locate_dev(l, w, (int)W, a.rW, a.small_ix, &il, &sfr);        // il == i
        float v = (float)spline_eval_m6(yi, yj, Mi, Mj, q, spline_pos(sfr));
        if (a.apply_mul) v = fmul(v, a.mul);
        op[l] = v;
      }
    }
  }
}

void spline_rows_needed(int row0, int rows, int hg, int Hg, const SplineSys& sy, int* lo, int* hi) {
  int ia, ib;
  double sfr;
  spline_locate(row0, hg, Hg, &ia, &sfr);
  spline_locate(row0 + rows - 1, hg, Hg, &ib, &sfr);
  const SplineWindow wy = spline_window(ia, ib + 1, hg, sy.conv, kSplineWarm);
  *lo = wy.FS;
  *hi = wy.RF + 3;
}

bool launch_spline(const Img& in, const Img& out, float mul, const SplineSys& sy, const SplineSys& sx, const ImgD& M1,
                   const ImgD& D1, cudaStream_t s, LaunchCounter& lc, int in_row0, int hg, int row0, int Hg) {
  if (hg <= 0) hg = in.H;
  if (Hg <= 0) Hg = out.H;
  const int w = in.W, A = in_row0;
  // ---- axis 0: unknowns M1[ia .. ib + 1] of every column -----------------------------------------------------------
  int ia, ib;
  double sfr;
  spline_locate(row0, hg, Hg, &ia, &sfr);
  spline_locate(row0 + out.H - 1, hg, Hg, &ib, &sfr);
  const SplineWindow wy = spline_window(ia, ib + 1, hg, sy.conv, kSplineWarm);
  if (wy.FS < A || wy.RF + 2 > A + in.H - 1) return false;
  if (M1.H < in.H || D1.H < in.H || M1.W < w || D1.W < w) return false;
  {
    const int range = wy.RF - wy.RA + 1;
    long want = 300000L / ((long)w * in.batch);                 // ~2 threads per lane of the GPU
    if (want < 1) want = 1;
    if (want > range / 32) want = range / 32 > 0 ? range / 32 : 1;
    int C = (int)((range + want - 1) / want);
    if (C < 2) C = 2;
    const int nch = spline_num_chunks(wy, C);
    dim3 b(128), g((w + 127) / 128, nch, in.batch);
    spline_cols_fwd_kernel<<<g, b, 0, s>>>(in, D1, A, hg, sy, wy, C);
    spline_cols_bwd_kernel<<<g, b, 0, s>>>(D1, M1, A, hg, sy, wy, C);
  }
  // ---- axes 0 (evaluation) + 1, fused per output row ------------------------------------------------------------------
  {
    constexpr int NT = 256;
    constexpr int MAXE = 4352;                                    // doubles per row and array: 2 x 34 KB -> 3 blocks / SM
    SplineRowsArgs a;
    a.A = A; a.hg = hg; a.row0 = row0; a.Hg = Hg;
    a.rHg = 1.0 / (double)Hg;
    a.rW = 1.0 / (double)out.W;
    a.mul = mul;
    a.apply_mul = mul != 1.0f ? 1 : 0;
    a.small_ix = (long)out.W * (long)w < 0x7fffffffL ? 1 : 0;
    const int margin = 2 * kSplineWarm + 4;
    a.nseg = w <= MAXE ? 1 : (w + (MAXE - margin) - 1) / (MAXE - margin);
    a.seg_len = a.nseg == 1 ? w : (w + a.nseg - 1) / a.nseg;
    const int ne = a.nseg == 1 ? w : a.seg_len + margin;
    a.ne_max = ne | 1;                                            // odd row pitch
    int rpb = (int)(65536 / ((long)a.ne_max * 16));
    a.rpb = rpb < 1 ? 1 : (rpb > 8 ? 8 : rpb);
    const int range = a.nseg == 1 ? w - 2 : a.seg_len + kSplineWarm + 1;
    const int per_row = NT / a.rpb;                               // solver threads per row
    int C = (range + per_row - 1) / per_row;
    if (C < 9) C = 9;
    a.C = C | 1;
    const size_t smem = (size_t)2 * a.rpb * a.ne_max * sizeof(double);
    if (cudaFuncSetAttribute(spline_rows_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    dim3 g(a.nseg, (out.H + a.rpb - 1) / a.rpb, in.batch);
    spline_rows_kernel<NT><<<g, NT, smem, s>>>(in, M1, out, sx, a);
  }
  lc.n += 3;
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Bilinear warp (GPOF:70-116, 200-201)
// ---------------------------------------------------------------------------------------------------------------
// us/vs/o1/o2: rows [row0, row0 + us.H) of the image; im1/im2: rows [img_row0, img_row0 + im1.H); Hg = image height
__global__ void warp_pair_kernel(Img im1, Img im2, Img us, Img vs, Img o1, Img o2, int row0, int img_row0, int Hg) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= us.W || y >= us.H) return;
  float u = us.p[(long)b * us.stride + (long)y * us.pitch + x];
  float v = vs.p[(long)b * vs.stride + (long)y * vs.pitch + x];
  float cy1 = warp_coord(y + row0, v, -1.0f), cx1 = warp_coord(x, u, -1.0f);
  float cy2 = warp_coord(y + row0, v, +1.0f), cx2 = warp_coord(x, u, +1.0f);
  o1.p[(long)b * o1.stride + (long)y * o1.pitch + x] =
      warp_sample(im1.p + (long)b * im1.stride, im1.pitch, im1.H, im1.W, cy1, cx1, img_row0, Hg);
  o2.p[(long)b * o2.stride + (long)y * o2.pitch + x] =
      warp_sample(im2.p + (long)b * im2.stride, im2.pitch, im2.H, im2.W, cy2, cx2, img_row0, Hg);
}
__global__ void warp_coords_kernel(Img img, Img cy, Img cx, Img o) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (x >= img.W || y >= img.H) return;
  float fy = cy.p[(long)b * cy.stride + (long)y * cy.pitch + x];
  float fx = cx.p[(long)b * cx.stride + (long)y * cx.pitch + x];
  o.p[(long)b * o.stride + (long)y * o.pitch + x] =
      warp_sample(img.p + (long)b * img.stride, img.pitch, img.H, img.W, fy, fx);
}
void launch_warp_pair(const Img& im1, const Img& im2, const Img& us, const Img& vs, const Img& out1, const Img& out2,
                      cudaStream_t s, LaunchCounter& lc, int row0, int img_row0, int Hg) {
  dim3 b(32, 8);
  if (Hg <= 0) Hg = im1.H;
  warp_pair_kernel<<<grid2d(us.W, us.H, us.batch, b), b, 0, s>>>(im1, im2, us, vs, out1, out2, row0, img_row0, Hg);
  lc.n += 1;
}
void launch_warp_coords(const Img& img, const Img& cy, const Img& cx, const Img& out, cudaStream_t s,
                        LaunchCounter& lc) {
  dim3 b(32, 8);
  warp_coords_kernel<<<grid2d(img.W, img.H, img.batch, b), b, 0, s>>>(img, cy, cx, out);
  lc.n += 1;
}

// ---------------------------------------------------------------------------------------------------------------
// "Liu-Shen warp" (biLinear = False; GPOF:190-196, 204-221): integer-shift scatter of frame 1 along the rounded flow,
// then the optical-flow equation with the Gaussian-smoothed sub-pixel parts.
// ---------------------------------------------------------------------------------------------------------------
// Scatter `im1[vsSwap, usSwap] = im1[ysMesh, xsMesh]`: sources are visited in row-major order and the last writer of
// a target wins -> atomicMax of the source's linear index per target.  Negative targets wrap around (numpy indexing);
// targets >= size or < -size raise IndexError in the reference -> flag.  Also emits the sub-pixel parts.
__global__ void lsw_prepare_kernel(Img us, Img vs, int* winner, Img dU, Img dV, int* flag) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  const int W = us.W, H = us.H;
  if (x >= W || y >= H) return;
  const float u = us.p[(long)b * us.stride + (long)y * us.pitch + x];
  const float v = vs.p[(long)b * vs.stride + (long)y * vs.pitch + x];
  const float fu = floorf(fadd(u, 0.5f)), fv = floorf(fadd(v, 0.5f));
  dU.p[(long)b * dU.stride + (long)y * dU.pitch + x] = fsub(u, fu);
  dV.p[(long)b * dV.stride + (long)y * dV.pitch + x] = fsub(v, fv);
  double txd = dadd((double)x, (double)fu), tyd = dadd((double)y, (double)fv);     // int32 + float32 -> float64 -> int32
  if (!(txd >= -(double)W && txd < (double)W && tyd >= -(double)H && tyd < (double)H)) {
    atomicExch(flag, 1);
    return;
  }
  int tx = (int)txd, ty = (int)tyd;
  if (tx < 0) tx += W;
  if (ty < 0) ty += H;
  atomicMax(winner + (long)b * H * W + (long)ty * W + tx, y * W + x);
}
__global__ void lsw_gather_kernel(Img im1, const int* winner, Img out) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  const int W = im1.W, H = im1.H;
  if (x >= W || y >= H) return;
  const float* I = im1.p + (long)b * im1.stride;
  const int w = winner[(long)b * H * W + (long)y * W + x];
  const int sy = w >= 0 ? w / W : y, sx = w >= 0 ? w - (w / W) * W : x;     // untouched targets keep their value
  out.p[(long)b * out.stride + (long)y * out.pitch + x] = I[(long)sy * im1.pitch + sx];
}
// im1[0:-1,0:-1] -= (tempDx + tempDy) with tempDx = I[y,x+1] dU[y,x+1] - I[y,x] dU[y,x], tempDy likewise along y (f32)
__global__ void lsw_ofeq_kernel(Img in, Img dU, Img dV, Img out) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  const int W = in.W, H = in.H;
  if (x >= W || y >= H) return;
  const float* I = in.p + (long)b * in.stride;
  const float* A = dU.p + (long)b * dU.stride;
  const float* B = dV.p + (long)b * dV.stride;
  float r = I[(long)y * in.pitch + x];
  if (x < W - 1 && y < H - 1) {
    const float c = r;
    const float tdx = fsub(fmul(I[(long)y * in.pitch + x + 1], A[(long)y * dU.pitch + x + 1]), fmul(c, A[(long)y * dU.pitch + x]));
    const float tdy = fsub(fmul(I[(long)(y + 1) * in.pitch + x], B[(long)(y + 1) * dV.pitch + x]), fmul(c, B[(long)y * dV.pitch + x]));
    r = fsub(c, fadd(tdx, tdy));
  }
  out.p[(long)b * out.stride + (long)y * out.pitch + x] = r;
}
// im1 -> out.  Scratch: winner [batch][H][W] ints, dU / dV (raw sub-pixel parts), fU / fV (filtered), sc (scattered
// frame), tmp (Gaussian scratch); taps = the 73-tap kernel of gaussian_filter(x, 0.6*3, truncate=4/0.6*3).
void launch_liu_shen_warp(const Img& im1, const Img& us, const Img& vs, const Img& out, int* winner, const Img& dU,
                          const Img& dV, const Img& fU, const Img& fV, const Img& sc, const Img& tmp,
                          const GaussTaps& taps, int* flag, cudaStream_t s, LaunchCounter& lc) {
  dim3 b(32, 8);
  dim3 g = grid2d(im1.W, im1.H, im1.batch, b);
  cudaMemsetAsync(winner, 0xff, sizeof(int) * (size_t)im1.H * im1.W * im1.batch, s);
  lsw_prepare_kernel<<<g, b, 0, s>>>(us, vs, winner, dU, dV, flag);
  lsw_gather_kernel<<<g, b, 0, s>>>(im1, winner, sc);
  lc.n += 2;
  launch_gauss(dU, tmp, fU, taps, s, lc);
  launch_gauss(dV, tmp, fV, taps, s, lc);
  lsw_ofeq_kernel<<<g, b, 0, s>>>(sc, fU, fV, out);
  lc.n += 1;
}

// ---------------------------------------------------------------------------------------------------------------
// element-wise helpers
// ---------------------------------------------------------------------------------------------------------------
__global__ void axpy_kernel(Img acc, Img x) {
  int xx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (xx >= acc.W || y >= acc.H) return;
  long ia = (long)b * acc.stride + (long)y * acc.pitch + xx;
  acc.p[ia] = fadd(acc.p[ia], x.p[(long)b * x.stride + (long)y * x.pitch + xx]);
}
__global__ void scale_kernel(Img x, float mul) {
  int xx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (xx >= x.W || y >= x.H) return;
  long i = (long)b * x.stride + (long)y * x.pitch + xx;
  x.p[i] = fmul(x.p[i], mul);
}
__global__ void copy_kernel(Img d, Img sr) {
  int xx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (xx >= d.W || y >= d.H) return;
  d.p[(long)b * d.stride + (long)y * d.pitch + xx] = sr.p[(long)b * sr.stride + (long)y * sr.pitch + xx];
}
__global__ void fill_kernel(Img d, float v) {
  int xx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y, b = blockIdx.z;
  if (xx >= d.W || y >= d.H) return;
  d.p[(long)b * d.stride + (long)y * d.pitch + xx] = v;
}
void launch_axpy(const Img& acc, const Img& x, cudaStream_t s, LaunchCounter& lc) {
  dim3 b(32, 8);
  axpy_kernel<<<grid2d(acc.W, acc.H, acc.batch, b), b, 0, s>>>(acc, x);
  lc.n += 1;
}
void launch_scale(const Img& x, float mul, cudaStream_t s, LaunchCounter& lc) {
  dim3 b(32, 8);
  scale_kernel<<<grid2d(x.W, x.H, x.batch, b), b, 0, s>>>(x, mul);
  lc.n += 1;
}
void launch_copy(const Img& dst, const Img& src, cudaStream_t s, LaunchCounter& lc) {
  dim3 b(32, 8);
  copy_kernel<<<grid2d(dst.W, dst.H, dst.batch, b), b, 0, s>>>(dst, src);
  lc.n += 1;
}
void launch_fill(const Img& dst, float v, cudaStream_t s, LaunchCounter& lc) {
  dim3 b(32, 8);
  fill_kernel<<<grid2d(dst.W, dst.H, dst.batch, b), b, 0, s>>>(dst, v);
  lc.n += 1;
}

const char* kernel_build_info() { return "libofri sm_100a " __DATE__ " " __TIME__; }

}  // namespace ofri
