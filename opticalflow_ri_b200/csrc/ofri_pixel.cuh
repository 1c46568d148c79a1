// ofri_pixel.cuh -- per-pixel arithmetic of every stage, shared by the simple kernels, the fused (temporally
// blocked) kernels and the test-only host harness (tests/hostcheck), so that index rules and rounding order are
// written exactly once.  All functions are __host__ __device__; nothing here touches memory spaces or threads.
//
// Rounding discipline: wherever the reference's result depends on separate rounding of each operation
// (numba f32 loops without FMA contraction, numpy f64 expressions) the explicit *_rn helpers are used so the
// compiler cannot contract.  The two iterative solvers use explicit fmaf() chains instead (fast path; within
// the 1e-4 px tolerance) -- explicit, so the simple and fused kernels are bit-identical.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OFRI_HD __host__ __device__ __forceinline__
#else
#define OFRI_HD inline
#endif

namespace ofri {

// ---- exactly-rounded scalar ops -------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
OFRI_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
OFRI_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
OFRI_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
OFRI_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
OFRI_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
OFRI_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
OFRI_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
OFRI_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
#else   // host build (tests/hostcheck is compiled with -ffp-contract=off)
OFRI_HD float fmul(float a, float b) { volatile float r = a * b; return r; }
OFRI_HD float fadd(float a, float b) { volatile float r = a + b; return r; }
OFRI_HD float fsub(float a, float b) { volatile float r = a - b; return r; }
OFRI_HD float fdiv(float a, float b) { volatile float r = a / b; return r; }
OFRI_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
OFRI_HD double dadd(double a, double b) { volatile double r = a + b; return r; }
OFRI_HD double dsub(double a, double b) { volatile double r = a - b; return r; }
OFRI_HD double ddiv(double a, double b) { volatile double r = a / b; return r; }
#endif

OFRI_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// scipy 'mirror' for a radius-1 stencil: -1 -> 1, n -> n-2 (degenerate n == 1 -> 0)
OFRI_HD int mirror1(int i, int n) {
  if (n == 1) return 0;
  return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i);
}

// ---- Gaussian pre-filter (gaussian_filter.py:54-85) ------------------------------------------------------
// Source index, in the unpadded line of n samples, of padded position p in [0, n+2h):
//   left/top  P[h-1-j] = a[j]        (mirror including the edge sample)
//   right/bot P[n+2h-1-j] = a[n-1-j] (forward copy of the last h samples -- reference quirk)
OFRI_HD int gauss_src_index(int p, int n, int h) {
  if (p < h) return h - 1 - p;
  if (p < h + n) return p - h;
  return p - 2 * h;
}
// out[x] = (((0 + P[x+2h] k0) + P[x+2h-1] k1) + ... + P[x] k[K-1]); separate f32 multiply and add
// (gaussian_filter.py:37-40).  `line` is addressed as line[idx*stride].
template <typename Taps>
OFRI_HD float gauss_point(const float* line, long stride, int x, int n, const Taps& k, int K) {
  const int h = K >> 1;
  float acc = 0.0f;
  for (int j = 0; j < K; ++j) {
    int src = gauss_src_index(x + 2 * h - j, n, h);
    acc = fadd(acc, fmul(line[(long)src * stride], k[j]));
  }
  return acc;
}

// ---- Pillow BICUBIC resample (libImaging/Resample.c, mode F) ---------------------------------------------
OFRI_HD double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
  if (x < 2.0) return (((x - 5.0) * x + 8.0) * x - 4.0) * a;
  return 0.0;
}
// one output sample: f64 accumulate, ascending taps, f32 store
OFRI_HD float resample_point(const float* line, long stride, int xmin, int cnt, const double* w) {
  double ss = 0.0;
  for (int t = 0; t < cnt; ++t) ss = dadd(ss, dmul((double)line[(long)(xmin + t) * stride], w[t]));
  return (float)ss;
}

// ---- not-a-knot cubic spline evaluation (FITPACK bispev restated) ------------------------------
// Correctly rounded x / d for a divisor d known in advance, r = RN(1/d): quotient estimate + two FMA residual
// corrections (the tail of every IEEE division routine; exact for the normal-range operands that occur here) -- the
// same double as ddiv(x, d), at a fraction of the cost.
OFRI_HD double ddiv_const(double x, double d, double r) {
  double q = dmul(x, r);
  double e = fma(-d, q, x);
  q = fma(e, r, q);
  e = fma(-d, q, x);
  return fma(e, r, q);
}
// position of output sample k of N over n input samples: i = floor(k n / N), s = frac, clamped to the last knot;
// rN = RN(1 / N)
OFRI_HD void spline_locate(int k, int n, int N, double rN, int* i_out, double* s_out) {
  long long num = (long long)k * (long long)n;
  long long i = num / N;
  double s = ddiv_const((double)(num - i * N), (double)N, rN);
  if (i >= n - 1) { i = n - 2; s = 1.0; }
  *i_out = (int)i;
  *s_out = s;
}
OFRI_HD void spline_locate(int k, int n, int N, int* i_out, double* s_out) {
  spline_locate(k, n, N, ddiv(1.0, (double)N), i_out, s_out);
}
// the part of the evaluation that depends on the position only (shared by all samples of one output row / column)
struct SplinePos { double s, t, s3, t3; };
OFRI_HD SplinePos spline_pos(double s) {
  SplinePos p;
  p.s = s;
  p.t = dsub(1.0, s);
  p.t3 = dmul(dmul(p.t, p.t), p.t);
  p.s3 = dmul(dmul(s, s), s);
  return p;
}
OFRI_HD double spline_eval_at(double yi, double yj, double Mi, double Mj, const SplinePos& p) {
  const double r6 = 0.16666666666666666;      // RN(1/6)
  double a = ddiv_const(dmul(Mi, p.t3), 6.0, r6);
  double b = ddiv_const(dmul(Mj, p.s3), 6.0, r6);
  double c = dmul(dsub(yi, ddiv_const(Mi, 6.0, r6)), p.t);
  double d = dmul(dsub(yj, ddiv_const(Mj, 6.0, r6)), p.s);
  return dadd(dadd(dadd(a, b), c), d);
}
// the same value with the position-independent quotients M/6 hoisted (one interval, several positions)
struct SplineM6 { double i6, j6; };
OFRI_HD SplineM6 spline_m6(double Mi, double Mj) {
  const double r6 = 0.16666666666666666;
  SplineM6 q;
  q.i6 = ddiv_const(Mi, 6.0, r6);
  q.j6 = ddiv_const(Mj, 6.0, r6);
  return q;
}
OFRI_HD double spline_eval_m6(double yi, double yj, double Mi, double Mj, const SplineM6& q, const SplinePos& p) {
  const double r6 = 0.16666666666666666;
  double a = ddiv_const(dmul(Mi, p.t3), 6.0, r6);
  double b = ddiv_const(dmul(Mj, p.s3), 6.0, r6);
  double c = dmul(dsub(yi, q.i6), p.t);
  double d = dmul(dsub(yj, q.j6), p.s);
  return dadd(dadd(dadd(a, b), c), d);
}
OFRI_HD double spline_eval(double yi, double yj, double Mi, double Mj, double s) {
  return spline_eval_at(yi, yj, Mi, Mj, spline_pos(s));
}

// ---- bilinear warp (GenericPyramidalOpticalFlow.py:70-116) --------------------------------------------------
// coordinate of GPOF:200-201: f32( f64(idx) -/+ f64(f32(flow/2)) )
OFRI_HD float warp_coord(int idx, float flow, float sign) {
  float half = fmul(flow, 0.5f);
  double c = sign < 0.0f ? dsub((double)idx, (double)half) : dadd((double)idx, (double)half);
  return (float)c;
}
// Band form (row-band domain decomposition): `img` holds rows [row0, row0 + H) of an image of Hg rows; cy is the GLOBAL
// row coordinate (its float32 rounding depends on the magnitude of the row index, so it must be formed globally).
// Rows are clamped to the image first (the reference's rule), then to the band (only reached when the band's halo is
// too small for the displacement; the driver checks that).  row0 = 0, Hg = H is the whole-image case.
OFRI_HD float warp_sample(const float* img, long pitch, int H, int W, float cy, float cx, int row0 = 0, int Hg = -1) {
  if (Hg < 0) Hg = H;
  // np.int32(np.round(c)): half-to-even.  Clamp first so absurd / NaN coordinates saturate instead of trapping
  // (the reference's behaviour there is undefined; inside +-1e9 the result is identical).
  int iy = (int)fminf(fmaxf(rintf(cy), -1.0e9f), 1.0e9f);
  int ix = (int)fminf(fmaxf(rintf(cx), -1.0e9f), 1.0e9f);
  double dy = dsub((double)cy, (double)iy);
  double dx = dsub((double)cx, (double)ix);
  int ny = dy < 0.0 ? iy - 1 : iy + 1;
  int nx = dx < 0.0 ? ix - 1 : ix + 1;
  dy = fabs(dy);
  dx = fabs(dx);
  iy = clampi(clampi(iy, 0, Hg - 1) - row0, 0, H - 1);
  ix = clampi(ix, 0, W - 1);
  ny = clampi(clampi(ny, 0, Hg - 1) - row0, 0, H - 1);
  nx = clampi(nx, 0, W - 1);
  double i00 = (double)img[(long)iy * pitch + ix];
  double i01 = (double)img[(long)iy * pitch + nx];
  double i10 = (double)img[(long)ny * pitch + ix];
  double i11 = (double)img[(long)ny * pitch + nx];
  double oy = dsub(1.0, dy), ox = dsub(1.0, dx);
  double r = dmul(dmul(oy, ox), i00);
  r = dadd(r, dmul(dmul(oy, dx), i01));
  r = dadd(r, dmul(dmul(dy, ox), i10));
  r = dadd(r, dmul(dmul(dy, dx), i11));
  return (float)r;
}

// ---- Horn-Schunck derivatives (HornSchunck.py:107-127 through the swaps of :37/:73/:84) ---------------------
// a.. = 2x2 block of frame1 (im1 of compute), b.. = same block of frame2; index i+1 / j+1 mirrored (n -> n-2).
OFRI_HD void hs_deriv_point(float a00, float a01, float a10, float a11, float b00, float b01, float b10, float b11,
                            float* fx, float* fy, float* ft) {
  // each scipy convolve: f64 accumulate (exact for these magnitudes), one rounding to f32; then f32 adds
  float gxa = (float)(dadd(dadd(dadd(dmul(a00, 0.25), dmul(a01, -0.25)), dmul(a10, 0.25)), dmul(a11, -0.25)));
  float gxb = (float)(dadd(dadd(dadd(dmul(b00, 0.25), dmul(b01, -0.25)), dmul(b10, 0.25)), dmul(b11, -0.25)));
  float gya = (float)(dadd(dadd(dadd(dmul(a00, 0.25), dmul(a01, 0.25)), dmul(a10, -0.25)), dmul(a11, -0.25)));
  float gyb = (float)(dadd(dadd(dadd(dmul(b00, 0.25), dmul(b01, 0.25)), dmul(b10, -0.25)), dmul(b11, -0.25)));
  float bxa = (float)(dadd(dadd(dadd(dmul(a00, 0.25), dmul(a01, 0.25)), dmul(a10, 0.25)), dmul(a11, 0.25)));
  float bxb = (float)(dadd(dadd(dadd(dmul(b00, -0.25), dmul(b01, -0.25)), dmul(b10, -0.25)), dmul(b11, -0.25)));
  *fx = fadd(gxb, gxa);
  *fy = fadd(gyb, gya);
  *ft = fadd(bxa, bxb);
}

// ---- Horn-Schunck Jacobi update (HornSchunck.py:52-71) -------------------------------------------------------
// inv = 1 / (alpha^2 + fx^2 + fy^2), iteration invariant
OFRI_HD float hs_inv_den(float fx, float fy, float alpha2) {
  return fdiv(1.0f, fmaf(fy, fy, fmaf(fx, fx, alpha2)));
}
// 3x3 weighted average: 1/6 on the 4 edge neighbours, 1/12 on the 4 corners (mirror applied by the caller)
OFRI_HD float hs_avg(float n, float s, float w, float e, float nw, float ne, float sw, float se) {
  float edges = fadd(fadd(n, s), fadd(w, e));
  float corners = fadd(fadd(nw, ne), fadd(sw, se));
  return fmaf(corners, 0.083333336f, fmul(edges, 0.16666667f));
}
// same average from column partial sums: vs_c = n_c + s_c of column c; m_c = centre-row sample of column c
OFRI_HD float hs_avg_cols(float vsl, float vsc, float vsr, float ml, float mr) {
  float edges = fadd(vsc, fadd(ml, mr));
  float corners = fadd(vsl, vsr);
  return fmaf(corners, 0.083333336f, fmul(edges, 0.16666667f));
}
OFRI_HD void hs_update(float ua, float va, float fx, float fy, float ft, float inv, float* u, float* v) {
  float der = fmul(fmaf(fx, ua, fmaf(fy, va, ft)), inv);
  *u = fmaf(-fx, der, ua);
  *v = fmaf(-fy, der, va);
}
// Fast path used by the kernels: NORMALISED coefficients a = fx n, b = fy n, c = ft n, n = 1/sqrt(alpha^2+fx^2+fy^2),
// prepared once per level.  Then fx (fx ua + fy va + ft)/(alpha^2+fx^2+fy^2) = a (a ua + b va + c): three
// coefficients per pixel instead of four, four FMAs per update, no division in the sweep.
OFRI_HD float rsqrt_rn(float x) {
#if defined(__CUDA_ARCH__)
  return __frsqrt_rn(x);
#else
  return (float)(1.0 / sqrt((double)x));
#endif
}
OFRI_HD void hs_normalise(float fx, float fy, float ft, float alpha2, float* a, float* b, float* c) {
  float n = rsqrt_rn(fmaf(fy, fy, fmaf(fx, fx, alpha2)));
  *a = fmul(fx, n);
  *b = fmul(fy, n);
  *c = fmul(ft, n);
}
OFRI_HD void hs_update_n(float ua, float va, float a, float b, float c, float* u, float* v) {
  float g = fmaf(a, ua, fmaf(b, va, c));
  *u = fmaf(-a, g, ua);
  *v = fmaf(-b, g, va);
}
// "precise" formulation = the reference's arithmetic, operation by operation:
//   stencil: scipy correlate accumulates the 8 products in float64 and rounds ONCE to float32.  The weights are
//   w6 = f32(1/6) and w12 = f32(1/12) = w6/2, so the exact value is w6*E + w12*C with E, C the (exact in f64) sums of
//   the edge / corner neighbours; evaluated here in f64 it differs from scipy's sequential f64 sum by ~1e-16
//   relative, i.e. the float32 results agree except in ~1e-9 of the evaluations (double-rounding ties).
//   update: numba evaluates every f32 operation separately rounded, with a true division (HornSchunck.py:55-58).
OFRI_HD float hs_den(float fx, float fy, float alpha2) { return fadd(fadd(alpha2, fmul(fx, fx)), fmul(fy, fy)); }
// w6 = 2 w12 exactly, so w6 E + w12 C = w12 (2E + C): X = 2E + C is exact in f64 whenever the eight neighbours span
// fewer than ~29 binary orders of magnitude, and then w12 X -- one rounding -- equals the value above (two exact products,
// one rounded sum) bit for bit; one DFMA + one DMUL instead of two DMUL + one DADD.
OFRI_HD float hs_avg_cols_precise(double vsl, double vsc, double vsr, double ml, double mr) {
  double E = dadd(vsc, dadd(ml, mr));
  double C = dadd(vsl, vsr);
  return (float)dmul(fma(2.0, E, C), (double)0.083333336f);
}
// RN(s / den) from rcp = RN(1/den): product, then two residual corrections (the sequence the hardware division
// expands to, without its range checks: den >= alpha^2 is always a normal number here)
OFRI_HD float div_rn_rcp(float s, float den, float rcp) {
  float q = fmul(s, rcp);
  float e = fmaf(-den, q, s);
  q = fmaf(e, rcp, q);
  e = fmaf(-den, q, s);
  return fmaf(e, rcp, q);
}
OFRI_HD float rcp_rn(float x) {
#if defined(__CUDA_ARCH__)
  return __frcp_rn(x);
#else
  return fdiv(1.0f, x);
#endif
}
OFRI_HD void hs_update_precise(float ua, float va, float fx, float fy, float ft, float den, float rcp, float* u,
                               float* v) {
  float der = div_rn_rcp(fadd(fadd(fmul(fx, ua), fmul(fy, va)), ft), den, rcp);
  *u = fsub(ua, fmul(fx, der));
  *v = fsub(va, fmul(fy, der));
}

// ---- Liu-Shen (PhysicsBasedOpticalFlowLiuShen.py:47-158) ------------------------------------------------------
// coefficient planes of one pixel from the 3x3 neighbourhoods (clamp-to-edge applied by the caller) of the
// NORMALISED images i1, i2; cnt = number of in-bounds 8-neighbours (8 / 5 / 3).  Every product and sum is a
// separately rounded f32 operation, as in numpy; stencil sums are f64-accumulated and rounded once, as in scipy.
struct LsCoef { float IIx, IIy, II, Ixt, Iyt, B11, B12, B22; };
OFRI_HD float ls_stencil_d(float lo, float hi) {   // (hi - lo)/2 : f64 accumulate of (-0.5 lo) + (0.5 hi)
  return (float)dadd(dmul((double)lo, -0.5), dmul((double)hi, 0.5));
}
OFRI_HD LsCoef ls_coef_point(const float a[3][3], const float d[3][3], float hpar, float cnt) {
  // a = i1 neighbourhood, d = (i2 - i1) neighbourhood (f32 difference), [row][col], centre [1][1]
  LsCoef c;
  float i1 = a[1][1];
  c.IIx = fmul(i1, ls_stencil_d(a[0][1], a[2][1]));
  c.IIy = fmul(i1, ls_stencil_d(a[1][0], a[1][2]));
  c.II = fmul(i1, i1);
  c.Ixt = fmul(i1, ls_stencil_d(d[0][1], d[2][1]));
  c.Iyt = fmul(i1, ls_stencil_d(d[1][0], d[1][2]));
  float d2r = (float)dadd(dadd((double)a[0][1], dmul((double)a[1][1], -2.0)), (double)a[2][1]);
  float d2c = (float)dadd(dadd((double)a[1][0], dmul((double)a[1][1], -2.0)), (double)a[1][2]);
  float mix = (float)dadd(dadd(dadd(dmul((double)a[0][0], 0.25), dmul((double)a[0][2], -0.25)),
                               dmul((double)a[2][0], -0.25)), dmul((double)a[2][2], 0.25));
  float two_i = fmul(2.0f, i1);
  float hc = fmul(hpar, cnt);
  float A11 = fsub(fmul(i1, fsub(d2r, two_i)), hc);
  float A22 = fsub(fmul(i1, fsub(d2c, two_i)), hc);
  float A12 = fmul(i1, mix);
  float det = fsub(fmul(A11, A22), fmul(A12, A12));
  c.B11 = fdiv(A22, det);
  c.B12 = fdiv(-A12, det);
  c.B22 = fdiv(A11, det);
  return c;
}
// one Jacobi-type sweep at one pixel.  u = ROW component, v = COLUMN component (adapter swap, LS:38-39).
// LsNb: the 8 neighbours with clamp-to-edge ('nearest') applied by the caller; h8u / h8v: the 8-neighbour sums with
// ZERO padding (the reference's H kernel uses mode='constant'), summed as ((nw+sw) + (n+s) + (ne+se)) + (w+e).
// Fast f32/FMA formulation; the power-of-two factors of the D / M stencils are folded into the coefficients, which is
// exact:   bu = 2 IIx Dr(u) + IIx Dc(v) + IIy Dr(v) + II Fr(u) + II Mx(v) + h H8(u) + Ixt          (LS:142-144)
//          bv = IIy Dr(u) + IIx Dc(u) + 2 IIy Dc(v) + II Mx(u) + II Fc(v) + h H8(v) + Iyt          (LS:146-148)
struct LsNb { float n, s, w, e, nw, ne, sw, se; };
OFRI_HD float ls_h8_cols(float vl, float vc, float vr, float w, float e) {   // vl = nw+sw, vc = n+s, vr = ne+se
  return fadd(fadd(fadd(vl, vc), vr), fadd(w, e));
}
// The stencil values of one pixel.  Written as differences / sums of COLUMN quantities (vertical difference s - n,
// centre-row samples), so a kernel that sweeps along a row forms each column quantity once and shares it between the
// three pixels it touches: 2 Dr = vd_c, 4 Mx = vd_e - vd_w with vd = s - n of the centre / east / west column.
struct LsSt { float dr_u, dc_u, dr_v, dc_v, fr_u, fc_v, mx_u, mx_v; };
OFRI_HD void ls_update3(const LsSt& t, float h8u, float h8v, const LsCoef& c, float hpar, float* un, float* vn) {
  float hx = fmul(0.5f, c.IIx), hy = fmul(0.5f, c.IIy), q = fmul(0.25f, c.II);
  float bu = fmul(c.IIx, t.dr_u);
  bu = fmaf(hx, t.dc_v, bu);
  bu = fmaf(hy, t.dr_v, bu);
  bu = fmaf(c.II, t.fr_u, bu);
  bu = fmaf(q, t.mx_v, bu);
  bu = fmaf(hpar, h8u, bu);
  bu = fadd(bu, c.Ixt);
  float bv = fmul(hy, t.dr_u);
  bv = fmaf(hx, t.dc_u, bv);
  bv = fmaf(c.IIy, t.dc_v, bv);
  bv = fmaf(q, t.mx_u, bv);
  bv = fmaf(c.II, t.fc_v, bv);
  bv = fmaf(hpar, h8v, bv);
  bv = fadd(bv, c.Iyt);
  *un = -fmaf(c.B11, bu, fmul(c.B12, bv));
  *vn = -fmaf(c.B12, bu, fmul(c.B22, bv));
}
OFRI_HD void ls_update2(const LsNb& u, const LsNb& v, float h8u, float h8v, const LsCoef& c, float hpar, float* un,
                        float* vn) {
  LsSt t;
  t.dr_u = fsub(u.s, u.n); t.dc_u = fsub(u.e, u.w); t.dr_v = fsub(v.s, v.n); t.dc_v = fsub(v.e, v.w);   // 2 Dr, 2 Dc
  t.fr_u = fadd(u.n, u.s); t.fc_v = fadd(v.w, v.e);
  t.mx_u = fsub(fsub(u.se, u.ne), fsub(u.sw, u.nw));                                                 // 4 Mx
  t.mx_v = fsub(fsub(v.se, v.ne), fsub(v.sw, v.nw));
  ls_update3(t, h8u, h8v, c, hpar, un, vn);
}
// convenience for per-pixel callers: 3x3 clamped neighbourhoods + in-bounds mask (bit 3r+c)
OFRI_HD void ls_update(const float uc[3][3], const float vc[3][3], unsigned inb, const LsCoef& c, float hpar,
                       float* un, float* vn) {
  LsNb u{uc[0][1], uc[2][1], uc[1][0], uc[1][2], uc[0][0], uc[0][2], uc[2][0], uc[2][2]};
  LsNb v{vc[0][1], vc[2][1], vc[1][0], vc[1][2], vc[0][0], vc[0][2], vc[2][0], vc[2][2]};
  float zu[3][3], zv[3][3];
  for (int r = 0; r < 3; ++r)
    for (int q = 0; q < 3; ++q) {
      bool in = (inb >> (3 * r + q)) & 1u;
      zu[r][q] = in ? uc[r][q] : 0.0f;
      zv[r][q] = in ? vc[r][q] : 0.0f;
    }
  float h8u = ls_h8_cols(fadd(zu[0][0], zu[2][0]), fadd(zu[0][1], zu[2][1]), fadd(zu[0][2], zu[2][2]), zu[1][0], zu[1][2]);
  float h8v = ls_h8_cols(fadd(zv[0][0], zv[2][0]), fadd(zv[0][1], zv[2][1]), fadd(zv[0][2], zv[2][2]), zv[1][0], zv[1][2]);
  ls_update2(u, v, h8u, h8v, c, hpar, un, vn);
}

}  // namespace ofri
