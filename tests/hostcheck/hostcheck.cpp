// hostcheck.cpp -- TEST-ONLY host harness.  Compiles the per-pixel arithmetic of the CUDA kernels
// (opticalflow_ri_b200/csrc/ofri_pixel.cuh, ofri_tables.h -- the very same source the kernels include) with g++
// and drives it with plain loops, so the index rules / rounding order can be checked against the golden vectors
// on a box without a GPU.  It is built by tests/test_hostcheck.py into tests/hostcheck/_build/ and is never
// loaded by the product: libofri.so has no CPU path.
#include <cmath>
#include <cstring>
#include <vector>
#include "../../opticalflow_ri_b200/csrc/ofri_pixel.cuh"
#include "../../opticalflow_ri_b200/csrc/ofri_tables.h"
#include "../../opticalflow_ri_b200/csrc/ofri_spline.cuh"

using namespace ofri;

struct Taps { const float* k; float operator[](int i) const { return k[i]; } };

// mirrors spline_solve_kernel: y[e*ye] (float or double) -> M[e*me]
template <typename T>
static void solve_line(const T* y, long ye, double* M, long me, int n, const HostSplineSys& s) {
  const int m = n - 2;
  double y0 = (double)y[0], y1 = (double)y[ye], y2, dp = 0.0;
  for (int i = 0; i < m; ++i) {
    y2 = (double)y[(long)(i + 2) * ye];
    double rhs = dmul(6.0, dadd(dsub(y0, dmul(2.0, y1)), y2));
    if (i == 0) dp = ddiv(rhs, s.den[0]);
    else dp = ddiv(dsub(rhs, dmul(s.lo[i], dp)), s.den[i]);
    M[(long)(i + 1) * me] = dp;
    y0 = y1; y1 = y2;
  }
  double next = M[(long)m * me];
  for (int i = m - 2; i >= 0; --i) {
    double v = dsub(M[(long)(i + 1) * me], dmul(s.cp[i], next));
    M[(long)(i + 1) * me] = v;
    next = v;
  }
  M[0] = dsub(dmul(2.0, M[me]), M[2 * me]);
  M[(long)(n - 1) * me] = dsub(dmul(2.0, M[(long)(n - 2) * me]), M[(long)(n - 3) * me]);
}
// mirrors the chunk-parallel kernels (spline_cols_fwd/bwd_kernel, spline_rows_kernel): the unknowns M[a .. b] of one
// line by windowed chunks of C rows (ofri_spline.cuh); D = forward-elimination scratch, same indexing as M
static SplineSysView sys_view(const HostSplineSys& s, int n) {
  SplineSysView v;
  const int m = n - 2;
  v.lo = s.lo.data(); v.cp = s.cp.data(); v.den = s.den.data();
  v.n = n; v.conv = s.conv;
  v.den_c = s.den[s.conv < m ? s.conv : m - 1];
  v.cp_c = s.cp[s.conv < m ? s.conv : m - 1];
  v.rcp_c = 1.0 / v.den_c;
  return v;
}
template <typename T>
static void solve_line_win(const T* y, long ye, double* D, double* M, long me, int n, int a, int b, int C, int Wm,
                           const SplineSysView& sv) {
  const int m = n - 2;
  const SplineWindow w = spline_window(a, b, n, sv.conv, Wm);
  const int nch = spline_num_chunks(w, C);
  for (int c = 0; c < nch; ++c) {
    SplineChunk k = spline_chunk(w, c, C, sv.conv, Wm);
    spline_chunk_forward(k, m, sv, [&](int e) { return (double)y[(long)e * ye]; },
                         [&](int e) -> double& { return D[(long)e * me]; });
  }
  for (int c = 0; c < nch; ++c) {
    SplineChunk k = spline_chunk(w, c, C, sv.conv, Wm);
    if (k.ra > w.RB) continue;
    auto Dr = [&](int e) { return D[(long)e * me]; };
    double next = spline_chunk_tail(k, w, m, Wm, sv, Dr);
    spline_chunk_back(k, n, sv, next, Dr, [&](int e) -> double& { return M[(long)e * me]; });
  }
}

extern "C" {

// full-line sequential solve and windowed / chunked solve of the unknowns [a, b] (others left untouched), f64 in/out
void hc_spline_line(const double* y, int n, double* M) {
  HostSplineSys s = build_spline_sys(n);
  solve_line<double>(y, 1, M, 1, n, s);
}
void hc_spline_line_win(const double* y, int n, int a, int b, int C, int Wm, double* M) {
  HostSplineSys s = build_spline_sys(n);
  std::vector<double> D(n, 0.0);
  solve_line_win<double>(y, 1, D.data(), M, 1, n, a, b, C, Wm, sys_view(s, n));
}
// coarse rows [*lo, *hi) a band must hold to up-sample output rows [row0, row0 + rows) of H (launch_spline's rule)
void hc_spline_rows_needed(int row0, int rows, int h, int H, int Wm, int* lo, int* hi) {
  HostSplineSys sy = build_spline_sys(h);
  int ia, ib;
  double sf;
  spline_locate(row0, h, H, &ia, &sf);
  spline_locate(row0 + rows - 1, h, H, &ib, &sf);
  const SplineWindow w = spline_window(ia, ib + 1, h, sy.conv, Wm);
  *lo = w.FS;
  *hi = w.RF + 3;
}
// rows [row0, row0 + rows) of the up-sampled plane by the windowed algorithm (what a row band computes)
void hc_spline_win(const float* in, int h, int w, int H, int W, float mul, int row0, int rows, int Cy, int Cx, int Wm,
                   float* out) {
  HostSplineSys sy = build_spline_sys(h), sx = build_spline_sys(w);
  const SplineSysView vy = sys_view(sy, h), vx = sys_view(sx, w);
  int ia, ib;
  double sf;
  spline_locate(row0, h, H, &ia, &sf);
  spline_locate(row0 + rows - 1, h, H, &ib, &sf);
  std::vector<double> M1((size_t)h * w, NAN), D1((size_t)h * w, NAN), T(w), D2(w), M2(w);
  for (int x = 0; x < w; ++x) solve_line_win<float>(in + x, w, D1.data() + x, M1.data() + x, w, h, ia, ib + 1, Cy, Wm, vy);
  for (int k = 0; k < rows; ++k) {
    int i; double s;
    spline_locate(k + row0, h, H, &i, &s);
    const SplinePos pos = spline_pos(s);
    for (int x = 0; x < w; ++x)
      T[x] = spline_eval_at((double)in[(size_t)i * w + x], (double)in[(size_t)(i + 1) * w + x], M1[(size_t)i * w + x],
                            M1[(size_t)(i + 1) * w + x], pos);
    solve_line_win<double>(T.data(), 1, D2.data(), M2.data(), 1, w, 0, w - 1, Cx, Wm, vx);
    for (int l = 0; l < W; ++l) {
      spline_locate(l, w, W, &i, &s);
      float r = (float)spline_eval(T[i], T[i + 1], M2[i], M2[i + 1], s);
      if (mul != 1.0f) r = fmul(r, mul);
      out[(size_t)k * W + l] = r;
    }
  }
}

void hc_gauss(const float* in, int H, int W, const float* taps, int K, float* out) {
  std::vector<float> tmp((size_t)H * W);
  Taps t{taps};
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) tmp[(size_t)y * W + x] = gauss_point(in + (size_t)y * W, 1, x, W, t, K);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) out[(size_t)y * W + x] = gauss_point(tmp.data() + x, W, y, H, t, K);
}

void hc_resize(const float* in, int H, int W, int oh, int ow, float* out) {
  HostResizeTaps tx = build_resize_taps(W, ow), ty = build_resize_taps(H, oh);
  std::vector<float> tmp((size_t)H * ow);
  for (int y = 0; y < H; ++y)
    for (int ox = 0; ox < ow; ++ox)
      tmp[(size_t)y * ow + ox] = resample_point(in + (size_t)y * W, 1, tx.xmin[ox], tx.cnt[ox], &tx.w[(size_t)ox * tx.kmax]);
  for (int oy = 0; oy < oh; ++oy)
    for (int x = 0; x < ow; ++x)
      out[(size_t)oy * ow + x] = resample_point(tmp.data() + x, ow, ty.xmin[oy], ty.cnt[oy], &ty.w[(size_t)oy * ty.kmax]);
}

int hc_level_size(int n, double scale) { return level_size_half_even(n, scale); }

void hc_spline(const float* in, int h, int w, int H, int W, float mul, float* out) {
  HostSplineSys sy = build_spline_sys(h), sx = build_spline_sys(w);
  std::vector<double> M1((size_t)h * w), T1((size_t)H * w), M2((size_t)H * w);
  for (int x = 0; x < w; ++x) solve_line<float>(in + x, w, M1.data() + x, w, h, sy);
  for (int k = 0; k < H; ++k)
    for (int x = 0; x < w; ++x) {
      int i; double s;
      spline_locate(k, h, H, &i, &s);
      T1[(size_t)k * w + x] = spline_eval((double)in[(size_t)i * w + x], (double)in[(size_t)(i + 1) * w + x],
                                          M1[(size_t)i * w + x], M1[(size_t)(i + 1) * w + x], s);
    }
  for (int k = 0; k < H; ++k) solve_line<double>(T1.data() + (size_t)k * w, 1, M2.data() + (size_t)k * w, 1, w, sx);
  for (int k = 0; k < H; ++k)
    for (int l = 0; l < W; ++l) {
      int i; double s;
      spline_locate(l, w, W, &i, &s);
      const double* tp = T1.data() + (size_t)k * w;
      const double* mp = M2.data() + (size_t)k * w;
      float r = (float)spline_eval(tp[i], tp[i + 1], mp[i], mp[i + 1], s);
      if (mul != 1.0f) r = fmul(r, mul);
      out[(size_t)k * W + l] = r;
    }
}

void hc_warp_coords(const float* img, const float* cy, const float* cx, int H, int W, float* out) {
  for (int i = 0; i < H * W; ++i) out[i] = warp_sample(img, W, H, W, cy[i], cx[i]);
}
void hc_warp_pair(const float* im1, const float* im2, const float* us, const float* vs, int H, int W, float* o1,
                  float* o2) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float u = us[(size_t)y * W + x], v = vs[(size_t)y * W + x];
      o1[(size_t)y * W + x] = warp_sample(im1, W, H, W, warp_coord(y, v, -1.f), warp_coord(x, u, -1.f));
      o2[(size_t)y * W + x] = warp_sample(im2, W, H, W, warp_coord(y, v, +1.f), warp_coord(x, u, +1.f));
    }
}

void hc_hs_derivs(const float* A, const float* B, int H, int W, float* fx, float* fy, float* ft) {
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int x1 = mirror1(x + 1, W), y1 = mirror1(y + 1, H);
      hs_deriv_point(A[(size_t)y * W + x], A[(size_t)y * W + x1], A[(size_t)y1 * W + x], A[(size_t)y1 * W + x1],
                     B[(size_t)y * W + x], B[(size_t)y * W + x1], B[(size_t)y1 * W + x], B[(size_t)y1 * W + x1],
                     fx + (size_t)y * W + x, fy + (size_t)y * W + x, ft + (size_t)y * W + x);
    }
}
void hc_hs_iterate_precise(const float* u0, const float* v0, const float* fx, const float* fy, const float* ft, int H,
                           int W, float alpha, int niter, float* uo, float* vo) {
  std::vector<float> U(u0, u0 + (size_t)H * W), V(v0, v0 + (size_t)H * W), Un((size_t)H * W), Vn((size_t)H * W);
  const float a2 = alpha * alpha;
  for (int it = 0; it < niter; ++it) {
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        int xl = mirror1(x - 1, W), xr = mirror1(x + 1, W), yu = mirror1(y - 1, H), yd = mirror1(y + 1, H);
        size_t ru = (size_t)yu * W, rm = (size_t)y * W, rd = (size_t)yd * W;
        float ua = hs_avg_cols_precise(dadd((double)U[ru + xl], (double)U[rd + xl]), dadd((double)U[ru + x], (double)U[rd + x]),
                                       dadd((double)U[ru + xr], (double)U[rd + xr]), (double)U[rm + xl], (double)U[rm + xr]);
        float va = hs_avg_cols_precise(dadd((double)V[ru + xl], (double)V[rd + xl]), dadd((double)V[ru + x], (double)V[rd + x]),
                                       dadd((double)V[ru + xr], (double)V[rd + xr]), (double)V[rm + xl], (double)V[rm + xr]);
        float dx = fx[rm + x], dy = fy[rm + x], dt = ft[rm + x];
        float den = hs_den(dx, dy, a2);
        hs_update_precise(ua, va, dx, dy, dt, den, rcp_rn(den), &Un[rm + x], &Vn[rm + x]);
      }
    U.swap(Un);
    V.swap(Vn);
  }
  memcpy(uo, U.data(), sizeof(float) * H * W);
  memcpy(vo, V.data(), sizeof(float) * H * W);
}
void hc_hs_iterate(const float* u0, const float* v0, const float* fx, const float* fy, const float* ft, int H, int W,
                   float alpha, int niter, float* uo, float* vo) {
  std::vector<float> U(u0, u0 + (size_t)H * W), V(v0, v0 + (size_t)H * W), Un((size_t)H * W), Vn((size_t)H * W);
  const float a2 = alpha * alpha;
  for (int it = 0; it < niter; ++it) {
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        int xl = mirror1(x - 1, W), xr = mirror1(x + 1, W), yu = mirror1(y - 1, H), yd = mirror1(y + 1, H);
        size_t ru = (size_t)yu * W, rm = (size_t)y * W, rd = (size_t)yd * W;
        float ua = hs_avg_cols(fadd(U[ru + xl], U[rd + xl]), fadd(U[ru + x], U[rd + x]), fadd(U[ru + xr], U[rd + xr]),
                               U[rm + xl], U[rm + xr]);
        float va = hs_avg_cols(fadd(V[ru + xl], V[rd + xl]), fadd(V[ru + x], V[rd + x]), fadd(V[ru + xr], V[rd + xr]),
                               V[rm + xl], V[rm + xr]);
        float dx = fx[rm + x], dy = fy[rm + x], dt = ft[rm + x];
        float ca, cb, cc;
        hs_normalise(dx, dy, dt, a2, &ca, &cb, &cc);
        hs_update_n(ua, va, ca, cb, cc, &Un[rm + x], &Vn[rm + x]);
      }
    U.swap(Un);
    V.swap(Vn);
  }
  memcpy(uo, U.data(), sizeof(float) * H * W);
  memcpy(vo, V.data(), sizeof(float) * H * W);
}

// coef: [8][H][W]
void hc_ls_coef(const float* im1, const float* im2, int H, int W, float hpar, float* coef) {
  float m1 = -INFINITY, m2 = -INFINITY;
  for (int i = 0; i < H * W; ++i) { m1 = fmaxf(m1, im1[i]); m2 = fmaxf(m2, im2[i]); }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float a[3][3], d[3][3];
      int cnt = 0;
      for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) {
          int yy = y + r - 1, xx = x + q - 1;
          bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
          if (in && !(r == 1 && q == 1)) ++cnt;
          yy = clampi(yy, 0, H - 1); xx = clampi(xx, 0, W - 1);
          float i1 = fdiv(im1[(size_t)yy * W + xx], m1), i2 = fdiv(im2[(size_t)yy * W + xx], m2);
          a[r][q] = i1; d[r][q] = fsub(i2, i1);
        }
      LsCoef c = ls_coef_point(a, d, hpar, (float)cnt);
      size_t o = (size_t)y * W + x, P = (size_t)H * W;
      coef[o] = c.IIx; coef[P + o] = c.IIy; coef[2 * P + o] = c.II; coef[3 * P + o] = c.Ixt; coef[4 * P + o] = c.Iyt;
      coef[5 * P + o] = c.B11; coef[6 * P + o] = c.B12; coef[7 * P + o] = c.B22;
    }
}
// u = ROW component, v = COLUMN component; returns sweeps run, *err = last total_error
int hc_ls_iterate(const float* u0, const float* v0, const float* coef, int H, int W, float hpar, int maxiter, double tol,
                  float* uo, float* vo, double* err) {
  size_t P = (size_t)H * W;
  std::vector<float> U(u0, u0 + P), V(v0, v0 + P), Un(P), Vn(P);
  double te = 1e8;
  int k = 0;
  while (te > tol && k < maxiter) {
    double su = 0, sv = 0;
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        float uc[3][3], vc[3][3];
        unsigned inb = 0;
        for (int r = 0; r < 3; ++r)
          for (int q = 0; q < 3; ++q) {
            int yy = y + r - 1, xx = x + q - 1;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) inb |= 1u << (3 * r + q);
            yy = clampi(yy, 0, H - 1); xx = clampi(xx, 0, W - 1);
            uc[r][q] = U[(size_t)yy * W + xx]; vc[r][q] = V[(size_t)yy * W + xx];
          }
        size_t o = (size_t)y * W + x;
        LsCoef c{coef[o], coef[P + o], coef[2 * P + o], coef[3 * P + o], coef[4 * P + o], coef[5 * P + o],
                 coef[6 * P + o], coef[7 * P + o]};
        ls_update(uc, vc, inb, c, hpar, &Un[o], &Vn[o]);
        float eu = fsub(Un[o], uc[1][1]), ev = fsub(Vn[o], vc[1][1]);
        su += (double)eu * eu; sv += (double)ev * ev;
      }
    te = ((double)(float)std::sqrt(su) + (double)(float)std::sqrt(sv)) / ((double)H * W);
    U.swap(Un); V.swap(Vn);
    ++k;
  }
  memcpy(uo, U.data(), sizeof(float) * P);
  memcpy(vo, V.data(), sizeof(float) * P);
  *err = te;
  return k;
}

}  // extern "C"
