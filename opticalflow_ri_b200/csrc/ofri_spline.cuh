// ofri_spline.cuh -- WINDOWED Thomas solve of the not-a-knot spline system (RectBivariateSpline restated, SURVEY A.3;
// reference call site GenericPyramidalOpticalFlow.py:155-162).  Shared by the spline kernels (ofri_stages.cu) and the
// test-only host harness (tests/hostcheck), so the recurrences are written once.
//
// The reduced system has rows i = 0 .. m-1 (m = n-2), row i <-> unknown M[i+1]:
//   forward   dp_i = (rhs_i - lo_i dp_{i-1}) / den_i,   rhs_i = 6 (y_i - 2 y_{i+1} + y_{i+2})
//   backward  M[m] = dp_{m-1};  M[i+1] = dp_i - cp_i M[i+2]
//   ends      M[0] = 2 M[1] - M[2],  M[n-1] = 2 M[n-2] - M[n-3]
// After a short head (SplineSys::conv rows) every interior row has the same constants (lo = 1, den_c, cp_c = 0.268),
// so both recurrences FORGET their start at the rate cp_c^k: a forward sweep started Wm rows early from dp = 0, or a
// back substitution started Wm rows late from M = dp, reproduces the values of the full-line solve to cp_c^Wm relative
// (Wm = 48: 3e-28, twelve orders below the float64 rounding of a single step, so the stored doubles are the SAME bit
// patterns as those of the sequential solve except for ~1e-12 of them, and those differ by one ulp).  That turns the
// sequential line solve into independent CHUNKS (parallelism inside a line) and lets a row band solve only the WINDOW
// of a column it needs from its own rows + a halo (no all-gather, no redundant full-column solve).
#pragma once
#include "ofri_pixel.cuh"

namespace ofri {

constexpr int kSplineWarm = 48;

// constants of one system size as the kernels see them (device pointers on the device, host pointers in hostcheck)
struct SplineSysView {
  const double* lo = nullptr;
  const double* cp = nullptr;
  const double* den = nullptr;
  int n = 0;
  int conv = 0;              // rows [conv, n-3) of the reduced system share (den_c, cp_c) and lo == 1
  double den_c = 0.0, cp_c = 0.0, rcp_c = 0.0;   // rcp_c = RN(1 / den_c)
};

OFRI_HD double spline_rhs(double y0, double y1, double y2) { return dmul(6.0, dadd(dsub(y0, dmul(2.0, y1)), y2)); }
// one forward-elimination step of row i (dp = dp_{i-1}; ignored for i == 0)
OFRI_HD double spline_fwd_step(int i, int m, double rhs, double dp, const SplineSysView& s) {
  if (i >= s.conv && i < m - 1) {            // steady state: lo == 1, division by the constant (correctly rounded)
    const double t = dsub(rhs, dp);
    return ddiv_const(t, s.den_c, s.rcp_c);
  }
  if (i == 0) return ddiv(rhs, s.den[0]);
  return ddiv(dsub(rhs, dmul(s.lo[i], dp)), s.den[i]);
}
OFRI_HD double spline_cp(int i, int m, const SplineSysView& s) { return (i >= s.conv && i < m - 1) ? s.cp_c : s.cp[i]; }

// Rows of the reduced system needed for the unknowns M[a .. b] (0 <= a <= b <= n-1):
//   RA..RB  rows whose unknowns are wanted;  RF >= RB  last row of the forward sweep (back-substitution warm-up);
//   FS <= RA  first row of the forward sweep (0 = exact start);  samples y[FS .. RF+2] are read.
struct SplineWindow { int RA, RB, RF, FS; };
OFRI_HD SplineWindow spline_window(int a, int b, int n, int conv, int Wm) {
  const int m = n - 2;
  SplineWindow w;
  w.RA = a - 1 < 0 ? 0 : a - 1;
  w.RB = b - 1 > m - 1 ? m - 1 : b - 1;
  if (a == 0 && w.RB < 1) w.RB = m - 1 < 1 ? m - 1 : 1;          // M[0] needs M[1], M[2]
  if (b == n - 1 && w.RA > m - 2) w.RA = m - 2 < 0 ? 0 : m - 2;  // M[n-1] needs M[n-2], M[n-3]
  if (w.RB < w.RA) w.RB = w.RA;
  w.RF = w.RB + Wm;
  if (w.RF >= m - 2) w.RF = m - 1;
  w.FS = w.RA - Wm;
  if (w.FS <= conv) w.FS = 0;
  return w;
}
// chunk c (size C) of the forward range [RA, RF]: own rows [ra, rb], forward start fs (with warm-up)
struct SplineChunk { int ra, rb, fs; };
OFRI_HD SplineChunk spline_chunk(const SplineWindow& w, int c, int C, int conv, int Wm) {
  SplineChunk k;
  k.ra = w.RA + c * C;
  k.rb = k.ra + C - 1 > w.RF ? w.RF : k.ra + C - 1;
  k.fs = k.ra - Wm;
  if (k.fs <= conv) k.fs = 0;
  return k;
}
OFRI_HD int spline_num_chunks(const SplineWindow& w, int C) { return (w.RF - w.RA + C) / C; }

// ---- the three passes of one chunk.  Y(e) -> sample e as double;  D(e) -> reference to the f64 slot of unknown e.
// Pass 1: forward sweep fs .. rb; dp of the chunk's own rows is stored in D(i+1).
template <class YF, class DF>
OFRI_HD void spline_chunk_forward(const SplineChunk& k, int m, const SplineSysView& s, YF Y, DF D) {
  double y0 = Y(k.fs), y1 = Y(k.fs + 1), dp = 0.0;
  for (int i = k.fs; i <= k.rb; ++i) {
    const double y2 = Y(i + 2);
    dp = spline_fwd_step(i, m, spline_rhs(y0, y1, y2), dp, s);
    if (i >= k.ra) D(i + 1) = dp;
    y0 = y1;
    y1 = y2;
  }
}
// Pass 2: back substitution through the Wm rows BEHIND the chunk (other chunks' dp, read only) -> M of row rb+1, i.e.
// the `next` of the chunk's own back substitution.  Chunks that own the last row of the system return dp_{m-1}.
template <class DF>
OFRI_HD double spline_chunk_tail(const SplineChunk& k, const SplineWindow& w, int m, int Wm, const SplineSysView& s, DF D) {
  if (k.rb >= m - 1) return D(m);
  int e = k.rb + Wm;
  if (e > w.RF) e = w.RF;                  // RF = m-1 (exact) or RB + Wm
  double next = D(e + 1);
  for (int i = e - 1; i > k.rb; --i) next = dsub(D(i + 1), dmul(spline_cp(i, m, s), next));
  return next;
}
// Pass 3: the chunk's own rows: dp read from D(i+1), M written to Mo(i+1) (Mo may alias D: in place), plus the two
// eliminated end unknowns.  `next` = result of pass 2.  In place only after EVERY chunk of the line finished pass 2.
template <class DF, class MF>
OFRI_HD void spline_chunk_back(const SplineChunk& k, int n, const SplineSysView& s, double next, DF D, MF Mo) {
  const int m = n - 2;
  double prev = next;                       // M of row i+1 while handling row i
  for (int i = k.rb; i >= k.ra; --i) {
    double v;
    if (i == m - 1) {
      v = D(m);                             // M[m] = dp_{m-1}
      Mo(m) = v;
    } else {
      v = dsub(D(i + 1), dmul(spline_cp(i, m, s), next));
      Mo(i + 1) = v;
      if (i == m - 2) Mo(n - 1) = dsub(dmul(2.0, next), v);    // M[n-1] = 2 M[n-2] - M[n-3]
    }
    prev = next;
    next = v;
  }
  if (k.ra == 0) Mo(0) = dsub(dmul(2.0, next), prev);          // M[0] = 2 M[1] - M[2]  (chunks hold >= 2 rows)
}

}  // namespace ofri
