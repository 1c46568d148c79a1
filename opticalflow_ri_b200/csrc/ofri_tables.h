// ofri_tables.h -- host-side (plain C++) builders of the per-size constant tables the stage kernels read:
// Pillow's bicubic tap windows and the Thomas-algorithm constants of the not-a-knot spline system.
// Shared by ofri_api.cu and the test-only host harness (tests/hostcheck).
#pragma once
#include <cmath>
#include <vector>
#include "ofri_pixel.cuh"

namespace ofri {

struct HostResizeTaps {
  int kmax = 0;
  std::vector<int> xmin, cnt;
  std::vector<double> w;   // [out][kmax]
};
// Pillow precompute_coeffs (libImaging/Resample.c), antialiased: the BICUBIC filter (support 2; the driver's
// down-sampling, GPOF:67-68) or, bilinear = true, the BILINEAR one (support 1; Farneback_PyCL.py:61-62, both directions)
inline HostResizeTaps build_resize_taps(int in_size, int out_size, bool bilinear = false) {
  HostResizeTaps t;
  const double scale = (double)in_size / (double)out_size;
  const double fs = scale < 1.0 ? 1.0 : scale;
  const double support = (bilinear ? 1.0 : 2.0) * fs;
  t.kmax = (int)std::ceil(support) * 2 + 1;
  t.xmin.assign(out_size, 0);
  t.cnt.assign(out_size, 0);
  t.w.assign((size_t)out_size * t.kmax, 0.0);
  const double ss = 1.0 / fs;
  for (int i = 0; i < out_size; ++i) {
    double center = (i + 0.5) * scale;
    int lo = (int)(center - support + 0.5);
    if (lo < 0) lo = 0;
    int hi = (int)(center + support + 0.5);
    if (hi > in_size) hi = in_size;
    int n = hi - lo;
    double tot = 0.0;
    double* k = &t.w[(size_t)i * t.kmax];
    for (int x = 0; x < n; ++x) {
      const double arg = (x + lo - center + 0.5) * ss;
      double v = bilinear ? (std::fabs(arg) < 1.0 ? 1.0 - std::fabs(arg) : 0.0) : bicubic_filter(arg);
      k[x] = v;
      tot += v;
    }
    for (int x = 0; x < n; ++x)
      if (tot != 0.0) k[x] /= tot;
    t.xmin[i] = lo;
    t.cnt[i] = n;
  }
  return t;
}

struct HostSplineSys {
  std::vector<double> lo, cp, den;
  int conv = 0;   // first index from which den[i], cp[i] are bitwise constant (and lo[i] == 1) up to index m-2
};
// not-a-knot system in the unknowns M_1..M_{n-2} (M_0, M_{n-1} eliminated): diag 4 (6 at both ends), off-diagonals 1
// (0 next to the ends); forward-elimination constants cp, den
inline HostSplineSys build_spline_sys(int n) {
  const int m = n - 2;
  HostSplineSys s;
  std::vector<double> di(m, 4.0), up(m, 1.0);
  s.lo.assign(m, 1.0);
  s.cp.assign(m, 0.0);
  s.den.assign(m, 0.0);
  di[0] = 6.0; up[0] = 0.0;
  di[m - 1] = 6.0; s.lo[m - 1] = 0.0;
  s.den[0] = di[0];
  s.cp[0] = up[0] / di[0];
  for (int i = 1; i < m; ++i) {
    s.den[i] = dsub(di[i], dmul(s.lo[i], s.cp[i - 1]));
    s.cp[i] = up[i] / s.den[i];
  }
  // den_i = 4 - 1/den_{i-1} converges (ratio 0.072) to a floating-point fixed point: from `conv` on the interior rows
  // share ONE (den, cp) pair, which the solve kernel keeps in registers.  The last row (i = m-1) is always special.
  s.conv = m - 1;
  for (int i = 1; i + 1 < m - 1; ++i)
    if (s.den[i] == s.den[i + 1] && s.cp[i] == s.cp[i + 1]) {
      bool all = true;
      for (int j = i; j < m - 1; ++j) all = all && s.den[j] == s.den[i] && s.cp[j] == s.cp[i] && s.lo[j] == 1.0;
      if (all) { s.conv = i; break; }
    }
  return s;
}

inline int level_size_half_even(int n, double scale) { return (int)std::nearbyint((double)n * scale); }

}  // namespace ofri
