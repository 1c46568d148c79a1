/* CPU restatement (plain C) of the reference's dense Lucas-Kanade adapter -- TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may build or call this file (through
 * oracle/ofri_lk_oracle.py); the product (opticalflow_ri_b200) never does.
 *
 * Follows the OpenCL kernel `lkDense` of /root/reference/src/pyrlkDenseLargeW.cl (non-CPU branch, cited as CL:line) as
 * launched by denseLucasKanade_PyCL.py:113-169 (cited LK:line): one 8 x 8 work-group per pixel, every work-item owning
 * a 4 x 4 grid of window samples 8 apart, work-group sums by the kernel's fixed tree.
 *
 * PARITY PINNING: **parity unpinned**.  The reference's adapter needs an OpenCL device and its result depends on
 * device-defined arithmetic that no CPU restatement can pin: the image sampler's bilinear filter (CL:236, weight
 * precision is implementation-defined; NVIDIA / AMD hardware uses 8-bit fixed-point weights), `mad` (fused or not),
 * and the precision of `/`.  This restatement fixes them the way the CUDA path does: bilinear weights in full float32
 * with the OpenCL specification's formula (i0 = floor(x - 0.5), a = frac(x - 0.5), clamp-to-edge addressing),
 * T = (((1-a)(1-b)) T00 + (a (1-b)) T10) + ((1-a) b) T01) + (a b) T11 with separately rounded products and sums;
 * `mad` = one fused multiply-add; IEEE division.  Pinned instead: agreement with OpenCV's pyramidal LK at level 0
 * (tests/test_lk_cpu.py) and the algebra of a pure translation.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared ofri_lk_oracle.c -o _build/libofri_lk_oracle.so -lm
 */
#include <math.h>
#include <stdlib.h>

static float texel(const float* im, int H, int W, int y, int x) {
  if (x < 0) x = 0;
  if (x > W - 1) x = W - 1;
  if (y < 0) y = 0;
  if (y > H - 1) y = H - 1;
  return im[(long)y * W + x];
}

/* read_imagef(.., CLK_NORMALIZED_COORDS_FALSE | CLK_ADDRESS_CLAMP_TO_EDGE | CLK_FILTER_LINEAR, (x, y)) (CL:236) */
static float sample_linear(const float* im, int H, int W, float x, float y) {
  float fx = x - 0.5f, fy = y - 0.5f;
  float ix = floorf(fx), iy = floorf(fy);
  float a = fx - ix, b = fy - iy;
  int x0 = (int)ix, y0 = (int)iy;
  float t00 = texel(im, H, W, y0, x0), t10 = texel(im, H, W, y0, x0 + 1);
  float t01 = texel(im, H, W, y0 + 1, x0), t11 = texel(im, H, W, y0 + 1, x0 + 1);
  float oa = 1.0f - a, ob = 1.0f - b;
  float r = (oa * ob) * t00;
  r = r + (a * ob) * t10;
  r = r + (oa * b) * t01;
  r = r + (a * b) * t11;
  return r;
}

/* the kernel's work-group sum over 64 work-items (CL:113-155): s[t] += s[t+32], +16, +8, +4, then (s0+s1)+(s2+s3) */
static float group_sum(float* s) {
  for (int half = 32; half >= 4; half >>= 1)
    for (int t = 0; t < half; ++t) s[t] = s[t] + s[t + half];
  return (s[0] + s[1]) + (s[2] + s[3]);
}

/* per-axis weights of the four grid columns / rows of work-item `id` (CL:339-372): win = window extent, lo / hi = the
 * asymmetric-window switches of that axis (left / right or top / bottom) */
static void axis_weights(int id, int win, int lo, int hi, float w[4]) {
  w[0] = 1.0f;
  if (win >= 16) {                                  /* WSX / WSY = 1 (LK:57-62) */
    w[1] = 1.0f;
    w[2] = (16 + id < win - hi) ? 1.0f : 0.0f;
    w[3] = (24 + id < win - hi) ? 1.0f : 0.0f;
    if (id == 0) w[1] = (float)(1 - lo);
  } else {
    w[1] = (8 + id < win - hi) ? 1.0f : 0.0f;
    if (id == 0) w[1] = (float)(1 - lo);
    w[2] = 0.0f;
    w[3] = 0.0f;
  }
}

/* one frame pair; u, v: initial flow in, refined flow out (in place, like the kernel's u / v buffers).
 * asym = {left, right, top, bottom} (LK:75-92).  Returns 0. */
int ofri_lk_oracle(const float* I, const float* J, float* u, float* v, int rows, int cols, int half_window, int iters,
                   const int* asym) {
  const int win = 2 * half_window + 1;
  const float hw = (float)((win - 1) >> 1);         /* c_halfWin (CL:374) */
  float wxs[8][4], wys[8][4];
  for (int id = 0; id < 8; ++id) {
    axis_weights(id, win, asym[0], asym[1], wxs[id]);
    axis_weights(id, win, asym[2], asym[3], wys[id]);
  }
  float* P = (float*)malloc(sizeof(float) * 64 * 16 * 3);   /* per work-item: patch value, Dx, Dy of its 16 samples */
  if (!P) return -1;
  float* Pv = P, *Px = P + 64 * 16, *Py = P + 2 * 64 * 16;
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < cols; ++j) {
      const long gid = (long)i * cols + j;
      const int px = j - (int)hw, py = i - (int)hw;  /* prevPt - c_halfWin: integers, exact in float */
      float s1[64], s2[64], s3[64];
      for (int yid = 0; yid < 8; ++yid)
        for (int xid = 0; xid < 8; ++xid) {
          const int tid = yid * 8 + xid;
          float A11 = 0.0f, A12 = 0.0f, A22 = 0.0f;
          for (int ty = 0; ty < 4; ++ty)
            for (int tx = 0; tx < 4; ++tx) {
              /* local patch entry (y, x) holds I at (py + y - 1, px + x - 1), clamped (CL:249-277: the sampler is hit
               * at texel centres, so the linear filter returns the texel itself) */
              const int y = py + ty * 8 + yid, x = px + tx * 8 + xid;
              const float w = wys[yid][ty] * wxs[xid][tx];
#define VAL(dy, dx) texel(I, rows, cols, y + (dy), x + (dx))
              float sx = VAL(-1, 1) + VAL(1, 1) - VAL(-1, -1) - VAL(1, -1);
              float sy = VAL(1, -1) + VAL(1, 1) - VAL(-1, -1) - VAL(-1, 1);
              float dx = fmaf(sx, 3.0f, (VAL(0, 1) - VAL(0, -1)) * 10.0f) * w;     /* CL:221 */
              float dy = fmaf(sy, 3.0f, (VAL(1, 0) - VAL(-1, 0)) * 10.0f) * w;     /* CL:222 */
              const int e = tid * 16 + ty * 4 + tx;
              Pv[e] = VAL(0, 0);
#undef VAL
              Px[e] = dx;
              Py[e] = dy;
              A11 = fmaf(dx, dx, A11);
              A12 = fmaf(dx, dy, A12);
              A22 = fmaf(dy, dy, A22);
            }
          s1[tid] = A11; s2[tid] = A12; s3[tid] = A22;
        }
      float A11 = group_sum(s1), A12 = group_sum(s2), A22 = group_sum(s3);
      float D = fmaf(A11, A22, -(A12 * A12));         /* CL:476 */
      if (D < 1.192092896e-07f) continue;              /* flow of this pixel stays as it came in (CL:478-484) */
      A11 /= D; A12 /= D; A22 /= D;
      float ppx = ((float)j + u[gid]) - hw, ppy = ((float)i + v[gid]) - hw;     /* CL:489 */
      float lx[8][4], ly[8][4];
      for (int id = 0; id < 8; ++id) {
        lx[id][0] = ppx + ((float)id + 0.5f);
        ly[id][0] = ppy + ((float)id + 0.5f);
        for (int t = 1; t < 4; ++t) { lx[id][t] = lx[id][t - 1] + 8.0f; ly[id][t] = ly[id][t - 1] + 8.0f; }
      }
      for (int k = 0; k < iters; ++k) {
        if (ppx < -hw || ppx >= (float)cols || ppy < -hw || ppy >= (float)rows) break;   /* CL:502 */
        for (int yid = 0; yid < 8; ++yid)
          for (int xid = 0; xid < 8; ++xid) {
            const int tid = yid * 8 + xid;
            float b1 = 0.0f, b2 = 0.0f;
            for (int ty = 0; ty < 4; ++ty)
              for (int tx = 0; tx < 4; ++tx) {
                const int e = tid * 16 + ty * 4 + tx;
                const float w = wys[yid][ty] * wxs[xid][tx];
                float diff = (sample_linear(J, rows, cols, lx[xid][tx], ly[yid][ty]) - Pv[e]) * w;   /* CL:234-238 */
                b1 = fmaf(diff, Px[e], b1);
                b2 = fmaf(diff, Py[e], b2);
              }
            s1[tid] = b1; s2[tid] = b2;
          }
        float b1 = group_sum(s1), b2 = group_sum(s2);
        float ddx = fmaf(A12, b2, -(A22 * b1)) * 32.0f;      /* CL:592-593 */
        float ddy = fmaf(A12, b1, -(A11 * b2)) * 32.0f;
        ppx += ddx; ppy += ddy;
        for (int id = 0; id < 8; ++id)
          for (int t = 0; t < 4; ++t) { lx[id][t] += ddx; ly[id][t] += ddy; }
        if (fabsf(ddx) < 0.01f && fabsf(ddy) < 0.01f) break;
      }
      ppx += hw; ppy += hw;                                   /* CL:658-661 */
      u[gid] = ppx - (float)j;
      v[gid] = ppy - (float)i;
    }
  free(P);
  return 0;
}
