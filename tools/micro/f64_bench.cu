// Microbenchmark: throughput of DADD / DFMA / F2F.F64.F32 / F2F.F32.F64 / integer widen (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(float* out, int iters, float s) {
  double d[8]; float f[8];
  for (int i = 0; i < 8; ++i) { d[i] = threadIdx.x * 0.001 + i; f[i] = threadIdx.x * 0.001f + i; }
  double ds = s;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { _Pragma("unroll") for (int i = 0; i < 8; ++i) d[i] = __dadd_rn(d[i], ds); }
    if (MODE == 1) { _Pragma("unroll") for (int i = 0; i < 8; ++i) d[i] = __fma_rn(d[i], ds, d[(i + 1) & 7]); }
    if (MODE == 2) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[i])); d[i] = t; f[i] = __int_as_float(__float_as_int(f[i]) + 1); } }
    if (MODE == 3) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { float t; asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t) : "d"(d[i])); f[i] = t; d[i] = __longlong_as_double(__double_as_longlong(d[i]) + 1); } }
    if (MODE == 4) { _Pragma("unroll") for (int i = 0; i < 8; ++i) {   // integer widen f32 -> f64 (normal numbers)
        unsigned b = __float_as_uint(f[i]);
        unsigned hi = (b & 0x80000000u) | ((((b >> 23) & 0xffu) + 896u) << 20) | ((b & 0x7fffffu) >> 3);
        unsigned lo = b << 29;
        d[i] = __hiloint2double(hi, lo); f[i] = __int_as_float(__float_as_int(f[i]) + 1); } }
    if (MODE == 5) { _Pragma("unroll") for (int i = 0; i < 8; ++i) d[i] = __dmul_rn(d[i], ds); }
  }
  float r = 0; for (int i = 0; i < 8; ++i) r += (float)d[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, int instr_per_iter) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
  int iters = 4000; dim3 g(148 * 4), b(256);
  k<MODE><<<g, b>>>(out, 100, 1.0001f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<g, b>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double winstr = (double)g.x * (b.x / 32) * iters * instr_per_iter;
  printf("%-26s %8.3f ms  %.3f warp-ops/clk/SM (at 1965 MHz) = %.1f lanes/clk/SM\n", name, ms, winstr / 148.0 / (ms * 1e-3 * 1.965e9), 32 * winstr / 148.0 / (ms * 1e-3 * 1.965e9));
  cudaFree(out);
}
int main() {
  run<0>("DADD", 8); run<1>("DFMA", 8); run<5>("DMUL", 8); run<2>("cvt.f64.f32 (+IADD)", 8); run<3>("cvt.rn.f32.f64 (+IADD64)", 8); run<4>("int widen f32->f64", 8);
  return 0;
}
