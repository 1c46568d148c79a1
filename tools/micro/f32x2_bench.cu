// Microbenchmark: issue throughput of scalar FADD/FFMA vs packed FADD2/FFMA2 (sm_100a), 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fadd_(float a, float b){ float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float ffma_(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
template <int MODE> __global__ void k(float* out, int iters, float s) {
  float a[16]; u64 p[8];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) p[i] = pk(a[2*i], a[2*i+1]);
  u64 ps = pk(s, s);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) { _Pragma("unroll") for (int i = 0; i < 16; ++i) a[i] = fadd_(a[i], s); }
    if (MODE == 1) { _Pragma("unroll") for (int i = 0; i < 16; ++i) a[i] = ffma_(a[i], s, a[(i+1)&15]); }
    if (MODE == 2) { _Pragma("unroll") for (int i = 0; i < 8; ++i) p[i] = add2(p[i], ps); }
    if (MODE == 3) { _Pragma("unroll") for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ps, p[(i+1)&7]); }
    if (MODE == 4) { _Pragma("unroll") for (int i = 0; i < 8; ++i) { a[2*i] = fadd_(a[2*i], s); a[2*i+1] = ffma_(a[2*i+1], s, a[2*i]); } }
  }
  float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
  for (int i = 0; i < 8; ++i) { float x, y; upk(p[i], x, y); r += x + y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char* name, int elems_per_instr, int instr_per_iter) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
  int iters = 20000; dim3 g(148 * 4), b(256);   // 32 warps / SM
  k<MODE><<<g, b>>>(out, 100, 1.0001f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<g, b>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double winstr = (double)g.x * (b.x / 32) * iters * instr_per_iter;
  double per_sm_clk = winstr / 148.0 / (ms * 1e-3 * 1.965e9);
  printf("%-28s %8.3f ms  %.2f warp-instr/clk/SM (at 1965 MHz)  %.1f Gelem-ops/s\n", name, ms, per_sm_clk, winstr * 32 * elems_per_instr / ms / 1e6);
  cudaFree(out);
}
int main() {
  run<0>("FADD  (16 chains)", 1, 16); run<1>("FFMA  (16 chains)", 1, 16);
  run<2>("FADD2 (8 chains)", 2, 8);   run<3>("FFMA2 (8 chains)", 2, 8);
  run<4>("FADD+FFMA mix", 1, 16);
  return 0;
}
