// ofri_tma.cuh -- TMA / mbarrier plumbing shared by the persistent kernels (ofri_hs_tma.cu, ofri_ls_tma.cu): host-side
// tensor-map encoding through the driver entry point (no link-time libcuda dependency) and the PTX wrappers.
#pragma once
#include <cuda.h>

#include "ofri_internal.h"

namespace ofri {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {      // thread-safe one-time lookup (row bands may be driven by several host threads)
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    EncodeTiledFn f = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      f = reinterpret_cast<EncodeTiledFn>(p);
    cudaGetLastError();
    return f;
  }();
  return fn;
}

// 3-D map over a plane stack: dims (W, H, batch), strides (pitch, stride) floats, box (128, box_rows, 1), zero OOB fill
inline bool make_map(CUtensorMap* m, const Img& img, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)img.W, (cuuint64_t)img.H, (cuuint64_t)img.batch};
  cuuint64_t strides[2] = {(cuuint64_t)img.pitch * 4, (cuuint64_t)img.stride * 4};
  cuuint32_t box[3] = {128, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, img.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int x, int y, int z, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(z), "r"(bar)
      : "memory");
}


// ---- optional phase timing of the persistent kernels (tools/phase_timing.py; built with -DOFRI_PHASE_TIMING into a
// separate libofri_phase.so, never into the product library): thread 0 of every CTA accumulates clock64() deltas per
// phase in shared memory and adds them to the translation unit's table (g_phase_acc) at exit.
#define OFRI_NPHASE 8
#if defined(OFRI_PHASE_TIMING)
__shared__ long long ph_acc_[OFRI_NPHASE];
__shared__ long long ph_last_;
static __device__ unsigned long long g_phase_acc[OFRI_NPHASE];
#define OFRI_PH_INIT do { if (threadIdx.x == 0) { for (int i_ = 0; i_ < OFRI_NPHASE; ++i_) ph_acc_[i_] = 0; ph_last_ = clock64(); } } while (0)
#define OFRI_PH(i) do { if (threadIdx.x == 0) { long long t_ = clock64(); ph_acc_[i] += t_ - ph_last_; ph_last_ = t_; } } while (0)
#define OFRI_PH_FLUSH do { if (threadIdx.x == 0) for (int i_ = 0; i_ < OFRI_NPHASE; ++i_) \
  atomicAdd(&g_phase_acc[i_], (unsigned long long)ph_acc_[i_]); } while (0)
#define OFRI_PH_READ(out) do { cudaMemcpyFromSymbol(out, g_phase_acc, sizeof(unsigned long long) * OFRI_NPHASE); \
  unsigned long long z_[OFRI_NPHASE] = {0}; cudaMemcpyToSymbol(g_phase_acc, z_, sizeof(z_)); } while (0)
#else
#define OFRI_PH_INIT do { } while (0)
#define OFRI_PH(i) do { } while (0)
#define OFRI_PH_FLUSH do { } while (0)
#define OFRI_PH_READ(out) do { for (int i_ = 0; i_ < OFRI_NPHASE; ++i_) (out)[i_] = 0; } while (0)
#endif

}  // namespace ofri
