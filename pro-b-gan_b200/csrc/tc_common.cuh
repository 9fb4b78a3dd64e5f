// tc_common.cuh -- constants and small device helpers shared by the tensor-core kernels (pass2_kernel.cuh, topk_kernel.cuh).
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace pbg {

constexpr int kBlockM = 128;  // rows of one CTA's half of a pair tile; also the granularity of the row-block counters
constexpr int kBlockK = 64;   // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kUmmaK = 16;    // K of one tcgen05.mma kind::f16

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace pbg
