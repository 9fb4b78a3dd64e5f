// occ_test.cu -- which (registers, dynamic smem) budgets a 2-CTA-cluster kernel of 320 threads can launch with.
#include <cstdio>
#include <cuda_runtime.h>
template <int R>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(R) k(int* out) {
  extern __shared__ unsigned char s[];
  if (threadIdx.x == 0 && out) out[blockIdx.x] = s[0];
}
template <int R>
void probe() {
  for (int smem = 232448; smem >= 196608; smem -= 1024) {
    if (cudaFuncSetAttribute(k<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) { cudaGetLastError(); continue; }
    k<R><<<148, 320, smem>>>(nullptr);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) { cudaDeviceSynchronize(); printf("maxnreg %d: launches with dynamic smem <= %d\n", R, smem); return; }
  }
  printf("maxnreg %d: no smem size in range launches\n", R);
}
int main() {
  probe<128>(); probe<168>(); probe<192>(); probe<200>();
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<192>);
  printf("k<192>: regs %d static smem %zu max threads %d\n", fa.numRegs, fa.sharedSizeBytes, fa.maxThreadsPerBlock);
  return 0;
}
