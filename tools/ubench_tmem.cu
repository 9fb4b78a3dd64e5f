// ubench_tmem.cu -- tcgen05.ld (TMEM -> registers) throughput per SM, alone and while the tensor core is
// accumulating into the other half of TMEM (the situation of a double-buffered epilogue).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench_tmem tools/ubench_tmem.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../pro-b-gan_b200/csrc/ptx.cuh"
using namespace pbg;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <int X>
__device__ __forceinline__ void ld_shape(uint32_t taddr, uint32_t* v);
template <>
__device__ __forceinline__ void ld_shape<32>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld_shape<16>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}

// warps 2 .. 2+EPI-1 read accumulator stage 0 (columns 0..255 of their lane quarter) over and over; with mma_iters > 0
// thread 0 issues back-to-back 128 x 256 x 16 MMAs into stage 1 (columns 256..511) at the same time.
template <int X, int EPI>
__global__ void __launch_bounds__(64 + 32 * EPI, 1) tmem_ld_kernel(int ld_iters, int mma_iters, long long* cyc_ld, long long* cyc_mma) {
  constexpr int kA = 128 * 64 * 2, kB = 256 * 64 * 2, kSt = 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kSt * (kA + kB));
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  uint32_t x = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
  for (int i = threadIdx.x; i < kSt * (kA + kB) / 4; i += blockDim.x) {
    x = x * 1664525u + 1013904223u;
    const uint32_t lo = 0x3C00u | ((x >> 8) & 0x83FFu), hi = 0x3C00u | ((x >> 20) & 0x83FFu);
    w[i] = (lo & 0xBFFFu) | ((hi & 0xBFFFu) << 16);
  }
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<512>(slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0 && mma_iters > 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 256);
    const long long t0 = clock64();
    for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
      for (int s = 0; s < kSt; ++s) {
        const uint32_t sa = smem_u32(smem + s * (kA + kB));
        const uint64_t da = make_kmajor_sw128_desc(sa), db = make_kmajor_sw128_desc(sa + kA);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + 256, da + 2 * k, db + 2 * k, idesc, (it | s | k) != 0);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    cyc_mma[blockIdx.x] = clock64() - t0;
  }
  if (warp >= 2) {
    const int wep = warp - 2, q = warp & 3, half = wep >> 2;
    const uint32_t taddr = tmem + (static_cast<uint32_t>(q * 32) << 16) + (EPI > 4 ? half * 128 : 0);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < ld_iters; ++it) {
      uint32_t v[64];
#pragma unroll
      for (int j = 0; j < 64 / X; ++j) ld_shape<X>(taddr + ((it * 64) & 64) + j * X, v + j * X);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 64; ++j) acc ^= v[j];
    }
    const long long t1 = clock64();
    if (lane == 0) cyc_ld[blockIdx.x * EPI + wep] = t1 - t0;
    if (acc == 0x12345678u) cyc_ld[0] = 0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

template <int X, int EPI>
int run(int sms, int ld_iters, int mma_iters, long long* d_ld, long long* d_mma) {
  const int smem = 2 * (128 * 64 * 2 + 256 * 64 * 2) + 64 + 1024;
  auto kern = tmem_ld_kernel<X, EPI>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaMemset(d_mma, 0, sizeof(long long) * sms));
  for (int rep = 0; rep < 2; ++rep) { kern<<<sms, 64 + 32 * EPI, smem>>>(ld_iters, mma_iters, d_ld, d_mma); CK(cudaDeviceSynchronize()); }
  std::vector<long long> ld(sms * EPI), mm(sms);
  CK(cudaMemcpy(ld.data(), d_ld, sizeof(long long) * sms * EPI, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(mm.data(), d_mma, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  double ld_mean = 0; for (auto c : ld) ld_mean += c; ld_mean /= ld.size();
  double mm_mean = 0; for (auto c : mm) mm_mean += c; mm_mean /= mm.size();
  const double bytes_per_sm = (double)EPI * ld_iters * 64 * 32 * 4;
  printf("tcgen05.ld 32x32b.x%-2d  %d warps  mma %s : %.0f clk per 64-column read per warp, %.1f B/clk/SM", X, EPI, mma_iters ? "on " : "off",
         ld_mean / ld_iters, bytes_per_sm / ld_mean);
  if (mma_iters) printf("   | mma %.1f clk per 128x256x16 (floor 128)", mm_mean / (mma_iters * 8.0));
  printf("\n");
  return 0;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  long long *d_ld, *d_mma; CK(cudaMalloc(&d_ld, sizeof(long long) * sms * 8)); CK(cudaMalloc(&d_mma, sizeof(long long) * sms));
  const int L = 2000;
  run<32, 8>(sms, L, 0, d_ld, d_mma);
  run<32, 4>(sms, L, 0, d_ld, d_mma);
  run<16, 8>(sms, L, 0, d_ld, d_mma);
  run<32, 8>(sms, L, 4000, d_ld, d_mma);   // long MMA stream: covers the whole read loop
  run<32, 4>(sms, L, 4000, d_ld, d_mma);
  run<16, 8>(sms, L, 4000, d_ld, d_mma);
  run<32, 8>(sms, 0, 4000, d_ld, d_mma);   // MMA alone
  printf("done\n");
  return 0;
}
