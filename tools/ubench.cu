// ubench.cu -- microbenchmarks that size the fused kernel's tiling on a B200 (results: profiles/ubench_r1.txt).
//   tma   : L2 -> shared-memory throughput of 128x64 bf16 SWIZZLE_128B boxes, persistent CTAs, 8-stage ring;
//           unicast (distinct tiles / tiles shared by all CTAs) and cluster multicast (each CTA loads 1/CSZ of the
//           box and multicasts it to the whole cluster).
//   mma   : tcgen05.mma issue throughput with operands already in shared memory, cta_group::1 (N = 64/128/256)
//           and cta_group::2 (M = 256, N = 256).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/ubench tools/ubench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define PBG_HANG_GUARD 1
#include "../pro-b-gan_b200/csrc/ptx.cuh"

using namespace pbg;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)),
      "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask) : "memory");
}

constexpr int kStages = 8;
constexpr int kBoxBytes = 128 * 64 * 2;

// mode 0: distinct tiles per cluster, 1: all clusters read the same 128 tiles (2 MiB)
template <int CSZ>
__global__ void __launch_bounds__(64, 1) tma_bw_kernel(const __grid_constant__ CUtensorMap tmap, int iters, int mode,
                                                       int row_blocks, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kBoxBytes);
  uint64_t* empty = full + kStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CSZ > 1 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CSZ;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CSZ); }
    fence_mbar_init();
  }
  __syncthreads();
  if (CSZ > 1) cluster_sync_all();
  const long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < iters; ++it) {
      const int t = mode == 0 ? (cluster_id * 131 + it) : it;
      const int rb = (t / 16) % row_blocks, cb = t % 16;
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full[stage], kBoxBytes);
      uint8_t* dst = smem + stage * kBoxBytes;
      if (CSZ == 1) {
        tma_load_2d(dst, &tmap, &full[stage], cb * 64, rb * 128);
      } else {
        constexpr int rows = 128 / CSZ;
        tma_load_2d_mc(dst + rank * rows * 128, &tmap, &full[stage], cb * 64, rb * 128 + rank * rows,
                       static_cast<uint16_t>((1u << CSZ) - 1));
      }
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[stage], phase);
      if (CSZ == 1) { if (lane == 0) mbar_arrive(&empty[stage]); }
      else if (lane < CSZ) mbar_arrive_remote(&empty[stage], lane);   // one lane per destination CTA
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (CSZ > 1) cluster_sync_all();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
}

// Unicast variants: BOX_ROWS x 64 boxes, STAGES-deep ring, optional 1-D bulk copy of the same byte count.
template <int BOX_ROWS, int STAGES, bool BULK1D>
__global__ void __launch_bounds__(64, 2) tma_var_kernel(const __grid_constant__ CUtensorMap tmap, const uint8_t* base,
                                                        int iters, int row_blocks, long long* cycles_out) {
  constexpr int kBytes = BOX_ROWS * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * kBytes);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < iters; ++it) {
      const int t = blockIdx.x * 131 + it;
      const int rb = (t / 16) % row_blocks, cb = t % 16;
      mbar_wait(&empty[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full[stage], kBytes);
      uint8_t* dst = smem + stage * kBytes;
      if (BULK1D) {
        const uint8_t* src = base + (static_cast<size_t>(t) % (static_cast<size_t>(row_blocks) * 16 * 16384 / kBytes)) * kBytes;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(kBytes), "r"(smem_u32(&full[stage])) : "memory");
      } else {
        tma_load_2d(dst, &tmap, &full[stage], cb * 64, rb * BOX_ROWS);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[stage], phase);
      mbar_arrive(&empty[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
}

// Self-consuming producers: NPROD threads (one per warp), each owns a ring of STAGES slots and waits for its own
// slot's previous load right before re-issuing into it -> no second thread in the loop, in-flight depth = STAGES.
template <int BOX_ROWS, int STAGES, int NPROD, bool GUARD>
__global__ void __launch_bounds__(32 * NPROD, 1) tma_self_kernel(const __grid_constant__ CUtensorMap tmap, int iters,
                                                                  int row_blocks, long long* cycles_out) {
  constexpr int kBytes = BOX_ROWS * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NPROD * STAGES * kBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NPROD * STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (lane == 0) {
    uint32_t stage = 0, phase = 0;
    uint64_t* myfull = full + warp * STAGES;
    uint8_t* mysmem = smem + warp * STAGES * kBytes;
    int rb = (blockIdx.x * 7 + warp * 3) % row_blocks, cb = 0;
    for (int it = 0; it < iters; ++it) {
      if (it >= STAGES) {
        if (GUARD) mbar_wait(&myfull[stage], phase ^ 1);
        else while (!mbar_try_wait(&myfull[stage], phase ^ 1)) {}
      }
      mbar_arrive_expect_tx(&myfull[stage], kBytes);
      tma_load_2d(mysmem + stage * kBytes, &tmap, &myfull[stage], cb * 64, rb * BOX_ROWS);
      if (++cb == 16) { cb = 0; if (++rb == row_blocks) rb = 0; }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    // drain
    for (int s = 0; s < STAGES && s < iters; ++s) {
      const int last_it = iters - 1 - ((iters - 1 - s) % STAGES == 0 ? 0 : 0);
      (void)last_it;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
  // let outstanding loads land before the CTA exits
  if (lane == 0) {
    uint64_t* myfull = full + warp * STAGES;
    for (int s = 0; s < STAGES; ++s) {
      const int uses = (iters - s + STAGES - 1) / STAGES;       // how many times slot s was armed
      if (uses > 0) while (!mbar_try_wait(&myfull[s], (uses - 1) & 1)) {}
    }
  }
}

// Plain LDG.128 streaming from L2 (no shared memory): the per-SM L2 read ceiling of the LSU path.
__global__ void __launch_bounds__(1024, 1) ldg_bw_kernel(const uint4* __restrict__ base, size_t n_vec, int iters,
                                                         uint32_t* sink, long long* cycles_out) {
  const long long t0 = clock64();
  uint4 acc = make_uint4(0, 0, 0, 0);
  size_t off = (static_cast<size_t>(blockIdx.x) * 131071u * 64u) % n_vec;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      size_t i = off + static_cast<size_t>(u) * blockDim.x + threadIdx.x;
      if (i >= n_vec) i -= n_vec;
      uint4 v;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(base + i));
      acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
    off += 8u * blockDim.x; if (off >= n_vec) off -= n_vec;
  }
  if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345u) sink[0] = 1;
  __syncthreads();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
}

// ---------------------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* cycles_out) {
  constexpr int kA = 128 * 64 * 2, kB = N * 64 * 2, kSt = 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kSt * (kA + kB));
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  // pseudo-random bf16 operands in (-2, 2): realistic toggling for the power / clock behaviour
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  uint32_t x = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
  for (int i = threadIdx.x; i < kSt * (kA + kB) / 4; i += blockDim.x) {
    x = x * 1664525u + 1013904223u;
    const uint32_t lo = 0x3C00u | ((x >> 8) & 0x83FFu), hi = 0x3C00u | ((x >> 20) & 0x83FFu);
    w[i] = (lo & 0xBFFFu) | ((hi & 0xBFFFu) << 16);
  }
  fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<512>(slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t idesc = make_idesc_bf16(128, N);
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < kSt; ++s) {
        const uint32_t sa = smem_u32(smem + s * (kA + kB));
        const uint64_t da = make_kmajor_sw128_desc(sa), db = make_kmajor_sw128_desc(sa + kA);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + (it & 1) * N, da + 2 * k, db + 2 * k, idesc, (it | s | k) != 0);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    cycles_out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

// cta_group::2: M = 256 (128 rows of A per CTA), N = 256 (128 rows of W per CTA); the leader CTA issues.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_rate_2cta_kernel(int iters, long long* cycles_out) {
  constexpr int kA = 128 * 64 * 2, kB = 128 * 64 * 2, kSt = 4;
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
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  const uint32_t tmem = *slot;
  constexpr uint32_t idesc = make_idesc_bf16(256, 256);
  if (rank == 0 && threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < kSt; ++s) {
        const uint32_t sa = smem_u32(smem + s * (kA + kB));
        const uint64_t da = make_kmajor_sw128_desc(sa), db = make_kmajor_sw128_desc(sa + kA);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = (it | s | k) != 0;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem + (it & 1) * 256),
              "l"(da + 2 * k), "l"(db + 2 * k), "r"(idesc), "r"(acc) : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
  }
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    mbar_wait(bar, 0);
    if (rank == 0) cycles_out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CSZ>
void run_tma(EncodeTiledFn enc, void* buf, int rows, int mode, int grid, int iters, long long* d_cyc) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {1024, (cuuint64_t)rows};
  cuuint64_t strides[1] = {2048};
  cuuint32_t box[2] = {64, (cuuint32_t)(128 / CSZ)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  const int smem = kStages * kBoxBytes + 2 * kStages * 8 + 1024;
  auto kern = tma_bw_kernel<CSZ>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (CSZ > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid / CSZ * CSZ); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CSZ; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  const int row_blocks = rows / 128;
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a));
    CK(cudaLaunchKernelEx(&cfg, kern, tm, iters, mode, row_blocks, d_cyc));
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  const int ctas = grid / CSZ * CSZ;
  const double landed = (double)ctas * iters * kBoxBytes;          // bytes written into shared memory
  const double from_l2 = landed / CSZ;                               // bytes requested from L2 (multicast: once per cluster)
  std::vector<long long> cyc(ctas);
  CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
  printf("tma  csz=%-2d mode=%s ctas=%3d iters=%d : %.3f ms  smem-landed %.2f TB/s (%.1f B/clk/SM)  L2-requested %.2f TB/s  [max %lld clk -> %.2f GHz]\n",
         CSZ, mode == 0 ? "distinct" : "shared  ", ctas, iters, best, landed / best * 1e-9, (double)iters * kBoxBytes / mx,
         from_l2 / best * 1e-9, mx, mx / (best * 1e6));
}

template <int BOX_ROWS, int STAGES, bool BULK1D>
void run_var(EncodeTiledFn enc, void* buf, int rows, int grid, int iters, long long* d_cyc, int blocks_per_sm = 1) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {1024, (cuuint64_t)rows};
  cuuint64_t strides[1] = {2048};
  cuuint32_t box[2] = {64, (cuuint32_t)BOX_ROWS};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  constexpr int kBytes = BOX_ROWS * 128;
  const int smem = STAGES * kBytes + 2 * STAGES * 8 + 1024;
  auto kern = tma_var_kernel<BOX_ROWS, STAGES, BULK1D>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  const int ctas = grid * blocks_per_sm;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a));
    kern<<<ctas, 64, smem>>>(tm, (const uint8_t*)buf, iters, rows / BOX_ROWS, d_cyc);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  std::vector<long long> cyc(ctas);
  CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
  const double landed = (double)ctas * iters * kBytes;
  printf("var  %s box=%3dx64 (%2d KB) stages=%2d ctas=%3d (x%d/SM) : %.3f ms  %.2f TB/s  %.1f B/clk/SM\n", BULK1D ? "bulk1d" : "tma2d ",
         BOX_ROWS, kBytes / 1024, STAGES, ctas, blocks_per_sm, best, landed / best * 1e-9,
         (double)iters * kBytes * blocks_per_sm / mx);
}

template <int BOX_ROWS, int STAGES, int NPROD, bool GUARD>
void run_self(EncodeTiledFn enc, void* buf, int rows, int grid, int iters, long long* d_cyc) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {1024, (cuuint64_t)rows};
  cuuint64_t strides[1] = {2048};
  cuuint32_t box[2] = {64, (cuuint32_t)BOX_ROWS};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
  constexpr int kBytes = BOX_ROWS * 128;
  const int smem = NPROD * STAGES * kBytes + NPROD * STAGES * 8 + 1024;
  auto kern = tma_self_kernel<BOX_ROWS, STAGES, NPROD, GUARD>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a));
    kern<<<grid, 32 * NPROD, smem>>>(tm, iters, rows / BOX_ROWS, d_cyc);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  std::vector<long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
  const double landed = (double)grid * NPROD * iters * kBytes;
  printf("self box=%3dx64 (%2d KB) stages=%2d producers=%d guard=%d ctas=%3d : %.3f ms  %.2f TB/s  %.1f B/clk/SM  (%.0f clk per op per producer)\n",
         BOX_ROWS, kBytes / 1024, STAGES, NPROD, (int)GUARD, grid, best, landed / best * 1e-9, (double)NPROD * iters * kBytes / mx,
         (double)mx / iters);
}

void run_ldg(void* buf, size_t bytes, int grid, int iters, long long* d_cyc) {
  uint32_t* sink; CK(cudaMalloc(&sink, 4));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a));
    ldg_bw_kernel<<<grid, 1024>>>((const uint4*)buf, bytes / 16, iters, sink, d_cyc);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  std::vector<long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
  const double total = (double)grid * iters * 8 * 1024 * 16;
  printf("ldg  v4 1024 thr x 8 in flight ctas=%3d : %.3f ms  %.2f TB/s  %.1f B/clk/SM\n", grid, best, total / best * 1e-9,
         (double)iters * 8 * 1024 * 16 / mx);
}

template <int N>
void run_mma(int grid, int iters, long long* d_cyc) {
  const int smem = 4 * (128 * 64 * 2 + N * 64 * 2) + 64 + 1024;
  auto kern = mma_rate_kernel<N>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a));
    kern<<<grid, 128, smem>>>(iters, d_cyc);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  std::vector<long long> cyc(grid);
  CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
  const double flop = (double)grid * iters * 16 * 2.0 * 128 * N * 16;
  printf("mma  cta_group=1 M=128 N=%-3d ctas=%3d iters=%d : %.3f ms  %.1f TFLOP/s  %.1f clk per MMA (floor %d)  [%.2f GHz]\n", N, grid, iters,
         best, flop / best * 1e-9, (double)mx / (iters * 16.0), N / 2, mx / (best * 1e6));
}

void run_mma2(int grid, int iters, long long* d_cyc) {
  const int smem = 4 * (128 * 64 * 2 * 2) + 64 + 1024;
  CK(cudaFuncSetAttribute(mma_rate_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  grid = grid / 2 * 2;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(a));
    mma_rate_2cta_kernel<<<grid, 128, smem>>>(iters, d_cyc);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  const double flop = (double)(grid / 2) * iters * 16 * 2.0 * 256 * 256 * 16;
  printf("mma  cta_group=2 M=256 N=256 ctas=%3d iters=%d : %.3f ms  %.1f TFLOP/s\n", grid, iters, best, flop / best * 1e-9);
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int rows = 32768;  // 32768 x 1024 bf16 = 64 MiB, L2-resident
  void* buf; CK(cudaMalloc(&buf, (size_t)rows * 2048)); CK(cudaMemset(buf, 0x11, (size_t)rows * 2048));
  long long* d_cyc; CK(cudaMalloc(&d_cyc, sizeof(long long) * 1024));
  const int iters = 4000;
  for (int mode = 0; mode < 2; ++mode) {
    run_tma<1>(enc, buf, rows, mode, sms, iters, d_cyc);
    run_tma<2>(enc, buf, rows, mode, sms, iters, d_cyc);
    run_tma<4>(enc, buf, rows, mode, sms, iters, d_cyc);
    run_tma<8>(enc, buf, rows, mode, sms, iters, d_cyc);
  }
  run_self<128, 1, 1, true>(enc, buf, rows, sms, iters, d_cyc);
  run_self<128, 2, 1, true>(enc, buf, rows, sms, iters, d_cyc);
  run_self<128, 4, 1, true>(enc, buf, rows, sms, iters, d_cyc);
  run_self<128, 8, 1, true>(enc, buf, rows, sms, iters, d_cyc);
  run_self<128, 8, 1, false>(enc, buf, rows, sms, iters, d_cyc);
  run_self<128, 4, 2, false>(enc, buf, rows, sms, iters, d_cyc);
  run_self<128, 3, 4, false>(enc, buf, rows, sms, iters, d_cyc);
  run_self<64, 4, 4, false>(enc, buf, rows, sms, iters, d_cyc);
  run_self<256, 3, 2, false>(enc, buf, rows, sms, iters / 2, d_cyc);
  run_self<128, 3, 4, false>(enc, buf, rows, 1, iters, d_cyc);
  return 0;
  run_var<128, 4, false>(enc, buf, rows, sms, iters, d_cyc);
  run_var<128, 12, false>(enc, buf, rows, sms, iters, d_cyc);
  run_var<256, 6, false>(enc, buf, rows, sms, iters / 2, d_cyc);
  run_var<64, 16, false>(enc, buf, rows, sms, iters * 2, d_cyc);
  run_var<32, 32, false>(enc, buf, rows, sms, iters * 4, d_cyc);
  run_var<128, 6, false>(enc, buf, rows, sms, iters, d_cyc, 2);
  run_var<128, 8, true>(enc, buf, rows, sms, iters, d_cyc);
  run_var<256, 6, true>(enc, buf, rows, sms, iters / 2, d_cyc);
  run_var<128, 8, true>(enc, buf, rows, 1, iters, d_cyc);
  run_ldg(buf, (size_t)rows * 2048, sms, 500, d_cyc);
  run_ldg(buf, (size_t)rows * 2048, 1, 500, d_cyc);
  if (argc > 1) {
    run_mma<256>(sms, 4000, d_cyc);
    run_mma<128>(sms, 4000, d_cyc);
    run_mma<64>(sms, 4000, d_cyc);
    run_mma2(sms, 4000, d_cyc);
  }
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
