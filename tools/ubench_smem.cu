// ubench_smem.cu -- is the pair main loop bound by shared-memory bandwidth?  The loop of tools/ubench_pipe.cu (TMA pair
// loads through a 5-stage ring feeding tcgen05.mma.cta_group::2, M 256 x N 256, no epilogue: 546 clk per K-block) with
// the 8 otherwise idle warps adding a controlled load of one kind every `period` clocks:
//   mode 1  STS.128 (conflict-free, 4 wavefronts each)      mode 2  LDS.128 broadcast (1 wavefront each)
//   mode 3  st.global 16 B per lane, lanes 2 KB apart        mode 4  tcgen05.ld 32x32b.x32 of the idle accumulator half
//   mode 5  LDS.128 conflict-free (4 wavefronts each)        mode 6  issue-slot load only (FFMA chain)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/ubench_smem tools/ubench_smem.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define PBG_HANG_GUARD 1
#include "../pro-b-gan_b200/csrc/ptx.cuh"
#include "../pro-b-gan_b200/csrc/pass_common.cuh"
using namespace pbg;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kBox = 128 * 64 * 2;
constexpr int STAGES = 5;

__device__ __forceinline__ uint4 lds128(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
  return v;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
smem_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, int n_kb, int a_blocks,
            int mode, int n_ops, int period, char* gscratch, long long* cyc_out, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* scratch = smem + STAGES * 2 * kBox;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(scratch + 32768);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* done_bar = empty_bar + STAGES;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  volatile int* stop = reinterpret_cast<volatile int*>(slot + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(done_bar, 1);
    *stop = 0;
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(slot);
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0) {
    if (lane == 0) {
      const uint32_t lead_full = mapa_u32(smem_u32(full_bar), 0);
      uint32_t stage = 0, phase = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 4u * kBox);
        uint8_t* st = smem + stage * 2 * kBox;
        const int ablk = (pair * 37 + kb / 16) % a_blocks;
        tma_load_2d_pair(st, &tm_a, lead_full + stage * 8, (kb % 16) * 64, ablk * 256 + static_cast<int>(rank) * 128);
        tma_load_2d_pair(st + kBox, &tm_w, lead_full + stage * 8, (kb % 16) * 64, ((kb / 16) % 4) * 256 + static_cast<int>(rank) * 128);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, 256);
      uint32_t stage = 0, phase = 0;
      const long long t0 = clock64();
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * 2 * kBox);
        const uint64_t da = make_kmajor_sw128_desc(sa), db = make_kmajor_sw128_desc(sa + kBox);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem, da + 2 * k, db + 2 * k, idesc, (kb % 16 | k) != 0);   // accumulator: columns 0..255
        umma_commit_pair(&empty_bar[stage], 3);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_pair(done_bar, 3);
      mbar_wait(done_bar, 0);
      cyc_out[pair] = clock64() - t0;
    }
    if (rank != 0 && lane == 0) mbar_wait(done_bar, 0);
    __syncwarp();
    if (lane == 0) *stop = 1;
  } else if (mode > 0) {
    uint8_t* mine = scratch + (warp - 2) * 4096;
    char* gmine = gscratch + (static_cast<size_t>(blockIdx.x) * 8 + (warp - 2)) * 65536;
    uint4 v = make_uint4(lane, warp, 3, 4);
    float acc = lane;
    long long next = clock64(), done_ops = 0;
    const uint32_t tq = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    while (!*stop) {
      for (int e = 0; e < n_ops; ++e) {
        const int t = e & 7;
        if (mode == 1) *reinterpret_cast<uint4*>(mine + lane * 128 + ((t ^ (lane & 7)) << 4)) = v;
        else if (mode == 2) { const uint4 x = lds128(mine + t * 16); v.x ^= x.x; }
        else if (mode == 3) *reinterpret_cast<uint4*>(gmine + lane * 2048 + t * 16) = v;
        else if (mode == 4 || mode == 7) {   // 4: the idle accumulator half (columns 256..511); 7: the half the MMA accumulates into
          uint32_t r[32];
          tmem_ld_32x32_ptr(tq + (mode == 4 ? 256 : 0) + (e & 7) * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v.y ^= r[i];
        }
        else if (mode == 5) { const uint4 x = lds128(mine + lane * 128 + ((t ^ (lane & 7)) << 4)); v.x ^= x.x; }
        else if (mode == 8) { v.z ^= static_cast<uint32_t>(ld_relaxed_gpu(reinterpret_cast<const int*>(gmine) + t)); }
        else if (mode == 9) { if (lane == 0) red_relaxed_gpu_add(reinterpret_cast<int*>(gmine) + 64 + t, 1); }
        else if (mode == 10) { v.w ^= mbar_test_wait(done_bar, 0) ? 1u : 0u; }
        else { for (int i = 0; i < 16; ++i) acc = acc * 1.0001f + 0.5f; }
      }
      done_ops += n_ops;
      next += period;
      while (clock64() < next && !*stop) __nanosleep(20);
    }
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(cyc_out + 512), static_cast<unsigned long long>(done_ops));
    if (v.x == 0x12345678u || acc == 1.2345f) sink[threadIdx.x] = acc + v.y + v.z + v.w;
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc_pair<512>(tmem); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn enc, void* base, uint64_t rows, uint64_t cols) {
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
  cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
  return m;
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int a_blocks = 128;
  void *abuf, *wbuf; char* gs; float* sink;
  CK(cudaMalloc(&abuf, (size_t)a_blocks * 256 * 2048)); CK(cudaMemset(abuf, 0x11, (size_t)a_blocks * 256 * 2048));
  CK(cudaMalloc(&wbuf, (size_t)1024 * 2048)); CK(cudaMemset(wbuf, 0x11, (size_t)1024 * 2048));
  CK(cudaMalloc(&gs, (size_t)148 * 8 * 65536)); CK(cudaMalloc(&sink, 4096));
  long long* d_cyc; CK(cudaMalloc(&d_cyc, sizeof(long long) * 1024));
  const CUtensorMap ta = make_map(enc, abuf, (uint64_t)a_blocks * 256, 1024), tw = make_map(enc, wbuf, 1024, 1024);
  const int n_kb = 2048, grid = 48;
  const int smem = STAGES * 2 * kBox + 32768 + 256 + 1024;
  CK(cudaFuncSetAttribute(smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct Cfg { int mode, n_ops, period; const char* what; };
  const Cfg cfgs[] = {
    {0, 0, 512, "nothing"},
    {1, 16, 512, "STS.128 x16 per warp per 512 clk (512 wavefronts / 512 clk / SM)"},
    {5, 16, 512, "LDS.128 conflict-free x16 (512 wf)"},
    {3, 2, 512, "st.global 16 B scattered x2 per warp per 512 clk"},
    {3, 8, 512, "st.global 16 B scattered x8"},
    {3, 8, 2048, "st.global 16 B scattered x8 per 2048 clk"},
    {4, 1, 4096, "tcgen05.ld x32 (idle half) x1 per warp per 4096 clk"},
    {4, 1, 1024, "tcgen05.ld x32 (idle half) x1 per warp per 1024 clk"},
    {4, 1, 256, "tcgen05.ld x32 (idle half) x1 per warp per 256 clk"},
    {4, 4, 256, "tcgen05.ld x32 (idle half) x4 per warp per 256 clk"},
    {7, 1, 1024, "tcgen05.ld x32 (live half) x1 per warp per 1024 clk"},
    {8, 1, 256, "ld.relaxed.gpu poll x1 per warp per 256 clk"},
    {9, 1, 256, "red.relaxed.gpu x1 per warp per 256 clk"},
    {10, 8, 128, "mbarrier.test_wait x8 per warp per 128 clk"},
    {6, 32, 128, "FFMA chains (issue slots only)"},
  };

  for (const Cfg& c : cfgs) {
    float best = 1e30f;
    CK(cudaMemset(d_cyc + 512, 0, sizeof(long long)));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(a));
      smem_kernel<<<grid, 320, smem>>>(ta, tw, n_kb, a_blocks, c.mode, c.n_ops, c.period, gs, d_cyc, sink);
      CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (rep > 0 && ms < best) best = ms;
    }
    std::vector<long long> cyc(64);
    CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * 63, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&cyc[63], d_cyc + 512, sizeof(long long), cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < grid / 2; ++i) mean += cyc[i]; mean /= grid / 2;
    printf("%-70s : %6.0f clk per K-block (floor 512), %.2f ops per warp per K-block\n", c.what, mean / n_kb,
           (double)cyc[63] / (grid * 8.0) / n_kb / 3.0);
  }
  printf("done\n");
  return 0;
}
