// ubench_pipe.cu -- the pass kernel's main loop in isolation: CTA pairs, tcgen05.mma.cta_group::2 (M 256 x N 256,
// K-blocks of 64) fed by cp.async.bulk.tensor pair loads (A: 128 rows x 64, W: 128 rows x 64 per CTA and K-block)
// through an S-stage ring, no epilogue.  Answers: what does this loop sustain per K-block (MMA floor: 512 clk) as a
// function of ring depth and of how many SMs pull from L2 at once?  Optional extra shared-memory traffic per K-block
// (generic stores by 8 idle warps) imitates the epilogues' staging writes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DPBG_TRY_WAIT_HINT_NS=1000 -o tools/ubench_pipe tools/ubench_pipe.cu -lcuda
//        (the short wait hint keeps the hang guard at 4 M rounds: the peer CTA waits on one barrier across the whole kernel)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define PBG_HANG_GUARD 1
#include "../pro-b-gan_b200/csrc/ptx.cuh"
using namespace pbg;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kBox = 128 * 64 * 2;   // 16 KB: 128 rows x 64 bf16, SWIZZLE_128B

template <int STAGES, int EXTRA>     // EXTRA: 16-byte shared-memory stores per epilogue-warp lane and K-block
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
pipe_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, int n_kb, int a_blocks,
            long long* cyc_out, long long* wait_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* scratch = smem + STAGES * 2 * kBox;                     // 32 KB for the imitation epilogue traffic
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
        // A: a fresh 256-row block every 16 K-blocks (activations: distinct addresses); W: a 2 MB weight matrix re-read
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
      long long waited = 0;
      const long long t0 = clock64();
      for (int kb = 0; kb < n_kb; ++kb) {
        const long long w0 = clock64();
        mbar_wait(&full_bar[stage], phase);
        waited += clock64() - w0;
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * 2 * kBox);
        const uint64_t da = make_kmajor_sw128_desc(sa), db = make_kmajor_sw128_desc(sa + kBox);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem + ((kb / 16) & 1) * 256, da + 2 * k, db + 2 * k, idesc, (kb % 16 | k) != 0);
        umma_commit_pair(&empty_bar[stage], 3);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_pair(done_bar, 3);
      mbar_wait(done_bar, 0);
      cyc_out[pair] = clock64() - t0;
      wait_out[pair] = waited;
    }
    if (rank != 0 && lane == 0) mbar_wait(done_bar, 0);
    __syncwarp();
    if (lane == 0) *stop = 1;
  } else if (EXTRA > 0) {
    // imitation epilogue: EXTRA x 16 B generic stores per lane per ~500 clk into a private 4 KB tile
    uint4* mine = reinterpret_cast<uint4*>(scratch + (warp - 2) * 4096);
    uint4 v = make_uint4(lane, warp, 3, 4);
    while (!*stop) {
#pragma unroll
      for (int e = 0; e < EXTRA; ++e) mine[(e * 32 + lane) & 255] = v;
      v.x += 1;
      __nanosleep(200);
    }
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

template <int STAGES, int EXTRA>
void run(const CUtensorMap& ta, const CUtensorMap& tw, int grid, int n_kb, int a_blocks, long long* d_cyc) {
  const int smem = STAGES * 2 * kBox + 32768 + 256 + 1024;
  auto kern = pipe_kernel<STAGES, EXTRA>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  grid = grid / 2 * 2;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(a));
    kern<<<grid, 320, smem>>>(ta, tw, n_kb, a_blocks, d_cyc, d_cyc + 512);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  std::vector<long long> cyc(1024);
  CK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * 1024, cudaMemcpyDeviceToHost));
  double mean = 0, mx = 0, wt = 0;
  for (int i = 0; i < grid / 2; ++i) { mean += cyc[i]; mx = cyc[i] > mx ? cyc[i] : mx; wt += cyc[512 + i]; }
  mean /= grid / 2; wt /= grid / 2;
  const double flop = (double)(grid / 2) * n_kb * 2.0 * 256 * 256 * 64;
  printf("pipe stages=%d extra_sts=%2d ctas=%3d : %.3f ms  %7.1f TFLOP/s  %6.0f clk per K-block (mean; slowest pair %.0f; floor 512)  "
         "MMA thread waited for data %.0f clk per K-block  %.1f B/clk/SM landed\n",
         STAGES, EXTRA, grid, best, flop / best * 1e-9, mean / n_kb, mx / n_kb, wt / n_kb, 2.0 * kBox * n_kb / mean);
}

// random bf16 values in (-1, 1) with full-entropy mantissas (a hash of the element index), two per 32-bit word
__global__ void fill_random_bf16(uint32_t* p, size_t n_words, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = static_cast<uint32_t>(i) * 2654435761u + seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    // per half: sign from the hash, exponent 0x7c..0x7e (|v| in [0.125, 1)), 7 random mantissa bits
    const uint32_t lo = (x & 0x8000u) | ((0x7cu + ((x >> 7) % 3u)) << 7) | (x & 0x7fu);
    const uint32_t hi = ((x >> 16) & 0x8000u) | ((0x7cu + ((x >> 23) % 3u)) << 7) | ((x >> 16) & 0x7fu);
    p[i] = lo | (hi << 16);
  }
}

int main(int argc, char** argv) {
  // ./ubench_pipe [random]: constant operand values (0x1111 everywhere) or random bf16 -- the same addresses, bytes and
  // instructions either way; what changes is how many bits toggle in the tensor pipe
  const bool random_data = argc > 1 && argv[1][0] == 'r';
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int a_blocks = 128;                      // 128 x 256 rows x 1024 bf16 = 64 MiB of "activations" (L2-resident)
  void *abuf, *wbuf;
  CK(cudaMalloc(&abuf, (size_t)a_blocks * 256 * 2048)); CK(cudaMemset(abuf, 0x11, (size_t)a_blocks * 256 * 2048));
  CK(cudaMalloc(&wbuf, (size_t)1024 * 2048)); CK(cudaMemset(wbuf, 0x11, (size_t)1024 * 2048));
  if (random_data) {
    fill_random_bf16<<<1184, 256>>>(static_cast<uint32_t*>(abuf), (size_t)a_blocks * 256 * 2048 / 4, 1u);
    fill_random_bf16<<<1184, 256>>>(static_cast<uint32_t*>(wbuf), (size_t)1024 * 2048 / 4, 2u);
    CK(cudaDeviceSynchronize());
  }
  printf("operand values: %s\n", random_data ? "random bf16 in (-1, 1)" : "constant (0x1111)");
  long long* d_cyc; CK(cudaMalloc(&d_cyc, sizeof(long long) * 1024));
  const CUtensorMap ta = make_map(enc, abuf, (uint64_t)a_blocks * 256, 1024), tw = make_map(enc, wbuf, 1024, 1024);
  const int n_kb = 4096;
  for (int grid : {sms, 48, 2}) {
    run<3, 0>(ta, tw, grid, n_kb, a_blocks, d_cyc);
    run<4, 0>(ta, tw, grid, n_kb, a_blocks, d_cyc);
    run<5, 0>(ta, tw, grid, n_kb, a_blocks, d_cyc);
    run<6, 0>(ta, tw, grid, n_kb, a_blocks, d_cyc);
  }
  run<5, 8>(ta, tw, sms, n_kb, a_blocks, d_cyc);
  run<5, 32>(ta, tw, sms, n_kb, a_blocks, d_cyc);
  CK(cudaDeviceSynchronize());
  printf("done\n");
  return 0;
}
