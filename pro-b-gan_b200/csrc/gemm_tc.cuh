// gemm_tc.cuh -- the Linear layers of the generator / discriminator as one persistent,
// warp-specialised tcgen05 GEMM with the layer's bias + activation fused in the epilogue.
//
//   out[M, N] = act( A[M, K] (bf16, K-major)  x  W[N, K]^T (bf16, K-major = nn.Linear layout) + bias[N] )
//
//   warp 0      TMA producer   : cp.async.bulk.tensor 128x64 A box + BLOCK_Nx64 W box per stage (SWIZZLE_128B)
//   warp 1      MMA issuer     : one thread, tcgen05.mma.cta_group::1.kind::f16, 128 x BLOCK_N x 16 per instruction,
//                                fp32 accumulators in TMEM, two accumulator stages (2 x BLOCK_N columns)
//   warps 2..5  epilogue       : tcgen05.ld 32x32b (thread <-> accumulator row), bias + activation, then
//                                  EPI_LEAKY  -> bf16 activations for the next layer
//                                  EPI_TANH   -> tanh, bf16 or fp32 generator output (optionally + cosine vs tail)
//                                  EPI_ROWDOT -> LeakyReLU then dot with the final [H/2 -> 1] weight: the
//                                                discriminator's last Linear is folded in as a per-row reduction,
//                                                the logit and sigmoid(logit) are the only things written.
//   Pipelines: smem full/empty ring (TMA <-> MMA), TMEM full/empty pair (MMA <-> epilogue), static persistent
//   tile schedule (work item = blockIdx.x + i * gridDim.x).
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace pbg {

enum : int { EPI_LEAKY = 0, EPI_TANH = 1, EPI_ROWDOT = 2 };

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;
constexpr int kNumEpiThreads = 128;

struct GemmParams {
  int M;              // valid rows of A / out
  int N;              // padded output width (multiple of BLOCK_N)
  int K;              // padded reduction width (multiple of 64)
  int n_valid;        // columns of out that exist (EPI_TANH store guard), <= N
  const float* bias;  // [N] fp32 (zero in the padding)
  void* out;          // EPI_LEAKY: bf16 [M, ldo]; EPI_TANH: bf16 / fp32 [M, ldo]
  long long ldo;      // leading dimension of out, elements
  float slope;        // LeakyReLU negative slope
  int out_f32;        // EPI_TANH: 1 = fp32 output, 0 = bf16
  // EPI_ROWDOT
  const float* w3;    // [N] fp32 final-layer weight (zero in the padding)
  float b3;
  float* logits;      // [M]
  float* probs;       // [M] or nullptr
  // EPI_TANH optional fused generator score: cosine(out_row, tail_row), pro_b_gan_infer.py:202
  const float* tail_tab;     // node_emb [Nent, n_valid] fp32 or nullptr
  const long long* tail_idx; // tail ids, element stride tail_stride
  long long tail_stride;
  long long n_ent;
  float* cosine;             // [M]
  long long* trace;          // diagnostics: 16 clock64 slots per CTA (pbg_debug_trace), or nullptr
};

template <int BLOCK_N, int STAGES>
struct GemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 4) * 8 + 16 + 1024 /*alignment slack*/;
};

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int BLOCK_N, int STAGES, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmParams p) {
  using L = GemmSmem<BLOCK_N, STAGES>;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N >= 32 && BLOCK_N <= 256, "UMMA N");
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages
  static_assert(kTmemCols == 64 || kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM cols");
  constexpr uint32_t kIdesc = make_idesc_bf16(kBlockM, BLOCK_N);

  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte aligned bases
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_tiles = p.N / BLOCK_N;
  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int num_kb = p.K / kBlockK;
  // EPI_LEAKY: a work item is one (m, n) tile.  Row-wise epilogues (ROWDOT, TANH + cosine) need the whole
  // output row in one thread, so their work item is one m block and the n tiles run back to back in it.
  constexpr bool kRowItems = (EPI != EPI_LEAKY);
  const int units_per_item = kRowItems ? n_tiles : 1;
  const int num_items = kRowItems ? m_tiles : m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kNumEpiThreads);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long* tr = p.trace ? p.trace + 256 * blockIdx.x : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      long long w_empty = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        for (int u = 0; u < units_per_item; ++u) {
          const int m_blk = kRowItems ? item : item / n_tiles;
          const int n_blk = kRowItems ? u : item % n_tiles;
          for (int kb = 0; kb < num_kb; ++kb) {
            if (tr) { const long long t = clock64(); mbar_wait(&empty_bar[stage], phase ^ 1); w_empty += clock64() - t; }
            else mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
            tma_load_2d(sb, &tmap_b, &full_bar[stage], kb * kBlockK, n_blk * BLOCK_N);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (tr) { tr[1] = w_empty; tr[2] = clock64(); }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      long long w_full = 0, w_tmem = 0, t_first = 0, n_kb = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        for (int u = 0; u < units_per_item; ++u) {
          if (tr) { const long long t = clock64(); mbar_wait(&tmem_empty[acc], acc_phase ^ 1); w_tmem += clock64() - t; }
          else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          for (int kb = 0; kb < num_kb; ++kb) {
            if (tr) {
              const long long t = clock64(); mbar_wait(&full_bar[stage], phase); const long long t2 = clock64();
              w_full += t2 - t; if (n_kb++ == 0) t_first = t2;
            } else mbar_wait(&full_bar[stage], phase);  // TMA bytes have landed
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
            const uint64_t da = make_kmajor_sw128_desc(sa);
            const uint64_t db = make_kmajor_sw128_desc(sa + L::kABytes);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
              umma_bf16(d_tmem, da + 2 * k, db + 2 * k, kIdesc, (kb | k) != 0);
            }
            umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
      if (tr) { tr[3] = w_full; tr[4] = w_tmem; tr[5] = t_first; tr[6] = clock64(); tr[9] = n_kb; }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t acc = 0, acc_phase = 0;
    long long w_acc = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int m_blk = kRowItems ? item : item / n_tiles;
      const long long grow = static_cast<long long>(m_blk) * kBlockM + q * 32 + lane;
      const bool row_ok = grow < p.M;
      float rowdot = 0.f;
      float cs_dot = 0.f, cs_pp = 0.f, cs_tt = 0.f;
      const float* trow = nullptr;
      if (EPI == EPI_TANH && p.cosine != nullptr && row_ok) {
        long long tid = p.tail_idx[grow * p.tail_stride];
        tid = (tid < 0 || tid >= p.n_ent) ? 0 : tid;  // gather kernel has already flagged it
        trow = p.tail_tab + tid * p.n_valid;
      }
      for (int u = 0; u < units_per_item; ++u) {
        const int n_blk = kRowItems ? u : item % n_tiles;
        if (tr && threadIdx.x == 64) { const long long t = clock64(); mbar_wait(&tmem_full[acc], acc_phase); w_acc += clock64() - t; tr[10] = clock64(); }
        else mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait();
          if (c == BLOCK_N / 32 - 1) {
            // every column of this accumulator stage is now in registers: hand it back to the MMA warp
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
          }
          const int col0 = n_blk * BLOCK_N + c * 32;
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
          float f[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b.x;
            f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
            f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
            f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
          }
          if (EPI == EPI_LEAKY) {
            if (row_ok) {
              uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + grow * p.ldo + col0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 w;
                w.x = pack_bf16x2(leaky(f[8 * j + 0], p.slope), leaky(f[8 * j + 1], p.slope));
                w.y = pack_bf16x2(leaky(f[8 * j + 2], p.slope), leaky(f[8 * j + 3], p.slope));
                w.z = pack_bf16x2(leaky(f[8 * j + 4], p.slope), leaky(f[8 * j + 5], p.slope));
                w.w = pack_bf16x2(leaky(f[8 * j + 6], p.slope), leaky(f[8 * j + 7], p.slope));
                dst[j] = w;
              }
            }
          } else if (EPI == EPI_ROWDOT) {
            const float4* w4 = reinterpret_cast<const float4*>(p.w3 + col0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 w = __ldg(w4 + j);
              rowdot = fmaf(leaky(f[4 * j + 0], p.slope), w.x, rowdot);
              rowdot = fmaf(leaky(f[4 * j + 1], p.slope), w.y, rowdot);
              rowdot = fmaf(leaky(f[4 * j + 2], p.slope), w.z, rowdot);
              rowdot = fmaf(leaky(f[4 * j + 3], p.slope), w.w, rowdot);
            }
          } else {  // EPI_TANH
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = tanh_fast(f[j]);
            if (row_ok && col0 < p.n_valid) {  // n_valid is a multiple of 8; chunks are 32 wide
              const int nv = min(32, p.n_valid - col0);
              if (trow != nullptr) {
                for (int j = 0; j < nv; j += 4) {
                  const float4 t = __ldg(reinterpret_cast<const float4*>(trow + col0 + j));
                  cs_dot += f[j] * t.x + f[j + 1] * t.y + f[j + 2] * t.z + f[j + 3] * t.w;
                  cs_pp += f[j] * f[j] + f[j + 1] * f[j + 1] + f[j + 2] * f[j + 2] + f[j + 3] * f[j + 3];
                  cs_tt += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
                }
              }
              if (p.out != nullptr) {
                if (p.out_f32) {
                  float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + grow * p.ldo + col0);
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    if (4 * j < nv) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                } else {
                  uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + grow * p.ldo + col0);
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    if (8 * j < nv) {
                      uint4 w;
                      w.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
                      w.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
                      w.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
                      w.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
                      dst[j] = w;
                    }
                }
              }
            }
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (EPI == EPI_ROWDOT && row_ok) {
        const float logit = rowdot + p.b3;
        p.logits[grow] = logit;
        if (p.probs != nullptr) p.probs[grow] = 1.f / (1.f + __expf(-logit));
      }
      if (EPI == EPI_TANH && trow != nullptr) {
        // F.cosine_similarity(pred, t, dim=1), eps = 1e-8 on each norm (pro_b_gan_infer.py:202)
        const float np = fmaxf(sqrtf(cs_pp), 1e-8f), nt = fmaxf(sqrtf(cs_tt), 1e-8f);
        p.cosine[grow] = cs_dot / (np * nt);
      }
    }
    if (tr && threadIdx.x == 64) { tr[7] = w_acc; tr[8] = clock64(); }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace pbg
