// gemm_f32.cuh -- fp32 parity mode of the Linear layers (BASELINE config 2: max-abs 1e-4 vs the CPU oracle).
// kind::tf32 (10-bit mantissa) over K = 1024 cannot hold 1e-4, so this mode is a plain SIMT FFMA GEMM with
// fp32 operands and fp32 accumulation; it exists for numerical parity, not for throughput.
//
//   out[M, N] = act( A[M, K] x W[N, K]^T + bias[N] ),  act in {LeakyReLU, tanh}
//
// 64 x 64 output tile per 256-thread CTA, 4 x 4 micro-tile per thread, K consumed 16 at a time through
// shared memory (tiles stored k-major so the inner product reads are conflict-free broadcasts).
#pragma once
#include <stdint.h>

namespace pbg {

enum : int { ACT_LEAKY = 0, ACT_TANH = 1, ACT_NONE = 2 };   // ACT_NONE: the plain product, no bias (top-k general path)

struct F32GemmParams {
  const float* A; long long lda;
  const float* W; long long ldw;  // [N, K] row-major
  const float* bias;
  float* out; long long ldo;
  int M, N, K;
  float slope;
};

template <int ACT>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const F32GemmParams p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4 x 4 outputs
  const long long m0 = static_cast<long long>(blockIdx.y) * BM;
  const int n0 = blockIdx.x * BN;
  const int lrow = tid >> 2;          // 0..63: tile row this thread loads
  const int lk = (tid & 3) * 4;       // 0,4,8,12: k offset of its float4

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), wv = av;
    if (m0 + lrow < p.M && k0 + lk < p.K)
      av = *reinterpret_cast<const float4*>(p.A + (m0 + lrow) * p.lda + k0 + lk);
    if (n0 + lrow < p.N && k0 + lk < p.K)
      wv = *reinterpret_cast<const float4*>(p.W + static_cast<long long>(n0 + lrow) * p.ldw + k0 + lk);
    As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
    Ws[lk + 0][lrow] = wv.x; Ws[lk + 1][lrow] = wv.y; Ws[lk + 2][lrow] = wv.z; Ws[lk + 3][lrow] = wv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w};
      const float wr[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (ACT != ACT_NONE) {
        v += p.bias[n];
        v = (ACT == ACT_LEAKY) ? (v > 0.f ? v : v * p.slope) : tanhf(v);
      }
      p.out[m * p.ldo + n] = v;
    }
  }
}

// Final discriminator Linear (H/2 -> 1) + sigmoid as a warp-per-row dot product with a shuffle reduction.
__global__ void __launch_bounds__(256) rowdot_f32_kernel(const float* __restrict__ x, long long ldx,
                                                         const float* __restrict__ w, float b, int K,
                                                         long long M, float* __restrict__ logits,
                                                         float* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long num_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long m = warp_global; m < M; m += num_warps) {
    const float* row = x + m * ldx;
    float s = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      const float4 a = *reinterpret_cast<const float4*>(row + k);
      const float4 c = __ldg(reinterpret_cast<const float4*>(w + k));
      s = fmaf(a.x, c.x, s); s = fmaf(a.y, c.y, s); s = fmaf(a.z, c.z, s); s = fmaf(a.w, c.w, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      const float logit = s + b;
      logits[m] = logit;
      if (probs != nullptr) probs[m] = 1.f / (1.f + expf(-logit));
    }
  }
}

// Generator score: F.cosine_similarity(pred, node_emb[tail], dim=1), eps 1e-8 (pro_b_gan_infer.py:202).
__global__ void __launch_bounds__(256) cosine_f32_kernel(const float* __restrict__ pred, long long ldp,
                                                         const float* __restrict__ node_emb, long long n_ent,
                                                         const long long* __restrict__ tails, long long tail_stride,
                                                         int E, long long M, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long num_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long m = warp_global; m < M; m += num_warps) {
    long long t = tails[m * tail_stride];
    if (t < 0) t += n_ent;  // same wrap as the gather
    t = (t < 0 || t >= n_ent) ? 0 : t;
    const float* prow = pred + m * ldp;
    const float* trow = node_emb + t * E;
    float d = 0.f, pp = 0.f, tt = 0.f;
    for (int k = lane * 4; k < E; k += 128) {
      const float4 a = *reinterpret_cast<const float4*>(prow + k);
      const float4 c = __ldg(reinterpret_cast<const float4*>(trow + k));
      d += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
      pp += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
      tt += c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d += __shfl_xor_sync(0xffffffffu, d, o);
      pp += __shfl_xor_sync(0xffffffffu, pp, o);
      tt += __shfl_xor_sync(0xffffffffu, tt, o);
    }
    if (lane == 0) out[m] = d / (fmaxf(sqrtf(pp), 1e-8f) * fmaxf(sqrtf(tt), 1e-8f));
  }
}

}  // namespace pbg
