// gather.cuh -- embedding gather + concat (+ cast) that feeds the first Linear of G and D.
//
//   xg[b, :] = [ node_emb[head_b] | rel_emb[rel_b] | z[b] ]          (generator input, width 2E+Z)
//   xd[b, :] = [ node_emb[head_b] | rel_emb[rel_b] | node_emb[tail_b] ]   (discriminator input, width 3E)
//
// replaces `self.node_emb[heads]`, `self.rel_emb(relations)`, `self.node_emb[tails]` and the torch.cat
// inside the modules (pro_b_gan_infer.py:139-140, :186-188).  HBM-bound, no reuse: one warp per sample,
// each lane moves 128-bit vectors (a 128-wide fp32 row is exactly one float4 per lane), rows are read once
// and feed both outputs.  The gathered values are copied bit-exactly (fp32 out) or rounded once to bf16
// (round-to-nearest-even, identical to torch's .to(bfloat16)).
// Out-of-range ids are clamped to row 0 and raise *err_flag (the host turns it into IndexError).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace pbg {

struct GatherParams {
  const float* node_emb;  // [N, E]
  const float* rel_emb;   // [R, E]
  long long N, R;
  int E, Z;
  // index mode (any of the three may be null -> the matching direct pointer is used)
  const long long* heads; long long head_stride;
  const long long* rels;  long long rel_stride;
  const long long* tails; long long tail_stride;
  // direct mode: already-gathered rows [B, E]
  const float* h; const float* r; const float* t;
  const float* z;         // [B, Z] or null
  void* xg; int ldg;      // generator input  [B, ldg] (ldg >= 2E+Z, padding zero-filled) or null
  void* xd; int ldd;      // discriminator input [B, ldd] (ldd >= 3E) or null
  long long B;
  int* err_flag;
};

template <typename T>
__device__ __forceinline__ void store4(T* dst, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* dst, float4 v) {
  *reinterpret_cast<float4*>(dst) = v;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* dst, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 w;
  w.x = *reinterpret_cast<uint32_t*>(&lo);
  w.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(dst) = w;
}

// streaming 128-bit read: rows are touched once, keep them out of L1 (the epilogues' bias lines live there)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

template <typename T>
__global__ void __launch_bounds__(256) gather_concat_kernel(const GatherParams p) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long num_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int E4 = p.E >> 2, Z4 = p.Z >> 2;
  T* xg = static_cast<T*>(p.xg);
  T* xd = static_cast<T*>(p.xd);

  for (long long b = warp_global; b < p.B; b += num_warps) {
    const float *hrow, *rrow, *trow = nullptr;
    bool bad = false;
    if (p.heads != nullptr) {
      long long i = p.heads[b * p.head_stride];
      if (i < 0) i += p.N;  // node_emb[idx] is tensor indexing: -N..-1 count from the end (pro_b_gan_infer.py:139)
      if (i < 0 || i >= p.N) { bad = true; i = 0; }
      hrow = p.node_emb + i * p.E;
    } else {
      hrow = p.h + b * p.E;
    }
    if (p.rels != nullptr) {
      long long i = p.rels[b * p.rel_stride];
      if (i < 0 || i >= p.R) { bad = true; i = 0; }
      rrow = p.rel_emb + i * p.E;
    } else {
      rrow = p.r + b * p.E;
    }
    // tail ids are range-checked whenever they are given -- a generator-only pass reads them too (cosine against
    // node_emb[tail], pro_b_gan_infer.py:188, :202), and the reference raises IndexError for every method
    if (p.tails != nullptr) {
      long long i = p.tails[b * p.tail_stride];
      if (i < 0) i += p.N;
      if (i < 0 || i >= p.N) { bad = true; i = 0; }
      trow = p.node_emb + i * p.E;
    } else if (xd != nullptr) {
      trow = p.t + b * p.E;
    }
    if (bad && lane == 0) atomicOr(p.err_flag, 1);

    T* g = xg ? xg + b * p.ldg : nullptr;
    T* d = xd ? xd + b * p.ldd : nullptr;
    for (int v = lane; v < E4; v += 32) {
      const float4 hv = ld_stream4(hrow + 4 * v);
      const float4 rv = ld_stream4(rrow + 4 * v);
      if (g) { store4<T>(g + 4 * v, hv); store4<T>(g + p.E + 4 * v, rv); }
      if (d) {
        store4<T>(d + 4 * v, hv);
        store4<T>(d + p.E + 4 * v, rv);
        store4<T>(d + 2 * p.E + 4 * v, ld_stream4(trow + 4 * v));
      }
    }
    if (g) {
      const float* zrow = p.z + b * p.Z;
      for (int v = lane; v < Z4; v += 32) store4<T>(g + 2 * p.E + 4 * v, ld_stream4(zrow + 4 * v));
      for (int v = (2 * p.E + p.Z) / 4 + lane; v < p.ldg / 4; v += 32)
        store4<T>(g + 4 * v, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    if (d) {
      for (int v = (3 * p.E) / 4 + lane; v < p.ldd / 4; v += 32)
        store4<T>(d + 4 * v, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}

// Request staging (pbg_stage_triplets): the same gather + concat + bf16 cast as above for one request, plus an fp32 copy
// of the tail rows in request order (the cosine epilogue of the pass then reads rows, not indices).  Runs on the
// caller's ingest stream BESIDE the pass kernels of earlier requests: small blocks (128 threads, <= 40 registers, no
// shared memory) fit into what a resident pass CTA leaves free on an SM (~11k registers, 1700 threads).  One warp per
// row, every load of the row in flight before the first store.  HBM/L2-bound: (3E + Z) * 4 bytes read and
// (2E + Z + 3E) * 2 + 4E bytes written per row.
struct StageParams {
  GatherParams g;
  float* xt;   // [B, E] fp32 tail rows, or null
};
__global__ void __launch_bounds__(128, 12) stage_rows_kernel(const __grid_constant__ StageParams sp) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long num_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int E = sp.g.E, Z = sp.g.Z, E4 = E >> 2, Z4 = Z >> 2;
  const long long N = sp.g.N;
  __nv_bfloat16* const xg = static_cast<__nv_bfloat16*>(sp.g.xg);
  __nv_bfloat16* const xd = static_cast<__nv_bfloat16*>(sp.g.xd);
  for (long long b = warp_global; b < sp.g.B; b += num_warps) {
    bool bad = false;
    long long hi = sp.g.heads[b * 3], ri = sp.g.rels[b * 3], ti = sp.g.tails[b * 3];   // [B, 3] int64 rows
    if (hi < 0) hi += N;   // node_emb[idx] is tensor indexing: -N..-1 count from the end (pro_b_gan_infer.py:186, :188)
    if (ti < 0) ti += N;
    if (hi < 0 || hi >= N) { bad = true; hi = 0; }
    if (ri < 0 || ri >= sp.g.R) { bad = true; ri = 0; }   // nn.Embedding rejects negative ids (:187)
    if (ti < 0 || ti >= N) { bad = true; ti = 0; }
    if (bad && lane == 0) atomicOr(sp.g.err_flag, 1);
    const float* hrow = sp.g.node_emb + hi * E;
    const float* rrow = sp.g.rel_emb + ri * E;
    const float* trow = sp.g.node_emb + ti * E;
    __nv_bfloat16* g = xg ? xg + b * sp.g.ldg : nullptr;
    __nv_bfloat16* d = xd ? xd + b * sp.g.ldd : nullptr;
    for (int v = lane; v < E4; v += 32) {
      const float4 hv = ld_stream4(hrow + 4 * v);
      const float4 rv = ld_stream4(rrow + 4 * v);
      const float4 tv = ld_stream4(trow + 4 * v);
      if (g) {
        store4<__nv_bfloat16>(g + 4 * v, hv);
        store4<__nv_bfloat16>(g + E + 4 * v, rv);
      }
      if (d) {
        store4<__nv_bfloat16>(d + 4 * v, hv);
        store4<__nv_bfloat16>(d + E + 4 * v, rv);
        store4<__nv_bfloat16>(d + 2 * E + 4 * v, tv);
      }
      if (sp.xt) store4<float>(sp.xt + b * E + 4 * v, tv);
    }
    if (g) {
      for (int v = lane; v < Z4; v += 32) store4<__nv_bfloat16>(g + 2 * E + 4 * v, ld_stream4(sp.g.z + b * Z + 4 * v));
      for (int v = (2 * E + Z) / 4 + lane; v < sp.g.ldg / 4; v += 32)
        store4<__nv_bfloat16>(g + 4 * v, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    if (d)
      for (int v = (3 * E) / 4 + lane; v < sp.g.ldd / 4; v += 32)
        store4<__nv_bfloat16>(d + 4 * v, make_float4(0.f, 0.f, 0.f, 0.f));
  }
}

// fp32 [rows, cols] -> bf16 [rows_p, cols_p], zero padded (weight packing at load time)
__global__ void pack_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols,
                                 int rows_p, int cols_p) {
  const long long total = static_cast<long long>(rows_p) * cols_p;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols_p), c = static_cast<int>(i % cols_p);
    dst[i] = __float2bfloat16_rn((r < rows && c < cols) ? src[static_cast<long long>(r) * cols + c] : 0.f);
  }
}

// fp32 [n] -> fp32 [n_p], zero padded
__global__ void pad_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int n_p) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_p; i += gridDim.x * blockDim.x)
    dst[i] = i < n ? src[i] : 0.f;
}

}  // namespace pbg
