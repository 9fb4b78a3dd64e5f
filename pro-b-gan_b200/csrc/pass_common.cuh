// pass_common.cuh -- pieces shared by the fused pass kernel (pass2_kernel.cuh): item / dependency kinds, the scheduler
// record, memory-ordering helpers, the register-path gather group and the arrival helpers.
#pragma once
#include <cuda.h>
#include "ptx.cuh"
#include "gather.cuh"
#include "tc_common.cuh"

namespace pbg {

enum : int { IT_G_L0 = 0, IT_D_L0 = 1, IT_G_L1 = 2, IT_D_L1 = 3, IT_G_L2 = 4, IT_GATHER = 5, IT_END = 255 };
enum : int { DEP_X = 0, DEP_G0 = 1, DEP_D0 = 2, DEP_G1 = 3, DEP_KINDS = 4 };
enum : int { FIN_G = 0, FIN_D = 1, FIN_KINDS = 2 };
enum : int { PEPI_STORE = 0, PEPI_TANH = 1, PEPI_ROWDOT = 2 };

constexpr int kEpiWarps = 8;
constexpr int kPassThreads = 64 + 32 * kEpiWarps;
constexpr int kTraceSlots = 256;                      // diagnostics slots per CTA (pbg_debug_trace)
constexpr int kTraceItems = 54;                       // slots 16 .. 231; 232 .. 239 and 240 .. 255: epilogue phase stamps / sums


struct PassSched {
  int q_head;   // next pop ticket
  int q_tail;   // next push slot
  int p0_next;  // next phase-0 gather group
  int init;     // 0 -> 1 by the CTA that seeds the queue
  int done;     // CTAs that have finished
  int pad_;
  // SM clock of the last launch, measured by CTA 0: clock64() ticks and globaltimer nanoseconds between the start of
  // its roles and its exit (pbg_last_pass_sm_clock; NVML's clock reading is a sample of a slower loop)
  long long clk_ticks, clk_ns;
};



__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_gpu_add(int* p, int v) {
  asm volatile("red.relaxed.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int atom_relaxed_gpu_add(int* p, int v) {
  int old;
  asm volatile("atom.relaxed.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ int atom_release_gpu_add(int* p, int v) {
  int old;
  asm volatile("atom.release.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_32x32_ptr(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// LeakyReLU for 0 <= slope <= 1 (the host rejects anything else): max(x, slope * x)
__device__ __forceinline__ float leaky_max(float x, float slope) { return fmaxf(x, x * slope); }

// ------------------------------------------------------------------------------------------------ gather item
// One warp gathers a group of 4 consecutive rows: indices first, then every row load in flight, then the stores.
// A gather item is 32 rows = 8 groups (one per epilogue warp); the first row blocks of a pass are gathered by all
// warps of all CTAs before the roles start (phase 0), one group per warp.
// MODE 0: both layouts behind a run-time test; 1: only the E = 128, Z <= 128 register path; 2: only the generic loop
template <int MODE = 0>
__device__ __forceinline__ void pass_gather_group(const GatherParams& g, long long group, int lane) {
  const long long r0 = group * 4;
  __nv_bfloat16* xg = static_cast<__nv_bfloat16*>(g.xg);
  __nv_bfloat16* xd = static_cast<__nv_bfloat16*>(g.xd);
  // lane l < 12 resolves the source row of (row l / 3, operand l % 3): 0 head, 1 relation, 2 tail
  const float* src = nullptr;
  bool bad = false;
  if (lane < 12) {
    const long long row = r0 + lane / 3;
    const int which = lane % 3;
    if (row < g.B) {
      if (which == 0) {
        if (g.heads) { long long i = g.heads[row * g.head_stride]; if (i < 0) i += g.N; if (i < 0 || i >= g.N) { bad = true; i = 0; } src = g.node_emb + i * g.E; }
        else src = g.h + row * g.E;
      } else if (which == 1) {
        if (g.rels) { long long i = g.rels[row * g.rel_stride]; if (i < 0 || i >= g.R) { bad = true; i = 0; } src = g.rel_emb + i * g.E; }
        else src = g.r + row * g.E;
      } else if (g.tails) {   // range-checked whenever tail ids are given: a generator-only pass reads them too (cosine)
        long long i = g.tails[row * g.tail_stride]; if (i < 0) i += g.N; if (i < 0 || i >= g.N) { bad = true; i = 0; } src = g.node_emb + i * g.E;
      } else if (xd != nullptr) {
        src = g.t + row * g.E;
      }
    }
  }
  if (bad) atomicOr(g.err_flag, 1);
  const unsigned long long sp = reinterpret_cast<unsigned long long>(src);
  const int E4 = g.E >> 2, Z4 = g.Z >> 2;
  if (MODE == 1 || (MODE == 0 && E4 == 32 && Z4 <= 32)) {
    float4 hv[4], rv[4], tv[4], zv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* ph = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, sp, j * 3 + 0));
      const float* pr = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, sp, j * 3 + 1));
      const float* pt = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, sp, j * 3 + 2));
      const long long row = r0 + j;
      hv[j] = rv[j] = tv[j] = zv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < g.B) {
        hv[j] = ld_stream4(ph + 4 * lane);
        rv[j] = ld_stream4(pr + 4 * lane);
        if (xd != nullptr) tv[j] = ld_stream4(pt + 4 * lane);
        if (xg != nullptr && lane < Z4) zv[j] = ld_stream4(g.z + row * g.Z + 4 * lane);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long row = r0 + j;
      if (row >= g.B) continue;
      if (xg != nullptr) {
        __nv_bfloat16* o = xg + row * g.ldg;
        store4<__nv_bfloat16>(o + 4 * lane, hv[j]);
        store4<__nv_bfloat16>(o + g.E + 4 * lane, rv[j]);
        if (lane < Z4) store4<__nv_bfloat16>(o + 2 * g.E + 4 * lane, zv[j]);
        for (int v = (2 * g.E + g.Z) / 4 + lane; v < g.ldg / 4; v += 32)
          store4<__nv_bfloat16>(o + 4 * v, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      if (xd != nullptr) {
        __nv_bfloat16* o = xd + row * g.ldd;
        store4<__nv_bfloat16>(o + 4 * lane, hv[j]);
        store4<__nv_bfloat16>(o + g.E + 4 * lane, rv[j]);
        store4<__nv_bfloat16>(o + 2 * g.E + 4 * lane, tv[j]);
        for (int v = (3 * g.E) / 4 + lane; v < g.ldd / 4; v += 32)
          store4<__nv_bfloat16>(o + 4 * v, make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
  } else {
    for (int j = 0; j < 4; ++j) {
      const float* ph = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, sp, j * 3 + 0));
      const float* pr = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, sp, j * 3 + 1));
      const float* pt = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, sp, j * 3 + 2));
      const long long row = r0 + j;
      if (row >= g.B) continue;
      __nv_bfloat16* og = xg ? xg + row * g.ldg : nullptr;
      __nv_bfloat16* od = xd ? xd + row * g.ldd : nullptr;
      for (int v = lane; v < E4; v += 32) {
        const float4 h4 = ld_stream4(ph + 4 * v), r4 = ld_stream4(pr + 4 * v);
        if (og) { store4<__nv_bfloat16>(og + 4 * v, h4); store4<__nv_bfloat16>(og + g.E + 4 * v, r4); }
        if (od) {
          store4<__nv_bfloat16>(od + 4 * v, h4);
          store4<__nv_bfloat16>(od + g.E + 4 * v, r4);
          store4<__nv_bfloat16>(od + 2 * g.E + 4 * v, ld_stream4(pt + 4 * v));
        }
      }
      if (og) {
        for (int v = lane; v < Z4; v += 32) store4<__nv_bfloat16>(og + 2 * g.E + 4 * v, ld_stream4(g.z + row * g.Z + 4 * v));
        for (int v = (2 * g.E + g.Z) / 4 + lane; v < g.ldg / 4; v += 32)
          store4<__nv_bfloat16>(og + 4 * v, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      if (od)
        for (int v = (3 * g.E) / 4 + lane; v < g.ldd / 4; v += 32)
          store4<__nv_bfloat16>(od + 4 * v, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  }
}

// A warp publishes "my part of this item is in global memory": the warp barrier orders every lane's stores before
// lane 0's release-increment at gpu scope (no sequentially-consistent fence, no L1 invalidation).
__device__ __forceinline__ void warp_publish(int* counter, int lane) {
  __syncwarp();
  if (lane == 0) red_release_gpu_add(counter, 1);
}
// Same, returning the previous count.  Release only: the one warp that turns out to be the last arriver issues the
// acquire fence itself before it reads the others' partials (an acquire on every arrival would invalidate L1).
__device__ __forceinline__ int warp_publish_fetch(int* counter, int lane) {
  __syncwarp();
  int old = 0;
  if (lane == 0) old = atom_release_gpu_add(counter, 1);
  return __shfl_sync(0xffffffffu, old, 0);
}


}  // namespace pbg
