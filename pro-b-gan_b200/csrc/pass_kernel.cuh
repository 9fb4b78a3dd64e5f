// pass_kernel.cuh -- the whole generator + discriminator pass as ONE persistent, warp-specialised kernel.
//
//   gather/concat -> G.L0 -> G.L1 -> G.L2 (tanh, cosine)         (pro_b_gan_infer.py:186-188, :201-202)
//                 -> D.L0 -> D.L1 (+ final H/2 -> 1 dot, sigmoid) (pro_b_gan_infer.py:207, :302)
//
// One CTA per SM.  The pass is cut into work items (gather units of 32 rows, GEMM tiles of 128 rows x BLOCK_N
// columns) that flow through a ready queue in global memory: every (buffer, 128-row block) has an arrival counter,
// the warp whose arrival completes a block pushes the tiles that consume it, and CTAs pop tickets from the queue.
// An item is therefore never claimed before its inputs exist -- no CTA sits on an unready tile, nothing depends on
// which CTAs are resident (several passes can share the GPU from different streams), and the order adapts to the
// timing it finds.  Layers hand activations to each other through L2 (bf16, row-major): no grid-wide barrier, no
// launch per layer.
//
//   warp 0      scheduler + TMA producer : pops a ticket, waits for that queue slot to be filled, publishes the item
//                                          to the CTA through a small smem ring, streams A (128x64) and W (BLOCK_N x
//                                          64) boxes into a 4-stage SWIZZLE_128B ring
//   warp 1      MMA issuer              : one thread, tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators in
//                                          TMEM, two accumulator stages of 256 columns
//   warps 2..9  epilogue (8 warps)      : tcgen05.ld -> bias + activation -> swizzled smem staging -> coalesced
//                                          128-bit global stores (bf16 activations / generator output), or the
//                                          row reductions (final discriminator dot, cosine vs the tail embedding)
//                                          written as per-tile partials that the last-arriving warp of the 128-row
//                                          block sums in a fixed order (deterministic), plus the gather items.
#pragma once
#include <cuda.h>
#include "ptx.cuh"
#include "gather.cuh"
#include "gemm_tc.cuh"

namespace pbg {

enum : int { IT_G_L0 = 0, IT_D_L0 = 1, IT_G_L1 = 2, IT_D_L1 = 3, IT_G_L2 = 4, IT_GATHER = 5, IT_END = 255 };
enum : int { DEP_X = 0, DEP_G0 = 1, DEP_D0 = 2, DEP_G1 = 3, DEP_KINDS = 4 };
enum : int { FIN_G = 0, FIN_D = 1, FIN_KINDS = 2 };
enum : int { PEPI_STORE = 0, PEPI_TANH = 1, PEPI_ROWDOT = 2 };

constexpr int kPassStages = 4;
constexpr int kSchedRing = 4;
constexpr int kEpiWarps = 8;
constexpr int kPassThreads = 64 + 32 * kEpiWarps;
constexpr int kTraceSlots = 256;                      // diagnostics slots per CTA (pbg_debug_trace)
constexpr int kTraceItems = 54;                       // slots 16 .. 231; 232 .. 239 and 240 .. 255: epilogue phase stamps / sums
constexpr int kGatherRows = 32;                       // rows per gather item
constexpr int kGatherPerBlock = kBlockM / kGatherRows;  // gather items per 128-row block

struct PassLayer {
  int num_kb;        // K / 64
  int block_n;       // UMMA N of this layer's tiles (64 / 128 / 256)
  int n_tiles;       // output width / block_n
  int epi;           // PEPI_*
  int dep_kind;      // DEP_* counter that gates this layer's A operand
  int out_kind;      // DEP_* counter this layer's stores bump, or -1
  int ldo;           // leading dimension of out (elements)
  int pad_;
  const float* bias;     // [n_tiles * block_n] fp32, zero padded
  __nv_bfloat16* out;    // PEPI_STORE: next layer's A operand
};

struct PassSched {
  int q_head;   // next pop ticket
  int q_tail;   // next push slot
  int p0_next;  // next phase-0 gather group
  int init;     // 0 -> 1 by the CTA that seeds the queue
  int done;     // CTAs that have finished
};

struct alignas(64) PassParams {
  CUtensorMap tm_a[5];  // A operand of layer i: xg0, xd0, actG0, actD0, actG1   (box 64 x 128, SWIZZLE_128B)
  CUtensorMap tm_w[5];  // weights of layer i                                      (box 64 x block_n)
  CUtensorMap tm_o[5];  // PEPI_STORE layers: the activation buffer they write      (box 64 x 32, SWIZZLE_128B)
  PassLayer layer[5];
  GatherParams gather;
  unsigned layer_mask;  // bit i: layer i takes part in this pass (its tensor maps are valid)
  int poll_ns;          // back-off between polls of a dependency counter
  int phase0_groups;    // 4-row gather groups done by all warps before the roles start (a multiple of 32 = whole row blocks)
  unsigned long long* queue;  // ready queue: (kind + 1) | n_blk << 8 | row block << 32 ; 0 = not pushed yet
  int n_total;          // items this launch will push (and pop) in total
  int mb;               // 128-row blocks in this pass
  int p0_blocks;        // row blocks gathered in phase 0 (phase0_groups / 32)
  int gather_ahead;     // gather items run this many row blocks ahead of the first-layer tiles
  int M;                // rows in this pass
  int mb_cap;           // stride of the counter arrays (row blocks)
  float slope;
  PassSched* sched;
  int* ready;           // [DEP_KINDS][mb_cap] arrival counters (zero between launches)
  int* fin;             // [FIN_KINDS][mb_cap]
  // generator output (PEPI_TANH)
  void* gen_out; int out_f32; int n_valid; int ld_gen;
  float* cosine; const float* tail_tab; const long long* tail_idx; long long tail_stride; long long n_ent;
  const float* tail_rows;  // direct mode: already gathered tail rows [M, n_valid] (or nullptr)
  float* part_g;        // [mb][slots_g][3][128]
  int slots_g;
  // discriminator output (PEPI_ROWDOT)
  const float* w3; float b3; float* logits; float* probs;
  float* part_d;        // [mb][slots_d][128]
  int slots_d;
  long long* trace;
};

struct PassSmem {
  static constexpr int kA = kBlockM * kBlockK * 2;
  static constexpr int kB = 256 * kBlockK * 2;
  static constexpr int kStage = kA + kB;
  static constexpr int kStagingOff = kPassStages * kStage;
  static constexpr int kStagingPerWarp = 4096;
  static constexpr int kBarOff = kStagingOff + kEpiWarps * kStagingPerWarp;
  static constexpr int kTotal = kBarOff + 256 + 1024 /*alignment slack*/;
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
        if (g.heads) { long long i = g.heads[row * g.head_stride]; if (i < 0 || i >= g.N) { bad = true; i = 0; } src = g.node_emb + i * g.E; }
        else src = g.h + row * g.E;
      } else if (which == 1) {
        if (g.rels) { long long i = g.rels[row * g.rel_stride]; if (i < 0 || i >= g.R) { bad = true; i = 0; } src = g.rel_emb + i * g.E; }
        else src = g.r + row * g.E;
      } else if (xd != nullptr) {
        if (g.tails) { long long i = g.tails[row * g.tail_stride]; if (i < 0 || i >= g.N) { bad = true; i = 0; } src = g.node_emb + i * g.E; }
        else src = g.t + row * g.E;
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

// ------------------------------------------------------------------------------------------------ ready queue
__device__ __forceinline__ unsigned long long pass_item(int kind, int n_blk, int m_blk) {
  return static_cast<unsigned long long>((kind + 1) | (n_blk << 8)) | (static_cast<unsigned long long>(m_blk) << 32);
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int pass_gather_units(const PassParams& p, int m) {
  const int r = min(kBlockM, p.M - m * kBlockM);
  return (r + kGatherRows - 1) / kGatherRows;
}
__device__ __forceinline__ int pass_dep_target(const PassParams& p, int dep_kind, int m) {
  if (dep_kind == DEP_X) return m < p.p0_blocks ? 32 : pass_gather_units(p, m) * kEpiWarps;
  const int producer = dep_kind == DEP_G0 ? IT_G_L0 : (dep_kind == DEP_D0 ? IT_D_L0 : IT_G_L1);
  return p.layer[producer].n_tiles * kEpiWarps;
}
__device__ __forceinline__ void pass_push_gather(const PassParams& p, int m) {
  const int n = pass_gather_units(p, m);
  const int base = atomicAdd(&p.sched->q_tail, n);
  for (int u = 0; u < n; ++u) st_release_gpu_u64(p.queue + base + u, pass_item(IT_GATHER, u, m));
}
// One thread, after its arrival completed block m of buffer dep_kind: everything the producers wrote is ordered
// before the pushes (acquire fence after the completing RMW, release stores into the queue).
__device__ __forceinline__ void pass_group_done(const PassParams& p, int dep_kind, int m) {
  fence_acq_rel_gpu();
  if (dep_kind == DEP_X) {
    const int ng = (p.layer_mask & (1u << IT_G_L0)) ? p.layer[IT_G_L0].n_tiles : 0;
    const int nd = (p.layer_mask & (1u << IT_D_L0)) ? p.layer[IT_D_L0].n_tiles : 0;
    const int base = atomicAdd(&p.sched->q_tail, ng + nd);
    for (int n = 0; n < ng; ++n) st_release_gpu_u64(p.queue + base + n, pass_item(IT_G_L0, n, m));        // G chain first
    for (int n = 0; n < nd; ++n) st_release_gpu_u64(p.queue + base + ng + n, pass_item(IT_D_L0, n, m));
    if (m >= p.p0_blocks && m + p.gather_ahead < p.mb) pass_push_gather(p, m + p.gather_ahead);
  } else {
    const int kind = dep_kind == DEP_G0 ? IT_G_L1 : (dep_kind == DEP_D0 ? IT_D_L1 : IT_G_L2);
    const int nt = p.layer[kind].n_tiles;
    const int base = atomicAdd(&p.sched->q_tail, nt);
    for (int n = 0; n < nt; ++n) st_release_gpu_u64(p.queue + base + n, pass_item(kind, n, m));
  }
}
// A warp announces that its part of (dep_kind, m) is in global memory; the arrival that completes the block pushes
// the block's consumers.  `async_stores`: the data went out through bulk (async-proxy) stores issued by lane 0.
__device__ __forceinline__ void pass_arrive(const PassParams& p, int dep_kind, int m, int lane, bool async_stores) {
  __syncwarp();
  if (lane == 0) {
    if (async_stores) { tma_store_wait<0>(); fence_proxy_async_all(); }
    const int old = atom_release_gpu_add(p.ready + dep_kind * p.mb_cap + m, 1);
    if (old + 1 == pass_dep_target(p, dep_kind, m)) pass_group_done(p, dep_kind, m);
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kPassThreads, 1) pbg_pass_kernel(const __grid_constant__ PassParams p) {
  using L = PassSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + kPassStages;
  uint64_t* tmem_full = empty_bar + kPassStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* sched_full = tmem_empty + 2;
  uint64_t* sched_empty = sched_full + kSchedRing;
  uint2* ring = reinterpret_cast<uint2*>(sched_empty + kSchedRing);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ring + kSchedRing);
  int* last_flag = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 5; ++i) {
      if (p.layer_mask & (1u << i)) {
        prefetch_tmap(&p.tm_a[i]);
        prefetch_tmap(&p.tm_w[i]);
        if (p.layer[i].epi == PEPI_STORE) prefetch_tmap(&p.tm_o[i]);
      }
    }
    for (int s = 0; s < kPassStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps);
    }
    for (int s = 0; s < kSchedRing; ++s) {
      mbar_init(&sched_full[s], 1);
      mbar_init(&sched_empty[s], 1 + kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // diagnostics: 256 slots per CTA = 16 header slots + 60 items x 4 stamps (claim, dependency met, accumulator
  // ready, epilogue done), all clock64 of this SM; header slot 14 holds %globaltimer at slot 0's clock64
  long long* tr = p.trace ? p.trace + kTraceSlots * blockIdx.x : nullptr;
  if (tr && threadIdx.x == 0) { tr[0] = clock64(); tr[14] = static_cast<long long>(globaltimer_ns()); }

  // ---- queue seeding: the first CTA to get here pushes the gather items of the first row blocks past phase 0
  if (threadIdx.x == 0 && p.p0_blocks < p.mb && atomicCAS(&p.sched->init, 0, 1) == 0) {
    for (int m = p.p0_blocks; m < min(p.mb, p.p0_blocks + p.gather_ahead); ++m) pass_push_gather(p, m);
  }
  // ---- phase 0: warps of all CTAs claim 4-row gather groups of the first row blocks, so that the first-layer tiles
  // of a small batch wait for one index -> row -> store round trip instead of a queue of gather items
  for (;;) {
    int g = 0;
    if (lane == 0) g = atomicAdd(&p.sched->p0_next, 1);
    g = __shfl_sync(0xffffffffu, g, 0);
    if (g >= p.phase0_groups) break;
    pass_gather_group(p.gather, g, lane);
    pass_arrive(p, DEP_X, g >> 5, lane, false);
  }
  if (tr && threadIdx.x == 0) tr[5] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------ scheduler + TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, slot = 0, sphase = 0;
      long long w_dep = 0, w_empty = 0, n_items = 0;
      for (;;) {
        // pop a ticket (after the previous item's loads are all issued: the MMA still has a full ring to chew on)
        const int ticket = atomicAdd(&p.sched->q_head, 1);
        uint2 it = make_uint2(IT_END, 0u);
        if (ticket < p.n_total) {
          const long long t = tr ? clock64() : 0;
          unsigned long long d;
          uint32_t spins = 0;
          while ((d = ld_relaxed_gpu_u64(p.queue + ticket)) == 0ull) {
            __nanosleep(p.poll_ns);
            if (++spins > 4000000u) { printf("pbg: ready-queue wait timed out (block %d ticket %d)\n", blockIdx.x, ticket); __trap(); }
          }
          if (tr) w_dep += clock64() - t;
          // The item was pushed after its inputs were complete (release stores after an acquire of the block's
          // counter).  The inputs are read only by TMA (async proxy, from L2), so no generic-proxy acquire fence is
          // issued -- it would invalidate this SM's L1 (the epilogue's bias lines); the proxy fence orders the
          // queue read before the bulk loads below.
          fence_proxy_async_all();
          it = make_uint2((static_cast<uint32_t>(d) & 0xff) - 1u | (static_cast<uint32_t>(d) & 0xff00u), static_cast<uint32_t>(d >> 32));
        }
        mbar_wait(&sched_empty[slot], sphase ^ 1);
        ring[slot] = it;
        mbar_arrive(&sched_full[slot]);
        if (++slot == kSchedRing) { slot = 0; sphase ^= 1; }
        const int kind = it.x & 0xff;
        if (kind == IT_END) break;
        long long* ti = (tr && n_items < kTraceItems) ? tr + 16 + 4 * n_items : nullptr;
        if (ti) { ti[0] = (clock64() << 20) | (static_cast<long long>(it.y & 0xfff) << 8) | kind; ti[1] = 0; }
        ++n_items;
        if (kind == IT_GATHER) {
          // no loads to issue: wait until the epilogue warps have picked the item up before popping another ticket,
          // so that an idle scheduler cannot hoard gather items while other CTAs have none
          const uint32_t used = slot == 0 ? kSchedRing - 1 : slot - 1;
          mbar_wait(&sched_empty[used], slot == 0 ? sphase ^ 1 : sphase);
          continue;
        }
        const int n_blk = (it.x >> 8) & 0xff;
        const int m_blk = static_cast<int>(it.y);
        const PassLayer& ly = p.layer[kind];
        const uint32_t bytes = L::kA + static_cast<uint32_t>(ly.block_n) * kBlockK * 2;
        for (int kb = 0; kb < ly.num_kb; ++kb) {
          if (tr) { const long long t = clock64(); mbar_wait(&empty_bar[stage], phase ^ 1); w_empty += clock64() - t; }
          else mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * L::kStage;
          uint8_t* sb = sa + L::kA;
          mbar_arrive_expect_tx(&full_bar[stage], bytes);
          tma_load_2d(sb, &p.tm_w[kind], &full_bar[stage], kb * kBlockK, n_blk * ly.block_n);  // weights: no dependency
          tma_load_2d(sa, &p.tm_a[kind], &full_bar[stage], kb * kBlockK, m_blk * kBlockM);
          if (++stage == kPassStages) { stage = 0; phase ^= 1; }
        }
      }
      if (tr) { tr[1] = w_empty; tr[2] = clock64(); tr[11] = w_dep; tr[12] = n_items; }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, slot = 0, sphase = 0;
      long long w_full = 0, w_tmem = 0, n_kb = 0;
      for (;;) {
        mbar_wait(&sched_full[slot], sphase);
        const uint2 it = ring[slot];
        mbar_arrive(&sched_empty[slot]);
        if (++slot == kSchedRing) { slot = 0; sphase ^= 1; }
        const int kind = it.x & 0xff;
        if (kind == IT_END) break;
        if (kind == IT_GATHER) continue;
        const PassLayer& ly = p.layer[kind];
        const uint32_t idesc = make_idesc_bf16(kBlockM, static_cast<uint32_t>(ly.block_n));
        if (tr) { const long long t = clock64(); mbar_wait(&tmem_empty[acc], acc_phase ^ 1); w_tmem += clock64() - t; }
        else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < ly.num_kb; ++kb) {
          if (tr) { const long long t = clock64(); mbar_wait(&full_bar[stage], phase); w_full += clock64() - t; ++n_kb; }
          else mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * L::kStage);
          const uint64_t da = make_kmajor_sw128_desc(sa);
          const uint64_t db = make_kmajor_sw128_desc(sa + L::kA);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (++stage == kPassStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (tr) { tr[3] = w_full; tr[4] = w_tmem; tr[6] = clock64(); tr[9] = n_kb; }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ epilogue + gather warps
    const int wep = warp - 2;          // 0..7
    const int q = warp & 3;            // TMEM lane quarter this warp may read
    const int half = wep >> 2;         // which half of a tile's column chunks this warp takes
    uint8_t* st = smem + L::kStagingOff + wep * L::kStagingPerWarp;
    uint32_t acc = 0, acc_phase = 0, slot = 0, sphase = 0;
    long long w_acc = 0, busy = 0, ph_ld = 0, ph_math = 0, ph_n = 0;
    int item_no = 0;
    int pend_kind = -1, pend_m = 0;  // the previous tile's block: announced once its bulk stores have completed
    for (;;) {
      // Deferred arrival: the previous tile's bulk stores were issued a while ago; waiting for their completion is
      // cheapest while this warp would idle anyway.  It must not be postponed past a point where the warp blocks on
      // the ring: the next ring item may be exactly what this arrival releases.
      if (pend_kind >= 0) { pass_arrive(p, pend_kind, pend_m, lane, true); pend_kind = -1; }
      mbar_wait(&sched_full[slot], sphase);
      const uint2 it = ring[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sched_empty[slot]);
      if (++slot == kSchedRing) { slot = 0; sphase ^= 1; }
      const int kind = it.x & 0xff;
      if (kind == IT_END) break;
      const int n_blk = (it.x >> 8) & 0xff;
      const int m_blk = static_cast<int>(it.y);
      long long* ti = (tr && threadIdx.x == 64 && item_no < kTraceItems) ? tr + 16 + 4 * item_no : nullptr;
      ++item_no;
      if (kind == IT_GATHER) {
        if (ti) ti[2] = clock64();
        pass_gather_group(p.gather, (static_cast<long long>(m_blk) * kGatherPerBlock + n_blk) * kEpiWarps + wep, lane);
        pass_arrive(p, DEP_X, m_blk, lane, false);
        if (ti) ti[3] = clock64();
        continue;
      }
      const PassLayer& ly = p.layer[kind];
      const int row_in_blk = q * 32 + lane;
      const long long grow = static_cast<long long>(m_blk) * kBlockM + row_in_blk;
      const bool row_ok = grow < p.M;
      const int n0 = n_blk * ly.block_n;
      {
        const long long t = (tr && lane == 0) ? clock64() : 0;
        mbar_wait(&tmem_full[acc], acc_phase);
        if (tr && lane == 0) { w_acc += clock64() - t; }
        if (ti) ti[2] = clock64();
      }
      const long long t_busy0 = (tr && lane == 0) ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256;

      if (ly.epi == PEPI_STORE) {
        // ---- bias + LeakyReLU -> bf16 -> swizzled staging tile -> one TMA store per 32-row x 64-column chunk
        const int n_chunks = ly.block_n >> 6;
        const bool tp = tr && threadIdx.x == 64;
        long long tq0 = 0, tq1 = 0, tq2 = 0;
        const float slope = p.slope;
        const float* const bias_tile = ly.bias + n0;
        const int row0 = m_blk * kBlockM + q * 32;
        for (int c = half; c < n_chunks; c += 2) {
          if (tp) tq0 = clock64();
          const float4* b4 = reinterpret_cast<const float4*>(bias_tile + c * 64);
          uint32_t v[32];
          float4 bv[8];
          tmem_ld_32x32_ptr(taddr + c * 64, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = __ldg(b4 + j);
          // the staging tile is free once the previous chunk's bulk store has read it
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
          tmem_ld_wait();
          if (tp) tq1 = clock64();
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t vn[32];
            float4 bn[8];
            if (h2 == 0) {  // second half of the chunk: in flight while the first half is converted
              tmem_ld_32x32_ptr(taddr + c * 64 + 32, vn);
#pragma unroll
              for (int j = 0; j < 8; ++j) bn[j] = __ldg(b4 + 8 + j);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float4 ba = bv[2 * t], bb = bv[2 * t + 1];
              uint4 w;
              w.x = pack_bf16x2(leaky_max(__uint_as_float(v[8 * t + 0]) + ba.x, slope), leaky_max(__uint_as_float(v[8 * t + 1]) + ba.y, slope));
              w.y = pack_bf16x2(leaky_max(__uint_as_float(v[8 * t + 2]) + ba.z, slope), leaky_max(__uint_as_float(v[8 * t + 3]) + ba.w, slope));
              w.z = pack_bf16x2(leaky_max(__uint_as_float(v[8 * t + 4]) + bb.x, slope), leaky_max(__uint_as_float(v[8 * t + 5]) + bb.y, slope));
              w.w = pack_bf16x2(leaky_max(__uint_as_float(v[8 * t + 6]) + bb.z, slope), leaky_max(__uint_as_float(v[8 * t + 7]) + bb.w, slope));
              const int piece = h2 * 4 + t;
              *reinterpret_cast<uint4*>(st + lane * 128 + ((piece ^ (lane & 7)) << 4)) = w;
            }
            if (h2 == 0) {
              tmem_ld_wait();
              if (c + 2 >= n_chunks) {  // this warp's last read of the accumulator stage
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = vn[j];
#pragma unroll
              for (int j = 0; j < 8; ++j) bv[j] = bn[j];
            }
          }
          fence_proxy_async_smem();  // generic-proxy writes of the staging tile -> visible to the bulk store
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&p.tm_o[kind], st, n0 + c * 64, row0);
            tma_store_commit();
          }
          if (tp) { tq2 = clock64(); ph_ld += tq1 - tq0; ph_math += tq2 - tq1; ph_n += 1; }
        }
        if (half >= n_chunks) {  // narrow tile: this warp had no chunk, still has to release the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        pend_kind = ly.out_kind; pend_m = m_blk;
      } else if (ly.epi == PEPI_ROWDOT) {
        // ---- bias + LeakyReLU, dotted with the final [H/2 -> 1] weight; one partial per 64 columns, summed in a
        //      fixed order by the last warp to arrive for this row block (deterministic, tile-width independent)
        const int n_chunks = ly.block_n >> 6;
        const float slope = p.slope;
        float* part = p.part_d + (static_cast<size_t>(m_blk) * p.slots_d) * kBlockM;
        for (int c = half; c < n_chunks; c += 2) {
          float rowdot = 0.f;
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t v[32];
            float4 bv[8], wv[8];
            tmem_ld_32x32_ptr(taddr + c * 64 + h2 * 32, v);
            const float4* b4 = reinterpret_cast<const float4*>(ly.bias + n0 + c * 64 + h2 * 32);
            const float4* w4 = reinterpret_cast<const float4*>(p.w3 + n0 + c * 64 + h2 * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) { bv[j] = __ldg(b4 + j); wv[j] = __ldg(w4 + j); }
            tmem_ld_wait();
            if (h2 == 1 && c + 2 >= n_chunks) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = bv[j], w = wv[j];
              rowdot = fmaf(leaky_max(__uint_as_float(v[4 * j + 0]) + b.x, slope), w.x, rowdot);
              rowdot = fmaf(leaky_max(__uint_as_float(v[4 * j + 1]) + b.y, slope), w.y, rowdot);
              rowdot = fmaf(leaky_max(__uint_as_float(v[4 * j + 2]) + b.z, slope), w.z, rowdot);
              rowdot = fmaf(leaky_max(__uint_as_float(v[4 * j + 3]) + b.w, slope), w.w, rowdot);
            }
          }
          part[((n0 >> 6) + c) * kBlockM + row_in_blk] = rowdot;
        }
        if (half >= n_chunks) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        const int old = warp_publish_fetch(p.fin + FIN_D * p.mb_cap + m_blk, lane);
        if (old == ly.n_tiles * kEpiWarps - 1) {
          fence_acq_rel_gpu();
          __syncwarp();
          for (int r = lane; r < kBlockM; r += 32) {
            const long long gr = static_cast<long long>(m_blk) * kBlockM + r;
            float s = 0.f;
            for (int k = 0; k < p.slots_d; ++k) s += __ldcg(part + k * kBlockM + r);
            if (gr < p.M) {
              const float logit = s + p.b3;
              p.logits[gr] = logit;
              if (p.probs != nullptr) p.probs[gr] = 1.f / (1.f + __expf(-logit));
            }
          }
        }
      } else {
        // ---- PEPI_TANH: bias + tanh -> generator output (fp32 / bf16), optional cosine vs the tail embedding
        const bool want_cos = p.cosine != nullptr;
        const bool want_out = p.gen_out != nullptr;
        const int n_chunks = ly.block_n >> 5;  // 32-column chunks
        const float* trow = nullptr;
        if (want_cos && row_ok) {
          if (p.tail_idx != nullptr) {
            long long tid = p.tail_idx[grow * p.tail_stride];
            tid = (tid < 0 || tid >= p.n_ent) ? 0 : tid;  // the gather item has already flagged it
            trow = p.tail_tab + tid * p.n_valid;
          } else {
            trow = p.tail_rows + grow * p.n_valid;
          }
        }
        const unsigned long long trow_bits = reinterpret_cast<unsigned long long>(trow);
        float* part = p.part_g + (static_cast<size_t>(m_blk) * p.slots_g) * 3 * kBlockM;
        for (int c = half; c < n_chunks; c += 2) {
          float cs_dot = 0.f, cs_pp = 0.f, cs_tt = 0.f;
          uint32_t v[32];
          tmem_ld_32x32_ptr(taddr + c * 32, v);
          const int col0 = n0 + c * 32;
          const float4* b4 = reinterpret_cast<const float4*>(ly.bias + col0);
          float4 bv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = __ldg(b4 + j);
          // tail rows of this chunk: one coalesced 16-byte piece per lane and iteration, all in flight together
          float4 tvv[8];
          if (want_cos && col0 < p.n_valid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + (lane >> 3), t = lane & 7;
              const float* tp2 = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, trow_bits, r));
              tvv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (tp2 != nullptr && col0 + t * 4 < p.n_valid) tvv[i] = __ldg(reinterpret_cast<const float4*>(tp2 + col0 + t * 4));
            }
          }
          tmem_ld_wait();
          if (c + 2 >= n_chunks) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = bv[j];
            f[4 * j + 0] = tanh_fast(__uint_as_float(v[4 * j + 0]) + b.x);
            f[4 * j + 1] = tanh_fast(__uint_as_float(v[4 * j + 1]) + b.y);
            f[4 * j + 2] = tanh_fast(__uint_as_float(v[4 * j + 2]) + b.z);
            f[4 * j + 3] = tanh_fast(__uint_as_float(v[4 * j + 3]) + b.w);
          }
          if (col0 >= p.n_valid) continue;  // padding columns (warp-uniform)
          if (want_cos) {
            // tail rows: coalesced into the staging tile, then read back row-per-thread
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + (lane >> 3), t = lane & 7;
              *reinterpret_cast<float4*>(st + r * 128 + ((t ^ (r & 7)) << 4)) = tvv[i];
            }
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const float4 x = *reinterpret_cast<const float4*>(st + lane * 128 + ((t ^ (lane & 7)) << 4));
              if (col0 + t * 4 < p.n_valid) {
                cs_dot += f[4 * t] * x.x + f[4 * t + 1] * x.y + f[4 * t + 2] * x.z + f[4 * t + 3] * x.w;
                cs_pp += f[4 * t] * f[4 * t] + f[4 * t + 1] * f[4 * t + 1] + f[4 * t + 2] * f[4 * t + 2] + f[4 * t + 3] * f[4 * t + 3];
                cs_tt += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
              }
            }
            __syncwarp();
            float* mine = part + (col0 >> 5) * 3 * kBlockM;  // one partial triple per 32-column chunk
            mine[row_in_blk] = cs_dot;
            mine[kBlockM + row_in_blk] = cs_pp;
            mine[2 * kBlockM + row_in_blk] = cs_tt;
          }
          if (want_out) {
            if (p.out_f32) {
#pragma unroll
              for (int t = 0; t < 8; ++t)
                *reinterpret_cast<float4*>(st + lane * 128 + ((t ^ (lane & 7)) << 4)) =
                    make_float4(f[4 * t], f[4 * t + 1], f[4 * t + 2], f[4 * t + 3]);
              __syncwarp();
              float* obase = static_cast<float*>(p.gen_out) + col0;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + (lane >> 3), t = lane & 7;
                const float4 x = *reinterpret_cast<const float4*>(st + r * 128 + ((t ^ (r & 7)) << 4));
                const long long gr = static_cast<long long>(m_blk) * kBlockM + q * 32 + r;
                if (gr < p.M && col0 + t * 4 < p.n_valid) *reinterpret_cast<float4*>(obase + gr * p.ld_gen + t * 4) = x;
              }
            } else {
              // bf16: 32 columns = 64 B per row, 4 x 16 B pieces, swizzled by (row >> 1) & 3
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                uint4 w;
                w.x = pack_bf16x2(f[8 * t + 0], f[8 * t + 1]);
                w.y = pack_bf16x2(f[8 * t + 2], f[8 * t + 3]);
                w.z = pack_bf16x2(f[8 * t + 4], f[8 * t + 5]);
                w.w = pack_bf16x2(f[8 * t + 6], f[8 * t + 7]);
                *reinterpret_cast<uint4*>(st + lane * 64 + ((t ^ ((lane >> 1) & 3)) << 4)) = w;
              }
              __syncwarp();
              __nv_bfloat16* obase = static_cast<__nv_bfloat16*>(p.gen_out) + col0;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int r = i * 8 + (lane >> 2), t = lane & 3;
                const uint4 w = *reinterpret_cast<const uint4*>(st + r * 64 + ((t ^ ((r >> 1) & 3)) << 4));
                const long long gr = static_cast<long long>(m_blk) * kBlockM + q * 32 + r;
                if (gr < p.M && col0 + t * 8 < p.n_valid) *reinterpret_cast<uint4*>(obase + gr * p.ld_gen + t * 8) = w;
              }
            }
            __syncwarp();
          }
        }
        if (half >= n_chunks) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (want_cos) {
          const int old = warp_publish_fetch(p.fin + FIN_G * p.mb_cap + m_blk, lane);
          if (old == ly.n_tiles * kEpiWarps - 1) {
            fence_acq_rel_gpu();
            __syncwarp();
            for (int r = lane; r < kBlockM; r += 32) {
              const long long gr = static_cast<long long>(m_blk) * kBlockM + r;
              float d = 0.f, pp = 0.f, tt = 0.f;
              for (int k = 0; k < p.slots_g; ++k) {
                d += __ldcg(part + (k * 3 + 0) * kBlockM + r);
                pp += __ldcg(part + (k * 3 + 1) * kBlockM + r);
                tt += __ldcg(part + (k * 3 + 2) * kBlockM + r);
              }
              // F.cosine_similarity(pred, t, dim=1), eps = 1e-8 on each norm (pro_b_gan_infer.py:202)
              if (gr < p.M) p.cosine[gr] = d / (fmaxf(sqrtf(pp), 1e-8f) * fmaxf(sqrtf(tt), 1e-8f));
            }
          }
        }
      }
      if (tr && lane == 0) busy += clock64() - t_busy0;
      if (ti) ti[3] = clock64();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (tr && threadIdx.x == 64) { tr[7] = w_acc; tr[8] = clock64(); tr[10] = busy; tr[240] = ph_ld; tr[241] = ph_math; tr[243] = ph_n; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  // the last CTA to finish re-arms the scheduler and zeroes the arrival counters for the next launch
  if (threadIdx.x == 0) {
    __threadfence();
    const int old = atomicAdd(&p.sched->done, 1);
    *last_flag = (old == static_cast<int>(gridDim.x) - 1);
  }
  __syncthreads();
  if (*last_flag) {
    for (int k = 0; k < DEP_KINDS; ++k)
      for (int i = threadIdx.x; i < p.mb; i += blockDim.x) p.ready[k * p.mb_cap + i] = 0;
    for (int k = 0; k < FIN_KINDS; ++k)
      for (int i = threadIdx.x; i < p.mb; i += blockDim.x) p.fin[k * p.mb_cap + i] = 0;
    for (int i = threadIdx.x; i < p.n_total; i += blockDim.x) p.queue[i] = 0ull;
    if (threadIdx.x == 0) { p.sched->q_head = 0; p.sched->q_tail = 0; p.sched->p0_next = 0; p.sched->init = 0; p.sched->done = 0; }
    if (tr && threadIdx.x == 0) tr[13] = clock64();
  }
}

}  // namespace pbg
