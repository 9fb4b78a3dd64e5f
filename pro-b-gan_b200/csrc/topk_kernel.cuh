// topk_kernel.cuh -- cosine scoring of query embeddings against the whole entity table with a fused candidate filter:
//
//   pred_norm = F.normalize(pred, dim=-1); entity_norm = F.normalize(node_emb, dim=-1)
//   similarities = pred_norm @ entity_norm.T;  top_scores, top_indices = similarities.topk(k, dim=1)
//                                               (pro_b_gan_infer.py:146-151, :231-236)
//
// The [B, N] similarity matrix (1 GiB at B = 4096, N = 65536) never exists.  Three steps:
//   1. prepare   : row norms; bf16 copies of the normalised table / queries, zero padded to 256 rows   (HBM-bound)
//   2. filter    : bf16 tensor-core scores, CTA pairs (tcgen05.mma.cta_group::2, M = 256 queries x N = 256 entities
//                  per MMA group, K = E = 128), the query tile stationary in shared memory, entity tiles streamed by
//                  TMA; the matrix goes TMEM -> registers -> compare only.  Per query row and per (entity range,
//                  column half) the epilogue first tracks the 6 best scores of a sample (the range's first 16 tiles)
//                  in registers, branch-free, as packed (score | position) integer keys; the 6th becomes the list's
//                  cut-off and every later score above it is appended to a thread-private shared-memory list
//                  (about 19 entries; a warp vote keeps the common case at three instructions per score)
//   3. rescore   : exact fp32 cosine of every candidate, top-k among them (ties: lower index first), and a proof
//                  obligation per row: the k-th exact score must beat every list's cut-off by more than the bf16 error
//                  bound, else the row is flagged and re-done by an exact scan of the table (rare; always used for
//                  k > 16 or E != 128)
// so the returned indices are those of an exact fp32 evaluation, not of the bf16 scores.
#pragma once
#include <cuda.h>
#include "ptx.cuh"
#include "gemm_tc.cuh"
#include "pass_common.cuh"

namespace pbg {

constexpr int kTkStages = 4;          // entity tiles in flight per CTA (32 KB each: 128 entities x 128 dims bf16)
constexpr int kTkCand = 32;           // candidates kept per (row, entity range, column half)
constexpr int kTkMaxRanges = 16;
constexpr int kTkMaxK = 16;           // largest k the filter path proves exact; above it the exact scan runs
constexpr int kTkSample = 16;         // tiles of a range whose 6 best scores set the list's cut-off
constexpr float kTkErrBound = 0.009f; // |bf16 score - exact score| <= 2^-8 (unit vectors, 8-bit mantissas) + key truncation + slack

struct alignas(64) TopkParams {
  CUtensorMap tm_q;      // normalised queries bf16 [Bpad, 128]: box 64 x 128 rows
  CUtensorMap tm_t;      // normalised table   bf16 [Npad, 128]: box 64 x 128 rows (one CTA's half of a 256-entity tile)
  int n_rb;              // 256-row query blocks
  int n_ranges;          // entity ranges (work item = query block x range)
  int tiles_per_range;   // 256-entity tiles per range
  int n_tiles;           // Npad / 256
  int n_items;           // n_rb * n_ranges
  long long N;           // valid entities
  float* cand_score;     // [Bpad][n_ranges * 2][kTkCand]
  int* cand_idx;
  float* cand_tau;       // [Bpad][n_ranges * 2]: the list's cut-off (its smallest kept approximate score)
};

struct TkSmem {
  static constexpr int kQ = 2 * 128 * kBlockK * 2;        // this CTA's 128 query rows, 2 k-blocks
  static constexpr int kT = 2 * 128 * kBlockK * 2;        // this CTA's 128 entities of a tile, 2 k-blocks
  static constexpr int kStageOff = kQ;
  static constexpr int kListOff = kStageOff + kTkStages * kT;
  static constexpr int kListPerWarp = kTkCand * 32 * 8;   // [entry][lane] scores, then indices
  static constexpr int kBarOff = kListOff + kEpiWarps * kListPerWarp;
  static constexpr int kTotal = kBarOff + 256 + 1024;
};
static_assert(TkSmem::kTotal <= 232448, "topk: shared memory budget");

// ------------------------------------------------------------------------------------------------ 1. prepare
// One warp per row: inv = 1 / max(||x||, 1e-12) (F.normalize's eps), bf16(x * inv) into a [rows_pad, E] matrix whose
// padding rows are zero.
__global__ void __launch_bounds__(256) topk_prepare_kernel(const float* __restrict__ x, long long rows, long long rows_pad, int E,
                                                           __nv_bfloat16* __restrict__ xn, float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long r = warp0; r < rows_pad; r += nwarp) {
    if (r >= rows) {
      for (int c = lane * 4; c < E; c += 128) store4<__nv_bfloat16>(xn + r * E + c, make_float4(0.f, 0.f, 0.f, 0.f));
      continue;
    }
    float ss = 0.f;
    for (int c = lane * 4; c < E; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(x + r * E + c);
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float iv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    if (lane == 0) inv[r] = iv;
    for (int c = lane * 4; c < E; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(x + r * E + c);
      store4<__nv_bfloat16>(xn + r * E + c, make_float4(v.x * iv, v.y * iv, v.z * iv, v.w * iv));
    }
  }
}

// ------------------------------------------------------------------------------------------------ 2. filter
// Thread-private candidate list in shared memory, laid out [entry][lane] so that a warp's accesses never conflict.
// Insert: replace the smallest kept score, return the new cut-off.  Out of line: it runs a few hundred times per list,
// the compare in front of it 8192 times per tile.
__device__ __noinline__ float tk_insert(float* sc, int* ix, int lane, float s, int idx) {
  int at = 0; float mn = sc[lane];
#pragma unroll
  for (int e = 1; e < kTkCand; ++e) { const float v = sc[e * 32 + lane]; if (v < mn) { mn = v; at = e; } }
  sc[at * 32 + lane] = s; ix[at * 32 + lane] = idx;
  mn = sc[lane];
#pragma unroll
  for (int e = 1; e < kTkCand; ++e) mn = fminf(mn, sc[e * 32 + lane]);
  return mn;
}

// A warp reaches this when some lane's score beat its cut-off: the passing lanes append to their lists (or, list
// full, replace its smallest and raise the cut-off).  Returns (entries << 32) | cut-off bits.  Out of line: about one
// score in eight gets here, and the inlined alternative is 128 copies of this body per tile.
__device__ __noinline__ unsigned long long tk_pass(float* sc, int* ix, int lane, bool pass, float sv, int idx, int cnt, float tau) {
  if (pass) {
    if (cnt < kTkCand) {
      sc[cnt * 32 + lane] = sv; ix[cnt * 32 + lane] = idx; ++cnt;
      if (cnt == kTkCand) {
        float mn = sc[lane];
#pragma unroll
        for (int e = 1; e < kTkCand; ++e) mn = fminf(mn, sc[e * 32 + lane]);
        tau = fmaxf(tau, mn);
      }
    } else {
      tau = tk_insert(sc, ix, lane, sv, idx);
    }
  }
  return (static_cast<unsigned long long>(static_cast<unsigned>(cnt)) << 32) | __float_as_uint(tau);
}

__device__ __noinline__ float tk_list_min(const float* sc, int lane) {
  float mn = sc[lane];
#pragma unroll
  for (int e = 1; e < kTkCand; ++e) mn = fminf(mn, sc[e * 32 + lane]);
  return mn;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPassThreads, 1)
pbg_topk_filter_kernel(const __grid_constant__ TopkParams p) {
  using L = TkSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
  uint64_t* empty_bar = full_bar + kTkStages;
  uint64_t* tmem_full = empty_bar + kTkStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* q_full = tmem_empty + 2;
  uint64_t* q_empty = q_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = static_cast<int>(blockIdx.x) >> 1;
  const int npairs = static_cast<int>(gridDim.x) >> 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&p.tm_q);
    prefetch_tmap(&p.tm_t);
    for (int s = 0; s < kTkStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 2 * kEpiWarps); }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lead_tmem_empty = mapa_u32(smem_u32(tmem_empty), 0);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs, their halves)
    if (lane == 0) {
      const uint32_t lead_full = mapa_u32(smem_u32(full_bar), 0);
      const uint32_t lead_q_full = mapa_u32(smem_u32(q_full), 0);
      uint32_t stage = 0, phase = 0, qphase = 0;
      for (int item = pair; item < p.n_items; item += npairs) {
        const int rb = item / p.n_ranges, rg = item % p.n_ranges;
        // the query tile of this item (the previous item's MMAs have finished with the buffer)
        mbar_wait(q_empty, qphase ^ 1);
        if (leader) mbar_arrive_expect_tx(q_full, 2u * L::kQ);
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d_pair(smem + kb * (L::kQ / 2), &p.tm_q, lead_q_full, kb * kBlockK, rb * 256 + static_cast<int>(rank) * 128);
        qphase ^= 1;
        const int t0 = rg * p.tiles_per_range, t1 = min(p.n_tiles, t0 + p.tiles_per_range);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * L::kT);
          uint8_t* st = smem + L::kStageOff + stage * L::kT;
          for (int kb = 0; kb < 2; ++kb)
            tma_load_2d_pair(st + kb * (L::kT / 2), &p.tm_t, lead_full + stage * 8, kb * kBlockK, t * 256 + static_cast<int>(rank) * 128);
          if (++stage == kTkStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, 256);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, qphase = 0;
      for (int item = pair; item < p.n_items; item += npairs) {
        const int rg = item % p.n_ranges;
        mbar_wait(q_full, qphase);
        qphase ^= 1;
        tc_fence_after();
        const int t0 = rg * p.tiles_per_range, t1 = min(p.n_tiles, t0 + p.tiles_per_range);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * 256;
          const uint32_t sq = smem_u32(smem), stt = smem_u32(smem + L::kStageOff + stage * L::kT);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t da = make_kmajor_sw128_desc(sq + kb * (L::kQ / 2));
            const uint64_t db = make_kmajor_sw128_desc(stt + kb * (L::kT / 2));
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) umma_bf16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit_pair(&empty_bar[stage], 3);
          umma_commit_pair(&tmem_full[acc], 3);
          if (++stage == kTkStages) { stage = 0; phase ^= 1; }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        umma_commit_pair(q_empty, 3);   // the query tile may be overwritten once these MMAs are done
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------ filter warps (both CTAs)
    const int wep = warp - 2, q = warp & 3, half = wep >> 2;
    float* lsc = reinterpret_cast<float*>(smem + L::kListOff + wep * L::kListPerWarp);
    int* lix = reinterpret_cast<int*>(lsc + kTkCand * 32);
    uint32_t acc = 0, acc_phase = 0;
    for (int item = pair; item < p.n_items; item += npairs) {
      const int rb = item / p.n_ranges, rg = item % p.n_ranges;
#pragma unroll
      for (int e = 0; e < kTkCand; ++e) { lsc[e * 32 + lane] = -3.0e38f; lix[e * 32 + lane] = -1; }
      const int t0 = rg * p.tiles_per_range, t1 = min(p.n_tiles, t0 + p.tiles_per_range);
      const int ts_end = min(t1, t0 + kTkSample);
      int m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0, m6 = 0;   // sample phase: the 6 largest keys, sorted
      float tau = 0.f;                                       // main phase: cut-off
      int cnt = 0;                                           // entries in the list
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * 256 + half * 128;
        const int ent0 = t * 256 + half * 128;
        const bool sampling = t < ts_end;
        const int tsb = (t - t0) << 7;
        uint32_t va[32], vb[32];
        // one 32-column group: sample phase -> packed (score | position) keys through a 6-deep compare-exchange chain;
        // main phase -> compare with the cut-off, a warp vote, the rare append out of line
        auto process = [&](const uint32_t (&cur)[32], int g) {
          if (sampling) {
            // non-negative floats order like integers; negatives clamp to 0; the low 12 bits carry the position
            const int pb = tsb | (g << 5);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int k = (max(static_cast<int>(cur[j]), 0) & ~0xFFF) | (pb | j);
              // sorted insert, every level from the OLD values (depth 2 instead of a 11-deep dependent chain):
              // new m_i = max(m_i, min(m_{i-1}, k))
              const int n6 = max(m6, min(m5, k)), n5 = max(m5, min(m4, k)), n4 = max(m4, min(m3, k));
              const int n3 = max(m3, min(m2, k)), n2 = max(m2, min(m1, k)), n1 = max(m1, k);
              m1 = n1; m2 = n2; m3 = n3; m4 = n4; m5 = n5; m6 = n6;
            }
          } else {
            // four scores per warp vote: the control flow around a vote costs more than the compares
            const int e0 = ent0 + g * 32;
#pragma unroll
            for (int j4 = 0; j4 < 32; j4 += 4) {
              const float s0 = __uint_as_float(cur[j4]), s1 = __uint_as_float(cur[j4 + 1]);
              const float s2 = __uint_as_float(cur[j4 + 2]), s3 = __uint_as_float(cur[j4 + 3]);
              if (__any_sync(0xffffffffu, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)) > tau)) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float sv = __uint_as_float(cur[j4 + u]);
                  const bool pass = sv > tau;
                  if (__any_sync(0xffffffffu, pass)) {
                    const unsigned long long r = tk_pass(lsc, lix, lane, pass && e0 + j4 + u < p.N, sv, e0 + j4 + u, cnt, tau);
                    cnt = static_cast<int>(r >> 32); tau = __uint_as_float(static_cast<unsigned>(r));
                  }
                }
              }
            }
          }
        };
        tmem_ld_32x32_ptr(taddr, va);
#pragma unroll 1
        for (int gp = 0; gp < 2; ++gp) {
          tmem_ld_wait();
          tmem_ld_32x32_ptr(taddr + (2 * gp + 1) * 32, vb);
          process(va, 2 * gp);
          tmem_ld_wait();
          if (gp == 0) {
            tmem_ld_32x32_ptr(taddr + 64, va);
          } else {  // last read of this accumulator stage
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(lead_tmem_empty + acc * 8);
          }
          process(vb, 2 * gp + 1);
        }
        if (t + 1 == ts_end) {
          // end of the sample: the 6 winners open the list, the 6th (truncated) score is the cut-off
          const int ms[6] = {m1, m2, m3, m4, m5, m6};
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const int pos = ms[i] & 0xFFF;
            const int ent = (t0 + (pos >> 7)) * 256 + half * 128 + (pos & 127);
            if ((ms[i] >> 12) != 0 && ent < p.N) { lsc[cnt * 32 + lane] = __int_as_float(ms[i] & ~0xFFF); lix[cnt * 32 + lane] = ent; ++cnt; }
          }
          tau = __int_as_float(m6 & ~0xFFF);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      // hand this (row, range, half) list to the rescoring kernel
      const long long row = static_cast<long long>(rb) * 256 + static_cast<int>(rank) * 128 + q * 32 + lane;
      const long long lbase = (row * (p.n_ranges * 2) + rg * 2 + half);
#pragma unroll 4
      for (int e = 0; e < kTkCand; ++e) {
        p.cand_score[lbase * kTkCand + e] = lsc[e * 32 + lane];
        p.cand_idx[lbase * kTkCand + e] = lix[e * 32 + lane];
      }
      p.cand_tau[lbase] = tau;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ 3. rescore
// One warp per query row (E = 128): exact scores of the candidates, the k best (descending; ties: lower index), and the
// proof that nothing outside the candidate lists can belong to them.  flag[row] = 1 asks for the exact scan.
// Each lane scores one candidate at a time (its whole 128-dim dot, the normalised query broadcast from shared memory):
// 32 candidates and 32 independent row reads in flight per warp, no shuffles.
__global__ void __launch_bounds__(256) topk_rescore_kernel(const float* __restrict__ q, const float* __restrict__ inv_q,
                                                           const float* __restrict__ table, const float* __restrict__ inv_t,
                                                           const float* __restrict__ cand_score, const int* __restrict__ cand_idx,
                                                           const float* __restrict__ cand_tau, int n_lists, long long B, int k,
                                                           long long* __restrict__ out_idx, float* __restrict__ out_score,
                                                           int* __restrict__ flag) {
  extern __shared__ uint8_t sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ncand = n_lists * kTkCand;
  float* qn = reinterpret_cast<float*>(sm) + w * 128;                                 // [8][128] normalised queries
  float* es = reinterpret_cast<float*>(sm) + 8 * 128 + static_cast<size_t>(w) * ncand * 2;   // exact scores
  int* ei = reinterpret_cast<int*>(es + ncand);
  const long long row = static_cast<long long>(blockIdx.x) * 8 + w;
  if (row >= B) return;
  {
    const float iq = inv_q[row];
    float4 v = *reinterpret_cast<const float4*>(q + row * 128 + 4 * lane);
    v.x *= iq; v.y *= iq; v.z *= iq; v.w *= iq;
    *reinterpret_cast<float4*>(qn + 4 * lane) = v;
  }
  const long long cb = row * n_lists * kTkCand;
  float tau_max = -3.0e38f;
  for (int l = lane; l < n_lists; l += 32) tau_max = fmaxf(tau_max, cand_tau[row * n_lists + l]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tau_max = fmaxf(tau_max, __shfl_xor_sync(0xffffffffu, tau_max, o));
  __syncwarp();
  for (int c = lane; c < ncand; c += 32) {
    const int idx = cand_idx[cb + c];
    float s = -3.0e38f;
    if (idx >= 0) {
      const float it = inv_t[idx];
      const float4* tr = reinterpret_cast<const float4*>(table + static_cast<long long>(idx) * 128);
      float d = 0.f;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float4 tv = __ldg(tr + j);
        const float4 qv = *reinterpret_cast<const float4*>(qn + 4 * j);
        d += qv.x * (tv.x * it) + qv.y * (tv.y * it) + qv.z * (tv.z * it) + qv.w * (tv.w * it);
      }
      s = d;
    }
    es[c] = s; ei[c] = idx;
  }
  __syncwarp();
  float kth = -3.0e38f;
  for (int r = 0; r < k; ++r) {
    float best = -3.0e38f; int bi = 0x7fffffff, bc = -1;
    for (int c = lane; c < ncand; c += 32) {
      const float s = es[c]; const int idx = ei[c];
      if (idx >= 0 && (s > best || (s == best && idx < bi))) { best = s; bi = idx; bc = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o), oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; bc = oc; }
    }
    if (lane == 0) {
      out_idx[row * k + r] = bc >= 0 ? bi : -1;
      out_score[row * k + r] = best;
      if (bc >= 0) ei[bc] = -1;   // taken
    }
    kth = best;
    __syncwarp();
  }
  // nothing outside the lists can reach the k-th exact score: every such entity's bf16 score is <= its list's cut-off
  if (lane == 0) flag[row] = (kth > tau_max + kTkErrBound) ? 0 : 1;
}

// Exact scan of the whole table for the rows whose flag is set (or all rows: always == 1): one CTA per row, every thread
// keeps the k best of its entities, the CTA merges.  Slow by design (fp32 SIMT over N x E) -- the proof above fails rarely.
__global__ void __launch_bounds__(256) topk_exact_kernel(const float* __restrict__ q, const float* __restrict__ inv_q,
                                                         const float* __restrict__ table, const float* __restrict__ inv_t, long long N,
                                                         int E, int k, const int* __restrict__ flag, int always,
                                                         long long* __restrict__ out_idx, float* __restrict__ out_score) {
  extern __shared__ uint8_t sm[];
  const long long row = blockIdx.x;
  if (!always && flag[row] == 0) return;
  float* qn = reinterpret_cast<float*>(sm);                 // [E]
  float* ms = qn + E;                                       // [256 * k] merged candidates
  int* mi = reinterpret_cast<int*>(ms + 256 * k);
  const float iq = inv_q[row];
  for (int c = threadIdx.x; c < E; c += blockDim.x) qn[c] = q[row * E + c] * iq;
  __syncthreads();
  float* mys = ms + threadIdx.x * k; int* myi = mi + threadIdx.x * k;
  for (int j = 0; j < k; ++j) { mys[j] = -3.0e38f; myi[j] = -1; }
  float tau = -3.0e38f;
  for (long long e = threadIdx.x; e < N; e += blockDim.x) {
    const float it = inv_t[e];
    const float* tr = table + e * E;
    float d = 0.f;
    for (int c = 0; c < E; c += 4) {
      const float4 tv = *reinterpret_cast<const float4*>(tr + c);
      d += qn[c] * (tv.x * it) + qn[c + 1] * (tv.y * it) + qn[c + 2] * (tv.z * it) + qn[c + 3] * (tv.w * it);
    }
    if (d > tau) {   // replace this thread's smallest
      int at = 0; float mn = mys[0];
      for (int j = 1; j < k; ++j) if (mys[j] < mn) { mn = mys[j]; at = j; }
      mys[at] = d; myi[at] = static_cast<int>(e);
      mn = mys[0];
      for (int j = 1; j < k; ++j) mn = fminf(mn, mys[j]);
      tau = mn;
    }
  }
  __syncthreads();
  // k rounds of a block-wide arg-max over the 256 * k kept candidates (warp 0 does the final step)
  __shared__ float rs[8]; __shared__ int ri[8], rc[8];
  for (int r = 0; r < k; ++r) {
    float best = -3.0e38f; int bi = 0x7fffffff, bc = -1;
    for (int c = threadIdx.x; c < 256 * k; c += blockDim.x) {
      const float s = ms[c]; const int idx = mi[c];
      if (idx >= 0 && (s > best || (s == best && idx < bi))) { best = s; bi = idx; bc = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o), oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; bc = oc; }
    }
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = best; ri[threadIdx.x >> 5] = bi; rc[threadIdx.x >> 5] = bc; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int j = 1; j < 8; ++j)
        if (rs[j] > best || (rs[j] == best && ri[j] < bi)) { best = rs[j]; bi = ri[j]; bc = rc[j]; }
      out_idx[row * k + r] = bc >= 0 ? bi : -1;
      out_score[row * k + r] = best;
      if (bc >= 0) mi[bc] = -1;
    }
    __syncthreads();
  }
}

}  // namespace pbg
