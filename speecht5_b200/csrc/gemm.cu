// Batched bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands staged by
// TMA into 128B-swizzled shared memory through an mbarrier ring. One CTA computes a 128 x BN output tile.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..5 =
// epilogue (TMEM -> registers -> fused bias / activation / dropout / residual -> global).
//
// This one kernel is the contraction engine for every dense op on the SpeechT5 path: the q/k/v/out projections
// (reference: speecht5/models/modules/multihead_attention.py:213-231,397), the FFN (transformer_layer.py:127-132,
// 385-391), pre-/post-net Linear layers, the post-net Conv1d stack expressed as an overlapping-window GEMM, and all
// of their backward contractions (MN-major operands avoid explicit transposes).
#include "gemm.cuh"
#include "ptx.cuh"
#include "kernels.cuh"
#include <cuda.h>
#include "tma_map.cuh"
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <unordered_map>

namespace st5 {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 16;  // epilogue warps: 4 per TMEM lane quarter, each taking every 4th 32-column chunk
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;

struct EpiParams {
  int M, N, nb1;
  void* C; int c_fp32; long c_ld, c_bs1, c_bs2;
  void* C_pre;
  const float* bias;
  const float* bias2; int bias2_rows;
  const void* residual;
  int act; float alpha; int accumulate;
  uint32_t drop_thr; float drop_scale; uint64_t drop_seed, drop_offset;
  int num_k_blocks;
  int a_m1, a_m2, b_m1, b_m2;  // 0 when the operand is broadcast over that batch dim (stride 0), else 1
  int c_m1, c_m2;              // the same for the output (accumulate == 2: split-K batches share one output)
  const void* ag_pre; int ag_act;   // optional: multiply by act'(ag_pre[m][n]) (activation backward fused into dX)
  int tma_store;                    // 1: outputs leave through shared memory + TMA store (maps map_c / map_cpre)
  int tiles_m, tiles_n, num_tiles;  // persistent schedule: tile = (z * tiles_n + n_blk) * tiles_m + m_blk
};

// Activation / activation-derivative over one 32-column chunk with the kind fixed at compile time: the runtime switch
// sits outside the unrolled element loop (one uniform branch per chunk instead of three per element).
template <int ACT>
__device__ __forceinline__ void act_chunk(float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if constexpr (ACT == ACT_RELU) v[j] = fmaxf(v[j], 0.f);
    if constexpr (ACT == ACT_GELU) v[j] = gelu_fwd(v[j]);
    if constexpr (ACT == ACT_TANH) v[j] = tanhf(v[j]);
    if constexpr (ACT == ACT_GELU_TANH) v[j] = gelu_tanh_fwd(v[j]);
  }
}
template <int ACT>
__device__ __forceinline__ void actgrad8(float* v, const uint4 u) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = __bfloat1622float2(h[t]);
    if constexpr (ACT == ACT_GATE) {  // the forward epilogue stored keep * scale * act'(pre): one multiply per output
      v[2 * t] *= f.x;
      v[2 * t + 1] *= f.y;
    } else {
      v[2 * t] *= act_grad(f.x, ACT);
      v[2 * t + 1] *= act_grad(f.y, ACT);
    }
  }
}

// CTAS = 2 (SURVEY section 8f / DESIGN section 8 item 1): the two CTAs of a cluster -- the two SMs of one TPC -- share a
// 256 x BN output tile. Each stages its own 128 rows of A and BN/2 columns of B (a third less operand traffic per
// output element than 128 x 256), the leader's MMA warp issues tcgen05.mma.cta_group::2 for both, and each CTA drains
// its own 128 accumulator rows. All pipeline barriers that cross the pair live in the leader's shared memory.
template <int BN, int STAGES, bool A_MN, bool B_MN, int CTAS>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap map_a,
                                                                  const __grid_constant__ CUtensorMap map_b,
                                                                  const __grid_constant__ CUtensorMap map_c,
                                                                  const __grid_constant__ CUtensorMap map_cpre,
                                                                  const EpiParams p) {
  constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  constexpr int B_ROWS = BN / CTAS;  // B columns (= rows of the [N, K] operand) this CTA stages
  constexpr uint32_t B_BYTES = B_ROWS * BLOCK_K * 2;
  constexpr uint32_t CHUNK_BYTES = 64 * BLOCK_K * 2;  // one 64(mn) x 64(k) MN-major box
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* stg_all = smem_b + STAGES * B_BYTES;  // EPI_WARPS x 4 KB staging blocks for the TMA-store epilogue
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_all + EPI_WARPS * 4096);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2] accumulator stage complete (MMA -> epilogue)
  uint64_t* tempty_bar = tfull_bar + 2;       // [2] accumulator stage drained (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  // Persistent CTA: loops over output tiles (m fastest, so concurrently running CTAs share the B/weight tile in L2);
  // two TMEM accumulator stages let the epilogue of tile i overlap the MMA main loop of tile i+1.
  const int warp = threadIdx.x >> 5;
  const int nkb = p.num_k_blocks;
  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int rank = CTAS == 2 ? (int)cluster_ctarank() : 0;           // 0 = leader of the pair
// first tile / stride of this CTA (pair) in the persistent schedule -- kept as expressions, not locals, so that the
// single-CTA instantiations compile to the code they had before the pair variant existed
#define ST5_TILE0 (CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x)
#define ST5_TILE_STEP (CTAS == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x)
  constexpr int TILE_M = BLOCK_M * CTAS;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS * CTAS);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    fence_mbar_init();
  }
  if constexpr (CTAS == 2) cluster_sync_all();  // both CTAs of the pair are resident before the paired TMEM allocation
  if (warp == 1) {
    if constexpr (CTAS == 2) {
      tmem_alloc_pair(tmem_slot, 2 * BN);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, 2 * BN);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();  // the peer's barriers exist before anything is signalled across
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: everything above touched only this CTA's shared / tensor memory, so it may run
  // while the preceding grid drains. Wait for that grid's memory here, before the first global access, and let the
  // grid behind us start its own prologue (both are no-ops when the launch carries no programmatic attribute).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = ST5_TILE0; tile < p.num_tiles; tile += ST5_TILE_STEP) {
      const int z = tile / tiles_mn, rmn = tile - z * tiles_mn;
      const int m0 = (rmn % p.tiles_m) * TILE_M + rank * BLOCK_M, n0 = (rmn / p.tiles_m) * BN + rank * B_ROWS;
      const int b1 = z % p.nb1, b2 = z / p.nb1;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem_a + stage * A_BYTES;
        uint8_t* sb = smem_b + stage * B_BYTES;
        const int k0 = kb * BLOCK_K;
        if constexpr (CTAS == 2) {
          // both producers credit the LEADER's full barrier, which expects the bytes of both halves of the stage
          const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * (A_BYTES + B_BYTES));
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_4d_pair(sa + c * CHUNK_BYTES, &map_a, fb, m0 + c * 64, k0, b1 * p.a_m1, b2 * p.a_m2);
          } else {
            tma_load_4d_pair(sa, &map_a, fb, k0, m0, b1 * p.a_m1, b2 * p.a_m2);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < B_ROWS / 64; ++c)
              tma_load_4d_pair(sb + c * CHUNK_BYTES, &map_b, fb, n0 + c * 64, k0, b1 * p.b_m1, b2 * p.b_m2);
          } else {
            tma_load_4d_pair(sb, &map_b, fb, k0, n0, b1 * p.b_m1, b2 * p.b_m2);
          }
        } else {
        mbar_expect_tx(&full_bar[stage], A_BYTES + B_BYTES);
        if (A_MN) {
#pragma unroll
          for (int c = 0; c < BLOCK_M / 64; ++c)
            tma_load_4d(sa + c * CHUNK_BYTES, &map_a, &full_bar[stage], m0 + c * 64, k0, b1 * p.a_m1, b2 * p.a_m2);
        } else {
          tma_load_4d(sa, &map_a, &full_bar[stage], k0, m0, b1 * p.a_m1, b2 * p.a_m2);
        }
        if (B_MN) {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            tma_load_4d(sb + c * CHUNK_BYTES, &map_b, &full_bar[stage], n0 + c * 64, k0, b1 * p.b_m1, b2 * p.b_m2);
        } else {
          tma_load_4d(sb, &map_b, &full_bar[stage], k0, n0, b1 * p.b_m1, b2 * p.b_m2);
        }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      }
      if constexpr (CTAS == 2) {
        // tail: the leader's multicast commits arrive on THIS CTA's empty barriers; stay until every slot has been
        // released so that none of them lands in the shared memory of a CTA that already left
#pragma unroll 1
        for (int i = 0; i < STAGES; ++i) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    // (pair: only the leader's MMA warp issues; the peer's warp 1 just owns its half of the TMEM allocation)
    for (int tile = (CTAS == 2 && rank != 0) ? p.num_tiles : ST5_TILE0; tile < p.num_tiles; tile += ST5_TILE_STEP, ++local) {
    const int acc = local & 1;
    const uint32_t acc_phase = (uint32_t)(local >> 1) & 1u;
    const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BN);
    mbar_wait(&tempty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator stage
    tc_fence_after();
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem_a + stage * A_BYTES);
        const uint32_t sb = smem_u32(smem_b + stage * B_BYTES);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          // K-major: advance 16 elements (32 B) along the swizzled row. MN-major: advance 16 k-rows (2 KB).
          const uint64_t da = A_MN ? umma_smem_desc(sa + k * (UMMA_K * 128), CHUNK_BYTES, 1024)
                                   : umma_smem_desc(sa + k * (UMMA_K * 2), 16, 1024);
          const uint64_t db = B_MN ? umma_smem_desc(sb + k * (UMMA_K * 128), CHUNK_BYTES, 1024)
                                   : umma_smem_desc(sb + k * (UMMA_K * 2), 16, 1024);
          if constexpr (CTAS == 2) umma_bf16_pair(tmem_acc, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          else umma_bf16(tmem_acc, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        if constexpr (CTAS == 2) {  // the same barrier offsets in both CTAs of the pair
          umma_commit_pair(&empty_bar[stage], 3);
          if (kb == nkb - 1) umma_commit_pair(&tfull_bar[acc], 3);
        } else {
        umma_commit(&empty_bar[stage]);                     // smem slot reusable once these MMAs retire
        if (kb == nkb - 1) umma_commit(&tfull_bar[acc]);    // accumulator complete
        }
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    }
  } else {
    // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
    const int q = warp & 3;                      // TMEM lane quarter this warp may read (hardware: warp id % 4)
    constexpr int NGRP = EPI_WARPS / 4;
    const int grp = (warp - 2) >> 2;             // which 32-column chunks this warp owns: c = grp, grp + NGRP, ...
    uint8_t* stg = stg_all + (warp - 2) * 4096;
    uint64_t dseed = p.drop_seed, doffset = p.drop_offset;
    if (p.drop_thr != 0) resolve_seed(dseed, doffset);
    int local = 0;
    for (int tile = ST5_TILE0; tile < p.num_tiles; tile += ST5_TILE_STEP, ++local) {
    const int z = tile / tiles_mn, rmn = tile - z * tiles_mn;
    const int m0 = (rmn % p.tiles_m) * TILE_M + rank * BLOCK_M, n0 = (rmn / p.tiles_m) * BN;
    const int b1 = z % p.nb1, b2 = z / p.nb1;
    const int acc = local & 1;
    const uint32_t acc_phase = (uint32_t)(local >> 1) & 1u;
    const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BN);
    const int row = m0 + q * 32 + (int)lane_id();
    const bool row_ok = row < p.M;
    const long zoff = (long)b1 * p.c_bs1 + (long)b2 * p.c_bs2;
    const long roff = zoff + (long)row * p.c_ld;
    // (Requesting the gate / residual / bias of a chunk ahead of the accumulator wait was tried -- A/B build, round 2:
    // the 16-24 extra live registers cost more in spills than the hidden latency gave back, +0.3 ms per step.)
    mbar_wait(&tfull_bar[acc], acc_phase);
    tc_fence_after();
    const float* bias2_row = (p.bias2 != nullptr && row_ok) ? p.bias2 + (long)(row / p.bias2_rows) * p.N : nullptr;
    const uint64_t drop_row = ((uint64_t)z * (uint64_t)p.M + (uint64_t)row) * (uint64_t)p.N;
    if (grp >= BN / 32) {  // narrow tiles: this warp has no chunk, it only releases the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) {
        if constexpr (CTAS == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
        else mbar_arrive(&tempty_bar[acc]);
      }
    }
#pragma unroll 1
    for (int c = grp; c < BN / 32; c += NGRP) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      if (c + NGRP >= BN / 32) {  // last TMEM read of this warp for this tile: hand the stage back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane_id() == 0) {
          if constexpr (CTAS == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), 0));
          else mbar_arrive(&tempty_bar[acc]);
        }
      }
      const int nb = n0 + c * 32;
      if (nb >= p.N) continue;  // warp-uniform; rows beyond M keep going (their loads are guarded, stores clipped)
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      if (p.alpha != 1.f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
      }
      const bool full = (nb + 32 <= p.N);
      if (p.accumulate == 1 && row_ok) {  // partial sums of a multi-pass (split-precision) product live in C (fp32)
        const float* src = reinterpret_cast<const float*>(p.C) + roff + nb;
        if (full && ((p.c_ld & 3) == 0) && ((zoff & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 o = *reinterpret_cast<const float4*>(src + j);
            v[j] += o.x; v[j + 1] += o.y; v[j + 2] += o.z; v[j + 3] += o.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || nb + j < p.N) v[j] += src[j];
        }
      }
      if (p.bias != nullptr) {
        if (full && ((reinterpret_cast<uintptr_t>(p.bias + nb) & 15) == 0)) {
          // every lane needs the same 32 values: eight 16-byte loads of one address per warp (L1 broadcast)
          const float4* bp = reinterpret_cast<const float4*>(p.bias + nb);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(bp + j);
            v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
        } else {  // ragged edge: one coalesced load per warp, then register shuffles
          const int jn = nb + (int)lane_id();
          const float bl = jn < p.N ? __ldg(p.bias + jn) : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bl, j);
        }
      }
      if (bias2_row != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (full || nb + j < p.N) v[j] += __ldg(bias2_row + nb + j);
      }
      const bool vec_ok = full && ((p.c_ld & 7) == 0) && ((zoff & 7) == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                          ((reinterpret_cast<uintptr_t>(p.C_pre) & 15) == 0);
      // ---- output path: stage the 32x32 block in shared memory and let TMA write it (coalesced, asynchronous,
      // clipped at the tensor edges); fall back to per-thread stores when the output layout is not TMA-addressable.
      auto emit = [&](void* base, const CUtensorMap* tmap) {
        if (p.tma_store) {
          if (lane_id() == 0) bulk_wait_read0();  // previous block of this warp has left the staging buffer
          __syncwarp();
          const int rr = (int)lane_id();
          if (p.c_fp32) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<float4*>(stg + rr * 128 + ((g ^ (rr & 7)) << 4)) =
                  make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
          } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 pk;
              __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * g], v[8 * g + 1]);
              __nv_bfloat162 t1 = __floats2bfloat162_rn(v[8 * g + 2], v[8 * g + 3]);
              __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * g + 4], v[8 * g + 5]);
              __nv_bfloat162 t3 = __floats2bfloat162_rn(v[8 * g + 6], v[8 * g + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
              pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
              *reinterpret_cast<uint4*>(stg + rr * 64 + ((g ^ ((rr >> 1) & 3)) << 4)) = pk;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane_id() == 0) {
            if (p.accumulate == 2) tma_reduce_add_4d(tmap, stg, nb, m0 + q * 32, b1 * p.c_m1, b2 * p.c_m2);  // C += tile, at the L2
            else tma_store_4d(tmap, stg, nb, m0 + q * 32, b1 * p.c_m1, b2 * p.c_m2);
            bulk_commit();
          }
          return;
        }
        if (!row_ok) return;
        if (p.c_fp32) {
          float* dst = reinterpret_cast<float*>(base) + roff + nb;
          if (full && ((p.c_ld & 3) == 0) && ((zoff & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || nb + j < p.N) dst[j] = v[j];
          }
        } else {
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(base) + roff + nb;
          if (vec_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]);
              __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
              __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
              __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
              pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
              *reinterpret_cast<uint4*>(dst + j) = pk;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || nb + j < p.N) dst[j] = __float2bfloat16(v[j]);
          }
        }
      };
      if (p.act == ACT_GELU_TANH_GATE) {
        // fc1 of the FFN in throughput mode: out = drop(gelu_tanh(x)), and INSTEAD of the pre-activation the second
        // output is the multiplier the backward pass needs, gate = keep * scale * gelu_tanh'(x) (tanh(u) is shared by
        // both). The dH = dO.W2 GEMM of the backward then multiplies by it -- no Philox, no tanh in that epilogue.
        // Eight columns at a time straight into the two halves of this warp's staging block (bf16: 2 KB each), so
        // the chunk is never held twice in registers. Launcher guarantees: bf16 output, TMA-store layout, N % 8 == 0.
        if (lane_id() == 0) bulk_wait_read0();
        __syncwarp();
        const int rr = (int)lane_id();
        const uint64_t e0 = drop_row + (uint64_t)nb;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float o[8], d[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float x = v[8 * g + t];
            const float x2 = x * x;
            const float th = fast_tanh(x * fmaf(0.0356774081f, x2, 0.7978845608f));
            const float hx = 0.5f * x;
            o[t] = fmaf(hx, th, hx);
            d[t] = fmaf(hx * fmaf(-th, th, 1.f), fmaf(0.1070322243f, x2, 0.7978845608f), fmaf(0.5f, th, 0.5f));
          }
          if (p.drop_thr != 0) {
            const Philox4 r4 = philox4x32(dseed, doffset, (e0 + 8 * g) >> 3);
            const uint32_t w[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float k0 = (w[t] & 0xFFFFu) >= p.drop_thr ? p.drop_scale : 0.f;
              const float k1 = (w[t] >> 16) >= p.drop_thr ? p.drop_scale : 0.f;
              o[2 * t] *= k0; d[2 * t] *= k0;
              o[2 * t + 1] *= k1; d[2 * t + 1] *= k1;
            }
          }
          uint4 po, pd;
          {
            __nv_bfloat162 t0 = __floats2bfloat162_rn(o[0], o[1]), t1 = __floats2bfloat162_rn(o[2], o[3]);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(o[4], o[5]), t3 = __floats2bfloat162_rn(o[6], o[7]);
            po.x = *reinterpret_cast<uint32_t*>(&t0); po.y = *reinterpret_cast<uint32_t*>(&t1);
            po.z = *reinterpret_cast<uint32_t*>(&t2); po.w = *reinterpret_cast<uint32_t*>(&t3);
            __nv_bfloat162 u0 = __floats2bfloat162_rn(d[0], d[1]), u1 = __floats2bfloat162_rn(d[2], d[3]);
            __nv_bfloat162 u2 = __floats2bfloat162_rn(d[4], d[5]), u3 = __floats2bfloat162_rn(d[6], d[7]);
            pd.x = *reinterpret_cast<uint32_t*>(&u0); pd.y = *reinterpret_cast<uint32_t*>(&u1);
            pd.z = *reinterpret_cast<uint32_t*>(&u2); pd.w = *reinterpret_cast<uint32_t*>(&u3);
          }
          const int so = rr * 64 + ((g ^ ((rr >> 1) & 3)) << 4);
          *reinterpret_cast<uint4*>(stg + so) = po;
          *reinterpret_cast<uint4*>(stg + 2048 + so) = pd;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane_id() == 0) {
          tma_store_4d(&map_c, stg, nb, m0 + q * 32, b1, b2);
          tma_store_4d(&map_cpre, stg + 2048, nb, m0 + q * 32, b1, b2);
          bulk_commit();
        }
        continue;
      }
      if (p.C_pre != nullptr) emit(p.C_pre, &map_cpre);
      if (p.act == ACT_GELU_TANH) act_chunk<ACT_GELU_TANH>(v);
      else if (p.act == ACT_GELU) act_chunk<ACT_GELU>(v);
      else if (p.act == ACT_RELU) act_chunk<ACT_RELU>(v);
      else if (p.act == ACT_TANH) act_chunk<ACT_TANH>(v);
      if (p.drop_thr != 0) {
        const uint64_t e0 = drop_row + (uint64_t)nb;
        if ((e0 & 7) == 0) {  // aligned: one Philox call per 8 elements
#pragma unroll
          for (int j = 0; j < 32; j += 8) dropout8_apply(v + j, e0 + j, p.drop_thr, p.drop_scale, dseed, doffset);
        } else {  // odd row pitch: assemble the 32 keep bits from at most five calls
          const uint32_t keep = dropout_keep_mask32(dseed, doffset, e0, p.drop_thr);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ((keep >> j) & 1u) ? v[j] * p.drop_scale : 0.f;
        }
      }
      if (p.ag_pre != nullptr && row_ok) {
        if (p.c_fp32) {
          const float* pr = reinterpret_cast<const float*>(p.ag_pre) + roff + nb;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || nb + j < p.N) v[j] *= p.ag_act == ACT_GATE ? pr[j] : act_grad(pr[j], p.ag_act);
        } else {
          const __nv_bfloat16* pr = reinterpret_cast<const __nv_bfloat16*>(p.ag_pre) + roff + nb;
          if (vec_ok && ((reinterpret_cast<uintptr_t>(p.ag_pre) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 u = *reinterpret_cast<const uint4*>(pr + j);
              if (p.ag_act == ACT_GATE) actgrad8<ACT_GATE>(v + j, u);
              else if (p.ag_act == ACT_GELU_TANH) actgrad8<ACT_GELU_TANH>(v + j, u);
              else if (p.ag_act == ACT_GELU) actgrad8<ACT_GELU>(v + j, u);
              else if (p.ag_act == ACT_RELU) actgrad8<ACT_RELU>(v + j, u);
              else actgrad8<ACT_TANH>(v + j, u);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || nb + j < p.N)
                v[j] *= p.ag_act == ACT_GATE ? __bfloat162float(pr[j]) : act_grad(__bfloat162float(pr[j]), p.ag_act);
          }
        }
      }
      if (p.residual != nullptr && row_ok) {
        if (p.c_fp32) {
          const float* rs = reinterpret_cast<const float*>(p.residual) + roff + nb;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || nb + j < p.N) v[j] += rs[j];
        } else {
          const __nv_bfloat16* rs = reinterpret_cast<const __nv_bfloat16*>(p.residual) + roff + nb;
          if (vec_ok && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 u = *reinterpret_cast<const uint4*>(rs + j);
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float2 f = __bfloat1622float2(h[t]);
                v[j + 2 * t] += f.x;
                v[j + 2 * t + 1] += f.y;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || nb + j < p.N) v[j] += __bfloat162float(rs[j]);
          }
        }
      }
      emit(p.C, &map_c);
    }
    }
    if (p.tma_store && lane_id() == 0) bulk_wait_read0();  // staging buffer must outlive the last bulk store
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();  // neither CTA may leave while the other's MMAs / TMA still target it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTAS == 2) tmem_dealloc_pair(tmem_base, 2 * BN);
    else tmem_dealloc(tmem_base, 2 * BN);
  }
}

#undef ST5_TILE0
#undef ST5_TILE_STEP

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

int device_sm_count() { return num_sms(); }

int encode_map_4d(CUtensorMap* map, const void* ptr, int is_f32, const uint64_t dims[4],
                  const uint64_t strides_bytes[3], const uint32_t box[4], int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -10;
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (s[0] & 15) || (s[1] & 15) || (s[2] & 15)) return -11;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                   const_cast<void*>(ptr), d, s, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -12;
}

int encode_bf16_map_4d(CUtensorMap* map, const void* ptr, const uint64_t dims[4], const uint64_t strides_bytes[3],
                       const uint32_t box[4]) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -10;
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (s[0] & 15) || (s[1] & 15) || (s[2] & 15)) return -11;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), d, s, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -12;
}

// rows x K operand. K-major: memory [rows][ld] (k contiguous). MN-major: memory [K][ld] (row index contiguous).
static int make_operand_map(CUtensorMap* map, const void* ptr, int mn_major, int rows, int K, long ld, int nb1,
                            long bs1, int nb2, long bs2, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -10;
  cuuint64_t dims[4];
  cuuint64_t strides[3];
  cuuint32_t box[4];
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (mn_major) {
    dims[0] = (cuuint64_t)rows; dims[1] = (cuuint64_t)K;
    box[0] = 64; box[1] = BLOCK_K;
  } else {
    dims[0] = (cuuint64_t)K; dims[1] = (cuuint64_t)rows;
    box[0] = BLOCK_K; box[1] = (cuuint32_t)box_rows;
  }
  if (nb1 > 1 && bs1 == 0) nb1 = 1;  // broadcast operand: the kernel pins that coordinate to 0
  if (nb2 > 1 && bs2 == 0) nb2 = 1;
  dims[2] = (cuuint64_t)nb1; dims[3] = (cuuint64_t)nb2;
  box[2] = 1; box[3] = 1;
  strides[0] = (cuuint64_t)ld * 2;
  strides[1] = (cuuint64_t)(nb1 > 1 ? bs1 : ld) * 2;
  strides[2] = (cuuint64_t)(nb2 > 1 ? bs2 : ld) * 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15))
    return -11;  // TMA needs 16-byte aligned base and strides
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -12;
}

template <int BN, int STAGES, bool A_MN, bool B_MN, int CTAS = 1>
static int launch_variant(const GemmDesc& g, const EpiParams& ep, cudaStream_t stream) {
  constexpr int TILE_M = BLOCK_M * CTAS;  // CTAS = 2: one CTA pair per 256 x BN tile, B columns split between the CTAs
  CUtensorMap ma, mb;
  int rc = make_operand_map(&ma, g.A, g.a_mn, g.M, g.K, g.a_ld, g.nb1, g.a_bs1, g.nb2, g.a_bs2, BLOCK_M);
  if (rc) return rc;
  rc = make_operand_map(&mb, g.B, g.b_mn, g.N, g.K, g.b_ld, g.nb1, g.b_bs1, g.nb2, g.b_bs2, BN / CTAS);
  if (rc) return rc - 10;
  constexpr size_t smem = (size_t)STAGES * (BLOCK_M * BLOCK_K * 2 + (BN / CTAS) * BLOCK_K * 2) + EPI_WARPS * 4096 +
                          (2 * STAGES + 4) * 8 + 16 + 1024;
  auto kern = gemm_bf16_tcgen05<BN, STAGES, A_MN, B_MN, CTAS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  EpiParams e2 = ep;
  e2.tiles_m = (g.M + TILE_M - 1) / TILE_M;
  e2.tiles_n = (g.N + BN - 1) / BN;
  const long total = (long)e2.tiles_m * e2.tiles_n * g.nb1 * g.nb2;
  if (total > 0x7fffffffL) return -4;
  e2.num_tiles = (int)total;
  // Output maps for the TMA-store epilogue: 32x32 element boxes, 64B (bf16) / 128B (fp32) swizzle. Needs 16-byte
  // aligned bases and row / batch pitches; otherwise the epilogue falls back to per-thread stores.
  CUtensorMap mc, mcp;
  memset(&mc, 0, sizeof(mc));
  memset(&mcp, 0, sizeof(mcp));
  e2.tma_store = 0;
  {
    const long es = g.c_fp32 ? 4 : 2;
    const bool ok = ((g.c_ld * es) % 16 == 0) && (g.nb1 <= 1 || (g.c_bs1 * es) % 16 == 0) &&
                    (g.nb2 <= 1 || (g.c_bs2 * es) % 16 == 0) && (reinterpret_cast<uintptr_t>(g.C) % 16 == 0) &&
                    (g.C_pre == nullptr || reinterpret_cast<uintptr_t>(g.C_pre) % 16 == 0);
    if (ok) {
      const int cn1 = (g.nb1 > 1 && g.c_bs1 != 0) ? g.nb1 : 1, cn2 = (g.nb2 > 1 && g.c_bs2 != 0) ? g.nb2 : 1;
      const uint64_t dims[4] = {(uint64_t)g.N, (uint64_t)g.M, (uint64_t)cn1, (uint64_t)cn2};
      const uint64_t strides[3] = {(uint64_t)(g.c_ld * es), (uint64_t)((cn1 > 1 ? g.c_bs1 : g.c_ld) * es),
                                   (uint64_t)((cn2 > 1 ? g.c_bs2 : g.c_ld) * es)};
      const uint32_t box[4] = {32, 32, 1, 1};
      int r2 = encode_map_4d(&mc, g.C, g.c_fp32, dims, strides, box, g.c_fp32 ? 128 : 64);
      if (!r2 && g.C_pre != nullptr) r2 = encode_map_4d(&mcp, g.C_pre, g.c_fp32, dims, strides, box, g.c_fp32 ? 128 : 64);
      e2.tma_store = r2 == 0 ? 1 : 0;
    }
  }
  if (g.accumulate == 2 && !e2.tma_store) return -6;  // the L2-side accumulate is a TMA reduce: needs that output layout
  if (e2.act == ACT_GELU_TANH_GATE &&
      (!e2.tma_store || g.c_fp32 || g.C_pre == nullptr || (g.N & 7) != 0 || g.residual != nullptr || g.ag_pre != nullptr))
    return -5;  // the gate epilogue needs the TMA-store layout, bf16 outputs and an 8-aligned row length
  const long slots = num_sms() / CTAS;
  const int grid = (int)(total < slots ? total : slots) * CTAS;  // one persistent CTA (or CTA pair per TPC) per SM
  static const bool pdl = [] {
    const char* e = getenv("ST5_PDL");
    return e == nullptr || atoi(e) != 0;
  }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CTAS == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, mcp, e2);
  if (le != cudaSuccess) return (int)le;
  return (int)cudaGetLastError();
}

template <int BN, int STAGES, int CTAS = 1>
static int launch_major(const GemmDesc& g, const EpiParams& ep, cudaStream_t stream) {
  if (g.a_mn) {
    if (g.b_mn) return launch_variant<BN, STAGES, true, true, CTAS>(g, ep, stream);
    return launch_variant<BN, STAGES, true, false, CTAS>(g, ep, stream);
  }
  if (g.b_mn) return launch_variant<BN, STAGES, false, true, CTAS>(g, ep, stream);
  return launch_variant<BN, STAGES, false, false, CTAS>(g, ep, stream);
}

int gemm_launch(const GemmDesc& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.nb1 <= 0 || g.nb2 <= 0) return 0;
  if (g.K <= 0) return -2;
  if (g.accumulate && !g.c_fp32) return -3;
  // a batch dimension with output stride 0 (split-K: several partial products into ONE output) is only sound with the
  // L2-side accumulate
  if (g.accumulate != 2 && ((g.nb1 > 1 && g.c_bs1 == 0) || (g.nb2 > 1 && g.c_bs2 == 0))) return -7;
  EpiParams ep;
  ep.M = g.M; ep.N = g.N; ep.nb1 = g.nb1;
  ep.C = g.C; ep.c_fp32 = g.c_fp32; ep.c_ld = g.c_ld; ep.c_bs1 = g.c_bs1; ep.c_bs2 = g.c_bs2;
  ep.C_pre = g.C_pre; ep.bias = g.bias; ep.bias2 = g.bias2; ep.bias2_rows = g.bias2_rows > 0 ? g.bias2_rows : 1;
  ep.ag_pre = g.ag_pre; ep.ag_act = g.ag_act;
  ep.residual = g.residual; ep.act = g.act; ep.alpha = g.alpha; ep.accumulate = g.accumulate;
  ep.drop_thr = drop_threshold(g.drop_p);
  ep.drop_scale = g.drop_p > 0.f ? 1.f / (1.f - g.drop_p) : 1.f;
  ep.drop_seed = g.drop_seed; ep.drop_offset = g.drop_offset;
  ep.num_k_blocks = (g.K + BLOCK_K - 1) / BLOCK_K;
  ep.a_m1 = (g.nb1 > 1 && g.a_bs1 == 0) ? 0 : 1; ep.a_m2 = (g.nb2 > 1 && g.a_bs2 == 0) ? 0 : 1;
  ep.b_m1 = (g.nb1 > 1 && g.b_bs1 == 0) ? 0 : 1; ep.b_m2 = (g.nb2 > 1 && g.b_bs2 == 0) ? 0 : 1;
  ep.c_m1 = (g.nb1 > 1 && g.c_bs1 == 0) ? 0 : 1; ep.c_m2 = (g.nb2 > 1 && g.c_bs2 == 0) ? 0 : 1;
  // Tile width: minimise (rounds of the persistent grid) x (time per tile). A 128x256 tile reads 12 KB of smem per
  // 128-cycle MMA (under the 128 B/clk port); 128x128 and 128x64 tiles are smem-port bound, hence the >1 factors.
  const long tiles_m = (g.M + BLOCK_M - 1) / BLOCK_M;
  const long batch = (long)g.nb1 * g.nb2;
  const long sms = num_sms();
  auto cost = [&](int bn, double factor) {
    const long tiles = tiles_m * ((g.N + bn - 1) / bn) * batch;
    const long rounds = (tiles + sms - 1) / sms;
    return (double)rounds * (bn * factor + 24.0);
  };
  static const int force_bn = [] {  // tuning / profiling knob: ST5_GEMM_BN=64|128|256 pins the tile width
    const char* e = getenv("ST5_GEMM_BN");
    return e ? atoi(e) : 0;
  }();
  // ST5_GEMM_PAIR=1|2 (experimental, off by default until measured on the GPU; 2 = 5-stage ring): 256 x 256 tiles on CTA pairs for every
  // problem the 256-wide tile would have been chosen for and that has at least one full pair of row blocks.
  static const int pair_mode = [] {
    const char* e = getenv("ST5_GEMM_PAIR");
    return e ? atoi(e) : -1;  // -1 = automatic (measured rule below), 0 = never, 1 / 2 = wherever the 256-wide tile applies
  }();
  if (force_bn == 256) return launch_major<256, 3>(g, ep, stream);
  if (force_bn == 128) return launch_major<128, 5>(g, ep, stream);
  if (force_bn == 64) return launch_major<64, 6>(g, ep, stream);
  const double c256 = g.N > 128 ? cost(256, 1.0) : 1e30;
  const double c128 = g.N > 64 ? cost(128, 1.12) : 1e30;
  const double c64 = cost(64, 1.35);
  // Weight gradients (accumulate == 2: both operands MN-major, long contraction split over the batch dimension by the
  // caller so that the CTA pairs fill the machine): 256 x 256 pair tiles halve the operand bytes per FLOP, which is what
  // bounds these GEMMs (128 x 128 tiles: 36 % tensor-pipe activity, the library reaches 1.45 PFLOP/s on the same shapes)
  if (pair_mode != 0 && force_bn == 0 && g.accumulate == 2 && g.M >= 2 * BLOCK_M && g.N >= 192)
    return launch_major<256, 4, 2>(g, ep, stream);
  if (c256 <= c128 && c256 <= c64) {
    if (pair_mode == 2 && g.M > BLOCK_M) return launch_major<256, 5, 2>(g, ep, stream);  // deeper ring (tuning)
    if (pair_mode == 1 && g.M > BLOCK_M) return launch_major<256, 4, 2>(g, ep, stream);
    // Automatic: the CTA pair wins where operand feed is the limit and there are several rounds of tiles (measured on a
    // B200, profiles/r02_pair_gemm_check.log: 8192^3 x1.15, M=10016 N=3072 K=768 x1.18, K=3072 x1.04), loses on
    // single-round grids and on the two-output GELU epilogue (x0.89-0.97).
    if (pair_mode == -1 && g.M >= 4096 && !g.a_mn && batch == 1 && g.C_pre == nullptr &&
        ((long)g.K >= 3072 || ((long)g.N >= 3072 && (long)g.M >= 8192) || (g.ag_pre != nullptr && (long)g.N >= 3072)))
      return launch_major<256, 4, 2>(g, ep, stream);
    return launch_major<256, 3>(g, ep, stream);
  }
  if (c128 <= c64) return launch_major<128, 5>(g, ep, stream);
  return launch_major<64, 6>(g, ep, stream);
}

}  // namespace st5
