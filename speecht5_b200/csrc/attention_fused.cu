// Fused attention forward on tcgen05: one CTA = (128 query rows, head, utterance); QK^T and PV run on the tensor
// cores with the whole score row block resident in TMEM (no HBM round trip for scores / probabilities):
//
//   TMA: Q tile [128x64], K [Tk x 64], V [Tk x 64] (read in place from the fused projection buffers, 128B swizzle)
//   MMA: S[128 x Tk] = Q K^T                                   -> TMEM columns [0, Tk)
//   softmax warps (1 thread = 1 query row = 1 TMEM lane): scale, causal / key-padding mask, exact two-pass softmax
//        straight out of TMEM, dropout, P -> shared memory as the K-major A operand (bf16, 128B swizzle),
//        optional probability output (fp32 for need_head_weights, multihead_attention.py:399-405) and log-sum-exp
//   MMA: O[128 x 64] = dropout(P) V   (V consumed as an MN-major B operand: no transpose)  -> TMEM columns [448, 512)
//   epilogue: O / rowsum -> bf16 -> out[b, i, h*64 : h*64+64]
//
// Reference semantics: speecht5/models/modules/multihead_attention.py:340-389. Tk <= 320 (S must fit TMEM next to O).
//
// RPE variant (encoder self-attention, encoder.py:239-246 + multihead_attention.py:356-364; Tq, Tk <= maxpos <= 160 so
// that clamp(i-j) never clips): the bias q_i . pe[i-j+maxpos] is a *skewed* read of QP = Q PE^T. The kernel computes
//   MMA: QP[128 x 288] = Q PE'^T, PE' = the 288 table rows this query tile can reach -> TMEM columns [160, 448)
// and every softmax thread pulls its row's 64-column window through a private shared-memory row (conflict-free odd
// pitch), reads it back reversed and shifted by its lane, adds it to S and parks the sum in TMEM (tcgen05.st) for the
// second softmax pass. Nothing of size T x T or T x 2*maxpos goes to HBM except the bf16 probabilities the backward
// pass consumes.
#include "../../include/speecht5_b200.h"
#include "kernels.cuh"
#include "ptx.cuh"
#include "tma_map.cuh"

namespace st5 {

int set_error(int code, const char* where);

// TMA warp, MMA warp, then NG groups of 4 softmax warps (one per TMEM lane quarter); group g owns the 32-column chunks
// c with c % NG == g. 16 softmax warps (4 per scheduler) hide the MUFU / TMEM-load latencies of the row passes; the
// RPE variant keeps 8 (its per-warp window staging would not fit shared memory otherwise).
template <bool RPE> constexpr int fa_groups() { return RPE ? 2 : 4; }
template <bool RPE> constexpr int fa_threads() { return 64 + fa_groups<RPE>() * 128; }
constexpr int FA_BM = 128;
constexpr int FA_MAX_TK = 320;
constexpr int FA_KBOX = 160;        // K rows per TMA box (two boxes cover 320 keys)
constexpr uint32_t FA_O_COL = 448;  // TMEM column of the O accumulator (S occupies [0, 320))
constexpr size_t FA_SMEM = 16384 + 40960 + 40960 + 81920 + 64 + 4096 + 1024;
// RPE variant: Tk <= 160. Q 16K | K 20K | V 3x8K | PE' 288 rows x 128 B | P 3x16K | 8 warps x (32 rows x 68 floats) staging
constexpr int FR_MAX_T = 160;
constexpr int FR_PE_ROWS = 288;
constexpr uint32_t FR_QP_COL = 160;
constexpr int FR_STAGE_PITCH = 68;  // floats per staged row: 16-byte aligned, 4*lane mod 32 banks -> conflict-free
constexpr size_t FR_STAGE_BYTES = 32 * FR_STAGE_PITCH * 4;
constexpr size_t FR_SMEM = 16384 + 20480 + 24576 + FR_PE_ROWS * 128 + 49152 + 8 * FR_STAGE_BYTES + 64 + 4096 + 1024;

struct FusedFwdParams {
  int B, H, Tq, Tk, causal;
  float scale_log2;        // scale * log2(e)
  const uint8_t* key_pad;  // [B][Tk] or null
  __nv_bfloat16* out; long o_ld, o_bs;
  float* lse;              // [B][H][Tq] natural-log sum-exp of the scaled scores
  __nv_bfloat16* psave;    // [B][H][Tq][p_ld] for the backward pass: exp(s - rowmax) (NOT normalised), sign bit = dropped
  float* inv_l;            // [B][H][Tq] 1 / sum_j exp(s - rowmax): the backward's normaliser of psave
  float* out_f32;          // optional [B][Tq][H*64]: the un-rounded output, for the backward's row constant dO.O
  void* probs; int probs_fp32; long p_ld;  // optional undropped probabilities [B][H][Tq][p_ld]
  int probs_heads;         // > 0: only heads < probs_heads get their probabilities written
  uint32_t drop_thr; float drop_scale; uint64_t seed, offset;
  int pe_row0;             // RPE: table row held by PE' row 0 of query tile 0 (= 1 + maxpos - 160; may be negative)
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <bool RPE>
__global__ void __launch_bounds__(fa_threads<RPE>(), 1)
    attn_fused_fwd_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                          const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_pe,
                          const FusedFwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                             // 128 x 128 B
  uint8_t* sK = sQ + 16384;                       // up to 320 (RPE: 160) x 128 B (K-major B operand of S)
  uint8_t* sV = sK + (RPE ? 20480 : 40960);       // blocks of [64 keys][128 B] (MN-major B operand of O)
  uint8_t* sPE = sV + (RPE ? 24576 : 40960);      // RPE only: 288 x 128 B (K-major B operand of QP)
  uint8_t* sP = sPE + (RPE ? FR_PE_ROWS * 128 : 0);  // blocks of [128 rows][128 B] (K-major A operand of O)
  uint8_t* sStage = sP + (RPE ? 49152 : 81920);   // RPE only: per-warp window staging
  // One CTA = one (head, utterance); it walks the query tiles with K / V resident. Barrier phases count query tiles.
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(sStage + (RPE ? 8 * FR_STAGE_BYTES : 0));  // K, V landed (once)
  uint64_t* bar_q = bar_kv + 1;      // Q (+ PE') tile landed
  uint64_t* bar_s = bar_kv + 2;      // S (+ QP) MMAs complete  (=> sQ / sPE may be refilled)
  uint64_t* bar_p = bar_kv + 3;      // P in shared memory
  uint64_t* bar_sfree = bar_kv + 4;  // last read of S done (probabilities written)  (=> next S MMA may start)
  uint64_t* bar_o = bar_kv + 5;      // PV MMA complete  (=> sP may be rewritten)
  uint64_t* bar_ofree = bar_kv + 6;  // O read out  (=> next PV MMA may start)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_kv + 7);
  constexpr int NG = fa_groups<RPE>();
  float* red_max = reinterpret_cast<float*>(bar_kv + 8);  // [NG][128] partial row maxima of the column groups
  float* red_sum = red_max + 512;                          // [NG][128] partial row sums

  const int warp = threadIdx.x >> 5;
  const int h = blockIdx.x, b = blockIdx.y;
  const int nqt = (p.Tq + FA_BM - 1) / FA_BM;
  const int nkb_all = (p.Tk + 63) >> 6;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
    if constexpr (RPE) tma_prefetch_desc(&map_pe);
    mbar_init(bar_kv, 1);
    mbar_init(bar_q, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, NG * 4);
    mbar_init(bar_sfree, NG * 4);
    mbar_init(bar_o, 1);
    mbar_init(bar_ofree, NG * 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_sync();  // barriers and tensor memory are set up: wait for the producers of q / k / v before the first load

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const int kboxes = ((p.Tk + 15) & ~15) > FA_KBOX ? 2 : 1;
      mbar_expect_tx(bar_kv, (uint32_t)kboxes * (FA_KBOX * 128u) + (uint32_t)nkb_all * 8192u);
      tma_load_4d(sK, &map_k, bar_kv, 0, 0, h, b);
      if (kboxes == 2) tma_load_4d(sK + FA_KBOX * 128, &map_k, bar_kv, 0, FA_KBOX, h, b);
      for (int kb = 0; kb < nkb_all; ++kb) tma_load_4d(sV + kb * 8192, &map_v, bar_kv, 0, kb * 64, h, b);
      for (int n = 0; n < nqt; ++n) {
        if (n > 0) mbar_wait(bar_s, (uint32_t)((n - 1) & 1));  // the MMAs that read sQ / sPE have completed
        mbar_expect_tx(bar_q, 16384u + (RPE ? (uint32_t)FR_PE_ROWS * 128u : 0u));
        if constexpr (RPE) {  // rows outside the table (negative / beyond 2*maxpos) arrive as zeros, never selected
          tma_load_4d(sPE, &map_pe, bar_q, 0, p.pe_row0 + n * FA_BM, 0, 0);
          tma_load_4d(sPE + 144 * 128, &map_pe, bar_q, 0, p.pe_row0 + n * FA_BM + 144, 0, 0);
        }
        tma_load_4d(sQ, &map_q, bar_q, 0, n * FA_BM, h, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(bar_kv, 0);
    for (int n = 0; n < nqt; ++n) {
      int tk = p.Tk;  // keys this query tile can see
      if (p.causal && n * FA_BM + FA_BM < tk) tk = n * FA_BM + FA_BM;
      const int tk16 = (tk + 15) & ~15;
      const int nkb = (tk + 63) >> 6;
      const int n1 = tk16 > FA_KBOX ? FA_KBOX : tk16;  // S is issued as one or two MMAs
      const int n2 = tk16 - n1;
      mbar_wait(bar_q, (uint32_t)(n & 1));
      if (n > 0) mbar_wait(bar_sfree, (uint32_t)((n - 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK);
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // head dim 64 = 4 x UMMA_K
          const uint64_t da = umma_smem_desc(aq + k * 32, 16, 1024);
          umma_bf16(tmem, da, umma_smem_desc(ak + k * 32, 16, 1024), umma_idesc_bf16(128, n1, 0, 0), k != 0);
          if (n2 > 0)
            umma_bf16(tmem + (uint32_t)n1, da, umma_smem_desc(ak + FA_KBOX * 128 + k * 32, 16, 1024),
                      umma_idesc_bf16(128, n2, 0, 0), k != 0);
          if constexpr (RPE) {  // QP = Q PE'^T: 288 columns as 160 + 128
            const uint32_t ape = smem_u32(sPE);
            umma_bf16(tmem + FR_QP_COL, da, umma_smem_desc(ape + k * 32, 16, 1024), umma_idesc_bf16(128, 160, 0, 0),
                      k != 0);
            umma_bf16(tmem + FR_QP_COL + 160, da, umma_smem_desc(ape + 160 * 128 + k * 32, 16, 1024),
                      umma_idesc_bf16(128, 128, 0, 0), k != 0);
          }
        }
        umma_commit(bar_s);
      }
      __syncwarp();
      mbar_wait(bar_p, (uint32_t)(n & 1));
      if (n > 0) mbar_wait(bar_ofree, (uint32_t)((n - 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ap = smem_u32(sP), av = smem_u32(sV);
        const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
        for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 64 keys = 4 x UMMA_K
            const uint64_t da = umma_smem_desc(ap + kb * 16384 + k * 32, 16, 1024);
            const uint64_t db = umma_smem_desc(av + kb * 8192 + k * 2048, 8192, 1024);
            umma_bf16(tmem + FA_O_COL, da, db, idesc, (kb | k) != 0);
          }
        }
        umma_commit(bar_o);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax + epilogue: NG x 4 warps, thread = (query row, every NG-th 32-column chunk) =====
    const int q = warp & 3;                 // TMEM lane quarter (hardware: warp id % 4)
    const int half = (warp - 2) >> 2;       // column group: chunks c with c % NG == half
    const int r = q * 32 + (int)lane_id();  // row within the tile == TMEM lane
    const uint8_t* kp = p.key_pad != nullptr ? p.key_pad + (int64_t)b * p.Tk : nullptr;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    uint64_t dseed = p.seed, doffset = p.offset;
    if (p.drop_thr != 0) resolve_seed(dseed, doffset);
    // O / rowsum of a finished tile (each column group writes 64 / NG of the 64 channels). Deferred by one tile so the
    // PV MMA and the next S MMA run under the first softmax pass of the following tile.
    auto epilogue = [&](int n, float inv) {
      const int i = n * FA_BM + r;
      const bool row_ok = i < p.Tq;
      mbar_wait(bar_o, (uint32_t)(n & 1));
      tc_fence_after();
      constexpr int CW = 64 / NG;
      uint32_t v[CW];
      if constexpr (NG == 2) tmem_ld_32x32(trow + FA_O_COL + (uint32_t)(half * CW), v);
      else tmem_ld_32x16(trow + FA_O_COL + (uint32_t)(half * CW), v);
      tmem_ld_wait();
      if (row_ok && p.out_f32 != nullptr) {
        float* d32 = p.out_f32 + ((int64_t)b * p.Tq + i) * (p.H * 64) + h * 64 + half * CW;
#pragma unroll
        for (int t = 0; t < CW; t += 4)
          *reinterpret_cast<float4*>(d32 + t) = make_float4(__uint_as_float(v[t]) * inv, __uint_as_float(v[t + 1]) * inv,
                                                            __uint_as_float(v[t + 2]) * inv, __uint_as_float(v[t + 3]) * inv);
      }
      if (row_ok) {
        __nv_bfloat16* dst = p.out + (int64_t)b * p.o_bs + (int64_t)i * p.o_ld + h * 64 + half * CW;
#pragma unroll
        for (int t = 0; t < CW; t += 8) {
          uint4 pk;
          pk.x = pack_bf16(__uint_as_float(v[t]) * inv, __uint_as_float(v[t + 1]) * inv);
          pk.y = pack_bf16(__uint_as_float(v[t + 2]) * inv, __uint_as_float(v[t + 3]) * inv);
          pk.z = pack_bf16(__uint_as_float(v[t + 4]) * inv, __uint_as_float(v[t + 5]) * inv);
          pk.w = pack_bf16(__uint_as_float(v[t + 6]) * inv, __uint_as_float(v[t + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + t) = pk;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(bar_ofree);
    };
    float inv_prev = 0.f;
    for (int n = 0; n < nqt; ++n) {
    const int i0 = n * FA_BM;
    int tk = p.Tk;
    if (p.causal && i0 + FA_BM < tk) tk = i0 + FA_BM;
    const int nkb = (tk + 63) >> 6;
    const int i = i0 + r;
    const bool row_ok = i < p.Tq;
    const int64_t prow = ((int64_t)b * p.H + h) * p.Tq + i;
    // warp-uniform: none of this warp's 32 rows exists (T = 160: three of the four lane quarters of the second tile;
    // T = 313: two of the third). Such a warp only keeps the barrier protocol going: its P rows stay whatever they
    // were (rows of P are independent in P V, and nothing of these rows is stored).
    const bool warp_ok = i0 + q * 32 < p.Tq;
    const int nchunks = warp_ok ? nkb * 2 : 0;  // 32-column chunks (TMEM columns beyond tk16 hold garbage and are masked)
    // validity bits of chunk c for THIS row: key exists, not padded (one coalesced byte load per lane + ballot), causal
    auto valid_bits = [&](int c) -> uint32_t {
      const int j = c * 32 + (int)lane_id();
      const bool ok = j < tk && !(kp != nullptr && kp[j] != 0);
      uint32_t m = __ballot_sync(0xffffffffu, ok);
      if (p.causal) {
        const int lim = i - c * 32;  // columns 0..lim are visible
        m &= lim >= 31 ? 0xffffffffu : (lim < 0 ? 0u : ((2u << lim) - 1u));
      }
      return m;
    };
    mbar_wait(bar_s, (uint32_t)(n & 1));
    tc_fence_after();
    // pass 1: row maximum of the masked, scaled (log2 domain) scores
    float m = -INFINITY;
#pragma unroll 1
    for (int c = half; c < nchunks; c += NG) {
      if (c * 32 >= tk) continue;  // warp-uniform: nothing visible in this chunk (pass 2 writes its zeros)
      uint32_t v[32];
      tmem_ld_32x32(trow + (uint32_t)(c * 32), v);
      const uint32_t vb = valid_bits(c);
      if constexpr (RPE) {
        // row i = i0 + 32q + lane, key j = 32c + u  ->  PE' column (32q - 32c + 128) + (lane + 31 - u)
        const uint32_t w0c = FR_QP_COL + (uint32_t)(q * 32 - c * 32 + 128);
        float* st = reinterpret_cast<float*>(sStage + (size_t)(warp - 2) * FR_STAGE_BYTES) +
                    (int)lane_id() * FR_STAGE_PITCH;
        {
          uint32_t w[32];
          tmem_ld_32x32(trow + w0c, w);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(st + 4 * g) = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
          tmem_ld_32x32(trow + w0c + 32, w);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(st + 32 + 4 * g) = make_uint4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]);
        }
        __syncwarp();  // (each thread re-reads only its own row; this is the compiler/memory fence)
        const float* rd = st + (int)lane_id() + 31;
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const float comb = __uint_as_float(v[t]) + rd[-t];
          v[t] = __float_as_uint(comb);
          if ((vb >> t) & 1u) m = fmaxf(m, comb * p.scale_log2);
        }
        __syncwarp();
        tmem_st_32x32(trow + (uint32_t)(c * 32), v);  // S <- S + bias, read back by pass 2 (same warp, same lanes)
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 32; ++t)
          if ((vb >> t) & 1u) m = fmaxf(m, __uint_as_float(v[t]) * p.scale_log2);
      }
    }
    if constexpr (RPE) tmem_st_wait();
    red_max[half * 128 + r] = m;
    asm volatile("bar.sync 1, %0;" ::"n"(NG * 128) : "memory");
    m = red_max[r];
#pragma unroll
    for (int g = 1; g < NG; ++g) m = fmaxf(m, red_max[g * 128 + r]);
    const float mm = m == -INFINITY ? 0.f : m;
    if (n > 0) epilogue(n - 1, inv_prev);  // (waits for PV(n-1): sP is free again after this)
    // pass 2: exponentials, partial row sum, dropout, P -> smem as the K-major SW128 A operand of the PV MMA.
    // The normaliser is applied to O at the end (PV is linear in P).
    float sum = 0.f;
    for (int c = half; c < nchunks; c += NG) {
      uint8_t* blk = sP + (c >> 1) * 16384 + r * 128;  // block = 64 keys, row r at r*128 B, chunk XOR (r & 7)
      const int cbase = (c & 1) * 4;
      if (c * 32 >= tk) {
        // warp-uniform: the whole chunk lies beyond the keys (T_k = 160: the second half of the third 64-key block).
        // P is zero there -- the operand tile needs the zeros, the saved exponentials too -- nothing else to do
        __nv_bfloat16* pz = (p.psave != nullptr && row_ok) ? p.psave + prow * p.p_ld + c * 32 : nullptr;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          *reinterpret_cast<uint4*>(blk + (((cbase + g) ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
          if (pz != nullptr && c * 32 + 8 * g + 8 <= p.p_ld) *reinterpret_cast<uint4*>(pz + 8 * g) = make_uint4(0u, 0u, 0u, 0u);
        }
        continue;
      }
      uint32_t v[32];
      tmem_ld_32x32(trow + (uint32_t)(c * 32), v);
      const uint32_t vb = valid_bits(c);
      uint32_t kb_ = 0xffffffffu;
      if (p.drop_thr != 0)
        kb_ = dropout_keep_mask32(dseed, doffset, (uint64_t)prow * attn_drop_pitch(p.Tk) + (uint64_t)(c * 32), p.drop_thr);
      tmem_ld_wait();
      // what the backward pass reads back: the un-normalised exponentials, a dropped element carries the sign bit
      // (probabilities are non-negative, so the bit is free; -0 for a dropped zero): no Philox and no exp there
      __nv_bfloat16* psv = (p.psave != nullptr && row_ok) ? p.psave + prow * p.p_ld + c * 32 : nullptr;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float ev[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int u = 8 * g + t;
          ev[t] = ((vb >> u) & 1u) ? fast_ex2(__uint_as_float(v[u]) * p.scale_log2 - mm) : 0.f;
          sum += ev[t];
        }
        const uint32_t k8 = kb_ >> (8 * g);
        uint4 pk, ps;
        pk.x = pack_bf16((k8 & 1u) ? ev[0] * p.drop_scale : 0.f, (k8 & 2u) ? ev[1] * p.drop_scale : 0.f);
        pk.y = pack_bf16((k8 & 4u) ? ev[2] * p.drop_scale : 0.f, (k8 & 8u) ? ev[3] * p.drop_scale : 0.f);
        pk.z = pack_bf16((k8 & 16u) ? ev[4] * p.drop_scale : 0.f, (k8 & 32u) ? ev[5] * p.drop_scale : 0.f);
        pk.w = pack_bf16((k8 & 64u) ? ev[6] * p.drop_scale : 0.f, (k8 & 128u) ? ev[7] * p.drop_scale : 0.f);
        *reinterpret_cast<uint4*>(blk + (((cbase + g) ^ (r & 7)) << 4)) = pk;
        if (psv != nullptr && c * 32 + 8 * g + 8 <= p.p_ld) {
          ps.x = pack_bf16((k8 & 1u) ? ev[0] : -ev[0], (k8 & 2u) ? ev[1] : -ev[1]);
          ps.y = pack_bf16((k8 & 4u) ? ev[2] : -ev[2], (k8 & 8u) ? ev[3] : -ev[3]);
          ps.z = pack_bf16((k8 & 16u) ? ev[4] : -ev[4], (k8 & 32u) ? ev[5] : -ev[5]);
          ps.w = pack_bf16((k8 & 64u) ? ev[6] : -ev[6], (k8 & 128u) ? ev[7] : -ev[7]);
          *reinterpret_cast<uint4*>(psv + 8 * g) = ps;
        }
      }
    }
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(bar_p);
    red_sum[half * 128 + r] = sum;
    asm volatile("bar.sync 1, %0;" ::"n"(NG * 128) : "memory");
    sum = red_sum[r];
#pragma unroll
    for (int g = 1; g < NG; ++g) sum += red_sum[g * 128 + r];
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    if (half == 0 && row_ok) {
      if (p.lse != nullptr) p.lse[prow] = (sum > 0.f) ? (mm + log2f(sum)) * 0.6931471805599453f : -INFINITY;
      if (p.inv_l != nullptr) p.inv_l[prow] = inv;
    }
    if (p.probs != nullptr && (p.probs_heads <= 0 || h < p.probs_heads)) {
      // normalised, undropped probabilities for the caller (overlaps the PV MMA)
      const int pchunks = warp_ok ? (int)((p.p_ld + 31) / 32) : 0;
      for (int c = half; c < pchunks; c += NG) {
        float pr[32];
        if (c < nchunks && c * 32 < tk) {
          uint32_t v[32];
          tmem_ld_32x32(trow + (uint32_t)(c * 32), v);
          const uint32_t vb = valid_bits(c);
          tmem_ld_wait();
#pragma unroll
          for (int t = 0; t < 32; ++t)
            pr[t] = ((vb >> t) & 1u) ? fast_ex2(__uint_as_float(v[t]) * p.scale_log2 - mm) * inv : 0.f;
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) pr[t] = 0.f;
        }
        if (row_ok) {
          const int j0 = c * 32;
          if (p.probs_fp32) {
            float* dst = reinterpret_cast<float*>(p.probs) + prow * p.p_ld + j0;
            if (j0 + 32 <= p.p_ld) {
#pragma unroll
              for (int t = 0; t < 32; t += 4)
                *reinterpret_cast<float4*>(dst + t) = make_float4(pr[t], pr[t + 1], pr[t + 2], pr[t + 3]);
            } else {
#pragma unroll
              for (int t = 0; t < 32; ++t)
                if (j0 + t < p.p_ld) dst[t] = pr[t];
            }
          } else {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.probs) + prow * p.p_ld + j0;
            if ((p.p_ld & 7) == 0 && (reinterpret_cast<uintptr_t>(p.probs) & 15) == 0) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                if (j0 + 8 * g + 8 <= p.p_ld)
                  *reinterpret_cast<uint4*>(dst + 8 * g) =
                      make_uint4(pack_bf16(pr[8 * g], pr[8 * g + 1]), pack_bf16(pr[8 * g + 2], pr[8 * g + 3]),
                                 pack_bf16(pr[8 * g + 4], pr[8 * g + 5]), pack_bf16(pr[8 * g + 6], pr[8 * g + 7]));
            } else {
#pragma unroll
              for (int t = 0; t < 32; ++t)
                if (j0 + t < p.p_ld) dst[t] = __float2bfloat16(pr[t]);
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane_id() == 0) mbar_arrive(bar_sfree);
    inv_prev = inv;
    }
    epilogue(nqt - 1, inv_prev);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

static int make_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t ld, int64_t bs, int H, int B, int box_rows) {
  const uint64_t dims[4] = {64, (uint64_t)rows, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)ld * 2, 128, (uint64_t)bs * 2};
  const uint32_t box[4] = {64, (uint32_t)box_rows, 1, 1};
  return encode_bf16_map_4d(m, ptr, dims, strides, box);
}

}  // namespace st5

using namespace st5;

extern "C" int st5_attn_fused_fwd(const st5_attn_args* a, float* lse, void* psave, float* inv_l, float* out_f32,
                                  void* stream) {
  const bool rpe = a->pe_k != nullptr;
  if (a->dtype != ST5_BF16 || a->Tk > FA_MAX_TK || a->Tk <= 0 || a->Tq <= 0)
    return set_error(-2, "st5_attn_fused_fwd: needs bf16 and Tk <= 320");
  if (rpe && (a->causal || a->maxpos <= 0 || a->maxpos > FR_MAX_T || a->Tk > a->maxpos || a->Tq > a->maxpos))
    return set_error(-5, "st5_attn_fused_fwd: relative positions need Tq, Tk <= maxpos <= 160 (no clipping), no causal mask");
  if ((a->probs != nullptr || psave != nullptr) && a->p_ld < a->Tk) return set_error(-3, "st5_attn_fused_fwd");
  if (psave != nullptr && ((a->p_ld & 7) || (reinterpret_cast<uintptr_t>(psave) & 15) || inv_l == nullptr))
    return set_error(-3, "st5_attn_fused_fwd: psave needs a row pitch that is a multiple of 8, 16-byte alignment and inv_l");
  if ((a->o_ld & 7) || (a->o_bs & 7) || (reinterpret_cast<uintptr_t>(a->out) & 15))
    return set_error(-4, "st5_attn_fused_fwd: out must be 16-byte aligned");
  CUtensorMap mq, mk, mv;
  int rc = make_map(&mq, a->q, a->Tq, a->q_ld, a->q_bs, a->H, a->B, FA_BM);
  if (!rc) rc = make_map(&mk, a->k, a->Tk, a->k_ld, a->k_bs, a->H, a->B, FA_KBOX);
  if (!rc) rc = make_map(&mv, a->v, a->Tk, a->v_ld, a->v_bs, a->H, a->B, 64);
  CUtensorMap mpe = mq;
  if (!rc && rpe) {  // the bf16 table [2*maxpos][64] as a rank-4 map with unit outer dimensions
    const uint64_t dims[4] = {64, (uint64_t)(2 * a->maxpos), 1, 1};
    const uint64_t strides[3] = {128, (uint64_t)(2 * a->maxpos) * 128, (uint64_t)(2 * a->maxpos) * 128};
    const uint32_t box[4] = {64, 144, 1, 1};
    rc = encode_bf16_map_4d(&mpe, a->pe_k, dims, strides, box);
  }
  if (rc) return set_error(rc, "st5_attn_fused_fwd: tensor map");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fused_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)FA_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_fused_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FR_SMEM);
    if (e != cudaSuccess) return set_error((int)e, "st5_attn_fused_fwd");
    attr_set = true;
  }
  FusedFwdParams p;
  p.B = a->B; p.H = a->H; p.Tq = a->Tq; p.Tk = a->Tk; p.causal = a->causal;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.key_pad = a->key_pad;
  p.out = (__nv_bfloat16*)a->out; p.o_ld = a->o_ld; p.o_bs = a->o_bs;
  p.lse = lse;
  p.psave = reinterpret_cast<__nv_bfloat16*>(psave);
  p.inv_l = inv_l;
  p.out_f32 = out_f32;
  p.probs = a->probs; p.probs_fp32 = a->probs_dtype == ST5_F32; p.p_ld = a->p_ld;
  p.probs_heads = a->probs_heads;
  p.drop_thr = drop_threshold(a->drop_p);
  p.drop_scale = a->drop_p > 0.f ? 1.f / (1.f - a->drop_p) : 1.f;
  p.seed = a->seed; p.offset = a->offset;
  p.pe_row0 = 1 + a->maxpos - FR_MAX_T;
  dim3 grid(a->H, a->B);  // one persistent CTA per (head, utterance)
  if (rpe)
    launch_pdl(attn_fused_fwd_kernel<true>, dim3(grid), dim3(fa_threads<true>()), FR_SMEM, (cudaStream_t)stream, mq, mk, mv, mpe, p);
  else
    launch_pdl(attn_fused_fwd_kernel<false>, dim3(grid), dim3(fa_threads<false>()), FA_SMEM, (cudaStream_t)stream, mq, mk, mv, mpe, p);
  return set_error((int)cudaGetLastError(), "st5_attn_fused_fwd");
}
