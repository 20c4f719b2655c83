// Row kernels that sit between the tensor-core GEMMs of the bf16 attention path:
//   forward : S = scale*Q.K^T (GEMM) [+ QP = scale*Q.PE^T (GEMM)] -> softmax_rpe (here) -> O = dropout(P).V (GEMM)
//   backward: dP = dO.V^T (GEMM) -> ds (here) -> dV = Pd^T.dO, dQ = dS.K, dK = dS^T.Q (GEMMs)
//             [-> dqp_scatter (here) -> dQ += dQP.PE, dPE = dQP^T.Q (GEMMs)]
// Semantics: speecht5/models/modules/multihead_attention.py:340-389; the relative-position bias (:343-353) is read from
// QP[i][clamp(i-j,-maxpos,maxpos-1)+maxpos] instead of materialising pos_k [T,T,64] (encoder.py:239-246).
// All are HBM-bound: one warp per attention row, coalesced row reads/writes, fp32 math, bf16 storage for GEMM operands.
#include "kernels.cuh"
#include "ptx.cuh"

namespace st5 {

constexpr int ROW_WARPS = 4;
constexpr int MAX_PER_LANE = 16;  // Tk <= 512

__device__ __forceinline__ int rel_idx(int i, int j, int maxpos) {
  int d = i - j;
  d = d < -maxpos ? -maxpos : (d >= maxpos ? maxpos - 1 : d);
  return d + maxpos;
}

// one warp per row (b,h,i)
__global__ void __launch_bounds__(ROW_WARPS * 32)
    attn_softmax_fwd_kernel(const float* __restrict__ S, const float* __restrict__ QP, int64_t qp_ld,
                            const uint8_t* __restrict__ key_pad, __nv_bfloat16* __restrict__ P,
                            float* __restrict__ probs_f32, __nv_bfloat16* __restrict__ Pd, int B, int H, int Tq, int Tk,
                            int64_t p_ld, int causal, int maxpos, uint32_t thr, float dscale, uint64_t seed,
                            uint64_t offset) {
  if (thr != 0) resolve_seed(seed, offset);
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int64_t nrows = (int64_t)B * H * Tq;
  if (row >= nrows) return;
  const int i = (int)(row % Tq);
  const int b = (int)(row / ((int64_t)H * Tq));
  const float* srow = S + row * p_ld;
  const float* qrow = QP != nullptr ? QP + row * qp_ld : nullptr;
  const uint8_t* kp = key_pad != nullptr ? key_pad + (int64_t)b * Tk : nullptr;
  float v[MAX_PER_LANE];
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < MAX_PER_LANE; ++k) {
    const int j = k * 32 + lane;
    float s = -INFINITY;
    if (j < Tk && !(causal && j > i) && !(kp != nullptr && kp[j] != 0)) {
      s = srow[j];
      if (qrow != nullptr) s += qrow[rel_idx(i, j, maxpos)];
    }
    v[k] = s;
    m = fmaxf(m, s);
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < MAX_PER_LANE; ++k) {
    const float e = (k * 32 < Tk) ? __expf(v[k] - m) : 0.f;
    v[k] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  __nv_bfloat16* prow = P + row * p_ld;
  __nv_bfloat16* pdrow = Pd != nullptr ? Pd + row * p_ld : nullptr;
  float* frow = probs_f32 != nullptr ? probs_f32 + row * p_ld : nullptr;
#pragma unroll
  for (int k = 0; k < MAX_PER_LANE; ++k) {
    const int j = k * 32 + lane;
    if (j < (int)p_ld) {
      const float p = j < Tk ? v[k] * inv : 0.f;
      const __nv_bfloat16 pb = __float2bfloat16(p);
      prow[j] = pb;
      if (frow != nullptr) frow[j] = p;
      if (pdrow != nullptr) {
        float pd = __bfloat162float(pb);
        if (thr != 0 && j < Tk)
          pd = dropout_keep(seed, offset, (uint64_t)row * attn_drop_pitch(Tk) + (uint64_t)j, thr) ? pd * dscale : 0.f;
        pdrow[j] = __float2bfloat16(pd);
      }
    }
  }
}

// dS = P * (dropout_bwd(dP) + dP_ext - delta), delta = sum_j P * (...)
__global__ void __launch_bounds__(ROW_WARPS * 32)
    attn_ds_kernel(const __nv_bfloat16* __restrict__ P, const float* __restrict__ dP, const float* __restrict__ dPx,
                   __nv_bfloat16* __restrict__ dS, __nv_bfloat16* __restrict__ Pd, int64_t nrows, int Tk, int64_t p_ld,
                   uint32_t thr, float dscale, uint64_t seed, uint64_t offset) {
  if (thr != 0) resolve_seed(seed, offset);
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= nrows) return;
  const __nv_bfloat16* prow = P + row * p_ld;
  const float* dprow = dP + row * p_ld;
  const float* dxrow = dPx != nullptr ? dPx + row * p_ld : nullptr;
  float pv[MAX_PER_LANE], g[MAX_PER_LANE];
  float delta = 0.f;
#pragma unroll
  for (int k = 0; k < MAX_PER_LANE; ++k) {
    const int j = k * 32 + lane;
    float p = 0.f, d = 0.f;
    if (j < Tk) {
      p = __bfloat162float(prow[j]);
      d = dprow[j];
      bool keep = true;
      if (thr != 0) {
        keep = dropout_keep(seed, offset, (uint64_t)row * attn_drop_pitch(Tk) + (uint64_t)j, thr);
        d = keep ? d * dscale : 0.f;
      }
      if (Pd != nullptr) Pd[row * p_ld + j] = __float2bfloat16(keep ? p * dscale : 0.f);
      if (dxrow != nullptr) d += dxrow[j];
    } else if (j < (int)p_ld && Pd != nullptr) {
      Pd[row * p_ld + j] = __float2bfloat16(0.f);
    }
    pv[k] = p; g[k] = d;
    delta += p * d;
  }
  delta = warp_sum(delta);
  __nv_bfloat16* dsrow = dS + row * p_ld;
#pragma unroll
  for (int k = 0; k < MAX_PER_LANE; ++k) {
    const int j = k * 32 + lane;
    if (j < (int)p_ld) dsrow[j] = __float2bfloat16(j < Tk ? pv[k] * (g[k] - delta) : 0.f);
  }
}

// dQP[row][r] = sum_{j : idx(i,j) = r} dS[row][j]  (one warp per row; interior offsets map 1:1, the two clamped ends sum)
__global__ void __launch_bounds__(ROW_WARPS * 32)
    attn_dqp_scatter_kernel(const __nv_bfloat16* __restrict__ dS, __nv_bfloat16* __restrict__ dQP, int64_t nrows, int Tq,
                            int Tk, int64_t p_ld, int maxpos, int B, int H, int h_major) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  if (row >= nrows) return;
  const int i = (int)(row % Tq);
  const int R = 2 * maxpos;
  const __nv_bfloat16* dsrow = dS + row * p_ld;
  // output row: (b, h, i) like dS, or head-major (h, b, i) -- then one head's rows are equidistant in memory and the
  // two table contractions run as 12 long GEMMs instead of B x H short ones
  int64_t orow_idx = row;
  if (h_major) {
    const int h = (int)((row / Tq) % H), b = (int)(row / ((int64_t)Tq * H));
    orow_idx = ((int64_t)h * B + b) * Tq + i;
  }
  __nv_bfloat16* orow = dQP + orow_idx * R;
  // clamped ends (nothing clips while the sequence fits the table: T <= maxpos at the TTS shapes -- no reduction then)
  float lo = 0.f, hi = 0.f;
  if (i + maxpos < Tk || i >= maxpos - 1) {  // (warp-uniform: one row per warp)
    for (int j = i + maxpos + lane; j < Tk; j += 32) lo += __bfloat162float(dsrow[j]);       // i-j <= -maxpos
    for (int j = lane; j <= i - (maxpos - 1) && j < Tk; j += 32) hi += __bfloat162float(dsrow[j]);  // i-j >= maxpos-1
    lo = warp_sum(lo);
    hi = warp_sum(hi);
  }
  if ((R & 7) == 0 && (reinterpret_cast<uintptr_t>(dQP) & 15) == 0) {
    // eight table rows per lane and 16-byte store (the reads are the same 2-byte gathers, reversed within the group)
    for (int g = lane; g < (R >> 3); g += 32) {
      float v[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int r = g * 8 + t;
        const int j = i - (r - maxpos);
        v[t] = r == 0 ? lo : (r == R - 1 ? hi : ((j >= 0 && j < Tk) ? __bfloat162float(dsrow[j]) : 0.f));
      }
      uint4 pk;
      __nv_bfloat162 t0 = __floats2bfloat162_rn(v[0], v[1]), t1 = __floats2bfloat162_rn(v[2], v[3]);
      __nv_bfloat162 t2 = __floats2bfloat162_rn(v[4], v[5]), t3 = __floats2bfloat162_rn(v[6], v[7]);
      pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
      pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
      *reinterpret_cast<uint4*>(orow + g * 8) = pk;
    }
    return;
  }
  for (int r = lane; r < R; r += 32) {
    float v;
    if (r == 0) v = lo;
    else if (r == R - 1) v = hi;
    else {
      const int j = i - (r - maxpos);
      v = (j >= 0 && j < Tk) ? __bfloat162float(dsrow[j]) : 0.f;
    }
    orow[r] = __float2bfloat16(v);
  }
}

}  // namespace st5

using namespace st5;

namespace st5 { int set_error(int code, const char* where); }

extern "C" {

int st5_attn_softmax_fwd(const float* s, const float* qp, int64_t qp_ld, const uint8_t* key_pad, void* p_bf16,
                         float* probs_f32, void* pdrop_bf16, int32_t B, int32_t H, int32_t Tq, int32_t Tk, int64_t p_ld,
                         int32_t causal, int32_t maxpos, float drop_p, uint64_t seed, uint64_t offset, void* stream) {
  if (Tk > 32 * MAX_PER_LANE || p_ld > 32 * MAX_PER_LANE || p_ld < Tk) return set_error(-2, "st5_attn_softmax_fwd");
  const int64_t nrows = (int64_t)B * H * Tq;
  if (nrows == 0) return 0;
  const uint32_t thr = drop_threshold(drop_p);
  const float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  attn_softmax_fwd_kernel<<<(unsigned)((nrows + ROW_WARPS - 1) / ROW_WARPS), ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(
      s, qp, qp_ld, key_pad, (__nv_bfloat16*)p_bf16, probs_f32, (__nv_bfloat16*)pdrop_bf16, B, H, Tq, Tk, p_ld, causal,
      maxpos, thr, ds, seed, offset);
  return set_error((int)cudaGetLastError(), "st5_attn_softmax_fwd");
}

int st5_attn_ds(const void* p_bf16, const float* dp, const float* dp_ext, void* ds_bf16, void* pdrop_bf16, int32_t B,
                int32_t H, int32_t Tq, int32_t Tk, int64_t p_ld, float drop_p, uint64_t seed, uint64_t offset,
                void* stream) {
  if (Tk > 32 * MAX_PER_LANE || p_ld > 32 * MAX_PER_LANE || p_ld < Tk) return set_error(-2, "st5_attn_ds");
  const int64_t nrows = (int64_t)B * H * Tq;
  if (nrows == 0) return 0;
  const uint32_t thr = drop_threshold(drop_p);
  const float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  attn_ds_kernel<<<(unsigned)((nrows + ROW_WARPS - 1) / ROW_WARPS), ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)p_bf16, dp, dp_ext, (__nv_bfloat16*)ds_bf16, (__nv_bfloat16*)pdrop_bf16, nrows, Tk, p_ld, thr,
      ds, seed, offset);
  return set_error((int)cudaGetLastError(), "st5_attn_ds");
}

int st5_attn_dqp_scatter(const void* ds_bf16, void* dqp_bf16, int32_t B, int32_t H, int32_t Tq, int32_t Tk,
                         int64_t p_ld, int32_t maxpos, int32_t h_major, void* stream) {
  const int64_t nrows = (int64_t)B * H * Tq;
  if (nrows == 0) return 0;
  launch_pdl(attn_dqp_scatter_kernel, dim3((unsigned)((nrows + ROW_WARPS - 1) / ROW_WARPS)), dim3(ROW_WARPS * 32), 0,
             (cudaStream_t)stream, (const __nv_bfloat16*)ds_bf16, (__nv_bfloat16*)dqp_bf16, nrows, Tq, Tk, p_ld, maxpos, B, H,
             h_major);
  return set_error((int)cudaGetLastError(), "st5_attn_dqp_scatter");
}

}  // extern "C"
