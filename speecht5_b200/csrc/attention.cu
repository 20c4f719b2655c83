// Exact (fp32-math) multi-head attention with the SpeechT5 relative-position bias, forward and backward.
//
// Reference: speecht5/models/modules/multihead_attention.py:232-405 and the Toeplitz position table of
// encoder.py:40-59,239-246. The reference materialises pos_k = pe_k[clamp(i-j)+maxpos] as a [T,T,64] tensor and
// runs T small matmuls (:346-353); here the bias is folded into the score as q_i . (k_j + pe_k[idx(i,j)]) so nothing
// of size T*T*64 ever exists.
//
// These are the "row" kernels: one CTA per (batch, head, query row) -- simple, exact, any T, any mask. They serve the
// fp32 parity mode and as the fallback for shapes the tensor-core path does not cover.
#include "kernels.cuh"
#include "ptx.cuh"

namespace st5 {

constexpr int AT = 128;  // threads per CTA
constexpr int HD = 64;   // head dim

template <typename T> __device__ __forceinline__ void load_row64(const T* p, float* out);
template <> __device__ __forceinline__ void load_row64<float>(const float* p, float* out) {
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float4 v = reinterpret_cast<const float4*>(p)[k];
    out[4 * k] = v.x; out[4 * k + 1] = v.y; out[4 * k + 2] = v.z; out[4 * k + 3] = v.w;
  }
}
template <> __device__ __forceinline__ void load_row64<__nv_bfloat16>(const __nv_bfloat16* p, float* out) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const uint4 v = reinterpret_cast<const uint4*>(p)[k];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = __bfloat1622float2(h[t]);
      out[8 * k + 2 * t] = f.x; out[8 * k + 2 * t + 1] = f.y;
    }
  }
}

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  v = is_max ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int w = 1; w < AT / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}
__device__ __forceinline__ int rel_index(int i, int j, int maxpos) {
  int d = i - j;
  d = d < -maxpos ? -maxpos : (d >= maxpos ? maxpos - 1 : d);
  return d + maxpos;
}

template <typename T, typename PT>
__global__ void __launch_bounds__(AT) attn_fwd_row(const st5_attn_args a) {
  extern __shared__ float sm[];
  float* q = sm;            // [64]
  float* red = sm + HD;     // [4]
  float* part = red + 8;    // [2][64]
  float* sc = part + 2 * HD;  // [Tk]
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const T* qp = (const T*)a.q + (int64_t)b * a.q_bs + (int64_t)i * a.q_ld + h * HD;
  if (tid < HD) q[tid] = ldf(qp + tid);
  __syncthreads();
  const T* kbase = (const T*)a.k + (int64_t)b * a.k_bs + h * HD;
  float lmax = -INFINITY;
  for (int j = tid; j < a.Tk; j += AT) {
    float s;
    const bool masked = (a.causal && j > i) || (a.key_pad != nullptr && a.key_pad[(int64_t)b * a.Tk + j] != 0);
    if (masked) {
      s = -INFINITY;
    } else {
      float kr[HD];
      load_row64<T>(kbase + (int64_t)j * a.k_ld, kr);
      float acc = 0.f;
      if (a.pe_k != nullptr) {
        const float* pr = a.pe_k + (int64_t)rel_index(i, j, a.maxpos) * HD;
#pragma unroll
        for (int c = 0; c < HD; ++c) acc += q[c] * (kr[c] + pr[c]);
      } else {
#pragma unroll
        for (int c = 0; c < HD; ++c) acc += q[c] * kr[c];
      }
      s = acc * a.scale;
    }
    sc[j] = s;
    lmax = fmaxf(lmax, s);
  }
  const float m = block_reduce(lmax, red, true);
  float lsum = 0.f;
  for (int j = tid; j < a.Tk; j += AT) {
    const float e = expf(sc[j] - m);  // exp(-inf) = 0 for masked keys; a fully masked row gives NaN like the reference
    sc[j] = e;
    lsum += e;
  }
  const float denom = block_reduce(lsum, red, false);
  const float inv = 1.f / denom;
  const uint32_t thr = drop_threshold(a.drop_p);
  const float dscale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  uint64_t dseed = a.seed, doffset = a.offset;
  if (thr != 0) resolve_seed(dseed, doffset);
  const int64_t prow = (((int64_t)b * a.H + h) * a.Tq + i);
  PT* pout = a.probs != nullptr ? (PT*)a.probs + prow * a.p_ld : nullptr;
  for (int j = tid; j < (int)a.p_ld; j += AT) {
    if (j < a.Tk) {
      float p = sc[j] * inv;
      if (pout != nullptr) { stf(pout + j, p); p = ldf(pout + j); }
      if (thr != 0) p = dropout_keep(dseed, doffset, (uint64_t)prow * attn_drop_pitch(a.Tk) + (uint64_t)j, thr) ? p * dscale : 0.f;
      sc[j] = p;
    } else if (pout != nullptr) {
      stf(pout + j, 0.f);
    }
  }
  __syncthreads();
  // out_c = sum_j p_j v_jc : two halves of the key range, threads of a half cover the 64 channels (coalesced)
  const int c = tid & (HD - 1), half = tid >> 6;
  const T* vbase = (const T*)a.v + (int64_t)b * a.v_bs + h * HD + c;
  const int jend = a.causal ? (i + 1 < a.Tk ? i + 1 : a.Tk) : a.Tk;
  float acc = 0.f;
  for (int j = half; j < jend; j += 2) acc += sc[j] * ldf(vbase + (int64_t)j * a.v_ld);
  part[half * HD + c] = acc;
  __syncthreads();
  if (tid < HD) {
    T* op = (T*)a.out + (int64_t)b * a.o_bs + (int64_t)i * a.o_ld + h * HD;
    stf(op + tid, part[tid] + part[HD + tid]);
  }
}

// backward, per query row: dP -> dS (stored fp32) and dq
template <typename T, typename PT>
__global__ void __launch_bounds__(AT) attn_bwd_q_row(const st5_attn_args a) {
  extern __shared__ float sm[];
  float* dO = sm;             // [64]
  float* red = sm + HD;       // [8]
  float* part = red + 8;      // [2][64]
  float* dsr = part + 2 * HD; // [Tk]
  const int i = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const T* dop = (const T*)a.dout + (int64_t)b * a.o_bs + (int64_t)i * a.o_ld + h * HD;
  if (tid < HD) dO[tid] = ldf(dop + tid);
  __syncthreads();
  const int64_t prow = (((int64_t)b * a.H + h) * a.Tq + i);
  const PT* prob = (const PT*)a.probs + prow * a.p_ld;
  const float* dpe = a.dprobs_ext != nullptr ? a.dprobs_ext + prow * a.p_ld : nullptr;
  const uint32_t thr = drop_threshold(a.drop_p);
  const float dscale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  uint64_t dseed = a.seed, doffset = a.offset;
  if (thr != 0) resolve_seed(dseed, doffset);
  const T* vbase = (const T*)a.v + (int64_t)b * a.v_bs + h * HD;
  float ldelta = 0.f;
  for (int j = tid; j < a.Tk; j += AT) {
    const float p = ldf(prob + j);
    float dp = 0.f;
    if (p != 0.f) {
      float vr[HD];
      load_row64<T>(vbase + (int64_t)j * a.v_ld, vr);
#pragma unroll
      for (int c = 0; c < HD; ++c) dp += dO[c] * vr[c];
      if (thr != 0) dp = dropout_keep(dseed, doffset, (uint64_t)prow * attn_drop_pitch(a.Tk) + (uint64_t)j, thr) ? dp * dscale : 0.f;
    }
    if (dpe != nullptr) dp += dpe[j];
    dsr[j] = dp;
    ldelta += p * dp;
  }
  const float delta = block_reduce(ldelta, red, false);
  float* ds_out = a.ds + prow * a.p_ld;
  for (int j = tid; j < (int)a.p_ld; j += AT) {
    float d = 0.f;
    if (j < a.Tk) {
      d = ldf(prob + j) * (dsr[j] - delta);
      dsr[j] = d;
    }
    ds_out[j] = d;
  }
  __syncthreads();
  const int c = tid & (HD - 1), half = tid >> 6;
  const T* kbase = (const T*)a.k + (int64_t)b * a.k_bs + h * HD + c;
  const int jend = a.causal ? (i + 1 < a.Tk ? i + 1 : a.Tk) : a.Tk;
  float acc = 0.f;
  for (int j = half; j < jend; j += 2) {
    float kv = ldf(kbase + (int64_t)j * a.k_ld);
    if (a.pe_k != nullptr) kv += a.pe_k[(int64_t)rel_index(i, j, a.maxpos) * HD + c];
    acc += dsr[j] * kv;
  }
  part[half * HD + c] = acc;
  __syncthreads();
  if (tid < HD) {
    T* dqp = (T*)a.dq + (int64_t)b * a.q_bs + (int64_t)i * a.q_ld + h * HD;
    stf(dqp + tid, (part[tid] + part[HD + tid]) * a.scale);
  }
}

// backward, per key row: dk_j = scale * sum_i dS_ij q_i ; dv_j = sum_i dropout(P)_ij dO_i
template <typename T, typename PT>
__global__ void __launch_bounds__(AT) attn_bwd_kv_row(const st5_attn_args a) {
  __shared__ float part[2][2][HD];
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, c = tid & (HD - 1), half = tid >> 6;
  const uint32_t thr = drop_threshold(a.drop_p);
  const float dscale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  uint64_t dseed = a.seed, doffset = a.offset;
  if (thr != 0) resolve_seed(dseed, doffset);
  const int64_t bh = (int64_t)b * a.H + h;
  const T* qb = (const T*)a.q + (int64_t)b * a.q_bs + h * HD + c;
  const T* dob = (const T*)a.dout + (int64_t)b * a.o_bs + h * HD + c;
  float dk = 0.f, dv = 0.f;
  const int i0 = a.causal ? j : 0;
  for (int i = i0 + half; i < a.Tq; i += 2) {
    const int64_t prow = bh * a.Tq + i;
    const float ds = a.ds[prow * a.p_ld + j];
    float p = ldf((const PT*)a.probs + prow * a.p_ld + j);
    if (thr != 0) p = dropout_keep(dseed, doffset, (uint64_t)prow * attn_drop_pitch(a.Tk) + (uint64_t)j, thr) ? p * dscale : 0.f;
    dk += ds * ldf(qb + (int64_t)i * a.q_ld);
    dv += p * ldf(dob + (int64_t)i * a.o_ld);
  }
  part[0][half][c] = dk;
  part[1][half][c] = dv;
  __syncthreads();
  if (tid < HD) {
    T* dkp = (T*)a.dk + (int64_t)b * a.k_bs + (int64_t)j * a.k_ld + h * HD;
    T* dvp = (T*)a.dv + (int64_t)b * a.v_bs + (int64_t)j * a.v_ld + h * HD;
    stf(dkp + tid, (part[0][0][tid] + part[0][1][tid]) * a.scale);
    stf(dvp + tid, part[1][0][tid] + part[1][1][tid]);
  }
}

// backward of the position table: dpe_k[r] += scale * sum_{b,h} sum_{(i,j): idx(i,j)=r} dS_ij q_i
template <typename T>
__global__ void __launch_bounds__(AT) attn_bwd_pe(const st5_attn_args a) {
  __shared__ float part[2][HD];
  const int r = blockIdx.x;
  const int64_t bh = blockIdx.y;
  const int b = (int)(bh / a.H), h = (int)(bh % a.H);
  const int tid = threadIdx.x, c = tid & (HD - 1), half = tid >> 6;
  const T* qb = (const T*)a.q + (int64_t)b * a.q_bs + h * HD + c;
  const int d = r - a.maxpos;  // i - j (clamped)
  float acc = 0.f;
  for (int i = half; i < a.Tq; i += 2) {
    int jlo, jhi;  // keys j with clamp(i-j) == d
    if (r == 0) { jlo = i + a.maxpos; jhi = a.Tk - 1; }
    else if (r == 2 * a.maxpos - 1) { jlo = 0; jhi = i - (a.maxpos - 1); }
    else { jlo = jhi = i - d; }
    if (jlo < 0) jlo = 0;
    if (jhi > a.Tk - 1) jhi = a.Tk - 1;
    if (jlo > jhi) continue;
    const float* dsrow = a.ds + (bh * a.Tq + i) * a.p_ld;
    float s = 0.f;
    for (int j = jlo; j <= jhi; ++j) s += dsrow[j];
    acc += s * ldf(qb + (int64_t)i * a.q_ld);
  }
  part[half][c] = acc;
  __syncthreads();
  if (tid < HD) {
    const float v = (part[0][tid] + part[1][tid]) * a.scale;
    if (v != 0.f) atomicAdd(a.dpe_k + (int64_t)r * HD + tid, v);
  }
}

static size_t row_smem(const st5_attn_args& a) { return sizeof(float) * (HD + 8 + 2 * HD + (size_t)a.Tk + 8); }

template <typename T, typename PT>
static int fwd_dispatch(const st5_attn_args& a, cudaStream_t s) {
  const size_t smem = row_smem(a);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_row<T, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  attn_fwd_row<T, PT><<<dim3(a.Tq, a.H, a.B), AT, smem, s>>>(a);
  return (int)cudaGetLastError();
}
template <typename T, typename PT>
static int bwd_dispatch(const st5_attn_args& a, cudaStream_t s) {
  const size_t smem = row_smem(a);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_q_row<T, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  attn_bwd_q_row<T, PT><<<dim3(a.Tq, a.H, a.B), AT, smem, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  attn_bwd_kv_row<T, PT><<<dim3(a.Tk, a.H, a.B), AT, 0, s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (a.pe_k != nullptr && a.dpe_k != nullptr) {
    attn_bwd_pe<T><<<dim3(2 * a.maxpos, a.B * a.H), AT, 0, s>>>(a);
    e = cudaGetLastError();
  }
  return (int)e;
}

static int check_args(const st5_attn_args& a, bool bwd) {
  if (a.B <= 0 || a.H <= 0 || a.Tq <= 0 || a.Tk <= 0) return -2;
  if (a.p_ld < a.Tk && (a.probs != nullptr || bwd)) return -3;
  if (a.pe_k != nullptr && a.maxpos <= 0) return -4;
  if (bwd && (a.probs == nullptr || a.ds == nullptr || a.dout == nullptr || a.dq == nullptr || a.dk == nullptr ||
              a.dv == nullptr))
    return -5;
  // 16-byte row loads
  const int esz = a.dtype == ST5_F32 ? 4 : 2;
  if (((a.k_ld * esz) & 15) || ((a.v_ld * esz) & 15) || ((a.k_bs * esz) & 15) || ((a.v_bs * esz) & 15)) return -6;
  if ((reinterpret_cast<uintptr_t>(a.k) & 15) || (reinterpret_cast<uintptr_t>(a.v) & 15)) return -6;
  return 0;
}

int attn_fwd_launch(const st5_attn_args& a, cudaStream_t s) {
  int rc = check_args(a, false);
  if (rc) return rc;
  if (a.dtype == ST5_F32) {
    return a.probs_dtype == ST5_F32 ? fwd_dispatch<float, float>(a, s) : fwd_dispatch<float, __nv_bfloat16>(a, s);
  }
  return a.probs_dtype == ST5_F32 ? fwd_dispatch<__nv_bfloat16, float>(a, s)
                                  : fwd_dispatch<__nv_bfloat16, __nv_bfloat16>(a, s);
}
int attn_bwd_launch(const st5_attn_args& a, cudaStream_t s) {
  int rc = check_args(a, true);
  if (rc) return rc;
  if (a.dtype == ST5_F32) {
    return a.probs_dtype == ST5_F32 ? bwd_dispatch<float, float>(a, s) : bwd_dispatch<float, __nv_bfloat16>(a, s);
  }
  return a.probs_dtype == ST5_F32 ? bwd_dispatch<__nv_bfloat16, float>(a, s)
                                  : bwd_dispatch<__nv_bfloat16, __nv_bfloat16>(a, s);
}

}  // namespace st5
