// CTC negative log-likelihood and its gradient with respect to the LOGITS of the encoder's CTC head, log-softmax fused in
// (reference: speecht5/criterions/speech_to_text_loss.py:303-335 -- F.log_softmax via get_normalized_probs_for_ctc, then
// F.ctc_loss(reduction="sum", zero_infinity=...) with cuDNN off). One CTA per utterance, one thread per position s of
// the extended label sequence l' = (blank, l1, blank, ..., lL, blank); the T time steps are sequential:
//   forward : alpha_t(s) = lp_t(l'_s) + logsumexp(alpha_{t-1}(s), alpha_{t-1}(s-1), [alpha_{t-1}(s-2) if l'_s != blank
//             and l'_s != l'_{s-2}]), kept in global scratch for the backward sweep; nll = -logsumexp of the two final states
//   backward: beta recursion from the end; at every step the gradient row is finished on the spot,
//             dlogits[t, k] = softmax_t(k) - sum_{s: l'_s = k} exp(alpha_t(s) + beta_t(s) - lp_t(k) + nll)
//             (the per-symbol sums are accumulated in shared memory in the linear domain: every term is <= 1).
// Rows t >= input_length get a zero gradient; an infeasible utterance gives nll = +inf, or 0 with zero gradient under
// zero_infinity. The algorithm is restated and checked against torch on the CPU in tests/test_kernel_algorithms_cpu.py.
//
// Written at the end of round 1 without GPU time (EXPERIMENTAL: the criterion uses it only with ST5_CTC_KERNEL=1).
#include "kernels.cuh"
#include <math_constants.h>

namespace st5 {

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + log1pf(__expf(-fabsf(a - b)));
}

// lse[t * B + b] = logsumexp_k logits[t, b, k]; one warp per row
__global__ void ctc_lse_kernel(const float* __restrict__ logits, int64_t ld_t, int64_t ld_b, float* __restrict__ lse,
                               int T, int B, int V) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= T * B) return;
  const int t = row / B, b = row - t * B;
  const float* x = logits + (int64_t)t * ld_t + (int64_t)b * ld_b;
  const int lane = threadIdx.x & 31;
  float m = -CUDART_INF_F;
  for (int k = lane; k < V; k += 32) m = fmaxf(m, x[k]);
  m = warp_max(m);
  float s = 0.f;
  for (int k = lane; k < V; k += 32) s += __expf(x[k] - m);
  s = warp_sum(s);
  if (lane == 0) lse[row] = m + __logf(s);
}

__global__ void ctc_alpha_beta_kernel(const float* __restrict__ logits, int64_t ld_t, int64_t ld_b,
                                      const float* __restrict__ lse, const int64_t* __restrict__ targets,
                                      const int64_t* __restrict__ tgt_offsets, const int64_t* __restrict__ input_lengths,
                                      const int64_t* __restrict__ target_lengths, float* __restrict__ nll_out,
                                      float* __restrict__ grad, float* __restrict__ alpha_ws, int T, int B, int V,
                                      int S_max, int blank, int zero_infinity) {
  extern __shared__ float sm[];
  float* prev = sm;                 // [S_max] state of the previous time step
  float* cur = sm + S_max;          // [S_max]
  float* acc = sm + 2 * S_max;      // [V] per-symbol posterior mass of the current step
  __shared__ float nll_s;
  const int b = blockIdx.x, s = threadIdx.x;
  const int Tn = (int)min((int64_t)T, input_lengths[b]);
  const int L = (int)target_lengths[b];
  const int S = 2 * L + 1;
  const int64_t* tg = targets + tgt_offsets[b];
  const bool on = s < S && S <= S_max;  // (an utterance longer than the scratch pitch is reported as infeasible)
  const int sym = on ? ((s & 1) ? (int)tg[s >> 1] : blank) : blank;
  const bool skip = on && (s & 1) && s >= 2 && (int)tg[(s >> 1) - 1] != sym;       // alpha: from s-2
  const bool skip_f = on && (s & 1) && s + 2 < S && (int)tg[(s >> 1) + 1] != sym;  // beta: from s+2
  const float* xb = logits + (int64_t)b * ld_b;
  float* gb = grad != nullptr ? grad + (int64_t)b * ld_b : nullptr;
  float* aw = alpha_ws + (int64_t)b * T * S_max;
  const bool feasible_len = Tn >= 1 && S <= S_max;

  // ---------------- forward sweep
  float a = -CUDART_INF_F;
  if (feasible_len && on && s < 2) a = xb[sym] - lse[b];  // t = 0: only the first blank and the first label
  if (on && feasible_len) aw[s] = a;
  if (s < S_max) prev[s] = a;
  __syncthreads();
  for (int t = 1; t < Tn; ++t) {
    float v = -CUDART_INF_F;
    if (on) {
      v = prev[s];
      if (s >= 1) v = lse2(v, prev[s - 1]);
      if (skip) v = lse2(v, prev[s - 2]);
      v += xb[(int64_t)t * ld_t + sym] - lse[t * B + b];
      aw[(int64_t)t * S_max + s] = v;
    }
    if (s < S_max) cur[s] = v;
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }
  if (s == 0) {
    float ll = -CUDART_INF_F;
    if (feasible_len) ll = S >= 2 ? lse2(prev[S - 1], prev[S - 2]) : prev[0];
    nll_s = -ll;
  }
  __syncthreads();
  const float nll = nll_s;
  const bool finite = nll < CUDART_INF_F;
  if (s == 0) nll_out[b] = finite ? nll : (zero_infinity ? 0.f : CUDART_INF_F);
  if (gb == nullptr) return;

  // ---------------- rows past the utterance (and every row of an infeasible one): zero gradient
  for (int t = finite ? Tn : 0; t < T; ++t)
    for (int k = s; k < V; k += blockDim.x) gb[(int64_t)t * ld_t + k] = 0.f;
  if (!finite) return;

  // ---------------- backward sweep with the gradient rows finished on the fly
  __syncthreads();
  float be = -CUDART_INF_F;
  for (int t = Tn - 1; t >= 0; --t) {
    const float lse_t = lse[t * B + b];
    const float* xt = xb + (int64_t)t * ld_t;
    float v = -CUDART_INF_F;
    if (on) {
      if (t == Tn - 1) {
        v = (s >= S - 2) ? 0.f : -CUDART_INF_F;  // final states: last blank and last label
      } else {
        v = prev[s];
        if (s + 1 < S) v = lse2(v, prev[s + 1]);
        if (skip_f) v = lse2(v, prev[s + 2]);
      }
      v += xt[sym] - lse_t;
    }
    be = v;
    for (int k = s; k < V; k += blockDim.x) acc[k] = 0.f;
    __syncthreads();
    if (on) {
      const float e = aw[(int64_t)t * S_max + s] + be - (xt[sym] - lse_t) + nll;
      if (e > -80.f) atomicAdd(&acc[sym], __expf(e));
    }
    if (s < S_max) cur[s] = be;
    __syncthreads();
    for (int k = s; k < V; k += blockDim.x) gb[(int64_t)t * ld_t + k] = __expf(xt[k] - lse_t) - acc[k];
    float* tmp = prev; prev = cur; cur = tmp;
    __syncthreads();
  }
}

int64_t ctc_ws_floats(int32_t T, int32_t B, int32_t S_max) { return (int64_t)T * B + (int64_t)B * T * S_max; }

int ctc_loss_launch(const float* logits, int64_t ld_t, int64_t ld_b, const int64_t* targets, const int64_t* tgt_offsets,
                    const int64_t* input_lengths, const int64_t* target_lengths, float* nll, float* grad, float* ws,
                    int32_t T, int32_t B, int32_t V, int32_t S_max, int32_t blank, int32_t zero_infinity,
                    cudaStream_t st) {
  if (T <= 0 || B <= 0 || V <= 0 || S_max <= 0 || S_max > 1024 || blank < 0 || blank >= V) return -2;
  float* lse = ws;
  float* alpha = ws + (int64_t)T * B;
  const int rows = T * B;
  ctc_lse_kernel<<<(rows + 7) / 8, 256, 0, st>>>(logits, ld_t, ld_b, lse, T, B, V);
  int threads = (S_max + 31) / 32 * 32;
  if (threads < 64) threads = 64;
  const size_t smem = (size_t)(2 * S_max + V) * sizeof(float);
  if (smem > 48 * 1024) return -5;
  ctc_alpha_beta_kernel<<<B, threads, smem, st>>>(logits, ld_t, ld_b, lse, targets, tgt_offsets, input_lengths,
                                                  target_lengths, nll, grad, alpha, T, B, V, S_max, blank, zero_infinity);
  return (int)cudaGetLastError();
}

}  // namespace st5
