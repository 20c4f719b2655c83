// LayerNorm (fused residual + dropout) and BatchNorm1d (fused tanh + dropout) forward/backward.
// Reference semantics: fairseq LayerNorm == torch.nn.LayerNorm (fairseq/modules/layer_norm.py:30-35), post-LN residual
// blocks of transformer_layer.py:112-132 / :343-391; espnet Tacotron2 Postnet BatchNorm1d blocks
// (speech_decoder_postnet.py:39-51) with training statistics over every row, padded frames included.
#include "kernels.cuh"
#include "ptx.cuh"
#include "gemm.cuh"

namespace st5 {

constexpr int LN_MAX_PER_LANE = 32;  // C <= 1024
constexpr int LN_WARPS = 4;

template <typename T>
__global__ void ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ residual, const float* __restrict__ gamma,
                              const float* __restrict__ beta, T* __restrict__ y, T* __restrict__ s_out,
                              float* __restrict__ mean, float* __restrict__ rstd, int64_t rows, int C, float eps,
                              uint32_t thr, float dscale, uint64_t seed, uint64_t offset) {
  resolve_seed(seed, offset);
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int per = (C + 31) / 32;
  float v[LN_MAX_PER_LANE];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
    if (k < per) {
      const int c = k * 32 + lane;
      float t = 0.f;
      if (c < C) {
        const int64_t i = row * C + c;
        t = ldf(x + i);
        if (thr != 0) t = dropout_keep(seed, offset, (uint64_t)i, thr) ? t * dscale : 0.f;
        if (residual != nullptr) t += ldf(residual + i);
        if (s_out != nullptr) stf(s_out + i, t);
        if (s_out != nullptr) t = ldf(s_out + i);  // normalise exactly what backward will read
        sum += t;
      }
      v[k] = t;
    }
  }
  sum = warp_sum(sum);
  const float mu = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
    if (k < per) {
      const int c = k * 32 + lane;
      if (c < C) {
        const float d = v[k] - mu;
        sq += d * d;
      }
    }
  }
  sq = warp_sum(sq);
  const float rs = rsqrtf(sq / (float)C + eps);
  if (lane == 0) {
    if (mean != nullptr) mean[row] = mu;
    if (rstd != nullptr) rstd[row] = rs;
  }
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
    if (k < per) {
      const int c = k * 32 + lane;
      if (c < C) stf(y + row * C + c, (v[k] - mu) * rs * gamma[c] + beta[c]);
    }
  }
}

int ln_fwd_launch(const void* x, const void* residual, const float* gamma, const float* beta, void* y, void* s_out,
                  float* mean, float* rstd, int dtype, int64_t rows, int64_t C, float eps, float drop_p, uint64_t seed,
                  uint64_t offset, cudaStream_t s) {
  if (rows == 0) return 0;
  if (C > 32 * LN_MAX_PER_LANE || C <= 0) return -2;
  const uint32_t thr = drop_threshold(drop_p);
  const float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const unsigned grid = (unsigned)((rows + LN_WARPS - 1) / LN_WARPS);
  if (dtype == ST5_F32)
    ln_fwd_kernel<float><<<grid, LN_WARPS * 32, 0, s>>>((const float*)x, (const float*)residual, gamma, beta, (float*)y,
                                                        (float*)s_out, mean, rstd, rows, (int)C, eps, thr, ds, seed,
                                                        offset);
  else
    ln_fwd_kernel<__nv_bfloat16><<<grid, LN_WARPS * 32, 0, s>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, gamma, beta, (__nv_bfloat16*)y, (__nv_bfloat16*)s_out,
        mean, rstd, rows, (int)C, eps, thr, ds, seed, offset);
  return (int)cudaGetLastError();
}

int64_t ln_bwd_blocks(int64_t rows) {
  int64_t b = (rows + LN_WARPS - 1) / LN_WARPS;
  if (b > 296) b = 296;
  if (b < 1) b = 1;
  return b;
}

template <typename T>
__global__ void ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ s_in, const float* __restrict__ mean,
                              const float* __restrict__ rstd, const float* __restrict__ gamma, T* __restrict__ ds,
                              T* __restrict__ dx, float* __restrict__ partials, int64_t rows, int C, uint32_t thr,
                              float dscale, uint64_t seed, uint64_t offset) {
  resolve_seed(seed, offset);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = (C + 31) / 32;
  float dg[LN_MAX_PER_LANE], db[LN_MAX_PER_LANE];
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) dg[k] = db[k] = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * LN_WARPS + warp; row < rows; row += (int64_t)gridDim.x * LN_WARPS) {
    const float mu = mean[row], rs = rstd[row];
    float g[LN_MAX_PER_LANE], xh[LN_MAX_PER_LANE];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
      if (k < per) {
        const int c = k * 32 + lane;
        float gg = 0.f, xx = 0.f;
        if (c < C) {
          const float d = ldf(dy + row * C + c);
          xx = (ldf(s_in + row * C + c) - mu) * rs;
          gg = d * gamma[c];
          dg[k] += d * xx;
          db[k] += d;
        }
        g[k] = gg; xh[k] = xx;
        c1 += gg; c2 += gg * xx;
      }
    }
    c1 = warp_sum(c1) / (float)C;
    c2 = warp_sum(c2) / (float)C;
#pragma unroll
    for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
      if (k < per) {
        const int c = k * 32 + lane;
        if (c < C) {
          const int64_t i = row * C + c;
          const float r = rs * (g[k] - c1 - xh[k] * c2);
          if (ds != nullptr) stf(ds + i, r);
          if (dx != nullptr) {
            float rr = r;
            if (thr != 0) rr = dropout_keep(seed, offset, (uint64_t)i, thr) ? r * dscale : 0.f;
            stf(dx + i, rr);
          }
        }
      }
    }
  }
  // reduce the per-warp column partials across the block's warps, then one partial row per block
  extern __shared__ float sm[];  // [LN_WARPS][2][C]
#pragma unroll
  for (int k = 0; k < LN_MAX_PER_LANE; ++k) {
    if (k < per) {
      const int c = k * 32 + lane;
      if (c < C) {
        sm[(warp * 2 + 0) * C + c] = dg[k];
        sm[(warp * 2 + 1) * C + c] = db[k];
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) {
      a += sm[(w * 2 + 0) * C + c];
      b += sm[(w * 2 + 1) * C + c];
    }
    partials[((int64_t)blockIdx.x * 2 + 0) * C + c] = a;
    partials[((int64_t)blockIdx.x * 2 + 1) * C + c] = b;
  }
}
__global__ void ln_bwd_finalize(const float* __restrict__ partials, float* __restrict__ dgamma,
                                float* __restrict__ dbeta, int nblk, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int k = 0; k < nblk; ++k) {
    a += partials[((int64_t)k * 2 + 0) * C + c];
    b += partials[((int64_t)k * 2 + 1) * C + c];
  }
  if (dgamma != nullptr) dgamma[c] += a;
  if (dbeta != nullptr) dbeta[c] += b;
}

int ln_bwd_launch(const void* dy, const void* s_in, const float* mean, const float* rstd, const float* gamma, void* ds,
                  void* dx, float* dgamma, float* dbeta, float* partials, int dtype, int64_t rows, int64_t C,
                  float drop_p, uint64_t seed, uint64_t offset, cudaStream_t s) {
  if (rows == 0) return 0;
  if (C > 32 * LN_MAX_PER_LANE || C <= 0) return -2;
  const uint32_t thr = drop_threshold(drop_p);
  const float dsc = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int nblk = (int)ln_bwd_blocks(rows);
  const size_t smem = sizeof(float) * LN_WARPS * 2 * C;
  if (dtype == ST5_F32)
    ln_bwd_kernel<float><<<nblk, LN_WARPS * 32, smem, s>>>((const float*)dy, (const float*)s_in, mean, rstd, gamma,
                                                           (float*)ds, (float*)dx, partials, rows, (int)C, thr, dsc,
                                                           seed, offset);
  else
    ln_bwd_kernel<__nv_bfloat16><<<nblk, LN_WARPS * 32, smem, s>>>(
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)s_in, mean, rstd, gamma, (__nv_bfloat16*)ds,
        (__nv_bfloat16*)dx, partials, rows, (int)C, thr, dsc, seed, offset);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ln_bwd_finalize<<<(unsigned)((C + 127) / 128), 128, 0, s>>>(partials, dgamma, dbeta, nblk, (int)C);
  return (int)cudaGetLastError();
}

// =============================================================================================== BatchNorm1d
// channels-last rows [rows][C]; block (32 channels, 8 row lanes); per-channel partial sums via fp32 atomics.
template <typename T, int MODE>  // MODE 0: sum(x) ; 1: sum((x-mean)^2) with mean = aux[c]
__global__ void bn_stat_kernel(const T* __restrict__ x, int64_t ld, const float* __restrict__ aux,
                               float* __restrict__ out, int64_t rows, int C, int64_t rows_per_block) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc = 0.f;
  if (c < C) {
    const float mu = MODE == 1 ? aux[c] : 0.f;
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float v = ldf(x + r * ld + c) - mu;
      acc += MODE == 1 ? v * v : v;
    }
  }
  __shared__ float red[8][33];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    atomicAdd(out + c, v);
  }
}
__global__ void bn_finalize_stats(float* __restrict__ scratch, float* __restrict__ running_mean,
                                  float* __restrict__ running_var, float* __restrict__ save_mean,
                                  float* __restrict__ save_rstd, int64_t rows, int C, float momentum, float eps,
                                  int stage) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (stage == 0) {
    scratch[c] = scratch[c] / (float)rows;  // mean
  } else {
    const float mu = scratch[c];
    const float var = scratch[C + c] / (float)rows;
    save_mean[c] = mu;
    save_rstd[c] = rsqrtf(var + eps);
    if (running_mean != nullptr) {
      const float unbiased = rows > 1 ? scratch[C + c] / (float)(rows - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  }
}
__global__ void bn_eval_stats(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                              float* __restrict__ save_mean, float* __restrict__ save_rstd, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  save_mean[c] = running_mean[c];
  save_rstd[c] = rsqrtf(running_var[c] + eps);
}
template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ x, int64_t x_ld, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ mean,
                                const float* __restrict__ rstd, T* __restrict__ y, int64_t y_ld, T* __restrict__ y_pre,
                                int64_t rows, int C, int act, uint32_t thr, float dscale, uint64_t seed,
                                uint64_t offset) {
  resolve_seed(seed, offset);
  const int64_t n = rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    float v = (ldf(x + r * x_ld + c) - mean[c]) * rstd[c] * gamma[c] + beta[c];
    if (y_pre != nullptr) { stf(y_pre + i, v); v = ldf(y_pre + i); }
    if (act == ACT_TANH) v = tanhf(v);
    else if (act == ACT_RELU) v = fmaxf(v, 0.f);
    if (thr != 0) v = dropout_keep(seed, offset, (uint64_t)i, thr) ? v * dscale : 0.f;
    stf(y + r * y_ld + c, v);
  }
}

static inline int bn_row_blocks(int64_t rows, int64_t* rpb) {
  int64_t nb = (rows + 255) / 256;
  if (nb > 128) nb = 128;
  if (nb < 1) nb = 1;
  *rpb = (rows + nb - 1) / nb;
  return (int)nb;
}

int bn_fwd_launch(const void* x, int64_t x_ld, const float* gamma, const float* beta, float* running_mean,
                  float* running_var, float* save_mean, float* save_rstd, void* y, int64_t y_ld, void* y_pre, int dtype,
                  int64_t rows, int64_t C, int training, float momentum, float eps, int act, float drop_p,
                  uint64_t seed, uint64_t offset, float* scratch, cudaStream_t s) {
  if (rows == 0 || C == 0) return 0;
  const unsigned cblk = (unsigned)((C + 127) / 128);
  if (training) {
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * C, s);
    if (e != cudaSuccess) return (int)e;
    int64_t rpb;
    const int nb = bn_row_blocks(rows, &rpb);
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)nb), block(32, 8);
    if (dtype == ST5_F32) bn_stat_kernel<float, 0><<<grid, block, 0, s>>>((const float*)x, x_ld, nullptr, scratch, rows, (int)C, rpb);
    else bn_stat_kernel<__nv_bfloat16, 0><<<grid, block, 0, s>>>((const __nv_bfloat16*)x, x_ld, nullptr, scratch, rows, (int)C, rpb);
    bn_finalize_stats<<<cblk, 128, 0, s>>>(scratch, nullptr, nullptr, nullptr, nullptr, rows, (int)C, momentum, eps, 0);
    if (dtype == ST5_F32) bn_stat_kernel<float, 1><<<grid, block, 0, s>>>((const float*)x, x_ld, scratch, scratch + C, rows, (int)C, rpb);
    else bn_stat_kernel<__nv_bfloat16, 1><<<grid, block, 0, s>>>((const __nv_bfloat16*)x, x_ld, scratch, scratch + C, rows, (int)C, rpb);
    bn_finalize_stats<<<cblk, 128, 0, s>>>(scratch, running_mean, running_var, save_mean, save_rstd, rows, (int)C, momentum, eps, 1);
  } else {
    bn_eval_stats<<<cblk, 128, 0, s>>>(running_mean, running_var, save_mean, save_rstd, (int)C, eps);
  }
  const uint32_t thr = drop_threshold(drop_p);
  const float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int64_t n = rows * C;
  int64_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (dtype == ST5_F32)
    bn_apply_kernel<float><<<(unsigned)g, 256, 0, s>>>((const float*)x, x_ld, gamma, beta, save_mean, save_rstd,
                                                       (float*)y, y_ld, (float*)y_pre, rows, (int)C, act, thr, ds, seed,
                                                       offset);
  else
    bn_apply_kernel<__nv_bfloat16><<<(unsigned)g, 256, 0, s>>>((const __nv_bfloat16*)x, x_ld, gamma, beta, save_mean,
                                                               save_rstd, (__nv_bfloat16*)y, y_ld,
                                                               (__nv_bfloat16*)y_pre, rows, (int)C, act, thr, ds, seed,
                                                               offset);
  return (int)cudaGetLastError();
}

// backward: g = dropout_bwd(dy) * act'(y_pre); pass 1: sum(g), sum(g * xhat) per channel; pass 2: dx.
template <typename T>
__device__ __forceinline__ float bn_g(const T* dy, int64_t dy_ld, const T* y_pre, int64_t r, int c, int C, int act,
                                      uint32_t thr, float dscale, uint64_t seed, uint64_t offset) {
  const int64_t i = r * C + c;
  float g = ldf(dy + r * dy_ld + c);
  if (thr != 0) g = dropout_keep(seed, offset, (uint64_t)i, thr) ? g * dscale : 0.f;
  if (act == ACT_TANH) {
    const float t = tanhf(ldf(y_pre + i));
    g *= 1.f - t * t;
  } else if (act == ACT_RELU) {
    g = ldf(y_pre + i) > 0.f ? g : 0.f;
  }
  return g;
}
template <typename T>
__global__ void bn_bwd_stat_kernel(const T* __restrict__ dy, int64_t dy_ld, const T* __restrict__ x, int64_t x_ld,
                                   const T* __restrict__ y_pre, const float* __restrict__ mean,
                                   const float* __restrict__ rstd, float* __restrict__ scratch, int64_t rows, int C,
                                   int64_t rows_per_block, int act, uint32_t thr, float dscale, uint64_t seed,
                                   uint64_t offset) {
  resolve_seed(seed, offset);
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float a = 0.f, b = 0.f;
  if (c < C) {
    const float mu = mean[c], rs = rstd[c];
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const float g = bn_g(dy, dy_ld, y_pre, r, c, C, act, thr, dscale, seed, offset);
      a += g;
      b += g * (ldf(x + r * x_ld + c) - mu) * rs;
    }
  }
  __shared__ float red[2][8][33];
  red[0][threadIdx.y][threadIdx.x] = a;
  red[1][threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float va = 0.f, vb = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { va += red[0][k][threadIdx.x]; vb += red[1][k][threadIdx.x]; }
    atomicAdd(scratch + c, va);
    atomicAdd(scratch + C + c, vb);
  }
}
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dy, int64_t dy_ld, const T* __restrict__ x, int64_t x_ld,
                                    const T* __restrict__ y_pre, const float* __restrict__ gamma,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ scratch, T* __restrict__ dx, int64_t dx_ld,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int C, int act,
                                    uint32_t thr, float dscale, uint64_t seed, uint64_t offset) {
  resolve_seed(seed, offset);
  const int64_t n = rows * C;
  const float inv_n = 1.f / (float)rows;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const float g = bn_g(dy, dy_ld, y_pre, r, c, C, act, thr, dscale, seed, offset);
    const float xh = (ldf(x + r * x_ld + c) - mean[c]) * rstd[c];
    stf(dx + r * dx_ld + c, gamma[c] * rstd[c] * (g - scratch[c] * inv_n - xh * scratch[C + c] * inv_n));
  }
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta != nullptr) dbeta[c] += scratch[c];
      if (dgamma != nullptr) dgamma[c] += scratch[C + c];
    }
}

int bn_bwd_launch(const void* dy, int64_t dy_ld, const void* x, int64_t x_ld, const void* y_pre, const float* gamma,
                  const float* save_mean, const float* save_rstd, void* dx, int64_t dx_ld, float* dgamma, float* dbeta,
                  int dtype, int64_t rows, int64_t C, int act, float drop_p, uint64_t seed, uint64_t offset,
                  float* scratch, cudaStream_t s) {
  if (rows == 0 || C == 0) return 0;
  if (act != ACT_NONE && y_pre == nullptr) return -2;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(float) * 2 * C, s);
  if (e != cudaSuccess) return (int)e;
  const uint32_t thr = drop_threshold(drop_p);
  const float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  int64_t rpb;
  const int nb = bn_row_blocks(rows, &rpb);
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)nb), block(32, 8);
  const int64_t n = rows * C;
  int64_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (dtype == ST5_F32) {
    bn_bwd_stat_kernel<float><<<grid, block, 0, s>>>((const float*)dy, dy_ld, (const float*)x, x_ld,
                                                     (const float*)y_pre, save_mean, save_rstd, scratch, rows, (int)C,
                                                     rpb, act, thr, ds, seed, offset);
    bn_bwd_apply_kernel<float><<<(unsigned)g, 256, 0, s>>>((const float*)dy, dy_ld, (const float*)x, x_ld,
                                                           (const float*)y_pre, gamma, save_mean, save_rstd, scratch,
                                                           (float*)dx, dx_ld, dgamma, dbeta, rows, (int)C, act, thr, ds,
                                                           seed, offset);
  } else {
    bn_bwd_stat_kernel<__nv_bfloat16><<<grid, block, 0, s>>>(
        (const __nv_bfloat16*)dy, dy_ld, (const __nv_bfloat16*)x, x_ld, (const __nv_bfloat16*)y_pre, save_mean,
        save_rstd, scratch, rows, (int)C, rpb, act, thr, ds, seed, offset);
    bn_bwd_apply_kernel<__nv_bfloat16><<<(unsigned)g, 256, 0, s>>>(
        (const __nv_bfloat16*)dy, dy_ld, (const __nv_bfloat16*)x, x_ld, (const __nv_bfloat16*)y_pre, gamma, save_mean,
        save_rstd, scratch, (__nv_bfloat16*)dx, dx_ld, dgamma, dbeta, rows, (int)C, act, thr, ds, seed, offset);
  }
  return (int)cudaGetLastError();
}

}  // namespace st5
