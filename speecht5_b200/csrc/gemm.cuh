// Parameters shared between the host launcher (gemm.cu) and callers inside the library.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace st5 {

enum Act : int {
  ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_TANH = 3, ACT_GELU_TANH = 4,
  ACT_GATE = 5,            // ag_act only: ag_pre already holds the backward multiplier (see ACT_GELU_TANH_GATE)
  ACT_GELU_TANH_GATE = 6   // act only: tanh-form GELU + dropout, and C_pre receives keep * scale * GELU'(pre) instead of pre
};

// D[z][m][n] = epilogue( alpha * sum_k A[z][m][k] * B[z][n][k] )
// Operands are bf16. "K-major" = k is the contiguous index (row-major [rows][K]); "MN-major" = the m (or n)
// index is contiguous (i.e. the operand is stored transposed, [K][rows]).
struct GemmDesc {
  int M, N, K;
  int nb1, nb2;  // batch = nb1 * nb2, z = b2 * nb1 + b1
  // A operand
  const void* A; int a_mn; long a_ld; long a_bs1, a_bs2;  // ld = stride (elements) of the non-contiguous index
  const void* B; int b_mn; long b_ld; long b_bs1, b_bs2;
  // output (row-major [M][N] per batch)
  void* C; int c_fp32; long c_ld; long c_bs1, c_bs2;
  void* C_pre;            // optional: value before activation/dropout (same layout & dtype as C)
  const float* bias;      // [N] or null
  const float* bias2;     // optional row-group bias [M / bias2_rows][N]
  int bias2_rows;
  const void* residual;   // optional, same layout & dtype as C, added after activation/dropout
  int act;
  float alpha;
  int accumulate;         // 1: C += result, read-modify-write in the epilogue; 2: C += result as a TMA reduce at the L2
                          //    (batch entries may then share one output: split-K). C must be fp32
  float drop_p; uint64_t drop_seed, drop_offset;
  const void* ag_pre; int ag_act;  // optional: out *= act'(ag_pre[m][n]) after the dropout mask (same layout as C)
};

int gemm_launch(const GemmDesc& g, cudaStream_t stream);

}  // namespace st5
