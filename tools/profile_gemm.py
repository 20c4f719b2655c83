"""A handful of representative st5_gemm_bf16 launches of the TTS step, for `ncu --set full -k regex:gemm_bf16`.
Each shape is warmed once, then launched once per variant (the profiled instances are the 2nd launch of each)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speecht5_b200 import kernels as K  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def run(M, N, Kd, a_mn=False, b_mn=False, out_dtype=torch.bfloat16, **epi):
    A = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).to(torch.bfloat16)
    B = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).to(torch.bfloat16)
    out = torch.zeros(M, N, device=dev, dtype=out_dtype)
    kw = {}
    if epi.get("bias"):
        kw["bias"] = torch.randn(N, device=dev)
    if epi.get("act"):
        kw["act"] = epi["act"]
        kw["c_pre"] = torch.empty_like(out)
    if epi.get("res"):
        kw["residual"] = torch.randn(M, N, device=dev).to(out_dtype)
    if epi.get("drop"):
        kw.update(drop_p=0.1, seed=1, offset=3)
    if epi.get("acc"):
        kw["accumulate"] = True
    if epi.get("ag"):
        kw["actgrad_pre"] = torch.randn(M, N, device=dev).to(out_dtype)
        kw["actgrad_act"] = epi["ag"]
    for _ in range(2):
        K.gemm(A, B, out, M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn, **kw)
    torch.cuda.synchronize()


run(10016, 768, 768, bias=True)                                   # decoder out_proj forward (plain bias epilogue)
run(5120, 3072, 768, bias=True, act="gelu_tanh_gate", drop=True)  # encoder fc1: GELU + backward gate store + dropout
run(5120, 768, 3072, bias=True, res=True, drop=True)              # encoder fc2 (+ dropout + residual)
run(5120, 3072, 768, b_mn=True, ag="gate")                        # dH = (dY . W2) * gate
run(5120, 768, 3072, b_mn=True)                                   # dX of fc1
run(768, 768, 10016, a_mn=True, b_mn=True, out_dtype=torch.float32, acc=True)   # dW of a 768x768 projection
run(3072, 768, 5120, a_mn=True, b_mn=True, out_dtype=torch.float32, acc=True)   # dW of fc1
run(2304, 768, 5120, a_mn=True, b_mn=True, out_dtype=torch.float32, acc=True)   # dW of q|k|v
run(8192, 8192, 8192)                                              # large square reference point
print("done")
