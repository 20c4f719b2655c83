"""Timing sweep of st5_gemm_bf16 (warm, back-to-back, CUDA events): separates fixed per-launch cost from per-K cost and
epilogue cost. Usage: ST5_GEMM_BN=<64|128|256> python tools/probe_gemm_timing.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speecht5_b200 import kernels as K  # noqa: E402

dev = "cuda"


def t(M, N, Kd, out_dtype=torch.bfloat16, iters=30, a_mn=False, b_mn=False, **epi):
    A = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).to(torch.bfloat16)
    B = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).to(torch.bfloat16)
    out = torch.zeros(M, N, device=dev, dtype=out_dtype)
    kw = {}
    if epi.get("bias"):
        kw["bias"] = torch.randn(N, device=dev)
    if epi.get("act"):
        kw["act"] = epi["act"]; kw["c_pre"] = torch.empty_like(out)
    if epi.get("drop"):
        kw.update(drop_p=0.1, seed=1, offset=3)
    if epi.get("acc"):
        kw["accumulate"] = True
    for _ in range(3):
        K.gemm(A, B, out, M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn, **kw)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()  # launches replayed from a graph: device time only, no Python/ctypes launch cost
    with torch.cuda.graph(g):
        for _ in range(iters):
            K.gemm(A, B, out, M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn, **kw)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(json.dumps(dict(M=M, N=N, K=Kd, f32=out_dtype == torch.float32, mn=f"{int(a_mn)}{int(b_mn)}", epi=epi,
                          us=round(us, 1), tflops=round(2.0 * M * N * Kd / us / 1e6, 1))), flush=True)


print("BN override:", os.environ.get("ST5_GEMM_BN"))
for Kd in (64, 256, 768, 1536, 3072):
    t(10016, 768, Kd)
t(10016, 768, 768, bias=True)
t(10016, 768, 768, out_dtype=torch.float32)
t(10016, 2304, 768, bias=True)
t(5120, 3072, 768, bias=True)
t(5120, 3072, 768, bias=True, act="gelu")
t(5120, 3072, 768, bias=True, act="gelu", drop=True)
t(768, 768, 10016, a_mn=True, b_mn=True, out_dtype=torch.float32, acc=True)
t(768, 768, 10016, a_mn=True, b_mn=True, out_dtype=torch.float32)
