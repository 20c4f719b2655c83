"""Where the tcgen05 GEMM stands against the library on the step's own shapes: every distinct (M, N, K, layout) of
profiles/r02_gemm_shapes_v1.json timed back to back under one CUDA graph, st5_gemm_bf16 (plain epilogue, bf16 or fp32
output as in the step) next to torch.matmul (cuBLASLt bf16, fp32 accumulate). The library number is a reference point,
not a product path."""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speecht5_b200 import kernels as K  # noqa: E402

dev = "cuda"
shapes = json.load(open(os.path.join(ROOT, "profiles", "r02_gemm_shapes_v1.json")))
cnt = collections.Counter((s["M"], s["N"], s["K"], s["nb"], s["a_mn"], s["b_mn"], s["c_fp32"]) for s in shapes)
SIDE = torch.cuda.Stream()


def timed(fn, reps=10):
    SIDE.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(SIDE):
        for _ in range(2):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=SIDE):
        for _ in range(reps):
            fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps * 1e3)
    return best


rows = []
tot_ours = tot_lib = 0.0
for (M, N, Kd, nb, a_mn, b_mn, f32), n in sorted(cnt.items(), key=lambda kv: -kv[1] * kv[0][0] * kv[0][1] * kv[0][2] * kv[0][3]):
    if nb != 1 or M * N * Kd < 1e8 or N % 8:
        continue
    A = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).to(torch.bfloat16)
    B = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    t_ours = timed(lambda: K.gemm(A, B, out, M=M, N=N, K=Kd, a_mn=bool(a_mn), b_mn=bool(b_mn)))
    Am = A.t() if a_mn else A          # [M, K] view
    Bm = B if b_mn else B.t()          # [K, N] view
    lib_out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    t_lib = timed(lambda: torch.matmul(Am, Bm, out=lib_out))
    fl = 2.0 * M * N * Kd
    rows.append(dict(M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn, f32=f32, per_step=n, ours_us=round(t_ours, 1), lib_us=round(t_lib, 1),
                     ours_tf=round(fl / t_ours * 1e-6), lib_tf=round(fl / t_lib * 1e-6)))
    tot_ours += n * t_ours
    tot_lib += n * t_lib
    print(json.dumps(rows[-1]), flush=True)
print(json.dumps(dict(step_ms_ours=round(tot_ours * 1e-3, 3), step_ms_library=round(tot_lib * 1e-3, 3),
                      note="plain epilogues only: the step's fused bias / activation / dropout / residual / accumulate work is "
                           "extra for st5_gemm_bf16 in the real step and would be separate kernels for the library")))
