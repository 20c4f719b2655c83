"""GPU probe: correctness of st5_gemm_bf16 across operand layouts / epilogues, then timing. Run under gpurun."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speecht5_b200 import kernels as K  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
results = []


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def run_case(name, M, N, Kd, a_mn=False, b_mn=False, out_dtype=torch.bfloat16, nb1=1, nb2=1, **epi):
    nb = nb1 * nb2
    A = torch.randn(nb, M, Kd, device=dev).to(torch.bfloat16)
    B = torch.randn(nb, N, Kd, device=dev).to(torch.bfloat16)
    ref = torch.matmul(A.float(), B.float().transpose(1, 2))
    a_st = A.transpose(1, 2).contiguous() if a_mn else A
    b_st = B.transpose(1, 2).contiguous() if b_mn else B
    a_ld = M if a_mn else Kd
    b_ld = N if b_mn else Kd
    out = torch.full((nb, M, N), float("nan"), device=dev, dtype=out_dtype)
    kw = {}
    bias = None
    if epi.get("bias"):
        bias = torch.randn(N, device=dev)
        ref = ref + bias
        kw["bias"] = bias
    if epi.get("act"):
        kw["act"] = epi["act"]
        pre = ref.clone()
        if epi["act"] == "gelu":
            ref = torch.nn.functional.gelu(ref)
        elif epi["act"] == "relu":
            ref = torch.relu(ref)
        elif epi["act"] == "tanh":
            ref = torch.tanh(ref)
        if epi.get("c_pre"):
            kw["c_pre"] = torch.empty_like(out)
    if epi.get("residual"):
        res = torch.randn(nb, M, N, device=dev).to(out_dtype)
        ref = ref + res.float()
        kw["residual"] = res
    if epi.get("alpha"):
        kw["alpha"] = epi["alpha"]
        ref = None  # handled below
    try:
        K.gemm(a_st, b_st, out, M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn, a_ld=a_ld, b_ld=b_ld, c_ld=N, nb1=nb1, nb2=nb2,
               a_bs=(a_st.stride(0), a_st.stride(0) * nb1), b_bs=(b_st.stride(0), b_st.stride(0) * nb1),
               c_bs=(M * N, M * N * nb1), **kw)
        torch.cuda.synchronize()
        err = rel(out, ref)
        extra = {}
        if "c_pre" in kw:
            extra["pre_err"] = rel(kw["c_pre"], pre)
        ok = err < (2e-2 if out_dtype == torch.bfloat16 else 1e-3) and not torch.isnan(out.float()).any().item()
        results.append(dict(case=name, ok=bool(ok), err=err, **extra))
    except Exception as e:  # noqa: BLE001
        results.append(dict(case=name, ok=False, error=str(e)[:300]))
    print(json.dumps(results[-1]), flush=True)


run_case("kk_128x128x64_f32", 128, 128, 64, out_dtype=torch.float32)
run_case("kk_256x256x128_f32", 256, 256, 128, out_dtype=torch.float32)
run_case("kk_256x256x128_bf16", 256, 256, 128)
run_case("kk_300x162x400_f32", 300, 162, 400, out_dtype=torch.float32)
run_case("kk_5120x768x768_bf16", 5120, 768, 768)
run_case("kk_5120x2304x768_bf16_bn256", 5120, 2304, 768)
run_case("kk_1000x64x313_f32", 1000, 64, 320, out_dtype=torch.float32)
run_case("mnA_256x256x128_f32", 256, 256, 128, a_mn=True, out_dtype=torch.float32)
run_case("mnB_256x256x128_f32", 256, 256, 128, b_mn=True, out_dtype=torch.float32)
run_case("mnAB_256x256x128_f32", 256, 256, 128, a_mn=True, b_mn=True, out_dtype=torch.float32)
run_case("mnAB_768x3072x5120_f32", 768, 3072, 5120, a_mn=True, b_mn=True, out_dtype=torch.float32)
run_case("mnB_5120x768x3072_bf16", 5120, 768, 3072, b_mn=True)
run_case("mnB_320x64x160_f32_batched", 320, 64, 160, b_mn=True, out_dtype=torch.float32, nb1=3, nb2=2)
run_case("kk_batched_160x160x64", 160, 160, 64, out_dtype=torch.float32, nb1=4, nb2=3)
run_case("bias_gelu_pre", 512, 3072, 768, bias=True, act="gelu", c_pre=True)
run_case("bias_relu_res_f32", 384, 256, 80, bias=True, act="relu", residual=True, out_dtype=torch.float32)
run_case("bias_res_bf16", 384, 768, 768, bias=True, residual=True)

# accumulate (split-precision triple product) and dropout statistics
try:
    M, N, Kd = 256, 384, 512
    x = torch.randn(M, Kd, device=dev)
    w = torch.randn(N, Kd, device=dev)
    xh = torch.empty(M, Kd, device=dev, dtype=torch.bfloat16); xl = torch.empty_like(xh)
    wh = torch.empty(N, Kd, device=dev, dtype=torch.bfloat16); wl = torch.empty_like(wh)
    K.cast_bf16(x, xh, xl); K.cast_bf16(w, wh, wl)
    out = torch.empty(M, N, device=dev)
    K.gemm(xh, wh, out, M=M, N=N, K=Kd)
    K.gemm(xh, wl, out, M=M, N=N, K=Kd, accumulate=True)
    K.gemm(xl, wh, out, M=M, N=N, K=Kd, accumulate=True)
    ref = (x.double() @ w.double().t()).float()
    e3 = rel(out, ref)
    out1 = torch.empty(M, N, device=dev)
    K.gemm(xh, wh, out1, M=M, N=N, K=Kd)
    e1 = rel(out1, ref)
    results.append(dict(case="split3_accumulate", ok=bool(e3 < 1e-4), err_bf16x3=e3, err_bf16=e1,
                        err_torch_fp32=rel(x @ w.t(), ref)))
    print(json.dumps(results[-1]), flush=True)
    a = torch.ones(1024, 64, device=dev, dtype=torch.bfloat16)
    b = torch.ones(1024, 64, device=dev, dtype=torch.bfloat16)
    o = torch.empty(1024, 1024, device=dev)
    K.gemm(a, b, o, M=1024, N=1024, K=64, drop_p=0.25, seed=1234, offset=7)
    keep = (o != 0).float().mean().item()
    val = o.max().item()
    o2 = torch.empty_like(o)
    K.gemm(a, b, o2, M=1024, N=1024, K=64, drop_p=0.25, seed=1234, offset=7)
    ones = torch.ones(1024 * 1024, device=dev)
    d = torch.empty_like(ones)
    K.dropout(ones, d, 0.25, 1234, 7)
    same_mask = bool(((d.view(1024, 1024) != 0) == (o != 0)).all().item())
    results.append(dict(case="dropout", ok=bool(abs(keep - 0.75) < 0.01 and abs(val - 64 / 0.75) < 0.5
                                                 and torch.equal(o, o2) and same_mask),
                        keep=keep, val=val, same_mask=same_mask))
    print(json.dumps(results[-1]), flush=True)
except Exception as e:  # noqa: BLE001
    results.append(dict(case="split3/dropout", ok=False, error=str(e)[:300]))
    print(json.dumps(results[-1]), flush=True)


def bench(name, M, N, Kd, a_mn=False, b_mn=False, out_dtype=torch.bfloat16, iters=50):
    A = torch.randn((Kd, M) if a_mn else (M, Kd), device=dev).to(torch.bfloat16)
    B = torch.randn((Kd, N) if b_mn else (N, Kd), device=dev).to(torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=out_dtype)
    for _ in range(5):
        K.gemm(A, B, out, M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        K.gemm(A, B, out, M=M, N=N, K=Kd, a_mn=a_mn, b_mn=b_mn)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    Af = A.t() if a_mn else A
    Bf = B.t() if b_mn else B
    for _ in range(5):
        torch.matmul(Af, Bf.t())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(Af, Bf.t())
    e1.record()
    torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * Kd / ms / 1e9
    r = dict(bench=name, ms=ms, tflops=tf, cublas_ms=ms_t, cublas_tflops=2.0 * M * N * Kd / ms_t / 1e9)
    results.append(r)
    print(json.dumps(r), flush=True)


try:
    bench("qkv_5120x2304x768", 5120, 2304, 768)
    bench("fc1_5120x3072x768", 5120, 3072, 768)
    bench("fc2_5120x768x3072", 5120, 768, 3072)
    bench("dec_fc1_10016x3072x768", 10016, 3072, 768)
    bench("dx_5120x768x3072_mnB", 5120, 768, 3072, b_mn=True)
    bench("dw_3072x768x5120_mnAB", 3072, 768, 5120, a_mn=True, b_mn=True, out_dtype=torch.float32)
    bench("big_8192^3", 8192, 8192, 8192, iters=10)
except Exception as e:  # noqa: BLE001
    print(json.dumps(dict(bench="failed", error=str(e)[:300])), flush=True)

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/probe_gemm.json", "w") as fh:
    json.dump(results, fh, indent=1)
nfail = sum(1 for r in results if r.get("ok") is False)
print("FAILED" if nfail else "ALL OK", nfail)
