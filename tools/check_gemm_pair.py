"""Validation + timing of the experimental CTA-pair GEMM (ST5_GEMM_PAIR=1: 256 x 256 tiles, tcgen05.mma.cta_group::2)
against the default single-CTA kernel. The switch is read once per process, so each mode runs in its own subprocess;
the parent compares the outputs BIT FOR BIT (both kernels accumulate the k blocks of an output element in the same order
in fp32 TMEM) and prints the per-shape times.
usage (GPU box): timeout 300 python tools/check_gemm_pair.py            # exit 0 = identical everywhere
       worker : python tools/check_gemm_pair.py --worker OUT.pt          (mode from the environment)
       quick  : ST5_GEMM_PAIR=1 python tools/check_gemm_pair.py --quick  (one process, fp32 reference only)"""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (M, N, K, a_mn, b_mn, fp32_out, nb, epilogue) -- the step's GEMM classes plus ragged / tiny edges
SHAPES = [
    (5120, 768, 768, 0, 0, 0, 1, dict(bias=True)),
    (10016, 768, 768, 0, 0, 0, 1, dict(bias=True)),                      # M % 256 = 32: peer CTA fully out of range
    (5120, 3072, 768, 0, 0, 0, 1, dict(bias=True, act="gelu_tanh", drop=True)),
    (5120, 768, 3072, 0, 0, 0, 1, dict(bias=True)),
    (5120, 3072, 768, 0, 1, 0, 1, dict(drop=True, ag="gelu_tanh")),
    (10016, 3072, 768, 0, 1, 0, 1, dict()),
    (3072, 768, 5120, 1, 1, 1, 1, dict(acc=True)),
    (768, 3072, 10016, 1, 1, 1, 1, dict(acc=True)),
    (768, 768, 1252, 1, 1, 1, 8, dict()),                                # batched split-K weight gradient
    (300, 1000, 136, 0, 0, 0, 1, dict(bias=True)),                       # ragged in every dimension
    (129, 257, 64, 0, 0, 1, 1, dict()),
    (8192, 8192, 8192, 0, 0, 0, 1, dict()),
]


def worker(out_path, quick=False):
    from speecht5_b200 import kernels as K
    dev = "cuda"
    res = {}
    for i, (M, N, Kd, a_mn, b_mn, f32, nb, epi) in enumerate(SHAPES):
        if quick and M * N * Kd > 2 ** 36:
            continue
        g = torch.Generator(device=dev).manual_seed(100 + i)
        A = (torch.randn((nb, Kd, M) if a_mn else (nb, M, Kd), device=dev, generator=g) * 0.25).to(torch.bfloat16)
        B = (torch.randn((nb, Kd, N) if b_mn else (nb, N, Kd), device=dev, generator=g) * 0.25).to(torch.bfloat16)
        dt = torch.float32 if f32 else torch.bfloat16
        out = torch.randn(nb, M, N, device=dev, generator=g).to(dt) if epi.get("acc") else torch.zeros(
            nb, M, N, device=dev, dtype=dt)
        out0 = out.clone()
        kw = dict(nb1=nb, a_bs=(A.stride(0), 0), b_bs=(B.stride(0), 0), c_bs=(M * N, 0))
        if epi.get("bias"):
            kw["bias"] = torch.randn(N, device=dev, generator=g)
        pre = None
        if epi.get("act"):
            pre = torch.zeros_like(out)
            kw.update(act=epi["act"], c_pre=pre)
        if epi.get("drop"):
            kw.update(drop_p=0.1, seed=7, offset=11)
        if epi.get("acc"):
            kw["accumulate"] = True
        if epi.get("ag"):
            kw.update(actgrad_pre=torch.randn(nb, M, N, device=dev, generator=g).to(dt), actgrad_act=epi["ag"])

        def launch():
            K.gemm(A, B, out, M=M, N=N, K=Kd, a_mn=bool(a_mn), b_mn=bool(b_mn), **kw)
        launch()
        torch.cuda.synchronize()
        first = out.clone()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 2 if quick else 20
        if epi.get("acc"):
            out.copy_(out0)
        ev[0].record()
        for _ in range(reps):
            launch()
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        ref = None
        plain = not (set(epi) - {"bias"})
        if plain and M * N * Kd <= 2 ** 35:  # plain product (+ bias): fp32 statement of the same contraction
            Af = (A.transpose(1, 2) if a_mn else A).float()
            Bf = (B.transpose(1, 2) if b_mn else B).float()
            ref = torch.bmm(Af, Bf.transpose(1, 2))
            if epi.get("bias"):
                ref = ref + kw["bias"]
            err = ((first.float() - ref).norm() / ref.norm()).item()
        else:
            err = None
        res[i] = dict(out=first.cpu(), pre=None if pre is None else pre.cpu(), ms=ms, err=err)
        print(f"shape {i} {SHAPES[i][:7]} {ms * 1e3:8.1f} us  {2.0 * nb * M * N * Kd / ms / 1e9:7.1f} TFLOP/s"
              f"  rel-err-vs-fp32 {err}", flush=True)
    if quick:
        errs = [r["err"] for r in res.values() if r["err"] is not None]
        finite = all(bool(torch.isfinite(r["out"].float()).all()) for r in res.values())
        ok = finite and len(errs) > 0 and max(errs) < 1e-2
        print(f"QUICK mode={os.environ.get('ST5_GEMM_PAIR', '0')}: max rel err {max(errs):.3e}, finite={finite} ->",
              "OK" if ok else "FAILED")
        return 0 if ok else 1
    torch.save(res, out_path)
    return 0


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        return worker(sys.argv[2])
    if len(sys.argv) > 1 and sys.argv[1] == "--quick":  # one process, mode from the environment, fp32 reference only
        return worker(None, quick=True)
    outs = {}
    for mode in ("0", "1"):
        path = f"/tmp/gemm_pair_{mode}.pt"
        env = dict(os.environ, ST5_GEMM_PAIR=mode)
        print(f"--- ST5_GEMM_PAIR={mode}", flush=True)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", path], env=env, timeout=240)
        if r.returncode != 0:
            print(f"worker for mode {mode} failed rc={r.returncode}")
            return 2
        outs[mode] = torch.load(path)
    bad = 0
    for i in range(len(SHAPES)):
        a, b = outs["0"][i], outs["1"][i]
        same = torch.equal(a["out"], b["out"]) and (a["pre"] is None or torch.equal(a["pre"], b["pre"]))
        if not same:
            bad += 1
            d = (a["out"].float() - b["out"].float()).abs().max().item()
            print(f"shape {i} {SHAPES[i][:7]}: MISMATCH max|diff|={d}")
        print(f"shape {i}: single {a['ms'] * 1e3:8.1f} us   pair {b['ms'] * 1e3:8.1f} us   x{a['ms'] / b['ms']:.2f}"
              f"   {'identical' if same else 'DIFFERENT'}")
    print("PAIR GEMM OK" if bad == 0 else f"PAIR GEMM: {bad} shapes differ")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
