"""CUDA-event timings of the attention launches of one TTS step (B=32, H=12): encoder self-attention with relative
positions (T=160), decoder causal self-attention (313), cross-attention 313 x 160 with returned probabilities (and, for
the two guided layers, an external gradient on them) -- forward and backward separately, per attention path.
`--asr` adds the speech-input shapes (T=499 with clipped relative positions, cross 160 x 499)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speecht5_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--asr", action="store_true")
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--out", default=None)
args = ap.parse_args()
dev = "cuda"
H, d = 12, 768
ops.RT.dtype = torch.bfloat16
ops.RT.manual_seed(3)


SIDE = torch.cuda.Stream()


def timed(fn, reps):
    """`reps` calls captured into ONE CUDA graph (no host gaps between the launches), replayed three times."""
    torch.cuda.synchronize()
    side = SIDE  # every eager call and the capture on ONE stream (autograd binds AccumulateGrad nodes to it)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
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
    return best  # us


def case(name, B, Tq, Tk, kind, ext=False, drop=0.1):
    torch.manual_seed(0)
    pe = torch.nn.Parameter(torch.randn(320, 64, device=dev) * 0.1) if kind == "rpe" else None
    if kind == "cross":
        q = (torch.randn(B, Tq, d, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
        kv = (torch.randn(B, Tk, 2 * d, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
        pad = torch.zeros(B, Tk, dtype=torch.bool, device=dev)
        pad[:, Tk - 10:] = True

        def fwd():
            return ops.attention(q, kv, H=H, d=d, q_col=0, k_col=0, v_col=1, scale=0.125, key_pad=pad, drop_p=drop,
                                 return_probs=True)
    else:
        q = (torch.randn(B, Tq, 3 * d, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)

        def fwd():
            return ops.attention(q, None, H=H, d=d, q_col=0, k_col=1, v_col=2, scale=0.125, causal=kind == "causal",
                                 pe_k=pe, maxpos=160 if pe is not None else 0, drop_p=drop)
    with torch.no_grad():
        t_inf = timed(fwd, args.reps)
    t_fwd = timed(fwd, args.reps)
    with torch.cuda.stream(SIDE):
        out, probs = fwd()
    torch.cuda.synchronize()
    g = torch.randn_like(out)
    gp = torch.randn_like(probs) * 1e-3 if (ext and probs is not None) else None

    def bwd():
        outs, grads = [out], [g]
        if gp is not None:
            outs.append(probs)
            grads.append(gp)
        torch.autograd.backward(outs, grads, retain_graph=True)
    t_bwd = timed(bwd, args.reps)
    flops_f = 4.0 * B * H * Tq * Tk * 64 * (0.5 if kind == "causal" else 1.0)
    return dict(name=name, B=B, Tq=Tq, Tk=Tk, fwd_us=round(t_fwd, 1), fwd_nograd_us=round(t_inf, 1), bwd_us=round(t_bwd, 1),
                fwd_tflops=round(flops_f / t_fwd * 1e-6, 1), bwd_tflops=round(2.5 * flops_f / t_bwd * 1e-6, 1))


cases = [("enc_rpe_160", 32, 160, 160, "rpe", False), ("dec_self_313", 32, 313, 313, "causal", False),
         ("cross_313x160", 32, 313, 160, "cross", False), ("cross_313x160_extdP", 32, 313, 160, "cross", True)]
if args.asr:
    cases += [("enc_rpe_499", 8, 499, 499, "rpe", False), ("dec_self_160", 8, 160, 160, "causal", False),
              ("cross_160x499", 8, 160, 499, "cross", False), ("enc_rpe_781", 4, 781, 781, "rpe", False)]
res = []
for c in cases:
    try:
        r = case(*c)
    except Exception as e:  # noqa: BLE001
        r = dict(name=c[0], error=str(e).splitlines()[0][:200])
    res.append(r)
    print(json.dumps(r), flush=True)
tot_f = sum(r.get("fwd_us", 0) * n for r, n in zip(res[:4], (12, 6, 4, 2)))
tot_b = sum(r.get("bwd_us", 0) * n for r, n in zip(res[:4], (12, 6, 4, 2)))
print(json.dumps(dict(step_attention_ms=round((tot_f + tot_b) * 1e-3, 3), fwd_ms=round(tot_f * 1e-3, 3), bwd_ms=round(tot_b * 1e-3, 3),
                      note="12 encoder + 6 decoder self + 6 cross launches (2 with an external dP) per direction")))
if args.out:
    json.dump(res, open(args.out, "w"), indent=1)
