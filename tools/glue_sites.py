"""Which lines of this repository issue the ATen (non-st5::) kernels of one update? One eager update of the bench workload
under a TorchDispatchMode: every ATen op that touches a CUDA tensor and is not a pure view is counted against the
innermost frame of this repository on the Python stack (ops issued by autograd's built-in nodes have no such frame below
`backward` and are listed under the op that called the engine). Complements tools/profile_glue.py, whose profiler stacks
do not resolve for most ops. Output: gpurun_out/glue_sites.txt.
usage (GPU box): python tools/glue_sites.py [--batch 32]"""
import argparse
import collections
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VIEW_OPS = ("view", "as_strided", "transpose", "slice", "select", "detach", "alias", "expand", "unsqueeze", "squeeze",
            "permute", "_unsafe_view", "reshape", "t.default", "empty", "narrow", "split", "unbind", "unfold",
            "_local_scalar_dense", "is_pinned", "_to_copy_meta", "lift_fresh", "set_", "resize_", "chunk", "movedim",
            "is_same_size", "stride", "size", "sym_", "_reshape_alias", "new_empty", "result_type", "record_stream",
            "_has_compatible_shallow_copy_type", "is_nonzero", "diagonal", "view_as", "contiguous", "flatten", "unflatten")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.set_defaults(workload="tts")
    args = ap.parse_args()
    import torch
    from torch.utils._python_dispatch import TorchDispatchMode
    import bench
    from speecht5_b200.ops import RT
    dev = torch.device("cuda", 0)
    RT.dtype = torch.bfloat16
    from speecht5_b200.criterions import SpeechT5Criterion
    from speecht5_b200.data import synthetic_tts_batch
    from speecht5_b200.models import make_args
    from speecht5_b200.tasks import SpeechT5Task
    from speecht5_b200.trainer import B200Trainer, _to_device
    W = bench.WORKLOAD
    RT.manual_seed(1)
    torch.manual_seed(1337)
    margs = make_args(W["arch"], encoder_layerdrop=0.0, decoder_layerdrop=0.0, bert_init=True,
                      decoder_layers=W["decoder_layers"], share_input_output_embed=True, max_text_positions=600,
                      max_speech_positions=1876)
    task = SpeechT5Task(margs)
    model = task.build_model(margs).to(dev).train()
    trainer = B200Trainer(model, SpeechT5Criterion(task, use_guided_attn_loss=True), task, lr=1e-4, betas=(0.9, 0.98),
                          eps=1e-8, clip_norm=25.0, use_cuda_graph=False)
    batches = [_to_device(synthetic_tts_batch(args.batch, W["text_len"], W["mel_frames"], seed=i), dev) for i in range(3)]
    for i in range(2):
        trainer.train_step([batches[i % len(batches)]])
    torch.cuda.synchronize()
    sites = collections.Counter()

    class Count(TorchDispatchMode):
        def __torch_dispatch__(self, func, types, a=(), kw=None):
            out = func(*a, **(kw or {}))
            name = str(func)
            short = name.replace("aten.", "")
            if any(short.startswith(v) for v in VIEW_OPS):
                return out
            flat = list(a) + list((kw or {}).values()) + (list(out) if isinstance(out, (tuple, list)) else [out])
            if not any(isinstance(t, torch.Tensor) and t.is_cuda for t in flat):
                return out
            site = "?"
            for fr in reversed(traceback.extract_stack()[:-1]):
                fn = fr.filename
                if fn.startswith(ROOT) and "/tools/" not in fn and not fn.endswith("bench.py"):
                    site = f"{fn[len(ROOT) + 1:]}:{fr.lineno} {fr.name}"
                    break
            if site == "?":  # issued by one of autograd's built-in nodes: name the node and the tensor it produced
                node = torch._C._current_autograd_node()
                first = next((t for t in flat if isinstance(t, torch.Tensor)), None)
                site = f"? autograd node {type(node).__name__ if node is not None else None} {tuple(first.shape) if first is not None else ''}"
            sites[(site, short)] += 1
            return out

    with Count():
        trainer.train_step([batches[2 % len(batches)]])
    torch.cuda.synchronize()
    lines = [f"{sum(sites.values())} ATen ops on CUDA tensors in one eager {args.workload} update (views excluded), by call site"]
    by_site = collections.defaultdict(list)
    for (site, op), n in sites.items():
        by_site[site].append((n, op))
    for site, ops_ in sorted(by_site.items(), key=lambda kv: -sum(n for n, _ in kv[1])):
        tot = sum(n for n, _ in ops_)
        lines.append(f"{tot:5d}  {site}   " + ", ".join(f"{op} x{n}" for n, op in sorted(ops_, reverse=True)))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", f"glue_sites_{args.workload}.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
