"""Where do the NON-st5 kernels of a step come from? One eager (no CUDA graph) update of the bench workload under
torch.profiler with Python stacks; every device kernel that is not one of ours (at::*, CatArrayBatchedCopy, fills,
copies -- 8-9 % of the step in profiles/r01_summary_v12.txt) is attributed to the innermost frame inside this repository
that launched it. Output: a table sorted by device time, written to gpurun_out/glue_profile.txt.
usage (GPU box): python tools/profile_glue.py [--batch 32]"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench
    from speecht5_b200.criterions import SpeechT5Criterion
    from speecht5_b200.data import synthetic_tts_batch
    from speecht5_b200.models import make_args
    from speecht5_b200.ops import RT
    from speecht5_b200.tasks import SpeechT5Task
    from speecht5_b200.trainer import B200Trainer, _to_device
    W = bench.WORKLOAD
    dev = torch.device("cuda", 0)
    RT.dtype = torch.bfloat16
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
        trainer.train_step([batches[i]])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        trainer.train_step([batches[2]])
        torch.cuda.synchronize()
    by_site = collections.defaultdict(lambda: [0.0, 0, collections.Counter()])
    total_glue = total_ours = 0.0
    for evt in prof.events():
        kernels = getattr(evt, "kernels", None) or []
        if not kernels:
            continue
        dur = sum(k.duration for k in kernels)  # us
        names = [k.name for k in kernels]
        if all("st5::" in n for n in names):
            total_ours += dur
            continue
        site = "?"
        for fr in (evt.stack or []):
            if ROOT in fr and "/tools/" not in fr and "site-packages" not in fr:
                site = fr.replace(ROOT + "/", "")
                break
        ent = by_site[(site, evt.name)]
        ent[0] += dur
        ent[1] += len(kernels)
        ent[2].update(n.split("<")[0][:48] for n in names)
        total_glue += dur
    lines = [f"glue kernels (not st5::): {total_glue / 1e3:.2f} ms in one eager step; st5:: launched from ops with a "
             f"profiler record: {total_ours / 1e3:.2f} ms (C-ABI launches are not torch ops and are not listed here)",
             f"{'us':>9} {'n':>4}  op @ innermost repo frame   [kernels]"]
    for (site, op), (dur, n, names) in sorted(by_site.items(), key=lambda kv: -kv[1][0])[:60]:
        lines.append(f"{dur:9.1f} {n:4d}  {op} @ {site}   {dict(names.most_common(2))}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = os.path.join(ROOT, "gpurun_out", "glue_profile.txt")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:25]))


if __name__ == "__main__":
    main()
