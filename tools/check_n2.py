"""Run under torchrun with 2+ ranks (NCCL): one update of a small TTS model from the same initial state and the same
per-rank batches, (a) exchange=allreduce, (b) exchange=shard (reduce-scatter + sharded Adam + bf16 all-gather), both
eager and under the captured graph; the updated bf16 shadows (what the next forward reads) and, after consolidate(), the
fp32 masters must agree between the modes, and every rank must hold the same values. Rank 0 also checks the exchanged
gradient against the mean of the ranks' local gradients (legacy_distributed_data_parallel.py:76-165 semantics)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.speecht5_oracle import synthetic_tts_batch  # noqa: E402  (input generator only)
from speecht5_b200.criterions import SpeechT5Criterion  # noqa: E402
from speecht5_b200.models import make_args  # noqa: E402
from speecht5_b200.ops import RT  # noqa: E402
from speecht5_b200.tasks import SpeechT5Task  # noqa: E402
from speecht5_b200.trainer import B200Trainer, _to_device  # noqa: E402

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
over = dict(encoder_layers=2, decoder_layers=2, dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
            encoder_layerdrop=0.0, decoder_layerdrop=0.0, postnet_dropout_rate=0.0, dprenet_dropout_rate=0.0,
            transformer_enc_positional_dropout_rate=0.0, transformer_dec_positional_dropout_rate=0.0, bert_init=True)
sample = _to_device(synthetic_tts_batch(4, 40, 64, seed=10 + rank), dev)
results = {}
for mode, graph in (("allreduce", False), ("shard", False), ("shard", True), ("allreduce", True)):
    RT.dtype = torch.bfloat16
    RT.manual_seed(1)
    RT.disable_device_seed()
    RT.clear_static()
    RT.invalidate_shadows()
    torch.manual_seed(5)
    args = make_args("t5_transformer_base_asr", **over)
    task = SpeechT5Task(args)
    model = task.build_model(args).to(dev).train()
    crit = SpeechT5Criterion(task, use_guided_attn_loss=True)
    # eps = 1: the update is ~linear in the gradient (Adam's default eps turns every noise-level gradient into a +-lr step,
    # and the modes then differ by the order of fp32 atomics, not by the exchange)
    tr = B200Trainer(model, crit, task, lr=5e-2, eps=1.0, clip_norm=25.0, use_cuda_graph=graph, exchange=mode)
    for _ in range(2):
        losses, _ = tr.train_step([sample])
    torch.cuda.synchronize()
    shadow = tr.fp.shadow.float().clone()
    tr.consolidate()
    flat = tr.fp.flat.clone()
    ref = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(ref, flat)
    same = all(torch.equal(ref[0], r) for r in ref)
    results[(mode, graph)] = (shadow, flat, float(losses[0]))
    if rank == 0:
        print(f"{mode:9s} graph={graph}: loss {float(losses[0]):.5f}  ranks identical: {same}  overlapped stages {tr.overlapped_stages}",
              flush=True)
    assert same, (mode, graph)
    del tr, model
base = results[("allreduce", False)]
ok = True
for key, (shadow, flat, loss) in results.items():
    ds = ((shadow - base[0]).norm() / base[0].norm()).item()
    df = ((flat - base[1]).norm() / base[1].norm()).item()
    if rank == 0:
        print(f"{key}: shadow rel diff {ds:.2e}, master rel diff {df:.2e}", flush=True)
    ok = ok and ds < 1e-3 and df < 2e-5
if rank == 0:
    print("N2 CHECK", "OK" if ok else "FAILED", flush=True)
dist.barrier()
os._exit(0 if ok else 1)
