"""Locate the op of the ASR update that breaks CUDA-graph capture ('capturing stream has unjoined work'): capture the
forward, forward + criterion, + backward, + the optimizer tail separately."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speecht5_b200.criterions import SpeechT5Criterion  # noqa: E402
from speecht5_b200.data import synthetic_asr_batch  # noqa: E402
from speecht5_b200.models import make_args  # noqa: E402
from speecht5_b200.ops import RT  # noqa: E402
from speecht5_b200.tasks import SpeechT5Task  # noqa: E402
from speecht5_b200.trainer import B200Trainer, _to_device  # noqa: E402

dev = torch.device("cuda")
RT.dtype = torch.bfloat16
RT.manual_seed(1)
small = os.environ.get("SMALL", "1") == "1"
margs = make_args("t5_transformer_base_asr", build_speech_encoder=True, build_text_decoder=True, bert_init=True,
                  feature_grad_mult=1.0, max_text_positions=600, **(dict(encoder_layers=2, decoder_layers=2) if small else {}))
task = SpeechT5Task(margs)
model = task.build_model(margs).to(dev).train()
crit = SpeechT5Criterion(task, label_smoothing=0.1, ce_weight=0.5, ctc_weight=0.5, zero_infinity=True)
trainer = B200Trainer(model, crit, task, lr=1e-4, clip_norm=25.0, use_cuda_graph=False)
s = trainer._with_host_draws(synthetic_asr_batch(2, 32000, 24, seed=1, pin=True))
s = _to_device(s, dev)
trainer._draw_layerdrop()
for c in (crit.text_to_speech_loss, crit.speech_to_text_loss):
    if c is not None:
        c.defer_logging = True
cap = torch.cuda.Stream()
cap.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(cap):
    trainer.train_step([s])
torch.cuda.current_stream().wait_stream(cap)
torch.cuda.synchronize()
print("eager step ok (on the capture stream)")


def attempt(name, fn):
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    try:
        with torch.cuda.graph(g, stream=cap):
            fn()
        torch.cuda.synchronize()
        print(f"capture[{name}]: ok")
    except Exception as e:  # noqa: BLE001
        print(f"capture[{name}]: FAILED {str(e).splitlines()[0]}")
        torch.cuda.synchronize()


def fwd():
    return model(**s["net_input"])


def fwd_crit():
    return crit(model, s)[0]


def fwd_bwd():
    trainer.fp.grads.zero_()
    loss = crit(model, s)[0]
    loss.backward()


def enc_only():  # (argument names as SpeechEncoderPrenet.forward)
    ni = s["net_input"]
    return model.speech_encoder_prenet(ni["source"], padding_mask=ni.get("padding_mask"), mask=True,
                                       **{k: ni[k] for k in ("mask_indices", "mask_channel_indices") if k in ni})


with torch.no_grad():
    attempt("prenet no_grad", enc_only)
    attempt("forward no_grad", fwd)
attempt("forward", fwd)
attempt("forward+criterion", fwd_crit)
attempt("forward+criterion+backward", fwd_bwd)
attempt("update", lambda: trainer._update([s]))
