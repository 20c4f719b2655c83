"""bf16 gradient error of the reference-vector TTS fixture under each attention path (fused / unfused tensor-core / row
kernels): separates a kernel bug from ill-conditioning of a gradient (all three paths then show the same error)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import NO_DROPOUT, TINY, rel  # noqa: E402
from test_ref_pin_gpu import _build, load, state_of  # noqa: E402
from speecht5_b200.criterions import TexttoSpeechLoss  # noqa: E402
from speecht5_b200.ops import RT  # noqa: E402

cuda = torch.device("cuda")
for name, pre_ln in (("ref_tts_tiny", False), ("ref_tts_preln_tiny", True)):
    blob = load(name)
    for label, flags in (("fused", {}), ("fused_fwd_only", dict(attn_fused_bwd=False)), ("unfused_tc", dict(attn_fused=False)),
                         ("row", dict(attn_tensor_core=False)), ("nogate", dict(ffn_gate=False)),
                         ("nostream", dict(fp32_stream=False))):
        for k, v in dict(attn_fused=True, attn_fused_bwd=True, attn_tensor_core=True, ffn_gate=True, fp32_stream=True).items():
            setattr(RT, k, v)
        for k, v in flags.items():
            setattr(RT, k, v)
        over = dict(TINY, **NO_DROPOUT, bert_init=True)
        if pre_ln:
            over.update(layer_norm_first=True, decoder_normalize_before=True)
        model = _build(cuda, torch.bfloat16, **over).train()
        model.load_state_dict(state_of(blob))
        ni = {k[3:]: torch.from_numpy(v).to(cuda) for k, v in blob.items() if k.startswith("in/")}
        before, after, logits, attn = model(**ni, task_name="t2s")
        sample = {k[7:]: torch.from_numpy(v).to(cuda) for k, v in blob.items() if k.startswith("sample/")}
        loss = TexttoSpeechLoss(None, use_guided_attn_loss=True).compute_loss(model, (before, after, logits, attn), sample)[0]
        loss.backward()
        params = dict(model.named_parameters())
        errs = {k[5:]: rel(params[k[5:]].grad, torch.from_numpy(v)) for k, v in blob.items() if k.startswith("grad/")}
        print(name, label, "after", f"{rel(after, torch.from_numpy(blob['out/after'])):.2e}",
              " ".join(f"{k.split('.')[-3] if k.count('.') > 2 else k}.{k.split('.')[-2]}={e:.3f}" for k, e in errs.items()))
