"""Per-layer error growth of the CUDA path on the FULL-DEPTH Base model (12 + 6 layers), BASELINE config 1 (one 4 s
utterance: 64 text tokens, 250 mel frames -> T_dec 125), eval mode, prenet dropout 0: each encoder / decoder layer
output, the prenet, `before` and `after` of the CUDA path in both numeric modes against the fp32 CPU oracle on the same
weights. Output: gpurun_out/bf16_error_layers.txt (one line per tap: relative L2, parity mode | bf16 mode).
Run:  python tools/bf16_error_layers.py [--layers 12 6]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def taps(model, store, is_oracle):
    hs = []

    def hook(name, tbc):
        def fn(mod, inp, out):
            o = out[0] if isinstance(out, (tuple, list)) else out
            if isinstance(o, dict):
                return
            o = o.detach().float().cpu()
            if tbc and o.dim() == 3:
                o = o.transpose(0, 1)
            store[name] = o
        return fn
    for i, l in enumerate(model.encoder.layers):
        hs.append(l.register_forward_hook(hook(f"enc{i:02d}", is_oracle)))
    for i, l in enumerate(model.decoder.layers):
        hs.append(l.register_forward_hook(hook(f"dec{i:02d}", is_oracle)))
    hs.append(model.speech_decoder_prenet.register_forward_hook(hook("dprenet", False)))
    hs.append(model.text_encoder_prenet.register_forward_hook(hook("eprenet", False)))
    return hs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, nargs=2, default=[12, 6])
    a = ap.parse_args()
    from oracle.speecht5_oracle import T5TransformerModelOracle, base_args, synthetic_tts_batch
    from speecht5_b200.models import make_args
    from speecht5_b200.ops import RT
    from speecht5_b200.tasks import SpeechT5Task
    from speecht5_b200.trainer import _to_device
    over = dict(encoder_layers=a.layers[0], decoder_layers=a.layers[1], dropout=0.0, attention_dropout=0.0,
                activation_dropout=0.0, encoder_layerdrop=0.0, decoder_layerdrop=0.0, postnet_dropout_rate=0.0,
                dprenet_dropout_rate=0.0, transformer_enc_positional_dropout_rate=0.0,
                transformer_dec_positional_dropout_rate=0.0, bert_init=True)
    torch.manual_seed(1337)
    oracle = T5TransformerModelOracle(base_args(**over)).eval()
    sample = synthetic_tts_batch(1, 64, 250, seed=1, ragged=False)
    ref = {}
    hs = taps(oracle, ref, True)
    with torch.no_grad():
        o = oracle(**sample["net_input"])
    ref["before"], ref["after"], ref["logits"] = o[0], o[1], o[2]
    for h in hs:
        h.remove()
    dev = torch.device("cuda:0")
    res = {}
    for dtype in (torch.float32, torch.bfloat16):
        RT.dtype = dtype
        RT.manual_seed(1)
        RT.disable_device_seed()
        RT.clear_static()
        RT.invalidate_shadows()
        args = make_args("t5_transformer_base_asr", **over)
        task = SpeechT5Task(args)
        model = task.build_model(args).to(dev).eval()
        model.load_state_dict(oracle.state_dict())
        got = {}
        hs = taps(model, got, False)
        with torch.no_grad():
            out = model(**_to_device(sample, dev)["net_input"])
        got["before"], got["after"], got["logits"] = [t.float().cpu() for t in out[:3]]
        for h in hs:
            h.remove()
        res[dtype] = {k: ((got[k] - ref[k]).norm() / ref[k].norm()).item() for k in ref if k in got}
    RT.dtype = torch.bfloat16
    lines = [f"{'tap':10s} {'parity(fp32x3)':>16s} {'bf16':>12s}"]
    for k in sorted(ref, key=lambda s: (s[:3] not in ("epr",), s[:3] != "enc", s[:3] != "dpr", s[:3] != "dec", s)):
        lines.append(f"{k:10s} {res[torch.float32].get(k, float('nan')):16.3e} {res[torch.bfloat16].get(k, float('nan')):12.3e}")
    text = "\n".join(lines)
    print(text)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "bf16_error_layers.txt"), "w").write(text + "\n")


if __name__ == "__main__":
    main()
