"""Generates tests/golden/tts_tiny.npz from the CPU oracle (fp64 math, stored fp32): a tiny SpeechT5 t2s model
(1 head of 64, 2+2 layers) with its weights, a seeded ragged batch, the forward outputs, the loss terms and a few
gradients. Run in the build container: `python tests/golden/make_golden.py`. The oracle itself is pinned against the
independent HuggingFace port by tests/test_oracle_cpu.py::test_oracle_matches_hf_port."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.speecht5_oracle import T5TransformerModelOracle, base_args, synthetic_tts_batch, tts_loss  # noqa: E402

TINY = dict(encoder_embed_dim=64, encoder_ffn_embed_dim=128, encoder_layers=2, encoder_attention_heads=1,
            decoder_embed_dim=64, decoder_ffn_embed_dim=128, decoder_layers=2, decoder_attention_heads=1,
            postnet_chans=32, dprenet_units=32, encoder_max_relative_position=8, decoder_max_relative_position=8,
            dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, encoder_layerdrop=0.0, decoder_layerdrop=0.0,
            postnet_dropout_rate=0.0, dprenet_dropout_rate=0.0, transformer_enc_positional_dropout_rate=0.0,
            transformer_dec_positional_dropout_rate=0.0, bert_init=True)


def main():
    torch.manual_seed(1337)
    model = T5TransformerModelOracle(base_args(**TINY)).double().train()
    with torch.no_grad():  # non-trivial norms / alphas / BN so every parameter matters
        for n, p in model.named_parameters():
            if n.endswith("alpha"):
                p.fill_(1.2)
            elif "layer_norm" in n or "postnet.postnet" in n and p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.1)
            elif "q_proj.weight" in n or "k_proj.weight" in n or "pe_k" in n:
                p.mul_(8.0)  # peaky attention: softmax gradients are then well conditioned in fp32
    sample = synthetic_tts_batch(3, 21, 30, seed=1)
    ni = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in sample["net_input"].items()}
    out = model(**ni)
    s64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in sample.items()}
    loss, l1, l2, bce, ga = tts_loss(out, s64)
    loss.backward()
    blob = {"state/" + k: v.detach().float().numpy() for k, v in model.state_dict().items()}
    for k, v in sample["net_input"].items():
        if torch.is_tensor(v):
            blob["in/" + k] = v.numpy()
    for k in ("labels", "dec_target", "dec_target_lengths", "src_lengths"):
        blob["sample/" + k] = sample[k].numpy()
    blob["out/before"], blob["out/after"], blob["out/logits"] = [t.detach().float().numpy() for t in out[:3]]
    blob["out/attn"] = torch.stack(out[3]).detach().float().numpy()
    blob["loss"] = np.array([loss.item(), l1.item(), l2.item(), bce.item(), ga.item()])
    for n in ("encoder.pos_emb.pe_k.weight", "encoder.layers.0.self_attn.q_proj.weight",
              "text_encoder_prenet.encoder_prenet.1.alpha", "speech_decoder_postnet.postnet.postnet.0.0.weight",
              "decoder.layers.1.encoder_attn.v_proj.bias", "speech_decoder_prenet.spkembs_layer.0.weight"):
        blob["grad/" + n] = dict(model.named_parameters())[n].grad.float().numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tts_tiny.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
