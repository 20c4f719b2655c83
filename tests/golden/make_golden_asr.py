"""Generates tests/golden/asr_tiny.npz from the CPU oracle of the speech-in / text-out path (fp64 math, stored fp32):
a tiny SpeechT5 s2t model (1 head of 64, 2+2 layers, 32-channel conv front-end) with its weights, a seeded ragged
batch (waveform padding mask + ragged targets), fixed time/channel mask draws, the forward outputs, the loss terms
(label-smoothed CE + CTC) and a few gradients. Run in the build container: `python tests/golden/make_golden_asr.py`.
The oracle itself is pinned against the independent HuggingFace port by
tests/test_oracle_cpu.py::test_asr_oracle_matches_hf_port."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.speecht5_oracle_asr import (T5TransformerModelASROracle, asr_loss, base_asr_args,  # noqa: E402
                                        compute_mask_indices_static, synthetic_asr_batch)

TINY = dict(encoder_embed_dim=64, encoder_ffn_embed_dim=128, encoder_layers=2, encoder_attention_heads=1,
            decoder_embed_dim=64, decoder_ffn_embed_dim=128, decoder_layers=2, decoder_attention_heads=1,
            encoder_max_relative_position=8, decoder_max_relative_position=8, dropout=0.0, attention_dropout=0.0,
            activation_dropout=0.0, conv_feature_layers=[(32, 10, 5)] + [(32, 3, 2)] * 4 + [(32, 2, 2)] * 2,
            conv_pos=16, conv_pos_groups=4, feature_grad_mult=1.0, bert_init=True)


def main():
    torch.manual_seed(4242)
    model = T5TransformerModelASROracle(base_asr_args(**TINY), vocab_size=41).double().train()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "layer_norm" in n or n.endswith("conv_layers.0.2.weight") or n.endswith("conv_layers.0.2.bias"):
                p.add_(torch.randn_like(p) * 0.1)
            elif "q_proj.weight" in n or "k_proj.weight" in n or "pe_k" in n:
                p.mul_(8.0)
    sample = synthetic_asr_batch(3, 6000, 9, vocab=41, seed=5)
    ni = dict(sample["net_input"])
    ni["source"] = ni["source"].double()
    with torch.no_grad():
        T = int(model.speech_encoder_prenet.feature_extractor.get_out_seq_lens_tensor(torch.tensor([6000]))[0])
        enc_pad = model.speech_encoder_prenet.forward_padding_mask(torch.zeros(3, T, 1), ni["padding_mask"])
    rng = np.random.default_rng(11)
    ni["mask_indices"] = compute_mask_indices_static(3, T, enc_pad, 0.3, 3, rng)
    ni["mask_channel_indices"] = compute_mask_indices_static(3, 64, None, 0.25, 8, rng, min_masks=0)
    s64 = dict(sample, net_input=ni)
    loss, ce, ctc, ss = asr_loss(model, s64)
    loss.backward()
    with torch.no_grad():
        (logits, _), enc = model(**ni)
    blob = {"state/" + k: v.detach().float().numpy() for k, v in model.state_dict().items()}
    for k, v in ni.items():
        if torch.is_tensor(v):
            blob["in/" + k] = v.float().numpy() if v.is_floating_point() else v.numpy()
    blob["sample/target"], blob["sample/target_lengths"] = sample["target"].numpy(), sample["target_lengths"].numpy()
    blob["out/logits"] = logits.float().numpy()
    blob["out/encoder_out"] = enc["encoder_out"][0].float().numpy()
    blob["out/encoder_padding_mask"] = enc["encoder_padding_mask"][0].numpy()
    blob["out/encoder_out_for_ctc"] = enc["encoder_out_for_ctc"][0].float().numpy()
    blob["loss"] = np.array([loss.item(), ce.item(), ctc.item(), float(ss), enc["features_pen"].item()])
    params = dict(model.named_parameters())
    for n in ("speech_encoder_prenet.feature_extractor.conv_layers.0.0.weight",
              "speech_encoder_prenet.feature_extractor.conv_layers.3.0.weight",
              "speech_encoder_prenet.pos_conv_v", "speech_encoder_prenet.pos_conv_g", "speech_encoder_prenet.mask_emb",
              "speech_encoder_prenet.post_extract_proj.weight", "encoder.proj.weight",
              "text_decoder_prenet.embed_tokens.weight", "text_decoder_postnet.output_projection.weight",
              "decoder.layers.1.encoder_attn.v_proj.bias"):
        blob["grad/" + n] = params[n].grad.float().numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "asr_tiny.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path), "loss", loss.item(), ce.item(), ctc.item())


if __name__ == "__main__":
    main()
