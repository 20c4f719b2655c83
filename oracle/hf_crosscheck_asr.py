"""Pins the speech-in / text-out oracle (oracle/speecht5_oracle_asr.py) against the INDEPENDENT HuggingFace port
(transformers SpeechT5ForSpeechToText; not part of the reference tree): random oracle weights are remapped onto the HF
model and both run in eval mode on the same batch.

The two implementations derive the frame-level padding mask differently (reference: a frame is padding iff ALL of its
320 samples are, speech_encoder_prenet.py:219-229; HF: frame index >= conv_out_length(valid samples)), so the
cross-check uses equal-length waveforms (no frame mask) and ragged TEXT targets; the reference rule itself is covered
by a direct test of forward_padding_mask.

Test infrastructure only (see oracle/speecht5_oracle.py header)."""
import torch


def oracle_asr_to_hf_state(sd, n_enc, n_dec):
    m = {}
    p = "speecht5.encoder.prenet."
    o = "speech_encoder_prenet."
    for i in range(7):
        m[p + f"feature_encoder.conv_layers.{i}.conv.weight"] = sd[o + f"feature_extractor.conv_layers.{i}.0.weight"]
    for wb in ("weight", "bias"):
        m[p + f"feature_encoder.conv_layers.0.layer_norm.{wb}"] = sd[o + f"feature_extractor.conv_layers.0.2.{wb}"]
        m[p + f"feature_projection.layer_norm.{wb}"] = sd[o + f"layer_norm.{wb}"]
        m[p + f"feature_projection.projection.{wb}"] = sd[o + f"post_extract_proj.{wb}"]
    m[p + "pos_conv_embed.conv.bias"] = sd[o + "pos_conv_bias"]
    m[p + "pos_conv_embed.conv.parametrizations.weight.original0"] = sd[o + "pos_conv_g"]
    m[p + "pos_conv_embed.conv.parametrizations.weight.original1"] = sd[o + "pos_conv_v"]
    # shared encoder / decoder stacks (same key scheme as oracle/hf_crosscheck.py)
    e = "speecht5.encoder.wrapped_encoder."
    for wb in ("weight", "bias"):
        m[e + f"layer_norm.{wb}"] = sd[f"encoder.layer_norm.{wb}"]
    m[e + "embed_positions.pe_k.weight"] = sd["encoder.pos_emb.pe_k.weight"]
    for i in range(n_enc):
        for wb in ("weight", "bias"):
            for pr in ("q_proj", "k_proj", "v_proj", "out_proj"):
                m[e + f"layers.{i}.attention.{pr}.{wb}"] = sd[f"encoder.layers.{i}.self_attn.{pr}.{wb}"]
            m[e + f"layers.{i}.layer_norm.{wb}"] = sd[f"encoder.layers.{i}.self_attn_layer_norm.{wb}"]
            m[e + f"layers.{i}.feed_forward.intermediate_dense.{wb}"] = sd[f"encoder.layers.{i}.fc1.{wb}"]
            m[e + f"layers.{i}.feed_forward.output_dense.{wb}"] = sd[f"encoder.layers.{i}.fc2.{wb}"]
            m[e + f"layers.{i}.final_layer_norm.{wb}"] = sd[f"encoder.layers.{i}.final_layer_norm.{wb}"]
    d = "speecht5.decoder.wrapped_decoder."
    for i in range(n_dec):
        for wb in ("weight", "bias"):
            for att in ("self_attn", "encoder_attn"):
                for pr in ("q_proj", "k_proj", "v_proj", "out_proj"):
                    m[d + f"layers.{i}.{att}.{pr}.{wb}"] = sd[f"decoder.layers.{i}.{att}.{pr}.{wb}"]
            for ln in ("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm"):
                m[d + f"layers.{i}.{ln}.{wb}"] = sd[f"decoder.layers.{i}.{ln}.{wb}"]
            m[d + f"layers.{i}.feed_forward.intermediate_dense.{wb}"] = sd[f"decoder.layers.{i}.fc1.{wb}"]
            m[d + f"layers.{i}.feed_forward.output_dense.{wb}"] = sd[f"decoder.layers.{i}.fc2.{wb}"]
    m["speecht5.decoder.prenet.embed_tokens.weight"] = sd["text_decoder_prenet.embed_tokens.weight"]
    m["text_decoder_postnet.lm_head.weight"] = sd["text_decoder_postnet.output_projection.weight"]
    return m


def build_hf_asr(oracle_model, n_enc, n_dec, vocab=81):
    from transformers import SpeechT5Config, SpeechT5ForSpeechToText
    cfg = SpeechT5Config(vocab_size=vocab, encoder_layers=n_enc, decoder_layers=n_dec, encoder_layerdrop=0.0,
                         decoder_layerdrop=0.0, mask_time_prob=0.0, mask_feature_prob=0.0, apply_spec_augment=False,
                         max_text_positions=600, tie_word_embeddings=False)
    hf = SpeechT5ForSpeechToText(cfg).eval()
    missing, unexpected = hf.load_state_dict(oracle_asr_to_hf_state(oracle_model.state_dict(), n_enc, n_dec),
                                             strict=False)
    assert not unexpected, unexpected
    assert all(("embed_positions" in k and "pe_k" not in k) or "masked_spec_embed" in k for k in missing), missing
    return hf


def compare(n_enc=2, n_dec=2, B=2, n_samples=8000, T_tgt=11, seed=0):
    """Returns relative L2 errors {decoder logits, encoder states} between the oracle and the HF port."""
    from .speecht5_oracle_asr import T5TransformerModelASROracle, base_asr_args, synthetic_asr_batch
    torch.manual_seed(seed)
    args = base_asr_args(encoder_layers=n_enc, decoder_layers=n_dec, bert_init=True)
    oracle = T5TransformerModelASROracle(args).eval()
    with torch.no_grad():  # give the zero-initialised biases / unit norms some signal
        for n, p_ in oracle.named_parameters():
            if n.endswith("bias") or "layer_norm" in n or n.endswith("conv_layers.0.2.weight"):
                p_.add_(0.05 * torch.randn_like(p_))
    hf = build_hf_asr(oracle, n_enc, n_dec)
    s = synthetic_asr_batch(B, n_samples, T_tgt, seed=seed + 1, ragged=True)
    ni = dict(s["net_input"])
    ni["source"] = torch.randn(B, n_samples, generator=torch.Generator().manual_seed(seed + 2)) * 0.1
    ni["padding_mask"] = None  # equal-length waveforms (see module docstring)
    with torch.no_grad():
        (logits, _), enc = oracle(**ni)
        dec_mask = ni["prev_output_tokens"].ne(1).long()
        out = hf(input_values=ni["source"], decoder_input_ids=ni["prev_output_tokens"], decoder_attention_mask=dec_mask,
                 output_hidden_states=False)
    keep = s["target"].ne(1)

    def rel(a, b):
        return float((a - b).norm() / b.norm())
    return {"logits": rel(logits[keep], out.logits[keep]),
            "encoder": rel(enc["encoder_out"][0].transpose(0, 1), out.encoder_last_hidden_state)}


if __name__ == "__main__":
    print(compare())
