"""Pins the oracle against an INDEPENDENT implementation of the same model: the HuggingFace port
(transformers.models.speecht5, installed in the image; not part of the reference tree). Random oracle weights are
remapped onto SpeechT5ForTextToSpeech; both are run in eval mode (BatchNorm running statistics, dropout off, the
always-on decoder-prenet dropout disabled on both sides) on the same ragged batch.

Test infrastructure only (see oracle/speecht5_oracle.py header)."""
import torch


def oracle_to_hf_state(sd, n_enc, n_dec):
    m = {}
    m["speecht5.encoder.prenet.embed_tokens.weight"] = sd["text_encoder_prenet.encoder_prenet.0.weight"]
    m["speecht5.encoder.prenet.encode_positions.alpha"] = sd["text_encoder_prenet.encoder_prenet.1.alpha"]
    e = "speecht5.encoder.wrapped_encoder."
    for wb in ("weight", "bias"):
        m[e + f"layer_norm.{wb}"] = sd[f"encoder.layer_norm.{wb}"]
    m[e + "embed_positions.pe_k.weight"] = sd["encoder.pos_emb.pe_k.weight"]
    for i in range(n_enc):
        for wb in ("weight", "bias"):
            for p in ("q_proj", "k_proj", "v_proj", "out_proj"):
                m[e + f"layers.{i}.attention.{p}.{wb}"] = sd[f"encoder.layers.{i}.self_attn.{p}.{wb}"]
            m[e + f"layers.{i}.layer_norm.{wb}"] = sd[f"encoder.layers.{i}.self_attn_layer_norm.{wb}"]
            m[e + f"layers.{i}.feed_forward.intermediate_dense.{wb}"] = sd[f"encoder.layers.{i}.fc1.{wb}"]
            m[e + f"layers.{i}.feed_forward.output_dense.{wb}"] = sd[f"encoder.layers.{i}.fc2.{wb}"]
            m[e + f"layers.{i}.final_layer_norm.{wb}"] = sd[f"encoder.layers.{i}.final_layer_norm.{wb}"]
    p = "speecht5.decoder.prenet."
    for wb in ("weight", "bias"):
        m[p + f"layers.0.{wb}"] = sd[f"speech_decoder_prenet.decoder_prenet.0.0.prenet.0.0.{wb}"]
        m[p + f"layers.1.{wb}"] = sd[f"speech_decoder_prenet.decoder_prenet.0.0.prenet.1.0.{wb}"]
        m[p + f"final_layer.{wb}"] = sd[f"speech_decoder_prenet.decoder_prenet.0.1.{wb}"]
        m[p + f"speaker_embeds_layer.{wb}"] = sd[f"speech_decoder_prenet.spkembs_layer.0.{wb}"]
    m[p + "encode_positions.alpha"] = sd["speech_decoder_prenet.decoder_prenet.1.alpha"]
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
    for wb in ("weight", "bias"):
        m[f"speech_decoder_postnet.feat_out.{wb}"] = sd[f"speech_decoder_postnet.feat_out.{wb}"]
        m[f"speech_decoder_postnet.prob_out.{wb}"] = sd[f"speech_decoder_postnet.prob_out.{wb}"]
    for i in range(5):
        m[f"speech_decoder_postnet.layers.{i}.conv.weight"] = sd[f"speech_decoder_postnet.postnet.postnet.{i}.0.weight"]
        for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked"):
            m[f"speech_decoder_postnet.layers.{i}.batch_norm.{k}"] = sd[f"speech_decoder_postnet.postnet.postnet.{i}.1.{k}"]
    return m


def build_hf(oracle_model, n_enc, n_dec, vocab=81):
    from transformers import SpeechT5Config, SpeechT5ForTextToSpeech
    cfg = SpeechT5Config(vocab_size=vocab, encoder_layers=n_enc, decoder_layers=n_dec, speech_decoder_prenet_dropout=0.0,
                         encoder_layerdrop=0.0, decoder_layerdrop=0.0)
    hf = SpeechT5ForTextToSpeech(cfg).eval()
    missing, unexpected = hf.load_state_dict(oracle_to_hf_state(oracle_model.state_dict(), n_enc, n_dec), strict=False)
    assert not unexpected, unexpected
    assert all("embed_positions" in k or "num_batches" in k for k in missing), missing
    # p = 0 "consistent dropout" in the HF port zeroes its input (bernoulli(p=0) mask); make it the identity
    hf.speecht5.decoder.prenet._consistent_dropout = lambda x, p: x
    return hf


@torch.no_grad()
def run_hf(hf, net_input, pad=1):
    att = net_input["src_tokens"].ne(pad).long()
    T = net_input["prev_output_tokens"].shape[1]
    dec_att = (torch.arange(T)[None, :] < net_input["tgt_lengths"][:, None]).long()
    out = hf.speecht5(input_values=net_input["src_tokens"], attention_mask=att,
                      decoder_input_values=net_input["prev_output_tokens"], decoder_attention_mask=dec_att,
                      speaker_embeddings=net_input["spkembs"], output_attentions=True, return_dict=True)
    before, after, logits = hf.speech_decoder_postnet(out.last_hidden_state)
    return before, after, logits, list(out.cross_attentions), out.encoder_last_hidden_state
