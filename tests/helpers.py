"""Shared test helpers: build the product model and the oracle with identical weights / inputs."""
import numpy as np
import torch

TINY = dict(encoder_embed_dim=64, encoder_ffn_embed_dim=128, encoder_layers=2, encoder_attention_heads=1,
            decoder_embed_dim=64, decoder_ffn_embed_dim=128, decoder_layers=2, decoder_attention_heads=1,
            postnet_chans=32, dprenet_units=32, encoder_max_relative_position=8, decoder_max_relative_position=8)
NO_DROPOUT = dict(dropout=0.0, attention_dropout=0.0, activation_dropout=0.0, encoder_layerdrop=0.0,
                  decoder_layerdrop=0.0, postnet_dropout_rate=0.0, dprenet_dropout_rate=0.0,
                  transformer_enc_positional_dropout_rate=0.0, transformer_dec_positional_dropout_rate=0.0)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def load_golden(path):
    z = np.load(path)
    state = {k[len("state/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state/")}
    net_input = {k[len("in/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in/")}
    net_input["task_name"] = "t2s"
    sample = {k[len("sample/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sample/")}
    sample["net_input"] = net_input
    sample["task_name"] = "t2s"
    sample["target"] = sample["dec_target"]
    sample["ntokens"] = int(sample["src_lengths"].sum())
    out = {k[len("out/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("out/")}
    grads = {k[len("grad/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    return state, sample, out, torch.from_numpy(z["loss"]), grads


def to_device(obj, dev):
    if torch.is_tensor(obj):
        return obj.to(dev)
    if isinstance(obj, dict):
        return {k: to_device(v, dev) for k, v in obj.items()}
    return obj


def speech_pretrain_fixture_case(dev):
    """Product model + criterion + the batch of tests/golden/ref_speech_pretrain_tiny.npz (the REFERENCE model's own
    speech pre-training update, make_golden_from_ref.py:case_speech_pretrain) on `dev`."""
    import os
    import numpy as np
    from speecht5_b200.criterions import SpeechT5Criterion
    from speecht5_b200.models import T5TransformerModel, make_args
    blob = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_speech_pretrain_tiny.npz")))
    over = dict(TINY, **NO_DROPOUT, bert_init=True, build_speech_encoder=True,
                conv_feature_layers="[(32, 10, 5)] + [(32, 3, 2)] * 4 + [(32, 2, 2)] * 2", feature_grad_mult=1.0,
                conv_pos=16, conv_pos_groups=4, use_conv_pos=True, use_sinc_pos=True, mask_prob=0.5,
                hubert_mask_length=4, mask_channel_prob=0.0, use_codebook=True, latent_vars=10, latent_groups=2,
                codebook_prob=0.5, final_dim=16, untie_final_proj=True, hubert_num_classes=[23])
    model = T5TransformerModel.build_model(make_args("t5_transformer_base_asr", **over)).to(dev).train()
    missing = model.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in blob.items() if k.startswith("state/")},
                                    strict=False)
    assert not missing.unexpected_keys and all(
        k.startswith(("text_encoder_prenet.", "text_decoder")) or "num_batches_tracked" in k or "version" in k
        for k in missing.missing_keys), missing
    t = lambda k: torch.from_numpy(blob[k]).to(dev)  # noqa: E731
    ni = dict(source=t("in/source"), padding_mask=t("in/padding_mask"), prev_output_tokens=t("in/prev_output_tokens"),
              spkembs=t("in/spkembs"), tgt_lengths=t("in/tgt_lengths"), mask_indices=t("in/mask_indices"),
              task_name="speech_pretrain")
    sample = {"id": torch.arange(3), "task_name": "speech_pretrain", "net_input": ni,
              "target_list": [t("sample/target_list0")], "labels": t("sample/labels"), "dec_target": t("sample/dec_target"),
              "dec_target_lengths": t("sample/dec_target_lengths"), "src_lengths": [blob["in/source"].shape[1]] * 3}
    model._gumbel_noise, model._codebook_perm = t("in/gumbel_noise"), t("in/codebook_perm")
    crit = SpeechT5Criterion(None, pred_masked_weight=1.0, pred_nomask_weight=0.5, loss_weights=[10.0, 0.1],
                             hubert_weight=1.0, dec_weight=0.5)
    return blob, model, crit, sample


def text_pretrain_fixture_case(dev):
    """Product model + criterion + the batch of tests/golden/ref_text_pretrain_tiny.npz (the REFERENCE model's own text
    pre-training update, make_golden_from_ref.py:case_text_pretrain) on `dev`."""
    import os
    from speecht5_b200.criterions import SpeechT5Criterion
    from speecht5_b200.models import T5TransformerModel, make_args
    blob = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_text_pretrain_tiny.npz")))
    over = dict(TINY, **NO_DROPOUT, bert_init=True, build_text_decoder=True, share_input_output_embed=True,
                use_codebook=True, latent_vars=10, latent_groups=2, codebook_prob=0.5, max_text_positions=600)
    model = T5TransformerModel.build_model(make_args("t5_transformer_base_asr", **over)).to(dev).train()
    missing = model.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in blob.items() if k.startswith("state/")},
                                    strict=False)
    assert not missing.unexpected_keys and all(
        k.startswith(("speech_decoder_prenet.", "speech_decoder_postnet.")) or "num_batches_tracked" in k or "version" in k
        for k in missing.missing_keys), missing
    t = lambda k: torch.from_numpy(blob[k]).to(dev)  # noqa: E731
    sample = {"id": torch.arange(3), "task_name": "text_pretrain", "target": t("sample/target"),
              "ntokens": int(blob["loss"][2]), "nsentences": 3,
              "net_input": dict(src_tokens=t("in/src_tokens"), prev_output_tokens=t("in/prev_output_tokens"))}
    model._gumbel_noise, model._codebook_perm = t("in/gumbel_noise"), t("in/codebook_perm")
    crit = SpeechT5Criterion(None, bart_weight=1.0, loss_weights=[10.0, 0.1])
    return blob, model, crit, sample
