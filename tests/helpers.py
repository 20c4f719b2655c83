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
