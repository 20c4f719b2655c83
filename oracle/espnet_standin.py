"""TEST INFRASTRUCTURE. espnet is an un-vendored, unpinned dependency of the reference (`SpeechT5/README.md:32`,
`pip install espnet`); its source is not under /root/reference. These are restatements of the five espnet classes the
reference's hot path constructs, WITH ESPNET'S OWN CONSTRUCTOR SIGNATURES, so that the reference's unmodified modules
(`speech_decoder_prenet.py:41-67`, `speech_decoder_postnet.py:39-51`, `text_encoder_prenet.py:36-42`,
`text_to_speech_loss.py:370`) can instantiate them through `oracle/ref_loader.py`. Semantics follow espnet >= 0.10
(`espnet/nets/pytorch_backend/{tacotron2/decoder.py, transformer/embedding.py, nets_utils.py, e2e_tts_tacotron2.py}`)
and are pinned against the independent HuggingFace port in `oracle/hf_crosscheck.py` (same parameter layout:
`prenet.{i}.0`, `postnet.{i}.{0,1}`, `alpha`)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def make_pad_mask(lengths, xs=None, length_dim=-1, maxlen=None):
    if not isinstance(lengths, list):
        lengths = torch.as_tensor(lengths).long().tolist()
    bs = len(lengths)
    if maxlen is None:
        maxlen = int(max(lengths)) if xs is None else xs.size(length_dim)
    seq = torch.arange(0, maxlen, dtype=torch.int64).unsqueeze(0).expand(bs, maxlen)
    mask = seq >= seq.new_tensor(lengths).unsqueeze(-1)
    if xs is not None:
        assert length_dim in (-1, 1) and xs.dim() == 2, "stand-in: only the [B, T] form is used on the path"
        mask = mask.to(xs.device)
    return mask


def make_non_pad_mask(lengths, xs=None, length_dim=-1):
    return ~make_pad_mask(lengths, xs, length_dim)


class PositionalEncoding(nn.Module):
    """x * sqrt(d_model) + pe[:T], then dropout; pe[:, 0::2] = sin, pe[:, 1::2] = cos."""

    def __init__(self, d_model, dropout_rate, max_len=5000, reverse=False):
        super().__init__()
        self.d_model = d_model
        self.reverse = reverse
        self.xscale = math.sqrt(self.d_model)
        self.dropout = nn.Dropout(p=dropout_rate)
        self.pe = None
        self.extend_pe(torch.tensor(0.0).expand(1, max_len))

    def extend_pe(self, x):
        if self.pe is not None and self.pe.size(1) >= x.size(1):
            if self.pe.dtype != x.dtype or self.pe.device != x.device:
                self.pe = self.pe.to(dtype=x.dtype, device=x.device)
            return
        pe = torch.zeros(x.size(1), self.d_model)
        if self.reverse:
            position = torch.arange(x.size(1) - 1, -1, -1.0, dtype=torch.float32).unsqueeze(1)
        else:
            position = torch.arange(0, x.size(1), dtype=torch.float32).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, self.d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / self.d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.pe = pe.unsqueeze(0).to(device=x.device, dtype=x.dtype)

    def forward(self, x):
        self.extend_pe(x)
        return self.dropout(x * self.xscale + self.pe[:, : x.size(1)])


class ScaledPositionalEncoding(PositionalEncoding):
    """x + alpha * pe[:T], then dropout; alpha is a learned scalar initialised to 1."""

    def __init__(self, d_model, dropout_rate, max_len=5000):
        super().__init__(d_model=d_model, dropout_rate=dropout_rate, max_len=max_len)
        self.alpha = nn.Parameter(torch.tensor(1.0))

    def reset_parameters(self):
        self.alpha.data = torch.tensor(1.0)

    def forward(self, x):
        self.extend_pe(x)
        return self.dropout(x + self.alpha * self.pe[:, : x.size(1)])


class Prenet(nn.Module):
    """(Linear -> ReLU) x n_layers, each followed by F.dropout(x, p) with its default training=True (always on)."""

    def __init__(self, idim, n_layers=2, n_units=256, dropout_rate=0.5):
        super().__init__()
        self.dropout_rate = dropout_rate
        self.prenet = nn.ModuleList()
        for layer in range(n_layers):
            n_inputs = idim if layer == 0 else n_units
            self.prenet += [nn.Sequential(nn.Linear(n_inputs, n_units), nn.ReLU())]

    def forward(self, x):
        for i in range(len(self.prenet)):
            x = F.dropout(self.prenet[i](x), self.dropout_rate)
        return x


class Postnet(nn.Module):
    """(Conv1d(no bias) -> [BatchNorm1d] -> Tanh -> Dropout) x (n_layers-1), then Conv1d -> [BatchNorm1d] -> Dropout."""

    def __init__(self, idim, odim, n_layers=5, n_chans=512, n_filts=5, dropout_rate=0.5, use_batch_norm=True):
        super().__init__()
        self.postnet = nn.ModuleList()
        for layer in range(n_layers - 1):
            ichans = odim if layer == 0 else n_chans
            ochans = odim if layer == n_layers - 1 else n_chans
            conv = nn.Conv1d(ichans, ochans, n_filts, stride=1, padding=(n_filts - 1) // 2, bias=False)
            mods = [conv] + ([nn.BatchNorm1d(ochans)] if use_batch_norm else []) + [nn.Tanh(), nn.Dropout(dropout_rate)]
            self.postnet += [nn.Sequential(*mods)]
        ichans = n_chans if n_layers != 1 else odim
        conv = nn.Conv1d(ichans, odim, n_filts, stride=1, padding=(n_filts - 1) // 2, bias=False)
        mods = [conv] + ([nn.BatchNorm1d(odim)] if use_batch_norm else []) + [nn.Dropout(dropout_rate)]
        self.postnet += [nn.Sequential(*mods)]

    def forward(self, xs):
        for i in range(len(self.postnet)):
            xs = self.postnet[i](xs)
        return xs


class GuidedAttentionLoss(nn.Module):
    """Base class only: the reference overrides forward and every mask builder
    (`criterions/text_to_speech_loss.py:370-427`)."""

    def __init__(self, sigma=0.4, alpha=1.0, reset_always=True):
        super().__init__()
        self.sigma = sigma
        self.alpha = alpha
        self.reset_always = reset_always
        self.guided_attn_masks = None
        self.masks = None

    def _reset_masks(self):
        self.guided_attn_masks = None
        self.masks = None
