"""Pre-/post-nets on the TTS path, mirroring the reference modules (same parameter names / nesting so reference
checkpoints load):
  TextEncoderPrenet     <- speecht5/models/modules/text_encoder_prenet.py:16-45
  SpeechDecoderPrenet   <- speecht5/models/modules/speech_decoder_prenet.py:21-110  (espnet Tacotron2 Prenet inside)
  SpeechDecoderPostnet  <- speecht5/models/modules/speech_decoder_postnet.py:17-76  (espnet Tacotron2 Postnet inside)
The espnet containers (Prenet / Postnet / ScaledPositionalEncoding) are re-declared as bare parameter holders with
espnet's attribute names; espnet itself is not needed."""
import contextlib
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...ops import RT


def sinusoid_table(length, d_model, device=None):
    """espnet PositionalEncoding.extend_pe: pe[:,0::2]=sin(pos*div), pe[:,1::2]=cos(pos*div)."""
    position = torch.arange(0, length, dtype=torch.float64).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float64) * -(math.log(10000.0) / d_model))
    pe = torch.zeros(length, d_model, dtype=torch.float64)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.float().to(device) if device is not None else pe.float()


class ScaledPositionalEncoding(nn.Module):
    """espnet ScaledPositionalEncoding parameter holder: learnable `alpha`, sinusoid table cached per device."""

    def __init__(self, d_model, dropout_rate, max_len=5000):
        super().__init__()
        self.d_model, self.dropout_rate, self.max_len = d_model, dropout_rate, max_len
        self.alpha = nn.Parameter(torch.tensor(1.0))
        self._pe = None

    def reset_parameters(self):
        self.alpha.data = torch.tensor(1.0)

    def table(self, length, device):
        if self._pe is None or self._pe.shape[0] < length or self._pe.device != device:
            self._pe = sinusoid_table(max(length, min(self.max_len, 4096)), self.d_model, device)
        return self._pe


class TextEncoderPrenet(nn.Module):
    def __init__(self, embed_tokens, args):
        super().__init__()
        self.padding_idx = embed_tokens.padding_idx
        assert args.enc_use_scaled_pos_enc, "only ScaledPositionalEncoding (the reference default) is implemented"
        self.encoder_prenet = nn.Sequential(
            embed_tokens,
            ScaledPositionalEncoding(args.encoder_embed_dim, args.transformer_enc_positional_dropout_rate,
                                     max_len=args.max_text_positions))

    def forward(self, src_tokens):
        emb, pos = self.encoder_prenet[0], self.encoder_prenet[1]
        pe = pos.table(src_tokens.shape[1], src_tokens.device)
        x = ops.scaled_posenc(pe, pos.alpha, pos.dropout_rate if self.training else 0.0, tokens=src_tokens.contiguous(),
                              emb=emb.weight, padding_idx=self.padding_idx)
        return x, src_tokens.eq(self.padding_idx)


def fairseq_sinusoid_table(num_embeddings, dim, padding_idx, device):
    """fairseq/modules/sinusoidal_positional_embedding.py:36-58: [sin | cos] halves, divisor half_dim - 1, zero row at
    padding_idx (fp64 build, fp32 storage)."""
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float64) * -step)
    ang = torch.arange(num_embeddings, dtype=torch.float64)[:, None] * freq[None, :]
    emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)
    if dim % 2 == 1:
        emb = torch.cat([emb, torch.zeros(num_embeddings, 1, dtype=torch.float64)], dim=1)
    emb[padding_idx, :] = 0
    return emb.float().to(device)


class TextDecoderPrenet(nn.Module):
    """models/modules/text_decoder_prenet.py:29-124: embed_scale * E[tok] + sinusoidal positions (fairseq layout),
    dropout. One launch (the embedding + positional kernel of the text encoder prenet with a unit alpha).

    Targets are right-padded (fairseq collaters), so the position of the t-th symbol is padding_idx + 1 + t
    (fairseq/utils.py:247-257); padded slots get a position row too, but they are masked as keys and ignored by
    every loss. Built configuration: no_scale_embedding (the arch default), sinusoidal positions, no LayerNorm on the
    embedding, no quant noise -- anything else raises at construction."""

    def __init__(self, embed_tokens, args):
        super().__init__()
        assert args.no_scale_embedding, "embed_scale != 1 is not built (arch default: no_scale_embedding=True)"
        assert not getattr(args, "decoder_learned_pos", False) and not getattr(args, "layernorm_embedding", False)
        assert not getattr(args, "no_token_positional_embeddings", False) and getattr(args, "quant_noise_pq", 0) == 0
        self.embed_tokens = embed_tokens
        self.padding_idx = embed_tokens.padding_idx
        self.embed_dim = args.decoder_embed_dim
        self.dropout_p = args.dropout
        self.max_positions = args.max_text_positions
        self.register_buffer("_unit", torch.ones(()), persistent=False)
        self._pe = None
        self.num_updates = 0
        self.freeze_decoder_updates = args.freeze_decoder_updates

    def _table(self, length, device):
        need = self.padding_idx + 1 + length
        if self._pe is None or self._pe.shape[0] < need or self._pe.device != device:
            self._pe = fairseq_sinusoid_table(max(need, self.padding_idx + 1 + self.max_positions), self.embed_dim,
                                              self.padding_idx, device)
        return self._pe[self.padding_idx + 1: self.padding_idx + 1 + length]

    def forward(self, prev_output_tokens, incremental_state=None):
        if incremental_state is not None:
            raise NotImplementedError("fairseq's incremental_state protocol is not used here: the key/value cache and "
                                      "the per-step CUDA graphs live in speecht5_b200/incremental.py")
        ft = self.freeze_decoder_updates <= self.num_updates
        with torch.no_grad() if not ft else contextlib.ExitStack():
            pe = self._table(prev_output_tokens.shape[1], prev_output_tokens.device)
            x = ops.scaled_posenc(pe, self._unit, self.dropout_p if self.training else 0.0,
                                  tokens=prev_output_tokens.contiguous(), emb=self.embed_tokens.weight,
                                  padding_idx=self.padding_idx)
        return x, prev_output_tokens.eq(self.padding_idx), incremental_state

    def set_num_updates(self, num_updates):
        self.num_updates = num_updates


class TextDecoderPostnet(nn.Module):
    """models/modules/text_decoder_postnet.py:21-93: vocabulary projection, tied to the embedding under
    --share-input-output-embed (no adaptive softmax)."""

    def __init__(self, embed_tokens, vocab_size, args):
        super().__init__()
        assert getattr(args, "adaptive_softmax_cutoff", None) is None, "adaptive softmax is not built"
        d = getattr(args, "decoder_output_dim", args.decoder_embed_dim)
        self.output_projection = nn.Linear(d, vocab_size, bias=False)
        if args.share_input_output_embed:
            self.output_projection.weight = embed_tokens.weight
        else:
            nn.init.normal_(self.output_projection.weight, mean=0, std=d ** -0.5)
        self.num_updates = 0
        self.freeze_decoder_updates = args.freeze_decoder_updates

    def forward(self, x):
        ft = self.freeze_decoder_updates <= self.num_updates
        with torch.no_grad() if not ft else contextlib.ExitStack():
            return ops.linear(x, self.output_projection.weight, (), out_dtype=torch.float32)

    def set_num_updates(self, num_updates):
        self.num_updates = num_updates


class _TacotronPrenet(nn.Module):
    """espnet tacotron2.decoder.Prenet holder: prenet.{i}.0 = Linear; dropout is ALWAYS active (F.dropout default)."""

    def __init__(self, idim, n_layers=2, n_units=256, dropout_rate=0.5):
        super().__init__()
        self.dropout_rate = dropout_rate
        self.prenet = nn.ModuleList(
            [nn.Sequential(nn.Linear(idim if i == 0 else n_units, n_units), nn.ReLU()) for i in range(n_layers)])


class SpeechDecoderPrenet(nn.Module):
    def __init__(self, odim, args):
        super().__init__()
        assert args.dprenet_layers != 0 and args.dec_use_scaled_pos_enc
        decoder_input_layer = nn.Sequential(
            _TacotronPrenet(idim=odim, n_layers=args.dprenet_layers, n_units=args.dprenet_units,
                            dropout_rate=args.dprenet_dropout_rate),
            nn.Linear(args.dprenet_units, args.decoder_embed_dim))
        self.decoder_prenet = nn.Sequential(
            decoder_input_layer,
            ScaledPositionalEncoding(args.decoder_embed_dim, args.transformer_dec_positional_dropout_rate,
                                     max_len=args.max_speech_positions))
        if args.spk_embed_integration_type == "pre":
            self.spkembs_layer = nn.Sequential(
                nn.Linear(args.spk_embed_dim + args.decoder_embed_dim, args.decoder_embed_dim), nn.ReLU())
        self.embed_dim = args.decoder_embed_dim
        self.num_updates = 0
        self.freeze_decoder_updates = args.freeze_decoder_updates

    def forward(self, prev_output_tokens, tgt_lengths_in=None, spkembs=None):
        ft = self.freeze_decoder_updates <= self.num_updates
        with torch.no_grad() if not ft else contextlib.ExitStack():
            taco, lin = self.decoder_prenet[0][0], self.decoder_prenet[0][1]
            pos = self.decoder_prenet[1]
            x = prev_output_tokens.to(RT.dtype)
            for layer in taco.prenet:  # dropout applied in train AND eval (espnet semantics)
                x = ops.linear(x, layer[0].weight, layer[0].bias, act="relu", drop_p=taco.dropout_rate)
            x = ops.linear(x, lin.weight, lin.bias)
            pe = pos.table(x.shape[1], x.device)
            x = ops.scaled_posenc(pe, pos.alpha, pos.dropout_rate if self.training else 0.0, x=x)
            if spkembs is not None:
                # cat([x, normalize(spk)]) @ W^T == x @ W[:, :d]^T + (normalize(spk) @ W[:, d:]^T) as a per-utterance bias
                W, b = self.spkembs_layer[0].weight, self.spkembs_layer[0].bias
                d = self.embed_dim
                spk = F.normalize(spkembs.float()).to(RT.dtype)
                spk_bias = ops.linear(spk, W[:, d:], (), out_dtype=torch.float32, key=("spk_w", id(W)), need_dx=False)
                x = ops.linear(x, W[:, :d], b, act="relu", bias2=spk_bias, bias2_rows=x.shape[1], key=("spk_h", id(W)))
            tgt_frames_mask = None
            if tgt_lengths_in is not None:
                T = x.shape[1]
                tgt_frames_mask = torch.arange(T, device=x.device)[None, :] >= tgt_lengths_in.to(x.device)[:, None]
            return x, tgt_frames_mask

    def set_num_updates(self, num_updates):
        self.num_updates = num_updates


class _TacotronPostnet(nn.Module):
    """espnet tacotron2.decoder.Postnet holder: postnet.{i}.0 = Conv1d(no bias), postnet.{i}.1 = BatchNorm1d."""

    def __init__(self, odim, n_layers=5, n_chans=512, n_filts=5, dropout_rate=0.5, use_batch_norm=True):
        super().__init__()
        assert use_batch_norm, "the reference default (use_batch_norm=True) is implemented"
        assert n_filts == 5, "conv-as-window-GEMM path is written for the reference's 5-tap filters"
        self.dropout_rate = dropout_rate
        self.postnet = nn.ModuleList()
        for layer in range(n_layers - 1):
            ichans = odim if layer == 0 else n_chans
            self.postnet.append(nn.Sequential(
                nn.Conv1d(ichans, n_chans, n_filts, stride=1, padding=(n_filts - 1) // 2, bias=False),
                nn.BatchNorm1d(n_chans), nn.Tanh(), nn.Dropout(dropout_rate)))
        ichans = n_chans if n_layers != 1 else odim
        self.postnet.append(nn.Sequential(
            nn.Conv1d(ichans, odim, n_filts, stride=1, padding=(n_filts - 1) // 2, bias=False),
            nn.BatchNorm1d(odim), nn.Dropout(dropout_rate)))


class SpeechDecoderPostnet(nn.Module):
    def __init__(self, odim, args):
        super().__init__()
        self.feat_out = nn.Linear(args.decoder_embed_dim, odim * args.reduction_factor)
        self.prob_out = nn.Linear(args.decoder_embed_dim, args.reduction_factor)
        self.postnet = (None if args.postnet_layers == 0 else _TacotronPostnet(
            odim=odim, n_layers=args.postnet_layers, n_chans=args.postnet_chans, n_filts=args.postnet_filts,
            use_batch_norm=args.use_batch_norm, dropout_rate=args.postnet_dropout_rate))
        self.odim = odim
        self.num_updates = 0
        self.freeze_decoder_updates = args.freeze_decoder_updates

    def project(self, zs):
        """zs [B, L/r, C] -> (before [B,L,odim] fp32, stop logits [B,L] fp32): feat_out | prob_out as one GEMM."""
        B = zs.size(0)
        nf = self.feat_out.weight.shape[0]
        fo = ops.linear(zs, (self.feat_out.weight, self.prob_out.weight), (self.feat_out.bias, self.prob_out.bias),
                        out_dtype=torch.float32)
        return fo[..., :nf].reshape(B, -1, self.odim), fo[..., nf:].reshape(B, -1)

    def refine(self, before_outs):
        """before + Postnet(before) (speech_decoder_postnet.py:66-70), channels-last."""
        if self.postnet is None:
            return before_outs
        x = before_outs.to(RT.dtype)  # channels-last [B, L, odim]: the conv GEMM wants time-major rows
        n = len(self.postnet.postnet)
        for i, blk in enumerate(self.postnet.postnet):
            c = ops.conv1d_k5(x, blk[0].weight)
            x = ops.batch_norm_act(c, blk[1], training=self.training, act="tanh" if i < n - 1 else None,
                                   drop_p=self.postnet.dropout_rate if self.training else 0.0)
        return before_outs + x.float()

    def forward(self, zs):
        """zs [B, L/r, C] -> before [B,L,odim] fp32, after [B,L,odim] fp32, logits [B,L] fp32."""
        ft = self.freeze_decoder_updates <= self.num_updates
        with torch.no_grad() if not ft else contextlib.ExitStack():
            before_outs, logits = self.project(zs)
            after_outs = self.refine(before_outs)
        return before_outs, after_outs, logits

    def set_num_updates(self, num_updates):
        self.num_updates = num_updates
