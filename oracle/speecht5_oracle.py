"""CPU ORACLE (test infrastructure, NOT product code) -- a dependency-free PyTorch restatement of the SpeechT5
forward path named by BASELINE.json:north_star, written from the reference sources under
/root/reference/SpeechT5/speecht5 (cited per class as file:line). Parameter names equal the reference checkpoint keys.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.

Parity status: the reference ships no tests/golden vectors for this path and cannot be imported here (fairseq needs
omegaconf<2.1 / numpy<1.24, espnet absent) => "parity unpinned" by the reference itself. The oracle is pinned instead
against the independent HuggingFace port (transformers.models.speecht5) with remapped random weights, see
oracle/hf_crosscheck.py and tests/test_oracle_cpu.py, and frozen as fixtures under tests/golden/.

espnet pieces (not vendored by the reference; restated from the published espnet>=0.10 sources):
Prenet / Postnet (espnet/nets/pytorch_backend/tacotron2/decoder.py), ScaledPositionalEncoding
(.../transformer/embedding.py), make_non_pad_mask (.../nets_utils.py), GuidedAttentionLoss
(.../e2e_tts_tacotron2.py).
"""
import math
from argparse import Namespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ arguments
def base_args(**overrides):
    """models/speecht5.py:1252-1383 base_architecture + :1427-1447 t5_transformer_base_asr, as a Namespace."""
    a = dict(
        encoder_embed_dim=768, encoder_ffn_embed_dim=3072, encoder_layers=12, encoder_attention_heads=12,
        decoder_embed_dim=768, decoder_ffn_embed_dim=3072, decoder_layers=6, decoder_attention_heads=12,
        decoder_normalize_before=False, layer_norm_first=False, dropout=0.1, attention_dropout=0.1,
        activation_dropout=0.1, activation_fn="gelu", encoder_layerdrop=0.1, decoder_layerdrop=0.1,
        max_text_positions=600, max_speech_positions=4000, use_batch_norm=True, enc_use_scaled_pos_enc=True,
        dec_use_scaled_pos_enc=True, postnet_layers=5, postnet_chans=256, postnet_filts=5, postnet_dropout_rate=0.5,
        dprenet_dropout_rate=0.5, dprenet_layers=2, dprenet_units=256, spk_embed_integration_type="pre",
        spk_embed_dim=512, reduction_factor=2, transformer_enc_positional_dropout_rate=0.1,
        transformer_dec_positional_dropout_rate=0.1, layer_norm_eps=1e-5, share_input_output_embed=False,
        share_ctc_embed=False, freeze_encoder_updates=0, freeze_decoder_updates=0, no_freeze_encoder_layer=None,
        relative_position_embedding=True, encoder_max_relative_position=160, decoder_max_relative_position=160,
        use_sent_enc_layer=True, speech_odim=80, bert_init=False, unb_enc_layer=-1,
    )
    a.update(overrides)
    return Namespace(**a)


def make_non_pad_mask(lengths, maxlen=None):
    """espnet nets_utils.make_non_pad_mask: [B, Tmax] bool, True where t < length."""
    lengths = torch.as_tensor(lengths).long()
    maxlen = int(lengths.max()) if maxlen is None else maxlen
    return torch.arange(maxlen)[None, :] < lengths[:, None]


# ------------------------------------------------------------------------------------------------ espnet modules
class ScaledPositionalEncoding(nn.Module):
    """espnet transformer/embedding.py ScaledPositionalEncoding: x + alpha * pe, then dropout."""

    def __init__(self, d_model, dropout_rate, max_len=5000):
        super().__init__()
        self.d_model = d_model
        self.dropout = nn.Dropout(p=dropout_rate)
        self.alpha = nn.Parameter(torch.tensor(1.0))
        self.max_len = max_len

    @staticmethod
    def table(length, d_model, dtype=torch.float32):
        pe = torch.zeros(length, d_model, dtype=torch.float64)
        position = torch.arange(0, length, dtype=torch.float64).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float64) * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        return pe.to(dtype)

    def forward(self, x):
        pe = self.table(x.size(1), self.d_model, x.dtype).unsqueeze(0)
        return self.dropout(x + self.alpha * pe)


class TacotronPrenet(nn.Module):
    """espnet tacotron2/decoder.py Prenet: dropout is applied with training=True ALWAYS (also at eval)."""

    def __init__(self, idim, n_layers=2, n_units=256, dropout_rate=0.5):
        super().__init__()
        self.dropout_rate = dropout_rate
        self.prenet = nn.ModuleList()
        for layer in range(n_layers):
            n_in = idim if layer == 0 else n_units
            self.prenet.append(nn.Sequential(nn.Linear(n_in, n_units), nn.ReLU()))

    def forward(self, x):
        for layer in self.prenet:
            x = F.dropout(layer(x), self.dropout_rate)
        return x


class TacotronPostnet(nn.Module):
    """espnet tacotron2/decoder.py Postnet: (Conv1d no-bias -> BatchNorm1d -> Tanh -> Dropout) x (n-1), then
    (Conv1d -> BatchNorm1d -> Dropout)."""

    def __init__(self, odim, n_layers=5, n_chans=256, n_filts=5, dropout_rate=0.5):
        super().__init__()
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

    def forward(self, xs):
        for layer in self.postnet:
            xs = layer(xs)
        return xs


# ------------------------------------------------------------------------------------------------ attention
class MultiheadAttention(nn.Module):
    """models/modules/multihead_attention.py:23-407 (training/eval path without incremental state)."""

    def __init__(self, embed_dim, num_heads, kdim=None, dropout=0.0, self_attention=False,
                 encoder_decoder_attention=False, has_relative_attention_bias=False):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        self.scaling = self.head_dim ** -0.5
        self.dropout_p = dropout
        self.self_attention, self.encoder_decoder_attention = self_attention, encoder_decoder_attention
        self.has_relative_attention_bias = has_relative_attention_bias
        kdim = embed_dim if kdim is None else kdim
        self.k_proj = nn.Linear(kdim, embed_dim)
        self.v_proj = nn.Linear(kdim, embed_dim)
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        g = 1 / math.sqrt(2)  # :100-115
        nn.init.xavier_uniform_(self.k_proj.weight, gain=g)
        nn.init.xavier_uniform_(self.v_proj.weight, gain=g)
        nn.init.xavier_uniform_(self.q_proj.weight, gain=g)
        nn.init.xavier_uniform_(self.out_proj.weight)
        nn.init.constant_(self.out_proj.bias, 0.0)

    def forward(self, query, key, value, key_padding_mask=None, attn_mask=None, need_head_weights=False,
                position_bias=None):
        """Time x Batch x Channel in, returns (attn [T,B,C], weights [H,B,T,S] or None)."""
        tgt_len, bsz, _ = query.shape
        H, hd = self.num_heads, self.head_dim
        q = self.q_proj(query) * self.scaling  # :213-232
        k = self.k_proj(key)
        v = self.v_proj(key if self.encoder_decoder_attention or self.self_attention else value)
        src_len = k.size(0)
        q = q.contiguous().view(tgt_len, bsz * H, hd).transpose(0, 1)
        k = k.contiguous().view(-1, bsz * H, hd).transpose(0, 1)
        v = v.contiguous().view(-1, bsz * H, hd).transpose(0, 1)
        attn_weights = torch.bmm(q, k.transpose(1, 2))  # :340
        if position_bias is not None and self.has_relative_attention_bias:  # :343-353
            reshape_q = q.contiguous().view(bsz * H, -1, hd).transpose(0, 1)
            Bm = torch.matmul(reshape_q, position_bias.transpose(-2, -1))
            attn_weights = attn_weights + Bm.transpose(0, 1).view(bsz * H, position_bias.size(0), position_bias.size(1))
        if attn_mask is not None:  # :359-363
            attn_weights = attn_weights + attn_mask.unsqueeze(0)
        if key_padding_mask is not None:  # :365-377
            attn_weights = attn_weights.view(bsz, H, tgt_len, src_len).masked_fill(
                key_padding_mask.unsqueeze(1).unsqueeze(2).to(torch.bool), float("-inf")).view(bsz * H, tgt_len, src_len)
        attn_weights_float = F.softmax(attn_weights.float(), dim=-1)  # :382-385
        attn_probs = F.dropout(attn_weights_float.type_as(attn_weights), p=self.dropout_p, training=self.training)
        attn = torch.bmm(attn_probs, v)  # :389
        attn = attn.transpose(0, 1).contiguous().view(tgt_len, bsz, self.embed_dim)
        attn = self.out_proj(attn)
        w = attn_weights_float.view(bsz, H, tgt_len, src_len).transpose(1, 0) if need_head_weights else None
        return attn, w


class RelativePositionalEncoding(nn.Module):
    """models/modules/encoder.py:40-59."""

    def __init__(self, d_model, maxlen):
        super().__init__()
        self.maxlen = maxlen
        self.pe_k = nn.Embedding(2 * maxlen, d_model)

    def forward(self, pos_seq):
        pos_seq = pos_seq.clamp(-self.maxlen, self.maxlen - 1) + self.maxlen
        return self.pe_k(pos_seq)


# ------------------------------------------------------------------------------------------------ encoder
class TransformerSentenceEncoderLayer(nn.Module):
    """models/modules/transformer_layer.py:23-134."""

    def __init__(self, args):
        super().__init__()
        d, Hh = args.encoder_embed_dim, args.encoder_attention_heads
        self.layer_norm_first = args.layer_norm_first
        self.self_attn = MultiheadAttention(d, Hh, dropout=args.attention_dropout, self_attention=True,
                                            has_relative_attention_bias=args.relative_position_embedding)
        self.dropout1 = nn.Dropout(args.dropout)
        self.dropout2 = nn.Dropout(args.activation_dropout)
        self.dropout3 = nn.Dropout(args.dropout)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, args.encoder_ffn_embed_dim)
        self.fc2 = nn.Linear(args.encoder_ffn_embed_dim, d)
        self.final_layer_norm = nn.LayerNorm(d)
        if args.relative_position_embedding:
            self.norm_k = nn.LayerNorm(d // Hh)

    def forward(self, x, self_attn_padding_mask=None, pos_bias=None):
        residual = x
        if self.layer_norm_first:  # :90-111
            x = self.self_attn_layer_norm(x)
            if pos_bias is not None:
                pos_bias = self.norm_k(pos_bias)
            x, _ = self.self_attn(x, x, x, key_padding_mask=self_attn_padding_mask, position_bias=pos_bias)
            x = residual + self.dropout1(x)
            residual = x
            x = self.final_layer_norm(x)
            x = self.fc2(self.dropout2(F.gelu(self.fc1(x).float()).type_as(x)))
            x = residual + self.dropout3(x)
        else:  # :112-132
            x, _ = self.self_attn(x, x, x, key_padding_mask=self_attn_padding_mask, position_bias=pos_bias)
            x = residual + self.dropout1(x)
            x = self.self_attn_layer_norm(x)
            residual = x
            x = self.fc2(self.dropout2(F.gelu(self.fc1(x).float()).type_as(x)))
            x = residual + self.dropout3(x)
            x = self.final_layer_norm(x)
        return x


class TransformerEncoder(nn.Module):
    """models/modules/encoder.py:61-291."""

    def __init__(self, args, vocab_size=None, embed_tokens=None):
        super().__init__()
        self.args = args
        self.register_buffer("version", torch.Tensor([3]))
        self.dropout_p = args.dropout
        self.encoder_layerdrop = args.encoder_layerdrop
        self.layers = nn.ModuleList([TransformerSentenceEncoderLayer(args) for _ in range(args.encoder_layers)])
        self.layer_norm_first = args.layer_norm_first
        self.layer_norm = nn.LayerNorm(args.encoder_embed_dim, eps=args.layer_norm_eps)
        if args.share_ctc_embed and embed_tokens is not None:
            self.proj = nn.Linear(embed_tokens.weight.shape[1], embed_tokens.weight.shape[0], bias=False)
            self.proj.weight = embed_tokens.weight
        elif vocab_size is not None:
            self.proj = nn.Linear(args.encoder_embed_dim, vocab_size)
            nn.init.xavier_uniform_(self.proj.weight)  # encoder.py:30-37 Linear()
            nn.init.constant_(self.proj.bias, 0.0)
        else:
            self.proj = None
        if args.relative_position_embedding:
            self.pos_emb = RelativePositionalEncoding(args.encoder_embed_dim // args.encoder_attention_heads,
                                                      args.encoder_max_relative_position)

    def forward(self, encoder_in, encoder_padding_mask):
        if not self.layer_norm_first:  # :226-227
            encoder_in = self.layer_norm(encoder_in)
        encoder_in = F.dropout(encoder_in, self.dropout_p, self.training)
        x = encoder_in.transpose(0, 1)
        pos_k = None
        if self.args.relative_position_embedding:  # :239-246
            pos_seq = torch.arange(0, x.shape[0]).long()
            pos_k = self.pos_emb(pos_seq[:, None] - pos_seq[None, :])
        for layer in self.layers:  # :251-271 (numpy RNG layer drop)
            dropout_probability = np.random.random()
            if not self.training or dropout_probability > self.encoder_layerdrop:
                x = layer(x, self_attn_padding_mask=encoder_padding_mask, pos_bias=pos_k)
        if self.layer_norm_first:
            x = self.layer_norm(x.transpose(0, 1)).transpose(0, 1)
        out = {"encoder_out": [x], "encoder_padding_mask": [encoder_padding_mask], "encoder_states": [],
               "src_tokens": [], "decoder_input": [None]}
        out["encoder_out_for_ctc"] = [self.proj(F.dropout(x, self.dropout_p, self.training))
                                      if self.proj is not None else None]  # :173-179
        return out


# ------------------------------------------------------------------------------------------------ decoder
class TransformerDecoderLayer(nn.Module):
    """models/modules/transformer_layer.py:137-404 (no incremental state)."""

    def __init__(self, args):
        super().__init__()
        d, Hh = args.decoder_embed_dim, args.decoder_attention_heads
        self.dropout_p = args.dropout
        self.activation_dropout_p = args.activation_dropout
        self.normalize_before = args.decoder_normalize_before
        # decoder self-attention is built WITHOUT relative bias (:229-242, kwarg commented out at :241)
        self.self_attn = MultiheadAttention(d, Hh, dropout=args.attention_dropout, self_attention=True)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.encoder_attn = MultiheadAttention(d, Hh, kdim=args.encoder_embed_dim, dropout=args.attention_dropout,
                                               encoder_decoder_attention=True)
        self.encoder_attn_layer_norm = nn.LayerNorm(d)
        self.fc1 = nn.Linear(d, args.decoder_ffn_embed_dim)
        self.fc2 = nn.Linear(args.decoder_ffn_embed_dim, d)
        self.final_layer_norm = nn.LayerNorm(d)
        if args.relative_position_embedding:
            self.norm_k = nn.LayerNorm(d // Hh)  # dead parameter, kept for checkpoint parity (:219-221)

    def forward(self, x, encoder_out, encoder_padding_mask, self_attn_mask, self_attn_padding_mask, need_head_weights):
        residual = x
        if self.normalize_before:
            x = self.self_attn_layer_norm(x)
        x, _ = self.self_attn(x, x, x, key_padding_mask=self_attn_padding_mask, attn_mask=self_attn_mask)
        x = residual + F.dropout(x, self.dropout_p, self.training)
        if not self.normalize_before:
            x = self.self_attn_layer_norm(x)
        residual = x
        if self.normalize_before:
            x = self.encoder_attn_layer_norm(x)
        x, attn = self.encoder_attn(x, encoder_out, encoder_out, key_padding_mask=encoder_padding_mask,
                                    need_head_weights=need_head_weights)
        x = residual + F.dropout(x, self.dropout_p, self.training)
        if not self.normalize_before:
            x = self.encoder_attn_layer_norm(x)
        residual = x
        if self.normalize_before:
            x = self.final_layer_norm(x)
        x = F.gelu(self.fc1(x).float()).type_as(x)
        x = F.dropout(x, self.activation_dropout_p, self.training)
        x = self.fc2(x)
        x = residual + F.dropout(x, self.dropout_p, self.training)
        if not self.normalize_before:
            x = self.final_layer_norm(x)
        return x, attn


class TransformerDecoder(nn.Module):
    """models/modules/decoder.py:33-288."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.register_buffer("version", torch.Tensor([3]))
        self.decoder_layerdrop = args.decoder_layerdrop
        self.layers = nn.ModuleList([TransformerDecoderLayer(args) for _ in range(args.decoder_layers)])
        self.layer_norm = (nn.LayerNorm(args.decoder_embed_dim, eps=args.layer_norm_eps)
                           if args.decoder_normalize_before else None)
        if args.relative_position_embedding:  # dead table (:83-84)
            self.pos_emb = RelativePositionalEncoding(args.encoder_embed_dim // args.encoder_attention_heads,
                                                      args.decoder_max_relative_position)

    def forward(self, prev_output_tokens, tgt_mask, encoder_out, alignment_layer=None):
        n = len(self.layers)
        if alignment_layer is None:
            alignment_layer = n - 1
        enc = encoder_out["encoder_out"][0]
        padding_mask = encoder_out["encoder_padding_mask"][0]
        x = prev_output_tokens.transpose(0, 1)
        T = x.size(0)
        future = torch.triu(torch.full((T, T), float("-inf"), dtype=x.dtype), 1)  # :275-288
        attn_list, attn = [], None
        for idx, layer in enumerate(self.layers):
            if self.training and self.decoder_layerdrop > 0 and torch.empty(1).uniform_().item() <= self.decoder_layerdrop:
                continue  # fairseq LayerDropModuleList
            want = idx == alignment_layer or alignment_layer == -1
            x, layer_attn = layer(x, enc, padding_mask, future, tgt_mask, need_head_weights=want)
            if layer_attn is not None and want:
                attn = layer_attn.float().to(x)
                attn_list.append(attn.transpose(0, 1))  # [B,H,T,S]
        if attn is not None and len(attn_list) == 1:
            attn = attn.mean(dim=0)
        if self.layer_norm is not None:
            x = self.layer_norm(x)
        return x.transpose(0, 1), {"attn": [attn if len(attn_list) <= 1 else attn_list]}


# ------------------------------------------------------------------------------------------------ pre/post nets
class TextEncoderPrenet(nn.Module):
    """models/modules/text_encoder_prenet.py:16-45."""

    def __init__(self, embed_tokens, args):
        super().__init__()
        self.padding_idx = embed_tokens.padding_idx
        self.encoder_prenet = nn.Sequential(
            embed_tokens,
            ScaledPositionalEncoding(args.encoder_embed_dim, args.transformer_enc_positional_dropout_rate,
                                     max_len=args.max_text_positions))

    def forward(self, src_tokens):
        return self.encoder_prenet(src_tokens), src_tokens.eq(self.padding_idx)


class SpeechDecoderPrenet(nn.Module):
    """models/modules/speech_decoder_prenet.py:21-110."""

    def __init__(self, odim, args):
        super().__init__()
        inp = nn.Sequential(
            TacotronPrenet(idim=odim, n_layers=args.dprenet_layers, n_units=args.dprenet_units,
                           dropout_rate=args.dprenet_dropout_rate),
            nn.Linear(args.dprenet_units, args.decoder_embed_dim))
        self.decoder_prenet = nn.Sequential(
            inp, ScaledPositionalEncoding(args.decoder_embed_dim, args.transformer_dec_positional_dropout_rate,
                                          max_len=args.max_speech_positions))
        self.spkembs_layer = nn.Sequential(
            nn.Linear(args.spk_embed_dim + args.decoder_embed_dim, args.decoder_embed_dim), nn.ReLU())

    def forward(self, prev_output_tokens, tgt_lengths_in=None, spkembs=None):
        x = self.decoder_prenet(prev_output_tokens)
        if spkembs is not None:
            s = F.normalize(spkembs).unsqueeze(1).expand(-1, x.size(1), -1)
            x = self.spkembs_layer(torch.cat([x, s], dim=-1))
        mask = None if tgt_lengths_in is None else ~make_non_pad_mask(tgt_lengths_in, x.size(1))
        return x, mask


class SpeechDecoderPostnet(nn.Module):
    """models/modules/speech_decoder_postnet.py:17-76."""

    def __init__(self, odim, args):
        super().__init__()
        self.feat_out = nn.Linear(args.decoder_embed_dim, odim * args.reduction_factor)
        self.prob_out = nn.Linear(args.decoder_embed_dim, args.reduction_factor)
        self.postnet = TacotronPostnet(odim, args.postnet_layers, args.postnet_chans, args.postnet_filts,
                                       args.postnet_dropout_rate)
        self.odim = odim

    def forward(self, zs):
        before = self.feat_out(zs).view(zs.size(0), -1, self.odim)
        logits = self.prob_out(zs).view(zs.size(0), -1)
        after = before + self.postnet(before.transpose(1, 2)).transpose(1, 2)
        return before, after, logits


# ------------------------------------------------------------------------------------------------ model
def init_bert_params(module):
    """fairseq/modules/transformer_sentence_encoder.py:21-53."""
    def normal_(data):
        data.copy_(data.cpu().normal_(mean=0.0, std=0.02).to(data.device))
    if isinstance(module, nn.Linear):
        normal_(module.weight.data)
        if module.bias is not None:
            module.bias.data.zero_()
    if isinstance(module, nn.Embedding):
        normal_(module.weight.data)
        if module.padding_idx is not None:
            module.weight.data[module.padding_idx].zero_()
    if isinstance(module, MultiheadAttention):
        normal_(module.q_proj.weight.data)
        normal_(module.k_proj.weight.data)
        normal_(module.v_proj.weight.data)


class T5TransformerModelOracle(nn.Module):
    """models/speecht5.py:47-116 (constructor), :786-963 (forward) restricted to text-in / speech-out (t2s)."""

    def __init__(self, args, vocab_size=81, padding_idx=1):
        super().__init__()
        self.args = args
        d = args.encoder_embed_dim

        def embedding():  # fairseq/models/transformer.py:1054 Embedding()
            m = nn.Embedding(vocab_size, d, padding_idx=padding_idx)
            nn.init.normal_(m.weight, mean=0, std=d ** -0.5)
            nn.init.constant_(m.weight[padding_idx], 0)
            return m

        dec_embed = embedding()
        enc_embed = dec_embed if args.share_input_output_embed else embedding()
        self.encoder = TransformerEncoder(args, vocab_size, enc_embed)
        self.decoder = TransformerDecoder(args)
        self.text_encoder_prenet = TextEncoderPrenet(enc_embed, args)
        self.speech_decoder_prenet = SpeechDecoderPrenet(args.speech_odim, args)
        self.speech_decoder_postnet = SpeechDecoderPostnet(args.speech_odim, args)
        self.reduction_factor = args.reduction_factor
        if args.bert_init:
            self.apply(init_bert_params)

    def forward(self, src_tokens=None, src_lengths=None, prev_output_tokens=None, tgt_lengths=None, spkembs=None,
                task_name=None, **unused):
        encoder_input, encoder_padding_mask = self.text_encoder_prenet(src_tokens)
        encoder_output = self.encoder(encoder_input, encoder_padding_mask)
        dec_in, tgt_mask = self.speech_decoder_prenet(prev_output_tokens, tgt_lengths, spkembs)
        decoder_output, extra = self.decoder(dec_in, tgt_mask, encoder_output, alignment_layer=-1)  # :921-923
        return self.speech_decoder_postnet(decoder_output) + (extra["attn"][0],)

    @torch.no_grad()
    def generate_speech(self, src_tokens=None, spkembs=None, **kwargs):
        """models/speecht5.py:1188-1249 for text input: greedy frame-by-frame synthesis until a stop probability of
        the current r-frame group reaches the threshold (or maxlen). The reference reads the "threshold" key for all
        three knobs (:1190-1199: threshold, minlenratio AND maxlenratio) -- restated as is: without that key the
        values are 0.5 / 0.0 / 20.0. The decoder is re-run on the whole prefix each step, which equals the
        reference's incremental state because the self-attention is causal (identical only with the always-on
        prenet dropout set to 0: the reference draws a fresh mask for the whole prefix at every step but feeds just
        the last frame). Returns (mel [L, odim], stop probabilities [L], cross-attention [layers, heads, L/r, T])."""
        assert src_tokens is not None and src_tokens.size(0) == 1
        threshold = kwargs.get("threshold", 0.5)
        minlenratio = kwargs.get("threshold", 0.0)
        maxlenratio = kwargs.get("threshold", 20.0)
        encoder_input, encoder_padding_mask = self.text_encoder_prenet(src_tokens)
        encoder_out = self.encoder(encoder_input, encoder_padding_mask)
        r, odim = self.reduction_factor, self.speech_decoder_postnet.odim
        T_enc = encoder_out["encoder_out"][0].size(0)
        maxlen, minlen = int(T_enc * maxlenratio / r), int(T_enc * minlenratio / r)
        ys = encoder_out["encoder_out"][0].new_zeros(1, 1, odim)
        outs, probs, attns, idx = [], [], [], 0
        while True:
            idx += 1
            decoder_in, _ = self.speech_decoder_prenet(ys, spkembs=spkembs)
            z, extra = self.decoder(decoder_in, None, encoder_out, alignment_layer=-1)
            outs.append(self.speech_decoder_postnet.feat_out(z[0, -1]).view(r, odim))
            probs.append(torch.sigmoid(self.speech_decoder_postnet.prob_out(z[0, -1])))
            ys = torch.cat((ys, outs[-1][-1].view(1, 1, odim)), dim=1)
            attns.append(torch.stack([a[0, :, -1:, :] for a in extra["attn"][0]], dim=0))  # [layers, H, 1, T]
            if int((probs[-1] >= threshold).sum()) > 0 or idx >= maxlen:
                if idx < minlen:
                    continue
                mel = torch.cat(outs, dim=0).unsqueeze(0).transpose(1, 2)  # [1, odim, L]
                mel = mel + self.speech_decoder_postnet.postnet(mel)
                return mel.transpose(2, 1).squeeze(0), torch.cat(probs, dim=0), torch.cat(attns, dim=2)


# ------------------------------------------------------------------------------------------------ criterion
def guided_attention_loss(att_ws, ilens, olens, sigma=0.4, alpha=1.0):
    """criterions/text_to_speech_loss.py:370-427 GuidedMultiHeadAttentionLoss. att_ws [B, H*, T_out, T_in]."""
    Bn = len(ilens)
    max_i, max_o = int(max(ilens)), int(max(olens))
    ga = torch.zeros((Bn, max_o, max_i), dtype=att_ws.dtype)
    for idx, (il, ol) in enumerate(zip(ilens.tolist(), olens.tolist())):
        gx, gy = torch.meshgrid(torch.arange(ol), torch.arange(il), indexing="ij")
        ga[idx, :ol, :il] = 1.0 - torch.exp(-((gy.float() / il - gx.float() / ol) ** 2) / (2 * sigma ** 2))
    masks = make_non_pad_mask(olens, max_o).unsqueeze(-1) & make_non_pad_mask(ilens, max_i).unsqueeze(-2)
    losses = ga.unsqueeze(1) * att_ws
    return alpha * torch.mean(losses.masked_select(masks.unsqueeze(1)))


def tts_loss(model_out, sample, reduction_factor=2, bce_pos_weight=5.0, use_guided_attn_loss=True,
             guided_sigma=0.4, guided_lambda=1.0, num_heads_applied=2):
    """criterions/text_to_speech_loss.py:154-214 + Tacotron2Loss :296-345 (use_masking=True, loss_type L1)."""
    before, after, logits, attn = model_out
    labels, ys, olens, ilens = sample["labels"], sample["dec_target"], sample["dec_target_lengths"], sample["src_lengths"]
    r = reduction_factor
    if r > 1:
        olens_in = torch.div(olens, r, rounding_mode="floor")
        olens = olens - olens % r
        max_olen = int(olens.max())
        ys, labels = ys[:, :max_olen], labels[:, :max_olen]
        labels = torch.scatter(labels, 1, (olens - 1).unsqueeze(1), 1.0)
    else:
        olens_in = olens
    masks = make_non_pad_mask(olens, ys.size(1)).unsqueeze(-1)
    ys_m = ys.masked_select(masks)
    after_m, before_m = after.masked_select(masks), before.masked_select(masks)
    labels_m, logits_m = labels.masked_select(masks[:, :, 0]), logits.masked_select(masks[:, :, 0])
    l1 = F.l1_loss(after_m, ys_m) + F.l1_loss(before_m, ys_m)
    l2 = F.mse_loss(after_m, ys_m) + F.mse_loss(before_m, ys_m)
    bce = F.binary_cross_entropy_with_logits(logits_m, labels_m, pos_weight=torch.tensor(bce_pos_weight))
    loss = l1 + bce
    attn_loss = None
    if use_guided_attn_loss:
        att_ws = torch.cat([a[:, :num_heads_applied] for a in attn], dim=1)
        attn_loss = guided_attention_loss(att_ws, ilens, olens_in, guided_sigma, guided_lambda)
        loss = loss + attn_loss
    return loss, l1, l2, bce, attn_loss


# ------------------------------------------------------------------------------------------------ synthetic data
def synthetic_tts_batch(B, T_txt, T_mel, vocab=81, odim=80, r=2, seed=1, ragged=True, pad=1):
    """Synthetic TTS batch following data/text_to_speech_dataset.py:223-281 (collater): right-padded text, mel targets,
    prev_output_tokens = [0; thinned target][:-1], stop labels, x-vectors."""
    g = torch.Generator().manual_seed(seed)
    src_lengths = (torch.randint(int(0.75 * T_txt), T_txt + 1, (B,), generator=g) if ragged and B > 1
                   else torch.full((B,), T_txt))
    src_lengths[0] = T_txt
    src_tokens = torch.randint(4, vocab, (B, T_txt), generator=g)
    for b in range(B):
        src_tokens[b, src_lengths[b]:] = pad
    mel_lengths = (torch.randint(int(0.9 * T_mel), T_mel + 1, (B,), generator=g) if ragged and B > 1
                   else torch.full((B,), T_mel))
    mel_lengths[0] = T_mel
    fbank = torch.randn(B, T_mel, odim, generator=g)
    for b in range(B):
        fbank[b, mel_lengths[b]:] = 0.0
    fb_in = fbank[:, r - 1::r]
    len_in = torch.div(mel_lengths, r, rounding_mode="floor")
    prev = torch.cat([fb_in.new_zeros((B, 1, odim)), fb_in[:, :-1]], dim=1)
    labels = fbank.new_zeros(B, T_mel)
    for b in range(B):
        labels[b, mel_lengths[b] - 1:] = 1.0
    spk = torch.randn(B, 512, generator=g)
    net_input = dict(src_tokens=src_tokens, src_lengths=src_lengths, prev_output_tokens=prev, tgt_lengths=len_in,
                     spkembs=spk, task_name="t2s")
    return dict(net_input=net_input, labels=labels, dec_target=fbank, dec_target_lengths=mel_lengths,
                src_lengths=src_lengths, task_name="t2s", ntokens=int(src_lengths.sum()), target=fbank)
