"""Incremental decoding with a key/value cache (the reference keeps fairseq incremental
state: speecht5/models/modules/multihead_attention.py:255-330, transformer_layer.py:262-404, decoder.py:171-269).

EXPERIMENTAL -- written at the end of round 1 without GPU time; opt-in (`use_cache=True` on generate_speech /
generate_text_greedy). Checked on the CPU against the prefix-recomputing path through the kernel emulation
(tests/test_frontend_cpu.py); nothing on the training path imports this module.

Per utterance batch the cross-attention keys / values of every decoder layer are projected ONCE; every step projects
q | k | v of the single new row, appends k | v to the layer's [B, T_max, 2C] cache and attends over the cache with a
one-row query -- O(L) work per step instead of the O(L^2) prefix recomputation. All contractions run on the existing
GEMM, the one-row attention on the exact row kernels (st5_attn_fwd), which take strided key / value views."""
import torch

from . import ops
from .ops import RT


def _act_dtype(x):
    return x if x.dtype == RT.dtype else x.to(RT.dtype)


def _attend(*a, **kw):
    """One-row queries: the row kernels (the fused tcgen05 path works on 128-query tiles)."""
    prev = RT.attn_tensor_core
    RT.attn_tensor_core = False
    try:
        return ops.attention(*a, **kw)
    finally:
        RT.attn_tensor_core = prev


class DecoderCache:
    def __init__(self, decoder, encoder_out, max_len):
        enc = encoder_out.get("_encoder_out_btc")
        if enc is None:
            enc = encoder_out["encoder_out"][0].transpose(0, 1).contiguous()
        enc = _act_dtype(enc)
        pm = encoder_out["encoder_padding_mask"]
        self.enc_pad = pm[0] if len(pm) > 0 else None
        B, _, C = enc.shape
        self.t = 0
        self.max_len = max_len
        self.cross, self.self_kv = [], []
        for layer in decoder.layers:
            ca = layer.encoder_attn
            self.cross.append(ops.linear(enc, (ca.k_proj.weight, ca.v_proj.weight), (ca.k_proj.bias, ca.v_proj.bias)))
            self.self_kv.append(torch.zeros((B, max_len, 2 * C), dtype=RT.dtype, device=enc.device))

    def reorder(self, new_order):
        """decoder.reorder_incremental_state_scripting (beam reordering): select utterances along the batch axis."""
        self.cross = [c.index_select(0, new_order) for c in self.cross]
        self.self_kv = [c.index_select(0, new_order) for c in self.self_kv]
        if self.enc_pad is not None:
            self.enc_pad = self.enc_pad.index_select(0, new_order)


@torch.no_grad()
def decoder_step(decoder, x_new, cache, need_head_weights=False):
    """x_new [B, 1, C] = decoder-prenet output of the newest position. Returns (x [B, 1, C], [attn [B, H, 1, S]] per
    layer or None). Evaluation semantics (no dropout, no LayerDrop) -- generation only."""
    assert not decoder.training and cache.t < cache.max_len
    x = _act_dtype(x_new).contiguous()
    t = cache.t
    attns = []
    for li, layer in enumerate(decoder.layers):
        sa, ca = layer.self_attn, layer.encoder_attn
        C = sa.embed_dim
        residual = x
        h = ops.residual_layer_norm(x, None, layer.self_attn_layer_norm) if layer.normalize_before else x
        qkv = ops.linear(h, (sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight),
                         (sa.q_proj.bias, sa.k_proj.bias, sa.v_proj.bias))  # [B, 1, 3C]
        cache.self_kv[li][:, t] = qkv[:, 0, C:]
        a, _ = _attend(qkv, cache.self_kv[li][:, : t + 1], H=sa.num_heads, d=C, q_col=0, k_col=0, v_col=1,
                       scale=sa.scaling)
        if layer.normalize_before:
            x = ops.linear(a, sa.out_proj.weight, sa.out_proj.bias, residual=residual)
        else:
            x = ops.residual_layer_norm(ops.linear(a, sa.out_proj.weight, sa.out_proj.bias), residual,
                                        layer.self_attn_layer_norm)
        residual = x
        h = ops.residual_layer_norm(x, None, layer.encoder_attn_layer_norm) if layer.normalize_before else x
        q = ops.linear(h, ca.q_proj.weight, ca.q_proj.bias)
        a, probs = _attend(q, cache.cross[li], H=ca.num_heads, d=C, q_col=0, k_col=0, v_col=1, scale=ca.scaling,
                           key_pad=cache.enc_pad, return_probs=need_head_weights)
        if need_head_weights:
            attns.append(probs.float())
        if layer.normalize_before:
            x = ops.linear(a, ca.out_proj.weight, ca.out_proj.bias, residual=residual)
            x = ops.ffn(ops.residual_layer_norm(x, None, layer.final_layer_norm), layer.fc1, layer.fc2,
                        layer.activation_fn, residual=x)
        else:
            x = ops.residual_layer_norm(ops.linear(a, ca.out_proj.weight, ca.out_proj.bias), residual,
                                        layer.encoder_attn_layer_norm)
            x = ops.residual_layer_norm(ops.ffn(x, layer.fc1, layer.fc2, layer.activation_fn), x,
                                        layer.final_layer_norm)
    if decoder.layer_norm is not None:
        x = ops.residual_layer_norm(x, None, decoder.layer_norm)
    cache.t = t + 1
    return x, (attns if need_head_weights else None)
