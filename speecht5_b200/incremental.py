"""Incremental decoding with a key/value cache (the reference keeps fairseq
incremental state: speecht5/models/modules/multihead_attention.py:255-330, transformer_layer.py:262-404,
decoder.py:171-269). Opt-in: `use_cache=True` on generate_speech / generate_text_greedy runs the step eagerly,
`use_cache="graph"` on generate_speech replays ONE captured CUDA graph per decoder step (SynthesisGraph below). Checked
on the CPU against the prefix-recomputing path through the kernel emulation (tests/test_frontend_cpu.py) and on the
device (tests/test_frontend_gpu.py); nothing on the training path imports this module.

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
def decoder_step(decoder, x_new, cache, need_head_weights=False, t_dev=None, span=None, self_pad=None):
    """x_new [B, 1, C] = decoder-prenet output of the newest position. Returns (x [B, 1, C], [attn [B, H, 1, S]] per
    layer or None). Evaluation semantics (no dropout, no LayerDrop) -- generation only.

    Device-side step index (the form a captured graph replays): `t_dev` int64 [1] holds the position, the new key / value
    row is written with index_copy_, and self-attention runs over the first `span` cache rows with `self_pad` (uint8
    [B, span], 1 = position > t) masking what has not been written yet -- no host scalar depends on the step."""
    assert not decoder.training
    x = _act_dtype(x_new).contiguous()
    t = cache.t
    assert t_dev is not None or t < cache.max_len
    attns = []
    for li, layer in enumerate(decoder.layers):
        sa, ca = layer.self_attn, layer.encoder_attn
        C = sa.embed_dim
        residual = x
        h = ops.residual_layer_norm(x, None, layer.self_attn_layer_norm) if layer.normalize_before else x
        qkv = ops.linear(h, (sa.q_proj.weight, sa.k_proj.weight, sa.v_proj.weight),
                         (sa.q_proj.bias, sa.k_proj.bias, sa.v_proj.bias))  # [B, 1, 3C]
        if t_dev is None:
            cache.self_kv[li][:, t] = qkv[:, 0, C:]
            a, _ = _attend(qkv, cache.self_kv[li][:, : t + 1], H=sa.num_heads, d=C, q_col=0, k_col=0, v_col=1,
                           scale=sa.scaling)
        else:
            cache.self_kv[li].index_copy_(1, t_dev, qkv[:, :, C:])
            a, _ = _attend(qkv, cache.self_kv[li][:, :span], H=sa.num_heads, d=C, q_col=0, k_col=0, v_col=1,
                           scale=sa.scaling, key_pad=self_pad)
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
    if t_dev is None:
        cache.t = t + 1
    return x, (attns if need_head_weights else None)


class SynthesisGraph:
    """Greedy speech synthesis (models/speecht5.py:1188-1249) with every decoder step = ONE CUDA-graph replay: prenet on
    the newest frame (always-on dropout drawn from the device-resident seed, advanced inside the graph), positional row
    gathered by the device step counter, the key/value-cached decoder, feat_out | prob_out, and the step's outputs
    written into preallocated result buffers at the step index. The host only replays and reads one stop flag per step
    (the reference's `int(sum(probs[-1] >= threshold)) > 0`, :1235). Self-attention spans are bucketed (128, 256, ...):
    one graph per bucket, captured on first use, so a step attends over at most 2x the keys it needs."""

    def __init__(self, model, encoder_out, spkembs, maxlen, threshold, capture=True):
        self.m = model
        self.capture = capture  # False: the same step body runs eagerly (CPU checks of the device-counter form)
        dec, post = model.decoder, model.speech_decoder_postnet
        dev = encoder_out["encoder_out"][0].device
        self.dev, self.r, self.odim = dev, model.reduction_factor, post.odim
        self.maxlen = max(int(maxlen), 1)
        self.cache = DecoderCache(dec, encoder_out, self.maxlen + 1)
        if self.cache.enc_pad is not None:
            self.cache.enc_pad = self.cache.enc_pad.to(torch.uint8).contiguous()
        S = self.cache.cross[0].shape[1]
        L, H = len(dec.layers), dec.layers[0].encoder_attn.num_heads
        self.t = torch.zeros(1, dtype=torch.int64, device=dev)
        self.ys_last = torch.zeros(1, 1, self.odim, dtype=torch.float32, device=dev)
        self.outs = torch.zeros(self.maxlen, self.r, self.odim, dtype=torch.float32, device=dev)
        self.probs = torch.zeros(self.maxlen, self.r, dtype=torch.float32, device=dev)
        self.attn = torch.zeros(self.maxlen, L, H, S, dtype=torch.float32, device=dev)
        self.stop = torch.zeros(1, dtype=torch.int32, device=dev)
        self.threshold = float(threshold)
        self.pos = torch.arange(self.maxlen + 1, device=dev)
        pre = model.speech_decoder_prenet
        self.pe = pre.decoder_prenet[1].table(self.maxlen + 1, dev)
        self.spk_bias = None
        if spkembs is not None:  # (speech_decoder_prenet.py:76-89) the speaker half of the merge layer: once per utterance
            W, d = pre.spkembs_layer[0].weight, pre.embed_dim
            spk = torch.nn.functional.normalize(spkembs.float()).to(RT.dtype)
            self.spk_bias = ops.linear(spk, W[:, d:], (), out_dtype=torch.float32, key=("spk_w", id(W)), need_dx=False)
        self.graphs = {}
        self.stream = torch.cuda.Stream(device=dev) if capture else None
        if capture:
            RT.enable_device_seed(dev)

    def _body(self, span):
        m, pre, post = self.m, self.m.speech_decoder_prenet, self.m.speech_decoder_postnet
        taco, lin, pos = pre.decoder_prenet[0][0], pre.decoder_prenet[0][1], pre.decoder_prenet[1]
        x = self.ys_last.to(RT.dtype)
        for layer in taco.prenet:  # dropout in eval too (espnet Prenet semantics)
            x = ops.linear(x, layer[0].weight, layer[0].bias, act="relu", drop_p=taco.dropout_rate)
        x = ops.linear(x, lin.weight, lin.bias)
        x = ops.scaled_posenc(self.pe.index_select(0, self.t), pos.alpha, 0.0, x=x)
        if self.spk_bias is not None:
            W, b, d = pre.spkembs_layer[0].weight, pre.spkembs_layer[0].bias, pre.embed_dim
            x = ops.linear(x, W[:, :d], b, act="relu", bias2=self.spk_bias, bias2_rows=1, key=("spk_h", id(W)))
        self_pad = (self.pos[:span] > self.t).to(torch.uint8)[None].contiguous()
        z, layer_attn = decoder_step(m.decoder, x, self.cache, need_head_weights=True, t_dev=self.t, span=span,
                                     self_pad=self_pad)
        before, logits = post.project(z.contiguous())  # [1, r, odim], [1, r]
        p = torch.sigmoid(logits)
        self.outs.index_copy_(0, self.t, before)
        self.probs.index_copy_(0, self.t, p)
        self.attn.index_copy_(0, self.t, torch.stack([a[0, :, 0, :] for a in layer_attn], 0)[None])
        self.ys_last.copy_(before[:, -1:, :])
        self.stop.copy_((p >= self.threshold).any().to(torch.int32).reshape(1))
        self.t += 1
        RT.advance_seed()

    def _span(self, t):
        span = 128
        while span < t + 1:
            span *= 2
        return min(span, self.maxlen + 1)

    @torch.no_grad()
    def step(self, t):
        """Run decoder step `t` (0-based; the device counter must hold the same value). Returns the stop flag."""
        span = self._span(t)
        if not self.capture:
            self._body(span)
            return bool(self.stop.item())
        g = self.graphs.get(span)
        if g is None:
            # one eager pass builds every weight shadow and scratch outside the capture, then its effects are undone:
            # the cache row / result rows it wrote are rewritten by the replay of the same step
            self.stream.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(self.stream):
                keep = self.ys_last.clone()
                self._body(span)
                self.t -= 1
                self.ys_last.copy_(keep)
                self.stream.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.stream):
                    self._body(span)
            torch.cuda.current_stream(self.dev).wait_stream(self.stream)
            self.graphs[span] = g
        g.replay()
        return bool(self.stop.item())

