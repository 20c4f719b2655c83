"""Data-parallel update step for the SpeechT5 hot path (host side).

Mirrors what fairseq does around the model for one update -- fairseq/trainer.py:675-952 (zero_grad -> per-micro-batch
task.train_step -> all_reduce_grads -> multiply_grads(world / sample_size) -> clip_grad_norm -> optimizer.step) with
LegacyDistributedDataParallel semantics for the gradient exchange (legacy_distributed_data_parallel.py:76-165: grads
/= world, all-reduce(sum), parameters without a gradient contribute zeros) -- re-designed for B200:

  * parameters, gradients, Adam moments live in single flat fp32 buffers plus a flat bf16 shadow; fused operand groups
    (q|k|v weights, k|v, feat_out|prob_out, and their biases) are laid out adjacently so the GEMMs read them in place;
  * the gradient exchange runs over the flat buffer in fixed-size buckets (NCCL over NVLink; gloo for CPU tests);
  * clip + Adam + bf16-shadow refresh is one kernel launch (st5_adam_step) with lr / step read from device memory;
  * the whole update (zero, forward, backward, exchange, norm, Adam) is captured once into a CUDA graph and replayed --
    the ~1.5k kernel launches per step otherwise leave the GPU waiting on Python.
"""
import os

import torch
import torch.distributed as dist

from . import kernels as K
from .models.modules.transformer import MultiheadAttention
from .ops import RT


class GradBucketer:
    """Mean-reduces a flat gradient buffer across ranks in fixed-size buckets. Device agnostic (NCCL or gloo)."""

    def __init__(self, flat_grads, bucket_elems=32 * 1024 * 1024, group=None):
        self.flat = flat_grads
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        n = flat_grads.numel()
        self.bounds = [(s, min(n, s + bucket_elems)) for s in range(0, n, bucket_elems)]

    def all_reduce_mean(self):
        if self.world == 1:
            return
        for s, e in self.bounds:
            chunk = self.flat[s:e]
            chunk.div_(self.world)  # legacy_ddp.py:110 (div before the sum keeps fp16/bf16 ranges safe)
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)

    def all_reduce_sum(self, lo=0, hi=None):
        """Sum (not mean) of flat[lo:hi] over the ranks, bucket by bucket; the caller folds 1/world into the gradient
        multiplier it applies anyway (fp32 buffer: no range concern, one elementwise pass over the buffer less)."""
        if self.world == 1:
            return
        hi = self.flat.numel() if hi is None else hi
        step = self.bounds[0][1] - self.bounds[0][0] if self.bounds else hi
        for s in range(lo, hi, max(1, step)):
            dist.all_reduce(self.flat[s:min(hi, s + step)], op=dist.ReduceOp.SUM, group=self.group)


def _fused_groups(model):
    """Parameter groups that must be adjacent in the flat buffers (read as one fused GEMM operand)."""
    groups = []
    for m in model.modules():
        if isinstance(m, MultiheadAttention):
            if m.self_attention:
                groups.append([m.q_proj.weight, m.k_proj.weight, m.v_proj.weight])
                groups.append([m.q_proj.bias, m.k_proj.bias, m.v_proj.bias])
            else:
                groups.append([m.k_proj.weight, m.v_proj.weight])
                groups.append([m.k_proj.bias, m.v_proj.bias])
    post = getattr(model, "speech_decoder_postnet", None)
    if post is not None:
        groups.append([post.feat_out.weight, post.prob_out.weight])
        groups.append([post.feat_out.bias, post.prob_out.bias])
    return [g for g in groups if all(p is not None for p in g)]


class FlatParams:
    def __init__(self, model):
        params = [p for p in model.parameters()]
        dev = params[0].device
        groups = _fused_groups(model)
        in_group = {id(p) for g in groups for p in g}
        order, seen = [], set()
        gmap = {id(g[0]): g for g in groups}
        for p in params:
            if id(p) in seen:
                continue
            if id(p) in gmap:
                order.append(gmap[id(p)])
                seen.update(id(q) for q in gmap[id(p)])
            elif id(p) not in in_group:
                order.append([p])
                seen.add(id(p))
        for g in groups:  # groups whose first member was not met first in parameters() order
            if id(g[0]) not in seen:
                order.append(g)
                seen.update(id(q) for q in g)
        # encoder-side groups first, everything whose gradient is complete once d(encoder_out) exists (decoder, decoder
        # prenet, postnet) behind them: the second range can be exchanged while the encoder is still in backward
        names = {id(p): n for n, p in model.named_parameters()}
        enc_side = ("encoder.", "text_encoder_prenet.", "speech_encoder_prenet.")
        is_enc = lambda g: names.get(id(g[0]), "").startswith(enc_side)  # noqa: E731
        order = [g for g in order if is_enc(g)] + [g for g in order if not is_enc(g)]
        offsets, off, split = {}, 0, None
        for g in order:
            off = (off + 7) // 8 * 8  # 16-byte alignment of every bf16 operand (TMA)
            if split is None and not is_enc(g):
                split = off
            for p in g:
                offsets[id(p)] = off
                off += p.numel()
        self.numel = (off + 7) // 8 * 8
        self.split = self.numel if split is None else split  # [0, split) encoder side, [split, numel) decoder side
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.grads = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.shadow = torch.zeros(self.numel, dtype=torch.bfloat16, device=dev)
        self.offsets = offsets
        with torch.no_grad():
            for p in params:
                o, n = offsets[id(p)], p.numel()
                self.flat[o:o + n].copy_(p.detach().reshape(-1))
                p.data = self.flat[o:o + n].view(p.shape)
                p.grad = self.grads[o:o + n].view(p.shape)
        self.refresh_shadow()
        # static shadows / fused biases for ops.linear
        RT.clear_static()
        for p in params:
            if p.dim() == 2:
                o = offsets[id(p)]
                RT.register_static(("lin", id(p)), self.shadow[o:o + p.numel()].view(p.shape))
                RT.register_static_grad(("lin", id(p)), self.grads[o:o + p.numel()].view(p.shape))
            elif p.dim() == 1:
                o = offsets[id(p)]
                RT.register_static_grad(("bias", id(p)), self.grads[o:o + p.numel()])
        for g in groups:
            o = offsets[id(g[0])]
            n = sum(p.numel() for p in g)
            if g[0].dim() == 2:
                RT.register_static(("lin",) + tuple(id(p) for p in g), self.shadow[o:o + n].view(-1, g[0].shape[1]))
                RT.register_static_grad(("lin",) + tuple(id(p) for p in g), self.grads[o:o + n].view(-1, g[0].shape[1]))
            else:
                RT.register_static(("bias",) + tuple(id(p) for p in g), self.flat[o:o + n])
                RT.register_static_grad(("bias",) + tuple(id(p) for p in g), self.grads[o:o + n])
        self.params = params

    def refresh_shadow(self):
        K.cast_bf16(self.flat.view(1, -1), self.shadow.view(1, -1))
        RT.invalidate_shadows()


class B200Trainer:
    """One process per GPU. `train_step(samples)` == fairseq Trainer.train_step for the speecht5 task."""

    def __init__(self, model, criterion, task, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.0, clip_norm=25.0,
                 process_group=None, use_cuda_graph=True, bucket_mb=128):
        self.model, self.criterion, self.task = model, criterion, task
        self.lr, self.betas, self.eps, self.weight_decay, self.clip_norm = lr, betas, eps, weight_decay, clip_norm
        self.device = next(model.parameters()).device
        self.criterion.to(self.device)  # criterion buffers (BCE pos_weight) must live on the device for capture
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.fp = FlatParams(model)
        self.bucketer = GradBucketer(self.fp.grads, bucket_elems=bucket_mb * 1024 * 1024 // 4, group=process_group)
        # overlap of the exchange with backward: when d(encoder_out) arrives, every decoder-side gradient is final
        # (autograd runs the later-created decoder nodes first), so that range is all-reduced on a side stream while
        # the encoder layers are still in backward. ST5_OVERLAP_AR=0 disables it.
        self._tail_launched, self._last_micro = False, False
        self._side = None
        if (self.world > 1 and self.device.type == "cuda" and os.environ.get("ST5_OVERLAP_AR", "1") != "0"
                and 0 < self.fp.split < self.fp.numel):
            self._side = torch.cuda.Stream(device=self.device)
            model._encoder_grad_hook = self._on_decoder_grads_final
        self.num_updates = 0
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.lr_dev = torch.full((1,), lr, dtype=torch.float32, device=self.device)
        self.gnorm_sq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._graph_sig = None
        self._static_samples = None
        self._static_out = None
        RT.enable_device_seed(self.device)

    def _on_decoder_grads_final(self, grad):
        if self._side is None or not self._last_micro or self._tail_launched:
            return None
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            self.bucketer.all_reduce_sum(self.fp.split, self.fp.numel)
        self._tail_launched = True
        return None

    # ------------------------------------------------------------------ the update, as a sequence of device work
    def _update(self, samples):
        self.fp.grads.zero_()
        losses, stats = [], []
        self._tail_launched = False
        for k, sample in enumerate(samples):  # --update-freq micro-batches
            self._last_micro = k == len(samples) - 1
            loss, sample_size, logging_output = self.task.train_step(sample, self.model, self.criterion, None,
                                                                     self.num_updates)
            losses.append(loss)
            stats.append(logging_output.get("_stats"))
        # exchange: sum over ranks (mean folded into grad_mul below). The decoder-side range may already be in flight
        # on the side stream (launched from the encoder_out gradient hook); join it, then reduce the rest.
        if self._tail_launched:
            torch.cuda.current_stream().wait_stream(self._side)
            self.bucketer.all_reduce_sum(0, self.fp.split)
        else:
            self.bucketer.all_reduce_sum()
        self._last_micro = False
        # legacy_ddp.py:110 divides by world before the sum; trainer.py:796 multiply_grads(world / sample_size) with
        # sample_size = world * n_micro (every micro-batch on every rank reports 1). Net factor on the summed gradient:
        grad_mul = 1.0 / float(self.world * len(samples))
        self.gnorm_sq.zero_()
        K.sumsq(self.fp.grads, self.gnorm_sq)
        self.step_dev += 1
        K.adam_step(self.fp.flat, self.fp.grads, self.fp.exp_avg, self.fp.exp_avg_sq, self.fp.shadow, self.lr,
                    self.betas[0], self.betas[1], self.eps, self.weight_decay, 1, self.gnorm_sq, self.clip_norm,
                    grad_mul, lr_dev=self.lr_dev, step_dev=self.step_dev)
        RT.advance_seed()
        return torch.stack(losses), (torch.stack(stats) if stats[0] is not None else None)

    def _signature(self, samples):
        shapes = tuple(tuple((k, tuple(v.shape)) for k, v in _flatten(s).items()) for s in samples)
        gates = tuple(getattr(m, "freeze_encoder_updates", 0) <= self.num_updates for m in self.model.modules()
                      if hasattr(m, "freeze_encoder_updates"))
        gates += tuple(getattr(m, "freeze_decoder_updates", 0) <= self.num_updates for m in self.model.modules()
                       if hasattr(m, "freeze_decoder_updates"))
        return shapes, gates, self.model.training

    def train_step(self, samples, lr=None):
        """samples: list of micro-batch dicts (tensors on host -- pinned for async copies -- or on the device).
        Returns (losses [n_micro] device tensor, stats [n_micro, 7] device tensor or None)."""
        if lr is not None:
            self.lr = lr
            self.lr_dev.fill_(lr)
        self.criterion.text_to_speech_loss.defer_logging = True
        if not self.use_cuda_graph:
            dev_samples = [_to_device(s, self.device) for s in samples]
            out = self._update(dev_samples)
            self.num_updates += 1
            return out
        sig = self._signature(samples)
        if self._graph is None or sig != self._graph_sig:
            self._capture(samples, sig)
        for st, s in zip(self._static_samples, samples):
            _copy_into(st, s)
        self._graph.replay()
        self.num_updates += 1
        return self._static_out

    def _capture(self, samples, sig):
        enc_ld = getattr(self.model.args, "encoder_layerdrop", 0)
        dec_ld = getattr(self.model.args, "decoder_layerdrop", 0)
        if self.model.training and (enc_ld > 0 or dec_ld > 0):
            raise RuntimeError("CUDA-graph capture needs LayerDrop 0 (host-side RNG decides the layer set); the "
                               "reference TTS recipe uses --encoder-layerdrop 0.0 --decoder-layerdrop 0.0")
        self._static_samples = [_to_device(s, self.device) for s in samples]
        # warm-up outside capture (lazy inits, allocator), on a side stream as torch requires. These are real updates
        # on the first batch; state is restored afterwards so capture does not change the training trajectory.
        snap = [t.clone() for t in (self.fp.flat, self.fp.exp_avg, self.fp.exp_avg_sq, self.step_dev)]
        bufs = [b.clone() for b in self.model.buffers()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._update(self._static_samples)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_out = self._update(self._static_samples)
        torch.cuda.synchronize()
        with torch.no_grad():
            for t, s in zip((self.fp.flat, self.fp.exp_avg, self.fp.exp_avg_sq, self.step_dev), snap):
                t.copy_(s)
            for b, s in zip(self.model.buffers(), bufs):
                b.copy_(s)
        self.fp.refresh_shadow()
        self._graph_sig = sig


def _flatten(sample, prefix=""):
    out = {}
    for k, v in sample.items():
        if isinstance(v, dict):
            out.update(_flatten(v, prefix + k + "."))
        elif torch.is_tensor(v):
            out[prefix + k] = v
    return out


def _to_device(sample, dev):
    if torch.is_tensor(sample):
        return sample.to(dev, non_blocking=True)
    if isinstance(sample, dict):
        return {k: _to_device(v, dev) for k, v in sample.items()}
    return sample


def _copy_into(static, new):
    for k, v in new.items():
        if isinstance(v, dict):
            _copy_into(static[k], v)
        elif torch.is_tensor(v):
            static[k].copy_(v, non_blocking=True)


def h2d_bytes(sample):
    return sum(v.numel() * v.element_size() for v in _flatten(sample).values())
