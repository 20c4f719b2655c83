"""CPU: host-side mirror of the reference plugin surface (registry names, signatures, checkpoint keys, arch presets),
the synthetic collater contract, and the data-parallel gradient exchange over gloo (world_size 2)."""
import inspect
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_registry_names_and_forward_signature():
    import speecht5_b200  # noqa: F401
    from speecht5_b200 import fairseq_shim as fs
    from speecht5_b200.models.speecht5 import T5TransformerModel
    if not fs.HAVE_FAIRSEQ:
        assert fs.MODEL_REGISTRY["t5_transformer"] is T5TransformerModel
        for arch in ("t5_transformer", "t5_transformer_base", "t5_transformer_large", "t5_transformer_base_asr"):
            assert arch in fs.ARCH_CONFIG_REGISTRY
        assert "speecht5" in fs.TASK_REGISTRY and "speecht5" in fs.CRITERION_REGISTRY
    sig = list(inspect.signature(T5TransformerModel.forward).parameters)
    assert sig == ["self", "source", "src_tokens", "src_lengths", "prev_output_tokens", "tgt_lengths", "spkembs",
                   "target_list", "task_name", "padding_mask", "only_hubert", "only_ctc", "feature_only",
                   "tgt_enc_layer", "mask",
                   "mask_indices", "mask_channel_indices"]  # reference signature + two optional extras (trailing, keyword)  # models/speecht5.py:786


def test_arch_presets_match_reference_defaults():
    from speecht5_b200.models import make_args
    a = make_args("t5_transformer_base_asr")
    assert (a.encoder_layers, a.decoder_layers, a.encoder_embed_dim, a.encoder_attention_heads) == (12, 6, 768, 12)
    assert a.relative_position_embedding and a.encoder_max_relative_position == 160 and a.reduction_factor == 2
    assert (a.dropout, a.activation_dropout, a.attention_dropout, a.encoder_layerdrop) == (0.1, 0.1, 0.1, 0.1)
    lg = make_args("t5_transformer_large")
    assert (lg.encoder_layers, lg.decoder_layers, lg.encoder_embed_dim, lg.encoder_attention_heads) == (24, 6, 1024, 16)
    assert lg.layer_norm_first and lg.decoder_normalize_before
    assert lg.extractor_mode == "layer_norm" and lg.final_dim == 768 and lg.use_conv_pos and lg.use_sinc_pos  # :1402-1425
    assert make_args("t5_transformer_base").use_conv_pos and a.extractor_mode == "default"


def test_state_dict_keys_equal_oracle_and_reference_names():
    from oracle.speecht5_oracle import T5TransformerModelOracle, base_args
    from speecht5_b200.models import T5TransformerModel, make_args
    m = T5TransformerModel.build_model(make_args("t5_transformer_base_asr", encoder_layers=1, decoder_layers=1))
    o = T5TransformerModelOracle(base_args(encoder_layers=1, decoder_layers=1))
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in o.state_dict().items()}
    for key in ("encoder.layers.0.self_attn.q_proj.weight", "encoder.pos_emb.pe_k.weight",
                "speech_decoder_postnet.postnet.postnet.0.1.running_mean",
                "speech_decoder_prenet.decoder_prenet.0.0.prenet.1.0.bias",
                "text_encoder_prenet.encoder_prenet.1.alpha", "decoder.layers.0.encoder_attn.k_proj.weight"):
        assert key in m.state_dict()  # SURVEY.md section 5 checkpoint-key contract


def test_unbuilt_branches_fail_loudly_not_silently():
    from speecht5_b200.models import T5TransformerModel, make_args
    m = T5TransformerModel.build_model(make_args("t5_transformer_base_asr", encoder_layers=1, decoder_layers=1))
    with pytest.raises(NotImplementedError):
        m(source=torch.zeros(1, 1600), padding_mask=torch.zeros(1, 1600, dtype=torch.bool),
          prev_output_tokens=torch.zeros(1, 3, dtype=torch.long), task_name="s2t")


def test_synthetic_batch_follows_collater_contract():
    from speecht5_b200.data import synthetic_tts_batch
    s = synthetic_tts_batch(4, 160, 626, seed=3)
    ni = s["net_input"]
    assert set(ni) == {"src_tokens", "src_lengths", "prev_output_tokens", "tgt_lengths", "spkembs", "task_name"}
    assert ni["prev_output_tokens"].shape == (4, 313, 80) and (ni["prev_output_tokens"][:, 0] == 0).all()
    assert torch.equal(ni["prev_output_tokens"][:, 1:], s["dec_target"][:, 1::2][:, :-1])
    for b in range(4):
        L = int(s["dec_target_lengths"][b])
        assert s["labels"][b, L - 1:].min() == 1 and s["labels"][b, :L - 1].max() == 0
        assert (ni["src_tokens"][b, int(ni["src_lengths"][b]):] == 1).all()
    from oracle.speecht5_oracle import synthetic_tts_batch as oracle_batch
    o = oracle_batch(4, 160, 626, seed=3)
    assert torch.equal(o["net_input"]["src_tokens"], ni["src_tokens"]) and torch.equal(o["labels"], s["labels"])


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from speecht5_b200.trainer import GradBucketer
    torch.manual_seed(rank)
    flat = torch.randn(1000)
    mine = flat.clone()
    GradBucketer(flat, bucket_elems=256).all_reduce_mean()
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ref = torch.stack(gathered).mean(0)
    err = float((flat - ref).abs().max())
    # the trainer's variant: two ranges (decoder side first, as the overlapped exchange issues them), summed, with the
    # 1/world folded into the gradient multiplier
    flat2 = mine.clone()
    b = GradBucketer(flat2, bucket_elems=256)
    b.all_reduce_sum(600, 1000)
    b.all_reduce_sum(0, 600)
    err = max(err, float((flat2 / world - ref).abs().max()))
    q.put((rank, err))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gradient_exchange_is_mean_over_ranks_gloo(world):
    """legacy_distributed_data_parallel.py:76-165 semantics: grads /= world, all-reduce(sum); bucketed; world_size 2 and
    4 (the scaling run goes to 8 ranks with the same code)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29511 + (os.getpid() * 7 + world) % 500
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert len(res) == world and all(err < 1e-6 for _, err in res), res


def test_model_exposes_the_reference_model_api():
    """SURVEY 8(b) Python face: the callers' methods exist with the reference behaviour; unbuilt branches raise
    NotImplementedError instead of falling back."""
    import pytest
    from speecht5_b200.models import T5TransformerModel
    for m in ("build_model", "forward", "set_num_updates", "load_state_dict", "max_positions", "forward_text_encoder",
              "generate_speech", "get_normalized_probs", "get_normalized_probs_for_ctc", "get_logits", "get_targets",
              "get_extra_losses", "forward_encoder", "forward_decoder"):
        assert callable(getattr(T5TransformerModel, m)), m
    dummy = T5TransformerModel.__new__(T5TransformerModel)
    logits = torch.randn(2, 5, 11)
    lp = T5TransformerModel.get_normalized_probs(dummy, (logits, None), log_probs=True)
    assert lp.batch_first and torch.allclose(lp.exp().sum(-1), torch.ones(2, 5), atol=1e-5)
    ctc = T5TransformerModel.get_normalized_probs_for_ctc(dummy, {"encoder_out_for_ctc": [logits]}, log_probs=False)
    assert torch.allclose(ctc.sum(-1), torch.ones(2, 5), atol=1e-5)
    assert T5TransformerModel.get_targets(dummy, {"target": 7}, {}) == 7
    losses, names = T5TransformerModel.get_extra_losses(dummy, {"features_pen": torch.tensor(2.0)})
    assert names == ["features_pen"] and float(losses[0]) == 2.0
    with pytest.raises(NotImplementedError):
        T5TransformerModel.forward_encoder(dummy, torch.zeros(1, 16000))
    with pytest.raises(NotImplementedError):
        T5TransformerModel.forward_decoder(dummy, None, None, None)


def test_model_add_args_parses_the_recipe_command_line():
    """models/speecht5.py:118-700 option surface: the TTS fine-tune recipe's model flags (SpeechT5/README.md) parse into
    the namespace the arch function completes; unset options are left to the arch defaults."""
    import argparse
    from speecht5_b200.models import T5TransformerModel
    from speecht5_b200.models.speecht5 import t5_transformer_base_asr
    parser = argparse.ArgumentParser()
    T5TransformerModel.add_args(parser)
    args = parser.parse_args("--share-input-output-embed --bert-init --relative-position-embedding "
                             "--encoder-layers 2 --decoder-layers 1 --relu-dropout 0.05 --dropout 0.2".split())
    assert args.share_input_output_embed and args.bert_init and args.relative_position_embedding
    assert args.encoder_layers == 2 and args.activation_dropout == 0.05 and not hasattr(args, "decoder_ffn_embed_dim")
    t5_transformer_base_asr(args)
    assert args.decoder_ffn_embed_dim == 3072 and args.encoder_max_relative_position == 160 and args.dropout == 0.2
    model = T5TransformerModel.build_model(args)
    assert len(model.encoder.layers) == 2 and len(model.decoder.layers) == 1
    assert model.text_encoder_prenet.encoder_prenet[0].weight.shape == (81, 768)


def test_task_add_args_parses_the_recipe_command_line():
    """tasks/speecht5.py:44-213 option surface (TTS fine-tune recipe, SpeechT5/README.md)."""
    import argparse
    from speecht5_b200.tasks import SpeechT5Task
    parser = argparse.ArgumentParser()
    SpeechT5Task.add_args(parser)
    args = parser.parse_args("/data/root --config-yaml config.yaml --t5-task t2s --max-speech-positions 1876 "
                             "--max-text-positions 600 --sample-rate 16000 --bpe-tokenizer spm_char.model".split())
    assert args.data == "/data/root" and args.t5_task == "t2s" and args.max_speech_positions == 1876
    assert args.sample_rate == 16000.0 and args.hubert_labels == ["km"] and args.ctc_weight == 0.0
    task = SpeechT5Task.setup_task(args)
    assert task.t5_task == "t2s"


def test_collate_tts_reproduces_the_reference_batch_contract():
    """text_to_speech_dataset.py:226-281: un-collating a synthetic batch and collating it again gives it back, and
    the derived fields follow the collater's rules (every r-th frame shifted right, stop labels from the last frame)."""
    from speecht5_b200.data import collate_tts, synthetic_tts_batch
    ref = synthetic_tts_batch(4, 23, 37, seed=9)
    items = []
    for b in range(4):
        n_txt, n_mel = int(ref["src_lengths"][b]), int(ref["dec_target_lengths"][b])
        items.append({"id": b, "source": [ref["net_input"]["src_tokens"][b, :n_txt]],
                      "target": ref["dec_target"][b, :n_mel], "spkembs": ref["net_input"]["spkembs"][b],
                      "audio_name": f"utt{b}"})
    got = collate_tts(items, reduction_factor=2)
    for k in ("src_tokens", "src_lengths", "prev_output_tokens", "tgt_lengths", "spkembs"):
        assert torch.equal(got["net_input"][k], ref["net_input"][k]), k
    for k in ("labels", "dec_target", "dec_target_lengths", "src_lengths", "target"):
        assert torch.equal(got[k], ref[k]), k
    assert got["ntokens"] == ref["ntokens"] and got["task_name"] == "t2s" and got["name"][2] == "utt2"
    assert collate_tts([{"source": None}]) == {}


def test_compute_mask_indices_properties_and_determinism():
    """fairseq/data/data_utils.py:393-517 (static spans, as the SpeechT5 recipes use them): every row ends up with the
    same number of masked frames, none on padding, spans of the requested length, reproducible from np.random.seed."""
    import numpy as np
    from speecht5_b200.data import compute_mask_indices
    B, T, L = 4, 499, 10
    lens = torch.tensor([499, 450, 400, 320])
    pad = torch.arange(T)[None, :] >= lens[:, None]
    np.random.seed(3)
    m1 = compute_mask_indices((B, T), pad, 0.75, L, "static", 0, min_masks=2)
    np.random.seed(3)
    m2 = compute_mask_indices((B, T), pad, 0.75, L, "static", 0, min_masks=2)
    assert m1.dtype == bool and m1.shape == (B, T) and (m1 == m2).all()
    counts = m1.sum(1)
    assert (counts == counts[0]).all() and counts[0] > 0
    assert not (m1 & pad.numpy()).any()
    # without thinning (single row) every masked run is a union of length-L spans
    np.random.seed(5)
    m = compute_mask_indices((1, 200), None, 0.3, L, "static", 0, min_masks=2)[0]
    runs, n = [], 0
    for v in list(m) + [False]:
        if v:
            n += 1
        elif n:
            runs.append(n)
            n = 0
    assert runs and all(r >= L for r in runs)
    # channel masks: no padding mask -> one shared span count
    np.random.seed(7)
    mc = compute_mask_indices((3, 768), None, 0.5, 64, "static", 0)
    assert (mc.sum(1) == mc.sum(1)[0]).all()


def test_collate_asr_follows_the_reference_collater():
    """speech_to_text_dataset.py:150-222: padding mask, eos-terminated targets, eos-first decoder inputs."""
    from speecht5_b200.data import collate_asr
    items = [{"id": 0, "source": torch.arange(1.0, 6.0), "label_list": [torch.tensor([7, 8, 9])]},
             {"id": 1, "source": torch.arange(1.0, 4.0), "label_list": [torch.tensor([5])]}]
    b = collate_asr(items)
    assert b["net_input"]["source"].tolist() == [[1, 2, 3, 4, 5], [1, 2, 3, 0, 0]]
    assert b["net_input"]["padding_mask"].tolist() == [[False] * 5, [False, False, False, True, True]]
    assert b["target"].tolist() == [[7, 8, 9, 2], [5, 2, 1, 1]]
    assert b["net_input"]["prev_output_tokens"].tolist() == [[2, 7, 8, 9], [2, 5, 1, 1]]
    assert b["target_lengths"].tolist() == [4, 2] and b["ntokens"] == 4 and b["task_name"] == "s2t"


def test_speech_to_text_criterion_matches_the_oracle_loss():
    """criterions/speech_to_text_loss.py restated (label-smoothed CE + CTC with target_lengths - 1, the unscaled
    single-term case, logging keys, greedy-CTC unit errors at eval) against oracle asr_loss on a stub model."""
    import torch.nn.functional as F
    from oracle.speecht5_oracle_asr import asr_loss, synthetic_asr_batch
    from speecht5_b200.criterions import SpeechT5Criterion, SpeechtoTextLoss
    torch.manual_seed(2)
    B, Td, Te, V = 3, 9, 40, 81
    s = synthetic_asr_batch(B, 4000, Td, seed=5)
    s["task_name"] = "s2t"
    dec_logits = torch.randn(B, Td, V, requires_grad=True)
    ctc_logits = torch.randn(Te, B, V, requires_grad=True)
    pm = torch.zeros(B, Te, dtype=torch.bool)
    pm[1, 30:] = True

    class Stub(torch.nn.Module):
        def forward(self, **kw):
            assert kw["task_name"] == "s2t" and "only_ctc" not in kw
            return (dec_logits, None), {"encoder_out_for_ctc": [ctc_logits], "encoder_padding_mask": [pm]}

        def get_normalized_probs(self, net_output, log_probs, sample=None):
            out = F.log_softmax(net_output[0].float(), dim=-1)
            out.batch_first = True
            return out

        def get_normalized_probs_for_ctc(self, net_output, log_probs):
            return F.log_softmax(net_output["encoder_out_for_ctc"][0].float(), dim=-1)

        def get_targets(self, sample, net_output):
            return sample["target"]

    m = Stub().train()
    want, ce, ctc, ss = asr_loss(m, s, ce_weight=0.5, ctc_weight=0.5, label_smoothing=0.1)
    crit = SpeechT5Criterion(None, label_smoothing=0.1, ce_weight=0.5, ctc_weight=0.5, report_accuracy=True)
    loss, sample_size, log = crit(m, s)
    assert sample_size == ss == B and abs(loss.item() - want.item()) < 1e-4 * abs(want.item())
    assert abs(log["ce_loss"] - ce.item()) < 1e-3 and abs(log["ctc_loss"] - ctc.item()) < 1e-3
    assert set(log) >= {"loss", "ce_loss", "ctc_loss", "nll_loss", "ntokens", "nsentences", "sample_size", "n_correct",
                        "total"} and log["total"] == int(s["target"].ne(1).sum())
    loss.backward()
    assert dec_logits.grad.abs().sum() > 0 and ctc_logits.grad.abs().sum() > 0
    # one term only: NOT multiplied by its weight (speech_to_text_loss.py:202-205); eval adds the unit error counts
    only = SpeechtoTextLoss(None, label_smoothing=0.1, ce_weight=0.0, ctc_weight=0.3)
    l2, _, log2 = only(m.eval(), s)
    assert abs(l2.item() - ctc.item()) < 1e-3 and s.get("only_ctc") is True
    assert log2["c_total"] == int(((s["target"] != 1) & (s["target"] != 2)).sum()) and 0 < log2["c_errors"]
    from speecht5_b200.criterions.speech_to_text_loss import _edit_distance
    assert _edit_distance([1, 2, 3, 4], [1, 3, 4, 5]) == 2 and _edit_distance([], [1, 2]) == 2


def test_model_builds_the_opt_in_speech_input_branch():
    """--build-speech-encoder / --build-text-decoder construct the waveform front end under the reference's parameter
    names, and the s2t dispatch reaches it (CPU tensors are refused by the kernels, not by a missing branch)."""
    import pytest
    from speecht5_b200.models import T5TransformerModel, make_args
    args = make_args("t5_transformer_base_asr", encoder_layers=1, decoder_layers=1, build_speech_encoder=True,
                     build_text_decoder=True, use_conv_pos=True, use_sinc_pos=True)
    model = T5TransformerModel.build_model(args)
    keys = set(model.state_dict())
    for k in ("speech_encoder_prenet.feature_extractor.conv_layers.0.0.weight",
              "speech_encoder_prenet.pos_conv.0.weight_g", "speech_encoder_prenet.mask_emb",
              "speech_encoder_prenet.post_extract_proj.weight", "text_decoder_prenet.embed_tokens.weight",
              "encoder.proj.weight"):
        assert k in keys, k
    with pytest.raises(RuntimeError, match="CUDA"):
        model(source=torch.zeros(1, 4000), padding_mask=torch.zeros(1, 4000, dtype=torch.bool),
              prev_output_tokens=torch.full((1, 3), 2), task_name="s2t")
    plain = T5TransformerModel.build_model(make_args("t5_transformer_base_asr", encoder_layers=1, decoder_layers=1))
    assert plain.speech_encoder_prenet is None
    with pytest.raises(NotImplementedError):
        plain(source=torch.zeros(1, 4000), padding_mask=torch.zeros(1, 4000, dtype=torch.bool),
              prev_output_tokens=torch.full((1, 3), 2), task_name="s2t")


def test_bias_gradient_is_handed_over_to_the_consuming_layer_norm(monkeypatch):
    """Post-LN tail y = LN(r + W a + b): the bias gradient of the projection (and of fc2 inside ops.ffn) is the column
    sum of the LayerNorm's dx, which st5_ln_bwd produces in its own pass (dxsum) -- no column-sum launch. Checked on the
    emulated kernels against plain autograd, with and without a trainer-style static gradient buffer, and that the
    projection still sums by itself when nobody consumes the hand-over."""
    import gemm_emulator
    from speecht5_b200 import kernels as K, ops
    from speecht5_b200.ops import RT
    gemm_emulator.install(monkeypatch)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    RT.invalidate_shadows()
    calls = []
    real_colsum = K.colsum
    monkeypatch.setattr(K, "colsum", lambda *a, **k: (calls.append(1), real_colsum(*a, **k))[1])
    torch.manual_seed(0)
    d, f = 16, 32
    proj, fc1, fc2 = torch.nn.Linear(d, d), torch.nn.Linear(d, f), torch.nn.Linear(f, d)
    ln1, ln2 = torch.nn.LayerNorm(d), torch.nn.LayerNorm(d)
    a = torch.randn(2, 5, d)
    r = torch.randn(2, 5, d)

    def ours(fold):
        for m in (proj, fc1, fc2, ln1, ln2):
            m.zero_grad()
        o = ops.linear(a, proj.weight, proj.bias, bias_grad_by_consumer=fold)
        x = ops.residual_layer_norm(o, r, ln1)
        o2 = ops.ffn(x, fc1, fc2, "relu", bias_grad_by_consumer=fold)
        y = ops.residual_layer_norm(o2, x, ln2)
        (y * torch.arange(d).float()).sum().backward()
        return {n: p.grad.clone() for n, p in (("pb", proj.bias), ("b2", fc2.bias), ("b1", fc1.bias), ("pw", proj.weight),
                                               ("g1", ln1.weight))}

    for m in (proj, fc1, fc2, ln1, ln2):
        m.zero_grad()
    F = torch.nn.functional
    x = F.layer_norm(r + F.linear(a, proj.weight, proj.bias), (d,), ln1.weight, ln1.bias)
    y = F.layer_norm(x + F.linear(F.relu(F.linear(x, fc1.weight, fc1.bias)), fc2.weight, fc2.bias), (d,), ln2.weight, ln2.bias)
    (y * torch.arange(d).float()).sum().backward()
    want = dict(pb=proj.bias.grad.clone(), b2=fc2.bias.grad.clone(), b1=fc1.bias.grad.clone(), pw=proj.weight.grad.clone(),
                g1=ln1.weight.grad.clone())
    n_plain = None
    for fold in (False, True):
        del calls[:]
        got = ours(fold)
        for k in want:
            assert torch.allclose(got[k], want[k], rtol=2e-2, atol=2e-2), (fold, k)  # (bf16 operands in the emulated GEMM)
        assert torch.allclose(got["pb"], want["pb"], rtol=1e-4, atol=1e-4) or not fold  # the folded sums are fp32-exact
        if not fold:
            n_plain = len(calls)
        else:
            assert len(calls) == n_plain - 2, (len(calls), n_plain)  # proj.bias and fc2.bias no longer launch a column sum
    # static gradient buffers (B200Trainer's flat buffer): the LayerNorm kernel accumulates straight into them
    gpb, gb2 = torch.zeros(d), torch.zeros(d)
    monkeypatch.setitem(RT._static_grad, ("bias", id(proj.bias)), gpb)
    monkeypatch.setitem(RT._static_grad, ("bias", id(fc2.bias)), gb2)
    for m in (proj, fc1, fc2, ln1, ln2):
        m.zero_grad()
    o = ops.linear(a, proj.weight, proj.bias, bias_grad_by_consumer=True)
    xx = ops.residual_layer_norm(o, r, ln1)
    o2 = ops.ffn(xx, fc1, fc2, "relu", bias_grad_by_consumer=True)
    yy = ops.residual_layer_norm(o2, xx, ln2)
    (yy * torch.arange(d).float()).sum().backward()
    assert proj.bias.grad is None and fc2.bias.grad is None
    assert torch.allclose(gpb, want["pb"], rtol=2e-2, atol=2e-2) and torch.allclose(gb2, want["b2"], rtol=2e-2, atol=2e-2)
    # nobody consumes the hand-over: the projection sums its own columns
    monkeypatch.delitem(RT._static_grad, ("bias", id(proj.bias)))
    proj.zero_grad()
    o = ops.linear(a, proj.weight, proj.bias, bias_grad_by_consumer=True)
    (o * r).sum().backward()
    assert torch.allclose(proj.bias.grad, r.sum((0, 1)), rtol=1e-5, atol=1e-5)
    RT.invalidate_shadows()


def test_fused_attention_function_leaves_no_reference_cycle(monkeypatch):
    """ops.AttentionTCFn (fused / streaming path) with the three kernel entry points stubbed: the output tensor reaches
    the backward through save_for_backward (same storage as the forward's), and once the caller drops its references
    the saved exponentials die by reference counting -- no object is left for Python's cycle collector (as a ctx
    attribute the output closed a cycle that kept ~12 GB per eager update of the Large pre-training step alive)."""
    import gc
    import weakref
    from speecht5_b200 import kernels as K
    from speecht5_b200 import ops
    seen = {}

    def fwd(a, lse, psave=None, inv_l=None, out_f32=None):
        seen["out"], seen["psave"] = a.out, weakref.ref(psave)
        psave.zero_(), inv_l.fill_(1.0), out_f32.zero_()

    def bwd(a, psave, inv_l, out_f32, delta, dq_acc, ext_heads=0):
        seen["bwd_out"], seen["bwd_psave"] = a.out, psave.data_ptr()

    monkeypatch.setattr(K, "attn_fused_fwd", fwd)
    monkeypatch.setattr(K, "attn_flash_fwd", fwd)
    monkeypatch.setattr(K, "attn_fused_bwd", bwd)
    B, T, H, d = 2, 24, 2, 128
    cfg = dict(H=H, d=d, q_col=0, k_col=1, v_col=2, scale=0.125, maxpos=0, causal=True, drop_p=0.0, return_probs=False,
               probs_grad_heads=0, probs_read_heads=0)
    gc.collect()
    gc.disable()
    try:
        before = len(gc.get_objects())
        q = torch.zeros(B, T, 3 * d, dtype=torch.bfloat16, requires_grad=True)
        out, probs = ops.AttentionTCFn.apply(q, None, None, None, cfg)
        assert probs is None and out.shape == (B, T, d)
        psave_ptr = seen["psave"]().data_ptr()
        out.float().sum().backward()
        assert seen["bwd_out"] == seen["out"] == out.data_ptr() and seen["bwd_psave"] == psave_ptr
        # a second pass whose graph is dropped WITHOUT running backward (the unmasked-frame head of the pre-training
        # criterion with weight 0, inference under enable_grad, ...)
        out2, _ = ops.AttentionTCFn.apply(q, None, None, None, cfg)
        ref = seen["psave"]
        assert ref() is not None
        del out, out2
        assert ref() is None, "saved exponentials must die with the last reference to the output"
        assert gc.collect() == 0
    finally:
        gc.enable()
    del before
