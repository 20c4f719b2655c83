"""-m gpu: end-to-end parity of the CUDA path (through the C ABI) against the CPU oracle and the golden fixture."""
import os

import pytest
import torch

from helpers import NO_DROPOUT, TINY, load_golden, rel, to_device

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "tts_tiny.npz")
MEL_TOL = 1e-3  # BASELINE.json north_star: mel L2 within 1e-3 relative


def _build(dev, dtype, **over):
    from speecht5_b200.models import T5TransformerModel, make_args
    from speecht5_b200.ops import RT
    RT.dtype = dtype
    RT.manual_seed(1)
    RT.invalidate_shadows()
    args = make_args("t5_transformer_base_asr", **over)
    return T5TransformerModel.build_model(args).to(dev)


def _criterion():
    from speecht5_b200.criterions import TexttoSpeechLoss
    return TexttoSpeechLoss(None, use_guided_attn_loss=True)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, MEL_TOL), (torch.bfloat16, 6e-2)])
def test_golden_fixture_forward_backward(cuda, dtype, tol):
    state, sample, out_ref, loss_ref, grads_ref = load_golden(GOLDEN)
    model = _build(cuda, dtype, **TINY, **NO_DROPOUT, bert_init=True).train()
    model.load_state_dict(state)
    s = to_device(sample, cuda)
    crit = _criterion()
    before, after, logits, attn = model(**s["net_input"])
    assert before.shape == out_ref["before"].shape and len(attn) == 2
    assert rel(after, out_ref["after"]) < tol
    assert rel(before, out_ref["before"]) < tol
    assert rel(logits, out_ref["logits"]) < tol * 3
    assert rel(torch.stack(attn), out_ref["attn"]) < tol * 3
    loss, l1, l2, bce, ga = crit.compute_loss(model, (before, after, logits, attn), s)
    got = torch.stack([loss, l1, l2, bce, ga]).detach().cpu().double()
    assert ((got - loss_ref).abs() / loss_ref.abs()).max().item() < tol * 3
    loss.backward()
    params = dict(model.named_parameters())
    gtol = 2e-3 if dtype == torch.float32 else 0.25
    for name, g in grads_ref.items():
        assert params[name].grad is not None, name
        assert rel(params[name].grad, g) < gtol, name


@pytest.mark.parametrize("dtype,tol", [(torch.float32, MEL_TOL), (torch.bfloat16, 6e-2)])
def test_base_dims_against_oracle(cuda, dtype, tol):
    """SpeechT5-Base widths (d=768, 12 heads, ffn 3072, RPE +-160), 2+2 layers, ragged batch, training-mode BN."""
    from oracle.speecht5_oracle import T5TransformerModelOracle, base_args, synthetic_tts_batch, tts_loss
    over = dict(encoder_layers=2, decoder_layers=2, bert_init=True, **NO_DROPOUT)
    torch.manual_seed(7)
    oracle = T5TransformerModelOracle(base_args(**over)).train()
    with torch.no_grad():
        for n, p in oracle.named_parameters():
            if n.endswith("alpha"):
                p.fill_(0.9)
            elif "q_proj.weight" in n or "k_proj.weight" in n or "pe_k" in n:
                p.mul_(6.0)  # peaky attention => well-conditioned softmax gradients for an fp32-vs-fp32 comparison
    sample = synthetic_tts_batch(3, 45, 64, seed=3)
    state0 = {k: v.clone() for k, v in oracle.state_dict().items()}  # before the oracle's own BN statistics update
    out_ref = oracle(**sample["net_input"])
    loss_ref = tts_loss(out_ref, sample)[0]
    loss_ref.backward()
    model = _build(cuda, dtype, **over).train()
    model.load_state_dict(state0)
    s = to_device(sample, cuda)
    before, after, logits, attn = model(**s["net_input"])
    assert rel(after, out_ref[1]) < tol, "mel (after postnet) L2"
    assert rel(before, out_ref[0]) < tol
    loss = _criterion().compute_loss(model, (before, after, logits, attn), s)[0]
    assert abs(loss.item() - loss_ref.item()) / abs(loss_ref.item()) < tol * 3
    loss.backward()
    ref_grads = dict(oracle.named_parameters())
    gmax = max(float(p.grad.norm()) for p in oracle.parameters() if p.grad is not None)
    worst, worst_name = 0.0, None
    for n, p in model.named_parameters():
        g_ref = ref_grads[n].grad
        if g_ref is None:
            continue
        assert p.grad is not None, n
        if float(g_ref.norm()) < 1e-4 * gmax:  # analytically-zero gradients (e.g. k_proj.bias): only roundoff noise
            assert float(p.grad.float().norm()) < 2e-3 * gmax, n
            continue
        e = rel(p.grad, g_ref)
        if e > worst:
            worst, worst_name = e, n
    assert worst < (5e-3 if dtype == torch.float32 else 0.25), (worst, worst_name)
    # BatchNorm running statistics follow the reference (local batch statistics, padded frames included)
    bn = "speech_decoder_postnet.postnet.postnet.0.1.running_var"
    assert rel(model.state_dict()[bn], oracle.state_dict()[bn]) < 1e-2


def test_eval_mode_matches_oracle(cuda):
    from oracle.speecht5_oracle import T5TransformerModelOracle, base_args, synthetic_tts_batch
    over = dict(encoder_layers=1, decoder_layers=1, dprenet_dropout_rate=0.0)
    torch.manual_seed(3)
    oracle = T5TransformerModelOracle(base_args(**over)).eval()
    sample = synthetic_tts_batch(2, 30, 40, seed=5)
    with torch.no_grad():
        ref = oracle(**sample["net_input"])
    model = _build(cuda, torch.float32, **over).eval()
    model.load_state_dict(oracle.state_dict())
    with torch.no_grad():
        got = model(**to_device(sample, cuda)["net_input"])
    assert rel(got[1], ref[1]) < MEL_TOL


def test_dropout_training_step_runs_and_is_reproducible(cuda):
    from oracle.speecht5_oracle import synthetic_tts_batch
    over = dict(encoder_layers=1, decoder_layers=2, encoder_layerdrop=0.0, decoder_layerdrop=0.0)
    sample = to_device(synthetic_tts_batch(2, 30, 40, seed=5), cuda)
    losses = []
    for _ in range(2):
        torch.manual_seed(0)
        model = _build(cuda, torch.bfloat16, **over).train()
        loss = _criterion().compute_loss(model, model(**sample["net_input"]), sample)[0]
        loss.backward()
        assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
        losses.append(loss.item())
    assert losses[0] == losses[1]


@pytest.mark.parametrize("dtype,tol", [(torch.float32, MEL_TOL), (torch.bfloat16, 6e-2)])
def test_generate_speech_matches_oracle(cuda, dtype, tol):
    """models/speecht5.py:1188-1249 greedy synthesis: (a) the reference's quirk -- kwargs["threshold"] also sets the
    length ratios, so threshold=0.5 runs exactly int(T * 0.5 / r) steps; (b) defaults with decisive stop logits."""
    from oracle.speecht5_oracle import T5TransformerModelOracle, base_args
    over = dict(encoder_layers=2, decoder_layers=2, dprenet_dropout_rate=0.0, bert_init=True)
    torch.manual_seed(11)
    oracle = T5TransformerModelOracle(base_args(**over)).eval()
    with torch.no_grad():
        for n, p in oracle.named_parameters():
            if n.endswith("alpha"):
                p.fill_(1.1)
            if "prob_out.weight" in n:
                p.mul_(20.0)  # decisive stop logits: no sigmoid(x) ~ 0.5 ties between the implementations
    model = _build(cuda, dtype, **over).eval()
    model.load_state_dict(oracle.state_dict())
    tok = torch.randint(4, 81, (1, 12))
    spk = torch.randn(1, 512)
    for kw in (dict(threshold=0.5), dict()):
        if not kw:  # (b): stop early for sure -- a large positive bias on the second frame's stop logit
            with torch.no_grad():
                oracle.speech_decoder_postnet.prob_out.bias[1] = 50.0
            model.load_state_dict(oracle.state_dict())
        mel_ref, probs_ref, attn_ref = oracle.generate_speech(src_tokens=tok, spkembs=spk, **kw)
        mel, probs, attn = model.generate_speech(src_tokens=tok.to(cuda), spkembs=spk.to(cuda), **kw)
        assert mel.shape == mel_ref.shape and probs.shape == probs_ref.shape and attn.shape == attn_ref.shape
        assert rel(mel, mel_ref) < tol
        assert rel(attn, attn_ref) < max(tol, 1e-3)
    assert mel_ref.shape[0] == 2  # case (b) stopped on the first step


@pytest.mark.parametrize("dtype,tol", [(torch.float32, MEL_TOL), (torch.bfloat16, 8e-2)])
def test_large_style_pre_layernorm_against_oracle(cuda, dtype, tol):
    """t5_transformer_large structure (models/speecht5.py:1402-1425): pre-LN encoder layers with norm_k on the
    relative-position table (transformer_layer.py:90-111), pre-LN decoder layers + final decoder LayerNorm
    (decoder_normalize_before), at reduced width; forward, loss and gradients incl. norm_k and the table."""
    from oracle.speecht5_oracle import T5TransformerModelOracle, base_args, synthetic_tts_batch, tts_loss
    over = dict(encoder_layers=2, decoder_layers=2, encoder_embed_dim=128, encoder_ffn_embed_dim=256,
                encoder_attention_heads=2, decoder_embed_dim=128, decoder_ffn_embed_dim=256, decoder_attention_heads=2,
                layer_norm_first=True, decoder_normalize_before=True, bert_init=True, **NO_DROPOUT)
    torch.manual_seed(13)
    oracle = T5TransformerModelOracle(base_args(**over)).train()
    with torch.no_grad():
        for n, p in oracle.named_parameters():
            if n.endswith("alpha"):
                p.fill_(0.9)
            elif "q_proj.weight" in n or "k_proj.weight" in n or "pe_k" in n:
                p.mul_(6.0)
            elif "norm_k" in n or "layer_norm" in n:
                p.add_(torch.randn_like(p) * 0.1)
    sample = synthetic_tts_batch(3, 37, 50, seed=4)
    state0 = {k: v.clone() for k, v in oracle.state_dict().items()}
    out_ref = oracle(**sample["net_input"])
    loss_ref = tts_loss(out_ref, sample)[0]
    loss_ref.backward()
    model = _build(cuda, dtype, **over).train()
    model.load_state_dict(state0)
    s = to_device(sample, cuda)
    before, after, logits, attn = model(**s["net_input"])
    assert rel(after, out_ref[1]) < tol and rel(before, out_ref[0]) < tol
    loss = _criterion().compute_loss(model, (before, after, logits, attn), s)[0]
    assert abs(loss.item() - loss_ref.item()) / abs(loss_ref.item()) < tol * 3
    loss.backward()
    ref = dict(oracle.named_parameters())
    got = dict(model.named_parameters())
    gtol = 5e-3 if dtype == torch.float32 else 0.25
    # (norm_k.bias shifts every logit of a row by q.b -- softmax is invariant to it: analytically zero gradient)
    gmax = max(float(p.grad.norm()) for p in oracle.parameters() if p.grad is not None)
    assert float(got["encoder.layers.1.norm_k.bias"].grad.float().norm()) < 2e-3 * gmax
    for n in ("encoder.layers.0.norm_k.weight", "encoder.layers.1.norm_k.weight", "encoder.pos_emb.pe_k.weight",
              "encoder.layers.0.self_attn_layer_norm.weight", "decoder.layer_norm.weight",
              "decoder.layers.1.fc1.weight", "encoder.layers.1.fc2.weight"):
        assert got[n].grad is not None, n
        assert rel(got[n].grad, ref[n].grad) < gtol, (n, rel(got[n].grad, ref[n].grad))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, MEL_TOL), (torch.bfloat16, 6e-2)])
def test_text_to_text_branch_against_oracle(cuda, dtype, tol):
    """SURVEY 8a rows 9 + 14 on the CUDA path (opt-in --build-text-decoder): text decoder prenet (embedding + fairseq
    sinusoidal positions), decoder, tied vocabulary projection; logits, label-smoothed CE and gradients vs the oracle."""
    from oracle.speecht5_oracle_asr import T5TransformerModelT2TOracle, base_asr_args, label_smoothed_nll_loss
    import torch.nn.functional as F
    over = dict(encoder_layers=2, decoder_layers=2, share_input_output_embed=True, bert_init=True, **NO_DROPOUT)
    torch.manual_seed(21)
    oracle = T5TransformerModelT2TOracle(base_asr_args(**over)).train()
    with torch.no_grad():
        for n, p in oracle.named_parameters():
            if n.endswith("alpha"):
                p.fill_(0.9)
            elif "q_proj.weight" in n or "k_proj.weight" in n or "pe_k" in n:
                p.mul_(6.0)
    g = torch.Generator().manual_seed(5)
    B, Ts, Tt, V, pad, eos = 3, 19, 14, 81, 1, 2
    src = torch.randint(4, V, (B, Ts), generator=g)
    src[1, 15:] = pad
    tgt = torch.randint(4, V, (B, Tt), generator=g)
    tgt[:, -1] = eos
    tgt[2, 9] = eos
    tgt[2, 10:] = pad
    prev = torch.full_like(tgt, pad)
    prev[:, 0] = eos
    prev[:, 1:] = tgt[:, :-1]
    prev[2, 10:] = pad

    def loss_of(logits):
        lp = F.log_softmax(logits.float(), dim=-1)
        return label_smoothed_nll_loss(lp.view(-1, V), tgt.to(logits.device).view(-1), 0.1, pad)[0]

    (logits_ref, _), _, _ = oracle(src_tokens=src, prev_output_tokens=prev)
    loss_ref = loss_of(logits_ref)
    loss_ref.backward()
    model = _build(cuda, dtype, build_text_decoder=True, **over).train()
    model.load_state_dict(oracle.state_dict())
    (logits, _), codebook_out, enc_out = model(src_tokens=src.to(cuda), prev_output_tokens=prev.to(cuda))
    keep = tgt.ne(pad)
    assert codebook_out == {} and logits.shape == logits_ref.shape
    assert rel(logits.cpu()[keep], logits_ref[keep]) < tol
    loss = loss_of(logits)
    assert abs(loss.item() - loss_ref.item()) / abs(loss_ref.item()) < tol * 3
    loss.backward()
    ref, got = dict(oracle.named_parameters()), dict(model.named_parameters())
    gtol = 5e-3 if dtype == torch.float32 else 0.25
    # the embedding is shared by the text encoder prenet, the text decoder prenet and the output projection
    e_got, e_ref = model.text_decoder_prenet.embed_tokens.weight, oracle.text_decoder_prenet.embed_tokens.weight
    assert model.text_decoder_postnet.output_projection.weight is e_got
    assert rel(e_got.grad, e_ref.grad) < gtol, rel(e_got.grad, e_ref.grad)
    for n in ("decoder.layers.0.self_attn.v_proj.weight",
              "decoder.layers.1.encoder_attn.q_proj.weight", "encoder.layers.1.fc1.weight",
              "text_encoder_prenet.encoder_prenet.1.alpha"):
        assert got[n].grad is not None, n
        assert rel(got[n].grad, ref[n].grad) < gtol, (n, rel(got[n].grad, ref[n].grad))


def test_trainer_graph_replay_and_prefetch_equal_the_eager_update(cuda):
    """B200Trainer on the device (bf16, dropout off): three updates as replays of the captured graph -- once with the
    batches copied by train_step itself, once through the input pipeline (`prefetch`: copy stream + staging buffers +
    device->device into the graph's inputs) -- give the parameters of three eager updates, and the losses agree."""
    from speecht5_b200.criterions import SpeechT5Criterion
    from speecht5_b200.data import synthetic_tts_batch
    from speecht5_b200.models import make_args
    from speecht5_b200.ops import RT
    from speecht5_b200.tasks import SpeechT5Task
    from speecht5_b200.trainer import B200Trainer
    host = [synthetic_tts_batch(3, 24, 40, seed=10 + i, pin=True) for i in range(2)]
    order = [0, 1, 0]
    results = []
    for mode in ("eager", "graph", "graph+prefetch"):
        RT.dtype = torch.bfloat16
        RT.manual_seed(1)
        RT.clear_static()
        RT.invalidate_shadows()
        torch.manual_seed(0)
        args = make_args("t5_transformer_base_asr", **TINY, **NO_DROPOUT, bert_init=True)
        task = SpeechT5Task(args)
        model = task.build_model(args).to(cuda).train()
        tr = B200Trainer(model, SpeechT5Criterion(task, use_guided_attn_loss=True), task, lr=1e-3,
                         use_cuda_graph=mode != "eager")
        losses = []
        for k, i in enumerate(order):
            out = tr.train_step([host[i]])
            if mode == "graph+prefetch" and k + 1 < len(order):
                tr.prefetch([host[order[k + 1]]])
            losses.append(float(out[0][0]))
        torch.cuda.synchronize()
        results.append((losses, tr.fp.flat.clone()))
        assert tr.graph_misses == (0 if mode == "eager" else 1)
    (l0, p0), (l1, p1), (l2, p2) = results
    for a, b, c in zip(l0, l1, l2):
        assert abs(a - b) < 2e-3 * abs(a) and abs(b - c) < 1e-5 * abs(b), (l0, l1, l2)
    assert rel(p1, p0) < 1e-3 and rel(p2, p1) < 1e-5
    RT.clear_static()
    RT.invalidate_shadows()
