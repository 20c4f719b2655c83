"""CPU: the host-side composition of the waveform front end (speecht5_b200/frontend.py) -- which operand views, row
pitches, batch strides and phase offsets it hands to st5_gemm_bf16 -- run through the GEMM emulator and compared with
torch's Conv1d + GELU autograd. (The kernels themselves are covered by the -m gpu tests.)"""
import pytest
import torch
import torch.nn.functional as F

import gemm_emulator


@pytest.mark.parametrize("k,s,T", [(3, 2, 41), (2, 2, 37), (3, 2, 40), (5, 3, 50), (4, 2, 23)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_strided_conv_gelu_composition(monkeypatch, k, s, T, dtype):
    from speecht5_b200 import frontend
    from speecht5_b200.ops import RT
    gemm_emulator.install(monkeypatch)
    monkeypatch.setattr(RT, "dtype", dtype)
    RT.invalidate_shadows()
    torch.manual_seed(k * 100 + s * 10 + T)
    B, Cin, Cout = 3, 16, 24
    x = (torch.randn(B, T, Cin) * 0.7).to(dtype).requires_grad_()
    w = torch.nn.Parameter(torch.randn(Cout, Cin, k) * 0.3)
    y = frontend.StridedConvGeluFn.apply(x, w, s)
    dy = torch.randn(y.shape).to(dtype)
    y.backward(dy)
    xr = x.detach().double().requires_grad_()
    wr = w.detach().double().requires_grad_()
    yr = F.gelu(F.conv1d(xr.transpose(1, 2), wr, stride=s)).transpose(1, 2)
    yr.backward(dy.double())
    assert y.shape == yr.shape == (B, (T - k) // s + 1, Cout)

    def rel(a, b):
        return ((a.double() - b).norm() / b.norm()).item()
    tol = 3e-5 if dtype == torch.float32 else 2e-2
    assert rel(y, yr) < tol
    assert rel(x.grad, xr.grad) < tol
    assert rel(w.grad, wr.grad) < tol
    RT.invalidate_shadows()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("T", [19, 64])
def test_grouped_positional_conv_composition(monkeypatch, dtype, T):
    """x + GELU(SamePad(grouped Conv1d(k even)) + bias) with a weight-normed weight: forward, dx, and the gradients
    that flow through the torch-side weight norm to g and v (speech_encoder_prenet.py:105-119,187-192)."""
    from speecht5_b200 import frontend
    from speecht5_b200.ops import RT
    gemm_emulator.install(monkeypatch)
    monkeypatch.setattr(RT, "dtype", dtype)
    torch.manual_seed(T)
    B, Cc, G, k = 2, 32, 4, 8
    x = (torch.randn(B, T, Cc) * 0.8).to(dtype).requires_grad_()
    v = torch.nn.Parameter(torch.randn(Cc, Cc // G, k) * 0.2)
    gpar = torch.nn.Parameter(v.detach().norm(dim=(0, 1), keepdim=True) * 1.1)
    bias = torch.nn.Parameter(torch.randn(Cc) * 0.1)
    w = gpar * v / v.norm(dim=(0, 1), keepdim=True)
    y = frontend.GroupedPosConvFn.apply(x, w, bias, G)
    dy = torch.randn(y.shape).to(dtype)
    y.backward(dy)
    xr = x.detach().double().requires_grad_()
    vr, gr, br = (t.detach().double().requires_grad_() for t in (v, gpar, bias))
    wr = gr * vr / vr.norm(dim=(0, 1), keepdim=True)
    pos = F.conv1d(xr.transpose(1, 2), wr, br, padding=k // 2, groups=G)[:, :, :-1]
    yr = xr + F.gelu(pos).transpose(1, 2)
    yr.backward(dy.double())

    def rel(a, b):
        return ((a.double() - b).norm() / b.norm()).item()
    tol = 3e-5 if dtype == torch.float32 else 2e-2
    assert rel(y, yr) < tol
    assert rel(x.grad, xr.grad) < tol
    assert rel(v.grad, vr.grad) < tol
    assert rel(gpar.grad, gr.grad) < tol
    assert rel(bias.grad, br.grad) < tol


def test_padding_mask_downsampling_and_positions_match_the_oracle():
    """Host logic of the speech prenet: frame mask = all-samples-padded after trimming the remainder
    (speech_encoder_prenet.py:219-229) and the positions the reference derives from the BOOLEAN mask (:196-198)."""
    from oracle import speecht5_oracle_asr as O
    from speecht5_b200 import frontend
    from speecht5_b200.models.modules.nets import fairseq_sinusoid_table
    torch.manual_seed(0)
    B, n, T = 3, 3217, 10
    lengths = torch.tensor([3217, 2000, 655])
    pm = torch.arange(n)[None, :] >= lengths[:, None]
    args = O.base_asr_args(encoder_layers=1, decoder_layers=1)
    want = O.SpeechEncoderPrenet(args).forward_padding_mask(torch.zeros(B, T, 4), pm)
    got = frontend.downsample_padding_mask(pm, T)
    assert torch.equal(got, want) and got.any() and not got.all()
    emb = O.SinusoidalPositionalEmbedding(16, 1)(got)
    table = fairseq_sinusoid_table(2 + T, 16, 1, "cpu")
    mine = table.index_select(0, frontend.padding_mask_positions(got, 1).view(-1)).view(B, T, -1)
    assert torch.allclose(mine, emb.float(), atol=1e-6)
    assert (mine[got] == 0).all()


def test_speech_prenet_state_dict_uses_the_reference_names():
    from oracle import speecht5_oracle_asr as O
    from speecht5_b200 import frontend
    args = O.base_asr_args(encoder_layers=1, decoder_layers=1)
    for k, v in dict(conv_feature_layers="[(512,10,5)] + [(512,3,2)]*4 + [(512,2,2)]*2", encoder_speech_prenet="conv",
                     mask_prob=0.65, hubert_mask_length=10, mask_selection="static", mask_other=0.0,
                     no_mask_overlap=False, mask_min_space=1, mask_channel_prob=0.0, freeze_encoder_updates=0).items():
        if not hasattr(args, k) or k == "conv_feature_layers":
            setattr(args, k, v)
    m = frontend.SpeechEncoderPrenet(args)
    keys = set(m.state_dict().keys())
    for name in ("feature_extractor.conv_layers.0.0.weight", "feature_extractor.conv_layers.0.2.weight",
                 "feature_extractor.conv_layers.0.2.bias", "feature_extractor.conv_layers.6.0.weight",
                 "post_extract_proj.weight", "post_extract_proj.bias", "layer_norm.weight", "layer_norm.bias",
                 "pos_conv.0.bias", "pos_conv.0.weight_g", "pos_conv.0.weight_v", "mask_emb"):
        assert name in keys, name
    assert m.state_dict()["pos_conv.0.weight_g"].shape == (1, 1, args.conv_pos)
    assert m.state_dict()["pos_conv.0.weight_v"].shape == (768, 768 // args.conv_pos_groups, args.conv_pos)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4000))  # CPU tensors: the product path refuses instead of falling back


def test_logmel_device_composition_matches_the_oracle(monkeypatch):
    """speecht5_b200/audio.py through the GEMM emulator: overlapping-frame STFT GEMM (row pitch = hop), magnitude, mel
    GEMM, log10 -- against oracle.audio_oracle.logmelfilterbank (librosa semantics), batched, n not a multiple of hop."""
    import numpy as np
    from oracle.audio_oracle import logmelfilterbank as ref_fn, mel_basis
    from speecht5_b200 import audio
    gemm_emulator.install(monkeypatch)
    audio._CONST.clear()
    assert np.allclose(audio.slaney_mel_basis(16000, 1024, 80, 80, 7600).numpy(), mel_basis(), atol=1e-7)
    rng = np.random.default_rng(3)
    n = 5000
    t = np.arange(n) / 16000.0
    waves = np.stack([0.3 * np.sin(2 * np.pi * 300 * t) + 0.02 * rng.standard_normal(n),
                      0.1 * np.sin(2 * np.pi * 2500 * t) + 0.05 * rng.standard_normal(n)]).astype(np.float32)
    got = audio.logmelfilterbank(torch.from_numpy(waves)).numpy()
    want = np.stack([ref_fn(w) for w in waves])
    assert got.shape == want.shape == (2, 1 + n // 256, 80)
    assert np.abs(got - want).max() < 1e-3, np.abs(got - want).max()
    one = audio.logmelfilterbank(torch.from_numpy(waves[0])).numpy()
    assert np.abs(one - want[0]).max() < 1e-3
    audio._CONST.clear()


def test_hifigan_device_composition_matches_the_oracle(monkeypatch):
    """speecht5_b200/vocoder.py through the GEMM emulator (bf16 activations) against oracle HifiGanGenerator (fp32):
    'same' convolutions, the stride-phase transposed convolutions, the de-interleaved dilated convolutions, ResBlock
    residuals, averaging and the final tanh -- reduced configuration, T not a multiple of the dilations."""
    from oracle.audio_oracle import HifiGanGenerator as Ref
    from speecht5_b200 import vocoder
    gemm_emulator.install(monkeypatch)
    cfg = dict(model_in_dim=16, upsample_initial_channel=32, upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4],
               resblock_kernel_sizes=[3, 7], resblock_dilation_sizes=[[1, 3], [1, 5]])
    torch.manual_seed(0)
    ref = Ref(cfg, std=0.15, seed=1).eval()
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if n.endswith("bias"):
                p.add_(0.05 * torch.randn_like(p))
        ref.mean.copy_(torch.randn(16) * 0.1)
        ref.scale.copy_(1.0 + 0.1 * torch.rand(16))
    gen = vocoder.HifiGanGenerator(ref.state_dict(), cfg, device="cpu")
    mel = torch.randn(2, 13, 16)
    with torch.no_grad():
        want = ref(mel)
    got = gen(mel)
    assert got.shape == want.shape == (2, 13 * 8)
    err = ((got.double() - want.double()).norm() / want.double().norm()).item()
    assert err < 3e-2, err


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_speech_encoder_prenet_module_forward_against_the_oracle(monkeypatch, dtype):
    """The whole speech prenet module (frontend.SpeechEncoderPrenet, eval mode) with every kernel entry point emulated
    on the CPU: conv stack, LayerNorm, projection, injected mask draw, positional conv, sinusoidal positions, frame
    padding mask and the feature penalty -- against oracle SpeechEncoderPrenet with the same weights."""
    from oracle import speecht5_oracle_asr as O
    from speecht5_b200 import frontend
    from speecht5_b200.ops import RT
    gemm_emulator.install(monkeypatch)
    monkeypatch.setattr(RT, "dtype", dtype)
    monkeypatch.setattr(frontend.ConvFeatureExtractor, "forward", _cpu_extractor_forward)
    RT.invalidate_shadows()
    torch.manual_seed(1)
    args = O.base_asr_args(encoder_layers=1, decoder_layers=1, dropout=0.0)
    for k, v in dict(encoder_speech_prenet="conv", mask_prob=0.0, hubert_mask_length=10, mask_selection="static",
                     mask_other=0.0, no_mask_overlap=False, mask_min_space=1, mask_channel_prob=0.0,
                     freeze_encoder_updates=0).items():
        setattr(args, k, v)
    args.conv_feature_layers = list(O.CONV_FEATURE_LAYERS)  # a list suits both (the CLI hands the product a string)
    ref = O.SpeechEncoderPrenet(args).double().eval()
    mine = frontend.SpeechEncoderPrenet(args).eval()
    sd = {k: v.float() for k, v in ref.state_dict().items()}
    for a, b in (("pos_conv_g", "pos_conv.0.weight_g"), ("pos_conv_v", "pos_conv.0.weight_v"),
                 ("pos_conv_bias", "pos_conv.0.bias")):
        sd[b] = sd.pop(a)
    missing, unexpected = mine.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    B, n = 2, 4000
    wave = torch.randn(B, n, dtype=torch.float64) * 0.3
    pm = torch.arange(n)[None, :] >= torch.tensor([4000, 2600])[:, None]
    T = int(ref.feature_extractor.get_out_seq_lens_tensor(torch.tensor([n]))[0])
    mi = torch.zeros(B, T, dtype=torch.bool)
    mi[0, 2:6] = True
    mi[1, 1:3] = True
    with torch.no_grad():
        xr, mr, pen_r = ref(wave, pm, mask_indices=mi)
        (x, pen, got_mi, _), m = mine(wave.float(), require_feat_pen=True, padding_mask=pm, mask=True, mask_indices=mi)
    assert torch.equal(m, mr) and got_mi is mi
    err = ((x.double() - xr).norm() / xr.norm()).item()
    assert err < (2e-4 if dtype == torch.float32 else 3e-2), err
    assert abs(pen.item() - pen_r.item()) / pen_r.item() < (1e-4 if dtype == torch.float32 else 2e-2)
    RT.invalidate_shadows()


@pytest.mark.parametrize("dims", ["[(32, 10, 5)] + [(32, 3, 2)] * 2 + [(32, 2, 2)]", "[(64, 10, 5), (48, 3, 2), (64, 2, 2)]"])
def test_layer_norm_extractor_composition_matches_the_oracle(monkeypatch, dims):
    """ConvFeatureExtractor(mode="layer_norm") (t5_transformer_large): layer-0 Function, window GEMM without epilogue
    activation + row LayerNorm + stand-alone GELU for the rest, forward and every gradient against the oracle's
    ConvFeatureExtractionModel (itself pinned to the reference's in both modes, tests/test_ref_pin_cpu.py); parameter
    names are the reference checkpoint's (conv_layers.{i}.2.1.*)."""
    from helpers import rel
    from oracle.speecht5_oracle_asr import ConvFeatureExtractionModel
    from speecht5_b200 import frontend
    from speecht5_b200.ops import RT
    gemm_emulator.install_autograd(monkeypatch)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    monkeypatch.setattr(frontend.ConvFeatureExtractor, "forward", _cpu_extractor_forward)
    RT.invalidate_shadows()
    torch.manual_seed(3)
    layers = eval(dims)
    ref = ConvFeatureExtractionModel(layers, "layer_norm", False).double()
    for blk in ref.conv_layers:
        torch.nn.init.normal_(blk[2].weight, 1.0, 0.2), torch.nn.init.normal_(blk[2].bias, 0.0, 0.2)
    mine = frontend.ConvFeatureExtractor(layers, "layer_norm", False)
    import re
    to_ref = lambda k: re.sub(r"^(conv_layers\.\d+\.2)\.", r"\1.1.", k)  # noqa: E731
    sd = {to_ref(k): v.float() for k, v in ref.state_dict().items()}
    assert set(sd) == set(mine.state_dict())
    mine.load_state_dict(sd)
    wave = torch.randn(2, 700, dtype=torch.float64) * 0.3
    want = ref(wave).transpose(1, 2)
    got = mine(wave.float())
    assert got.shape == want.shape and rel(got, want) < 1e-5
    probe = torch.randn_like(want)
    (want * probe).sum().backward()
    (got * probe.float()).sum().backward()
    named = dict(mine.named_parameters())
    for k, p in ref.named_parameters():
        g = named[to_ref(k)].grad
        assert g is not None and rel(g, p.grad) < 2e-5, (k, rel(g, p.grad))
    RT.invalidate_shadows()


def _cpu_extractor_forward(self, wave):
    """ConvFeatureExtractor.forward without its CUDA guard (the kernels underneath are emulated in these tests)."""
    return self._layers(wave)


def test_speech_to_text_model_forward_and_greedy_decoding_on_emulated_kernels(monkeypatch):
    """The opt-in s2t branch of T5TransformerModel.forward, forward_encoder / forward_decoder and generate_text_greedy
    with every kernel entry point emulated on the CPU (fp32 parity arithmetic): logits and CTC head against the ASR
    oracle, and beam-1 token ids against oracle greedy_decode."""
    from oracle import speecht5_oracle_asr as O
    from speecht5_b200 import frontend
    from speecht5_b200.models import T5TransformerModel, make_args
    from speecht5_b200.ops import RT
    gemm_emulator.install(monkeypatch)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    monkeypatch.setattr(frontend.ConvFeatureExtractor, "forward", _cpu_extractor_forward)
    RT.invalidate_shadows()
    torch.manual_seed(6)
    over = dict(encoder_layers=2, decoder_layers=2, bert_init=True, mask_prob=0.0, mask_channel_prob=0.0)
    oracle = O.T5TransformerModelASROracle(O.base_asr_args(**over)).eval()
    with torch.no_grad():
        oracle.text_decoder_postnet.output_projection.weight.mul_(8.0)
    args = make_args("t5_transformer_base_asr", build_speech_encoder=True, build_text_decoder=True, use_conv_pos=True,
                     use_sinc_pos=True, **over)
    model = T5TransformerModel.build_model(args).eval()
    sd = dict(oracle.state_dict())
    for a, b in (("pos_conv_g", "pos_conv.0.weight_g"), ("pos_conv_v", "pos_conv.0.weight_v"),
                 ("pos_conv_bias", "pos_conv.0.bias")):
        sd["speech_encoder_prenet." + b] = sd.pop("speech_encoder_prenet." + a)
    model.load_state_dict(sd)
    s = O.synthetic_asr_batch(2, 6000, 6, seed=9)
    with torch.no_grad():
        (want, _), enc_ref = oracle(**s["net_input"])
        (got, _), enc = model(**s["net_input"])
    keep = s["target"].ne(1)
    assert ((got[keep].double() - want[keep].double()).norm() / want[keep].double().norm()).item() < 2e-4
    ctc_ref, ctc = enc_ref["encoder_out_for_ctc"][0], enc["encoder_out_for_ctc"][0]
    valid = ~enc_ref["encoder_padding_mask"][0].t()  # [T, B]
    assert ((ctc[valid].double() - ctc_ref[valid].double()).norm() / ctc_ref[valid].double().norm()).item() < 2e-4
    assert torch.equal(enc["encoder_padding_mask"][0], enc_ref["encoder_padding_mask"][0])
    src, pm = s["net_input"]["source"], s["net_input"]["padding_mask"]
    ids_ref = O.greedy_decode(oracle, src, pm, max_len_b=10)
    ids = model.generate_text_greedy(src, pm, max_len_b=10)
    assert [t.tolist() for t in ids] == [t.tolist() for t in ids_ref]
    RT.invalidate_shadows()


def test_text_to_speech_forward_on_emulated_kernels_matches_the_golden_fixture(monkeypatch):
    """Regression guard for the DEFAULT path without a GPU: T5TransformerModel.forward (t2s) with every kernel entry
    point emulated on the CPU reproduces the committed golden outputs (tests/golden/tts_tiny.npz, training-mode
    BatchNorm, no dropout) -- the model-level Python is exercised here, the kernels in the -m gpu tests."""
    import os
    from helpers import NO_DROPOUT, TINY, load_golden, rel
    from speecht5_b200.models import T5TransformerModel, make_args
    from speecht5_b200.ops import RT
    gemm_emulator.install(monkeypatch)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    RT.invalidate_shadows()
    state, sample, out_ref, _, _ = load_golden(os.path.join(os.path.dirname(__file__), "golden", "tts_tiny.npz"))
    model = T5TransformerModel.build_model(make_args("t5_transformer_base_asr", **TINY, **NO_DROPOUT, bert_init=True))
    model.train()
    model.load_state_dict(state)
    with torch.no_grad():
        before, after, logits, attn = model(**sample["net_input"])
    assert rel(after, out_ref["after"]) < 1e-4 and rel(before, out_ref["before"]) < 1e-4
    assert rel(logits, out_ref["logits"]) < 3e-4 and rel(torch.stack(attn), out_ref["attn"]) < 3e-4
    RT.invalidate_shadows()


def test_kv_cache_decoding_equals_prefix_recomputation_on_emulated_kernels(monkeypatch):
    """speecht5_b200/incremental.py: (a) beam-1 text decoding with the cache gives the oracle's token ids, (b) greedy
    speech synthesis with the cache gives the same mel / stop probabilities / cross-attention rows as the
    prefix-recomputing path and as the oracle's generate_speech (prenet dropout off so the runs are comparable)."""
    from helpers import NO_DROPOUT, rel
    from oracle import speecht5_oracle as OT
    from oracle import speecht5_oracle_asr as O
    from speecht5_b200 import frontend
    from speecht5_b200.models import T5TransformerModel, make_args
    from speecht5_b200.ops import RT
    gemm_emulator.install(monkeypatch)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    monkeypatch.setattr(frontend.ConvFeatureExtractor, "forward", _cpu_extractor_forward)
    RT.invalidate_shadows()
    # (a) text decoding
    torch.manual_seed(6)
    over = dict(encoder_layers=2, decoder_layers=2, bert_init=True, mask_prob=0.0, mask_channel_prob=0.0)
    oracle = O.T5TransformerModelASROracle(O.base_asr_args(**over)).eval()
    with torch.no_grad():
        oracle.text_decoder_postnet.output_projection.weight.mul_(8.0)
    model = T5TransformerModel.build_model(make_args(
        "t5_transformer_base_asr", build_speech_encoder=True, build_text_decoder=True, use_conv_pos=True,
        use_sinc_pos=True, **over)).eval()
    sd = dict(oracle.state_dict())
    for a, b in (("pos_conv_g", "pos_conv.0.weight_g"), ("pos_conv_v", "pos_conv.0.weight_v"),
                 ("pos_conv_bias", "pos_conv.0.bias")):
        sd["speech_encoder_prenet." + b] = sd.pop("speech_encoder_prenet." + a)
    model.load_state_dict(sd)
    s = O.synthetic_asr_batch(2, 6000, 6, seed=9)
    src, pm = s["net_input"]["source"], s["net_input"]["padding_mask"]
    ids_ref = O.greedy_decode(oracle, src, pm, max_len_b=10)
    ids = model.generate_text_greedy(src, pm, max_len_b=10, use_cache=True)
    assert [t.tolist() for t in ids] == [t.tolist() for t in ids_ref]
    # the device-side form a CUDA graph replays (step counter, additive score masks, bookkeeping on tensors), run eagerly
    ids_g = model.generate_text_greedy(src, pm, max_len_b=10, use_cache="graph_body_eager")
    assert [t.tolist() for t in ids_g] == [t.tolist() for t in ids_ref]
    ids_m = model.generate_text_greedy(src, pm, max_len_b=10, min_len=4, unk_penalty=0.5, use_cache=True)
    ids_mg = model.generate_text_greedy(src, pm, max_len_b=10, min_len=4, unk_penalty=0.5, use_cache="graph_body_eager")
    assert [t.tolist() for t in ids_mg] == [t.tolist() for t in ids_m] and all(len(t) >= 4 for t in ids_mg)
    # the generator object fairseq's generate.py drives (task.build_generator / inference_step): hypotheses in the
    # SequenceGenerator's return shape, token log-probabilities equal on the three decoding paths and equal to the
    # oracle's teacher-forced log-probabilities of the same tokens
    from types import SimpleNamespace
    from speecht5_b200.dictionary import Vocabulary
    from speecht5_b200.tasks.speecht5 import SpeechT5Task
    vocab = Vocabulary()
    for i in range(model.text_decoder_postnet.output_projection.weight.shape[0] - len(vocab) - 2):
        vocab.add_symbol("s%d" % i)
    vocab.add_symbol("<mask>"), vocab.add_symbol("<ctc_blank>")
    task = SpeechT5Task.__new__(SpeechT5Task)
    task.args, task.dicts = SimpleNamespace(ctc_weight=0.0), {"text": vocab}
    task.blank_symbol_idx, task.mask_idx = 0, vocab.index("<mask>")
    gen_args = SimpleNamespace(beam=1, max_len_a=0, max_len_b=10, min_len=1, unnormalized=False, lenpen=1.0, unkpen=0.0)
    per_path = []
    for mode in (False, True, "graph_body_eager"):
        gen = task.build_generator([model], gen_args, extra_gen_cls_kwargs=dict(use_cache=mode))
        hypos = task.inference_step(gen, [model], s)
        assert [h[0]["tokens"].tolist() for h in hypos] == [t.tolist() for t in ids_ref]
        for h in hypos:
            assert len(h) == 1 and h[0]["positional_scores"].shape == h[0]["tokens"].shape
            assert abs(float(h[0]["score"]) - float(h[0]["positional_scores"].mean())) < 1e-6
        per_path.append(torch.cat([h[0]["positional_scores"] for h in hypos]))
    assert rel(per_path[1], per_path[0]) < 1e-4 and rel(per_path[2], per_path[0]) < 1e-4
    with torch.no_grad():  # teacher-forced oracle log-probabilities of the hypothesis tokens
        x, enc_pad, _ = oracle.speech_encoder_prenet(src, pm, None, None)
        enc = oracle.encoder(x, enc_pad)
        want = []
        for b, hyp in enumerate(ids_ref):
            prev = torch.cat([hyp.new_tensor([2]), hyp[:-1]])[None]
            dec_in, tgt_mask = oracle.text_decoder_prenet(prev)
            one = dict(enc, encoder_out=[enc["encoder_out"][0][:, b: b + 1]],
                       encoder_padding_mask=[enc["encoder_padding_mask"][0][b: b + 1]])
            z, _ = oracle.decoder(dec_in, tgt_mask, one, alignment_layer=None)
            lp = F.log_softmax(oracle.text_decoder_postnet(z)[0].float(), dim=-1)
            want.append(lp.gather(1, hyp[:, None])[:, 0])
    assert rel(per_path[0], torch.cat(want)) < 1e-3
    with pytest.raises(NotImplementedError):
        task.build_generator([model], SimpleNamespace(beam=5))
    # (b) speech synthesis
    RT.invalidate_shadows()
    torch.manual_seed(3)
    over = dict(encoder_layers=2, decoder_layers=2, bert_init=True, **NO_DROPOUT)
    tts_oracle = OT.T5TransformerModelOracle(OT.base_args(**over)).eval()
    with torch.no_grad():
        tts_oracle.speech_decoder_postnet.prob_out.bias.fill_(-2.0)
    tts = T5TransformerModel.build_model(make_args("t5_transformer_base_asr", **over)).eval()
    tts.load_state_dict(tts_oracle.state_dict())
    tok = torch.randint(4, 81, (1, 7))
    spk = torch.randn(1, 512)
    with torch.no_grad():
        want = tts_oracle.generate_speech(src_tokens=tok, spkembs=spk, threshold=0.9)
    plain = tts.generate_speech(src_tokens=tok, spkembs=spk, threshold=0.9)
    cached = tts.generate_speech(src_tokens=tok, spkembs=spk, threshold=0.9, use_cache=True)
    # the step body a CUDA graph replays (device-side step counter, index_copy_ cache writes, masked full-span
    # self-attention, results written at the step index), run eagerly here
    counted = tts.generate_speech(src_tokens=tok, spkembs=spk, threshold=0.9, use_cache="graph_body_eager")
    for a, b, c, d in zip(plain, cached, want, counted):
        assert a.shape == b.shape == c.shape == d.shape
        assert rel(b, a) < 1e-4 and rel(b, c) < 1e-3 and rel(d, b) < 1e-5
    RT.invalidate_shadows()


def test_speech_to_text_training_step_gradients_on_emulated_kernels(monkeypatch):
    """One s2t update (CE + CTC through the speecht5 criterion) back-propagated on the CPU: LinearFn / FFNFn / the
    front-end autograd Functions run their own backward compositions on the emulated GEMM (LayerNorm, embedding and
    attention are differentiable torch stand-ins); loss and parameter gradients against oracle asr_loss."""
    from oracle import speecht5_oracle_asr as O
    from speecht5_b200 import frontend
    from speecht5_b200.criterions import SpeechT5Criterion
    from speecht5_b200.models import T5TransformerModel, make_args
    from speecht5_b200.ops import RT
    gemm_emulator.install_autograd(monkeypatch)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    monkeypatch.setattr(frontend.ConvFeatureExtractor, "forward", _cpu_extractor_forward)
    RT.invalidate_shadows()
    torch.manual_seed(4)
    over = dict(encoder_layers=1, decoder_layers=1, bert_init=True, dropout=0.0, attention_dropout=0.0,
                activation_dropout=0.0, encoder_layerdrop=0.0, decoder_layerdrop=0.0, mask_prob=0.0,
                mask_channel_prob=0.0, feature_grad_mult=1.0)
    oracle = O.T5TransformerModelASROracle(O.base_asr_args(**over)).train()
    model = T5TransformerModel.build_model(make_args(
        "t5_transformer_base_asr", build_speech_encoder=True, build_text_decoder=True, use_conv_pos=True,
        use_sinc_pos=True, **over)).train()
    sd = dict(oracle.state_dict())
    for a, b in (("pos_conv_g", "pos_conv.0.weight_g"), ("pos_conv_v", "pos_conv.0.weight_v"),
                 ("pos_conv_bias", "pos_conv.0.bias")):
        sd["speech_encoder_prenet." + b] = sd.pop("speech_encoder_prenet." + a)
    model.load_state_dict(sd)
    s = O.synthetic_asr_batch(2, 5000, 7, seed=3)
    want, _, _, _ = O.asr_loss(oracle, s, ce_weight=0.5, ctc_weight=0.5, label_smoothing=0.1)
    want.backward()
    sample = dict(s, task_name="s2t")
    loss, _, log = SpeechT5Criterion(None, label_smoothing=0.1, ce_weight=0.5, ctc_weight=0.5)(model, sample)
    assert abs(loss.item() - want.item()) < 2e-4 * abs(want.item()), (loss.item(), want.item(), log)
    loss.backward()
    ref, got = dict(oracle.named_parameters()), dict(model.named_parameters())
    rename = {"speech_encoder_prenet.pos_conv.0.weight_g": "speech_encoder_prenet.pos_conv_g",
              "speech_encoder_prenet.pos_conv.0.weight_v": "speech_encoder_prenet.pos_conv_v",
              "speech_encoder_prenet.pos_conv.0.bias": "speech_encoder_prenet.pos_conv_bias"}
    checked = 0
    for name, p in got.items():
        r = ref.get(rename.get(name, name))
        if r is None or r.grad is None or p.grad is None or float(r.grad.norm()) == 0.0:
            continue
        if name.endswith("k_proj.bias"):  # a key bias shifts every score of a row equally: its gradient is rounding noise
            continue
        err = ((p.grad.double() - r.grad.double()).norm() / r.grad.double().norm()).item()
        assert err < 2e-3, (name, err)
        checked += 1
    assert checked > 40
    for must in ("speech_encoder_prenet.feature_extractor.conv_layers.0.0.weight",
                 "speech_encoder_prenet.feature_extractor.conv_layers.4.0.weight",
                 "speech_encoder_prenet.pos_conv.0.weight_v", "speech_encoder_prenet.mask_emb", "encoder.proj.weight"):
        assert got[must].grad is not None or ref[rename.get(must, must)].grad is None, must
    RT.invalidate_shadows()


def test_text_to_speech_training_step_on_emulated_kernels_matches_the_golden_gradients(monkeypatch):
    """Default-path regression guard including the backward: the t2s update (TexttoSpeechLoss with guided attention)
    on emulated kernels reproduces the golden fixture's loss terms and parameter gradients; LinearFn, FFNFn and
    Conv1dK5Fn run their own backward compositions on the emulated GEMM."""
    import os
    import torch.nn.functional as F
    from helpers import NO_DROPOUT, TINY, load_golden, rel
    from speecht5_b200 import ops
    from speecht5_b200.criterions import TexttoSpeechLoss
    from speecht5_b200.models import T5TransformerModel, make_args
    from speecht5_b200.ops import RT
    gemm_emulator.install_autograd(monkeypatch)

    def batch_norm_act(x, bn, training, act=None, drop_p=0.0):
        assert drop_p == 0.0 and training
        y = F.batch_norm(x.float().reshape(-1, x.shape[-1]), None, None, bn.weight, bn.bias, True, 0.0, bn.eps)
        y = torch.tanh(y) if act == "tanh" else y
        return y.reshape(x.shape).to(x.dtype)
    monkeypatch.setattr(ops, "batch_norm_act", batch_norm_act)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    RT.invalidate_shadows()
    state, sample, out_ref, loss_ref, grads_ref = load_golden(
        os.path.join(os.path.dirname(__file__), "golden", "tts_tiny.npz"))
    model = T5TransformerModel.build_model(make_args("t5_transformer_base_asr", **TINY, **NO_DROPOUT, bert_init=True))
    model.train()
    model.load_state_dict(state)
    crit = TexttoSpeechLoss(None, use_guided_attn_loss=True)
    out = model(**sample["net_input"])
    loss, l1, l2, bce, ga = crit.compute_loss(model, out, sample)
    got = torch.stack([loss, l1, l2, bce, ga]).detach().double()
    assert ((got - loss_ref).abs() / loss_ref.abs()).max().item() < 3e-4
    loss.backward()
    params = dict(model.named_parameters())
    for name, g in grads_ref.items():
        assert params[name].grad is not None, name
        assert rel(params[name].grad, g) < 2e-3, name
    RT.invalidate_shadows()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pretraining_head_and_quantizer_on_emulated_kernels(monkeypatch, dtype):
    """speecht5_b200/pretrain.py (projections through ops.linear on the emulated GEMM) against oracle/pretrain_oracle.py
    with the same weights, masks, Gumbel noise and time permutation: NCE logits of the masked / unmasked frames, the
    quantizer's codes, perplexities, the mixed encoder states, and the gradients that reach the projections."""
    from oracle import pretrain_oracle as P
    from speecht5_b200 import pretrain
    from speecht5_b200.ops import RT
    gemm_emulator.install_autograd(monkeypatch)
    monkeypatch.setattr(RT, "dtype", dtype)
    RT.invalidate_shadows()
    torch.manual_seed(5)
    B, T, d = 2, 12, 64
    tol = 2e-4 if dtype == torch.float32 else 4e-2
    x = torch.randn(B, T, d)
    pm = torch.zeros(B, T, dtype=torch.bool)
    pm[1, 9:] = True
    mi = torch.rand(B, T) < 0.4
    targets = [torch.randint(0, 17, (B, T))]
    ref_h = P.SpeechEncoderPostnet([17], encoder_embed_dim=d, final_dim=32)
    mine_h = pretrain.SpeechEncoderPostnet([17], encoder_embed_dim=d, final_dim=32)
    mine_h.load_state_dict(ref_h.state_dict())
    xr = x.clone().requires_grad_()
    xm = x.to(dtype).requires_grad_()
    out_r, out_m = ref_h(xr, pm, mi, targets), mine_h(xm, pm, mi, targets)
    def fin(t):  # the class equal to the positive is -inf by construction (compute_nce): compare the finite entries
        return torch.where(torch.isfinite(t), t, torch.zeros_like(t))
    for key in ("logit_m_list", "logit_u_list"):
        a, b = out_m[key][0].float(), out_r[key][0]
        assert a.shape == b.shape and torch.equal(torch.isfinite(a), torch.isfinite(b))
        assert (torch.isinf(b).sum(1) == 1).all()
        assert ((fin(a) - fin(b)).norm() / fin(b).norm()).item() < tol
    (fin(out_r["logit_m_list"][0]).sum() + fin(out_r["logit_u_list"][0]).square().sum()).backward()
    (fin(out_m["logit_m_list"][0].float()).sum() + fin(out_m["logit_u_list"][0].float()).square().sum()).backward()
    gw_r, gw_m = ref_h.final_proj.weight.grad, mine_h.final_proj.weight.grad
    assert ((gw_m - gw_r).norm() / gw_r.norm()).item() < max(tol, 2e-3)
    # quantizer (training mode, injected noise) + code mixing
    ref_q = P.GumbelVectorQuantizer(dim=d, num_vars=10, groups=2, vq_dim=d).train()
    mine_q = pretrain.GumbelVectorQuantizer(dim=d, num_vars=10, groups=2, vq_dim=d).train()
    mine_q.load_state_dict(ref_q.state_dict())
    noise = -torch.empty(B * T * 2, 10).exponential_().log()
    qr, qm = ref_q(x, noise), mine_q(x.to(dtype), noise)
    if dtype == torch.float32:  # identical code choices, hence identical vectors
        assert ((qm["x"].float() - qr["x"]).norm() / qr["x"].norm()).item() < 1e-5
        assert abs(float(qm["prob_perplexity"]) - float(qr["prob_perplexity"])) < 1e-3
        assert float(qm["code_perplexity"]) == float(qr["code_perplexity"])
    assert qm["num_vars"] == 20 and qm["x"].shape == (B, T, d)
    perm = torch.randperm(T)
    mixed_r = P.mix_codes(x, qr["x"], 0.5, perm)
    mixed_m = pretrain.mix_codes(x, qr["x"], 0.5, perm)
    assert torch.equal(mixed_m, mixed_r)
    RT.invalidate_shadows()


def test_speech_pretraining_branch_of_forward_on_emulated_kernels(monkeypatch):
    """models/speecht5.py:813-961 with target_list (speech pre-training): prenet with label alignment and feature penalty
    -> encoder -> masked-prediction head, Gumbel quantizer + code mixing on the encoder output, speech decoder on the
    mixed states -- T5TransformerModel.forward on emulated kernels against the same composition of the oracles
    (speech prenet / encoder / decoder oracles + oracle/pretrain_oracle.py), same Gumbel noise and time permutation."""
    from oracle import pretrain_oracle as P
    from oracle import speecht5_oracle as OT
    from oracle import speecht5_oracle_asr as O
    from speecht5_b200 import frontend
    from speecht5_b200.models import T5TransformerModel, make_args
    from speecht5_b200.ops import RT
    gemm_emulator.install(monkeypatch)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    monkeypatch.setattr(frontend.ConvFeatureExtractor, "forward", _cpu_extractor_forward)
    RT.invalidate_shadows()
    torch.manual_seed(11)
    over = dict(encoder_layers=2, decoder_layers=2, bert_init=True, mask_prob=0.0, mask_channel_prob=0.0, dropout=0.0,
                attention_dropout=0.0, activation_dropout=0.0, dprenet_dropout_rate=0.0, postnet_dropout_rate=0.0,
                transformer_enc_positional_dropout_rate=0.0, transformer_dec_positional_dropout_rate=0.0,
                encoder_layerdrop=0.0, decoder_layerdrop=0.0)
    oargs = O.base_asr_args(**over)
    tts = OT.T5TransformerModelOracle(oargs).train()          # encoder, decoder, speech decoder pre / post-net
    prenet = O.SpeechEncoderPrenet(oargs).train()
    d = oargs.encoder_embed_dim
    head = P.SpeechEncoderPostnet([23], encoder_embed_dim=d, final_dim=32).train()
    quant = P.GumbelVectorQuantizer(dim=d, num_vars=10, groups=2, vq_dim=d).train()
    args = make_args("t5_transformer_base_asr", build_speech_encoder=True, use_conv_pos=True, use_sinc_pos=True,
                     use_codebook=True, latent_vars=10, latent_groups=2, codebook_prob=0.5, hubert_num_classes=[23],
                     final_dim=32, **over)
    model = T5TransformerModel.build_model(args).train()
    sd = {k: v for k, v in tts.state_dict().items() if not k.startswith("text_encoder_prenet.")}
    for k, v in prenet.state_dict().items():
        k = {"pos_conv_g": "pos_conv.0.weight_g", "pos_conv_v": "pos_conv.0.weight_v", "pos_conv_bias": "pos_conv.0.bias"}.get(k, k)
        sd["speech_encoder_prenet." + k] = v
    sd.update({"hubert_layer." + k: v for k, v in head.state_dict().items()})
    sd.update({"quantizer." + k: v for k, v in quant.state_dict().items()})
    own = model.state_dict()
    assert all(k in own for k in sd if not k.startswith("encoder.proj")), [k for k in sd if k not in own][:5]
    model.load_state_dict(sd)
    B, n = 2, 6000
    wave = torch.randn(B, n) * 0.3
    pad = torch.zeros(B, n, dtype=torch.bool)
    pad[1, 5200:] = True
    with torch.no_grad():
        x_ref, enc_pad, fpen_ref = prenet(wave, pad, None, None)
    T = x_ref.shape[1]
    labels = [torch.randint(0, 23, (B, T + 3))]                 # longer than the frames: no trimming, ratio 1
    mask_idx = torch.zeros(B, T, dtype=torch.bool)
    mask_idx[0, 2:7] = True
    mask_idx[1, 4:9] = True
    prev = torch.randn(B, 9, 80)
    tgt_lengths = torch.tensor([9, 7])
    spk = torch.randn(B, 512)
    noise = -torch.empty(B * T * 2, 10).exponential_().log()
    perm = torch.randperm(T)
    with torch.no_grad():
        x_ref, enc_pad, fpen_ref = prenet(wave, pad, mask_idx, None)
        enc = tts.encoder(x_ref, enc_pad)
        enc_btc = enc["encoder_out"][0].transpose(0, 1)
        hub_ref = head(enc_btc, enc_pad, mask_idx, [labels[0][:, :T]])
        q = quant(enc_btc, noise)
        mixed = P.mix_codes(enc_btc, q["x"], 0.5, perm)
        enc["encoder_out"] = [mixed.transpose(0, 1)]
        dec_in, tgt_mask = tts.speech_decoder_prenet(prev, tgt_lengths, spk)
        dec_out, extra = tts.decoder(dec_in, tgt_mask, enc, alignment_layer=None)
        before_ref, after_ref, logits_ref = tts.speech_decoder_postnet(dec_out)
        model._gumbel_noise, model._codebook_perm = noise, perm
        hub, (before, after, logits, attn) = model(source=wave, padding_mask=pad, prev_output_tokens=prev,
                                                   tgt_lengths=tgt_lengths, spkembs=spk, target_list=labels,
                                                   task_name="speech_pretrain", mask_indices=mask_idx)

    def close(a, b, tol=3e-4):
        a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
        fin = torch.isfinite(b)
        if not torch.equal(torch.isfinite(a), fin):  # (-inf marks a negative that equals the positive, :68-70)
            return False
        return ((a[fin] - b[fin]).norm() / b[fin].norm().clamp_min(1e-12)).item() < tol
    assert close(hub["features_pen"], fpen_ref)
    assert close(hub["logit_m_list"][0], hub_ref["logit_m_list"][0]) and close(hub["logit_u_list"][0], hub_ref["logit_u_list"][0])
    assert close(hub["prob_perplexity"], q["prob_perplexity"]) and hub["num_vars"] == q["num_vars"]
    assert close(after, after_ref) and close(before, before_ref) and close(logits, logits_ref)
    # only_hubert returns the head alone (:884-885); feature_only the encoder states (:832-833)
    with torch.no_grad():
        only, none = model(source=wave, padding_mask=pad, target_list=labels, task_name="speech_pretrain",
                           only_hubert=True, mask_indices=mask_idx)
    assert none is None and close(only["logit_m_list"][0], hub_ref["logit_m_list"][0])
    RT.invalidate_shadows()


def test_joint_pretraining_update_through_the_trainer_on_emulated_kernels(monkeypatch):
    """BASELINE config 4 in miniature: t5_transformer_large's structure (layer_norm waveform extractor, pre-LN, tied
    embeddings, masked-prediction head, shared Gumbel quantizer) at tiny widths; one update = a speech_pretrain and a
    text_pretrain micro-batch (`--update-freq 2`) through B200Trainer and the `speecht5` criterion dispatcher. The
    pre-training criteria read their statistics back inside forward, so the trainer runs these updates eagerly even
    when graphs are on; the loss falls over a few updates of the same batches and every parameter group moves."""
    import numpy as np
    from helpers import NO_DROPOUT, TINY
    from speecht5_b200 import frontend
    from speecht5_b200.criterions import SpeechT5Criterion
    from speecht5_b200.data import synthetic_speech_pretrain_batch, synthetic_text_pretrain_batch
    from speecht5_b200.models import make_args
    from speecht5_b200.ops import RT
    from speecht5_b200.tasks import SpeechT5Task
    from speecht5_b200.trainer import B200Trainer
    gemm_emulator.install_autograd(monkeypatch)
    gemm_emulator.install_trainer(monkeypatch)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    monkeypatch.setattr(frontend.ConvFeatureExtractor, "forward", _cpu_extractor_forward)
    RT.disable_device_seed()
    RT.clear_static()
    RT.invalidate_shadows()
    torch.manual_seed(5)
    np.random.seed(5)
    V = 40
    args = make_args("t5_transformer_large", **dict(TINY, **NO_DROPOUT), bert_init=True, build_speech_encoder=True,
                     build_text_decoder=True, share_input_output_embed=True, use_codebook=True, latent_vars=10,
                     latent_groups=2, codebook_prob=0.5, hubert_num_classes=[23], final_dim=16, vocab_size=V,
                     conv_feature_layers="[(32, 10, 5)] + [(32, 3, 2)] * 4 + [(32, 2, 2)] * 2", conv_pos=16,
                     conv_pos_groups=4, mask_prob=0.5, hubert_mask_length=3, max_text_positions=600)
    assert args.extractor_mode == "layer_norm" and args.layer_norm_first
    task = SpeechT5Task(args)
    model = task.build_model(args).train()
    assert model.speech_encoder_prenet.feature_extractor.mode == "layer_norm"
    crit = SpeechT5Criterion(task, loss_weights=[10.0], dec_weight=0.5, bart_weight=1.0, hubert_weight=1.0)
    tr = B200Trainer(model, crit, task, lr=2e-3, clip_norm=10.0, use_cuda_graph=True)  # (pre-training -> eager anyway)
    speech = synthetic_speech_pretrain_batch(2, 6400, n_classes=23, seed=1)
    text = synthetic_text_pretrain_batch(3, 12, V, mask_idx=V - 2, seed=2)
    assert speech["target_list"][0].shape == (2, 20) and speech["net_input"]["prev_output_tokens"].shape == (2, 13, 80)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    losses = []
    for _ in range(6):
        out, stats = tr.train_step([speech, text])
        assert stats is None and out.shape == (2,) and bool(torch.isfinite(out).all())
        losses.append(out.tolist())
    assert tr.graph_misses == 0 and tr.num_updates == 6
    assert losses[-1][0] < losses[0][0] and losses[-1][1] < losses[0][1], losses
    moved = {n for n, p in model.named_parameters() if not torch.equal(p.detach(), before[n])}
    for key in ("speech_encoder_prenet.feature_extractor.conv_layers.0.0.weight",
                "speech_encoder_prenet.feature_extractor.conv_layers.3.2.1.weight", "hubert_layer.label_embs_concat",
                "quantizer.vars", "text_encoder_prenet.encoder_prenet.0.weight", "encoder.layers.0.fc1.weight",
                "decoder.layers.0.encoder_attn.k_proj.weight", "speech_decoder_postnet.feat_out.weight"):
        assert key in moved, key
    RT.clear_static()
    RT.invalidate_shadows()


@pytest.mark.parametrize("name", ["ref_asr_tiny", "ref_asr_large_style_tiny"])
def test_generator_scores_equal_the_reference_sequence_generator_on_emulated_kernels(monkeypatch, name):
    """speecht5_b200/generator.py against the reference's OWN SequenceGenerator (sequence_generator.py:207-655, beam 1)
    run on the reference model (tests/golden/ref_asr_*.npz, produced by make_golden_from_ref.py): same weights in the
    product model on emulated kernels -> the same token ids, per-token log-probabilities and length-normalised score,
    on the prefix-recomputing, cached and graph-body decoding paths."""
    import os
    import numpy as np
    from types import SimpleNamespace
    from helpers import NO_DROPOUT, TINY, rel
    from speecht5_b200 import frontend
    from speecht5_b200.generator import GreedyGenerator
    from speecht5_b200.models import T5TransformerModel, make_args
    from speecht5_b200.ops import RT
    gemm_emulator.install(monkeypatch)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    monkeypatch.setattr(frontend.ConvFeatureExtractor, "forward", _cpu_extractor_forward)
    RT.invalidate_shadows()
    blob = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz")))
    over = dict(TINY, **NO_DROPOUT, bert_init=True, build_speech_encoder=True, build_text_decoder=True,
                conv_feature_layers="[(32, 10, 5)] + [(32, 3, 2)] * 4 + [(32, 2, 2)] * 2", feature_grad_mult=1.0,
                conv_pos=16, conv_pos_groups=4, use_conv_pos=True, use_sinc_pos=True, mask_prob=0.5,
                hubert_mask_length=4, mask_channel_prob=0.25, mask_channel_length=8, max_text_positions=600)
    if "large_style" in name:
        over.update(extractor_mode="layer_norm", layer_norm_first=True, decoder_normalize_before=True,
                    share_input_output_embed=True)
    model = T5TransformerModel.build_model(make_args("t5_transformer_base_asr", **over)).eval()
    model.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in blob.items() if k.startswith("state/")}, strict=False)
    V = 81
    vocab = SimpleNamespace(pad=lambda: 1, eos=lambda: 2, unk=lambda: 3)
    sample = {"net_input": {"source": torch.from_numpy(blob["in/source"]),
                            "padding_mask": torch.from_numpy(blob["in/padding_mask"])}}
    for mode in (False, True, "graph_body_eager"):
        gen = GreedyGenerator([model], vocab, max_len_b=12, blank=V - 1, mask_idx=V - 2, use_cache=mode)
        for b, h in enumerate(gen.generate([model], sample)):
            n = int(blob["out/greedy_lengths"][b])
            assert h[0]["tokens"].tolist() == blob["out/greedy_tokens"][b, :n].tolist(), (mode, b)
            assert rel(h[0]["positional_scores"], torch.from_numpy(blob["out/greedy_pos_scores"][b, :n])) < 1e-4, (mode, b)
            assert abs(float(h[0]["score"]) - float(blob["out/greedy_scores"][b])) < 1e-4, (mode, b)
    RT.invalidate_shadows()


def test_speech_pretraining_update_equals_the_reference_model_on_emulated_kernels(monkeypatch):
    """SURVEY 8a row 22 pinned END TO END to the reference's own code: the reference T5TransformerModel's speech
    pre-training forward (prenet + its own mask draw, encoder, masked-prediction head, Gumbel quantizer + code mixing,
    speech decoder) under the reference SpeechPretrainCriterion gives the fixture's loss, sample size, logging values
    and gradients; the product model with the same weights, mask, Gumbel noise and permutation (kernels emulated on the
    CPU, parity arithmetic) reproduces them through the `speecht5` criterion dispatcher."""
    from helpers import rel, speech_pretrain_fixture_case
    from speecht5_b200 import frontend
    from speecht5_b200.ops import RT
    gemm_emulator.install_trainer(monkeypatch)  # (install_autograd + the post-net BatchNorm as a differentiable torch call)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    monkeypatch.setattr(frontend.ConvFeatureExtractor, "forward", _cpu_extractor_forward)
    RT.clear_static()
    RT.invalidate_shadows()
    blob, model, crit, sample = speech_pretrain_fixture_case(torch.device("cpu"))
    loss, n, log = crit(model, sample)
    assert n == int(blob["loss"][1])
    assert abs(loss.item() - blob["loss"][0]) < 1e-4 * abs(blob["loss"][0]), (loss.item(), blob["loss"])
    keys = [k[4:] for k in blob if k.startswith("log/")]
    assert len(keys) >= 15
    for k in keys:
        want = float(blob["log/" + k])
        assert k in log and abs(float(log[k]) - want) <= 2e-4 * max(1.0, abs(want)), (k, log.get(k), want)
    loss.backward()
    params = dict(model.named_parameters())
    grads = [k[5:] for k in blob if k.startswith("grad/")]
    assert len(grads) >= 12
    for k in grads:
        assert params[k].grad is not None, k
        assert rel(params[k].grad, torch.from_numpy(blob["grad/" + k])) < 2e-4, (k, rel(params[k].grad, torch.from_numpy(blob["grad/" + k])))
    RT.clear_static()
    RT.invalidate_shadows()


def test_text_pretraining_update_equals_the_reference_model_on_emulated_kernels(monkeypatch):
    """The text half of a pre-training update against the reference's own code: reference T5TransformerModel (text
    pre-net, encoder, the shared Gumbel quantizer + code mixing on the text states, text decoder, tied output embedding)
    under the reference TextPretrainCriterion -> the fixture; the product model with the same weights and draws on
    emulated kernels gives its loss, sample size, logging values and gradients (ragged sources, padded targets)."""
    from helpers import rel, text_pretrain_fixture_case
    from speecht5_b200.ops import RT
    gemm_emulator.install_trainer(monkeypatch)
    monkeypatch.setattr(RT, "dtype", torch.float32)
    RT.clear_static()
    RT.invalidate_shadows()
    blob, model, crit, sample = text_pretrain_fixture_case(torch.device("cpu"))
    loss, n, log = crit(model, sample)
    assert n == int(blob["loss"][1])
    assert abs(loss.item() - blob["loss"][0]) < 1e-4 * abs(blob["loss"][0]), (loss.item(), blob["loss"])
    for k in [k[4:] for k in blob if k.startswith("log/")]:
        want = float(blob["log/" + k])
        assert k in log and abs(float(log[k]) - want) <= 2e-4 * max(1.0, abs(want)), (k, log.get(k), want)
    loss.backward()
    params = dict(model.named_parameters())
    grads = [k[5:] for k in blob if k.startswith("grad/")]
    assert len(grads) >= 10
    for k in grads:
        assert params[k].grad is not None, k
        assert rel(params[k].grad, torch.from_numpy(blob["grad/" + k])) < 2e-4, (k, rel(params[k].grad, torch.from_numpy(blob["grad/" + k])))
    RT.clear_static()
    RT.invalidate_shadows()
