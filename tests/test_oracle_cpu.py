"""CPU (-m "not gpu"): the oracle against (a) the committed golden fixture and (b) the independent HuggingFace port --
the pin that stands in for the reference's missing tests (the reference ships none for this path)."""
import os

import pytest
import torch

from helpers import NO_DROPOUT, TINY, load_golden, rel

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "tts_tiny.npz")


def test_oracle_reproduces_golden_fixture():
    from oracle.speecht5_oracle import T5TransformerModelOracle, base_args, tts_loss
    state, sample, out_ref, loss_ref, grads_ref = load_golden(GOLDEN)
    model = T5TransformerModelOracle(base_args(**TINY, **NO_DROPOUT, bert_init=True)).double().train()
    model.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in state.items()})
    ni = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in sample["net_input"].items()}
    out = model(**ni)
    assert rel(out[1], out_ref["after"]) < 1e-6 and rel(out[0], out_ref["before"]) < 1e-6
    assert rel(torch.stack(out[3]), out_ref["attn"]) < 1e-6
    s64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in sample.items()}
    terms = tts_loss(out, s64)
    got = torch.stack([t.detach() for t in terms])
    assert ((got - loss_ref).abs() / loss_ref.abs()).max() < 1e-6
    terms[0].backward()
    params = dict(model.named_parameters())
    for n, g in grads_ref.items():
        assert rel(params[n].grad, g) < 1e-5, n


def test_oracle_matches_hf_port():
    pytest.importorskip("transformers")
    from oracle.hf_crosscheck import build_hf, run_hf
    from oracle.speecht5_oracle import T5TransformerModelOracle, base_args, synthetic_tts_batch
    torch.manual_seed(0)
    n_enc, n_dec = 2, 2
    oracle = T5TransformerModelOracle(base_args(encoder_layers=n_enc, decoder_layers=n_dec, dprenet_dropout_rate=0.0,
                                                bert_init=True)).eval()
    with torch.no_grad():
        for n, p in oracle.named_parameters():
            if n.endswith("alpha"):
                p.fill_(1.1)
            elif "q_proj.weight" in n or "k_proj.weight" in n or "pe_k" in n:
                p.mul_(5.0)
        for n, b in oracle.named_buffers():
            if "running_var" in n:
                b.uniform_(0.5, 1.5)
            if "running_mean" in n:
                b.normal_(0, 0.2)
    hf = build_hf(oracle, n_enc, n_dec)
    sample = synthetic_tts_batch(3, 40, 60, seed=2)
    with torch.no_grad():
        ref = oracle(**sample["net_input"])
    before, after, logits, cross, _ = run_hf(hf, sample["net_input"])
    assert rel(ref[0], before) < 1e-5 and rel(ref[1], after) < 1e-5 and rel(ref[2], logits) < 1e-5
    for a, b in zip(ref[3], cross):
        assert rel(a, b) < 1e-5


def test_oracle_rpe_matches_reference_formulation():
    """encoder.py:239-246 + multihead_attention.py:346-353 written out literally (gather pos_k [T,T,64], T matmuls) vs
    the oracle MultiheadAttention."""
    from oracle.speecht5_oracle import MultiheadAttention, RelativePositionalEncoding
    torch.manual_seed(0)
    T, B, H, d = 11, 2, 2, 128
    mha = MultiheadAttention(d, H, self_attention=True, has_relative_attention_bias=True).double().eval()
    pos = RelativePositionalEncoding(64, 4).double()
    x = torch.randn(T, B, d, dtype=torch.float64)
    seq = torch.arange(T)
    pos_k = pos(seq[:, None] - seq[None, :])
    out, _ = mha(x, x, x, position_bias=pos_k)
    q = (mha.q_proj(x) * mha.scaling).view(T, B * H, 64).transpose(0, 1)
    k = mha.k_proj(x).view(T, B * H, 64).transpose(0, 1)
    v = mha.v_proj(x).view(T, B * H, 64).transpose(0, 1)
    s = torch.bmm(q, k.transpose(1, 2))
    idx = (seq[:, None] - seq[None, :]).clamp(-4, 3) + 4
    for i in range(T):
        for j in range(T):
            s[:, i, j] += (q[:, i] * pos.pe_k.weight[idx[i, j]]).sum(-1)
    ref = mha.out_proj(torch.bmm(torch.softmax(s, -1), v).transpose(0, 1).reshape(T, B, d))
    assert rel(out, ref) < 1e-6  # the reference (and the oracle) take the softmax in fp32


# ---------------------------------------------------------------------------------------------- speech-in / text-out
ASR_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "asr_tiny.npz")


def _load_asr_golden():
    import numpy as np
    z = np.load(ASR_GOLDEN)
    state = {k[len("state/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("state/")}
    ni = {k[len("in/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in/")}
    sample = {"net_input": ni, "target": torch.from_numpy(z["sample/target"]),
              "target_lengths": torch.from_numpy(z["sample/target_lengths"])}
    out = {k[len("out/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("out/")}
    grads = {k[len("grad/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad/")}
    return state, sample, out, torch.from_numpy(z["loss"]), grads


def test_asr_oracle_reproduces_golden_fixture():
    """SURVEY 8a rows 2, 3, 9, 14, 18 (conv front-end, speech prenet with time/channel masks, text decoder pre/post-net,
    label-smoothed CE + CTC): fixture generated by tests/golden/make_golden_asr.py."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_asr import TINY as ASR_TINY
    from oracle.speecht5_oracle_asr import T5TransformerModelASROracle, asr_loss, base_asr_args
    state, sample, out_ref, loss_ref, grads_ref = _load_asr_golden()
    model = T5TransformerModelASROracle(base_asr_args(**ASR_TINY), vocab_size=41).double().train()
    model.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in state.items()})
    sample["net_input"]["source"] = sample["net_input"]["source"].double()
    loss, ce, ctc, ss = asr_loss(model, sample)
    got = torch.stack([loss.detach(), ce.detach(), ctc.detach()])
    assert ((got - loss_ref[:3]).abs() / loss_ref[:3].abs()).max() < 1e-6 and ss == int(loss_ref[3])
    loss.backward()
    with torch.no_grad():
        (logits, _), enc = model(**sample["net_input"])
    assert rel(logits, out_ref["logits"]) < 1e-6 and rel(enc["encoder_out"][0], out_ref["encoder_out"]) < 1e-6
    assert torch.equal(enc["encoder_padding_mask"][0], out_ref["encoder_padding_mask"])
    params = dict(model.named_parameters())
    for n, g in grads_ref.items():
        assert rel(params[n].grad, g) < 1e-5, n


def test_asr_oracle_matches_hf_port():
    """Independent pin: oracle weights remapped onto transformers' SpeechT5ForSpeechToText (conv front-end with
    GroupNorm, weight-normed positional conv, sinusoidal positions, shared stacks, LM head)."""
    pytest.importorskip("transformers")
    from oracle.hf_crosscheck_asr import compare
    err = compare(n_enc=2, n_dec=2, B=2, n_samples=8000, T_tgt=11, seed=0)
    assert err["logits"] < 1e-5 and err["encoder"] < 1e-5, err


def test_asr_frame_padding_mask_and_lengths_follow_the_reference_rule():
    """speech_encoder_prenet.py:219-229: a frame is padding iff ALL samples of its (equal-sized) chunk are; the trailing
    remainder samples are dropped. And :365-374: conv output lengths."""
    from oracle.speecht5_oracle_asr import ConvFeatureExtractionModel, SpeechEncoderPrenet, base_asr_args
    fe = ConvFeatureExtractionModel()
    assert fe.get_out_seq_lens_tensor(torch.tensor([160000, 64000, 250000])).tolist() == [499, 199, 781]
    pre = SpeechEncoderPrenet(base_asr_args())
    n, T = 1000, 3  # chunks of 333 samples, one remainder sample dropped
    lens = torch.tensor([1000, 667, 666, 1])
    pm = torch.arange(n)[None, :] >= lens[:, None]
    got = pre.forward_padding_mask(torch.zeros(4, T, 1), pm)
    assert got.tolist() == [[False, False, False], [False, False, False], [False, False, True], [False, True, True]]


def test_asr_label_smoothing_matches_closed_form():
    """criterions/speech_to_text_loss.py:93-110: eps spread over the V-1 other classes."""
    from oracle.speecht5_oracle_asr import label_smoothed_nll_loss
    torch.manual_seed(0)
    V, N, eps = 7, 5, 0.1
    lp = torch.log_softmax(torch.randn(N, V, dtype=torch.float64), -1)
    tgt = torch.tensor([3, 1, 6, 1, 0])
    loss, nll = label_smoothed_nll_loss(lp, tgt, eps, ignore_index=1)
    want = 0.0
    for i in range(N):
        if int(tgt[i]) == 1:
            continue
        q = torch.full((V,), eps / (V - 1), dtype=torch.float64)
        q[tgt[i]] = 1.0 - eps
        want += float(-(q * lp[i]).sum())
    assert abs(float(loss) - want) < 1e-12


# ---------------------------------------------------------------------------------------------- audio ends (rows 15, 16)
AUDIO_GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "audio_tiny.npz")


def test_logmel_oracle_matches_torchaudio_and_hf_feature_extractor():
    """text_to_speech_dataset.py:95-138 (librosa stft + slaney mel, unvendored): two independent implementations."""
    import numpy as np
    from oracle.audio_oracle import logmelfilterbank
    rng = np.random.default_rng(0)
    wav = (rng.standard_normal(64000) * 0.1).astype(np.float32)  # 4 s -> 251 frames (1 + n // 256)
    lm = logmelfilterbank(wav)
    assert lm.shape == (251, 80)
    ta = pytest.importorskip("torchaudio")
    ms = ta.transforms.MelSpectrogram(sample_rate=16000, n_fft=1024, win_length=1024, hop_length=256, f_min=80,
                                      f_max=7600, n_mels=80, power=1.0, norm="slaney", mel_scale="slaney", center=True,
                                      pad_mode="reflect", window_fn=torch.hann_window)
    ref = torch.log10(torch.clamp(ms(torch.from_numpy(wav)), min=1e-10)).T.numpy()
    assert np.linalg.norm(lm - ref) / np.linalg.norm(ref) < 1e-5
    tr = pytest.importorskip("transformers")
    hf = tr.SpeechT5FeatureExtractor()(audio_target=wav, sampling_rate=16000, return_tensors="np")["input_values"][0]
    assert np.linalg.norm(lm - hf) / np.linalg.norm(hf) < 1e-5


def test_hifigan_oracle_matches_hf_port_and_golden():
    """SpeechUT/fairseq/.../hifigan.py:13-170 restated; pinned against transformers.SpeechT5HifiGan at the SpeechT5
    vocoder configuration, and the committed small-config fixture."""
    import numpy as np
    import sys
    from oracle.audio_oracle import HifiGanGenerator, hifigan_to_hf_state, logmelfilterbank
    tr = pytest.importorskip("transformers")
    gen = HifiGanGenerator(seed=7).eval()
    with torch.no_grad():
        gen.mean.normal_(0, 0.5)
        gen.scale.uniform_(0.5, 2.0)
        for m in gen.modules():
            if isinstance(m, (torch.nn.Conv1d, torch.nn.ConvTranspose1d)):
                m.weight.mul_(3.0)
                m.bias.normal_(0, 0.01)
    hf = tr.SpeechT5HifiGan(tr.SpeechT5HifiGanConfig()).eval()
    missing, unexpected = hf.load_state_dict(hifigan_to_hf_state(gen.state_dict()), strict=False)
    assert not missing and not unexpected
    mel = torch.randn(2, 40, 80, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        a, b = gen(mel), hf(mel)
    assert a.shape == (2, 40 * 256) and rel(a, b) < 1e-6
    # committed fixture (small configuration)
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_audio import TINY_VOCODER
    z = np.load(AUDIO_GOLDEN)
    small = HifiGanGenerator(TINY_VOCODER).double().eval()
    small.load_state_dict({k[len("voc/state/"):]: torch.from_numpy(z[k]).double() for k in z.files
                           if k.startswith("voc/state/")})
    with torch.no_grad():
        out = small(torch.from_numpy(z["voc/in"]).double())
    assert rel(out, torch.from_numpy(z["voc/out"])) < 1e-6
    assert np.allclose(logmelfilterbank(z["mel/wav"]), z["mel/logmel"], atol=1e-5)


def test_tts_generate_speech_matches_hf_port():
    """models/speecht5.py:1188-1249 (greedy frame-by-frame synthesis with the stop threshold) against the HF port's
    generate_speech, prenet dropout 0 on both sides; one run to the length cap, one that stops on the threshold."""
    pytest.importorskip("transformers")
    from oracle.hf_crosscheck import build_hf
    from oracle.speecht5_oracle import T5TransformerModelOracle, base_args
    for bias, seed in ((-1.5, 0), (0.0, 3)):
        torch.manual_seed(seed)
        oracle = T5TransformerModelOracle(base_args(encoder_layers=2, decoder_layers=2, dprenet_dropout_rate=0.0,
                                                    bert_init=True)).eval()
        with torch.no_grad():
            for n, p in oracle.named_parameters():
                if n.endswith("alpha"):
                    p.fill_(1.1)
                if "prob_out.bias" in n:
                    p.fill_(bias)
                if "prob_out.weight" in n:
                    p.mul_(20.0)  # decisive stop logits: no borderline sigmoid(x) ~ 0.5 ties between implementations
        hf = build_hf(oracle, 2, 2)
        tok = torch.randint(4, 81, (1, 10))
        spk = torch.randn(1, 512)
        mel, probs, attn = oracle.generate_speech(src_tokens=tok, spkembs=spk)
        ref = hf.generate_speech(tok, speaker_embeddings=spk, threshold=0.5, minlenratio=0.0, maxlenratio=20.0)
        assert mel.shape == ref.shape and rel(mel, ref) < 1e-5
        assert probs.numel() == mel.shape[0] and attn.shape[2] == mel.shape[0] // 2


def test_asr_greedy_decode_matches_hf_generate():
    """sequence_generator.py greedy path (beam 1, no CTC/LM fusion) vs transformers' greedy generate with the same
    start token and pad/blank suppression: identical token ids up to the length cap, where the reference forces eos."""
    pytest.importorskip("transformers")
    from oracle.hf_crosscheck_asr import build_hf_asr
    from oracle.speecht5_oracle_asr import T5TransformerModelASROracle, base_asr_args, greedy_decode
    torch.manual_seed(1)
    oracle = T5TransformerModelASROracle(base_asr_args(encoder_layers=2, decoder_layers=2, bert_init=True)).eval()
    with torch.no_grad():
        oracle.text_decoder_postnet.output_projection.weight.mul_(30.0)  # decisive logits (no argmax ties)
    hf = build_hf_asr(oracle, 2, 2)
    wav = torch.randn(2, 8000) * 0.1
    cap = 12
    hyp = greedy_decode(oracle, wav, None, max_len_b=cap)
    gen = hf.generate(input_values=wav, do_sample=False, num_beams=1, max_new_tokens=cap, min_new_tokens=2,
                      decoder_start_token_id=2, eos_token_id=2, pad_token_id=1, suppress_tokens=[0, 1])
    for b in range(2):
        h = hyp[b].tolist()
        assert h[-1] == 2 and len(h) <= cap + 1
        ref = gen[b, 1:].tolist()
        n = min(len(h) - 1, len(ref))
        assert h[:n] == ref[:n]


# ---------------------------------------------------------------------------------------------- pre-training extras (row 22)
def test_gumbel_quantizer_oracle_matches_hf_wav2vec2_quantizer_in_eval():
    """fairseq GumbelVectorQuantizer restated; pinned in eval mode (hard codes) against the independent transformers
    Wav2Vec2GumbelVectorQuantizer (same variables layout [1, G*V, d/G], same projection)."""
    tr = pytest.importorskip("transformers")
    from transformers.models.wav2vec2.modeling_wav2vec2 import Wav2Vec2GumbelVectorQuantizer
    from oracle.pretrain_oracle import GumbelVectorQuantizer
    torch.manual_seed(0)
    d, V, G = 64, 10, 2
    q = GumbelVectorQuantizer(dim=d, num_vars=V, groups=G, vq_dim=d).eval()
    cfg = tr.Wav2Vec2Config(conv_dim=(d,), conv_stride=(2,), conv_kernel=(2,), num_codevectors_per_group=V,
                            num_codevector_groups=G, codevector_dim=d)
    hf = Wav2Vec2GumbelVectorQuantizer(cfg).eval()
    with torch.no_grad():
        hf.codevectors.copy_(q.vars)
        hf.weight_proj.weight.copy_(q.weight_proj.weight)
        hf.weight_proj.bias.copy_(q.weight_proj.bias)
    x = torch.randn(3, 17, d)
    with torch.no_grad():
        out = q(x)
        ref, ppl = hf(x)
    assert rel(out["x"], ref) < 1e-6
    assert abs(float(out["code_perplexity"]) - float(ppl)) / float(ppl) < 1e-5
    # training mode with an explicit noise draw: straight-through one-hot of argmax((logits + g) / tau)
    q.train()
    g = -torch.empty(3 * 17 * G, V).exponential_().log()
    out_t = q(x, gumbel_noise=g)
    logits = q.weight_proj(x.reshape(-1, d)).view(-1, V)
    idx = (logits + g).argmax(-1).view(3 * 17, G)
    want = torch.cat([q.vars[0, idx[:, gi] + gi * V] for gi in range(G)], dim=-1).view(3, 17, d)
    assert rel(out_t["x"], want) < 1e-5


def test_hubert_head_oracle_closed_form():
    """speech_encoder_postnet.py:61-74: logits[n, 0] = cos(proj(x_n), e[target_n]) / T and logits[n, 1 + c] =
    cos(proj(x_n), e[c]) / T, with the duplicate of the positive among the negatives masked to -inf."""
    import torch.nn.functional as F
    from oracle.pretrain_oracle import SpeechEncoderPostnet
    torch.manual_seed(1)
    head = SpeechEncoderPostnet([7], encoder_embed_dim=32, final_dim=16, logit_temp=0.1)
    x = torch.randn(2, 9, 32)
    pad = torch.zeros(2, 9, dtype=torch.bool)
    pad[1, 7:] = True
    mask = torch.rand(2, 9) < 0.5
    tgt = torch.randint(0, 7, (2, 9))
    out = head(x, pad, mask, [tgt])
    sel = ~pad & mask
    proj = head.final_proj(x[sel])
    e = head.label_embs_concat
    want = torch.stack([F.cosine_similarity(proj, e[tgt[sel]], dim=-1)] +
                       [F.cosine_similarity(proj, e[c].expand_as(proj), dim=-1) for c in range(7)], dim=1) / 0.1
    got = out["logit_m_list"][0]
    dup = torch.zeros_like(want, dtype=torch.bool)
    dup[torch.arange(want.size(0)), 1 + tgt[sel]] = True
    assert torch.isinf(got[dup]).all() and rel(got[~dup], want[~dup]) < 1e-6
    assert out["logit_u_list"][0].shape == (int((~pad & ~mask).sum()), 8)
