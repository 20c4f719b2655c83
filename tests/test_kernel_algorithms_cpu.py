"""CPU: executable design notes for the convolution rows that come next. Each test states, in
numpy/torch on the host, the exact operand VIEW the device GEMM will be given -- no im2col copies -- and checks it
against torch's convolution. The postnet's 5-tap convolution already runs this way (speecht5_b200/ops.py:Conv1dK5Fn)."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _window_view(x_cl, k, stride):
    """x_cl [B, T, C] channels-last -> [B, T_out, k*C] VIEW: row t starts at frame t*stride and spans k frames (k*C
    contiguous elements). This is the rank-4 TMA map of the A operand: row pitch stride*C, row length k*C."""
    B, T, C = x_cl.shape
    T_out = (T - k) // stride + 1
    s = x_cl.stride()
    return x_cl.as_strided((B, T_out, k * C), (s[0], stride * C, 1))


def test_strided_conv1d_is_one_gemm_over_an_overlapping_window_view():
    """wav2vec2-style front-end layers 1..6 (speech_encoder_prenet.py:277-354: Conv1d(512, 512, k in {3, 2}, stride 2,
    no bias)): out[b, t, :] = W2 . x[b, 2t : 2t + k, :].ravel() with W2 = W.permute(0, 2, 1).reshape(C_out, k*C_in)."""
    torch.manual_seed(0)
    for k, stride in ((3, 2), (2, 2), (10, 5)):
        B, T, Cin, Cout = 2, 57, 16, 24
        x = torch.randn(B, T, Cin).contiguous()
        W = torch.randn(Cout, Cin, k)
        ref = F.conv1d(x.transpose(1, 2), W, stride=stride).transpose(1, 2)  # [B, T_out, Cout]
        A = _window_view(x, k, stride)
        W2 = W.permute(0, 2, 1).reshape(Cout, k * Cin)
        out = A @ W2.t()
        assert out.shape == ref.shape and torch.allclose(out, ref, atol=1e-4)
        assert A.untyped_storage().data_ptr() == x.untyped_storage().data_ptr()  # a view, not a copy


def test_transposed_conv1d_splits_into_stride_phase_gemms():
    """HiFi-GAN up-sampling (hifigan.py:120-134: ConvTranspose1d(C, C/2, k=8, stride=4, padding=2)): output sample
    n = 4m + r only sees the taps j with (n + p - j) % 4 == 0, so each of the 4 phases r is an ordinary 2-tap
    convolution of the input -- a window GEMM over the zero-padded channels-last input, written with row stride 4."""
    torch.manual_seed(1)
    B, T, Cin, Cout, k, s, p = 2, 19, 12, 6, 8, 4, 2
    x = torch.randn(B, T, Cin)
    W = torch.randn(Cin, Cout, k)  # ConvTranspose1d weight layout
    ref = F.conv_transpose1d(x.transpose(1, 2), W, stride=s, padding=p).transpose(1, 2)  # [B, T*s, Cout]
    out = torch.zeros(B, T * s, Cout)
    taps = k // s  # 2 input frames contribute to every output sample
    xp = F.pad(x, (0, 0, taps - 1, taps - 1))  # one frame of zeros on both sides
    for r in range(s):
        # output n = s*m + r  <-  sum_i x[m + d_i] . W[:, :, j_i] with j_i = r + p - s*d_i in [0, k)
        ds_ = [d for d in range(-(taps - 1), taps) if 0 <= r + p - s * d < k]
        d0 = min(ds_)
        Wr = torch.stack([W[:, :, r + p - s * d] for d in ds_], dim=0)  # [taps, Cin, Cout], frame order d0, d0+1
        A = _window_view(xp.contiguous(), len(ds_), 1)  # rows: padded frames m' .. m'+taps-1
        rows = A[:, (taps - 1) + d0: (taps - 1) + d0 + T]  # row m uses frames m + d0 .. m + d0 + taps - 1
        out[:, r::s] = rows @ Wr.reshape(len(ds_) * Cin, Cout)
    assert torch.allclose(out, ref, atol=1e-4)


def test_dilated_conv1d_is_a_window_gemm_per_dilation_phase():
    """HiFi-GAN ResBlock convolutions (hifigan.py:20-102: kernel 3/7/11, dilation 1/3/5, "same" padding): with the time
    axis de-interleaved into d phases (frame t -> phase t % d, index t // d) a dilation-d convolution is a plain k-tap
    window GEMM inside every phase; d = 1 needs no regrouping."""
    torch.manual_seed(2)
    B, T, C, k, d = 2, 45, 8, 7, 3
    x = torch.randn(B, T, C)
    W = torch.randn(C, C, k)
    pad = (k * d - d) // 2
    ref = F.conv1d(x.transpose(1, 2), W, dilation=d, padding=pad).transpose(1, 2)
    xp = F.pad(x, (0, 0, pad, pad))
    W2 = W.permute(0, 2, 1).reshape(C, k * C)
    out = torch.zeros_like(ref)
    for ph in range(d):
        xph = xp[:, ph::d].contiguous()  # frames ph, ph + d, ... of the padded input
        A = _window_view(xph, k, 1)  # output t = ph + d*m reads padded frames t, t + d, ..., t + (k-1) d
        n = out[:, ph::d].shape[1]
        out[:, ph::d] = (A @ W2.t())[:, :n]
    assert torch.allclose(out, ref, atol=1e-4)


def test_groupnorm_with_one_channel_per_group_is_per_utterance_batchnorm_statistics():
    """Front-end layer 0 (speech_encoder_prenet.py:318-323: GroupNorm(512 groups, 512 channels)): per (utterance,
    channel) statistics over time == training-mode BatchNorm statistics of that utterance's [T, C] rows, which is what
    st5_bn_fwd computes (biased variance, eps inside the square root)."""
    torch.manual_seed(3)
    B, C, T = 3, 16, 101
    x = torch.randn(B, C, T) * 2 + 0.5
    g, b = torch.randn(C), torch.randn(C)
    ref = F.group_norm(x, C, g, b, eps=1e-5)
    rows = x.transpose(1, 2)  # [B, T, C] channels-last
    mean = rows.mean(dim=1, keepdim=True)
    var = rows.var(dim=1, unbiased=False, keepdim=True)
    out = ((rows - mean) / torch.sqrt(var + 1e-5) * g + b).transpose(1, 2)
    assert torch.allclose(out, ref, atol=1e-5)
    assert np.isfinite(out.numpy()).all()


def _ctc_numpy(logits, targets, input_len, target_len, blank=0):
    """CTC negative log-likelihood and its gradient wrt the logits of ONE utterance, the way the device kernel will do
    it (criterions/speech_to_text_loss.py:303-335 calls F.ctc_loss(reduction="sum", zero_infinity=...)): log-space
    alpha / beta recursions over the extended label sequence l' = (blank, l1, blank, ..., lL, blank), one thread per
    extended position s, T sequential steps; grad[t, k] = y[t, k] - sum_{s: l'_s = k} exp(alpha[t, s] + beta[t, s] -
    lp[t, k] + nll), rows t >= input_len zero."""
    T, V = logits.shape
    lp = logits - np.logaddexp.reduce(logits, axis=1, keepdims=True)
    L = int(target_len)
    ext = np.full(2 * L + 1, blank, dtype=np.int64)
    ext[1::2] = targets[:L]
    S, Tn = 2 * L + 1, int(input_len)
    ninf = -np.inf
    alpha = np.full((Tn, S), ninf)
    beta = np.full((Tn, S), ninf)
    alpha[0, 0] = lp[0, blank]
    if S > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, Tn):
        for s in range(S):
            a = alpha[t - 1, s]
            if s >= 1:
                a = np.logaddexp(a, alpha[t - 1, s - 1])
            if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                a = np.logaddexp(a, alpha[t - 1, s - 2])
            alpha[t, s] = a + lp[t, ext[s]]
    beta[Tn - 1, S - 1] = lp[Tn - 1, blank]
    if S > 1:
        beta[Tn - 1, S - 2] = lp[Tn - 1, ext[S - 2]]
    for t in range(Tn - 2, -1, -1):
        for s in range(S):
            b = beta[t + 1, s]
            if s + 1 < S:
                b = np.logaddexp(b, beta[t + 1, s + 1])
            if s + 2 < S and ext[s] != blank and ext[s] != ext[s + 2]:
                b = np.logaddexp(b, beta[t + 1, s + 2])
            beta[t, s] = b + lp[t, ext[s]]
    ll = alpha[Tn - 1, S - 1] if S == 1 else np.logaddexp(alpha[Tn - 1, S - 1], alpha[Tn - 1, S - 2])
    nll = -ll
    grad = np.zeros_like(logits)
    if np.isfinite(nll):
        grad[:Tn] = np.exp(lp[:Tn])
        for t in range(Tn):
            acc = np.full(V, ninf)
            for s in range(S):
                acc[ext[s]] = np.logaddexp(acc[ext[s]], alpha[t, s] + beta[t, s])
            grad[t] -= np.exp(acc - lp[t] + nll)
    return nll, grad


def test_ctc_recursion_and_gradient_match_torch():
    """The ASR criterion's CTC term (speech_to_text_loss.py:303-335): per-utterance alpha/beta recursion and the
    closed-form gradient wrt the logits, against torch's ctc_loss + log_softmax autograd (repeated labels, ragged
    input / target lengths, an infeasible utterance under zero_infinity)."""
    torch.manual_seed(4)
    T, B, V = 23, 4, 9
    logits = torch.randn(T, B, V, dtype=torch.float64, requires_grad=True)
    targets = [torch.tensor([3, 3, 5, 1, 1, 2]), torch.tensor([4, 2]), torch.tensor([7, 7, 7, 7, 7, 7, 7, 7]),
               torch.tensor([1, 2, 3])]
    input_lengths = torch.tensor([23, 15, 12, 20])  # utterance 2: 8 repeated labels need 15 frames > 12 -> infeasible
    target_lengths = torch.tensor([len(t) for t in targets])
    lp = F.log_softmax(logits, dim=-1)
    loss = F.ctc_loss(lp, torch.cat(targets), input_lengths, target_lengths, blank=0, reduction="sum",
                      zero_infinity=True)
    loss.backward()
    total = 0.0
    for b in range(B):
        nll, grad = _ctc_numpy(logits.detach().numpy()[:, b], targets[b].numpy(), input_lengths[b], target_lengths[b])
        if not np.isfinite(nll):  # zero_infinity: the utterance contributes neither loss nor gradient
            nll, grad = 0.0, np.zeros_like(grad)
        total += nll
        assert np.allclose(grad, logits.grad[:, b].numpy(), atol=1e-9), b
    assert abs(total - loss.item()) < 1e-9


def _split_bf16(x):
    hi = x.to(torch.bfloat16).float()
    lo = (x - hi).to(torch.bfloat16).float()
    return hi, lo


def test_logmel_as_two_gemms_with_split_precision():
    """SURVEY section 8a row 15 (text_to_speech_dataset.py:95-138) as the device path will run it: the framed,
    reflect-padded waveform [n_frames, 1024] times ONE windowed real-DFT matrix [1024, 2*513] (cos | -sin with the
    periodic hann folded in) on the existing GEMM kernel, a magnitude epilogue over column pairs, then the [513, 80] mel
    GEMM with a log10(max(eps, .)) epilogue. Checked against the oracle's FFT formulation in fp32, and in the GEMM
    kernel's split-bf16 parity arithmetic (hi*hi + hi*lo + lo*hi, fp32 accumulate) to bound what that mode costs."""
    from oracle.audio_oracle import logmelfilterbank, mel_basis
    rng = np.random.default_rng(11)
    n = 16000
    t = np.arange(n) / 16000.0
    audio = (0.3 * np.sin(2 * np.pi * 220 * t) + 0.1 * np.sin(2 * np.pi * 3100 * t) + 0.02 * rng.standard_normal(n))
    audio = audio.astype(np.float32)
    want = logmelfilterbank(audio)
    nfft, hop = 1024, 256
    x = np.pad(audio, (nfft // 2, nfft // 2), mode="reflect")
    n_frames = 1 + (x.size - nfft) // hop
    frames = torch.from_numpy(x[np.arange(nfft)[None, :] + hop * np.arange(n_frames)[:, None]])
    k = np.arange(nfft // 2 + 1)[None, :]
    ang = 2.0 * np.pi * (np.arange(nfft)[:, None] * k % nfft) / nfft  # exact argument reduction on the host
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(nfft) / nfft)
    dft = torch.from_numpy(np.concatenate([np.cos(ang), -np.sin(ang)], axis=1) * win[:, None]).float()
    basis = torch.from_numpy(mel_basis()).t().contiguous()

    def finish(spec2):
        mag = torch.sqrt(spec2[:, :513] ** 2 + spec2[:, 513:] ** 2)
        return mag

    def split_mm(a, b):
        ah, al = _split_bf16(a)
        bh, bl = _split_bf16(b)
        return ah @ bh + ah @ bl + al @ bh

    got32 = torch.log10(torch.clamp(finish(frames @ dft) @ basis, min=1e-10)).numpy()
    assert got32.shape == want.shape == (1 + n // hop, 80)
    assert np.abs(got32 - want).max() < 2e-4
    mag_s = finish(split_mm(frames, dft))
    got_s = torch.log10(torch.clamp(split_mm(mag_s, basis), min=1e-10)).numpy()
    # log10 domain: an absolute error of 1e-3 is a 0.23 % magnitude error, inside the TTS target noise floor
    assert np.abs(got_s - want).max() < 1e-3, np.abs(got_s - want).max()


def test_incremental_decode_with_kv_cache_equals_prefix_recompute():
    """Synthesis (models/speecht5.py:1188-1249; the reference keeps fairseq incremental state,
    multihead_attention.py:255-330): the planned device decode loop projects the cross-attention K/V of every decoder
    layer ONCE per utterance, appends one self-attention K/V row per step to a [layer, H, T_max, 64] cache and runs
    one-row attention -- O(L) work per step instead of the O(L^2) prefix recompute generate_speech does today. Same
    numbers as the prefix recompute at every step (causality), including the per-layer cross-attention rows the
    guided-attention diagnostics read."""
    from oracle.speecht5_oracle import T5TransformerModelOracle, base_args
    torch.manual_seed(5)
    args = base_args(encoder_layers=1, decoder_layers=2, dprenet_dropout_rate=0.0)
    m = T5TransformerModelOracle(args).double().eval()
    d, H = args.decoder_embed_dim, args.decoder_attention_heads
    hd = d // H
    src = torch.randint(4, 81, (1, 9))
    spk = torch.randn(1, 512, dtype=torch.float64)
    enc_in, enc_mask = m.text_encoder_prenet(src)
    enc_out = m.encoder(enc_in, enc_mask)
    enc = enc_out["encoder_out"][0]  # [S, 1, d]
    L = 7
    frames = torch.randn(1, L, 80, dtype=torch.float64)

    def heads(x):  # [T, 1, d] -> [H, T, hd]
        return x.view(x.size(0), H, hd).transpose(0, 1)

    with torch.no_grad():
        cross = [(heads(l.encoder_attn.k_proj(enc)), heads(l.encoder_attn.v_proj(enc))) for l in m.decoder.layers]
        cache_k = [torch.zeros(H, L, hd, dtype=torch.float64) for _ in m.decoder.layers]
        cache_v = [torch.zeros(H, L, hd, dtype=torch.float64) for _ in m.decoder.layers]
        for t in range(L):
            x_all, _ = m.speech_decoder_prenet(frames[:, : t + 1], spkembs=spk)
            want, extra = m.decoder(x_all, None, enc_out, alignment_layer=-1)
            x = x_all[:, t:].transpose(0, 1)  # the new row only, [1, 1, d]
            for li, l in enumerate(m.decoder.layers):
                assert not l.normalize_before
                a = l.self_attn
                cache_k[li][:, t] = heads(a.k_proj(x))[:, 0]
                cache_v[li][:, t] = heads(a.v_proj(x))[:, 0]
                q = heads(a.q_proj(x) * a.scaling)  # [H, 1, hd]
                p = torch.softmax(q @ cache_k[li][:, : t + 1].transpose(1, 2), dim=-1)
                o = (p @ cache_v[li][:, : t + 1]).transpose(0, 1).reshape(1, 1, d)
                x = l.self_attn_layer_norm(x + a.out_proj(o))
                c = l.encoder_attn
                q = heads(c.q_proj(x) * c.scaling)
                pc = torch.softmax(q @ cross[li][0].transpose(1, 2), dim=-1)  # [H, 1, S]
                o = (pc @ cross[li][1]).transpose(0, 1).reshape(1, 1, d)
                x = l.encoder_attn_layer_norm(x + c.out_proj(o))
                x = l.final_layer_norm(x + l.fc2(F.gelu(l.fc1(x).float()).type_as(x)))  # fp32 GELU as the reference
                assert torch.allclose(pc[:, 0], extra["attn"][0][li][0, :, -1], atol=1e-12)
            assert torch.allclose(x[0, 0], want[0, -1], atol=1e-6), t  # the fp32 GELU rounding bounds it


def test_fused_conv0_groupnorm_gelu_statement_matches_autograd():
    """csrc/conv_frontend.cu, restated step by step in numpy/torch on the CPU: chunked (sum, centred M2) statistics
    combined with Chan's formula, y = gelu((v - mean) rstd gamma + beta), and the backward the kernels use --
    g = dy gelu'(z), S1 = sum_t g, S2 = sum_t g xhat, dbeta = sum_b S1, dgamma = sum_b S2,
    dv = rstd gamma (g - S1/T - xhat S2/T), dW[c, k] = sum_{b,t} dv[b, t, c] wave[b, t*stride + k] -- against autograd
    through Conv1d -> GroupNorm(C, C) -> GELU (speech_encoder_prenet.py:290-327)."""
    torch.manual_seed(8)
    B, n, Cc, Kt, S, TCH = 2, 1403, 6, 10, 5, 128
    wave = torch.randn(B, n, dtype=torch.float64) + 0.3
    w = torch.randn(Cc, Kt, dtype=torch.float64, requires_grad=True)
    gamma = torch.randn(Cc, dtype=torch.float64, requires_grad=True)
    beta = torch.randn(Cc, dtype=torch.float64, requires_grad=True)
    eps = 1e-5
    v_ref = F.conv1d(wave[:, None], w[:, None], stride=S)  # [B, C, T0]
    y_ref = F.gelu(F.group_norm(v_ref, Cc, gamma, beta, eps)).transpose(1, 2)  # channels-last [B, T0, C]
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    T0 = (n - Kt) // S + 1
    assert y_ref.shape == (B, T0, Cc)
    with torch.no_grad():
        win = wave.unfold(1, Kt, S)  # [B, T0, K] overlapping windows of the waveform
        v = win @ w.t()  # [B, T0, C]
        # forward statistics: per chunk (sum, M2 about the chunk mean), then Chan's combination
        tot, parts = torch.zeros(B, Cc, dtype=torch.float64), []
        for t0 in range(0, T0, TCH):
            blk = v[:, t0:t0 + TCH]
            s = blk.sum(1)
            parts.append((blk.shape[1], s, ((blk - s[:, None] / blk.shape[1]) ** 2).sum(1)))
            tot += s
        mean = tot / T0
        m2 = sum(p[2] + p[0] * (p[1] / p[0] - mean) ** 2 for p in parts)
        rstd = 1.0 / torch.sqrt(m2 / T0 + eps)
        xh = (v - mean[:, None]) * rstd[:, None]
        z = xh * gamma + beta
        assert torch.allclose(F.gelu(z), y_ref, atol=1e-12)
        # backward
        cdf = 0.5 * (1 + torch.erf(z / math.sqrt(2)))
        g = dy * (cdf + z * torch.exp(-0.5 * z * z) / math.sqrt(2 * math.pi))
        s1, s2 = g.sum(1), (g * xh).sum(1)
        assert torch.allclose(s1.sum(0), beta.grad, atol=1e-10)
        assert torch.allclose(s2.sum(0), gamma.grad, atol=1e-10)
        dv = rstd[:, None] * gamma * (g - s1[:, None] / T0 - xh * s2[:, None] / T0)
        dw = torch.einsum("btc,btk->ck", dv, win)
        assert torch.allclose(dw, w.grad, atol=1e-9)


def _ctc_kernel_emulation(logits, tg, Tn, blank=0):
    """csrc/ctc.cu restated statement by statement in fp32 numpy, vectorised over the threads s of one CTA: the
    prev/cur state exchange, the skip-transition predicates, the backward sweep that finishes each gradient row in the
    linear domain (terms below e^-80 dropped), rows t >= Tn zeroed."""
    f = np.float32
    T, V = logits.shape
    L = len(tg)
    S = 2 * L + 1
    s = np.arange(S)
    sym = np.where(s & 1, np.asarray(tg, dtype=np.int64)[np.minimum(s >> 1, max(L - 1, 0))] if L else blank, blank)
    odd = (s & 1) == 1
    skip = np.zeros(S, bool)
    skip_f = np.zeros(S, bool)
    for i in range(S):
        if odd[i] and i >= 2:
            skip[i] = tg[(i >> 1) - 1] != sym[i]
        if odd[i] and i + 2 < S:
            skip_f[i] = tg[(i >> 1) + 1] != sym[i]
    lse = np.logaddexp.reduce(logits.astype(np.float64), axis=1).astype(f)
    ninf = f(-np.inf)

    def lse2(a, b):
        with np.errstate(invalid="ignore"):
            m = np.maximum(a, b)
            r = m + np.log1p(np.exp(-np.abs(a - b))).astype(f)
        return np.where(np.isneginf(m), ninf, r).astype(f)

    def shift(x, k):  # x[s - k] (k > 0) or x[s + |k|] (k < 0), -inf outside
        out = np.full(S, ninf, f)
        if k > 0:
            out[k:] = x[:-k] if k < S else []
        else:
            out[:k] = x[-k:]
        return out
    aw = np.full((T, S), ninf, f)
    prev = np.full(S, ninf, f)
    if Tn >= 1:
        prev[:2] = (logits[0, sym[:2]] - lse[0]).astype(f)
    aw[0] = prev
    for t in range(1, Tn):
        v = lse2(prev, shift(prev, 1))
        v = np.where(skip, lse2(v, shift(prev, 2)), v)
        prev = (v + (logits[t, sym] - lse[t])).astype(f)
        aw[t] = prev
    ll = ninf if Tn < 1 else (lse2(prev[S - 1:S], prev[S - 2:S - 1])[0] if S >= 2 else prev[0])
    nll = -ll
    grad = np.zeros((T, V), f)
    if not np.isfinite(nll):
        return nll, grad
    for t in range(Tn - 1, -1, -1):
        lp_sym = (logits[t, sym] - lse[t]).astype(f)
        if t == Tn - 1:
            v = np.where(s >= S - 2, f(0), ninf)
        else:
            v = lse2(prev, shift(prev, -1))
            v = np.where(skip_f, lse2(v, shift(prev, -2)), v)
        be = (v + lp_sym).astype(f)
        e = aw[t] + be - lp_sym + nll
        acc = np.zeros(V, f)
        with np.errstate(invalid="ignore"):
            np.add.at(acc, sym[e > -80], np.exp(e[e > -80]).astype(f))
        grad[t] = np.exp(logits[t] - lse[t]).astype(f) - acc
        prev = be
    return nll, grad


def test_ctc_kernel_program_matches_torch():
    """The thread program of csrc/ctc.cu (emulated in fp32) against torch's ctc_loss + log_softmax autograd: repeated
    labels, an empty target, a one-frame input, ragged lengths and an infeasible utterance (zero_infinity)."""
    torch.manual_seed(12)
    T, V = 19, 7
    cases = [([3, 3, 5, 1, 1, 2], 19), ([4, 2], 11), ([], 5), ([6], 1), ([2, 2, 2, 2, 2, 2], 10), ([1, 2, 3], 16)]
    for tg, Tn in cases:
        logits = torch.randn(T, 1, V, dtype=torch.float32, requires_grad=True)
        loss = F.ctc_loss(F.log_softmax(logits, -1), torch.tensor(tg, dtype=torch.long)[None] if tg else
                          torch.zeros(1, 0, dtype=torch.long), torch.tensor([Tn]), torch.tensor([len(tg)]), blank=0,
                          reduction="sum", zero_infinity=True)
        loss.backward()
        nll, grad = _ctc_kernel_emulation(logits.detach().numpy()[:, 0], tg, Tn)
        feasible = np.isfinite(nll)
        assert abs((nll if feasible else 0.0) - loss.item()) < 2e-4 * max(1.0, abs(loss.item())), (tg, Tn)
        assert np.abs(grad - logits.grad.numpy()[:, 0]).max() < 2e-5, (tg, Tn)
    assert not np.isfinite(_ctc_kernel_emulation(np.zeros((T, V), np.float32), [2, 2, 2, 2, 2, 2], 10)[0])


def _conv0_ln_kernel_program(wave, w, gamma, beta, dy, stride, eps, n_ctas=3):
    """csrc/conv_frontend.cu conv0_ln_fwd_kernel / conv0_ln_bwd_kernel restated step for step: a warp per frame, lane l
    owns the channel pairs (2l + 64j, 2l + 64j + 1); forward = convolution from k broadcast taps, mean, CENTRED variance,
    normalise, affine, GELU. Backward = a PAIR of warps per frame: both recompute x-hat, g = dy * gelu'(z) and the two
    LayerNorm sums; warp `role` accumulates the taps k = role (mod 2) of dW and one of dgamma (role 0) / dbeta (role 1)
    over its frames (frames are dealt to pairs round-robin); every pair writes ONE partial row [C*K | C | C] and a last
    pass sums the rows. Returns (y, mean, rstd, dW, dgamma, dbeta)."""
    B, n = wave.shape
    C, K = w.shape
    T0 = (n - K) // stride + 1
    frames = B * T0
    pairs = n_ctas * 4                                  # C0L_WARPS / 2 pairs per CTA
    lanes = [[c for j in range(8) for c in (2 * l + 64 * j, 2 * l + 64 * j + 1) if c < C] for l in range(32)]
    assert sorted(c for ln in lanes for c in ln) == list(range(C))  # the lanes partition the channels
    y = np.zeros((frames, C))
    mean, rstd = np.zeros(frames), np.zeros(frames)
    part = np.zeros((pairs, C * K + 2 * C))

    def gelu(z):
        return 0.5 * z * (1.0 + np.vectorize(math.erf)(z / math.sqrt(2.0)))

    def dgelu(z):
        return 0.5 * (1.0 + np.vectorize(math.erf)(z / math.sqrt(2.0))) + z * np.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)

    for f in range(frames):
        b, t = divmod(f, T0)
        x = wave[b, t * stride: t * stride + K]
        v = w @ x                                        # every lane: its channels, the same K taps
        mu = v.sum() / C                                 # warp_sum over the lanes' partial sums
        q = ((v - mu) ** 2).sum() / C
        rs = 1.0 / math.sqrt(q + eps)
        mean[f], rstd[f] = mu, rs
        y[f] = gelu((v - mu) * rs * gamma + beta)
    for p in range(pairs):
        for role in (0, 1):
            acc_w = np.zeros((C, K))
            acc_aff = np.zeros(C)
            for f in range(p, frames, pairs):
                b, t = divmod(f, T0)
                x = wave[b, t * stride: t * stride + K]
                xh = (w @ x - mean[f]) * rstd[f]
                g = dy[f] * dgelu(xh * gamma + beta)
                acc_aff += g if role else g * xh
                dxh = g * gamma
                m1, m2 = dxh.sum() / C, (dxh * xh).sum() / C
                du = rstd[f] * (dxh - m1 - xh * m2)
                for k in range(role, K, 2):               # taps 2i + role
                    acc_w[:, k] += du * x[k]
            for k in range(role, K, 2):
                part[p, np.arange(C) * K + k] = acc_w[:, k]
            off = C * K + (C if role else 0)
            part[p, off: off + C] = acc_aff
    tot = part.sum(0)                                    # conv0_reduce_w_kernel over the rows
    return (y.reshape(B, T0, C), mean, rstd, tot[: C * K].reshape(C, K), tot[C * K: C * K + C], tot[C * K + C:])


@pytest.mark.parametrize("C,K,stride", [(32, 10, 5), (70, 7, 3)])
def test_conv0_layer_norm_kernel_program_matches_autograd(C, K, stride):
    """The layer-0 kernel pair of the "layer_norm" waveform extractor (speech_encoder_prenet.py:308-318) as a program:
    channel ownership, per-frame statistics, the role split of the backward and the partial-row layout reproduce
    torch's conv1d + layer_norm + gelu and its gradients (the device kernels are checked against the same statement in
    tests/test_frontend_gpu.py)."""
    rng = np.random.default_rng(C)
    B, n = 2, 83
    wave, w = rng.normal(size=(B, n)) * 0.3, rng.normal(size=(C, K)) * 0.4
    gamma, beta = 1 + 0.2 * rng.normal(size=C), 0.2 * rng.normal(size=C)
    T0 = (n - K) // stride + 1
    dy = rng.normal(size=(B * T0, C))
    y, mean, rstd, dW, dg, db = _conv0_ln_kernel_program(wave, w, gamma, beta, dy, stride, 1e-5)
    tw, tg, tb = (torch.tensor(a, requires_grad=True) for a in (w, gamma, beta))
    v = F.conv1d(torch.tensor(wave)[:, None], tw[:, None], stride=stride).transpose(1, 2)
    want = F.gelu(F.layer_norm(v, (C,), tg, tb, 1e-5))
    gw, gg, gb = torch.autograd.grad(want, (tw, tg, tb), torch.tensor(dy).view(B, T0, C))
    np.testing.assert_allclose(y, want.detach().numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(mean, v.detach().mean(-1).reshape(-1).numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dW, gw.numpy(), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(dg, gg.numpy(), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(db, gb.numpy(), rtol=1e-8, atol=1e-10)
