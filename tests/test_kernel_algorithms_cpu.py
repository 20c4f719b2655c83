"""CPU: executable design notes for the convolution rows that come next. Each test states, in
numpy/torch on the host, the exact operand VIEW the device GEMM will be given -- no im2col copies -- and checks it
against torch's convolution. The postnet's 5-tap convolution already runs this way (speecht5_b200/ops.py:Conv1dK5Fn)."""
import numpy as np
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
