"""CPU restatement (oracle) of the two audio ends of the SpeechT5 TTS path -- SURVEY.md section 8a rows 15 and 16:
the log-mel filterbank that produces the decoder targets and the HiFi-GAN generator that turns mels into waveforms.
TEST INFRASTRUCTURE ONLY (see oracle/speecht5_oracle.py header).

Row 15 follows speecht5/data/text_to_speech_dataset.py:95-138, whose arithmetic lives in librosa (not vendored, version
unpinned): librosa.stft(center=True, pad_mode="reflect", periodic hann) and librosa.filters.mel(htk=False,
norm="slaney"); restated here in numpy from their published definitions and pinned against torchaudio and the
HuggingFace SpeechT5FeatureExtractor (tests/test_oracle_cpu.py). Row 16 follows the sibling tree's
SpeechUT/fairseq/fairseq/models/text_to_speech/hifigan.py:13-170 (weight norm folded: inference-time weights), with
the vocoder configuration of the SpeechT5 release (80 -> 512, upsample 4x4x4x4 with kernels 8, ResBlocks 3/7/11 with
dilations 1,3,5); pinned against transformers.SpeechT5HifiGan."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ row 15: log-mel
def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis(sr=16000, n_fft=1024, n_mels=80, fmin=80.0, fmax=7600.0):
    """librosa.filters.mel(htk=False, norm="slaney"): triangles on the Slaney mel scale, each scaled by
    2 / (f_right - f_left) (constant energy per channel). [n_mels, 1 + n_fft/2] float32."""
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, fftfreqs.size))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


def logmelfilterbank(audio, sampling_rate=16000, fft_size=1024, hop_size=256, num_mels=80, fmin=80, fmax=7600,
                     eps=1e-10):
    """text_to_speech_dataset.py:95-138: reflect-pad fft_size/2, frames of fft_size * periodic hann every hop_size,
    |rFFT|, mel projection, log10(max(eps, .)). audio [N] -> [1 + N // hop_size, num_mels] float32."""
    audio = np.asarray(audio, dtype=np.float32)
    pad = fft_size // 2
    x = np.pad(audio, (pad, pad), mode="reflect")
    n_frames = 1 + (x.size - fft_size) // hop_size
    idx = np.arange(fft_size)[None, :] + hop_size * np.arange(n_frames)[:, None]
    window = (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(fft_size) / fft_size)).astype(np.float32)
    spc = np.abs(np.fft.rfft(x[idx] * window[None, :], axis=1)).astype(np.float32)
    mel = spc @ mel_basis(sampling_rate, fft_size, num_mels, fmin, fmax).T
    return np.log10(np.maximum(eps, mel)).astype(np.float32)


# ------------------------------------------------------------------------------------------------ row 16: HiFi-GAN
HIFIGAN_CFG = dict(model_in_dim=80, upsample_initial_channel=512, upsample_rates=[4, 4, 4, 4],
                   upsample_kernel_sizes=[8, 8, 8, 8], resblock_kernel_sizes=[3, 7, 11],
                   resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])
LRELU_SLOPE = 0.1


class ResBlock(nn.Module):
    """hifigan.py:20-102: three (leaky_relu -> dilated conv -> leaky_relu -> conv) residual steps, "same" padding."""

    def __init__(self, channels, kernel_size, dilation):
        super().__init__()
        pad = lambda k, d: (k * d - d) // 2  # noqa: E731
        self.convs1 = nn.ModuleList([nn.Conv1d(channels, channels, kernel_size, 1, dilation=d,
                                               padding=pad(kernel_size, d)) for d in dilation])
        self.convs2 = nn.ModuleList([nn.Conv1d(channels, channels, kernel_size, 1, dilation=1,
                                               padding=pad(kernel_size, 1)) for _ in dilation])

    def forward(self, x):
        for c1, c2 in zip(self.convs1, self.convs2):
            xt = c1(F.leaky_relu(x, LRELU_SLOPE))
            xt = c2(F.leaky_relu(xt, LRELU_SLOPE))
            x = xt + x
        return x


class HifiGanGenerator(nn.Module):
    """hifigan.py:105-170 with weight norm folded into the weights. Input log-mel [B, T, 80] (optionally normalised by
    the vocoder's stored mean / scale, as the SpeechT5 release does) -> waveform [B, T * 256]."""

    def __init__(self, cfg=None, std=0.01, seed=None):
        super().__init__()
        cfg = dict(HIFIGAN_CFG, **(cfg or {}))
        self.cfg = cfg
        self.num_kernels = len(cfg["resblock_kernel_sizes"])
        self.num_upsamples = len(cfg["upsample_rates"])
        c0 = cfg["upsample_initial_channel"]
        self.conv_pre = nn.Conv1d(cfg["model_in_dim"], c0, 7, 1, padding=3)
        self.ups = nn.ModuleList([nn.ConvTranspose1d(c0 // 2 ** i, c0 // 2 ** (i + 1), k, u, padding=(k - u) // 2)
                                  for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"]))])
        self.resblocks = nn.ModuleList()
        ch = c0
        for i in range(self.num_upsamples):
            ch = c0 // 2 ** (i + 1)
            for k, d in zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"]):
                self.resblocks.append(ResBlock(ch, k, d))
        self.conv_post = nn.Conv1d(ch, 1, 7, 1, padding=3)
        self.register_buffer("mean", torch.zeros(cfg["model_in_dim"]))
        self.register_buffer("scale", torch.ones(cfg["model_in_dim"]))
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        for m in self.modules():  # hifigan.py:13-16 init_weights: N(0, 0.01) on every conv
            if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d)):
                with torch.no_grad():
                    m.weight.copy_(torch.randn(m.weight.shape, generator=g) * std)

    def forward(self, spectrogram, normalize_before=True):
        if normalize_before:
            spectrogram = (spectrogram - self.mean) / self.scale
        x = self.conv_pre(spectrogram.transpose(1, 2))
        for i in range(self.num_upsamples):
            x = self.ups[i](F.leaky_relu(x, LRELU_SLOPE))
            xs = None
            for j in range(self.num_kernels):
                y = self.resblocks[i * self.num_kernels + j](x)
                xs = y if xs is None else xs + y
            x = xs / self.num_kernels
        x = self.conv_post(F.leaky_relu(x))  # (default slope 0.01 here, as in the reference :165)
        return torch.tanh(x).squeeze(1)


def hifigan_to_hf_state(sd, cfg=None):
    """Key mapping onto transformers.SpeechT5HifiGan (independent implementation used as the pin)."""
    cfg = dict(HIFIGAN_CFG, **(cfg or {}))
    m = {"mean": sd["mean"], "scale": sd["scale"]}
    for wb in ("weight", "bias"):
        m[f"conv_pre.{wb}"] = sd[f"conv_pre.{wb}"]
        m[f"conv_post.{wb}"] = sd[f"conv_post.{wb}"]
        for i in range(len(cfg["upsample_rates"])):
            m[f"upsampler.{i}.{wb}"] = sd[f"ups.{i}.{wb}"]
        n_res = len(cfg["upsample_rates"]) * len(cfg["resblock_kernel_sizes"])
        for r in range(n_res):
            for j in range(3):
                m[f"resblocks.{r}.convs1.{j}.{wb}"] = sd[f"resblocks.{r}.convs1.{j}.{wb}"]
                m[f"resblocks.{r}.convs2.{j}.{wb}"] = sd[f"resblocks.{r}.convs2.{j}.{wb}"]
    return m


def fold_weight_norm(sd):
    """torch.nn.utils.weight_norm (dim 0) as the reference generator applies it to every conv (hifigan.py:116-150):
    weight = g * v / ||v||, the norm taken over all dims but the first. Returns a plain {.weight, .bias} state dict."""
    out = {}
    for k, v in sd.items():
        if k.endswith("weight_g"):
            base = k[: -len("weight_g")]
            wv = sd[base + "weight_v"]
            norm = wv.reshape(wv.size(0), -1).norm(dim=1).view(-1, *([1] * (wv.dim() - 1)))
            out[base + "weight"] = v * wv / norm
        elif not k.endswith("weight_v"):
            out[k] = v
    return out


def load_reference_hifigan_state(gen, sd):
    """Load a state dict of the reference's weight-normed Generator into the folded-weight oracle."""
    folded = fold_weight_norm(sd)
    folded.setdefault("mean", gen.mean)
    folded.setdefault("scale", gen.scale)
    gen.load_state_dict(folded)
    return gen
