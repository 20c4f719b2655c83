"""Generates tests/golden/audio_tiny.npz from oracle/audio_oracle.py: a seeded 0.5 s waveform with its log-mel
filterbank and a small HiFi-GAN generator (2 up-sampling stages, rows 16) with its weights, input
mels and output waveform. Run in the build container: `python tests/golden/make_golden_audio.py`. The oracle is pinned
against torchaudio / the HuggingFace feature extractor and transformers.SpeechT5HifiGan by tests/test_oracle_cpu.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.audio_oracle import HifiGanGenerator, logmelfilterbank  # noqa: E402

TINY_VOCODER = dict(upsample_initial_channel=32, upsample_rates=[4, 4], upsample_kernel_sizes=[8, 8],
                    resblock_kernel_sizes=[3, 7], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]])


def main():
    rng = np.random.default_rng(21)
    t = np.arange(8000) / 16000.0
    wav = (0.3 * np.sin(2 * np.pi * 220.0 * t) + 0.1 * rng.standard_normal(8000)).astype(np.float32)
    blob = {"mel/wav": wav, "mel/logmel": logmelfilterbank(wav)}
    gen = HifiGanGenerator(TINY_VOCODER, seed=3).double().eval()
    with torch.no_grad():
        g = torch.Generator().manual_seed(4)
        gen.mean.copy_(torch.randn(80, generator=g).double() * 0.5)
        gen.scale.copy_(torch.rand(80, generator=g).double() + 0.5)
        for m in gen.modules():
            if isinstance(m, (torch.nn.Conv1d, torch.nn.ConvTranspose1d)):
                m.weight.mul_(8.0)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g).double() * 0.05)
        mel = torch.randn(2, 24, 80, generator=g).double()
        out = gen(mel)
    for k, v in gen.state_dict().items():
        blob["voc/state/" + k] = v.float().numpy()
    blob["voc/in"], blob["voc/out"] = mel.float().numpy(), out.float().numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "audio_tiny.npz")
    np.savez_compressed(path, **blob)
    print(path, os.path.getsize(path), out.shape, float(out.abs().max()))


if __name__ == "__main__":
    main()
