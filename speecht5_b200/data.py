"""Synthetic batches with the reference collaters' key contract (host I/O itself -- soundfile, librosa log-mel,
manifests -- is outside the hot path). TTS: speecht5/data/text_to_speech_dataset.py:223-281 (right-padded text tokens,
mel targets [B, L, 80], prev_output_tokens = [0; every r-th frame][:-1], stop labels 1 from the last real frame on,
512-d x-vectors)."""
import torch


def synthetic_tts_batch(B, T_txt, T_mel, vocab=81, odim=80, r=2, seed=1, ragged=True, pad=1, pin=False):
    g = torch.Generator().manual_seed(seed)
    src_lengths = (torch.randint(int(0.75 * T_txt), T_txt + 1, (B,), generator=g) if ragged and B > 1
                   else torch.full((B,), T_txt))
    src_lengths[0] = T_txt
    src_tokens = torch.randint(4, vocab, (B, T_txt), generator=g)
    ar = torch.arange(T_txt)[None, :]
    src_tokens = torch.where(ar < src_lengths[:, None], src_tokens, torch.full_like(src_tokens, pad))
    mel_lengths = (torch.randint(int(0.9 * T_mel), T_mel + 1, (B,), generator=g) if ragged and B > 1
                   else torch.full((B,), T_mel))
    mel_lengths[0] = T_mel
    fbank = torch.randn(B, T_mel, odim, generator=g)
    am = torch.arange(T_mel)[None, :]
    fbank = fbank * (am < mel_lengths[:, None]).unsqueeze(-1)
    fb_in = fbank[:, r - 1::r]
    len_in = torch.div(mel_lengths, r, rounding_mode="floor")
    prev = torch.cat([fb_in.new_zeros((B, 1, odim)), fb_in[:, :-1]], dim=1).contiguous()
    labels = (am >= (mel_lengths[:, None] - 1)).float()
    spk = torch.randn(B, 512, generator=g)
    net_input = dict(src_tokens=src_tokens, src_lengths=src_lengths, prev_output_tokens=prev, tgt_lengths=len_in,
                     spkembs=spk, task_name="t2s")
    sample = dict(net_input=net_input, labels=labels, dec_target=fbank, dec_target_lengths=mel_lengths,
                  src_lengths=src_lengths, task_name="t2s", ntokens=int(src_lengths.sum()), target=fbank)
    if pin:
        sample = _pin(sample)
    return sample


def _pin(obj):
    if torch.is_tensor(obj):
        return obj.pin_memory()
    if isinstance(obj, dict):
        return {k: _pin(v) for k, v in obj.items()}
    return obj
