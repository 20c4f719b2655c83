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


def collate_frames(frames, is_audio_input=False):
    """text_to_speech_dataset.py:25-45 _collate_frames: zero-padded stack of [L_i, F] (or [L_i]) tensors."""
    max_len = max(f.size(0) for f in frames)
    shape = (len(frames), max_len) if is_audio_input else (len(frames), max_len, frames[0].size(1))
    out = frames[0].new_zeros(shape)
    for i, v in enumerate(frames):
        out[i, : v.size(0)] = v
    return out


def collate_tts(samples, reduction_factor=2, pad=1):
    """TextToSpeechDataset.collater (text_to_speech_dataset.py:226-281) over in-memory items
    {"id", "source": [LongTensor tokens], "target": FloatTensor [L, odim], "spkembs": FloatTensor [512],
    "audio_name"}: the batch dict the task, the criterion and the trainer consume. Host-side work only."""
    samples = [s for s in samples if s["source"] is not None]
    if len(samples) == 0:
        return {}
    fbanks = [s["target"] for s in samples]
    fbank_sizes = [len(f) for f in fbanks]
    collated = collate_frames(fbanks)
    sizes = torch.tensor(fbank_sizes, dtype=torch.long)
    r = reduction_factor
    if r > 1:  # thin out frames for the reduction factor: (B, Lmax, odim) -> (B, Lmax // r, odim)
        fb_in = collated[:, r - 1::r]
        sizes_in = torch.div(sizes, r, rounding_mode="floor")
    else:
        fb_in, sizes_in = collated, sizes
    prev = torch.cat([fb_in.new_zeros((fb_in.shape[0], 1, fb_in.shape[2])), fb_in[:, :-1]], dim=1)
    labels = collated.new_zeros(collated.size(0), collated.size(1))
    for i, n in enumerate(fbank_sizes):
        labels[i, n - 1:] = 1.0
    spkembs = collate_frames([s["spkembs"] for s in samples], is_audio_input=True)
    toks = [s["source"][0] for s in samples]
    lengths = torch.LongTensor([len(t) for t in toks])
    src = toks[0].new_full((len(toks), int(lengths.max())), pad)  # data_utils.collate_tokens(left_pad=False)
    for i, t in enumerate(toks):
        src[i, : len(t)] = t
    net_input = {"src_tokens": src, "src_lengths": lengths, "prev_output_tokens": prev, "tgt_lengths": sizes_in,
                 "spkembs": spkembs, "task_name": "t2s"}
    return {"id": torch.LongTensor([s["id"] for s in samples]), "name": [s.get("audio_name") for s in samples],
            "net_input": net_input, "labels": labels, "dec_target": collated, "dec_target_lengths": sizes,
            "src_lengths": lengths, "task_name": "t2s", "ntokens": int(lengths.sum().item()), "target": collated}


def collate_asr(samples, pad=1, eos=2):
    """SpeechToTextDataset.collater (speecht5/data/speech_to_text_dataset.py:150-222) over in-memory items
    {"id", "source": FloatTensor [N] waveform, "label_list": [LongTensor tokens]}: zero-padded waveforms with their
    boolean padding mask, targets = tokens + eos (right padded), prev_output_tokens = the same with eos moved to the
    front (fairseq collate_tokens(move_eos_to_beginning=True)). Host-side work only."""
    samples = [s for s in samples if s["source"] is not None]
    if len(samples) == 0:
        return {}
    audios = [s["source"] for s in samples]
    n = max(len(a) for a in audios)
    source = audios[0].new_zeros(len(audios), n)
    padding_mask = torch.zeros(len(audios), n, dtype=torch.bool)
    for i, a in enumerate(audios):
        source[i, : len(a)] = a
        padding_mask[i, len(a):] = True
    labels = [torch.cat((s["label_list"][0].long(), torch.tensor([eos]))) for s in samples]
    lengths = torch.tensor([len(t) for t in labels], dtype=torch.long)
    T = int(lengths.max())
    target = torch.full((len(labels), T), pad, dtype=torch.long)
    prev = torch.full((len(labels), T), pad, dtype=torch.long)
    for i, t in enumerate(labels):
        target[i, : len(t)] = t
        prev[i, 0] = eos
        prev[i, 1: len(t)] = t[:-1]
    ntokens = int(sum(len(s["label_list"][0]) for s in samples))
    return {"id": torch.LongTensor([s["id"] for s in samples]),
            "net_input": {"source": source, "padding_mask": padding_mask, "prev_output_tokens": prev,
                          "task_name": "s2t"},
            "target": target, "target_lengths": lengths, "task_name": "s2t", "ntokens": ntokens}


def compute_mask_indices(shape, padding_mask, mask_prob, mask_length, mask_type="static", mask_other=0.0, min_masks=0,
                         no_overlap=False, min_space=0):
    """Span masks for the speech prenet (speech_encoder_prenet.py:236-262 calls fairseq/data/data_utils.py:393-517).
    Host-side numpy, drawing from the GLOBAL np.random stream in the reference's order (one rand() for the whole
    batch, one more per row when a padding mask is given, the span lengths, the span starts, then the per-row
    thinning to the common count), so a run seeded like the reference masks the same frames. Returns a bool ndarray
    [B, T]. `no_overlap` (off in every SpeechT5 recipe) is not built."""
    import numpy as np
    if no_overlap:
        raise NotImplementedError("no_overlap span placement is not built (unused by the SpeechT5 recipes)")
    bsz, all_sz = shape
    mask = np.full((bsz, all_sz), False)
    all_num_mask = max(min_masks, int(mask_prob * all_sz / float(mask_length) + np.random.rand()))
    mask_idcs = []
    for i in range(bsz):
        if padding_mask is not None:
            sz = all_sz - int(padding_mask[i].long().sum().item())
            num_mask = max(min_masks, int(mask_prob * sz / float(mask_length) + np.random.rand()))
        else:
            sz, num_mask = all_sz, all_num_mask
        if mask_type == "static":
            lengths = np.full(num_mask, mask_length)
        elif mask_type == "uniform":
            lengths = np.random.randint(mask_other, mask_length * 2 + 1, size=num_mask)
        elif mask_type == "normal":
            lengths = [max(1, int(round(x))) for x in np.random.normal(mask_length, mask_other, size=num_mask)]
        elif mask_type == "poisson":
            lengths = [int(round(x)) for x in np.random.poisson(mask_length, size=num_mask)]
        else:
            raise Exception("unknown mask selection " + mask_type)
        if sum(lengths) == 0:
            lengths[0] = min(mask_length, sz - 1)
        min_len = min(lengths)
        if sz - min_len <= num_mask:
            min_len = sz - num_mask - 1
        starts = np.random.choice(sz - min_len, num_mask, replace=False)
        idc = np.asarray([starts[j] + off for j in range(len(starts)) for off in range(lengths[j])])
        mask_idcs.append(np.unique(idc[idc < sz]))
    min_len = min(len(m) for m in mask_idcs)
    for i, idc in enumerate(mask_idcs):
        if len(idc) > min_len:
            idc = np.random.choice(idc, min_len, replace=False)
        mask[i, idc] = True
    return mask


def _pin(obj):
    if torch.is_tensor(obj):
        return obj.pin_memory()
    if isinstance(obj, dict):
        return {k: _pin(v) for k, v in obj.items()}
    return obj
