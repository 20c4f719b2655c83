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
