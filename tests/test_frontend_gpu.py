"""-m gpu, OPT-IN (ST5_TEST_FRONTEND=1): the speech-input front end (speecht5_b200/frontend.py + csrc/conv_frontend.cu)
against oracle/speecht5_oracle_asr.py. These kernels and compositions were written at the end of round 1 without GPU
time; the gate comes off once they have run green on a B200."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("ST5_TEST_FRONTEND") != "1", reason="opt-in until first validated run")]


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _args():
    from oracle import speecht5_oracle_asr as O
    args = O.base_asr_args(encoder_layers=1, decoder_layers=1, dropout=0.0)
    for k, v in dict(encoder_speech_prenet="conv", mask_prob=0.0, hubert_mask_length=10, mask_selection="static",
                     mask_other=0.0, no_mask_overlap=False, mask_min_space=1).items():
        setattr(args, k, v)
    if not isinstance(getattr(args, "conv_feature_layers", None), (list, str)):
        args.conv_feature_layers = "[(512,10,5)] + [(512,3,2)]*4 + [(512,2,2)]*2"
    return args


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_feature_extractor_forward_backward(cuda, dtype):
    from oracle import speecht5_oracle_asr as O
    from speecht5_b200 import frontend
    from speecht5_b200.ops import RT
    RT.dtype = dtype
    RT.invalidate_shadows()
    torch.manual_seed(0)
    ref = O.ConvFeatureExtractionModel().double()
    mine = frontend.ConvFeatureExtractor().to(cuda)
    mine.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    wave = torch.randn(2, 9000, dtype=torch.float64) * 0.3
    yr = ref(wave).transpose(1, 2)
    dy = torch.randn_like(yr)
    yr.backward(dy)
    y = mine(wave.float().to(cuda))
    y.backward(dy.to(cuda).to(y.dtype))
    tol = 5e-5 if dtype == torch.float32 else 3e-2
    assert y.shape == yr.shape
    assert rel(y.cpu(), yr) < tol
    gr = dict(ref.named_parameters())
    for n, p in mine.named_parameters():
        assert rel(p.grad.cpu(), gr[n].grad) < tol * 4, n
    RT.dtype = torch.bfloat16


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_speech_encoder_prenet_against_the_oracle(cuda, dtype):
    from oracle import speecht5_oracle_asr as O
    from speecht5_b200 import frontend
    from speecht5_b200.ops import RT
    RT.dtype = dtype
    RT.invalidate_shadows()
    torch.manual_seed(1)
    args = _args()
    ref = O.SpeechEncoderPrenet(args).double().eval()
    mine = frontend.SpeechEncoderPrenet(args).to(cuda).eval()
    sd = {k: v.float() for k, v in ref.state_dict().items()}
    sd["pos_conv.0.weight_g"] = sd.pop("pos_conv_g")
    sd["pos_conv.0.weight_v"] = sd.pop("pos_conv_v")
    sd["pos_conv.0.bias"] = sd.pop("pos_conv_bias")
    missing, unexpected = mine.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    B, n = 2, 8000
    wave = torch.randn(B, n, dtype=torch.float64) * 0.3
    lengths = torch.tensor([8000, 5000])
    pm = torch.arange(n)[None, :] >= lengths[:, None]
    T = int(ref.feature_extractor.get_out_seq_lens_tensor(torch.tensor([n]))[0])
    mi = torch.zeros(B, T, dtype=torch.bool)
    mi[0, 3:9] = True
    mi[1, 1:4] = True
    xr, mr, pen_r = ref(wave, pm, mask_indices=mi)
    (x, pen, _, _), m = mine(wave.float().to(cuda), require_feat_pen=True, padding_mask=pm.to(cuda), mask=True,
                             mask_indices=mi.to(cuda))
    tol = 1e-4 if dtype == torch.float32 else 3e-2
    assert torch.equal(m.cpu(), mr)
    assert rel(x.cpu(), xr) < tol
    assert abs(pen.item() - pen_r.item()) / pen_r.item() < tol
    RT.dtype = torch.bfloat16
