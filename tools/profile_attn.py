"""Fused attention launches of the decoder (self: causal 313x313, cross: 313x160 with probabilities) forward + backward,
for `ncu --set full -k regex:attn_fused`. The second iteration's launches are the ones to read."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speecht5_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
B, H, d, Td, Te = 32, 12, 768, 313, 160
ops.RT.manual_seed(3)
for it in range(2):
    qkv = (torch.randn(B, Td, 3 * d, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    out, _ = ops.attention(qkv, None, H=H, d=d, q_col=0, k_col=1, v_col=2, scale=0.125, causal=True,
                           drop_p=0.1)
    out.float().square().sum().backward()
    q = (torch.randn(B, Td, d, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    kv = (torch.randn(B, Te, 2 * d, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    pad = torch.zeros(B, Te, dtype=torch.bool, device=dev)
    pad[:, 150:] = True
    out, probs = ops.attention(q, kv, H=H, d=d, q_col=0, k_col=0, v_col=1, scale=0.125, key_pad=pad, drop_p=0.1,
                               return_probs=True)
    (out.float().square().sum() + probs.square().sum()).backward()
    torch.cuda.synchronize()
print("done")
