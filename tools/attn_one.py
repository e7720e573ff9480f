"""Run attention / score_select / layernorm once at BASELINE config-2 shapes (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import ops  # noqa: E402

B, N, Np = 256, 197, 173
qkv = torch.randn(B * N, 2304, device="cuda").bfloat16()
idx = torch.stack([torch.sort(torch.randperm(N, device="cuda")[:Np]).values for _ in range(B)])
rmap = (idx + torch.arange(B, device="cuda")[:, None] * N).int().flatten()
out = torch.empty(B * Np, 768, device="cuda", dtype=torch.bfloat16)
x = torch.randn(B * N, 768, device="cuda").bfloat16()
g = torch.ones(768, device="cuda")
y = torch.empty_like(x)
for _ in range(3):
    ops.attention(qkv, rmap, B, N, Np, 768, 12, 0.125, out=out)
    ops.score_select(qkv.view(B, N, 2304), 12, Np - 1)
    ops.layernorm(x, g, g, 1e-6, B * N, 768, out=y)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
