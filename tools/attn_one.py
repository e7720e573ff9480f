"""Run the attention kernel a few times at BASELINE config-2 shapes (for ncu captures).
usage: attn_one.py [N Np]   (default 197 173; N == Np runs without a row map)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import ops  # noqa: E402

B = 256
N, Np = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (197, 173)
qkv = torch.randn(B * N, 2304, device="cuda").bfloat16()
rmap = None
if Np < N:
    idx = torch.stack([torch.sort(torch.randperm(N, device="cuda")[:Np]).values for _ in range(B)])
    rmap = (idx + torch.arange(B, device="cuda")[:, None] * N).int().flatten()
out = torch.empty(B * Np, 768, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.attention(qkv, rmap, B, N, Np, 768, 12, 0.125, out=out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
