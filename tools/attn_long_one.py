"""Run the long-sequence attention kernel a few times (for ncu captures).  usage: attn_long_one.py [B N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import ops  # noqa: E402

B, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 577)
qkv = torch.randn(B * N, 2304, device="cuda").bfloat16()
out = torch.empty(B * N, 768, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.attention(qkv, None, B, N, N, 768, 12, 0.125, out=out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
