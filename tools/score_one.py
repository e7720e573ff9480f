"""Run the fused score+select kernel a few times at BASELINE config-2 block-3 shape (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import ops  # noqa: E402

B, N, H = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 197, 12
keep = max(1, int(0.88 * (N - 1)))
qkv = torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(4):
    flush.zero_()
    s, idx, nxt, rmap = ops.score_select(qkv, H, keep)
torch.cuda.synchronize()
print("ok", int(idx.sum()))
