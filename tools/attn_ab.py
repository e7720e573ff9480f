"""A/B of the two short-sequence attention kernels (rajni_attention_fwd_ex: PIPE vs TC) at given shapes: CUDA-event time per
call over back-to-back launches rotating over several qkv buffers.  usage: attn_ab.py [B H] N:Np [N:Np ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import _lib, ops  # noqa: E402

args = sys.argv[1:]
B, H = 256, 12
if args and ":" not in args[0]:
    B, H = int(args[0]), int(args[1])
    args = args[2:]
shapes = [tuple(int(v) for v in a.split(":")) for a in args] or [(87, 87), (96, 96), (112, 112), (64, 64), (121, 87), (197, 197)]
C = H * 64
for N, Np in shapes:
    nbuf = max(2, int(300e6 // (B * N * 3 * C * 2)) + 1)
    qkvs = [torch.randn(B * N, 3 * C, device="cuda").bfloat16() for _ in range(nbuf)]
    rmap = None
    if Np < N:
        idx = torch.stack([torch.cat([torch.zeros(1, dtype=torch.long, device="cuda"), torch.sort(torch.randperm(N - 1, device="cuda")[:Np - 1]).values + 1]) for _ in range(B)])
        rmap = (idx + torch.arange(B, device="cuda")[:, None] * N).int().flatten()
    out = torch.empty(B * Np, C, device="cuda", dtype=torch.bfloat16)
    res = {}
    for name, impl in (("pipe", _lib.ATTN_PIPE), ("tc", _lib.ATTN_TC), ("auto", _lib.ATTN_AUTO)):
        if name == "pipe" and (Np + 15) // 16 * 16 > 224:
            continue
        for i in range(4):
            ops.attention(qkvs[i % nbuf], rmap, B, N, Np, C, H, 0.125, out=out, impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30):
            ops.attention(qkvs[i % nbuf], rmap, B, N, Np, C, H, 0.125, out=out, impl=impl)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 30 * 1e3
    print(f"B={B} H={H} N={N:3d} Np={Np:3d}: " + "  ".join(f"{k} {v:7.1f} us" for k, v in res.items()), flush=True)
