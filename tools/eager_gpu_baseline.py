"""What the reference's op sequence costs in eager PyTorch on the SAME B200 ("the honest bar to beat on
the same box").  The reference itself cannot travel to the GPU box, so this is a plain-torch restatement of its forward
(rajni/wrapper/model.py:30-69, attention.py:17-60, importance.py:5-34): pruned blocks run the un-fused
matmul-softmax-matmul attention the reference uses, un-pruned blocks use F.scaled_dot_product_attention like timm's block,
Linear/LayerNorm/GELU go to cuBLAS / ATen.  Reported for fp32 (the reference's default) and after .bfloat16().

    python tools/eager_gpu_baseline.py [--batch 256] [--steps 10]
"""
import argparse
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200.vit import create_model  # noqa: E402

SCHEDULE = {3: 0.88, 4: 0.88, 7: 0.8, 8: 0.72}


@torch.no_grad()
def importance(qkv, H):
    B, N, C3 = qkv.shape
    D = C3 // 3 // H
    qkv = qkv.view(B, N, 3, H, D)
    q_cls, k, v = qkv[:, 0, 0], qkv[:, :, 1], qkv[:, :, 2]
    a = (torch.einsum("bhd,bnhd->bhn", q_cls, k) / math.sqrt(D)).softmax(-1).mean(1)
    vm = v.mean(2)
    r = (vm - vm.mean(1, keepdim=True)).norm(dim=-1)
    return a * torch.sigmoid((r - r.mean(1, keepdim=True)) / (r.std(1, keepdim=True) + 1e-6))


@torch.no_grad()
def forward(m, x, schedule):
    x = m.patch_embed.proj(x).flatten(2).transpose(1, 2)
    x = torch.cat([m.cls_token.expand(x.shape[0], -1, -1), x], 1) + m.pos_embed[:, : x.shape[1] + 1]
    for i, blk in enumerate(m.blocks):
        B, N, C = x.shape
        H = blk.attn.num_heads
        if i in schedule:
            qkv = blk.attn.qkv(blk.norm1(x))
            keep = max(1, int(schedule[i] * (N - 1)))
            idx = torch.topk(importance(qkv, H)[:, 1:], keep, dim=1).indices.sort(dim=1).values + 1
            idx = torch.cat([torch.zeros_like(idx[:, :1]), idx], 1)
            qkv = torch.gather(qkv, 1, idx[:, :, None].expand(-1, -1, 3 * C))
            q, k, v = qkv.view(B, keep + 1, 3, H, C // H).permute(2, 0, 3, 1, 4)
            att = ((q @ k.transpose(-2, -1)) * blk.attn.scale).softmax(-1) @ v
            out = blk.attn.proj(att.transpose(1, 2).reshape(B, keep + 1, C))
            x = torch.gather(x, 1, idx[:, :, None].expand(-1, -1, C)) + out
        else:
            q, k, v = blk.attn.qkv(blk.norm1(x)).view(B, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
            att = F.scaled_dot_product_attention(q, k, v)
            x = x + blk.attn.proj(att.transpose(1, 2).reshape(B, N, C))
        x = x + blk.mlp.fc2(F.gelu(blk.mlp.fc1(blk.norm2(x))))
    return m.head(m.norm(x)[:, 0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    for dtype in (torch.float32, torch.bfloat16):
        m = create_model("vit_base_patch16_224", seed=0).cuda().to(dtype).eval()
        x = torch.randn(args.batch, 3, 224, 224, device="cuda", dtype=dtype)
        for _ in range(3):
            forward(m, x, SCHEDULE)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            forward(m, x, SCHEDULE)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(f"eager PyTorch on this GPU, {str(dtype).split('.')[-1]:8s} batch {args.batch}: {ms:8.2f} ms/step  {args.batch / ms * 1e3:9.0f} img/s", flush=True)


if __name__ == "__main__":
    main()
