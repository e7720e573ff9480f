"""Throughput / agreement sweep over pruning schedules (the reference leaves schedule choice to the user).

For every candidate schedule: images/s on this GPU, tensor work per image, and top-1 agreement + mean |dlogit| against the
UN-PRUNED model on the same inputs (with real weights and labels, pass a loader to rajni_vit_b200.evaluate_model instead:
agreement on random-init weights only says how much the pruning perturbs the logits).

    python tools/schedule_sweep.py [--model vit_base_patch16_224] [--batch 256] [--steps 20]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import RAJNIViTWrapper  # noqa: E402
from rajni_vit_b200.vit import create_model  # noqa: E402


def candidates(depth):
    yield "none", {}
    yield "README {3:.88,4:.88,7:.8,8:.72}", {3: {"keep_ratio": 0.88}, 4: {"keep_ratio": 0.88}, 7: {"keep_ratio": 0.8}, 8: {"keep_ratio": 0.72}}
    for r in (0.9, 0.8, 0.7):
        yield f"every block from 3, keep {r}", {i: {"keep_ratio": r} for i in range(3, depth)}
    for r in (0.7, 0.5):
        yield f"blocks 3,6,9 keep {r}", {i: {"keep_ratio": r} for i in (3, 6, 9) if i < depth}
    yield "blocks 3,6,9 keep .7, scores carried", {3: {"keep_ratio": 0.7}, 6: {"keep_ratio": 0.7, "update": False}, 9: {"keep_ratio": 0.7, "update": False}}


def flops_per_image(counts, C, hidden, P, classes=1000):
    total = 2.0 * P * C * 768 + 2.0 * C * classes
    for i, n in enumerate(counts):
        np_ = counts[i + 1] if i + 1 < len(counts) else n      # tokens after this block's pruning (last block: unknown, unchanged)
        total += 6.0 * n * C * C + 4.0 * np_ * np_ * C + 2.0 * np_ * C * C + 4.0 * np_ * C * hidden
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="vit_base_patch16_224")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    base = create_model(args.model, seed=0)
    C, depth = base.patch_embed.proj.out_channels, len(base.blocks)
    size = base.patch_embed.img_size[0]
    x = torch.randn(args.batch, 3, size, size, generator=torch.Generator().manual_seed(1234)).cuda()
    ref = None
    print(f"{'schedule':42s} {'img/s':>9s} {'GFLOP/img':>10s} {'final tokens':>12s} {'top-1 agree':>11s} {'mean|dlogit|':>12s}")
    for name, sched in candidates(depth):
        m = RAJNIViTWrapper(create_model(args.model, seed=0), sched).cuda().eval()
        y = m(x)
        for _ in range(3):
            m(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            m(x)
        e1.record()
        torch.cuda.synchronize()
        ips = args.batch * args.steps / (e0.elapsed_time(e1) * 1e-3)
        counts = m.get_last_stats()["token_counts"]
        if ref is None:
            ref = y
        agree = (y.argmax(1) == ref.argmax(1)).float().mean().item()
        dl = (y - ref).abs().mean().item()
        gf = flops_per_image(counts, C, base.blocks[0].mlp.fc1.out_features, (size // 16) ** 2) / 1e9
        print(f"{name:42s} {ips:9.0f} {gf:10.2f} {counts[-1]:12d} {agree:11.3f} {dl:12.4f}", flush=True)
        del m


if __name__ == "__main__":
    main()
