"""Throughput / accuracy sweep over pruning schedules + greedy search (rajni_vit_b200.schedule).

    python tools/schedule_sweep.py [--model vit_base_patch16_224] [--batch 256] [--batches 2] [--steps 10] [--floor 95] [--ckpt file]

Accuracy axis: agreement with the UN-PRUNED model on synthetic images (random-init or --ckpt weights); with real data use
rajni_vit_b200.schedule.sweep / search on batches of (images, labels) from the CLI's loader - the axis is then top-1.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import load_checkpoint, schedule as S  # noqa: E402
from rajni_vit_b200.vit import create_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="vit_base_patch16_224")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--batches", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--floor", type=float, default=None, help="also run the greedy search with this accuracy floor (%%)")
    ap.add_argument("--ckpt", default=None)
    args = ap.parse_args()

    def make():
        m = create_model(args.model, seed=0)
        if args.ckpt:
            load_checkpoint(args.ckpt, m)
        return m

    size = make().patch_embed.img_size[0]
    g = torch.Generator().manual_seed(1234)
    batches = [torch.randn(args.batch, 3, size, size, generator=g) for _ in range(args.batches)]
    res = S.sweep(make, batches, timing_steps=args.steps)
    print(f"{'schedule':48s} {'img/s':>9s} {'GFLOP/img':>10s} {'final tokens':>12s} {'accuracy %':>10s}")
    for r in res:
        print(f"{r['name']:48s} {r['img_s']:9.0f} {r['gflop_per_image']:10.2f} {r['token_counts'][-1]:12d} {r['accuracy']:10.2f}", flush=True)
    print("Pareto front (fastest first):", [r["name"] for r in S.pareto_front(res)])
    if args.floor is not None:
        sched, hist = S.search(make, batches, floor=args.floor, timing_steps=max(3, args.steps // 2))
        print(f"greedy search, accuracy >= {args.floor} %: {len(hist) - 1} accepted moves")
        for h in hist:
            print(f"  {h['img_s']:9.0f} img/s  {h['accuracy']:6.2f} %  {h['schedule']}")
        print("schedule:", sched)


if __name__ == "__main__":
    main()
