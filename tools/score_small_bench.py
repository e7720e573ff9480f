"""score_select at the per-GPU batches of the 8-GPU configs (C3: vit_small bs 64, C4: vit_large bs 32, C5: deit-384 bs 16):
split path (K/V pass spread over (image, row-block) CTAs + per-image selection kernel) against the fused one-CTA-per-image kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import ops  # noqa: E402


def timeit(fn, inner=50, iters=7):
    """GPU time per call: `inner` calls captured in a CUDA graph (no host launch cost), median over `iters` replays."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner):
            fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / inner)
    return sorted(ts)[len(ts) // 2] * 1e3


CASES = (("C3 vit_small bs64", 64, 197, 6, 0.7), ("C3 late layer", 64, 60, 6, 0.7), ("C4 vit_large bs32", 32, 197, 16, 0.9),
                         ("C4 late layer", 32, 40, 16, 0.9), ("C5 deit384 bs16", 16, 577, 12, 0.88), ("C2 bs256", 256, 197, 12, 0.88))
if len(sys.argv) > 1:        # score_small_bench.py B:N:H ...   (other shapes)
    CASES = tuple((a, *[int(v) for v in a.split(":")], 0.88) for a in sys.argv[1:])
for name, B, N, H, r in CASES:
    keep = max(1, int(r * (N - 1)))
    nbuf = max(1, min(4, int(400e6 // (B * N * 3 * H * 64 * 2)) + 1))     # rotate over > 126 MB of inputs: the pass comes from HBM
    qkvs = [torch.randn(B, N, 3 * H * 64, device="cuda").bfloat16() for _ in range(nbuf)]
    out = {}
    it = [0]
    for split in (True, False):
        idx = torch.empty(B, keep + 1, device="cuda", dtype=torch.int32)
        nxt = torch.empty(B, keep + 1, device="cuda")
        rmap = torch.empty(B * (keep + 1), device="cuda", dtype=torch.int32)
        def call():
            it[0] += 1
            ops.score_select(qkvs[it[0] % nbuf], H, keep, keep_idx=idx, next_scores=nxt, row_map=rmap, split=split)
        out[split] = timeit(call, inner=48)
    mb = B * (2 * N * H * 64 * 2) / 1e6
    print(f"{name:20s} B={B:3d} N={N:3d} C={H*64:4d}: split {out[True]:6.1f} us   fused {out[False]:6.1f} us   ({mb:6.1f} MB of K,V = {mb/6.4e3*1e3:5.1f} us at HBM peak)", flush=True)
