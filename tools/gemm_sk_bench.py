"""fc2-shaped GEMMs (bias + residual + row statistics, in place like the model) with and without the stream-K tail:
CUDA-event time per call over back-to-back launches, rotating over > 126 MB of operands.  tools/gemm_sk_bench.py [K N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import ops  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
N = int(sys.argv[2]) if len(sys.argv) > 2 else 768
ws = ops.gemm_workspace("cuda")
slots = ops.row_stats_slots(N)
flags = ops.EPI_BIAS | ops.EPI_RESIDUAL | ops.EPI_ROW_STATS


def bench(M, workspace, iters=40):
    nbuf = max(2, int(300e6 // (M * K * 2)) + 1)
    a = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(nbuf)]
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda")
    x = [torch.randn(M, N, device="cuda").bfloat16() for _ in range(nbuf)]
    stats = torch.zeros((slots, M, 2), device="cuda")
    def call(i):
        ops.gemm(a[i % nbuf], w, b, M, N, K, residual=x[i % nbuf], ldres=N, out=x[i % nbuf], ldd=N, row_stats=stats, workspace=workspace)
    for i in range(5):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


print(f"# N={N} K={K}; fix={os.environ.get('RAJNI_GEMM_SK_FIX', '10')}")
ONLY = [int(v) for v in os.environ.get("SK_BENCH_MS", "").split(",") if v]
for tokens in (197, 173, 152, 121, 87):
    for B in (256, 64, 32):
        M = tokens * B
        if ONLY and M not in ONLY:
            continue
        plan = ops.stream_k_plan(M, N, K, flags)
        t0 = bench(M, None)
        t1 = bench(M, ws) if plan[0] else float("nan")
        tf = 2.0 * M * N * K / 1e6
        print(f"M={M:6d} ({tokens:3d} x {B:3d}) tiles={((M + 255) // 256) * (N // 256):4d} plan={plan!s:9s}  plain {t0:7.1f} us {tf / t0:7.1f} TF/s   stream-K {t1:7.1f} us {tf / t1:7.1f} TF/s  x{t0 / t1:.3f}", flush=True)
