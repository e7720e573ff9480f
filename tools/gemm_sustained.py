"""Sustained (power-capped) GEMM rates: each variant runs back to back for ~0.4 s, CUDA events around the loop.
usage: gemm_sustained.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import ops  # noqa: E402


def run(name, M, N, K, iters=None, torch_ref=False, **kw):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16() if kw.pop("res", False) else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fn = (lambda: torch.nn.functional.linear(a, w, out=None)) if torch_ref else \
         (lambda: ops.gemm(a, w, bias, M, N, K, residual=res, out=out, **kw))
    for _ in range(20):
        fn()
    flops = 2.0 * M * N * K
    iters = iters or max(50, int(0.4 / (flops / 1.2e15)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / iters
    print(f"{name:34s} M={M:6d} N={N:5d} K={K:5d} {t*1e6:8.1f} us {flops/t/1e12:7.1f} TF/s  ({iters} back-to-back launches)", flush=True)


M = 50432
run("qkv   bias", M, 2304, 768)
run("fc1   bias+gelu", M, 3072, 768, gelu=True)
run("fc1   bias only", M, 3072, 768)
run("fc1   cuBLAS (no epilogue)", M, 3072, 768, torch_ref=True)
run("fc2   bias+res", M, 768, 3072, res=True)
run("fc2   cuBLAS (no epilogue)", M, 768, 3072, torch_ref=True)
run("proj  bias+res", M, 768, 768, res=True)
run("proj  bias only", M, 768, 768)
run("proj  cuBLAS (no epilogue)", M, 768, 768, torch_ref=True)
run("K=3072 N=3072 bias+gelu", M, 3072, 3072, gelu=True)
run("K=3072 N=3072 bias", M, 3072, 3072)
