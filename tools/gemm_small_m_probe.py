"""Tile-shape choice at small M (per-GPU shards of the 8-GPU configs): time the GEMM shapes of ViT-B at batch 32 under the
forced tile widths (RAJNI_GEMM_BN, RAJNI_GEMM_CG1 are read once per process: run once per setting).
usage: [RAJNI_GEMM_BN=256|192|128] [RAJNI_GEMM_CG1=1] gemm_small_m_probe.py [batch=32]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tag = f"BN={os.environ.get('RAJNI_GEMM_BN', 'auto'):>4s} CG1={os.environ.get('RAJNI_GEMM_CG1', '0')}"


def run(name, M, N, K, **kw):
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16() if kw.pop("res", False) else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, bias, M, N, K, residual=res, out=out, **kw)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 20)
    t = sorted(ts)[3] * 1e-3
    print(f"{tag} {name:10s} M={M:6d} N={N:5d} K={K:5d} {t * 1e6:7.1f} us {2.0 * M * N * K / t / 1e12:7.1f} TF/s", flush=True)


for N_tok in (197, 87):
    M = B * N_tok
    run("qkv", M, 2304, 768)
    run("proj", M, 768, 768, res=True)
    run("fc1", M, 3072, 768, gelu=True)
    run("fc2", M, 768, 3072, res=True)
