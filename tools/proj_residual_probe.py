"""Where does the residual read of the proj GEMM cost its 29 %?  Sustained (0.4 s per line), M = 50432, N = K = 768.
  bias only | + residual (77 MB, DRAM) | + residual through an identity row map | + residual rows folded into a 0.75 MB region (cache hits)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import ops  # noqa: E402

M, N, K = 50432, 768, 768
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
bias = torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
ident = torch.arange(M, device="cuda", dtype=torch.int32)
fold = (ident % 512).contiguous()


def run(name, **kw):
    fn = lambda: ops.gemm(a, w, bias, M, N, K, out=out, **kw)
    for _ in range(20):
        fn()
    iters = 6000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / iters
    print(f"{name:52s} {t * 1e6:7.1f} us  {2.0 * M * N * K / t / 1e12:7.1f} TF/s", flush=True)


run("bias only")
run("bias + residual (77 MB)", residual=res, ldres=N)
run("bias + residual, identity row map", residual=res, ldres=N, res_row_map=ident)
run("bias + residual, rows folded into 0.75 MB", residual=res, ldres=N, res_row_map=fold)
