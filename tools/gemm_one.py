"""Run one GEMM shape a few times (for ncu captures).  usage: gemm_one.py {qkv|proj|fc1|fc2}"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import ops  # noqa: E402

SHAPES = {"qkv": (50432, 2304, 768, {}), "proj": (50432, 768, 768, {"res": True}),
          "fc1": (50432, 3072, 768, {"gelu": True}), "fc2": (50432, 768, 3072, {"res": True})}
name = sys.argv[1] if len(sys.argv) > 1 else "qkv"
M, N, K, kw = SHAPES[name]
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
bias = torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda").bfloat16() if kw.get("res") else None
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.gemm(a, w, bias, M, N, K, residual=res, out=out, gelu=kw.get("gelu", False))
torch.cuda.synchronize()
print("ok", name, float(out.float().abs().mean()))
