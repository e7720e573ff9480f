import os, sys, torch
sys.path.insert(0, "/root/repo")
from rajni_vit_b200 import ops
M=50432
def run(name, M, N, K, **kw):
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda"); res = torch.randn(M, N, device="cuda").bfloat16() if kw.pop("res", False) else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, w, bias, M, N, K, residual=res, out=out, **kw)
    for _ in range(20): fn()
    flops = 2.0*M*N*K; iters = max(50, int(0.4/(flops/1.2e15)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1)*1e-3/iters
    print(f"half_feed={'RAJNI_GEMM_DEBUG_HALF_FEED' in os.environ!s:5s} {name:18s} {t*1e6:8.1f} us {flops/t/1e12:7.1f} TF/s", flush=True)
run("qkv bias", M, 2304, 768)
run("fc1 bias+gelu", M, 3072, 768, gelu=True)
run("fc2 bias+res", M, 768, 3072, res=True)
run("big K=N=3072 bias", M, 3072, 3072)
