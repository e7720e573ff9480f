"""Can attention run beside the GEMMs on a slice of the SMs?  (The GEMMs are power-capped: tools/gemm_sm_cap_probe.py shows
they lose only ~10 % on 116 of 148 SMs.)  Steady state of a two-half-batch pipeline, synthetic:
  serial : [attention, proj, fc1, fc2, qkv] at batch 256 on all SMs, one stream
  overlap: main stream = the four GEMMs at batch 128, twice, on RAJNI_GEMM_MAX_CTAS SMs; side stream = attention at batch 128,
           twice, on RAJNI_ATTN_MAX_CTAS SMs (set both in the environment; the caps are read once per process).
usage: overlap_probe.py serial|overlap [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import ops  # noqa: E402

mode = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 197
C, H = 768, 12
B = 256 if mode == "serial" else 128
M = B * N
dev = "cuda"
x = torch.randn(M, C, device=dev).bfloat16()
qkv = torch.randn(M, 3 * C, device=dev).bfloat16()
att = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
hid = torch.empty(M, 4 * C, device=dev, dtype=torch.bfloat16)
wq = (torch.randn(3 * C, C, device=dev) / 28).bfloat16()
wp = (torch.randn(C, C, device=dev) / 28).bfloat16()
w1 = (torch.randn(4 * C, C, device=dev) / 28).bfloat16()
w2 = (torch.randn(C, 4 * C, device=dev) / 55).bfloat16()
bq, bp, b1, b2 = (torch.randn(n, device=dev) for n in (3 * C, C, 4 * C, C))


def gemms():
    ops.gemm(att, wp, bp, M, C, C, residual=x, ldres=C, out=x, ldd=C)
    ops.gemm(x, w1, b1, M, 4 * C, C, gelu=True, out=hid, ldd=4 * C)
    ops.gemm(hid, w2, b2, M, C, 4 * C, residual=x, ldres=C, out=x, ldd=C)
    ops.gemm(x, wq, bq, M, 3 * C, C, out=qkv)


def attn():
    ops.attention(qkv, None, B, N, N, C, H, 0.125, out=att)


side = torch.cuda.Stream()
main = torch.cuda.current_stream()


def step():
    if mode == "serial":
        attn()
        gemms()
    else:
        for _ in range(2):                     # two half batches; their data dependencies are not modelled, only the occupancy
            side.wait_stream(main)
            with torch.cuda.stream(side):
                attn()
            gemms()
        main.wait_stream(side)


for _ in range(10):
    step()
torch.cuda.synchronize()
iters = 300
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    step()
e1.record()
torch.cuda.synchronize()
t = e0.elapsed_time(e1) / iters
print(f"{mode:8s} N={N} gemm_cap={os.environ.get('RAJNI_GEMM_MAX_CTAS', '-'):>4s} attn_cap={os.environ.get('RAJNI_ATTN_MAX_CTAS', '-'):>4s}: "
      f"{t * 1e3:8.1f} us per layer of 256 images", flush=True)
