"""Does running two half-batches on two streams fill the wave-quantisation / tail gaps of the persistent kernels?
Compares one stream at batch 256 with two streams at batch 128 each (same total images)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import RAJNIViTWrapper  # noqa: E402
from rajni_vit_b200.vit import create_model  # noqa: E402

SCHEDULE = {3: {"keep_ratio": 0.88}, 4: {"keep_ratio": 0.88}, 7: {"keep_ratio": 0.8}, 8: {"keep_ratio": 0.72}}
dev = torch.device("cuda")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100


def timed(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


m = RAJNIViTWrapper(create_model("vit_base_patch16_224", seed=0), SCHEDULE).to(dev).eval()
x = torch.randn(256, 3, 224, 224, device=dev)
t1 = timed(lambda: m(x), steps)
print(f"1 stream  x 256: {t1:.3f} ms/step  {256 / t1 * 1e3:.0f} img/s", flush=True)

ma = RAJNIViTWrapper(create_model("vit_base_patch16_224", seed=0), SCHEDULE).to(dev).eval()
mb = RAJNIViTWrapper(create_model("vit_base_patch16_224", seed=0), SCHEDULE).to(dev).eval()
xa, xb = x[:128].contiguous(), x[128:].contiguous()
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()


def two():
    with torch.cuda.stream(sa):
        ma(xa)
    with torch.cuda.stream(sb):
        mb(xb)


torch.cuda.synchronize()
for _ in range(5):
    two()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    two()
torch.cuda.synchronize()
t2 = (time.perf_counter() - t0) / steps * 1e3
print(f"2 streams x 128: {t2:.3f} ms/step  {256 / t2 * 1e3:.0f} img/s", flush=True)
t3 = timed(lambda: ma(xa), steps)
print(f"1 stream  x 128: {t3:.3f} ms/step  {128 / t3 * 1e3:.0f} img/s", flush=True)
