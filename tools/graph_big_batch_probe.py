import sys, torch
sys.path.insert(0, "/root/repo")
from rajni_vit_b200 import RAJNIViTWrapper
from rajni_vit_b200.vit import create_model
S = {3: {"keep_ratio": 0.88}, 4: {"keep_ratio": 0.88}, 7: {"keep_ratio": 0.8}, 8: {"keep_ratio": 0.72}}
m = RAJNIViTWrapper(create_model("vit_base_patch16_224", seed=0), S).cuda().eval()
xs = [torch.randn(256, 3, 224, 224, device="cuda") for _ in range(2)]
for mode in (False, True, False, True):
    m.use_cuda_graph = mode
    for i in range(10): m(xs[i & 1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(150): m(xs[i & 1])
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 150
    print(f"graph={mode!s:5s} {t:.3f} ms/step {256/t*1e3:.0f} img/s", flush=True)
