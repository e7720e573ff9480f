import sys, torch
sys.path.insert(0, "/root/repo")
from rajni_vit_b200 import RAJNIViTWrapper, ops
from rajni_vit_b200.vit import create_model
S = {3: {"keep_ratio": 0.88}, 4: {"keep_ratio": 0.88}, 7: {"keep_ratio": 0.8}, 8: {"keep_ratio": 0.72}}
m = RAJNIViTWrapper(create_model("vit_base_patch16_224", seed=0), S).cuda().eval()
for B in ([int(v) for v in sys.argv[1:]] or [32, 64]):
    x = torch.randn(B, 3, 224, 224, device="cuda")
    for mode in (False, None):
        m.use_cuda_graph = mode
        for _ in range(5): m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): m(x)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 200
        print(f"B={B} graph={mode!s:5s} {t:.3f} ms/step {B/t*1e3:.0f} img/s", flush=True)
    m.use_cuda_graph = False
    prof = ops.profile_steps(lambda: m(x), steps=3)
    tot = sum(v["ms"] for v in prof.values())
    print("   per class (eager, event-timed): " + "  ".join(f"{k}={v['ms']:.3f}/{v['launches']}" for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])) + f"  total={tot:.3f}")
