"""Per-kernel-class time of BASELINE configs 2-5 at their FULL global batch on one GPU (the `configs` key of bench.py),
every launch bracketed by its own CUDA-event pair (ops.profile_steps).  usage: config_profile.py [C3 C4 ...]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rajni_vit_b200 import RAJNIViTWrapper, ops  # noqa: E402
from rajni_vit_b200.vit import create_model  # noqa: E402

README = {3: {"keep_ratio": 0.88}, 4: {"keep_ratio": 0.88}, 7: {"keep_ratio": 0.8}, 8: {"keep_ratio": 0.72}}
CASES = [("C2", "vit_base_patch16_224", README, 256, 224),
         ("C3", "vit_small_patch16_224", {i: {"keep_ratio": 0.7} for i in range(3, 12)}, 512, 224),
         ("C4", "vit_large_patch16_224", {i: {"keep_ratio": 0.9} for i in range(24)}, 256, 224),
         ("C5", "deit_base_patch16_384", README, 128, 384)]
# config_profile.py C4 C5 ...      full global batch;   config_profile.py C4/8 ...   the batch of one of 8 shards
only = {a.split("/")[0]: int(a.split("/")[1]) if "/" in a else 1 for a in sys.argv[1:]}
for name, model_name, sched, B, S in CASES:
    if only and name not in only:
        continue
    B //= only.get(name, 1)
    m = RAJNIViTWrapper(create_model(model_name, seed=0), sched).cuda().eval()
    m.use_cuda_graph = False
    x = torch.randn(B, 3, S, S, device="cuda")
    for _ in range(3):
        m(x)
    prof = ops.profile_steps(lambda: m(x), steps=3)
    tot = sum(v["ms"] for v in prof.values())
    print(f"{name} {model_name} bs {B}: {tot:.3f} ms/step (sum of per-launch times), tokens {m.get_last_stats()['token_counts']}")
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        rate = v["work"] / (v["ms"] * 1e-3) if v["ms"] else 0
        unit = "TF/s" if k.startswith("gemm") or k == "attention" else "GB/s"
        print(f"    {k:14s} {v['ms']:7.3f} ms  {100 * v['ms'] / tot:5.1f} %  {v['launches']:3d} launches  {rate / (1e12 if unit == 'TF/s' else 1e9):8.1f} {unit}")
    del m, x
    torch.cuda.empty_cache()
