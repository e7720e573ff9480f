"""Generate tests/golden/* by running the UNMODIFIED reference from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/make_golden.py

The reference has no tests or fixtures of its own, so these
files are the pins for the oracle: every array below was produced by
``rajni.wrapper.{compute_importance, RAJNIAttention, RAJNIViTWrapper}`` imported
as-is, applied to the stand-in ViT of ``rajni_vit_b200/vit.py``.
Inputs are regenerated from seeds by the tests; each fixture also stores a
checksum of the seeded inputs/weights so a silent RNG change is caught.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from rajni.wrapper import RAJNIAttention, RAJNIViTWrapper, compute_importance  # noqa: E402  (the reference)
from rajni_vit_b200.vit import create_model  # noqa: E402
from tests.cases import (IMPORTANCE_CASES, SELECT_CASES, SCHEDULES, E2E_CASES,  # noqa: E402
                         make_qkv, make_scores, make_images, checksum)

OUT = os.path.join(ROOT, "tests", "golden")


def kat():
    qkv = torch.tensor([[[1, 0, 0, 1, 1, 0, 0, 1, 1, 2, 3, 4],
                         [0, 0, 0, 0, 2, 0, 0, 2, 0, 0, 0, 0],
                         [0, 0, 0, 0, 0, 2, 2, 0, 4, 0, 0, 4],
                         [0, 0, 0, 0, -2, 0, 0, -2, 2, 2, 2, 2]]], dtype=torch.float64)
    s = compute_importance(qkv, 2)
    np.savez(os.path.join(OUT, "importance_kat.npz"), qkv=qkv.numpy(), score=s.numpy())
    print("KAT", s)


def importance_rand():
    out = {}
    for name, (B, N, H, D, seed) in IMPORTANCE_CASES.items():
        qkv = make_qkv(B, N, H, D, seed)
        out[name + "_sum"] = checksum(qkv)
        out[name + "_f32"] = compute_importance(qkv, H).numpy()
        out[name + "_f64"] = compute_importance(qkv.double(), H).numpy()
    np.savez(os.path.join(OUT, "importance_rand.npz"), **out)


def select_cases():
    out = {}
    for name, (B, N, ratio, seed) in SELECT_CASES.items():
        scores = make_scores(B, N, seed)
        keep = max(1, int(ratio * (N - 1)))
        _, idx = torch.topk(scores[:, 1:], keep, dim=1)           # attention.py:34-39, verbatim semantics
        idx = torch.sort(idx, dim=1).values
        keep_idx = torch.cat([torch.zeros((B, 1), dtype=torch.long), idx + 1], dim=1)
        out[name + "_sum"] = checksum(scores)
        out[name + "_idx"] = keep_idx.numpy().astype(np.int32)
    np.savez(os.path.join(OUT, "select_cases.npz"), **out)


def trajectories():
    res = {}
    for cfg, (model_name, sched) in SCHEDULES.items():
        base = create_model(model_name, seed=0)
        m = RAJNIViTWrapper(base, sched).eval()
        size = base.patch_embed.img_size[0]
        with torch.no_grad():
            m(torch.randn(1, 3, size, size))
        res[cfg] = m.get_last_stats()["token_counts"]
        print(cfg, res[cfg])
    # string-key schedule => nothing pruned
    base = create_model("vit_tiny_patch16_224", seed=0)
    m = RAJNIViTWrapper(base, json.load(open("/root/reference/schedule.json"))).eval()
    with torch.no_grad():
        m(torch.randn(1, 3, 224, 224))
    res["C1_string_keys"] = m.get_last_stats()["token_counts"]
    json.dump(res, open(os.path.join(OUT, "trajectories.json"), "w"), indent=1)


def e2e():
    for name, (model_name, sched, batch, seed) in E2E_CASES.items():
        base = create_model(model_name, seed=0)
        wsum = checksum(torch.cat([p.detach().flatten() for p in base.parameters()]))
        images = make_images(batch, base.patch_embed.img_size[0], seed)
        m = RAJNIViTWrapper(base, sched).eval()
        rec = {}

        def hook(idx):
            def fn(mod, args, output):
                out, keep_idx, nxt = output
                rec[f"b{idx}_xnorm"] = args[0].numpy()
                if len(args) > 1 and args[1] is not None:
                    rec[f"b{idx}_prev"] = args[1].numpy()
                rec[f"b{idx}_out"] = out.numpy()
                rec[f"b{idx}_keep_idx"] = keep_idx.numpy().astype(np.int32)
                rec[f"b{idx}_next"] = nxt.numpy()
            return fn

        block_out = {}
        for i, blk in enumerate(m.blocks):
            if isinstance(blk.attn, RAJNIAttention):
                blk.attn.register_forward_hook(hook(i))
        with torch.no_grad():
            logits = m(images)
        keep = {k: v for k, v in rec.items()
                if k.endswith("_keep_idx") or k.endswith("_next") or name.startswith("micro")}
        np.savez_compressed(os.path.join(OUT, f"e2e_{name}.npz"), logits=logits.numpy(),
                            token_counts=np.array(m.get_last_stats()["token_counts"]),
                            weight_sum=wsum, image_sum=checksum(images), **keep, **block_out)
        print(name, m.get_last_stats(), logits.std().item())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    kat()
    importance_rand()
    select_cases()
    trajectories()
    e2e()
