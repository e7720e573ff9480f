"""Seeded test cases shared by the golden generator, the CPU tests and the GPU tests."""
import numpy as np
import torch

# name -> (B, N, H, D, seed)
IMPORTANCE_CASES = {
    "small": (3, 23, 2, 8, 11),
    "tiny197": (4, 197, 3, 64, 12),
    "base197": (2, 197, 12, 64, 13),
    "base173": (2, 173, 12, 64, 14),
    "large87": (2, 87, 16, 64, 15),
    "deit577": (1, 577, 12, 64, 16),
}

# name -> (B, N, keep_ratio, seed)
SELECT_CASES = {
    "n197_r88": (5, 197, 0.88, 21),
    "n173_r88": (5, 173, 0.88, 22),
    "n121_r72": (5, 121, 0.72, 23),
    "n577_r80": (3, 577, 0.80, 24),
    "n17_r50": (4, 17, 0.50, 25),
    "n5_r01": (2, 5, 0.01, 26),      # keep clamps to 1
    "n33_r100": (2, 33, 1.0, 27),    # identity gather
}

README_SCHEDULE = {3: {"keep_ratio": 0.88, "update": True}, 4: {"keep_ratio": 0.88, "update": True},
                   7: {"keep_ratio": 0.8, "update": True}, 8: {"keep_ratio": 0.72, "update": True}}
# /root/reference/schedule.json with int keys (the JSON file itself has string keys)
C1_SCHEDULE = {3: {"keep_ratio": 0.95, "update": False}, 4: {"keep_ratio": 0.95, "update": True},
               5: {"keep_ratio": 0.85, "update": True}, 6: {"keep_ratio": 0.85, "update": True},
               7: {"keep_ratio": 0.95, "update": True}}
C3_SCHEDULE = {i: {"keep_ratio": 0.7, "update": True} for i in range(3, 12)}
C4_SCHEDULE = {i: {"keep_ratio": 0.9, "update": True} for i in range(24)}
MICRO_SCHEDULE = {1: {"keep_ratio": 0.75, "update": True}, 2: {"keep_ratio": 0.6, "update": False},
                  3: {"keep_ratio": 0.5}}

# BASELINE.json configs -> (model, schedule)
SCHEDULES = {
    "C1": ("vit_tiny_patch16_224", C1_SCHEDULE),
    "C2": ("vit_base_patch16_224", README_SCHEDULE),
    "C3": ("vit_small_patch16_224", C3_SCHEDULE),
    "C4": ("vit_large_patch16_224", C4_SCHEDULE),
    "C5": ("deit_base_patch16_384", README_SCHEDULE),
}

# name -> (model, schedule, batch, image seed)
E2E_CASES = {
    "micro": ("vit_micro_patch16_64", MICRO_SCHEDULE, 4, 31),
    "tiny_c1": ("vit_tiny_patch16_224", C1_SCHEDULE, 8, 1234),
}


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def make_qkv(B, N, H, D, seed):
    """bf16-representable fp32 qkv [B,N,3*H*D]; same tile goes to oracle (fp32) and kernel (bf16)."""
    g = torch.Generator().manual_seed(seed)
    return bf16_round(torch.randn(B, N, 3 * H * D, generator=g))


def make_scores(B, N, seed):
    """Distinct positive scores of the magnitude importance() produces (O(1/N))."""
    g = torch.Generator().manual_seed(seed)
    s = torch.rand(B, N, generator=g) / N
    assert all(len(torch.unique(r)) == N for r in s)
    return s


def make_images(B, size, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, size, size, generator=g)


def checksum(t: torch.Tensor) -> float:
    return float(t.double().abs().sum())


def npz(path):
    return dict(np.load(path))
