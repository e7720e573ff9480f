"""rajni_vit_b200 — B200-native (sm_100a) implementation of RAJNI-ViT's token-pruning
forward path behind the reference's API (rajni/__init__.py:1-2):

    from rajni_vit_b200 import RAJNIViTWrapper, evaluate_model

PyTorch provides device memory, streams and torch.distributed; all arithmetic runs in
hand-written CUDA kernels behind the C ABI of include/rajni_b200.h.
"""
from .checkpoint import load_checkpoint
from .eval import evaluate_model
from .wrapper import RAJNIAttention, RAJNIViTWrapper, compute_importance

__all__ = ["RAJNIViTWrapper", "RAJNIAttention", "compute_importance", "evaluate_model", "load_checkpoint"]
