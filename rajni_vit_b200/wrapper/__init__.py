"""Same public names as the reference's ``rajni.wrapper`` (rajni/wrapper/__init__.py:1-3)."""
from .attention import RAJNIAttention
from .importance import compute_importance
from .model import RAJNIViTWrapper

__all__ = ["RAJNIViTWrapper", "RAJNIAttention", "compute_importance"]
