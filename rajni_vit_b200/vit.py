"""A timm-attribute-compatible ViT used as the base model in tests and benches.

The reference drives a ``timm`` VisionTransformer (``/root/reference/rajni/run.py:89-92``)
but only ever touches a small duck-typed surface of it
(``/root/reference/rajni/wrapper/model.py:10,34-37,45-66`` and
``/root/reference/rajni/wrapper/attention.py:8-12``).  timm is not installed in
this image and there is no network, so this module provides a model with exactly
those attribute names and timm's ViT semantics:

    patch_embed.proj : Conv2d(3, C, k=16, s=16)  -> flatten(2).transpose(1, 2)
    cls_token [1,1,C], pos_embed [1,1+P,C], pos_drop
    blocks[i] : norm1, attn{num_heads, scale, qkv, proj, proj_drop}, norm2,
                mlp{fc1, act=GELU(erf), fc2}, ls1/ls2/drop_path1/drop_path2 = Identity
    norm, head

It is a plain eager PyTorch module.  It is NOT on the accelerated path: the
B200 wrapper reads its leaf parameters and runs its own kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

# name -> (embed_dim, depth, heads, img_size)
VIT_CONFIGS = {
    "vit_tiny_patch16_224": (192, 12, 3, 224),
    "vit_small_patch16_224": (384, 12, 6, 224),
    "vit_base_patch16_224": (768, 12, 12, 224),
    "vit_large_patch16_224": (1024, 24, 16, 224),
    "deit_base_patch16_384": (768, 12, 12, 384),
    # a 4-block, 17-token model small enough for golden fixtures
    "vit_micro_patch16_64": (128, 4, 2, 64),
}


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch, in_chans, dim):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch, patch)
        self.grid_size = (img_size // patch, img_size // patch)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)
        self.norm = nn.Identity()

    def forward(self, x):
        return self.norm(self.proj(x).flatten(2).transpose(1, 2))


class Attention(nn.Module):
    fused_attn = True

    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        x = x.transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x))


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden, dim)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.ls1 = nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = nn.Identity()
        self.drop_path2 = nn.Identity()

    def forward(self, x):
        x = x + self.drop_path1(self.ls1(self.attn(self.norm1(x))))
        x = x + self.drop_path2(self.ls2(self.mlp(self.norm2(x))))
        return x


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000,
                 embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0):
        super().__init__()
        self.num_classes = num_classes
        self.embed_dim = self.num_features = embed_dim
        self.global_pool = "token"
        self.num_prefix_tokens = 1
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n_tok = self.patch_embed.num_patches + 1
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_tok, embed_dim))
        self.pos_drop = nn.Dropout(0.0)
        self.patch_drop = nn.Identity()
        self.norm_pre = nn.Identity()
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.fc_norm = nn.Identity()
        self.head_drop = nn.Dropout(0.0)
        self.head = nn.Linear(embed_dim, num_classes)
        nn.init.normal_(self.cls_token, std=0.02)
        nn.init.normal_(self.pos_embed, std=0.02)

    def forward_features(self, x):
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        x = self.pos_drop(x + self.pos_embed)
        x = self.blocks(self.norm_pre(self.patch_drop(x)))
        return self.norm(x)

    def forward(self, x):
        x = self.forward_features(x)
        return self.head(self.head_drop(self.fc_norm(x[:, 0])))


def create_model(name: str, seed: int | None = 0, bf16_round: bool = True,
                 num_classes: int = 1000) -> VisionTransformer:
    """Random-init stand-in for ``timm.create_model(name)`` (no pretrained weights exist here).

    seed: torch CPU RNG seed for the init (manual_seed(0)).
    bf16_round: round every parameter to a bf16-representable fp32 value so that the
        fp32 oracle and the bf16 kernel path share bit-identical weights.
    """
    dim, depth, heads, img = VIT_CONFIGS[name]
    if seed is not None:
        gen_state = torch.random.get_rng_state()
        torch.manual_seed(seed)
    model = VisionTransformer(img_size=img, embed_dim=dim, depth=depth, num_heads=heads,
                              num_classes=num_classes)
    if seed is not None:
        torch.random.set_rng_state(gen_state)
    if bf16_round:
        with torch.no_grad():
            for p in model.parameters():
                p.copy_(p.to(torch.bfloat16).to(torch.float32))
    return model.eval()


def randomize_trained_like(model: VisionTransformer, seed: int = 0, outlier_channels: int = 4, outlier_scale: float = 20.0,
                           bf16_round: bool = True) -> VisionTransformer:
    """Give a random-init stand-in the parameter STATISTICS of a trained ViT (no pretrained weights exist offline):
    LayerNorm gains spread over [0.3, 2.5] and non-zero shifts, biases of O(0.5), and a few "massive activation"
    channels - residual-stream channels that every block's fc2 / proj pushes far from zero, with the matching large
    LayerNorm gains trained models show there.  Exercises the folded-LayerNorm path (W*gamma, b + W beta, E[x^2]-mean^2
    statistics at large |mean|/sigma) that gamma = 1, beta = 0 never touches."""
    g = torch.Generator().manual_seed(seed)
    C = model.embed_dim
    hot = torch.randperm(C, generator=g)[:outlier_channels]
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name == "norm.weight":
                p.copy_(0.3 + 2.2 * torch.rand(p.shape, generator=g))
            elif name.endswith("norm1.bias") or name.endswith("norm2.bias") or name == "norm.bias":
                p.copy_(0.3 * torch.randn(p.shape, generator=g))
            elif name.endswith(".bias"):
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
        for blk in model.blocks:
            blk.mlp.fc2.bias[hot] += outlier_scale * (0.5 + torch.rand(outlier_channels, generator=g))
            blk.attn.proj.bias[hot] -= 0.25 * outlier_scale * torch.rand(outlier_channels, generator=g)
            for nrm in (blk.norm1, blk.norm2):
                nrm.weight[hot] *= 0.1                      # trained models damp their massive channels at the next norm
        model.pos_embed[..., hot] += outlier_scale
        if bf16_round:
            for p in model.parameters():
                p.copy_(p.to(torch.bfloat16).to(torch.float32))
    return model
