"""TEST INFRASTRUCTURE ONLY — stand-in for `timm.models.vision_transformer.{PatchEmbed, Attention, Mlp}`.

The reference imports these three classes (models/maskdit.py:16) from timm, which is unpinned (Dockerfile:3) and
not installed here.  This module restates their published semantics (timm 0.6-0.9, Apache-2.0) so that
`/root/reference/models/maskdit.py` can be imported UNMODIFIED when generating golden vectors
(tests/golden/make_golden.py) and when timing the reference's CPU path.  Parameter names (`proj`, `qkv`, `fc1`,
`fc2`) are the checkpoint contract.  Nothing in the product path imports this file.
"""
import sys
import types

import torch
import torch.nn as nn


class PatchEmbed(nn.Module):
    """Conv2d(k=p, s=p) patchifier -> [B, L, D]; attrs num_patches / patch_size (tuple) are read by the reference."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, bias=True):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class Attention(nn.Module):
    """qkv Linear -> [B,N,3,H,dh] -> softmax(q k^T dh^-0.5) v -> proj Linear."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        attn = ((q * self.scale) @ k.transpose(-2, -1)).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


def install():
    """Register the stand-in as `timm.models.vision_transformer` (plus an `lmdb` stub for sample.py:14)."""
    if "timm.models.vision_transformer" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        vt = types.ModuleType("timm.models.vision_transformer")
        vt.PatchEmbed, vt.Attention, vt.Mlp = PatchEmbed, Attention, Mlp
        timm.models, models.vision_transformer = models, vt
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.vision_transformer": vt})
    for missing in ("lmdb",):
        if missing not in sys.modules:
            try:
                __import__(missing)
            except ImportError:
                sys.modules[missing] = types.ModuleType(missing)
