"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference's SD-VAE DECODE path (autoencoder.py:306-453), the
sampler tail that `generate_with_net` runs on every batch of latents (sample.py:275: `images = vae.decode(z)`).
The product path (`maskdit_b200/vae.py`) never imports it; it is pinned against the unmodified reference module
(`autoencoder.FrozenAutoencoderKL`) by tests/golden/make_golden.py::vae_case + tests/test_oracle_golden.py.

Parameters are a flat dict keyed like `FrozenAutoencoderKL.state_dict()` (`decoder.*`, `post_quant_conv.*`).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CH, CH_MULT, NUM_RES_BLOCKS, Z_CH, OUT_CH = 128, (1, 2, 4, 4), 2, 4, 3   # autoencoder.py:466-479 (get_model)


def decoder_param_shapes():
    """Key -> shape of `post_quant_conv` + `decoder` (autoencoder.py:307-377, :424), in registration order."""
    out = {"post_quant_conv.weight": (Z_CH, 4, 1, 1), "post_quant_conv.bias": (Z_CH,)}

    def res(p, cin, cout):
        out[f"{p}.norm1.weight"], out[f"{p}.norm1.bias"] = (cin,), (cin,)
        out[f"{p}.conv1.weight"], out[f"{p}.conv1.bias"] = (cout, cin, 3, 3), (cout,)
        out[f"{p}.norm2.weight"], out[f"{p}.norm2.bias"] = (cout,), (cout,)
        out[f"{p}.conv2.weight"], out[f"{p}.conv2.bias"] = (cout, cout, 3, 3), (cout,)
        if cin != cout:
            out[f"{p}.nin_shortcut.weight"], out[f"{p}.nin_shortcut.bias"] = (cout, cin, 1, 1), (cout,)

    block_in = CH * CH_MULT[-1]
    out["decoder.conv_in.weight"], out["decoder.conv_in.bias"] = (block_in, Z_CH, 3, 3), (block_in,)
    res("decoder.mid.block_1", block_in, block_in)
    p = "decoder.mid.attn_1"
    out[f"{p}.norm.weight"], out[f"{p}.norm.bias"] = (block_in,), (block_in,)
    for n in ("q", "k", "v", "proj_out"):
        out[f"{p}.{n}.weight"], out[f"{p}.{n}.bias"] = (block_in, block_in, 1, 1), (block_in,)
    res("decoder.mid.block_2", block_in, block_in)
    ups = {}
    for lvl in reversed(range(len(CH_MULT))):
        block_out = CH * CH_MULT[lvl]
        for i in range(NUM_RES_BLOCKS + 1):
            ups[(lvl, i)] = (block_in, block_out)
            block_in = block_out
        ups[(lvl, "up")] = block_in
    for lvl in range(len(CH_MULT)):               # ModuleList order: up.0 .. up.3 (decoder prepends, :370)
        for i in range(NUM_RES_BLOCKS + 1):
            res(f"decoder.up.{lvl}.block.{i}", *ups[(lvl, i)])
        if lvl != 0:
            c = ups[(lvl, "up")]
            out[f"decoder.up.{lvl}.upsample.conv.weight"] = (c, c, 3, 3)
            out[f"decoder.up.{lvl}.upsample.conv.bias"] = (c,)
    out["decoder.norm_out.weight"], out["decoder.norm_out.bias"] = (CH,), (CH,)
    out["decoder.conv_out.weight"], out["decoder.conv_out.bias"] = (OUT_CH, CH, 3, 3), (OUT_CH,)
    return out


def make_vae_state_dict(seed=3):
    """Deterministic stand-in weights (the published autoencoder_kl.pth is not available offline): conv / linear weights
    N(0, 1/fan_in), GroupNorm scales 1 + 0.1 N(0,1), biases 0.05 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in decoder_param_shapes().items():
        if k.endswith("weight") and len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            sd[k] = torch.randn(shp, generator=g) * fan_in ** -0.5
        elif k.endswith("weight"):
            sd[k] = 1 + 0.1 * torch.randn(shp, generator=g)
        else:
            sd[k] = 0.05 * torch.randn(shp, generator=g)
    return sd


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[f"{p}.weight"], sd[f"{p}.bias"], eps=1e-6)          # Normalize, autoencoder.py:34-35


def _swish(x):
    return x * torch.sigmoid(x)                                                          # nonlinearity, :29-31


def _conv(sd, p, x, pad):
    return F.conv2d(x, sd[f"{p}.weight"], sd[f"{p}.bias"], padding=pad)


def _resblock(sd, p, x):
    """ResnetBlock.forward with temb = None (autoencoder.py:117-137)."""
    h = _conv(sd, f"{p}.conv1", _swish(_gn(sd, f"{p}.norm1", x)), 1)
    h = _conv(sd, f"{p}.conv2", _swish(_gn(sd, f"{p}.norm2", h)), 1)
    if f"{p}.nin_shortcut.weight" in sd:
        x = _conv(sd, f"{p}.nin_shortcut", x, 0)
    return x + h


def _attn(sd, p, x):
    """AttnBlock.forward (autoencoder.py:174-198): single head over h*w positions, head_dim = channels."""
    h_ = _gn(sd, f"{p}.norm", x)
    q, k, v = (_conv(sd, f"{p}.{n}", h_, 0) for n in ("q", "k", "v"))
    b, c, h, w = q.shape
    w_ = torch.bmm(q.reshape(b, c, h * w).permute(0, 2, 1), k.reshape(b, c, h * w)) * (int(c) ** -0.5)
    w_ = torch.softmax(w_, dim=2)
    h_ = torch.bmm(v.reshape(b, c, h * w), w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(sd, f"{p}.proj_out", h_, 0)


def decode(sd, z, scale_factor=0.18215):
    """FrozenAutoencoderKL.decode (autoencoder.py:449-453) + Decoder.forward (:379-412).  z [B,4,h,w] -> [B,3,8h,8w]."""
    z = _conv(sd, "post_quant_conv", (1.0 / scale_factor) * z, 0)
    h = _conv(sd, "decoder.conv_in", z, 1)
    h = _resblock(sd, "decoder.mid.block_1", h)
    h = _attn(sd, "decoder.mid.attn_1", h)
    h = _resblock(sd, "decoder.mid.block_2", h)
    for lvl in reversed(range(len(CH_MULT))):
        for i in range(NUM_RES_BLOCKS + 1):
            h = _resblock(sd, f"decoder.up.{lvl}.block.{i}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")                       # Upsample.forward, :49-53
            h = _conv(sd, f"decoder.up.{lvl}.upsample.conv", h, 1)
    return _conv(sd, "decoder.conv_out", _swish(_gn(sd, "decoder.norm_out", h)), 1)


def to_uint8(images):
    """sample.py:287."""
    return images.clone().add_(1).mul(127.5).clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1)
