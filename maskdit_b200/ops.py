"""Thin Python wrappers over the C ABI (one function per kernel entry point).  CUDA tensors only; no fallback."""
from __future__ import annotations

import torch

from . import _lib as L
from ._lib import (ACT_NONE, ACT_SILU, EPI_ATOMIC, EPI_DGELU, EPI_GATE_RESID, EPI_GELU, EPI_STORE, check, gemm, lib,
                   ptr, stream_ptr)

bf16, f32 = torch.bfloat16, torch.float32


def _c(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise L.MdtError("maskdit_b200 kernels need CUDA tensors (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise L.MdtError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise L.MdtError("expected a contiguous tensor")
    return t


def mask_indices(noise, len_keep):
    """get_mask for a given noise tensor (models/maskdit.py:88-113) -> dict like the reference's mask_dict."""
    _c(noise, f32)
    B, Lt = noise.shape
    ids_keep = torch.empty(B, len_keep, dtype=torch.int64, device=noise.device)
    ids_restore = torch.empty(B, Lt, dtype=torch.int64, device=noise.device)
    mask = torch.empty(B, Lt, dtype=f32, device=noise.device)
    check(lib().mdt_mask_indices(ptr(noise), B, Lt, len_keep, ptr(ids_keep), ptr(ids_restore), ptr(mask),
                                 stream_ptr()), "mdt_mask_indices")
    return {"mask": mask, "ids_keep": ids_keep, "ids_restore": ids_restore}


def patch_embed(x, sigma, sigma_data, W, bias, pos, ids_keep, p, D):
    _c(x, f32), _c(W, f32), _c(bias, f32), _c(pos, f32), _c(ids_keep, torch.int64), _c(sigma, f32)
    B, C, R, _ = x.shape
    T = ids_keep.shape[1] if ids_keep is not None else (R // p) ** 2
    out = torch.empty(B, T, D, dtype=f32, device=x.device)
    check(lib().mdt_patch_embed(ptr(x), ptr(sigma), sigma_data, ptr(W), ptr(bias), ptr(pos), ptr(ids_keep), ptr(out),
                                B, C, R, p, D, T, stream_ptr()), "mdt_patch_embed")
    return out


def patch_embed_bwd(x, sigma, sigma_data, ids_keep, g, gW, gb, p):
    _c(x, f32), _c(g, f32), _c(gW, f32), _c(gb, f32)
    B, C, R, _ = x.shape
    T, D = g.shape[1], g.shape[2]
    check(lib().mdt_patch_embed_bwd(ptr(x), ptr(sigma), sigma_data, ptr(ids_keep), ptr(g), ptr(gW), ptr(gb), B, C, R,
                                    p, D, T, stream_ptr()), "mdt_patch_embed_bwd")


def timestep_freq(sigma, dim=256):
    _c(sigma, f32)
    out = torch.empty(sigma.numel(), dim, dtype=bf16, device=sigma.device)
    check(lib().mdt_timestep_freq(ptr(sigma), sigma.numel(), dim, ptr(out), stream_ptr()), "mdt_timestep_freq")
    return out


def silu(a, b=None, want_sum=False):
    _c(a, f32), _c(b, f32)
    out = torch.empty(a.shape, dtype=bf16, device=a.device)
    s = torch.empty_like(a) if want_sum else None
    check(lib().mdt_silu(ptr(a), ptr(b), ptr(s), ptr(out), a.numel(), stream_ptr()), "mdt_silu")
    return (out, s) if want_sum else out


def silu_bwd(dy, x, want_f32=True, want_bf16=True):
    _c(dy, f32), _c(x, f32)
    d32 = torch.empty_like(x) if want_f32 else None
    d16 = torch.empty(x.shape, dtype=bf16, device=x.device) if want_bf16 else None
    check(lib().mdt_silu_bwd(ptr(dy), ptr(x), ptr(d32), ptr(d16), x.numel(), stream_ptr()), "mdt_silu_bwd")
    return d32, d16


def cast_bf16(x, out=None):
    _c(x, f32)
    if out is None:
        out = torch.empty(x.shape, dtype=bf16, device=x.device)
    check(lib().mdt_cast_f32_bf16(ptr(x), ptr(out), x.numel(), stream_ptr()), "mdt_cast_f32_bf16")
    return out


def colsum(x, out, M=None, N=None, ld=None):
    """out[N] += sum_m x[m, n]"""
    M = x.shape[0] if M is None else M
    N = x.shape[1] if N is None else N
    ld = x.stride(0) if ld is None else ld
    fn = lib().mdt_colsum_bf16 if x.dtype == bf16 else lib().mdt_colsum_f32
    check(fn(ptr(x), M, N, ld, ptr(out), stream_ptr()), "mdt_colsum")


def ln_modulate(x, shift, scale, ld_mod, rows_per_group, M, D, save_stats=True, eps=1e-6):
    out = torch.empty(M, D, dtype=bf16, device=x.device)
    mean = torch.empty(M, dtype=f32, device=x.device) if save_stats else None
    rstd = torch.empty(M, dtype=f32, device=x.device) if save_stats else None
    check(lib().mdt_ln_modulate(ptr(x), ptr(shift), ptr(scale), ld_mod, rows_per_group, ptr(out), ptr(mean),
                                ptr(rstd), M, D, eps, stream_ptr()), "mdt_ln_modulate")
    return out, mean, rstd


def ln_modulate_bwd(dxmod, x, mean, rstd, scale, ld_mod, rows_per_group, g, accumulate, dshift, dscale, ld_dmod, M, D):
    check(lib().mdt_ln_modulate_bwd(ptr(dxmod), ptr(x), ptr(mean), ptr(rstd), ptr(scale), ld_mod, rows_per_group,
                                    ptr(g), int(accumulate), ptr(dshift), ptr(dscale), ld_dmod, M, D, stream_ptr()),
          "mdt_ln_modulate_bwd")


def ln_modulate_bwd_gate(dxmod, x, mean, rstd, scale, ld_mod, rows_per_group, g, accumulate, dshift, dscale, ld_dmod,
                         M, D, gate_next=None):
    """LN-modulate backward; with `gate_next = (y, gate, ld_gate, dgate, ld_dgate, dbias)` also the gate backward of
    the branch that consumes the finished residual gradient (one pass over g).  Returns dy (bf16) or None."""
    dy = None
    y = gate = dgate = dbias = None
    ld_gate = ld_dgate = 0
    if gate_next is not None:
        y, gate, ld_gate, dgate, ld_dgate, dbias = gate_next
        dy = torch.empty(M, D, dtype=bf16, device=g.device)
    check(lib().mdt_ln_modulate_bwd_gate(ptr(dxmod), ptr(x), ptr(mean), ptr(rstd), ptr(scale), ld_mod, rows_per_group,
                                         ptr(g), int(accumulate), ptr(dshift), ptr(dscale), ld_dmod, ptr(y), ptr(gate),
                                         ld_gate, ptr(dy), ptr(dgate), ld_dgate, ptr(dbias), M, D, stream_ptr()),
          "mdt_ln_modulate_bwd_gate")
    return dy


def gate_bwd(g, y, gate, ld_gate, rows_per_group, dgate, ld_dgate, dbias, M, D):
    dy = torch.empty(M, D, dtype=bf16, device=g.device)
    check(lib().mdt_gate_bwd(ptr(g), ptr(y), ptr(gate), ld_gate, rows_per_group, ptr(dy), ptr(dgate), ld_dgate,
                             ptr(dbias), M, D, stream_ptr()), "mdt_gate_bwd")
    return dy


def attention_fwd(qkv, B, T, H, dh, need_lse=True):
    _c(qkv, bf16)
    out = torch.empty(B * T, H * dh, dtype=bf16, device=qkv.device)
    lse = torch.empty(2, B, H, T, dtype=f32, device=qkv.device) if need_lse else None  # [1] = scratch for delta
    check(lib().mdt_attention_fwd(ptr(qkv), ptr(out), ptr(lse), B, T, H, dh, stream_ptr()), "mdt_attention_fwd")
    return out, lse


def attention_bwd(qkv, out, dout, lse, B, T, H, dh):
    _c(qkv, bf16), _c(out, bf16), _c(dout, bf16), _c(lse, f32)
    dqkv = torch.empty_like(qkv)
    check(lib().mdt_attention_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dqkv), B, T, H, dh, stream_ptr()),
          "mdt_attention_bwd", 2)
    return dqkv


def unmask_tokens(u, mask_token, pos, ids_restore, B, T, Lt, D):
    out = torch.empty(B, Lt, D, dtype=f32, device=u.device)
    check(lib().mdt_unmask_tokens(ptr(u), ptr(mask_token), ptr(pos), ptr(ids_restore), ptr(out), B, T, Lt, D,
                                  stream_ptr()), "mdt_unmask_tokens")
    return out


def unmask_tokens_bwd(g, ids_restore, dmask_token, B, T, Lt, D):
    du = torch.empty(B * T, D, dtype=bf16, device=g.device)
    check(lib().mdt_unmask_tokens_bwd(ptr(g), 0, ptr(ids_restore), ptr(du), ptr(dmask_token), B, T, Lt, D,
                                      stream_ptr()), "mdt_unmask_tokens_bwd")
    return du


def edm_loss(F, xin, y, sigma, mask, gl, sigma_data, mae_coef, p, want_D=False, want_dF=True):
    B, C, R, _ = xin.shape
    loss = torch.empty(B, dtype=f32, device=xin.device)
    Dx = torch.empty_like(xin) if want_D else None
    dF = torch.empty(F.shape, dtype=bf16, device=xin.device) if want_dF else None
    check(lib().mdt_edm_loss(ptr(F), ptr(xin), ptr(y), ptr(sigma), ptr(mask), ptr(gl), sigma_data, mae_coef,
                             ptr(loss), ptr(Dx), ptr(dF), B, C, R, p, stream_ptr()), "mdt_edm_loss")
    return loss, Dx, dF


def step_front(moments, eps, rnd_normal, noise_unit, labels=None, drop_u=None, drop_prob=0.0, scale_factor=0.18215,
               P_mean=-1.2, P_std=1.2):
    """moments -> latent, label dropout (in place on `labels`), sigma draw, noise injection: one launch.
    Returns (y, yn, sigma)."""
    _c(moments, f32), _c(eps, f32), _c(rnd_normal, f32), _c(noise_unit, f32), _c(labels, f32), _c(drop_u, f32)
    B, C2, R, _ = moments.shape
    C = C2 // 2
    y = torch.empty(B, C, R, R, dtype=f32, device=moments.device)
    yn = torch.empty_like(y)
    sigma = torch.empty(B, dtype=f32, device=moments.device)
    nc = labels.shape[1] if labels is not None else 0
    check(lib().mdt_step_front(ptr(moments), ptr(eps), ptr(rnd_normal), ptr(noise_unit), ptr(drop_u), drop_prob,
                               scale_factor, P_mean, P_std, ptr(y), ptr(yn), ptr(sigma),
                               ptr(labels) if drop_u is not None else 0, B, C, R, nc, stream_ptr()), "mdt_step_front")
    return y, yn, sigma


def edm_precond_out(F, xin, sigma, sigma_data, p):
    B, C, R, _ = xin.shape
    Dx = torch.empty_like(xin)
    check(lib().mdt_edm_precond_out(ptr(F), ptr(xin), ptr(sigma), sigma_data, ptr(Dx), B, C, R, p, stream_ptr()),
          "mdt_edm_precond_out")
    return Dx


def edm_precond_out_bwd(gD, sigma, sigma_data, p):
    B, C, R, _ = gD.shape
    Lt = (R // p) ** 2
    dF = torch.empty(B * Lt, p * p * C, dtype=bf16, device=gD.device)
    check(lib().mdt_edm_precond_out_bwd(ptr(gD), ptr(sigma), sigma_data, ptr(dF), B, C, R, p, stream_ptr()),
          "mdt_edm_precond_out_bwd")
    return dF


def cfg_precond_out(F, xin, sigma, sigma_data, cfg_scale, p):
    B, C, R, _ = xin.shape
    Dx = torch.empty_like(xin)
    check(lib().mdt_cfg_precond_out(ptr(F), ptr(xin), ptr(sigma), sigma_data, cfg_scale, ptr(Dx), B, C, R, p,
                                    stream_ptr()), "mdt_cfg_precond_out")
    return Dx


def heun_update(mode, x_hat, denoised, d_cur, x_next, x_next_f32, t_hat, t_next):
    check(lib().mdt_heun_update(mode, ptr(x_hat), ptr(denoised), ptr(d_cur), ptr(x_next), ptr(x_next_f32),
                                float(t_hat), float(t_next), x_hat.numel(), stream_ptr()), "mdt_heun_update")


def lincomb_f64(a, x, b=0.0, y=None, c=0.0, z=None, out=None, out_f32=None, f32_scale=1.0):
    """out = a*x + b*y + c*z (x, y fp64; z fp32), out_f32 = float(out * f32_scale).  Returns (out, out_f32)."""
    _c(x, torch.float64), _c(y, torch.float64), _c(z, f32), _c(out, torch.float64), _c(out_f32, f32)
    check(lib().mdt_lincomb_f64(float(a), ptr(x), float(b), ptr(y), float(c), ptr(z), ptr(out), ptr(out_f32),
                                float(f32_scale), x.numel(), stream_ptr()), "mdt_lincomb_f64")
    return out, out_f32


def to_uint8_nhwc(img):
    """[B,C,H,W] f32 in [-1,1] -> uint8 [B,H,W,C] (sample.py:287)."""
    _c(img, f32)
    B, C, H, W = img.shape
    out = torch.empty(B, H, W, C, dtype=torch.uint8, device=img.device)
    check(lib().mdt_to_uint8_nhwc(ptr(img), ptr(out), B, C, H, W, stream_ptr()), "mdt_to_uint8_nhwc")
    return out


def adamw_ema(w, g, m, v, ema, w16, n, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0,
              ema_decay=0.9999, grad_scale=1.0, max_blocks=0):
    fn = lib().mdt_adamw_ema_g16 if g.dtype == bf16 else lib().mdt_adamw_ema   # bf16: all-reduced bf16 gradients
    check(fn(ptr(w), ptr(g), ptr(m), ptr(v), ptr(ema), ptr(w16), n, lr, beta1, beta2, eps, weight_decay, step,
             ema_decay, grad_scale, max_blocks, stream_ptr()), "mdt_adamw_ema")
