"""EDM Heun sampler (reference: sample.py:30-66) driving the B200 engine.

Same signature and semantics as the reference `edm_sampler` (fp64 state, 2N-1 network evaluations, `randn_like`
consumed once per step even when S_churn = 0).  With a `maskdit_b200.EDMPrecond` network each evaluation is one
eval-mode engine pass at batch 2B with the classifier-free-guidance combine fused into the output kernel, and the
fp64 Euler/Heun state updates are single fused kernels.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def edm_sampler(net, latents, class_labels=None, cfg_scale=None, feat=None, randn_like=torch.randn_like,
                num_steps=18, sigma_min=0.002, sigma_max=80, rho=7, S_churn=0, S_min=0, S_max=float("inf"),
                S_noise=1):
    sigma_min = max(sigma_min, net.sigma_min)
    sigma_max = min(sigma_max, net.sigma_max)
    dev = latents.device
    # Karras schedule in fp64 on the host (sample.py:40-43); t_N = 0
    i = np.arange(num_steps, dtype=np.float64)
    t = (sigma_max ** (1 / rho) + i / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    t = np.concatenate([t, [0.0]])

    x_next = (latents.to(torch.float64) * t[0]).contiguous()
    d_cur = torch.empty_like(x_next)
    x32 = torch.empty(x_next.shape, dtype=torch.float32, device=dev)
    for k in range(num_steps):
        t_cur, t_next = float(t[k]), float(t[k + 1])
        gamma = min(S_churn / num_steps, np.sqrt(2) - 1) if S_min <= t_cur <= S_max else 0
        t_hat = t_cur + gamma * t_cur
        noise = randn_like(x_next)  # drawn every step, as in the reference (sample.py:53)
        if gamma > 0:
            x_hat = (x_next + float(np.sqrt(t_hat ** 2 - t_cur ** 2)) * S_noise * noise).contiguous()
        else:
            x_hat = x_next.clone()
        den = net(x_hat.float(), torch.tensor(t_hat, dtype=torch.float64, device=dev), class_labels, cfg_scale,
                  feat=feat)["x"].float().contiguous()
        ops.heun_update(0, x_hat, den, d_cur, x_next, x32, t_hat, t_next)          # Euler step (sample.py:56-58)
        if k < num_steps - 1:
            den = net(x32, torch.tensor(t_next, dtype=torch.float64, device=dev), class_labels, cfg_scale,
                      feat=feat)["x"].float().contiguous()
            ops.heun_update(1, x_hat, den, d_cur, x_next, x32, t_hat, t_next)      # 2nd-order correction (:61-64)
    return x_next
