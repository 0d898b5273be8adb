"""GPU: further reference goldens on the CUDA path — another geometry (DiT-B/4: patch 4, 12 heads of 64, three of
four patches masked) and the stochastic (S_churn > 0) sampler."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from test_model_gpu import GoldenLoss, build, load, rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu


def test_b4_patch4_mask75_loss_and_grads_vs_reference_golden():
    g = load("b4_train_mask75")
    net, cfg, _ = build("DiT-B/4", 16, 7)
    net.train()
    lf = GoldenLoss(g)
    loss = lf(net, g["images"].cuda(), g["labels"].cuda(), mask_ratio=0.75, mae_loss_coef=0.1)
    for k in ("mask", "ids_keep", "ids_restore"):
        assert torch.equal(lf.last_mask_dict[k].cpu(), g[k]), k
    assert torch.allclose(loss.cpu(), g["loss"], rtol=1e-2), (loss, g["loss"])
    loss.mean().backward()
    for k, p in net.named_parameters():
        key = f"gnorm/{k}"
        if key not in g:
            continue
        gn, ref = p.grad.double().norm().item(), float(g[key])
        assert abs(gn - ref) <= 3e-2 * ref + 1e-7, (k, gn, ref)
        if f"grad/{k}" in g and ref > 0:
            assert rel_l2(p.grad, g[f"grad/{k}"]) <= 3e-2, k


def test_sampler_with_churn_vs_reference_golden():
    from maskdit_b200.sampler import edm_sampler
    g = load("s2_sampler_churn")
    net, cfg, _ = build()
    net.eval()
    noises = [n.cuda() for n in g["noises"]]
    calls = []
    orig = net.forward

    def spy(x, s, *a, **k):
        calls.append(float(s))
        return orig(x, s, *a, **k)

    net.forward = spy
    with torch.no_grad():
        z = edm_sampler(net, g["latents"].cuda(), g["labels"].cuda(), randn_like=lambda x: noises.pop(0),
                        num_steps=int(g["num_steps"]), S_churn=float(g["S_churn"]), S_min=float(g["S_min"]),
                        S_max=float(g["S_max"]), S_noise=float(g["S_noise"]))
    net.forward = orig
    assert len(calls) == 15 and not noises
    np.testing.assert_allclose(np.array(calls), g["sampler_sigmas"].numpy(), rtol=1e-12)
    assert z.dtype == torch.float64 and rel_l2(z, g["z"]) <= 2e-2
