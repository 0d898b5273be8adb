"""CPU: pin the oracle (oracle/maskdit_oracle.py) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py ran /root/reference through the timm stand-in).  fp32 vs fp32: tight tolerances."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import maskdit_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SMALL = O.Cfg(model_type="DiT-S/2", img_resolution=8, num_classes=10)


def load(name):
    return {k: v for k, v in np.load(os.path.join(GOLD, name + ".npz")).items()}


def t(a):
    return torch.from_numpy(np.asarray(a))


def test_tables_match_reference():
    g = load("tables")
    for D, grid in ((1152, 16), (512, 16), (384, 4), (512, 4), (1152, 32)):
        step = max(1, grid * grid // 16)
        np.testing.assert_allclose(O.sincos_pos_embed(D, grid).numpy()[::step], g[f"pos_{D}_{grid}"], atol=1e-6)
    np.testing.assert_allclose(O.timestep_embedding(t(g["tfreq_in"]), 256).numpy(), g["tfreq"], atol=1e-6)
    for L, r in ((256, 0.5), (1024, 0.5), (256, 0.75), (16, 0.5)):
        md = O.mask_from_noise(t(g[f"mask_noise_{L}_{r}"]), r)
        for k, v in md.items():
            assert np.array_equal(v.numpy(), g[f"mask_{k}_{L}_{r}"]), (L, r, k)


B4 = O.Cfg(model_type="DiT-B/4", img_resolution=16, num_classes=7)  # patch 4, 12 heads of 64, 16 patches


XL32 = O.Cfg(model_type="DiT-XL/2", img_resolution=32, num_classes=1000)
XL64 = O.Cfg(model_type="DiT-XL/2", img_resolution=64, num_classes=1000)


@pytest.mark.parametrize("name,CFG", [("s2_train_mask", SMALL), ("s2_train_nomask", SMALL), ("b4_train_mask75", B4),
                                      ("xl2_c1_grads", XL32), ("xl2_r64_grads", XL64),
                                      ("s2_uncond_mask30", O.Cfg(model_type="DiT-S/2", img_resolution=32, num_classes=0))])
def test_train_loss_and_grads_match_reference(name, CFG):
    g = load(name)
    SMALL = CFG  # noqa: N806 - the body below is written against the small config's name
    sd = {k: v.requires_grad_(not k.endswith("pos_embed")) for k, v in O.make_state_dict(SMALL, 1).items()}
    mr = float(g["mask_ratio"])
    md = O.mask_from_noise(t(g["mask_noise"]), mr) if mr > 0 else None
    if md is not None:
        for k in ("mask", "ids_keep", "ids_restore"):
            assert np.array_equal(md[k].numpy(), g[k])
    labels = t(g["labels"]) if "labels" in g else None           # class-unconditional golden: labels None
    loss, D = O.edm_loss(sd, SMALL, t(g["images"]), labels, t(g["rnd_normal"]), t(g["noise_unit"]), md,
                         SMALL.mae_loss_coef)
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(D.detach().numpy(), g["D"], rtol=1e-4, atol=2e-5)
    loss.mean().backward()
    checked = 0
    for k, v in g.items():
        if k.startswith("grad/"):
            gg = sd[k[5:]].grad  # None == zero gradient (reference: "+ 0 * sum(mask_token)", loss.py:57-58)
            gg = np.zeros_like(v) if gg is None else gg.numpy()
            np.testing.assert_allclose(gg, v, rtol=2e-3, atol=1e-6, err_msg=k)
            checked += 1
        elif k.startswith("gnorm/"):
            got = sd[k[6:]].grad
            got = 0.0 if got is None else got.double().norm().item()
            assert abs(got - float(v)) <= 2e-4 * (float(v) + 1e-9) + 1e-9, (k, got, float(v))
            checked += 1
        elif k.startswith("gslice/"):
            gg = sd[k[7:]].grad
            np.testing.assert_allclose(gg.reshape(gg.shape[0], -1)[:4, :8].numpy(), v, rtol=2e-3, atol=1e-6)
    assert checked > 100


def test_eval_cfg_and_sampler_match_reference():
    g = load("s2_eval")
    sd = O.make_state_dict(SMALL, 1)
    with torch.no_grad():
        plain = O.edm_precond(sd, SMALL, t(g["images"]), t(g["sigma"]), t(g["labels"]), training=False)
        np.testing.assert_allclose(plain.numpy(), g["D_plain"], rtol=1e-4, atol=2e-5)
        cfg = O.edm_precond(sd, SMALL, t(g["images"]), torch.tensor(1.7, dtype=torch.float64), t(g["labels"]),
                            cfg_scale=1.5, training=False)
        np.testing.assert_allclose(cfg.numpy(), g["D_cfg"], rtol=1e-4, atol=2e-5)
        lab = t(g["labels"])
        z, evals = O.edm_sampler(lambda x, s: O.edm_precond(sd, SMALL, x, s, lab, cfg_scale=1.5, training=False),
                                 t(g["latents"]))
    assert len(evals) == 35  # 2N-1 network evaluations
    np.testing.assert_allclose(np.array(evals), g["sampler_sigmas"], rtol=1e-12)
    np.testing.assert_allclose(z.numpy(), g["z"], rtol=1e-3, atol=1e-4)


def test_sampler_with_churn_matches_reference():
    """Stochastic sampler (S_churn > 0, sample.py:50-53): sigma sequence incl. the inflated t_hat, noise injection."""
    g = load("s2_sampler_churn")
    sd = O.make_state_dict(SMALL, 1)
    noises = [t(n) for n in g["noises"]]
    lab = t(g["labels"])
    with torch.no_grad():
        z, evals = O.edm_sampler(lambda x, s: O.edm_precond(sd, SMALL, x, s, lab, training=False), t(g["latents"]),
                                 num_steps=int(g["num_steps"]), S_churn=float(g["S_churn"]), S_min=float(g["S_min"]),
                                 S_max=float(g["S_max"]), S_noise=float(g["S_noise"]),
                                 randn_like=lambda x: noises.pop(0))
    assert len(evals) == 15 and not noises
    np.testing.assert_allclose(np.array(evals), g["sampler_sigmas"], rtol=1e-12)
    assert evals[1] < evals[2]  # the churned t_hat of step 1 exceeds its t_cur (the previous evaluation's t_next)
    np.testing.assert_allclose(z.numpy(), g["z"], rtol=1e-3, atol=1e-4)


def test_xl2_config1_forward_matches_reference():
    """BASELINE config 1: MaskDiT-XL/2 single forward, batch 2, 32x32x4 latents, mask_ratio 0.5, fp32 on CPU."""
    g = load("xl2_c1_fwd")
    cfg = O.Cfg(model_type="DiT-XL/2", img_resolution=32, num_classes=1000)
    sd = O.make_state_dict(cfg, 1)
    assert sum(v.numel() for v in sd.values()) == 730_541_200
    md = O.mask_from_noise(t(g["mask_noise"]), 0.5)
    with torch.no_grad():
        loss, D = O.edm_loss(sd, cfg, t(g["images"]), t(g["labels"]), t(g["rnd_normal"]), t(g["noise_unit"]), md,
                             cfg.mae_loss_coef)
    np.testing.assert_allclose(loss.numpy(), g["loss"], rtol=1e-4)
    np.testing.assert_allclose(D.numpy(), g["D"], rtol=1e-3, atol=1e-4)


def test_xl2_eval_cfg_and_short_sampler_match_reference():
    """XL/2 unmasked eval + CFG forward (BASELINE config 5's network evaluation: 256 tokens, 16 heads of 72) and a
    3-step (5-evaluation) CFG sampler run."""
    g = load("xl2_eval")
    sd = O.make_state_dict(XL32, 1)
    lab = t(g["labels"])
    with torch.no_grad():
        plain = O.edm_precond(sd, XL32, t(g["images"]), t(g["sigma"]), lab, training=False)
        np.testing.assert_allclose(plain.numpy(), g["D_plain"], rtol=1e-3, atol=1e-4)
        cfg = O.edm_precond(sd, XL32, t(g["images"]), torch.tensor(1.7, dtype=torch.float64), lab, cfg_scale=1.5,
                            training=False)
        np.testing.assert_allclose(cfg.numpy(), g["D_cfg"], rtol=1e-3, atol=1e-4)
        z, evals = O.edm_sampler(lambda x, s: O.edm_precond(sd, XL32, x, s, lab, cfg_scale=1.5, training=False),
                                 t(g["latents"]), num_steps=int(g["num_steps"]))
    np.testing.assert_allclose(np.array(evals), g["sampler_sigmas"], rtol=1e-12)
    np.testing.assert_allclose(z.numpy(), g["z"], rtol=1e-3, atol=1e-3)


def test_step_front_matches_reference():
    """utils.sample + label dropout + sigma draw + noise injection (train.py:206-209, loss.py:35-39)."""
    g = load("step_front")
    y, yn, sigma, lab = O.step_front(t(g["moments"]), t(g["eps"]), t(g["rnd_normal"]), t(g["noise_unit"]),
                                     t(g["labels"]), t(g["drop_u"]), float(g["drop_prob"]))
    assert np.array_equal(y.numpy(), g["y"]) and np.array_equal(lab.numpy(), g["labels_out"])
    np.testing.assert_allclose(sigma.numpy(), g["sigma"], rtol=1e-6)
    np.testing.assert_allclose(yn.numpy(), g["yn"], rtol=1e-6, atol=1e-6)
    assert 0 < int((lab.sum(1) == 0).sum()) < lab.shape[0]


def test_ablation_sampler_matches_reference():
    """sample.py:73-188: every discretization / schedule / scaling family, Euler and Heun, churn, alpha != 1."""
    g = load("s2_ablation")
    sd = O.make_state_dict(SMALL, 1)
    lab = t(g["labels"])
    for ci in range(int(g["n"])):
        kw = eval(str(g[f"kw{ci}"]))  # noqa: S307 - our own fixture
        noises = [t(n) for n in g[f"noises{ci}"]]
        cs = 1.5 if ci % 2 == 0 else None
        with torch.no_grad():
            z, evals = O.ablation_sampler(
                lambda x, s: O.edm_precond(sd, SMALL, x, s, lab, cfg_scale=cs, training=False), t(g[f"latents{ci}"]),
                randn_like=lambda x: noises.pop(0), num_steps=5, **kw)
        assert not noises
        np.testing.assert_allclose(np.array(evals), g[f"sigmas{ci}"], rtol=1e-9, err_msg=str(kw))
        scale = np.abs(g[f"z{ci}"]).max()
        np.testing.assert_allclose(z.numpy(), g[f"z{ci}"], rtol=2e-3, atol=2e-4 * scale, err_msg=str(kw))


def test_lr_schedule_and_rank_batches():
    assert O.lr_schedule(0, 1e-4, 1024, 0) == 0.0 and O.lr_schedule(1, 1e-4, 1024, 0) == 1e-4   # train.py:223
    assert abs(O.lr_schedule(5, 1e-4, 1024, 10) - 1e-4 * 5 * 1024 / 10000) < 1e-18
    seeds = list(range(100, 170))
    got = [O.rank_batches(seeds, 8, r, 4) for r in range(4)]
    assert sorted(s for rb in got for b in rb for s in b) == seeds
    assert all(len(b) <= 8 for rb in got for b in rb) and len({len(rb) for rb in got}) == 1
    assert got[1][0] == list(range(100 + 6, 100 + 12))   # 70 seeds -> 12 batches of 6/5, rank 1 takes batch 1, 5, 9


def test_vae_decode_matches_reference():
    """SD-VAE decode restatement (oracle/vae_oracle.py) vs the unmodified reference Decoder + post_quant_conv
    (autoencoder.py:306-453) on the stand-in weights; the 8-bit conversion of sample.py:287 bit for bit."""
    from oracle import vae_oracle as VO
    g = load("vae_decode")
    sd = VO.make_vae_state_dict(3)
    assert sum(v.numel() for k, v in sd.items() if k.startswith("decoder.")) == 49_490_179   # SD-VAE decoder size
    with torch.no_grad():
        img = VO.decode(sd, t(g["z"]))
    np.testing.assert_allclose(img.numpy(), g["images"], rtol=1e-3, atol=2e-4)
    u8 = VO.to_uint8(img).numpy()
    assert (u8 != g["u8"]).mean() < 1e-3 and np.abs(u8.astype(int) - g["u8"].astype(int)).max() <= 1
    assert np.array_equal(VO.to_uint8(t(g["images"])).numpy(), g["u8"])
