#!/usr/bin/env python
"""Sampling entry point with the reference's CLI / YAML surface (generate.py:54-87, configs/test/*.yaml).

Builds `Precond_models[config.model.precond]` from the YAML, loads `ckpt['ema']` (reference checkpoints load
unchanged: same state-dict keys), and runs `edm_sampler` on the B200 engine for the requested seeds with
per-sample generators (utils.StackedRandomGenerator, utils.py:119-133).  The SD-VAE decode + PNG writing of the
reference (sample.py:275-296) is outside the accelerated path: latents are saved as `.npy` per seed.
"""
import argparse
import os

import numpy as np
import torch

from maskdit_b200.config import build_net, load_config, parse_float_none, parse_int_list
from maskdit_b200.sampler import edm_sampler


class StackedRandomGenerator:
    def __init__(self, device, seeds):
        self.generators = [torch.Generator(device).manual_seed(int(s) % (1 << 32)) for s in seeds]

    def randn(self, size, **kw):
        return torch.stack([torch.randn(size[1:], generator=g, **kw) for g in self.generators])

    def randn_like(self, x):
        return self.randn(x.shape, dtype=x.dtype, layout=x.layout, device=x.device)

    def randint(self, *a, size, **kw):
        return torch.stack([torch.randint(*a, size=size[1:], generator=g, **kw) for g in self.generators])


def main():
    ap = argparse.ArgumentParser("Sample from a trained model")
    ap.add_argument("--config", required=True)
    ap.add_argument("--results_dir", default="samples")
    ap.add_argument("--ckpt_path", default=None)
    ap.add_argument("--seeds", type=parse_int_list, default="100-131")
    ap.add_argument("--class_idx", type=int, default=None)
    ap.add_argument("--cfg_scale", type=parse_float_none, default=None)
    ap.add_argument("--num_steps", type=int, default=40)
    ap.add_argument("--S_churn", type=int, default=0)
    ap.add_argument("--max_batch_size", type=int, default=32)
    args, _ = ap.parse_known_args()
    cfg = load_config(args.config)
    device = torch.device("cuda")
    net = build_net(cfg).to(device).eval()
    if args.ckpt_path:
        ck = torch.load(args.ckpt_path, map_location=device)
        net.load_state_dict({k.replace("_orig_mod.", ""): v for k, v in ck["ema"].items()})
    os.makedirs(args.results_dir, exist_ok=True)
    seeds = args.seeds
    for i in range(0, len(seeds), args.max_batch_size):
        bs = seeds[i:i + args.max_batch_size]
        rnd = StackedRandomGenerator(device, bs)
        latents = rnd.randn([len(bs), net.img_channels, net.img_resolution, net.img_resolution], device=device)
        labels = torch.eye(net.num_classes, device=device)[rnd.randint(net.num_classes, size=[len(bs)], device=device)]
        if args.class_idx is not None:
            labels[:, :] = 0
            labels[:, args.class_idx] = 1
        with torch.no_grad():
            z = edm_sampler(net, latents.float(), labels.float(), cfg_scale=args.cfg_scale,
                            randn_like=rnd.randn_like, num_steps=args.num_steps, S_churn=args.S_churn).float()
        for s, zi in zip(bs, z.cpu().numpy()):
            np.save(os.path.join(args.results_dir, f"{s:06d}.npy"), zi)
    print(f"wrote {len(seeds)} latents to {args.results_dir}")


if __name__ == "__main__":
    main()
