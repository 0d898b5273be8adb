"""YAML config + CLI helpers for the train.py / generate.py twins.

The reference's config files (configs/{train,finetune,test}/*.yaml, schema in SURVEY.md §5) are accepted unchanged;
OmegaConf is replaced by PyYAML + attribute access.  Known quirks handled: `model.mask_ratio_fn: cos4` in
configs/finetune/imagenet256-latent-cos.yaml is an alias the reference's helper does not know (helper.py:14);
test YAMLs lack `model.self_cond` / `model.mask_ratio_fn`.
"""
from __future__ import annotations

import math
import re

import yaml


class Node(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def get_path(self, path, default=None):
        cur = self
        for part in path.split("."):
            if not isinstance(cur, dict) or part not in cur:
                return default
            cur = cur[part]
        return cur


def _wrap(x):
    if isinstance(x, dict):
        return Node({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    if isinstance(x, str) and x == "None":
        return None
    return x


def load_config(path_or_text: str) -> Node:
    text = open(path_or_text).read() if "\n" not in path_or_text else path_or_text
    return _wrap(yaml.safe_load(text))


def mask_ratio_schedule(name="constant", ratio_scale=0.5, ratio_min=0.0):
    """get_mask_ratio_fn (train_utils/helper.py:9-27): progress in [0,1] -> mask ratio."""
    m = re.fullmatch(r"cos(?:ine)?(\d)", name or "constant")
    if m:
        k = int(m.group(1))
        return lambda x: (ratio_scale - ratio_min) * math.cos(math.pi * x / 2) ** k + ratio_min
    if name == "exp":
        return lambda x: (ratio_scale - ratio_min) * math.exp(-x * 7) + ratio_min
    if name == "linear":
        return lambda x: (ratio_scale - ratio_min) * x + ratio_min
    if name in ("constant", None):
        return lambda x: ratio_scale
    raise ValueError(f"Unknown mask ratio function: {name}")


def parse_int_list(s):
    """'1,2,5-10' -> [1,2,5,...,10]  (utils.py:140-151)."""
    if isinstance(s, list):
        return s
    out = []
    for part in s.split(","):
        m = re.fullmatch(r"(\d+)-(\d+)", part)
        out.extend(range(int(m.group(1)), int(m.group(2)) + 1) if m else [int(part)])
    return out


def parse_float_none(s):
    return None if s is None or str(s).lower() == "none" else float(s)


def build_net(cfg: Node, **extra):
    """Precond_models[config.model.precond](...) exactly as train.py:123-131 / generate.py:31-40 call it."""
    from .maskdit import Precond_models
    m = cfg.model
    return Precond_models[m.precond](img_resolution=m.in_size, img_channels=m.in_channels, num_classes=m.num_classes,
                                     model_type=m.model_type, use_decoder=m.use_decoder,
                                     mae_loss_coef=m.mae_loss_coef, pad_cls_token=m.pad_cls_token, **extra)
