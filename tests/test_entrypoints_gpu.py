"""GPU: the train.py / generate.py twins run end to end on the reference's YAML schema (small DiT-S/2 config)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

YAML = """
data: {dataset: imagenet256-latent, category: lmdb, resolution: 16, num_channels: 4, root: none, feat_path: None}
model:
  precond: edm
  model_type: DiT-S/2
  in_size: 16
  in_channels: 4
  num_classes: 1000
  use_decoder: True
  ext_feature_dim: 0
  pad_cls_token: False
  mask_ratio: 0.5
  mask_ratio_fn: constant
  mask_ratio_min: 0
  mae_loss_coef: 0.1
  class_dropout_prob: 0.1
train: {tf32: False, amp: True, batchsize: 8, grad_accum: 1, epochs: 1, lr: 0.0001, lr_rampup_kimg: 0, xflip: False,
        max_num_steps: 4}
log: {log_every: 2, ckpt_every: 4, tag: t}
"""


def run(cmd, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, *cmd], cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_train_then_generate(tmp_path):
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(YAML)
    out = run([os.path.join(ROOT, "train.py"), "--config", str(cfg), "--synthetic", "--max_steps", "4",
               "--results_dir", str(tmp_path / "res")], str(tmp_path))
    assert "Train Loss" in out
    ck = tmp_path / "res" / "checkpoints" / "0000004.pt"
    assert ck.exists()
    sd = torch.load(ck, map_location="cpu", weights_only=False)
    assert set(sd) >= {"model", "ema"} and "model.blocks.0.attn.qkv.weight" in sd["ema"]
    out = run([os.path.join(ROOT, "generate.py"), "--config", str(cfg), "--ckpt_path", str(ck), "--seeds", "0-3",
               "--num_steps", "6", "--cfg_scale", "1.5", "--results_dir", str(tmp_path / "samples")], str(tmp_path))
    z = np.load(tmp_path / "samples" / "000002.npy")
    assert z.shape == (4, 16, 16) and np.isfinite(z).all()


def test_train_from_lmdb_dataset_with_grad_accum_then_resume_and_ablation_generate(tmp_path):
    """train.py WITHOUT --synthetic: the reference's LMDB latent layout (z-{i} / y-{i} / length) feeds the fused step
    front; grad_accum = 2; the checkpoint stores `args` as a Namespace like the reference's and resumes; generate.py
    runs the ablation sampler rank-strided (single rank here)."""
    sys.path.insert(0, ROOT)
    from maskdit_b200.data import write_latent_lmdb
    rng = np.random.default_rng(0)
    write_latent_lmdb(str(tmp_path / "data"), rng.standard_normal((64, 8, 16, 16)).astype(np.float32),
                      rng.integers(0, 1000, 64))
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(YAML.replace("root: none", f"root: {tmp_path / 'data'}").replace("batchsize: 8, grad_accum: 1",
                                                                                  "batchsize: 4, grad_accum: 2"))
    out = run([os.path.join(ROOT, "train.py"), "--config", str(cfg), "--max_steps", "4", "--results_dir",
               str(tmp_path / "res")], str(tmp_path))
    assert "Dataset contains 64 images" in out and "Train Loss" in out
    ck = tmp_path / "res" / "checkpoints" / "0000004.pt"
    sd = torch.load(ck, map_location="cpu", weights_only=False)
    import argparse
    assert isinstance(sd["args"], argparse.Namespace) and min(sd["opt"]["state"]) == 2
    out = run([os.path.join(ROOT, "train.py"), "--config", str(cfg), "--max_steps", "2", "--results_dir",
               str(tmp_path / "res")], str(tmp_path))            # resumes from 0000004.pt
    assert "(step=0000006)" in out
    run([os.path.join(ROOT, "generate.py"), "--config", str(cfg), "--ckpt_path", str(ck), "--seeds", "0-2",
         "--num_steps", "4", "--solver", "euler", "--discretization", "vp", "--schedule", "vp", "--scaling", "vp",
         "--png_preview", "--results_dir", str(tmp_path / "samples")], str(tmp_path))
    z = np.load(tmp_path / "samples" / "000001.npy")
    assert z.shape == (4, 16, 16) and np.isfinite(z).all()
    assert open(tmp_path / "samples" / "000001.png", "rb").read(8) == b"\x89PNG\r\n\x1a\n"


def test_train_from_webdataset_shards(tmp_path):
    """train.py --wds: the reference's WebDataset layout (lmdb2wds.py:26) feeds the same fused step front."""
    sys.path.insert(0, ROOT)
    from maskdit_b200.data import write_wds_shard
    rng = np.random.default_rng(1)
    os.makedirs(tmp_path / "shards")
    for s in range(2):
        write_wds_shard(str(tmp_path / "shards" / f"latent-{s:04d}.tar"),
                        rng.standard_normal((24, 8, 16, 16)).astype(np.float32), rng.integers(0, 1000, 24), start=24 * s)
    cfg = tmp_path / "cfg.yaml"
    cfg.write_text(YAML.replace("root: none", f"root: {tmp_path / 'shards'}"))
    out = run([os.path.join(ROOT, "train.py"), "--config", str(cfg), "--wds", "--max_steps", "4", "--results_dir",
               str(tmp_path / "res")], str(tmp_path))
    assert "2 WebDataset shards" in out and "(step=0000004)" in out
