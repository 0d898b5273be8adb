"""torchrun --nproc-per-node N tools/dp_equivalence.py — data-parallel equivalence on real GPUs:
the gradient of a global batch split over N ranks and summed by the step's collective equals (x N) the gradient one
GPU computes on the whole batch, for every exchange mode of TrainStep: our own NCCL communicator (fp32, flat),
torch.distributed, the overlapped per-block exchange with an SM budget, and the bf16 buffer (bf16 tolerance).
Prints DP_EQUIV_OK on rank 0."""
import copy
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maskdit_b200.loss import EDMLoss  # noqa: E402
from maskdit_b200.maskdit import Precond_models  # noqa: E402
from maskdit_b200.train_step import TrainStep, shard_batch  # noqa: E402


class Draws(EDMLoss):
    """Fixed random draws, sliced to this call's rows."""

    def __init__(self, rnd, noise, mnoise, lo, hi):
        super().__init__()
        self.t, self.k, self.m = (rnd[lo:hi].contiguous(), noise[lo:hi].contiguous()), 0, mnoise[lo:hi].contiguous()

    def _randn(self, shape, device):
        t = self.t[self.k % 2]
        self.k += 1
        return t

    def _rand(self, shape, device):
        return self.m


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    R, ncls, Bper = 32, 1000, 4
    Bg = Bper * world
    g = torch.Generator().manual_seed(0)
    images = (torch.randn(Bg, 4, R, R, generator=g) * 0.5).to(dev)
    labels = torch.nn.functional.one_hot(torch.randint(0, ncls, (Bg,), generator=g), ncls).float().to(dev)
    rnd, noise = torch.randn(Bg, 1, 1, 1, generator=g).to(dev), torch.randn(Bg, 4, R, R, generator=g).to(dev)
    mnoise = torch.rand(Bg, 256, generator=g).to(dev)
    torch.manual_seed(1)
    base = Precond_models["edm"](img_resolution=R, img_channels=4, num_classes=ncls, model_type="DiT-B/2",
                                 use_decoder=True, mae_loss_coef=0.1, pad_cls_token=False)
    with torch.no_grad():
        gz = torch.Generator().manual_seed(2)
        for p in base.parameters():
            if p.requires_grad and float(p.abs().sum()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=gz) * 0.02)
    lo, hi = shard_batch(Bg, world, rank)

    # single-GPU gradient of the whole global batch (every rank computes it locally: same weights, same draws)
    ref_net = copy.deepcopy(base).to(dev).train()
    ts_ref = TrainStep(ref_net, None, lr=1e-3, loss_fn=Draws(rnd, noise, mnoise, 0, Bg), process_group=None,
                       grad_dtype="fp32")
    ts_ref.world, ts_ref.comm, ts_ref.g16 = 1, None, None      # a 1-GPU step inside the N-rank job
    ts_ref._grad_scale = 1.0
    ts_ref.step(images, labels, 0.5, 0.1)
    g_ref = ts_ref.st.grad.clone()
    w_ref = ts_ref.st.w32.clone()

    def run(**kw):
        net = copy.deepcopy(base).to(dev).train()
        ts = TrainStep(net, None, lr=1e-3, loss_fn=Draws(rnd, noise, mnoise, lo, hi), global_batch=Bg, **kw)
        ts.step(images[lo:hi].contiguous(), labels[lo:hi].contiguous(), 0.5, 0.1)
        torch.cuda.synchronize()
        gsum = (ts.g16.float() if ts.g16 is not None else ts.st.grad).clone()
        return gsum / world, ts.st.w32.clone(), ts

    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm()).item()

    results = {}
    for name, kw, tol in (("mdt fp32 flat", dict(collective="mdt", grad_dtype="fp32", overlap=False), 5e-5),
                          ("torch fp32 flat", dict(collective="torch", grad_dtype="fp32", overlap=False), 5e-5),
                          ("mdt fp32 overlapped", dict(collective="mdt", grad_dtype="fp32", overlap=True), 5e-5),
                          ("mdt bf16 flat", dict(collective="mdt", grad_dtype="bf16", overlap=False), 6e-3),
                          ("mdt bf16 overlapped", dict(collective="mdt", grad_dtype="bf16", overlap=True), 6e-3)):
        gm, w, ts = run(**kw)
        r = rel(gm, g_ref)
        dw = (w - w_ref).abs().max().item()
        results[name] = (r, dw)
        if rank == 0:
            print(f"{name:22s} gradient rel-L2 vs 1-GPU whole batch: {r:.3e}   max |dw| after the step: {dw:.3e}   "
                  f"[{ts.describe_collective()}]", flush=True)
        assert r <= tol, (name, r)
        assert dw <= 2.5e-3, (name, dw)      # one Adam step of lr 1e-3: sign-like, order noise flips only ~0 gradients
        ts.close()
    dist.barrier()
    if rank == 0:
        print("DP_EQUIV_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
