"""torchrun --nproc-per-node N tools/allreduce_probe.py — time of the step's ONE collective (sum-all-reduce of the flat
gradient buffer: 730 115 216 elements, fp32 2.92 GB / bf16 1.46 GB) through our own communicator (mdt_allreduce_grads)
and through torch.distributed, under whatever NCCL_* environment the caller set.  Prints one line on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskdit_b200.train_step import GradComm, ar_chunk_bounds  # noqa: E402


def timed(fn, n=8, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 730_115_216
    comm = GradComm(None)
    g32 = torch.zeros(n, device="cuda")
    g16 = torch.zeros(n, device="cuda", dtype=torch.bfloat16)
    res = {}
    res["mdt fp32"] = timed(lambda: comm.all_reduce(g32))
    res["mdt bf16"] = timed(lambda: comm.all_reduce(g16))
    res["torch fp32"] = timed(lambda: dist.all_reduce(g32))
    b4 = ar_chunk_bounds(n, 4)
    res["mdt fp32 4 chunks"] = timed(lambda: [comm.all_reduce(g32[lo:hi]) for lo, hi in b4])
    if rank == 0:
        tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("NCCL_")) or "(default NCCL env)"
        print(f"N={world} [{tag}] " + "  ".join(
            f"{k}: {v:.2f} ms ({(2.92 if 'fp32' in k else 1.46) / v * 1e3:.0f} GB/s algbw)" for k, v in res.items()), flush=True)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
