"""One profiled training step (between cudaProfilerStart/Stop) for ncu:  ncu --profile-from-start off ... python tools/profile_step.py"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from maskdit_b200.maskdit import Precond_models
from maskdit_b200.train_step import TrainStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda")
torch.manual_seed(0)
net = Precond_models["edm"](img_resolution=R, img_channels=4, num_classes=1000, model_type="DiT-XL/2",
                            use_decoder=True, mae_loss_coef=0.1, pad_cls_token=False)
bench.randomise_zero_init(net)
net = net.to(dev).train()
ema = copy.deepcopy(net).eval()
ts = TrainStep(net, ema)
x, y = bench.make_batches(1, B, R, 1000)[0]
x, y = x.to(dev), y.to(dev)
for _ in range(3):
    ts.step(x, y, 0.5, 0.1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ts.step(x, y, 0.5, 0.1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step")
