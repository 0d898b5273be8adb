"""GPU timeline of training steps via torch.profiler (CUPTI): busy vs idle time and per-kernel totals.
python tools/timeline_step.py [B] [R]   ->  prints a summary; writes gpurun_out/timeline_<R>.json (per-kernel table)"""
import copy
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from maskdit_b200.maskdit import Precond_models
from maskdit_b200.train_step import TrainStep

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
R = int(sys.argv[2]) if len(sys.argv) > 2 else 32
NSTEP = 3
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
for _ in range(4):
    ts.step(x, y, 0.5, 0.1)
torch.cuda.synchronize()
# host enqueue time vs device time, unprofiled
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(NSTEP):
    ts.step(x, y, 0.5, 0.1)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"unprofiled: host enqueue {1e3 * (t1 - t0) / NSTEP:.1f} ms/step, device {e0.elapsed_time(e1) / NSTEP:.1f} ms/step")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NSTEP):
        ts.step(x, y, 0.5, 0.1)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in evs)
span = iv[-1][1] - iv[0][0]
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
gaps = []
for s, e, n in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, n))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
tot = {}
for s, e, n in iv:
    k = n.split("(")[0][:70]
    a = tot.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e - s
print(f"span {span / 1e3 / NSTEP:.2f} ms/step, busy {busy / 1e3 / NSTEP:.2f} ms/step, idle {(span - busy) / 1e3 / NSTEP:.2f} ms/step, "
      f"kernels/step {len(iv) / NSTEP:.0f}")
gaps.sort(reverse=True)
print("largest gaps (us):", [(round(g, 1), n[:40]) for g, n in gaps[:8]])
print(f"gap histogram: >20us {sum(g > 20 for g, _ in gaps) / NSTEP:.0f}/step, 5-20us {sum(5 < g <= 20 for g, _ in gaps) / NSTEP:.0f}/step, "
      f"<=5us {sum(g <= 5 for g, _ in gaps) / NSTEP:.0f}/step; sum>20us {sum(g for g, _ in gaps if g > 20) / 1e3 / NSTEP:.2f} ms/step")
rows = sorted(tot.items(), key=lambda kv: -kv[1][1])
for k, (c, t) in rows[:24]:
    print(f"{t / 1e3 / NSTEP:8.3f} ms/step {c / NSTEP:6.0f}x {t / c:8.1f} us  {k}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump({k: {"launches_per_step": c / NSTEP, "ms_per_step": t / 1e3 / NSTEP} for k, (c, t) in rows},
          open(f"gpurun_out/timeline_{R}.json", "w"), indent=1)
