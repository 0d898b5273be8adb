"""GPU timeline of the EDM sampler (torch.profiler / CUPTI): busy vs idle time and per-kernel totals per net eval."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from maskdit_b200.maskdit import Precond_models
from maskdit_b200.sampler import edm_sampler

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
STEPS = 4
EVALS = 2 * STEPS - 1
dev = torch.device("cuda")
torch.manual_seed(0)
net = Precond_models["edm"](img_resolution=32, img_channels=4, num_classes=1000, model_type="DiT-XL/2",
                            use_decoder=True, mae_loss_coef=0.1, pad_cls_token=False)
bench.randomise_zero_init(net)
net = net.to(dev).eval()
lat = torch.randn(B, 4, 32, 32, device=dev)
lab = torch.nn.functional.one_hot(torch.randint(0, 1000, (B,), device=dev), 1000).float()
with torch.no_grad():
    for _ in range(2):
        edm_sampler(net, lat, lab, cfg_scale=1.5, num_steps=STEPS)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    edm_sampler(net, lat, lab, cfg_scale=1.5, num_steps=STEPS)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"unprofiled: host enqueue {1e3 * (t1 - t0) / EVALS:.1f} ms/eval, device {e0.elapsed_time(e1) / EVALS:.1f} ms/eval")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        edm_sampler(net, lat, lab, cfg_scale=1.5, num_steps=STEPS)
        torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
iv = sorted((e.time_range.start, e.time_range.end, e.name) for e in evs)
span = iv[-1][1] - iv[0][0]
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s, e, n in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
tot = {}
for s, e, n in iv:
    a = tot.setdefault(n.split("(")[0][:70], [0, 0.0])
    a[0] += 1
    a[1] += e - s
print(f"span {span / 1e3 / EVALS:.2f} ms/eval, busy {busy / 1e3 / EVALS:.2f}, idle {(span - busy) / 1e3 / EVALS:.2f}, kernels/eval {len(iv) / EVALS:.0f}")
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{t / 1e3 / EVALS:8.3f} ms/eval {c / EVALS:6.1f}x {t / c:8.1f} us  {k}")
