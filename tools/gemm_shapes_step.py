"""Per-shape GEMM timing inside one real training step (CUDA events around every mdt_gemm_bf16 launch)."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from maskdit_b200 import _lib
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
for _ in range(4):
    ts.step(x, y, 0.5, 0.1)
torch.cuda.synchronize()
_lib.GEMM_PROFILE = []
for _ in range(2):
    ts.step(x, y, 0.5, 0.1)
torch.cuda.synchronize()
prof, _lib.GEMM_PROFILE = _lib.GEMM_PROFILE, None
EPI = {0: "store", 1: "gelu", 2: "gate_resid", 3: "dgelu", 4: "atomic"}
agg = {}
for f, a, b, k in prof:
    e = agg.setdefault(k, [0, 0.0, 0.0])
    e[0] += 1; e[1] += a.elapsed_time(b); e[2] += f
tot = sum(v[1] for v in agg.values())
print(f"GEMM total {tot / 2:.2f} ms/step, {sum(v[2] for v in agg.values()) / tot / 1e9:.0f} TF/s (event-bracketed: includes launch gaps)")
print(f"{'M':>7} {'N':>7} {'K':>7} amn bmn {'epi':>10} f32 {'n/step':>6} {'avg us':>8} {'TF/s':>6} {'ms/step':>8}")
for k, (c, ms, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M, N, K, amn, bmn, epi, f32 = k
    print(f"{M:7d} {N:7d} {K:7d} {amn:3d} {bmn:3d} {EPI.get(epi, epi):>10} {f32:3d} {c / 2:6.0f} {1e3 * ms / c:8.1f} {f / ms / 1e9:6.0f} {ms / 2:8.3f}")
