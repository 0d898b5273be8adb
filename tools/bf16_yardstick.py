"""Dev-container only (needs /root/reference): rel-L2 of the UNMODIFIED reference run under PyTorch's own bf16 autocast
(CPU) against its fp32 output, on the inputs of the committed goldens.  This is the yardstick the CUDA path's
tolerances are stated against: no bf16 implementation can be closer to fp32 than bf16 rounding allows."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as mg  # noqa: E402  (installs the timm stand-in, imports the reference)
from oracle import maskdit_oracle as O  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def main():
    for name, cfg in (("s2_eval", O.Cfg(model_type="DiT-S/2", img_resolution=8, num_classes=10)),
                      ("xl2_eval", O.Cfg(model_type="DiT-XL/2", img_resolution=32, num_classes=1000))):
        g = {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(os.path.join(ROOT, "tests/golden", name + ".npz")).items()
             if v.dtype.kind in "fiub"}
        net = mg.build_ref(cfg).eval()
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            plain = net(g["images"], g["sigma"], g["labels"])["x"]
            c = net(g["images"], torch.tensor(1.7, dtype=torch.float64), g["labels"], 1.5)["x"]
        print(f"{name}: PyTorch bf16 autocast of the reference vs its fp32 output: plain rel-L2 {rel(plain, g['D_plain']):.3e}, "
              f"CFG rel-L2 {rel(c, g['D_cfg']):.3e}", flush=True)


if __name__ == "__main__":
    main()
