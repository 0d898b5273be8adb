#!/bin/bash
# round 2, GPU job I (1 GPU): VAE decode parity + timing at the real size, remaining targeted tests
timeout 900 python -m pytest tests -q -s -m gpu -k "vae or c_driver or gemm_epilogues or sampler_tail" 2>&1 | grep -v "^$" | tail -n 30
timeout 600 python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
from maskdit_b200.vae import AutoencoderKLDecoder
from maskdit_b200 import ops, _lib
torch.manual_seed(0)
vae = AutoencoderKLDecoder()
g = torch.Generator().manual_seed(1)
with torch.no_grad():
    for k, p in vae.named_parameters():
        if p.ndim == 4: p.copy_(torch.randn(p.shape, generator=g) * (p.shape[1] * p.shape[2] * p.shape[3]) ** -0.5)
        elif k.endswith('weight'): p.fill_(1.0)
vae = vae.cuda().eval()
for B in (8, 32):
    z = torch.randn(B, 4, 32, 32, device='cuda') * 0.18215 * 4
    img = vae.decode(z); torch.cuda.synchronize()
    n0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); img = vae.decode(z); u8 = ops.to_uint8_nhwc(img.contiguous()); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"VAE decode B={B} 32x32x4 -> {tuple(img.shape)}: {ms:.1f} ms = {B / ms * 1e3:.1f} img/s, {_lib.LAUNCHES - n0} launches, finite={bool(torch.isfinite(img).all())}, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
PY
