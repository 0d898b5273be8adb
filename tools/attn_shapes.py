"""Launch the attention kernels once at the config-2 (and sampler) shapes; used under ncu."""
import sys, torch
sys.path.insert(0, '.')
from maskdit_b200 import ops
shapes = [(256, 128, 16, 72), (256, 256, 16, 32)]
if len(sys.argv) > 1 and sys.argv[1] == "sampler":
    shapes = [(128, 256, 16, 72)]
for (B, T, H, dh) in shapes:
    qkv = (torch.randn(B * T, 3 * H * dh, device='cuda')).to(torch.bfloat16)
    for _ in range(2):
        out, lse = ops.attention_fwd(qkv, B, T, H, dh)
        dout = torch.randn_like(out)
        ops.attention_bwd(qkv, out, dout, lse, B, T, H, dh)
        torch.cuda.synchronize()
