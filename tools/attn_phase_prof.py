"""Phase cycle counts of the persistent attention forward (library built with -DMDT_ATTN_PROF; see attention_tc.cu)."""
import os, sys, torch
sys.path.insert(0, '.')
from maskdit_b200 import ops
names = ["load wait+bar", "S issue+prefetch issue", "S MMA wait", "max pass+exch", "exp pass+bar", "PV issue+lse", "PV wait", "readout"]
for (B, T, H, dh) in [(256, 128, 16, 72), (256, 256, 16, 32), (128, 256, 16, 72)]:
    qkv = (torch.randn(B * T, 3 * H * dh, device='cuda')).to(torch.bfloat16)
    for _ in range(2):
        out, lse = ops.attention_fwd(qkv, B, T, H, dh)
    torch.cuda.synchronize()
    items_per_cta = (B * H * (T // 128) + 147) // 148
    v = lse[1].flatten()[:16].tolist()
    print(f"B{B} T{T} dh{dh}: ~{items_per_cta} items/CTA; cycles per item, thread 0 | thread 200")
    for i, n in enumerate(names):
        print(f"   {n:26s} {v[i] / items_per_cta:8.0f} | {v[8 + i] / items_per_cta:8.0f}")
    print(f"   total                      {sum(v[:8]) / items_per_cta:8.0f}")
