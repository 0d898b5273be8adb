"""Phase cycle counts of the persistent attention backward (library built with -DMDT_ATTN_PROF; see attention_tc.cu):
nvcc ... -DMDT_ATTN_PROF -c attention_tc.cu, link as libmaskdit_b200_prof.so, run with MDT_LIB_PATH=<that library>."""
import os, sys, torch
sys.path.insert(0, '.')
from maskdit_b200 import ops
names = ["prefetch issue+load wait", "delta(O)+S/dP issue", "S/dP MMA wait", "delta pass+exch", "P/dS pass+bar",
         "grad MMA issue", "grad MMA wait", "dQ readout+bar", "dK/dV readout+bar"]
for (B, T, H, dh) in [(256, 128, 16, 72), (256, 256, 16, 32)]:
    qkv = (torch.randn(B * T, 3 * H * dh, device='cuda')).to(torch.bfloat16)
    out, lse = ops.attention_fwd(qkv, B, T, H, dh)
    dout = torch.randn_like(out)
    for _ in range(2):
        ops.attention_bwd(qkv, out, dout, lse, B, T, H, dh)
    torch.cuda.synchronize()
    items_per_cta = (B * H + 147) // 148
    v = lse[1].flatten()[:32].tolist()
    print(f"B{B} T{T} dh{dh}: ~{items_per_cta} items/CTA; cycles per item, thread 0 | thread 200")
    for i, n in enumerate(names):
        print(f"   {n:26s} {v[i] / items_per_cta:8.0f} | {v[16 + i] / items_per_cta:8.0f}")
    print(f"   total                      {sum(v[:9]) / items_per_cta:8.0f}")
