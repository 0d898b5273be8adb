import os, sys, torch
sys.path.insert(0, '.')
from maskdit_b200 import ops
def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (B,T,H,dh) in [(256,128,16,72),(256,256,16,32),(128,256,16,72),(128,512,16,72),(128,1024,16,32)]:
    qkv = (torch.randn(B*T, 3*H*dh, device='cuda')).to(torch.bfloat16)
    out, lse = ops.attention_fwd(qkv, B, T, H, dh)
    dout = torch.randn_like(out)
    tf = bench(lambda: ops.attention_fwd(qkv, B, T, H, dh))
    tb = bench(lambda: ops.attention_bwd(qkv, out, dout, lse, B, T, H, dh))
    print(f"attn B{B} T{T} H{H} dh{dh}: fwd {tf:.0f} us  bwd {tb:.0f} us  (HBM floor fwd {(qkv.numel()+out.numel())*2/6.5e6:.0f} bwd {(2*qkv.numel()+2*out.numel())*2/6.5e6:.0f} us)")
