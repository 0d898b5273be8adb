#!/bin/bash
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from maskdit_b200 import ops
B,T,H,dh = 256,128,16,72
qkv = (torch.randn(B*T, 3*H*dh, device='cuda')).to(torch.bfloat16)
out, lse = ops.attention_fwd(qkv, B, T, H, dh)
dout = torch.randn_like(out)
for _ in range(2):
    ops.attention_bwd(qkv, out, dout, lse, B, T, H, dh)
    torch.cuda.synchronize()
PY
