#!/bin/bash
# A/B (new vs prev library) of the attention forward change: parity, kernel timing, sampler + train bench
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention or gemm" 2>&1 | tail -n 3
for v in new prev; do
  if [ $v = new ]; then unset MDT_LIB_PATH; else export MDT_LIB_PATH=$PWD/maskdit_b200/libmaskdit_b200_$v.so; fi
  echo "=== $v"
  python - <<'PY'
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
for (B,T,H,dh) in [(256,128,16,72),(256,256,16,32),(128,256,16,72)]:
    qkv = (torch.randn(B*T, 3*H*dh, device='cuda')).to(torch.bfloat16)
    out, lse = ops.attention_fwd(qkv, B, T, H, dh)
    dout = torch.randn_like(out)
    tf = bench(lambda: ops.attention_fwd(qkv, B, T, H, dh))
    tb = bench(lambda: ops.attention_bwd(qkv, out, dout, lse, B, T, H, dh))
    print(f"attn B{B} T{T} H{H} dh{dh}: fwd {tf:.0f} us  bwd {tb:.0f} us  (fwd HBM floor {(qkv.numel()+out.numel())*2/6.5e6:.0f} us)")
PY
  timeout 600 python bench.py --workload sampler --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v sampler', round(d['value'],2), d['unit'], 'clk', d['clocks']['sm_mhz'])"
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v train', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'gemm', round(d['roofline']['achieved']), 'clk', d['clocks']['sm_mhz'])"
done
