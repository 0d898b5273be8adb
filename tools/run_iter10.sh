#!/bin/bash
# general-T tcgen05 attention: parity, timing at the 512-px (C4) shapes, then the train512 bench with / without it
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" 2>&1 | tail -n 5
for tc in 1 0; do
MDT_ATTN_LONG=$tc python - <<'PY'
import os, sys, torch
sys.path.insert(0, '.')
from maskdit_b200 import ops
def bench(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (B,T,H,dh) in [(128,512,16,72),(128,1024,16,32),(128,256,16,72)]:
    qkv = (torch.randn(B*T, 3*H*dh, device='cuda')).to(torch.bfloat16)
    out, lse = ops.attention_fwd(qkv, B, T, H, dh)
    dout = torch.randn_like(out)
    tf = bench(lambda: ops.attention_fwd(qkv, B, T, H, dh))
    tb = bench(lambda: ops.attention_bwd(qkv, out, dout, lse, B, T, H, dh))
    fl = 4 * B * H * T * T * dh
    print(f"long={os.environ['MDT_ATTN_LONG']} attn B{B} T{T} H{H} dh{dh}: fwd {tf:.0f} us ({fl/tf/1e6:.0f} TF/s)  bwd {tb:.0f} us ({2.5*fl/tb/1e6:.0f} TF/s)")
PY
done
for tc in 1 0; do
MDT_ATTN_LONG=$tc timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --workload train512 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train512 long=$tc', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'clk', d['clocks']['sm_mhz'])"
done
