#!/bin/bash
# First GPU job of round 2: run the opt-in code paths that were written after round 1's GPU budget was spent.
#   MDT_ATTN_SW64=1  decoder (head_dim 32, T=256) attention on SWIZZLE_64B tiles        (attention_sw.cu)
#   MDT_ATTN_SWL=1   blocked attention (T=512/1024) on split TMA tiles                  (attention_sw_long.cu)
#   MDT_TRAIN_GRAPH=1 zero-grad + forward + backward replayed from a CUDA graph         (train_step.py)
# Every kernel has the 2 s mbarrier trap (common.cuh) and each test process its own `timeout`.
echo "== baseline attention parity"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k attention 2>&1 | tail -n 2
echo "== SW64";  MDT_ATTN_SW64=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k attention 2>&1 | tail -n 6
echo "== SWL";   MDT_ATTN_SWL=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k attention 2>&1 | tail -n 6
echo "== train graph"; MDT_TRAIN_GRAPH=1 timeout 600 python -m pytest tests/test_model_gpu.py -q -k "train_step" 2>&1 | tail -n 6
for f in "" "MDT_ATTN_SW64=1" "MDT_ATTN_SWL=1"; do echo "== timing [$f]"; env $f timeout 300 python tools/run_attn_time.py 2>&1 | tail -n 5; done
for f in "" "MDT_ATTN_SW64=1" "MDT_TRAIN_GRAPH=1"; do
  env $f timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('train256 [$f]', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'clk', d['clocks']['sm_mhz'])"
done
for f in "" "MDT_ATTN_SWL=1"; do
  env $f timeout 900 python bench.py --workload train512 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('train512 [$f]', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms', 'clk', d['clocks']['sm_mhz'])"
done
