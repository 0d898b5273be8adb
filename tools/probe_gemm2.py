"""Discriminating experiments for the GEMM: mainloop-only (epi=99: no global stores) vs epilogue-only (K=64)."""
import os, sys
os.environ.setdefault("MDT_ALLOW_PARTIAL_LIB", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maskdit_b200 import _lib as L
dev = "cuda"
def rnd(*s): return torch.randn(*s, device=dev).to(torch.bfloat16)
def t(M, N, K, epi=0, out_dtype=torch.bfloat16, bn=0, n=10, **kw):
    A, B = rnd(M, K), rnd(N, K)
    out = torch.empty(M, N, device=dev, dtype=out_dtype)
    for _ in range(3): L.gemm(A, B, M, N, K, out=out, epi=epi, block_n=bn, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): L.gemm(A, B, M, N, K, out=out, epi=epi, block_n=bn, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return ms, 2 * M * N * K / ms / 1e9
cg = os.environ.get("MDT_GEMM_CG", "auto")
for (M, N, K) in [(32768, 4608, 1152), (32768, 1152, 4608), (65536, 2048, 512)]:
    for bn in (256,):
        ms, tf = t(M, N, K, epi=0, bn=bn)
        ms2, tf2 = t(M, N, K, epi=99, bn=bn)
        ms3, _ = t(M, N, 64, epi=0, bn=bn)
        tiles = (M // 128) * ((N + bn - 1) // bn)
        print(f"CG={cg} M{M} N{N} K{K} bn{bn}: full {ms:.3f} ms ({tf:.0f} TF/s) | no-store {ms2:.3f} ms ({tf2:.0f} TF/s) | "
              f"K=64 (epilogue-bound) {ms3:.3f} ms = {ms3*1e-3/ (tiles/148) * 1.85e9:.0f} cyc/tile", flush=True)
