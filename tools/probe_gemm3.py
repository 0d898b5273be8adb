"""Backward-shape GEMM timings: dgrad (K,MN) and wgrad (MN,MN) with tile width / CG / epilogue variations."""
import os, sys
os.environ.setdefault("MDT_ALLOW_PARTIAL_LIB", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maskdit_b200 import _lib as L
dev = "cuda"
def rnd(*s): return torch.randn(*s, device=dev).to(torch.bfloat16)
def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
cg = os.environ.get("MDT_GEMM_CG", "auto")
print("--- dgrad (K,MN): dX[M,N] = dY[M,K] W[K,N]")
for (M, N, K) in [(32768, 4608, 1152), (32768, 1152, 4608), (32768, 1152, 3456), (32768, 1152, 1152), (65536, 512, 2048)]:
    A, W = rnd(M, K), rnd(K, N)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for bn in (0, 256, 128):
        ms = bench(lambda: L.gemm(A, W, M, N, K, b_mn=True, out=out, block_n=bn))
        print(f"CG={cg} dgrad M{M} N{N} K{K} bn{bn}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TF/s", flush=True)
print("--- wgrad (MN,MN): dW[M,N] = dY[K,M]^T X[K,N]")
for (M, N, K) in [(3456, 1152, 32768), (4608, 1152, 32768), (1152, 4608, 32768), (1152, 1152, 32768), (2048, 512, 65536)]:
    A, B = rnd(K, M), rnd(K, N)
    out = torch.zeros(M, N, device=dev)
    for bn in (0, 256, 128):
        ms = bench(lambda: L.gemm(A, B, M, N, K, a_mn=True, b_mn=True, out=out, block_n=bn, epi=L.EPI_ATOMIC))
        ms2 = bench(lambda: L.gemm(A, B, M, N, K, a_mn=True, b_mn=True, out=out, block_n=bn, epi=L.EPI_STORE))
        print(f"CG={cg} wgrad M{M} N{N} K{K} bn{bn}: streamK+atomic {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TF/s | tile+store {ms2:.3f} ms = {2*M*N*K/ms2/1e9:.0f} TF/s", flush=True)
