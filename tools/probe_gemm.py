"""GPU probe: tcgen05 GEMM vs torch fp32 matmul on bf16-rounded operands.  Usage: probe_gemm.py <group>"""
import json
import os
import sys

os.environ.setdefault("MDT_ALLOW_PARTIAL_LIB", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from maskdit_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"
res = []


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


def report(name, got, ref, tol):
    got = got.float()
    err = (got - ref).abs().max().item()
    rel = err / (ref.abs().max().item() + 1e-12)
    bad = int(((got - ref).abs() > tol * (1 + ref.abs())).sum().item())
    ok = bool(rel < tol and torch.isfinite(got).all().item())
    res.append(dict(name=name, max_abs=err, rel=rel, n_bad=bad, ok=ok))
    print(f"{'OK ' if ok else 'BAD'} {name}: max_abs={err:.4g} rel={rel:.3g} bad={bad}/{got.numel()}", flush=True)
    if not ok:
        d = (got - ref).abs()
        idx = torch.nonzero(d > tol * (1 + ref.abs()))[:8]
        for i in idx:
            i = tuple(i.tolist())
            print("   ", i, got[i].item(), ref[i].item())


def kk(M, N, K, bn=0, out_dtype=torch.float32):
    A, B = rnd(M, K), rnd(N, K)
    out = torch.full((M, N), float("nan"), device=dev, dtype=out_dtype)
    L.gemm(A, B, M, N, K, out=out, block_n=bn)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    report(f"KK M{M} N{N} K{K} bn{bn} {out_dtype}", out, ref, 2e-2 if out_dtype == torch.bfloat16 else 1e-3)


def kmn(M, N, K, bn=0):  # dgrad: dX[M,N] = dY[M,K] @ W[K,N]   (B stored [K, N], N contiguous)
    A, W = rnd(M, K), rnd(K, N)
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32)
    L.gemm(A, W, M, N, K, b_mn=True, out=out, block_n=bn)
    torch.cuda.synchronize()
    report(f"K-MN M{M} N{N} K{K} bn{bn}", out, A.float() @ W.float(), 1e-3)


def mnmn(M, N, K, bn=0, atomic=True):  # wgrad: dW[M,N] = dY[K,M]^T @ X[K,N]
    A, B = rnd(K, M), rnd(K, N)
    out = torch.zeros((M, N), device=dev, dtype=torch.float32)
    L.gemm(A, B, M, N, K, a_mn=True, b_mn=True, out=out, block_n=bn, epi=L.EPI_ATOMIC if atomic else L.EPI_STORE)
    torch.cuda.synchronize()
    report(f"MN-MN M{M} N{N} K{K} bn{bn} atomic={atomic}", out, A.float().t() @ B.float(), 1e-3)


def epilogues():
    M, N, K, T = 512, 1152, 1152, 128
    A, B = rnd(M, K), rnd(N, K, scale=0.05)
    bias = torch.randn(N, device=dev)
    acc = A.float() @ B.float().t() + bias
    # bias store bf16
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    L.gemm(A, B, M, N, K, out=out, bias=bias)
    report("epi bias bf16", out, acc, 2e-2)
    # silu
    out = torch.empty(M, N, device=dev, dtype=torch.float32)
    L.gemm(A, B, M, N, K, out=out, bias=bias, act=L.ACT_SILU)
    report("epi bias+silu f32", out, torch.nn.functional.silu(acc), 2e-3)
    # resid add in EPI_STORE
    R = torch.randn(M, N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float32)
    L.gemm(A, B, M, N, K, out=out, bias=bias, resid=R, ld_resid=N)
    report("epi bias+resid f32", out, acc + R, 1e-3)
    # gelu
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    L.gemm(A, B, M, N, K, out=out, bias=bias, epi=L.EPI_GELU, aux=aux, ld_aux=N)
    pre = acc.to(torch.bfloat16).float()
    report("epi gelu pre", aux, acc, 2e-2)
    report("epi gelu out", out, torch.nn.functional.gelu(pre, approximate="tanh"), 2e-2)
    # gate resid
    gate = torch.randn(M // T, N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float32)
    aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    L.gemm(A, B, M, N, K, out=out, bias=bias, epi=L.EPI_GATE_RESID, aux=aux, ld_aux=N, resid=R, ld_resid=N, gate=gate,
           ld_gate=N, rows_per_group=T)
    report("epi gate_resid y", aux, acc, 2e-2)
    report("epi gate_resid out", out, R + gate.repeat_interleave(T, 0) * acc, 1e-3)
    # dgelu
    h = rnd(M, N)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    L.gemm(A, B, M, N, K, out=out, epi=L.EPI_DGELU, aux=h, ld_aux=N)
    hf = h.float().requires_grad_(True)
    torch.nn.functional.gelu(hf, approximate="tanh").sum().backward()
    report("epi dgelu", out, (A.float() @ B.float().t()) * hf.grad, 2e-2)


def timing():
    import time
    for (M, N, K) in [(32768, 4608, 1152), (32768, 1152, 4608), (32768, 3456, 1152), (32768, 1152, 1152),
                      (65536, 2048, 512), (65536, 512, 2048)]:
        A, B = rnd(M, K), rnd(N, K)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            L.gemm(A, B, M, N, K, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            L.gemm(A, B, M, N, K, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        tf = 2 * M * N * K / ms / 1e9
        t0 = time.time()
        for _ in range(3):
            ref = A @ B.t()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            ref = A @ B.t()
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / n
        print(f"TIME KK M{M} N{N} K{K}: {ms:.3f} ms = {tf:.0f} TFLOP/s   (cuBLAS {ms2:.3f} ms = {2*M*N*K/ms2/1e9:.0f})", flush=True)
        res.append(dict(name=f"time M{M} N{N} K{K}", ms=ms, tflops=tf, cublas_ms=ms2, ok=True))
    # wgrad stream-K timing
    for (M, N, K) in [(3456, 1152, 32768), (4608, 1152, 32768), (512, 512, 65536)]:
        A, B = rnd(K, M), rnd(K, N)
        out = torch.zeros(M, N, device=dev)
        for _ in range(3):
            L.gemm(A, B, M, N, K, a_mn=True, b_mn=True, out=out, epi=L.EPI_ATOMIC)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            L.gemm(A, B, M, N, K, a_mn=True, b_mn=True, out=out, epi=L.EPI_ATOMIC)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"TIME wgrad M{M} N{N} K{K}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
        res.append(dict(name=f"time wgrad M{M} N{N} K{K}", ms=ms, tflops=2 * M * N * K / ms / 1e9, ok=True))


group = sys.argv[1]
if group == "kk":
    kk(128, 128, 64, bn=128)
    kk(128, 256, 64, bn=256)
    kk(128, 192, 128, bn=192)
    kk(256, 256, 256)
    kk(512, 1152, 1152)
    kk(384, 3456, 1152)
    kk(300, 200, 1000)          # ragged everything (K tail via TMA zero fill)
    kk(2, 1152, 256)            # tiny M
    kk(1024, 512, 2048, out_dtype=torch.bfloat16)
    kk(148 * 128 * 2 + 128, 256, 128)  # multi-wave persistent loop
elif group == "kmn":
    kmn(128, 128, 64, bn=128)
    kmn(256, 256, 128, bn=256)
    kmn(512, 1152, 3456)
    kmn(512, 1152, 4608)
    kmn(256, 512, 16)
    kmn(300, 1000, 1152)
elif group == "mnmn":
    mnmn(128, 128, 64, bn=128, atomic=False)
    mnmn(256, 256, 128, bn=256, atomic=False)
    mnmn(128, 128, 64, bn=128)
    mnmn(3456, 1152, 2048)
    mnmn(512, 512, 8192)
    mnmn(1152, 1000, 256)
    mnmn(16, 512, 4096)
elif group == "epi":
    epilogues()
elif group == "time":
    timing()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/probe_gemm_{group}.json", "w"), indent=1)
