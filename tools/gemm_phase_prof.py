"""Epilogue phase cycle counts of the GEMM kernel (library built with -DMDT_GEMM_PROF, see gemm_tcgen05.cu):
cycles per 128x256 tile spent by ONE epilogue warp of CTA 0 in each phase, for the training-step GEMM shapes."""
import ctypes, os, sys, torch
sys.path.insert(0, '.')
from maskdit_b200 import _lib as L
from maskdit_b200.ops import EPI_STORE, EPI_GELU, EPI_GATE_RESID, EPI_DGELU, EPI_ATOMIC
lib = L.lib()
names = ["wait accumulator", "operand issue + tcgen05.ld", "staging stores", "read/math/global stores", "release"]
dev = "cuda"
bf, f32 = torch.bfloat16, torch.float32
def run(tag, M, N, K, epi, **kw):
    A = torch.randn(M, K, device=dev).to(bf); B = torch.randn(N, K, device=dev).to(bf)
    out = torch.empty(M, N, device=dev, dtype=kw.pop("odt", bf))
    for _ in range(3):
        L.gemm(A, B, M, N, K, out=out, epi=epi, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): L.gemm(A, B, M, N, K, out=out, epi=epi, **kw)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    buf = (ctypes.c_float * 12)()
    assert lib.mdt_debug_gemm_prof(buf) == 0
    v = list(buf); nt = max(v[5], 1)
    print(f"{tag}: {us:.0f} us, {2*M*N*K/us/1e6:.0f} TF/s, {nt:.0f} tiles/CTA; cycles per tile: " +
          ", ".join(f"{n} {v[i]/nt:.0f}" for i, n in enumerate(names)) + f"; total {sum(v[:5])/nt:.0f}" +
          f" | MMA issuer per tile: wait accumulator release {v[6]/nt:.0f}, wait full stage {v[7]/nt:.0f}, issue {v[8]/nt:.0f}")
M = 32768
bias = lambda n: torch.randn(n, device=dev)
run("qkv store   ", M, 3456, 1152, EPI_STORE, bias=bias(3456))
run("fc1 gelu    ", M, 4608, 1152, EPI_GELU, bias=bias(4608), aux=torch.empty(M, 4608, device=dev, dtype=bf), ld_aux=4608)
run("fc2 dgrad dgelu", M, 4608, 1152, EPI_DGELU, aux=torch.randn(M, 4608, device=dev).to(bf), ld_aux=4608)
gate = torch.randn(256, 1152, device=dev)
run("fc2 gate_res", M, 1152, 4608, EPI_GATE_RESID, odt=f32, bias=bias(1152), aux=torch.empty(M, 1152, device=dev, dtype=bf), ld_aux=1152,
    resid=torch.randn(M, 1152, device=dev), ld_resid=1152, gate=gate, ld_gate=1152, rows_per_group=128)
run("proj gate_res", M, 1152, 1152, EPI_GATE_RESID, odt=f32, bias=bias(1152), aux=torch.empty(M, 1152, device=dev, dtype=bf), ld_aux=1152,
    resid=torch.randn(M, 1152, device=dev), ld_resid=1152, gate=gate, ld_gate=1152, rows_per_group=128)
run("fc1 dgrad store", M, 1152, 4608, EPI_STORE)
run("dec gelu K512", 65536, 2048, 512, EPI_GELU, bias=bias(2048), aux=torch.empty(65536, 2048, device=dev, dtype=bf), ld_aux=2048)
