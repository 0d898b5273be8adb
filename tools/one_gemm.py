import os, sys
os.environ.setdefault("MDT_ALLOW_PARTIAL_LIB", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from maskdit_b200 import _lib as L
M, N, K = 32768, 4608, int(sys.argv[1]) if len(sys.argv) > 1 else 1152
A = torch.randn(M, K, device="cuda").to(torch.bfloat16); B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3): L.gemm(A, B, M, N, K, out=out)
torch.cuda.synchronize()
