"""One projection GEMM launch for ncu: python scripts/run_gemm.py [fwd|dgrad|wgrad] [M] [N] [K] [exact]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segger_b200 import ops
op = sys.argv[1] if len(sys.argv) > 1 else "fwd"
M, N, K = (int(v) for v in (sys.argv[2:5] if len(sys.argv) > 4 else (262144, 384, 256)))
exact = int(sys.argv[5]) if len(sys.argv) > 5 else 0
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda"); dy = torch.randn(M, N, device="cuda")
for _ in range(2):
    if op == "fwd": ops.linear_fwd(x, w, None, exact=exact)
    elif op == "dgrad": ops.linear_dgrad(dy, w)
    else: ops.linear_wgrad(dy, x)
torch.cuda.synchronize()
