"""Profiling aid: run one projection-shaped GEMM a few times (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segger_b200 import ops
M, N, K = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (1_000_000, 384, 256)))
mode = sys.argv[4] if len(sys.argv) > 4 else "fwd"
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
dy = torch.randn(M, N, device="cuda")
for _ in range(3):
    if mode == "fwd": ops.linear_fwd(x, w, b)
    elif mode == "dgrad": ops.linear_dgrad(dy, w)
    else: ops.linear_wgrad(dy, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    if mode == "fwd": ops.linear_fwd(x, w, b)
    elif mode == "dgrad": ops.linear_dgrad(dy, w)
    else: ops.linear_wgrad(dy, x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{mode} M={M} N={N} K={K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s (fp32-equivalent)")
