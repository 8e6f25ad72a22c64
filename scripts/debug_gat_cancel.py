"""Debug aid: conditioning of the fused GATv2 backward on SMOOTH features (neighbouring rows almost
equal, so sum_e delta_e cancels).  Compares every (fwd path, bwd path) combination of the two kernel
generations against an fp64 torch oracle; errors are relative to max|truth| per tensor and for the
column sums (what lin_r.bias / lin_l.bias gradients see)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle.pyg_ref import gatv2_aggregate
from segger_b200 import ops
from tests.util import rel_err

torch.manual_seed(0)
N, k, H, C = 20000, 5, 2, 64
F = H * C
smooth = float(os.environ.get("SMOOTH", "1e-2"))
# kNN-like graph on a line: i receives from i-2..i+2 (clipped), plus self
src = torch.arange(N).repeat_interleave(k)
dst = (src + torch.tensor([-2, -1, 0, 1, 2]).repeat(N)).clamp(0, N - 1)
ei = torch.stack([src, dst])
base = torch.randn(1, F)
walk = torch.cumsum(torch.randn(N, F) * smooth, 0)
x_l = (base + walk).contiguous()
x_r = (torch.randn(1, F) + torch.cumsum(torch.randn(N, F) * smooth, 0)).contiguous()
att = torch.randn(F) * 0.3
bias = torch.rand(F) * 0.4 - 0.2
g = torch.randn(N, F)

xl64 = x_l.double().view(N, H, C).requires_grad_()
xr64 = x_r.double().view(N, H, C).requires_grad_()
a64 = att.double().view(1, H, C).requires_grad_()
b64 = bias.double().requires_grad_()
out64 = gatv2_aggregate(xl64, xr64, ei, a64, b64)
(out64 * g.double()).sum().backward()
truth = {"gxl": xl64.grad.view(N, F), "gxr": xr64.grad.view(N, F), "gatt": a64.grad.view(F), "gbias": b64.grad}

# fp32 oracle for scale
xl32 = x_l.view(N, H, C).clone().requires_grad_()
xr32 = x_r.view(N, H, C).clone().requires_grad_()
a32 = att.view(1, H, C).clone().requires_grad_()
b32 = bias.clone().requires_grad_()
out32 = gatv2_aggregate(xl32, xr32, ei, a32, b32)
(out32 * g).sum().backward()
print(f"fp32 oracle: out {rel_err(out32, out64):.2e} gxl {rel_err(xl32.grad.view(N, F), truth['gxl']):.2e} "
      f"gxr {rel_err(xr32.grad.view(N, F), truth['gxr']):.2e} colsum gxr "
      f"{rel_err(xr32.grad.view(N, F).sum(0), truth['gxr'].sum(0)):.2e} gatt {rel_err(a32.grad.view(F), truth['gatt']):.2e}")
print("scale: max|gxr| %.3e  max|colsum gxr| %.3e  max|gxl| %.3e" % (
    truth["gxr"].abs().max(), truth["gxr"].sum(0).abs().max(), truth["gxl"].abs().max()))

dev = "cuda"
csr = ops.build_csr(ei.to(dev), N, N)
xl, xr, at, bi, gg = x_l.to(dev), x_r.to(dev), att.to(dev), bias.to(dev), g.to(dev)
for fpath in ("legacy", "quad"):
    os.environ["SEGGER_B200_GAT"] = fpath
    out, _, smax, sden = ops.gatv2_fwd(xl, xr, at, bi, csr, H, C, 0.2, 0.0, False, 0, False)
    torch.cuda.synchronize()
    for bpath in ("legacy", "quad"):
        os.environ["SEGGER_B200_GAT"] = bpath
        gl, gr, ga, gb = ops.gatv2_bwd(xl, xr, at, bi, out, gg, False, csr, H, C, 0.2, 0.0, False, 0, smax, sden)
        torch.cuda.synchronize()
        print(f"fwd={fpath:6s} bwd={bpath:6s}: out {rel_err(out, out64):.2e} gxl {rel_err(gl, truth['gxl']):.2e} "
              f"gxr {rel_err(gr, truth['gxr']):.2e} colsum gxr {rel_err(gr.sum(0), truth['gxr'].sum(0)):.2e} "
              f"colsum gxl {rel_err(gl.sum(0), truth['gxl'].sum(0)):.2e} gatt {rel_err(ga, truth['gatt']):.2e}")
