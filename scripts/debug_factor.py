"""Component check of the factored first layer (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segger_b200 import ops
torch.manual_seed(0)
dev = "cuda"
N, G, D1, D2, R = 6000, 500, 128, 128, 256
ids = torch.randint(0, G, (N,), device=dev, dtype=torch.int32)
table = torch.randn(G, D1, device=dev)
x = torch.randn(N, D2, device=dev)
w = torch.randn(R, D1 + D2, device=dev) / 16
b = torch.randn(R, device=dev)
g = torch.randn(N, R, device=dev)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


ref_y = (torch.cat([table[ids.long()], x], 1).double() @ w.double().t() + b.double())
for exact in (0, 1):
    tab, _ = ops.linear_fwd(table, w[:, :D1], None, exact=2)
    print("table vs fp64", rel(tab, table.double() @ w[:, :D1].double().t()))
    y, _ = ops.linear_fwd(x, w[:, D1:], b, exact=exact, gather=(ids, tab))
    yd, _ = ops.linear_fwd(torch.cat([table[ids.long()], x], 1), w, b, exact=exact)
    print(f"exact={exact}: gather fwd vs fp64 {rel(y, ref_y):.2e}   dense fwd vs fp64 {rel(yd, ref_y):.2e}")
seg = ops.segment_sum_rows(g, ids, G)
seg_ref = torch.zeros(G, R, device=dev, dtype=torch.float64).index_add_(0, ids.long(), g.double())
print("segment_sum_rows vs fp64", rel(seg, seg_ref))
dt = ops.linear_dgrad(seg, w[:, :D1])
print("d_table vs fp64", rel(dt, seg_ref @ w[:, :D1].double()))
dw, _ = ops.linear_wgrad(seg, table, want_db=False)
print("dw_table vs fp64", rel(dw, seg_ref.t() @ table.double()))
dxd = ops.linear_dgrad(g, w[:, D1:])
print("dx dense vs fp64", rel(dxd, g.double() @ w[:, D1:].double()))
dwd, db = ops.linear_wgrad(g, x)
print("dw dense vs fp64", rel(dwd, g.double().t() @ x.double()), rel(db, g.double().sum(0)))
