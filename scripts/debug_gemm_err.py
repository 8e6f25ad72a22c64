"""Debug aid: per-call error of the GEMM back end against float64 torch on the real encoder data."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segger_b200 import ops
from tests.util import make_models, synth_batch, to_dev, rel_err

orig_fwd, orig_dg, orig_wg = ops.linear_fwd, ops.linear_dgrad, ops.linear_wgrad

def fwd(x, w, b, act=0, y=None, y_act=None, want_pre=True):
    r = orig_fwd(x, w, b, act, y, y_act, want_pre)
    ref = x.double() @ w.double().t() + (b.double() if b is not None else 0)
    print(f"fwd   M={x.size(0):6d} N={w.size(0):4d} K={x.size(1):4d} err={rel_err(r[0], ref):.2e}")
    return r

def dg(dy, w, dx=None, accumulate=False, act=0, act_pre=None):
    r = orig_dg(dy, w, dx, accumulate, act, act_pre)
    if act == 0 and not accumulate:
        ref = dy.double() @ w.double()
        print(f"dgrad M={dy.size(0):6d} N={dy.size(1):4d} K={w.size(1):4d} err={rel_err(r, ref):.2e}  absmax dy={float(dy.abs().max()):.2e}")
    return r

def wg(dy, x, dw=None, db=None, want_db=True, accumulate=False):
    r = orig_wg(dy, x, dw, db, want_db, accumulate)
    ref = dy.double().t() @ x.double()
    print(f"wgrad M={dy.size(0):6d} N={dy.size(1):4d} K={x.size(1):4d} err={rel_err(r[0], ref):.2e}")
    return r

ops.linear_fwd, ops.linear_dgrad, ops.linear_wgrad = fwd, dg, wg
ts, x, edges, pos, bat = synth_batch(6000, 60, seed=1)
ref, prod = make_models(ts.n_genes, ts.bd_x.shape[1], 128, 64, 64, 0, 2, seed=3)
prod.eval()
out = prod(to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat))
g = torch.Generator().manual_seed(0)
loss = sum((out[k] * torch.randn(out[k].shape, generator=g).cuda()).sum() for k in ("tx", "bd"))
loss.backward()
