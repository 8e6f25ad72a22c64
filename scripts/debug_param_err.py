"""Debug aid: per-parameter gradient error of the product encoder vs the fp64 oracle."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import make_models, synth_batch, to_dev, rel_err

ts, x, edges, pos, bat = synth_batch(6000, 60, seed=1)
ref, prod = make_models(ts.n_genes, ts.bd_x.shape[1], 128, 64, 64, 0, 2, seed=3)
ref.eval(); prod.eval()
r64 = copy.deepcopy(ref).double()
gen = torch.Generator().manual_seed(0)
out64 = r64({"tx": x["tx"], "bd": x["bd"].double()}, edges, {k: v.double() for k, v in pos.items()}, bat)
g = {k: torch.randn(v.shape, generator=gen) for k, v in out64.items()}
sum((out64[k] * g[k].double()).sum() for k in out64).backward()
out = prod(to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat))
sum((out[k] * g[k].cuda()).sum() for k in out).backward()
print("GEMM =", os.environ.get("SEGGER_B200_GEMM", "tc"))
for k in out: print("out", k, f"{rel_err(out[k], out64[k]):.2e}")
g64 = {n: p.grad for n, p in r64.named_parameters()}
for n, p in prod.named_parameters():
    if p.grad is not None and n in g64:
        print(f"{rel_err(p.grad, g64[n]):.2e}  {n}")
