"""Where do the factored and the dense first layer differ? (run on the GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import make_models, synth_batch, to_dev
ts, x, edges, pos, bat = synth_batch(6000, 60, seed=31)
ref, prod = make_models(ts.n_genes, ts.bd_x.shape[1], 128, 64, 64, 0, 2, seed=2)
prod.eval(); ref.eval()
args = (to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat))
gen = torch.Generator().manual_seed(0)
g = {"tx": torch.randn(6000, 64, generator=gen), "bd": torch.randn(60, 64, generator=gen)}
import copy
r64 = copy.deepcopy(ref).double()
out = r64({"tx": x["tx"], "bd": x["bd"].double()}, edges, {k: v.double() for k, v in pos.items()}, bat)
sum((out[k] * g[k].double()).sum() for k in g).backward()
truth = {n: p.grad for n, p in r64.named_parameters()}
res = {}
for flag in ("0", "1"):
    os.environ["SEGGER_B200_FACTOR"] = flag
    prod.zero_grad()
    o = prod(*args)
    sum((o[k] * g[k].cuda()).sum() for k in g).backward()
    res[flag] = {n: p.grad.detach().cpu().double() for n, p in prod.named_parameters() if p.grad is not None}
for n in res["0"]:
    t = truth[n]
    e0 = float((res["0"][n] - t).abs().max() / t.abs().max())
    e1 = float((res["1"][n] - t).abs().max() / t.abs().max())
    if max(e0, e1) > 2e-6:
        print(f"{n:70s} dense {e0:.2e}  factored {e1:.2e}")
n = "lin_first.tx.weight"
d = (res["1"][n] - truth[n]).abs().max(1).values
top = torch.topk(d, 5)
cnt = torch.bincount(x["tx"].long(), minlength=ts.n_genes)
print("worst rows", top.indices.tolist(), top.values.tolist(), "counts", cnt[top.indices].tolist(), "row max", truth[n].abs().max(1).values[top.indices].tolist())
