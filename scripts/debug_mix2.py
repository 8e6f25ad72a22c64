"""Debug aid: which forward output (out/act vs softmax stats) carries the sensitivity?  MIX=out_legacy|stats_legacy"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from segger_b200 import ops
from tests.util import make_models, synth_batch, to_dev, rel_err

CFG = {"cfg0": (128, 64, 64, 0, 2), "cfg3": (64, 128, 128, 1, 4)}[os.environ.get("CFG", "cfg0")]
MIX = os.environ.get("MIX", "none")
of, ob = ops.gatv2_fwd, ops.gatv2_bwd
def fwd(*a, **k):
    os.environ["SEGGER_B200_GAT"] = "legacy"; rl = of(*a, **k)
    os.environ["SEGGER_B200_GAT"] = "quad"; rq = of(*a, **k)
    if a[4].n_dst > 100:
        d = (rl[0] - rq[0]).double()
        print(f"  out diff legacy-quad: max {float(d.abs().max()):.2e} mean signed {float(d.mean()):.2e} mean abs {float(d.abs().mean()):.2e}")
    if a[4].n_dst > 100 and rl[1] is not None:
        for nm, r in (("legacy", rl), ("quad", rq)):
            t = torch.nn.functional.gelu(r[0].double())
            d = (r[1].double() - t)
            i = int(d.abs().argmax())
            print(f"  {nm}: act vs fp64 gelu(out): max abs {float(d.abs().max()):.2e} at out={float(r[0].flatten()[i]):.4f} mean signed {float(d.mean()):.2e} rms {float(d.pow(2).mean().sqrt()):.2e}")
    if a[4].n_dst > 100 and rl[1] is not None:
        A, B = rl[1].double(), rq[1].double()
        rel = ((A - B).abs() / A.abs().clamp_min(1e-300))
        i = int(rel.argmax())
        print(f"  act legacy vs quad: n differing {int((A != B).sum())}/{A.numel()} max elementwise rel {float(rel.max()):.2e} at act={float(A.flatten()[i]):.3e}/{float(B.flatten()[i]):.3e} out={float(rl[0].flatten()[i]):.6f}/{float(rq[0].flatten()[i]):.6f}; nan {int(torch.isnan(B).sum())}")
        O1, O2 = rl[0].double(), rq[0].double()
        relo = ((O1 - O2).abs() / O1.abs().clamp_min(1e-300))
        print(f"  out legacy vs quad: n differing {int((O1 != O2).sum())} max elementwise rel {float(relo.max()):.2e}")
    if MIX == "out_legacy": return (rl[0], rl[1], rq[2], rq[3])
    if MIX == "stats_legacy": return (rq[0], rq[1], rl[2], rl[3])
    if MIX == "all_legacy": return rl
    if MIX == "noise":      # legacy out with random 1-ulp-level relative noise, act recomputed
        o = rl[0] * (1 + 6e-8 * torch.randn_like(rl[0]))
        return (o, torch.nn.functional.gelu(o), rl[2], rl[3])
    if MIX == "regelu":     # legacy out, act recomputed by torch
        return (rl[0], torch.nn.functional.gelu(rl[0]), rl[2], rl[3])
    if MIX == "quad_out_legacy_act": return (rq[0], rl[1], rq[2], rq[3])
    if MIX == "legacy_out_quad_act": return (rl[0], rq[1], rq[2], rq[3])
    return rq
def bwd(*a, **k):
    os.environ["SEGGER_B200_GAT"] = os.environ.get("BWD", "quad")
    if os.environ.get("CONSISTENT"):
        a = list(a); a[13] = 12345   # seed slot: debug trigger
    return ob(*a, **k)
ops.gatv2_fwd, ops.gatv2_bwd = fwd, bwd
ts, x, edges, pos, bat = synth_batch(6000, 60, seed=1)
ref, prod = make_models(ts.n_genes, ts.bd_x.shape[1], *CFG, seed=3)
ref.eval(); prod.eval()
r64 = copy.deepcopy(ref).double()
gen = torch.Generator().manual_seed(0)
out64 = r64({"tx": x["tx"], "bd": x["bd"].double()}, edges, {k: v.double() for k, v in pos.items()}, bat)
g = {k: torch.randn(v.shape, generator=gen) for k, v in out64.items()}
sum((out64[k] * g[k].double()).sum() for k in out64).backward()
out = prod(to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat))
sum((out[k] * g[k].cuda()).sum() for k in out).backward()
g64 = {n: p.grad for n, p in r64.named_parameters()}
out32 = ref(x, edges, pos, bat)
sum((out32[k] * g[k]).sum() for k in out32).backward()
g32 = {n: p.grad for n, p in ref.named_parameters()}
errs = []
for n, p in prod.named_parameters():
    if p.grad is None or n not in g64:
        continue
    noise = rel_err(g32[n], g64[n])
    e = rel_err(p.grad, g64[n])
    errs.append((e / max(1e-4, 2 * noise), e, noise, n))
errs.sort(reverse=True)
print(MIX, os.environ.get("BWD", "quad"), " | ".join(f"x{r:.2f} err {e:.1e} noise {nz:.1e} {n[12:50]}" for r, e, nz, n in errs[:3]))
r, e, nz, n = errs[0]
pg = dict(prod.named_parameters())[n].grad.double().cpu()
d = (pg - g64[n])
if d.dim() == 2:
    rows = d.pow(2).sum(1)
    print(f"error concentration for {n[12:60]}: top row {int(rows.argmax())} holds {float(rows.max() / rows.sum()):.3f} of squared error; "
          f"top-3 rows {[round(float(v), 3) for v in (rows.sort(descending=True).values[:3] / rows.sum())]}; max|d| {float(d.abs().max()):.2e} max|g| {float(g64[n].abs().max()):.2e}")
