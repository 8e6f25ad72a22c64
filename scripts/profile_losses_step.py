"""Per-kernel time of the training step with segger's own losses (configs[1]) minus the synthetic-loss step:
torch.profiler (CUPTI).  usage: python scripts/profile_losses_step.py [n_steps]"""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from segger_b200 import ops  # noqa: E402
from segger_b200.hetero import HeteroBatch  # noqa: E402
from segger_b200.lightning_model import LitISTEncoder  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    dev = torch.device("cuda", 0)
    n_tx, n_cells, k, in_c, hid, out_c, n_mid, heads = bench.WORKLOADS["cfg2"]
    ts, host = bench.build_workload("cfg2", 0, dev)
    torch.manual_seed(0)
    lit = LitISTEncoder(ts.n_genes, in_channels=in_c, hidden_channels=hid, out_channels=out_c, n_mid_layers=n_mid,
                        n_heads=heads).to(dev)
    lit.train()
    d = bench.to_device(host, dev, bench.TRAIN_KEYS)
    with torch.no_grad():
        lit.model(*bench.model_inputs(d))
    opt = torch.optim.Adam([p for p in lit.model.parameters() if p.requires_grad], lr=1e-3, fused=True)
    g = torch.Generator().manual_seed(5)

    def sim(c):
        a = torch.rand(c, c, generator=g) * 2 - 1
        return ((a + a.t()) / 2).contiguous()

    lit.setup_losses(sim(20), sim(10))
    lit.set_epoch(5, 10)
    b = HeteroBatch()
    b["tx"]["x"], b["tx"]["pos"], b["tx"]["batch"] = d["tx_x"], d["tx_pos"], d["tx_batch"]
    b["bd"]["x"], b["bd"]["pos"], b["bd"]["batch"] = d["bd_x"], d["bd_pos"], d["bd_batch"]
    b["tx"]["mask"] = torch.ones(n_tx, dtype=torch.bool, device=dev)
    b["bd"]["mask"] = torch.ones(n_cells, dtype=torch.bool, device=dev)
    b["tx"]["cluster"] = torch.randint(0, 20, (n_tx,), generator=g).to(dev)
    b["bd"]["cluster"] = torch.randint(0, 10, (n_cells,), generator=g).to(dev)
    b[bench.TT]["edge_index"], b[bench.TB]["edge_index"] = d["e_tt"], d["e_tb"]
    t_tx = torch.randn(n_tx, out_c, device=dev)
    t_bd = torch.randn(n_cells, out_c, device=dev)

    def step_losses():
        ops.CSR_CACHE.clear()
        opt.zero_grad(set_to_none=True)
        lit.training_step(b, 0).backward()
        opt.step()

    def step_synth():
        ops.CSR_CACHE.clear()
        opt.zero_grad(set_to_none=True)
        out = lit.model(*bench.model_inputs(d))
        ((out["tx"] * t_tx).sum() / n_tx + (out["bd"] * t_bd).sum() / n_cells).backward()
        opt.step()

    from torch.profiler import ProfilerActivity, profile
    tabs = {}
    for name, fn in (("losses", step_losses), ("synthetic", step_synth)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(steps):
                fn()
            torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for ev in prof.events():
            if ev.device_type.name != "CUDA":
                continue
            n = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").replace("sgb::", "").replace("at::native::", "")
            n = re.sub(r"\(.*", "", n)[:90]
            a = agg.setdefault(n, [0, 0.0])
            a[0] += 1
            a[1] += ev.device_time_total
        tabs[name] = {k_: (c / steps, t / steps) for k_, (c, t) in agg.items()}
    tot = {k_: sum(t for _, t in v.values()) for k_, v in tabs.items()}
    print(f"kernel time per step: losses {tot['losses']:.0f} us, synthetic {tot['synthetic']:.0f} us, difference {tot['losses'] - tot['synthetic']:.0f} us")
    diff = []
    for k_ in set(tabs["losses"]) | set(tabs["synthetic"]):
        cl, tl = tabs["losses"].get(k_, (0, 0.0))
        cs, t_s = tabs["synthetic"].get(k_, (0, 0.0))
        diff.append((tl - t_s, cl - cs, k_))
    for dt, dc, k_ in sorted(diff, reverse=True)[:30]:
        print(f"{dt:9.1f} us  x{dc:+5.1f}  {k_}")


if __name__ == "__main__":
    main()
