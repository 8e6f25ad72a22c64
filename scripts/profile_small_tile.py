"""Per-kernel time of ONE configs[0] training step (50k transcripts): torch.profiler (CUPTI) over eager steps.
usage: python scripts/profile_small_tile.py [n_steps] -> table on stdout (sum over steps / n_steps)."""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from segger_b200.lightning_model import LitISTEncoder  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    dev = torch.device("cuda", 0)
    n_tx, n_cells, k, in_c, hid, out_c, n_mid, heads = bench.WORKLOADS["cfg1"]
    ts, host = bench.build_workload("cfg1", 0, dev)
    torch.manual_seed(0)
    lit = LitISTEncoder(ts.n_genes, in_channels=in_c, hidden_channels=hid, out_channels=out_c, n_mid_layers=n_mid,
                        n_heads=heads).to(dev)
    lit.train()
    model = lit.model
    d = bench.to_device(host, dev, bench.TRAIN_KEYS)
    inputs = bench.model_inputs(d)
    with torch.no_grad():
        model(*inputs)
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3, fused=True, capturable=True)
    t_tx = torch.randn(n_tx, out_c, device=dev)
    t_bd = torch.randn(n_cells, out_c, device=dev)

    def step():
        opt.zero_grad(set_to_none=False)
        out = model(*inputs)
        ((out["tx"] * t_tx).sum() / n_tx + (out["bd"] * t_bd).sum() / n_cells).backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    tot = 0.0
    for ev in prof.events():
        if ev.device_type.name != "CUDA":
            continue
        n = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").replace("sgb::", "").replace("at::native::", "")
        n = re.sub(r"\(.*", "", n)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        tot += a[1] * 0
    tot = sum(v[1] for v in agg.values())
    print(f"{steps} steps: {tot / steps:.1f} us of kernel time per step, {sum(v[0] for v in agg.values()) / steps:.0f} launches per step")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"{t / steps:9.1f} us {100 * t / tot:5.1f}%  x{c / steps:5.1f}  {n[:110]}")


if __name__ == "__main__":
    main()
