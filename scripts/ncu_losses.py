"""Runs LitISTEncoder.training_step (real losses) on the cfg2 graph a few times, for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_losses.csv python scripts/ncu_losses.py
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from segger_b200 import ops  # noqa: E402
from segger_b200.distributed import FlatGradAllReduce, trainable_parameters  # noqa: E402
from segger_b200.lightning_model import LitISTEncoder  # noqa: E402

dev = torch.device("cuda", 0)
n_tx, n_cells, k, in_c, hid, out_c, n_mid, heads = bench.WORKLOADS["cfg2"]
ts, host = bench.build_workload("cfg2", 0, dev)
torch.manual_seed(0)
lit = LitISTEncoder(ts.n_genes, in_channels=in_c, hidden_channels=hid, out_channels=out_c, n_mid_layers=n_mid, n_heads=heads).to(dev)
d = bench.to_device(host, dev, bench.TRAIN_KEYS)
with torch.no_grad():
    lit.model.train()(*bench.model_inputs(d))
params = trainable_parameters(lit.model)
flat = FlatGradAllReduce(params)
opt = torch.optim.Adam(params, lr=1e-3, fused=True)

def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

print(bench.losses_step_time(lit, d, n_tx, n_cells, dev, flat, opt, timed, 2, 0.0))
