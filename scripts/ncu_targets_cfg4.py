"""ncu targets for BASELINE configs[3] (k=20, heads=4, hidden=128 -> F=512): the fused GATv2 forward / backward launched once.

    ncu --set full --clock-control none --import-source on -k regex:"gatv2_" -o gpurun_out/r2_cfg4 python scripts/ncu_targets_cfg4.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from segger_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    n_tx, n_cells, k, in_c, hid, out_c, n_mid, H = bench.WORKLOADS["cfg4"]
    C, F = hid, H * hid
    ts, host = bench.build_workload("cfg4", 0, dev)
    d = bench.to_device(host, dev, ("e_tt",))
    csr_tt = ops.build_csr(d["e_tt"], n_tx, n_tx)
    g = torch.Generator(device="cuda").manual_seed(0)
    y = torch.randn(n_tx, 2 * F, device=dev, generator=g)
    att = torch.randn(F, device=dev, generator=g) * 0.1
    bias = torch.randn(F, device=dev, generator=g) * 0.1
    gt = torch.randn(n_tx, F, device=dev, generator=g)
    G = torch.empty(n_tx, 2 * F, device=dev)
    torch.cuda.synchronize()
    out, _, smax, sden = ops.gatv2_fwd(y[:, :F], y[:, F:], att, bias, csr_tt, H, C, 0.2, 0.2, True, 7, True)
    ops.gatv2_bwd(y[:, :F], y[:, F:], att, bias, out, gt, True, csr_tt, H, C, 0.2, 0.2, True, 7, smax, sden,
                  grad_x_l=G[:, :F], grad_x_r=G[:, F:])
    torch.cuda.synchronize()
    print("cfg4 ncu targets launched, E =", csr_tt.E)


if __name__ == "__main__":
    main()
