"""Launches each dominant kernel of the cfg2 training step ONCE on realistic inputs, for `ncu --set full`:

    ncu --set full --clock-control none --import-source on -k regex:"gemm_tf32x3|gatv2_|posfreq|score_" \
        -o gpurun_out/r1c_top python scripts/ncu_targets.py

Order of the captured launches: gatv2 fwd (tt), gatv2 bwd dst + src (tt), gatv2 fwd (tb), bwd dst + src (tb),
gemm fwd 1Mx384x256 (4 terms), fwd (3 terms), dgrad 1Mx384x256, wgrad 1Mx384x256, fwd 2Mx64x256 (pos MLP),
wgrad 2Mx64x256, posfreq, score.  A number printed under ncu is never a bench value.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from segger_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    n_tx, n_cells, k, in_c, hid, out_c, n_mid, H = bench.WORKLOADS["cfg2"]
    C = hid
    F = H * C
    ts, host = bench.build_workload("cfg2", 0, dev)
    d = bench.to_device(host, dev, bench.PRED_KEYS + ("e_tt",))
    csr_tt = ops.build_csr(d["e_tt"], n_tx, n_tx)
    csr_tb = ops.build_csr(d["e_tb"], n_tx, n_cells)
    g = torch.Generator(device="cuda").manual_seed(0)
    y = torch.randn(n_tx, 3 * F, device=dev, generator=g)
    y_bd = torch.randn(n_cells, F, device=dev, generator=g)
    att = torch.randn(F, device=dev, generator=g) * 0.1
    bias = torch.randn(F, device=dev, generator=g) * 0.1
    gt = torch.randn(n_tx, F, device=dev, generator=g)
    gb = torch.randn(n_cells, F, device=dev, generator=g)
    G = torch.empty(n_tx, 3 * F, device=dev)
    torch.cuda.synchronize()
    # --- message passing
    out, _, smax, sden = ops.gatv2_fwd(y[:, :F], y[:, F:2 * F], att, bias, csr_tt, H, C, 0.2, 0.2, True, 7, True)
    ops.gatv2_bwd(y[:, :F], y[:, F:2 * F], att, bias, out, gt, True, csr_tt, H, C, 0.2, 0.2, True, 7, smax, sden,
                  grad_x_l=G[:, :F], grad_x_r=G[:, F:2 * F])
    out_b, _, smax_b, sden_b = ops.gatv2_fwd(y[:, 2 * F:], y_bd, att, bias, csr_tb, H, C, 0.2, 0.2, True, 9, True)
    ops.gatv2_bwd(y[:, 2 * F:], y_bd, att, bias, out_b, gb, True, csr_tb, H, C, 0.2, 0.2, True, 9, smax_b, sden_b,
                  grad_x_l=G[:, 2 * F:])
    # --- projections
    x = torch.nn.functional.gelu(torch.randn(n_tx, 256, device=dev, generator=g))
    w = torch.randn(3 * F, 256, device=dev, generator=g) / 16
    ops.linear_fwd(x, w, None, exact=1)
    ops.linear_fwd(x, w, None, exact=0)
    ops.linear_dgrad(G, w)
    ops.linear_wgrad(G, x)
    f2 = torch.randn(2 * n_tx, 256, device=dev, generator=g)
    w0 = torch.randn(64, 256, device=dev, generator=g) / 16
    ops.linear_fwd(f2, w0, None, exact=1)
    dy0 = torch.randn(2 * n_tx, 64, device=dev, generator=g)
    ops.linear_wgrad(dy0, f2)
    # --- input stage / scoring
    ops.posfreq(d["tx_pos"], d["tx_batch"], ts.n_tiles, 256, ops.sinusoid_freqs(256, 10000, dev))
    e_tx = torch.nn.functional.normalize(torch.randn(n_tx, out_c, device=dev, generator=g))
    e_bd = torch.nn.functional.normalize(torch.randn(n_cells, out_c, device=dev, generator=g))
    ops.score_argmax(e_tx, e_bd, d["e_pred"], d["bd_index"])
    torch.cuda.synchronize()
    print("ncu targets launched")


if __name__ == "__main__":
    main()
