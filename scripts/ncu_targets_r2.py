"""Round-2 ncu targets: each dominant kernel of the cfg2 training / predict step launched ONCE on realistic inputs.

    ncu --set full --clock-control none --import-source on -k regex:"gatv2_|gemm_tf32x3|score_vec|segment_sum" \
        -o gpurun_out/r2_top python scripts/ncu_targets_r2.py
    python scripts/ncu_traffic.py gpurun_out/r2_top.ncu-rep r2x        # -> profiles/ncu_traffic.json (+ a text summary)

Launch order (scripts/ncu_traffic.py relies on it): gatv2 fwd tt, gatv2 bwd tt (dst pass, src pass), gemm fwd
[1M,128]x[128->256] (exact), gemm dgrad [1M,256]->[128], gemm wgrad 256x128 over 1M rows, segment sum [1M,256] by gene,
score/arg-max.  A number printed under ncu is never a bench value."""
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
    C, F = hid, H * hid
    ts, host = bench.build_workload("cfg2", 0, dev)
    d = bench.to_device(host, dev, bench.PRED_KEYS + ("e_tt",))
    csr_tt = ops.build_csr(d["e_tt"], n_tx, n_tx)
    g = torch.Generator(device="cuda").manual_seed(0)
    y = torch.randn(n_tx, 2 * F, device=dev, generator=g)
    att = torch.randn(F, device=dev, generator=g) * 0.1
    bias = torch.randn(F, device=dev, generator=g) * 0.1
    gt = torch.randn(n_tx, F, device=dev, generator=g)
    G = torch.empty(n_tx, 2 * F, device=dev)
    torch.cuda.synchronize()
    out, _, smax, sden, lg = ops.gatv2_fwd(y[:, :F], y[:, F:], att, bias, csr_tt, H, C, 0.2, 0.2, True, 7, True, want_logits=True)
    ops.gatv2_bwd(y[:, :F], y[:, F:], att, bias, out, gt, True, csr_tt, H, C, 0.2, 0.2, True, 7, smax, sden,
                  grad_x_l=G[:, :F], grad_x_r=G[:, F:], e_logit=lg)
    x = torch.nn.functional.gelu(torch.randn(n_tx, 128, device=dev, generator=g))
    w = torch.randn(2 * F, 128, device=dev, generator=g) / 11
    ops.linear_fwd(x, w, None, exact=1)
    ops.linear_dgrad(G, w)
    ops.linear_wgrad(G, x)
    ops.segment_sum_rows(G, d["tx_x"], ts.n_genes)
    e_tx = torch.nn.functional.normalize(torch.randn(n_tx, out_c, device=dev, generator=g))
    e_bd = torch.nn.functional.normalize(torch.randn(n_cells, out_c, device=dev, generator=g))
    ops.score_argmax(e_tx, e_bd, d["e_pred"], d["bd_index"])
    torch.cuda.synchronize()
    print("ncu targets launched")


if __name__ == "__main__":
    main()
