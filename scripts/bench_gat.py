"""Times the fused GATv2 kernels alone (tx-neighbors-tx and tx-belongs-bd of a bench workload),
legacy row-per-warp path vs the cp.async quad path, L2 flushed between launches.

    python scripts/bench_gat.py [--workload cfg2] [--reps 10]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from segger_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--paths", default="legacy,quad")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    n_tx, n_cells, k, in_c, hid, out_c, n_mid, H = bench.WORKLOADS[a.workload]
    C = hid
    F = H * C
    ts, host = bench.build_workload(a.workload, 0, dev)
    d = bench.to_device(host, dev, bench.TRAIN_KEYS)
    csr_tt = ops.build_csr(d["e_tt"], n_tx, n_tx)
    csr_tb = ops.build_csr(d["e_tb"], n_tx, n_cells)
    g = torch.Generator(device="cuda").manual_seed(0)
    y = torch.randn(n_tx, 3 * F, device=dev, generator=g)
    y_bd = torch.randn(n_cells, F, device=dev, generator=g)
    att = torch.randn(F, device=dev, generator=g) * 0.1
    bias = torch.randn(F, device=dev, generator=g) * 0.1
    gt = torch.randn(n_tx, F, device=dev, generator=g)
    gb = torch.randn(n_cells, F, device=dev, generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    touched_tt = int(torch.unique(d["e_tt"][0]).numel())
    touched_tb = int(torch.unique(d["e_tb"][0]).numel())
    hbm = bench.peaks()[0]

    def time_it(fn):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(a.reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / a.reps * 1e-3

    res = {}
    outs = {}
    for path in a.paths.split(","):
        # "quad_nosort": the quad kernels with the warp's rows in chunk order (A/B of the degree-ordered quads)
        os.environ["SEGGER_B200_GAT"] = "quad" if path.startswith("quad") else path
        os.environ["SEGGER_B200_GAT_SORT"] = "0" if path == "quad_nosort" else "1"
        for name, (xl, xr, csr, gg, ns, nd, touched) in {
            "tt": (y[:, :F], y[:, F:2 * F], csr_tt, gt, n_tx, n_tx, touched_tt),
            "tb": (y[:, 2 * F:], y_bd, csr_tb, gb, n_tx, n_cells, touched_tb),
        }.items():
            fb, bb = bench.gat_bytes(touched, nd, ns, csr.E, H, C)
            # "quad_nolg": the backward recomputes the logits (A/B of the saved-logit dst pass)
            want_lg = path in ("quad", "quad_nosort")
            out, act, smax, sden, lg = ops.gatv2_fwd(xl, xr, att, bias, csr, H, C, 0.2, 0.2, True, 7, True, want_logits=True)
            if not want_lg:
                lg = None
            G = torch.zeros(ns, F, device=dev)
            Gr = torch.zeros(nd, F, device=dev)
            r = ops.gatv2_bwd(xl, xr, att, bias, out, gg, True, csr, H, C, 0.2, 0.2, True, 7, smax, sden,
                              grad_x_l=G, grad_x_r=Gr, e_logit=lg)
            outs[(path, name)] = (act.clone(), G.clone(), Gr.clone(), r[2].clone(), r[3].clone())
            tf = time_it(lambda: ops.gatv2_fwd(xl, xr, att, bias, csr, H, C, 0.2, 0.2, True, 7, True, want_logits=want_lg))
            tfe = time_it(lambda: ops.gatv2_fwd(xl, xr, att, bias, csr, H, C, 0.2, 0.0, False, 7, True))
            tb = time_it(lambda: ops.gatv2_bwd(xl, xr, att, bias, out, gg, True, csr, H, C, 0.2, 0.2, True, 7, smax,
                                               sden, grad_x_l=G, grad_x_r=Gr, e_logit=lg))
            res[f"{path}.{name}"] = {
                "fwd_ms": tf * 1e3, "fwd_frac": fb / tf / 1e9 / hbm, "fwd_eval_ms": tfe * 1e3,
                "bwd_ms": tb * 1e3, "bwd_frac": bb / tb / 1e9 / hbm, "E": csr.E}
    paths = a.paths.split(",")
    if len(paths) == 2:
        for name in ("tt", "tb"):
            a_, b_ = outs[(paths[0], name)], outs[(paths[1], name)]
            res[f"maxdiff.{name}"] = [float((u - v).abs().max() / v.abs().max().clamp_min(1e-30)) for u, v in zip(a_, b_)]
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
