"""Experiment: fused GATv2 kernels on the tx-neighbors-tx graph with transcripts in the synthetic (random within tile) order
vs Morton order within tile.  usage: python scripts/exp_morton.py cfg2|cfg4"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from segger_b200 import ops  # noqa: E402


def morton(pos, bits=12):
    lo, hi = pos.min(0).values, pos.max(0).values
    q = ((pos - lo) / (hi - lo).clamp_min(1e-9) * (2 ** bits - 1)).long()
    def spread(v):
        out = torch.zeros_like(v)
        for b in range(bits):
            out |= ((v >> b) & 1) << (2 * b)
        return out
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1)


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    dev = torch.device("cuda", 0)
    n_tx, n_cells, k, in_c, hid, out_c, n_mid, H = bench.WORKLOADS[wl]
    C, F = hid, H * hid
    ts, host = bench.build_workload(wl, 0, dev)
    d = bench.to_device(host, dev, ("e_tt", "tx_pos", "tx_batch"))
    g = torch.Generator(device="cuda").manual_seed(0)
    y = torch.randn(n_tx, 2 * F, device=dev, generator=g)
    att = torch.randn(F, device=dev, generator=g) * 0.1
    bias = torch.randn(F, device=dev, generator=g) * 0.1
    gt = torch.randn(n_tx, F, device=dev, generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def time_it(fn, reps=5):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps

    def run(ei, yy, gg, tag):
        csr = ops.build_csr(ei, n_tx, n_tx)
        out, _, smax, sden = ops.gatv2_fwd(yy[:, :F], yy[:, F:], att, bias, csr, H, C, 0.2, 0.2, True, 7, True)
        G = torch.empty(n_tx, 2 * F, device=dev)
        tf = time_it(lambda: ops.gatv2_fwd(yy[:, :F], yy[:, F:], att, bias, csr, H, C, 0.2, 0.2, True, 7, True))
        tb = time_it(lambda: ops.gatv2_bwd(yy[:, :F], yy[:, F:], att, bias, out, gg, True, csr, H, C, 0.2, 0.2, True, 7, smax,
                                           sden, grad_x_l=G[:, :F], grad_x_r=G[:, F:]))
        print(f"{wl} {tag:8s} fwd {tf:7.3f} ms  bwd {tb:7.3f} ms")
        return out

    o1 = run(d["e_tt"], y, gt, "as-is")
    key = d["tx_batch"].long() * (1 << 24) + morton(d["tx_pos"].double())
    perm = torch.argsort(key, stable=True)
    inv = torch.empty_like(perm); inv[perm] = torch.arange(n_tx, device=dev)
    ei2 = inv[d["e_tt"].long()].to(d["e_tt"].dtype)
    o2 = run(ei2.contiguous(), y[perm].contiguous(), gt[perm].contiguous(), "morton")
    print("max diff after un-permuting:", float((o2[inv] - o1).abs().max()))


if __name__ == "__main__":
    main()
