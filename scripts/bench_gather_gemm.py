"""Times the first-layer projection GEMM with and without the gene-table gather epilogue (cfg2 shape: 1M x 128 -> 256).

    python scripts/bench_gather_gemm.py [--rows 1000000]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segger_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    M, K, N, G = a.rows, 128, 256, 500
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.nn.functional.gelu(torch.randn(M, K, device=dev, generator=g))
    w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
    b = torch.randn(N, device=dev, generator=g)
    tab = torch.randn(G, N, device=dev, generator=g)
    ids = torch.randint(0, G, (M,), device=dev, generator=g)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def time_it(fn):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(a.reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / a.reps

    for exact in (0, 1):
        t0 = time_it(lambda: ops.linear_fwd(x, w, b, exact=exact))
        t1 = time_it(lambda: ops.linear_fwd(x, w, b, exact=exact, gather=(ids, tab)))
        y0 = ops.linear_fwd(x, w, b, exact=exact)[0]
        y1 = ops.linear_fwd(x, w, b, exact=exact, gather=(ids, tab))[0]
        err = float((y1 - (y0 + tab[ids])).abs().max())
        print(f"exact={exact}: plain {t0:.3f} ms, + table gather {t1:.3f} ms, max |gather - (plain + tab[ids])| = {err:.2e}")


if __name__ == "__main__":
    main()
