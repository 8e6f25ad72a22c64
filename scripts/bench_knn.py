"""kNN / radius graph construction sweep (BASELINE configs[4]): sgb_knn2d on uniform and clustered point sets,
N in {1, 2, 5, 10, 20, 50} M, k in {5, 20}, r = 5 um at Xenium-like density (0.5 points / um^2).

    python scripts/bench_knn.py [--max-n 50000000] [--check-n 1000000]

Per case: ms per build (plan + bin + query + table -> COO, CUDA events, best of 3), points/s, edges, achieved GB/s on the
algorithmic bytes of SURVEY 8d ((24 + 8k) N) against the measured HBM peak; the smallest case of every kind is checked
bit-exactly against scipy's cKDTree (the reference's own call, neighbors.py:139-150) when --check-n allows."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from segger_b200.neighbors import kdtree_neighbors  # noqa: E402


def points(kind, n, rng):
    side = (n / 0.5) ** 0.5
    if kind == "uniform":
        return rng.uniform(0, side, (n, 2)).astype(np.float32)
    n_c = max(1, n // 100)                                   # 100 transcripts per cell-like cluster, sigma 3 um
    c = rng.uniform(0, side, (n_c, 2))
    return (c[rng.integers(0, n_c, n)] + rng.normal(0, 3.0, (n, 2))).astype(np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-n", type=int, default=50_000_000)
    ap.add_argument("--check-n", type=int, default=1_000_000)
    a = ap.parse_args()
    hbm = bench.peaks()[0]
    rng = np.random.default_rng(0)
    rows = []
    for kind in ("uniform", "clustered"):
        for n in (1_000_000, 2_000_000, 5_000_000, 10_000_000, 20_000_000, 50_000_000):
            if n > a.max_n:
                continue
            pts = points(kind, n, rng)
            dev_pts = torch.from_numpy(pts).cuda()
            for k in (5, 20):
                best = 1e30
                for _ in range(3):
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ei, _ = kdtree_neighbors(dev_pts, k, 5.0, device_output=True, device="cuda")
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                row = dict(kind=kind, n=n, k=k, ms=best, points_per_s=n / best * 1e3, edges=int(ei.size(1)),
                           alg_gbs=(24 + 8 * k) * n / best / 1e6, frac_hbm=(24 + 8 * k) * n / best / 1e6 / hbm)
                if n <= a.check_n:
                    from oracle import neighbors_ref
                    canon, _, _, _ = neighbors_ref.canonical_knn_table(pts, k, 5.0)
                    ce, _ = neighbors_ref.knn_to_edge_index(torch.from_numpy(canon), padding_value=n)
                    row["bit_exact_vs_scipy"] = bool(torch.equal(ei.cpu(), ce))
                rows.append(row)
                print(json.dumps(row), flush=True)
                del ei
            del dev_pts


if __name__ == "__main__":
    main()
