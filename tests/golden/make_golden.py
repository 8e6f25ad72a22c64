"""Generates the golden fixtures in this directory.

The reference implementation (dpeerlab/segger) cannot be imported in the build image: it needs
torch_geometric, torch_scatter, cupy, rmm, lightning, polars (none installed, no network), and it
ships no tests or golden vectors of its own.  The fixtures therefore freeze
 * the scipy cKDTree result -- the reference's *actual* kNN call (neighbors.py:139-150) -- and
 * the outputs of the CPU oracle (oracle/), which restates PyG/torch_scatter semantics
   (PARITY UNPINNED for those, see oracle/__init__.py).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import neighbors_ref, pyg_ref  # noqa: E402
from oracle.ist_encoder_ref import ISTEncoderRef, TB, TT  # noqa: E402
from tests.util import random_graph  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = torch.Generator().manual_seed(1234)
    n_src, n_dst, E, H, C = 40, 30, 150, 2, 64
    x_l, x_r = torch.randn(n_src, H, C, generator=g), torch.randn(n_dst, H, C, generator=g)
    att, bias = torch.randn(1, H, C, generator=g) * 0.3, torch.randn(H * C, generator=g) * 0.1
    ei = random_graph(n_src, n_dst, E, seed=99)
    out = pyg_ref.gatv2_aggregate(x_l, x_r, ei, att, bias)
    torch.save(dict(x_l=x_l, x_r=x_r, att=att, bias=bias, edge_index=ei, out=out), os.path.join(HERE, "gatv2_small.pt"))

    torch.manual_seed(7)
    hp = dict(n_genes=20, bd_in=12, in_channels=16, hidden_channels=8, out_channels=8, n_mid_layers=1, n_heads=2)
    m = ISTEncoderRef(**hp).eval()
    N, M = 120, 9
    x = {"tx": torch.randint(0, 20, (N,), generator=g, dtype=torch.int32), "bd": torch.randn(M, 12, generator=g)}
    pos = {"tx": torch.rand(N, 2, generator=g) * 50, "bd": torch.rand(M, 2, generator=g) * 50}
    batch = {"tx": (torch.arange(N) // 60), "bd": torch.tensor([0] * 5 + [1] * 4)}
    ett, _ = neighbors_ref.kdtree_neighbors(pos["tx"].numpy(), 5, 8.0)
    etb = torch.stack([torch.arange(0, N, 3), torch.arange(0, N, 3) % M])
    edges = {TT: ett, TB: etb}
    with torch.no_grad():
        o = m(x, edges, pos, batch)
    torch.save(dict(hparams=hp, state_dict=m.state_dict(), x=x, pos=pos, batch=batch, edges=edges, out=o),
               os.path.join(HERE, "encoder_small.pt"))

    rng = np.random.default_rng(42)
    pts = rng.uniform(0, 40, (600, 2)).astype(np.float32)
    _, idx = neighbors_ref.kdtree_table(pts, 5, 5.0)
    np.savez(os.path.join(HERE, "knn_small.npz"), points=pts, k=5, max_dist=5.0, scipy_idx=idx)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
