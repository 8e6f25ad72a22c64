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
    make_losses()
    make_pip()
    print("golden fixtures written to", HERE)


def make_losses():
    """losses_small.pt: freezes oracle/triplet_loss_ref.py (sampler indices for injected uniforms, the three loss
    values and the embedding gradients)."""
    from oracle import triplet_loss_ref as R
    g = torch.Generator().manual_seed(4321)
    C, N, D, M = 6, 90, 16, 12
    a = torch.rand(C, C, generator=g) * 2 - 1
    sim = ((a + a.t()) / 2).contiguous()
    labels = torch.randint(0, 5, (N,), generator=g)
    uni = [torch.rand(N, generator=g) for _ in range(4)]
    emb = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    bd = torch.nn.functional.normalize(torch.randn(M, D, generator=g), dim=-1)
    ei = torch.stack([torch.randperm(N, generator=g)[:40], torch.randint(0, M, (40,), generator=g)])
    dst_neg = (ei[1] + torch.randint(1, M, (40,), generator=g)) % M
    pos, neg, dp, dn = R.FastTripletSelectorRef(sim.clone()).sample_triplets(labels, uni)
    e = emb.clone().requires_grad_()
    l_t = R.triplet_loss_ref(e, pos, neg, 0.3)
    l_m = R.metric_loss_ref(e, pos, neg, dp, dn)
    l_s = R.segmentation_loss_ref(e, bd, ei, dst_neg, "triplet", 0.4)
    l_b = R.segmentation_loss_ref(e, bd, ei, dst_neg, "bce", 0.4)
    (l_t + l_m + l_s + l_b).backward()
    torch.save(dict(similarity=sim, labels=labels, uniforms=uni, emb=emb, bd=bd, edge_index=ei, dst_neg=dst_neg,
                    positives=pos, negatives=neg, dists_pos=dp, dists_neg=dn, loss_triplet=l_t.detach(),
                    loss_metric=l_m.detach(), loss_seg_triplet=l_s.detach(), loss_seg_bce=l_b.detach(), grad=e.grad),
               os.path.join(HERE, "losses_small.pt"))


def make_pip():
    """pip_small.npz: freezes oracle/geometry_ref.py on 40 jittered 12-gons (some closed by repeating the first vertex,
    overlapping) and 3000 float32 points."""
    from oracle.geometry_ref import points_in_polygons_ref
    rng = np.random.default_rng(77)
    rings = []
    for i in range(40):
        c = rng.uniform(5, 95, 2)
        ang = np.linspace(0, 2 * np.pi, 12, endpoint=False)
        r = 6.0 * (1 + 0.35 * rng.uniform(-1, 1, 12))
        ring = np.stack([c[0] + r * np.cos(ang), c[1] + r * np.sin(ang)], 1)
        rings.append(np.concatenate([ring, ring[:1]]) if i % 3 == 0 else ring)
    off = np.zeros(41, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in rings])
    verts = np.concatenate(rings)
    pts = rng.uniform(0, 100, (3000, 2)).astype(np.float32)
    pairs = points_in_polygons_ref(pts, verts, off)
    np.savez(os.path.join(HERE, "pip_small.npz"), points=pts, verts=verts, ring_off=off, pairs=pairs)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "losses":
        make_losses()
    elif len(sys.argv) > 1 and sys.argv[1] == "pip":
        make_pip()
    else:
        main()
