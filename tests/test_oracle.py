"""CPU tests of the oracle itself (no GPU).  The GATv2 restatement is unpinned by the reference
(no tests / golden vectors exist, PyG is not installable here), so it is held to self-consistency
properties: fp64 gradcheck, equivalence with a dense masked-attention formulation, softmax row
sums, permutation equivariance, and the committed golden vectors.  The kNN oracle is the
reference's own scipy call and is cross-checked by brute force."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import neighbors_ref, pyg_ref
from oracle.ist_encoder_ref import (ISTEncoderRef, Positional2dEmbedderRef, predict_scores_ref, scatter_max_ref,
                                    sinusoidal_embedding)
from tests.util import random_graph

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _case(n_src, n_dst, E, H, C, seed, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    x_l = torch.randn(n_src, H, C, generator=g, dtype=dtype)
    x_r = torch.randn(n_dst, H, C, generator=g, dtype=dtype)
    att = torch.randn(1, H, C, generator=g, dtype=dtype)
    bias = torch.randn(H * C, generator=g, dtype=dtype)
    return x_l, x_r, att, bias, random_graph(n_src, n_dst, E, seed)


def test_gatv2_gradcheck_fp64():
    x_l, x_r, att, bias, ei = _case(6, 5, 14, 2, 3, seed=0)
    for t in (x_l, x_r, att, bias):
        t.requires_grad_()
    assert torch.autograd.gradcheck(lambda a, b, c, d: pyg_ref.gatv2_aggregate(a, b, ei, c, d), (x_l, x_r, att, bias))


def test_gatv2_equals_dense_masked_attention():
    n_src, n_dst, H, C = 9, 7, 2, 4
    x_l, x_r, att, bias, _ = _case(n_src, n_dst, 1, H, C, seed=1)
    g = torch.Generator().manual_seed(2)
    adj = torch.rand(n_dst, n_src, generator=g) < 0.4
    adj[3] = False                                      # isolated destination
    dst, src = adj.nonzero(as_tuple=True)
    out = pyg_ref.gatv2_aggregate(x_l, x_r, torch.stack([src, dst]), att, bias)
    e = torch.nn.functional.leaky_relu(x_l[None] + x_r[:, None], 0.2)       # [n_dst, n_src, H, C]
    a = (e * att).sum(-1).masked_fill(~adj[..., None], float("-inf"))
    alpha = torch.softmax(a, dim=1).nan_to_num(0.0)
    dense = torch.einsum("ijh,jhc->ihc", alpha, x_l).reshape(n_dst, H * C) + bias
    assert torch.allclose(out, dense, atol=1e-12)
    assert torch.equal(out[3], bias)                    # isolated row = bias


def test_softmax_rows_sum_to_one_and_permutation_equivariance():
    x_l, x_r, att, bias, ei = _case(30, 20, 200, 3, 5, seed=3)
    out, alpha = pyg_ref.gatv2_aggregate(x_l, x_r, ei, att, bias, return_alpha=True)
    s = torch.zeros(20, 3, dtype=torch.float64).index_add_(0, ei[1], alpha)
    deg = torch.bincount(ei[1], minlength=20)
    assert torch.allclose(s[deg > 0], torch.ones_like(s[deg > 0]))
    assert torch.all(s[deg == 0] == 0)
    perm = torch.randperm(200, generator=torch.Generator().manual_seed(0))
    out2 = pyg_ref.gatv2_aggregate(x_l, x_r, ei[:, perm], att, bias)
    assert torch.allclose(out, out2, atol=1e-12)


def test_dropout_mask_injection_matches_manual():
    x_l, x_r, att, bias, ei = _case(12, 10, 60, 2, 4, seed=4)
    keep = torch.rand(60, 2, generator=torch.Generator().manual_seed(1)) > 0.2
    out = pyg_ref.gatv2_aggregate(x_l, x_r, ei, att, None, dropout_p=0.2, training=True, keep_mask=keep)
    _, alpha = pyg_ref.gatv2_aggregate(x_l, x_r, ei, att, None, return_alpha=True)
    w = alpha * keep / 0.8
    manual = torch.zeros(10, 2, 4, dtype=torch.float64).index_add_(0, ei[1], x_l[ei[0]] * w[..., None]).reshape(10, 8)
    assert torch.allclose(out, manual, atol=1e-12)


def test_scatter_max_semantics():
    src = torch.tensor([0.5, 0.9, 0.9, -1.0, 0.1])
    index = torch.tensor([2, 0, 0, 3, 2])
    out, arg = scatter_max_ref(src, index, 5)
    assert out.tolist() == pytest.approx([0.9, 0.0, 0.5, -1.0, 0.0])
    assert arg.tolist() == [1, 5, 0, 3, 5]               # ties -> first; empty -> (0, E)
    out, arg = scatter_max_ref(torch.zeros(0), torch.zeros(0, dtype=torch.long), 3)
    assert out.tolist() == [0, 0, 0] and arg.tolist() == [0, 0, 0]


def test_predict_scores_ref_unassigned_and_threshold():
    tx = torch.nn.functional.normalize(torch.randn(6, 8, generator=torch.Generator().manual_seed(0)), dim=-1)
    bd = torch.nn.functional.normalize(torch.randn(3, 8, generator=torch.Generator().manual_seed(1)), dim=-1)
    ei = torch.tensor([[0, 0, 2, 5], [0, 1, 2, 1]], dtype=torch.int32)
    seg, sim, idx = predict_scores_ref(tx, bd, ei, torch.tensor([10, 11, 12], dtype=torch.int32))
    assert seg[1] == -1 and seg[3] == -1 and seg[4] == -1 and idx[1] == 4
    assert seg[2] == 12 and seg[5] == 11
    seg2, _, _ = predict_scores_ref(tx, bd, ei, torch.tensor([10, 11, 12]), min_similarity=2.0)
    assert torch.all(seg2 == -1)


def test_positional_embedder_matches_reference_loop_semantics():
    torch.manual_seed(0)
    pe = Positional2dEmbedderRef(16)
    pos = torch.rand(50, 2) * 100
    batch = torch.randint(0, 3, (50,))
    out = pe(pos, batch)
    assert out.shape == (50, 16)
    # a tile's embedding depends only on that tile's own points
    m = batch == 1
    out1 = pe(pos[m], torch.zeros(int(m.sum()), dtype=torch.long))
    assert torch.allclose(out[m], out1, atol=1e-6)
    emb = sinusoidal_embedding(torch.tensor([0.0, 1.0]), 256, max_period=10000)
    assert emb.shape == (2, 256) and torch.all(emb[0, :128] == 1) and torch.all(emb[0, 128:] == 0)


def test_istencoder_ref_shapes_and_state_dict_keys():
    m = ISTEncoderRef(50, 12, in_channels=16, hidden_channels=8, out_channels=8, n_mid_layers=1, n_heads=2)
    keys = set(m.state_dict().keys())
    for k in ("lin_first.tx.weight", "lin_first.bd.bias", "pos_emb.mlp.0.weight", "pos_emb.mlp.2.bias",
              "conv_layers.0.conv.convs.<tx___neighbors___tx>.lin_l.weight",
              "conv_layers.2.conv.convs.<tx___belongs___bd>.att", "lin_last.lins.bd.weight"):
        assert k in keys, k
    assert m.conv_layers[0].conv.convs["<tx___neighbors___tx>"].lin_l.weight.shape == (16, 32)
    assert m.conv_layers[1].conv.convs["<tx___neighbors___tx>"].lin_l.weight.shape == (16, 16)


# ------------------------------------------------------------------------------------------- kNN
def test_knn_oracle_vs_bruteforce_and_scipy_properties():
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 60, (1500, 2)).astype(np.float32)
    canon, n_tie, n_bf, raw = neighbors_ref.canonical_knn_table(pts, 5, 5.0)
    brute = neighbors_ref.brute_force_knn_table(pts, 5, 5.0)
    assert np.array_equal(canon, brute)
    assert np.array_equal(canon[:, 0], np.arange(1500))            # self is neighbour 0 (Appendix A.5)
    d, i = neighbors_ref.kdtree_table(np.array([[0, 0], [5, 0], [0, 3]], dtype=np.float32), 3, 5.0)
    assert i.tolist() == [[0, 2, 3], [1, 3, 3], [2, 0, 3]]           # d == max_dist is excluded, pad = n
    ei, _ = neighbors_ref.kdtree_neighbors(pts, 5, 5.0, chunk_size=400)
    ei2, _ = neighbors_ref.kdtree_neighbors(pts, 5, 5.0)
    assert torch.equal(ei, ei2)                                    # chunking does not change the result


def test_knn_canonical_handles_ties():
    g = np.arange(12, dtype=np.float32)
    lattice = np.stack(np.meshgrid(g, g), -1).reshape(-1, 2)
    canon, n_tie, n_bf, _ = neighbors_ref.canonical_knn_table(lattice, 5, 1.5)
    assert n_tie > 0 and n_bf > 0
    assert np.array_equal(canon, neighbors_ref.brute_force_knn_table(lattice, 5, 1.5))


def test_knn_to_edge_index_restatement():
    t = torch.tensor([[0, 2, 4], [1, 4, 4], [4, 4, 4], [3, 0, 1]])
    ei, ip = neighbors_ref.knn_to_edge_index(t)
    assert ei.tolist() == [[0, 0, 1, 3, 3, 3], [0, 2, 1, 3, 0, 1]]
    assert ip.tolist() == [0, 2, 3, 3, 6]


# ------------------------------------------------------------------------------------------- golden
def test_oracle_reproduces_committed_golden_vectors():
    """tests/golden/*.pt were produced by tests/golden/make_golden.py from this oracle (the reference
    itself cannot run here); they freeze the oracle so that later edits cannot silently change it."""
    path = os.path.join(GOLD, "gatv2_small.pt")
    g = torch.load(path)
    out = pyg_ref.gatv2_aggregate(g["x_l"], g["x_r"], g["edge_index"], g["att"], g["bias"])
    assert torch.allclose(out, g["out"], atol=1e-6)
    enc = torch.load(os.path.join(GOLD, "encoder_small.pt"))
    m = ISTEncoderRef(**enc["hparams"])
    m.load_state_dict(enc["state_dict"])
    m.eval()
    o = m(enc["x"], enc["edges"], enc["pos"], enc["batch"])
    for k in ("tx", "bd"):
        assert torch.allclose(o[k], enc["out"][k], atol=1e-5)
    kn = np.load(os.path.join(GOLD, "knn_small.npz"))
    _, idx = neighbors_ref.kdtree_table(kn["points"], int(kn["k"]), float(kn["max_dist"]))
    assert np.array_equal(idx, kn["scipy_idx"])


def test_losses_oracle_reproduces_committed_golden_vectors():
    from oracle import triplet_loss_ref as R
    gd = torch.load(os.path.join(GOLD, "losses_small.pt"))
    pos, neg, dp, dn = R.FastTripletSelectorRef(gd["similarity"].clone()).sample_triplets(gd["labels"], gd["uniforms"])
    assert torch.equal(pos, gd["positives"]) and torch.equal(neg, gd["negatives"])
    assert torch.equal(dp, gd["dists_pos"]) and torch.equal(dn, gd["dists_neg"])
    e = gd["emb"].clone().requires_grad_()
    l_t = R.triplet_loss_ref(e, pos, neg, 0.3)
    l_m = R.metric_loss_ref(e, pos, neg, dp, dn)
    l_s = R.segmentation_loss_ref(e, gd["bd"], gd["edge_index"], gd["dst_neg"], "triplet", 0.4)
    l_b = R.segmentation_loss_ref(e, gd["bd"], gd["edge_index"], gd["dst_neg"], "bce", 0.4)
    (l_t + l_m + l_s + l_b).backward()
    for got, key in ((l_t, "loss_triplet"), (l_m, "loss_metric"), (l_s, "loss_seg_triplet"), (l_b, "loss_seg_bce")):
        assert torch.allclose(got.detach(), gd[key], atol=1e-6)
    assert torch.allclose(e.grad, gd["grad"], atol=1e-6)


# ---------------------------------------------------------------------------------------------- losses (N1)
def test_triplet_selector_ref_semantics():
    """FastTripletSelectorRef against a direct restatement of what triplet_loss.py:88-125 computes: cluster drawn from
    the row CDF of the anchor's cluster over *present* clusters, member = floor(u * size) into the cluster's block."""
    from oracle import triplet_loss_ref as R
    g = torch.Generator().manual_seed(0)
    C, N = 7, 400
    sim = torch.rand(C, C, generator=g) * 2 - 1
    sim = (sim + sim.t()) / 2
    labels = torch.randint(0, 5, (N,), generator=g)              # clusters 5, 6 absent
    uni = [torch.rand(N, generator=g) for _ in range(4)]
    sel = R.FastTripletSelectorRef(sim.clone())
    pos, neg, dp, dn = sel.sample_triplets(labels, uni)
    s = sim.clone(); s.fill_diagonal_(1)
    simc, dis = s.clamp_min(1e-8), (-s).clamp_min(1e-8)
    present = [c for c in range(C) if (labels == c).any()]
    members = {c: torch.nonzero(labels == c).flatten() for c in present}
    for i in range(0, N, 7):
        for u_c, u_m, w, got in ((uni[0], uni[1], simc, pos), (uni[2], uni[3], dis, neg)):
            row = w[labels[i]][present]
            cdf = torch.cumsum(row / row.sum(), 0); cdf[-1] = 1.0
            c = present[int(torch.searchsorted(cdf, u_c[i]))]
            m = members[c][int((u_m[i] * float(len(members[c]))).floor())]
            assert int(got[i]) == int(m)
    assert torch.equal(dp, 1 - simc[labels, labels[pos]]) and torch.equal(dn, 1 - simc[labels, labels[neg]])
    assert int(labels[pos].max()) < 5 and int(labels[neg].max()) < 5


def test_loss_refs_match_manual_formulas():
    from oracle import triplet_loss_ref as R
    g = torch.Generator().manual_seed(1)
    e = torch.randn(50, 8, generator=g, dtype=torch.float64)
    p, n = torch.randint(0, 50, (50,), generator=g), torch.randint(0, 50, (50,), generator=g)
    d = lambda x, y: ((x - y + 1e-6) ** 2).sum(1).sqrt()
    manual = (0.3 + d(e, e[p]) - d(e, e[n])).clamp_min(0).mean()
    assert torch.allclose(R.triplet_loss_ref(e, p, n, 0.3), manual, atol=1e-12)
    dp, dn = torch.rand(50, generator=g), torch.rand(50, generator=g)
    cos = lambda x, y: (x * y).sum(1) / (x.norm(dim=1) * y.norm(dim=1))
    manual = ((cos(e, e[p]) - (1 - dp)) ** 2).mean() + ((cos(e, e[n]) - (1 - dn)) ** 2).mean()
    assert torch.allclose(R.metric_loss_ref(e, p, n, dp, dn), manual, atol=1e-12)
    w = R.scheduled_weights_ref(torch.tensor([1., 1., 0.]), torch.tensor([1., 1., .5]), 0, 10)
    assert torch.allclose(w, torch.tensor([.5, .5, 0.]), atol=1e-6)
    w = R.scheduled_weights_ref(torch.tensor([1., 1., 0.]), torch.tensor([1., 1., .5]), 99, 10)
    assert torch.allclose(w, torch.tensor([.4, .4, .2]), atol=1e-6)


# ---------------------------------------------------------------------------------------------- geometry (N2)
def test_points_in_polygons_ref_known_answers():
    """Even-odd crossing rule on hand-checkable shapes: convex square, concave L, a ring closed by repeating its first
    vertex, degenerate rings; agreement with an independent winding-number evaluation on random interior points."""
    from oracle.geometry_ref import points_in_polygons_ref
    sq = np.array([[0.0, 0.0], [4.0, 0.0], [4.0, 4.0], [0.0, 4.0]])
    lshape = np.array([[10.0, 0.0], [16.0, 0.0], [16.0, 2.0], [12.0, 2.0], [12.0, 6.0], [10.0, 6.0]])
    closed = np.concatenate([sq + 20.0, (sq + 20.0)[:1]])
    rings = [sq, lshape, closed, np.zeros((0, 2)), np.array([[1.0, 1.0], [2.0, 2.0]])]
    off = np.zeros(len(rings) + 1, dtype=np.int64); off[1:] = np.cumsum([len(r) for r in rings])
    verts = np.concatenate(rings)
    pts = np.array([[2.0, 2.0], [11.0, 5.0], [15.0, 1.0], [14.0, 4.0], [22.0, 22.0], [100.0, 100.0], [-1.0, 2.0]])
    got = set(map(tuple, points_in_polygons_ref(pts, verts, off).T))
    assert got == {(0, 0), (1, 1), (2, 1), (4, 2)}
    # random points vs winding number (angle sum) for a star-shaped 16-gon
    rng = np.random.default_rng(0)
    ang = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    rad = 5.0 * (1 + 0.4 * rng.uniform(-1, 1, 16))
    ring = np.stack([rad * np.cos(ang), rad * np.sin(ang)], 1)
    p = rng.uniform(-8, 8, (4000, 2))
    a = ring[None] - p[:, None]
    b = np.roll(ring, -1, 0)[None] - p[:, None]
    wind = np.arctan2(a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0], (a * b).sum(-1)).sum(1) / (2 * np.pi)
    inside = np.abs(wind) > 0.5
    res = points_in_polygons_ref(p, ring, np.array([0, 16]))
    mask = np.zeros(4000, bool); mask[res[0]] = True
    assert np.array_equal(mask, inside)


def test_pip_oracle_reproduces_committed_golden_vectors():
    from oracle.geometry_ref import points_in_polygons_ref
    gd = np.load(os.path.join(GOLD, "pip_small.npz"))
    got = points_in_polygons_ref(gd["points"], gd["verts"], gd["ring_off"])
    assert np.array_equal(got, gd["pairs"]) and got.shape[1] > 500
