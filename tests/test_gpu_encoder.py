"""GPU parity tests, module level: the drop-in ISTEncoder / SkipGAT / LitISTEncoder.predict_step /
kdtree_neighbors against the CPU oracle on identical synthetic inputs and identical weights."""
import numpy as np
import pytest
import torch

from oracle import neighbors_ref
from oracle.ist_encoder_ref import TB, TT, predict_scores_ref
from segger_b200 import ops
from segger_b200.hetero import HeteroBatch
from segger_b200.lightning_model import LitISTEncoder
from segger_b200.neighbors import kdtree_neighbors, knn_table, knn_to_edge_index
from tests.util import PRED, make_models, rel_err, synth_batch, to_dev

pytestmark = pytest.mark.gpu
TOL = 1e-4

CONFIGS = [  # (in, hidden, out, n_mid, heads)
    (128, 64, 64, 0, 2),     # BASELINE cfg-1/2 model ("2-layer hetero GATv2 hidden=64 heads=2")
    (128, 64, 64, 2, 2),     # `segger segment` default (4 SkipGAT layers)
    (16, 32, 32, 1, 3),      # ISTEncoder's own defaults -> generic (F=96) kernels
    (64, 128, 128, 1, 4),    # cfg-4 shape (F=512)
]


from tests.util import RELAXED, check_forward_backward  # noqa: E402


@pytest.mark.parametrize("cfg", CONFIGS)
def test_istencoder_forward_backward_vs_oracle(cfg):
    in_c, hid, out_c, n_mid, heads = cfg
    ts, x, edges, pos, bat = synth_batch(6000, 60, seed=1)
    ref, prod = make_models(ts.n_genes, ts.bd_x.shape[1], in_c, hid, out_c, n_mid, heads, seed=3)
    check_forward_backward(ref, prod, x, edges, pos, bat, "r2_grad_parity_%d_%d_%d_%d_%d" % cfg)


def test_istencoder_train_mode_dropout_statistics_and_determinism():
    ts, x, edges, pos, bat = synth_batch(4000, 40, seed=2)
    _, prod = make_models(ts.n_genes, ts.bd_x.shape[1], 128, 64, 64, 0, 2, seed=1)
    prod.train()
    args = (to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat))
    torch.manual_seed(5); a = prod(*args)["tx"].detach()
    torch.manual_seed(5); b = prod(*args)["tx"].detach()
    torch.manual_seed(6); c = prod(*args)["tx"].detach()
    assert torch.equal(a, b)            # same torch seed -> same dropout realisation, bit-identical
    assert not torch.equal(a, c)
    prod.eval()
    e = prod(*args)["tx"].detach()
    assert float((a - e).abs().max()) > 0


def test_skipgat_standalone_generic_and_fused_agree():
    from segger_b200.ist_encoder import SkipGAT
    torch.manual_seed(0)
    ts, x, edges, pos, bat = synth_batch(3000, 30, seed=4)
    layer = SkipGAT((-1, -1), 64, 2).cuda().eval()
    xd = {"tx": torch.randn(3000, 96).cuda(), "bd": torch.randn(30, 96).cuda()}
    ed = to_dev({TT: edges[TT], TB: edges[TB]})
    fused = layer(xd, ed)
    generic = layer.conv(xd, ed)
    for k in ("tx", "bd"):
        assert rel_err(fused[k], generic[k]) < 1e-6
    # a caller that does supply bd-contains-tx edges gets the third conv through the generic path
    ed[("bd", "contains", "tx")] = ed[TB].flip(0)
    out = layer(xd, ed)
    assert out["tx"].shape == (3000, 128)
    aw = layer.attention_weights[TT]
    assert aw.shape == (ed[TT].size(1), 2)
    s = torch.zeros(3000, 2, device="cuda").index_add_(0, ed[TT][1], aw)
    deg = torch.bincount(ed[TT][1], minlength=3000)
    assert float((s[deg > 0] - 1).abs().max()) < 1e-5


def test_skipgat_source_subset_path_equals_full_projection_path():
    """tx-belongs-bd edges with unique increasing sources (what setup_heterodata emits) take the row-subset
    projection; the same edges shuffled, or with a repeated source, take the full [N, 3F] projection.  Same graph,
    same weights -> same outputs and gradients (fixed-order sums: only the edge order inside a bd row differs)."""
    from segger_b200.ist_encoder import SkipGAT
    torch.manual_seed(1)
    ts, x, edges, pos, bat = synth_batch(4000, 40, seed=5)
    layer = SkipGAT((-1, -1), 64, 2).cuda().eval()
    e_tt, e_tb = edges[TT].cuda(), edges[TB].cuda()
    assert bool((e_tb[0][1:] > e_tb[0][:-1]).all())
    perm = torch.randperm(e_tb.size(1), generator=torch.Generator().manual_seed(0)).cuda()
    g = torch.Generator().manual_seed(2)
    x_tx0, x_bd0 = torch.randn(4000, 96, generator=g).cuda(), torch.randn(40, 96, generator=g).cuda()
    go_tx, go_bd = torch.randn(4000, 128, generator=g).cuda(), torch.randn(40, 128, generator=g).cuda()
    res = []
    for tb in (e_tb, e_tb[:, perm].contiguous()):
        ops.CSR_CACHE.clear()
        layer.zero_grad()
        x_tx, x_bd = x_tx0.clone().requires_grad_(), x_bd0.clone().requires_grad_()
        out = layer({"tx": x_tx, "bd": x_bd}, {TT: e_tt, TB: tb})
        ((out["tx"] * go_tx).sum() + (out["bd"] * go_bd).sum()).backward()
        csr = ops.CSR_CACHE.get(tb, 4000, 40, True)
        res.append((csr.sources_unique_increasing(), out["tx"].detach(), out["bd"].detach(), x_tx.grad, x_bd.grad,
                    {n: p.grad.clone() for n, p in layer.named_parameters() if p.grad is not None}))
    assert res[0][0] is True and res[1][0] is False
    for a, b in zip(res[0][1:5], res[1][1:5]):
        assert rel_err(a, b) < 1e-5
    assert res[0][5].keys() == res[1][5].keys() and len(res[0][5]) == 12
    for n in res[0][5]:
        assert rel_err(res[0][5][n], res[1][5][n]) < 1e-5, n


def test_predict_step_assignment_agreement():
    ts, x, edges, pos, bat = synth_batch(20000, 200, seed=5, train_edges=False)
    torch.manual_seed(0)
    lit = LitISTEncoder(ts.n_genes, in_channels=128, n_mid_layers=0)
    ref, _ = make_models(ts.n_genes, ts.bd_x.shape[1], 128, 64, 64, 0, 2, seed=3, device="cpu")
    lit.model.load_state_dict(ref.state_dict(), strict=False)
    lit = lit.cuda().eval(); ref.eval()
    b = HeteroBatch()
    for k in ("tx", "bd"):
        b[k]["x"], b[k]["pos"], b[k]["batch"] = x[k], pos[k], bat[k]
    b["tx"]["index"] = torch.from_numpy(ts.tx_index)
    b["bd"]["index"] = torch.from_numpy(ts.bd_index) + 7
    mask = torch.rand(20000, generator=torch.Generator().manual_seed(1)) < 0.8
    b["tx"]["predict_mask"] = mask
    for et in (TT, TB, PRED):
        b[et]["edge_index"] = edges[et]
    with torch.no_grad():
        src_idx, seg_idx, max_sim, gen_idx = lit.predict_step(b.cuda(), 0)
        emb = ref(x, edges, pos, bat)
    seg_r, sim_r, _ = predict_scores_ref(emb["tx"], emb["bd"], edges[PRED], b["bd"]["index"])
    assert not src_idx.is_cuda and src_idx.dtype == torch.int64 and seg_idx.dtype == torch.int64
    assert torch.equal(src_idx, b["tx"]["index"][mask]) and torch.equal(gen_idx, x["tx"][mask])
    agree = float((seg_idx == seg_r[mask]).float().mean())
    assert agree >= 0.9999, agree
    assert rel_err(max_sim, sim_r[mask]) < TOL
    assert int((seg_idx >= 0).sum()) > 10000
    # all-true mask: no gather, same contract
    b["tx"]["predict_mask"] = torch.ones(20000, dtype=torch.bool)
    with torch.no_grad():
        s2, g2, m2, x2 = lit.predict_step(b.cuda(), 0)
    assert s2.shape == (20000,) and torch.equal(s2, b["tx"]["index"]) and torch.equal(g2[mask], seg_idx)
    assert torch.equal(m2[mask], max_sim) and torch.equal(x2, x["tx"])


# ---------------------------------------------------------------------------------------------- kNN
def _check_knn(points, k, r, query=None):
    table, count = knn_table(points, k, r, query=query)
    canon, n_tie, n_bf, raw = neighbors_ref.canonical_knn_table(points, k, r, query=query)
    got = table.cpu().numpy()
    assert got.shape == canon.shape
    assert np.array_equal(got, canon), f"{(got != canon).any(1).sum()} rows differ"
    assert np.array_equal(count.cpu().numpy(), (canon != points.shape[0]).sum(1))
    return n_tie, raw, got


@pytest.mark.parametrize("n,k,r,dtype", [(2000, 5, 5.0, np.float32), (50000, 5, 5.0, np.float32),
                                         (30000, 20, 5.0, np.float32), (20000, 3, 2.5, np.float64),
                                         (5000, 1 + 1, 50.0, np.float32), (10, 5, 5.0, np.float32)])
def test_knn_table_bit_exact_vs_scipy(n, k, r, dtype):
    rng = np.random.default_rng(n + k)
    side = (n / 0.5) ** 0.5
    pts = rng.uniform(0, side, (n, 2)).astype(dtype)
    n_tie, raw, got = _check_knn(pts, k, r)
    if n_tie == 0:
        assert np.array_equal(got, raw)     # tie-free: identical to scipy's own row order


def test_knn_ties_duplicates_and_lattice():
    g = np.arange(40, dtype=np.float32)
    lattice = np.stack(np.meshgrid(g, g), -1).reshape(-1, 2)            # massive exact ties
    n_tie, _, _ = _check_knn(lattice, 5, 1.5)
    assert n_tie > 0
    rng = np.random.default_rng(0)
    pts = rng.uniform(0, 30, (1500, 2)).astype(np.float32)
    pts[100:110] = pts[100]                                             # coincident points
    _check_knn(pts, 5, 5.0)
    # strict radius: a neighbour at exactly max_dist is excluded (Appendix A.5)
    ex = np.array([[0, 0], [5, 0], [0, 3], [100, 100]], dtype=np.float32)
    t, _ = knn_table(ex, 3, 5.0)
    assert t.cpu().tolist() == [[0, 2, 4], [1, 4, 4], [2, 0, 4], [3, 4, 4]]


def test_knn_separate_query_set_and_clustered():
    rng = np.random.default_rng(5)
    centers = rng.uniform(0, 500, (50, 2))
    pts = (centers[rng.integers(0, 50, 20000)] + rng.normal(0, 4, (20000, 2))).astype(np.float32)
    _check_knn(pts, 5, 5.0)
    qry = rng.uniform(-20, 520, (3000, 2)).astype(np.float32)           # some queries outside the bbox
    _check_knn(pts, 4, 5.0, query=qry)


def test_kdtree_neighbors_edge_list_matches_reference_call():
    ts, x, edges, pos, bat = synth_batch(30000, 300, seed=7, train_edges=False)
    ei, none = kdtree_neighbors(ts.tx_pos, 5, 5.0)
    assert none is None and ei.dtype == torch.int64 and not ei.is_cuda
    ref, _ = neighbors_ref.kdtree_neighbors(ts.tx_pos, 5, 5.0)
    assert ei.shape == ref.shape
    # same edge multiset; identical order wherever scipy's rows are tie-free
    key = lambda e: np.unique(e[0].numpy() * (1 << 32) + e[1].numpy())
    canon, n_tie, _, _ = neighbors_ref.canonical_knn_table(ts.tx_pos, 5, 5.0)
    ce, _ = neighbors_ref.knn_to_edge_index(torch.from_numpy(canon), padding_value=30000)
    assert torch.equal(ei, ce)
    if n_tie == 0:
        assert torch.equal(ei, ref)
    else:
        assert len(np.setxor1d(key(ei), key(ref))) <= 2 * n_tie
    # self loop first for every transcript (A.5), query-major order
    assert torch.equal(ei[0], torch.sort(ei[0], stable=True).values)


def test_knn_to_edge_index_generic_padding():
    t = torch.tensor([[1, 9, 2], [9, 9, 9], [0, 1, 9], [2, 9, 0]])
    ei, ip = knn_to_edge_index(t.cuda(), padding_value=9)
    er, ir = neighbors_ref.knn_to_edge_index(t, padding_value=9)
    assert torch.equal(ei.cpu(), er) and torch.equal(ip.cpu(), ir)
    ei2, _ = knn_to_edge_index(torch.full((4, 3), 4).cuda())          # default padding = N
    assert ei2.shape == (2, 0)
