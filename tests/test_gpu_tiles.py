"""GPU parity of the device-side tiling against oracle/tiles_ref.py -- index work, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import tiles_ref
from segger_b200 import ops, tiles
from segger_b200.hetero import HeteroBatch
from segger_b200.lightning_model import LitISTEncoder
from tests.util import PRED, TB, TT, synth_batch

pytestmark = pytest.mark.gpu


def _graph(n_tx=12000, n_cells=120, seed=0):
    ts, x, edges, pos, bat = synth_batch(n_tx, n_cells, seed=seed, train_edges=False)
    nodes = {
        "tx": {"x": x["tx"], "pos": pos["tx"], "index": torch.from_numpy(ts.tx_index),
               "flag": torch.from_numpy(ts.tx_compartment > 0)},
        "bd": {"x": x["bd"], "pos": pos["bd"], "index": torch.from_numpy(ts.bd_index)},
    }
    b = HeteroBatch()
    for nt, st in nodes.items():
        for k, v in st.items():
            b[nt][k] = v
    for et in (TT, TB, PRED):
        b[et]["edge_index"] = edges[et]
    return ts, nodes, {et: edges[et] for et in (TT, TB, PRED)}, b


def _labels(ts, nodes, nx=3, ny=2):
    lab = {}
    for nt in ("tx", "bd"):
        p = nodes[nt]["pos"]
        ix = torch.clamp((p[:, 0] / (ts.side / nx)).long(), 0, nx - 1)
        iy = torch.clamp((p[:, 1] / (ts.side / ny)).long(), 0, ny - 1)
        lab[nt] = iy * nx + ix
    return lab, nx * ny


def test_tile_partition_matches_oracle():
    ts, nodes, edges, b = _graph()
    lab, P = _labels(ts, nodes)
    part = tiles.TilePartition(b.cuda(), {k: v.cuda() for k, v in lab.items()}, P)
    rn, re, indptr, e_indptr, perm = tiles_ref.partition_ref(nodes, edges, lab, P)
    for nt in ("tx", "bd"):
        assert part.node_indptr[nt] == indptr[nt].tolist()
        assert torch.equal(part.node_perm[nt].cpu().long(), perm[nt])
        for k, v in rn[nt].items():
            assert torch.equal(part.data[nt][k].cpu(), v), (nt, k)
    for et in (TT, TB, PRED):
        assert part.edge_indptr[et] == e_indptr[et].tolist(), et
        got = part.data[et]["edge_index"]
        assert got.dtype == edges[et].dtype and torch.equal(got.cpu().long(), re[et]), et
    # cross-tile edges were dropped, none invented
    assert part.data[TT]["edge_index"].size(1) < edges[TT].size(1)
    # tile slices and batches
    for i in (0, P - 1, 2):
        n_ref, e_ref = tiles_ref.get_tile_ref(rn, re, indptr, e_indptr, i)
        t = part[i]
        for nt in ("tx", "bd"):
            for k, v in n_ref[nt].items():
                assert torch.equal(t[nt][k].cpu(), v), (i, nt, k)
        for et in (TT, TB, PRED):
            assert torch.equal(t[et]["edge_index"].cpu().long(), e_ref[et]), (i, et)
    order = [4, 1, 5]
    cn, ce = tiles_ref.collate_ref([tiles_ref.get_tile_ref(rn, re, indptr, e_indptr, i) for i in order])
    batch = part.collate(order)
    assert batch.num_graphs == 3
    for nt in ("tx", "bd"):
        for k, v in cn[nt].items():
            assert torch.equal(batch[nt][k].cpu(), v), (nt, k)
    for et in (TT, TB, PRED):
        assert torch.equal(batch[et]["edge_index"].cpu().long(), ce[et]), et
    # packing the tiles into batches of <= max edges: every tile once, no batch over the limit
    w = part.weights("edge")
    bs = part.batches(max_num=max(w) * 2)
    assert sorted(i for bb in bs for i in bb) == list(range(P)) and all(sum(w[i] for i in bb) <= max(w) * 2 for bb in bs)
    with pytest.raises(IndexError):
        tiles.TilePartition(b.cuda(), {"tx": lab["tx"].cuda() + 1, "bd": lab["bd"].cuda()}, P)


def test_tile_predict_subset_matches_oracle_incl_boundaries_and_dtypes():
    ts, nodes, edges, b = _graph(seed=3)
    boxes = tiles.square_tiles(0.0, 0.0, ts.side, ts.side, 3, 3)
    # put transcripts exactly on inner / outer box borders: closed inner box, half-open outer box
    x0, y0, x1, y1 = boxes[4]
    nodes["tx"]["pos"][0] = torch.tensor([x0, y0]); nodes["tx"]["pos"][1] = torch.tensor([x1, y1])
    nodes["tx"]["pos"][2] = torch.tensor([x0 - 20.0, y0]); nodes["tx"]["pos"][3] = torch.tensor([x1 + 20.0, y1])
    b["tx"]["pos"] = nodes["tx"]["pos"]
    ds = tiles.TilePredictSet(b.cuda(), boxes, margin=20.0)
    assert len(ds) == 9
    for i in (4, 0, 8):
        n_ref, e_ref, kept = tiles_ref.subset_ref(nodes, edges, boxes[i], 20.0)
        t = ds[i]
        for nt in ("tx", "bd"):
            for k, v in n_ref[nt].items():
                assert torch.equal(t[nt][k].cpu(), v), (i, nt, k)
        for et in (TT, TB, PRED):
            got = t[et]["edge_index"]
            assert got.dtype == edges[et].dtype and torch.equal(got.cpu().long(), e_ref[et]), (i, et)
    with pytest.raises(IndexError):
        ds[9]
    # the grid-indexed path (scan restricted to the 3 x 3 neighbourhood of cells) returns the identical subgraph
    dsi = tiles.TilePredictSet(b.cuda(), boxes, margin=20.0, grid=(3, 3))
    assert dsi._index is not None
    for i in range(9):
        a, c = ds.subset(boxes[i]), dsi[i]
        for nt in ("tx", "bd"):
            assert set(a[nt]) == set(c[nt])
            for k in a[nt]:
                assert torch.equal(a[nt][k], c[nt][k]), (i, nt, k)
        for et in (TT, TB, PRED):
            assert torch.equal(a[et]["edge_index"], c[et]["edge_index"]), (i, et)
    assert all(int((m != -1).sum()) == 0 for m in dsi._index["maps"].values())      # persistent maps left clean
    assert tiles.TilePredictSet(b.cuda(), boxes, margin=1e6, grid=(3, 3))._index is None   # halo wider than a tile: full scan
    # float64 positions compare in float64
    b64 = HeteroBatch()
    for nt in ("tx", "bd"):
        for k, v in nodes[nt].items():
            b64[nt][k] = v.double() if k == "pos" else v
    for et in (TT, TB, PRED):
        b64[et]["edge_index"] = edges[et]
    n64 = {nt: {k: (v.double() if k == "pos" else v) for k, v in st.items()} for nt, st in nodes.items()}
    n_ref, e_ref, _ = tiles_ref.subset_ref(n64, edges, boxes[4], 7.5)
    t = tiles.TilePredictSet(b64.cuda(), boxes, margin=7.5)[4]
    assert torch.equal(t["tx"]["index"].cpu(), n_ref["tx"]["index"]) and torch.equal(t["tx"]["predict_mask"].cpu(), n_ref["tx"]["predict_mask"])


def test_batched_tile_cut_equals_collated_per_tile_subsets():
    """TilePredictSet.cut (all tiles of a group in two flag kernels per store) == collate of the per-tile subsets,
    element for element: node order, attributes, predict_mask, batch vectors, renumbered edges in original order.
    slice_slots / collate_tiles re-cut a group into batches."""
    for pos_dtype, e_dtype in ((torch.float32, torch.int64), (torch.float64, torch.int32)):
        ts, nodes, edges, b = _graph(n_tx=40000, n_cells=400, seed=9)
        for nt in ("tx", "bd"):
            b[nt]["pos"] = nodes[nt]["pos"].to(pos_dtype)
        for et in (TT, TB, PRED):
            b[et]["edge_index"] = edges[et].to(e_dtype)
        b[TT]["weight"] = torch.arange(edges[TT].size(1), dtype=torch.float32)       # an edge attribute rides along
        lo = nodes["tx"]["pos"].min(0).values - 1e-3
        hi = nodes["tx"]["pos"].max(0).values + 1e-3
        boxes = tiles.square_tiles(float(lo[0]), float(lo[1]), float(hi[0]), float(hi[1]), 4, 4)
        ds = tiles.TilePredictSet(b.cuda(), boxes, margin=15.0, grid=(4, 4))
        assert ds._index is not None
        ids = [5, 0, 15, 6, 10]
        want = tiles.collate_tiles([ds[t] for t in ids])
        got, info = ds.cut(ids)
        assert got.num_graphs == want.num_graphs == 5
        for nt in ("tx", "bd"):
            assert set(got[nt]) == set(want[nt])
            for k in want[nt]:
                assert got[nt][k].dtype == want[nt][k].dtype and torch.equal(got[nt][k], want[nt][k]), (nt, k)
            assert info["nodes"][nt] == [int(ds[t][nt]["pos"].size(0)) for t in ids]
        for et in (TT, TB, PRED):
            assert got[et]["edge_index"].dtype == e_dtype
            assert torch.equal(got[et]["edge_index"], want[et]["edge_index"]), et
            assert info["edges"][et] == [int(ds[t][et]["edge_index"].size(1)) for t in ids]
        assert torch.equal(got[TT]["weight"], want[TT]["weight"])
        # slots [1, 4) of the group == the collate of those three tiles; two slices collated back == the group
        mid = tiles.slice_slots(got, info, 1, 4)
        want_mid = tiles.collate_tiles([ds[t] for t in ids[1:4]])
        back = tiles.collate_tiles([tiles.slice_slots(got, info, 0, 2), tiles.slice_slots(got, info, 2, 5)])
        for a, c in ((mid, want_mid), (back, got)):
            assert a.num_graphs == c.num_graphs
            for nt in ("tx", "bd"):
                for k in c[nt]:
                    assert torch.equal(a[nt][k], c[nt][k]), (nt, k)
            for et in (TT, TB, PRED):
                assert torch.equal(a[et]["edge_index"], c[et]["edge_index"]), et
    # a group runs through predict_step like any collated batch
    torch.manual_seed(0)
    lit = LitISTEncoder(ts.n_genes, in_channels=32, n_mid_layers=0).cuda().eval()
    with torch.no_grad():
        r_got = lit.predict_step(got, 0)
        r_want = lit.predict_step(want, 0)
    for x, y in zip(r_got, r_want):
        assert torch.equal(x, y)


def test_tiled_prediction_with_halo_equals_whole_graph_prediction():
    """Tiles + 20 um halo are independent units (receptive field = n_layers x 5 um): predicting tile by tile and keeping
    the inner-tile rows gives exactly the assignments of one pass over the whole graph."""
    ts, nodes, edges, b = _graph(n_tx=30000, n_cells=300, seed=5)
    torch.manual_seed(0)
    lit = LitISTEncoder(ts.n_genes, in_channels=32, n_mid_layers=0).cuda().eval()
    b["tx"]["predict_mask"] = torch.ones(30000, dtype=torch.bool)
    # one tile for the positional normalisation: every tile must see the same normalised coordinates -> no pos-emb
    lit.model.use_positional_embeddings = False
    for nt in ("tx", "bd"):
        b[nt]["batch"] = torch.zeros(nodes[nt]["pos"].size(0), dtype=torch.long)
    full = b.cuda()
    with torch.no_grad():
        src, seg, sim, gene = lit.predict_step(full, 0)
    whole = {int(i): (int(s), float(v)) for i, s, v in zip(src, seg, sim)}
    lo = nodes["tx"]["pos"].min(0).values - 1e-3      # cells at the rim stick out of [0, side]
    hi = nodes["tx"]["pos"].max(0).values + 1e-3
    boxes = tiles.square_tiles(float(lo[0]), float(lo[1]), float(hi[0]), float(hi[1]), 3, 3)
    ds = tiles.TilePredictSet(full, boxes, margin=20.0)
    seen = {}
    for i in range(len(ds)):
        with torch.no_grad():
            s_i, g_i, v_i, _ = lit.predict_step(ds[i], 0)
        for a, c, v in zip(s_i.tolist(), g_i.tolist(), v_i.tolist()):
            if a not in seen or v > seen[a][1]:
                seen[a] = (c, v)
    assert set(seen) == set(whole)
    agree = np.mean([seen[k][0] == whole[k][0] for k in whole])
    assert agree >= 0.9999, agree
    assert max(abs(seen[k][1] - whole[k][1]) for k in whole) < 1e-5


def test_collated_training_batch_runs_sync_free_and_matches_per_tile_losses():
    """A batch assembled by TilePartition.collate carries tagged batch vectors: forward/backward over it equals the sum
    over its tiles (tiles share no edges) and needs no read-back for the batch metadata."""
    ts, nodes, edges, b = _graph(n_tx=9000, n_cells=90, seed=7)
    lab, P = _labels(ts, nodes)
    part = tiles.TilePartition(b.cuda(), {k: v.cuda() for k, v in lab.items()}, P)
    torch.manual_seed(0)
    from segger_b200.ist_encoder import ISTEncoder, _known_num_graphs
    model = ISTEncoder(ts.n_genes, 32, 64, 64, 0, 2).cuda().eval()
    batch = part.collate([1, 3, 4])
    assert _known_num_graphs(batch["tx"]["batch"]) == 3 and _known_num_graphs(batch["bd"]["batch"]) == 3
    out = model(batch.x_dict, {TT: batch[TT]["edge_index"], TB: batch[TB]["edge_index"]}, batch.pos_dict, batch.batch_dict)
    parts = []
    for t in (1, 3, 4):
        one = part[t]
        parts.append(model(one.x_dict, {TT: one[TT]["edge_index"], TB: one[TB]["edge_index"]}, one.pos_dict, one.batch_dict))
    for k in ("tx", "bd"):
        ref = torch.cat([p[k] for p in parts])
        assert float((out[k] - ref).abs().max()) < 1e-5, k
