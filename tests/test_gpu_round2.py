"""Round-2 GPU tests: Lightning-style inference mode, malformed edge lists, encoder-level dropout parity with the
kernels' own masks injected into the oracle, and parity at BASELINE.json's own sizes (configs[0] in full, the
configs[3] model on a 20k-transcript tile)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.ist_encoder_ref import TB, TT, predict_scores_ref
from segger_b200 import ops
from segger_b200.hetero import HeteroBatch
from segger_b200.lightning_model import LitISTEncoder
from tests.util import PRED, check_forward_backward, make_models, rel_err, synth_batch, to_dev

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _batch(ts, x, edges, pos, bat):
    b = HeteroBatch()
    for k in ("tx", "bd"):
        b[k]["x"], b[k]["pos"], b[k]["batch"] = x[k], pos[k], bat[k]
    b["tx"]["index"] = torch.from_numpy(ts.tx_index)
    b["bd"]["index"] = torch.from_numpy(ts.bd_index)
    b["tx"]["predict_mask"] = torch.ones(x["tx"].size(0), dtype=torch.bool)
    for et in (TT, TB, PRED):
        b[et]["edge_index"] = edges[et]
    return b


def test_forward_and_predict_step_under_inference_mode():
    """Lightning's predict / validation loops run under torch.inference_mode(): batches moved to the GPU inside it are
    inference tensors (no version counter).  Forward, predict_step and the generic conv path must all work."""
    ts, x, edges, pos, bat = synth_batch(3000, 30, seed=11, train_edges=False)
    torch.manual_seed(0)
    lit = LitISTEncoder(ts.n_genes, in_channels=32, n_mid_layers=0).cuda().eval()
    b = _batch(ts, x, edges, pos, bat)
    with torch.no_grad():
        want = lit.predict_step(b.cuda(), 0)
    ops.CSR_CACHE.clear()
    with torch.inference_mode():
        bi = b.cuda()
        assert bi[TT]["edge_index"].is_inference()
        got = lit.predict_step(bi, 0)
        emb = lit(bi)
        again = lit(bi)                                   # second call: CSR cache hit on an inference tensor
        generic = lit.model.conv_layers[0].conv(
            {"tx": torch.randn(3000, 64, device="cuda"), "bd": torch.randn(30, 64, device="cuda")},
            {TT: bi[TT]["edge_index"], TB: bi[TB]["edge_index"]})
    for g, w in zip(got, want):
        assert torch.equal(g, w)
    a = b = None
    assert torch.equal(emb["tx"], again["tx"]) and generic["tx"].shape == (3000, 128)
    # page-locked result memory is bounded: beyond the budget results come back as pageable copies, identical values
    import gc
    import segger_b200.lightning_model as L
    assert all(t.is_pinned() for t in got)
    live = L._pinned_live
    assert live >= sum(t.numel() * t.element_size() for t in got)
    old, L._PINNED_BUDGET = L._PINNED_BUDGET, 0
    try:
        with torch.inference_mode():
            paged = lit.predict_step(bi, 0)
    finally:
        L._PINNED_BUDGET = old
    assert all(not t.is_pinned() for t in paged) and all(torch.equal(a, b) for a, b in zip(paged, got))
    del got, want, paged, g, w, a, b
    gc.collect()
    with torch.inference_mode():
        extra = lit.predict_step(bi, 0)
    held = L._pinned_live
    del extra
    gc.collect()
    assert L._pinned_live < held                          # returned to the budget when the results are dropped


def test_malformed_edge_index_raises_index_error():
    """Out-of-range / negative node ids must raise like torch / PyG indexing does, not train on clamped edges."""
    ts, x, edges, pos, bat = synth_batch(2000, 20, seed=12)
    torch.manual_seed(0)
    _, prod = make_models(ts.n_genes, ts.bd_x.shape[1], 32, 64, 64, 0, 2, seed=0)
    prod.eval()
    args = lambda e: (to_dev(x), to_dev(e), to_dev(pos), to_dev(bat))
    prod(*args(edges))                                   # well-formed: fine
    for bad_val, et, row in ((-1, TT, 0), (2000, TT, 1), (20, TB, 1)):
        e = {k: v.clone() for k, v in edges.items()}
        e[et][row, 3] = bad_val
        ops.CSR_CACHE.clear()
        with pytest.raises(IndexError):
            prod(*args(e))
    # candidate edges of predict_step are checked too (folded into its result read-back)
    lit = LitISTEncoder(ts.n_genes, in_channels=32, n_mid_layers=0).cuda().eval()
    b = _batch(ts, x, edges, pos, bat)
    e = edges[PRED].clone(); e[1, 0] = 20
    b[PRED]["edge_index"] = e
    ops.CSR_CACHE.clear()
    with pytest.raises(IndexError), torch.no_grad():
        lit.predict_step(b.cuda(), 0)
    # a standalone CSR reports through validate()
    csr = ops.build_csr(torch.tensor([[0, 5], [1, 9]]).cuda(), 6, 9)
    with pytest.raises(IndexError):
        csr.validate()


def test_forward_is_sync_free_on_a_seen_batch():
    """A batch whose CSRs are validated and whose batch vectors are tagged runs without reading anything back:
    the forward can be captured in a CUDA graph."""
    ts, x, edges, pos, bat = synth_batch(4000, 40, seed=13)
    _, prod = make_models(ts.n_genes, ts.bd_x.shape[1], 32, 64, 64, 0, 2, seed=0)
    prod.eval()
    args = (to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat))
    with torch.no_grad():
        want = prod(*args)["tx"].clone()                  # first call: validates + tags (one read-back)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            prod(*args)
            torch.cuda.current_stream().synchronize()
            with torch.cuda.graph(g, stream=s):
                out = prod(*args)["tx"]
        g.replay()
        torch.cuda.synchronize()
    assert torch.equal(out, want)


def test_graphed_train_step_replays_equal_eager_steps_bit_for_bit():
    """segger_b200.graphs: a captured training step (forward + loss + backward + capturable Adam, dropout seeds from the
    device word) replayed twice leaves the parameters exactly where the same five eager steps leave them -- so the
    replays draw the dropout masks an eager run with the same seed word draws, and a new mask every replay."""
    import copy
    from segger_b200 import graphs
    ts, x, edges, pos, bat = synth_batch(6000, 60, seed=31)
    _, prod = make_models(ts.n_genes, ts.bd_x.shape[1], 32, 64, 64, 0, 2, seed=2)
    prod.train()
    twin = copy.deepcopy(prod)
    args = (to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat))
    gen = torch.Generator().manual_seed(3)
    t_tx = torch.randn(6000, 64, generator=gen).cuda()

    def make(model):
        opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True, fused=True)
        def fl():
            out = model(*args)
            return (out["tx"] * t_tx).sum() / 6000 + out["bd"].square().sum() / 60
        return opt, fl

    opt_a, fl_a = make(prod)
    prod(*args)                                         # resolve lazy layers and batch tags before the capture
    twin(*args)
    torch.manual_seed(5)
    step = graphs.graphed_train_step(fl_a, opt_a, warmup=3)
    wrap = lambda v: (v + (1 << 63)) % (1 << 64) - (1 << 63)          # int64 wrap-around, as the device add does
    word0 = wrap(int(step.seed_word.item()) - 3 * graphs._GOLDEN)    # 3 warm-up passes advanced it (a capture runs nothing)
    losses = [float(step.replay().item()), float(step.replay().item())]
    assert step.launches > 50
    assert losses[0] != losses[1]                                     # new dropout mask / new weights every replay
    # 3 warm-up + 2 replays = 5 steps, the word advanced once per step
    assert ((int(step.seed_word.item()) - word0) - 5 * graphs._GOLDEN) % (1 << 64) == 0

    opt_b, fl_b = make(twin)
    word = torch.tensor([word0], dtype=torch.int64).cuda()
    eager = []
    for i in range(5):
        word.add_(graphs._GOLDEN)
        with ops.device_seed(word):
            opt_b.zero_grad(set_to_none=True)
            loss = fl_b()
            loss.backward()
            opt_b.step()
            eager.append(float(loss.item()))
    assert eager[3:] == losses
    for (n, a), (_, b) in zip(prod.named_parameters(), twin.named_parameters()):
        if isinstance(a, torch.nn.parameter.UninitializedParameter) or "contains" in n:
            continue                                    # the dead bd-contains-tx conv: never materialised / never touched (B.1)
        assert torch.equal(a, b), n


def test_encoder_train_mode_parity_with_the_kernels_own_dropout_masks():
    """Train mode end to end: the masks the fused kernels regenerate (sgb_dropout_mask of the seeds drawn from torch's
    generator) are injected into the oracle layer by layer -> outputs and gradients within 1e-4."""
    in_c, hid, out_c, n_mid, heads = 128, 64, 64, 1, 2
    ts, x, edges, pos, bat = synth_batch(5000, 50, seed=21)
    ref, prod = make_models(ts.n_genes, ts.bd_x.shape[1], in_c, hid, out_c, n_mid, heads, seed=4)
    ref.train(); prod.train()
    n_layers = n_mid + 2
    torch.manual_seed(77)
    seeds = [(ops.new_seed(), ops.new_seed()) for _ in range(n_layers)]      # the order forward_fused draws them in
    E_tt, E_tb = edges[TT].size(1), edges[TB].size(1)
    keep = [{TT: ops.dropout_keep_mask(s_tt, E_tt, heads, 0.2, "cuda").cpu(),
             TB: ops.dropout_keep_mask(s_tb, E_tb, heads, 0.2, "cuda").cpu()} for s_tt, s_tb in seeds]
    assert 0.15 < 1 - float(keep[0][TT].float().mean()) < 0.25
    out_r = ref(x, edges, pos, bat, keep_masks=keep)
    gen = torch.Generator().manual_seed(0)
    g = {k: torch.randn(v.shape, generator=gen) for k, v in out_r.items()}
    sum((out_r[k] * g[k]).sum() for k in g).backward()
    torch.manual_seed(77)
    out_p = prod(to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat))
    sum((out_p[k] * g[k].cuda()).sum() for k in g).backward()
    for k in ("tx", "bd"):
        assert rel_err(out_p[k], out_r[k]) < TOL, k
    rg = dict(ref.named_parameters())
    errs = {n: rel_err(p.grad, rg[n].grad) for n, p in prod.named_parameters() if p.grad is not None}
    bad = {n: e for n, e in errs.items() if e >= TOL}
    assert not bad, bad


def _report(name, payload):
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", name), "w") as f:
        json.dump(payload, f, indent=1)


def test_config0_full_size_forward_backward_assignment_vs_oracle():
    """BASELINE.json configs[0] in full: 50k transcripts / 500 cells, k=5, 2-layer hetero GATv2 hidden=64 heads=2 --
    kNN graph built by the product, forward + backward + transcript->cell assignment against the CPU oracle."""
    from oracle import neighbors_ref
    from segger_b200.neighbors import kdtree_neighbors
    ts, x, edges, pos, bat = synth_batch(50_000, 500, seed=0, train_edges=False)
    ei, _ = kdtree_neighbors(ts.tx_pos, 5, 5.0)
    canon, _, _, _ = neighbors_ref.canonical_knn_table(ts.tx_pos, 5, 5.0)
    ce, _ = neighbors_ref.knn_to_edge_index(torch.from_numpy(canon), padding_value=50_000)
    assert torch.equal(ei, ce)                                            # graph edge list bit-exact
    edges = dict(edges); edges[TT] = ei
    ref, prod = make_models(ts.n_genes, ts.bd_x.shape[1], 128, 64, 64, 0, 2, seed=0)
    out_r, out_p = check_forward_backward(ref, prod, x, edges, pos, bat, "r2_grad_parity_cfg0_50k", grad_scale=1e-3,
                                          noise_trials=2)
    _, _, seg = ops.score_argmax(out_p["tx"].detach(), out_p["bd"].detach(), edges[PRED].cuda(),
                                 torch.from_numpy(ts.bd_index).cuda())
    seg_r, _, _ = predict_scores_ref(out_r["tx"].detach(), out_r["bd"].detach(), edges[PRED], torch.from_numpy(ts.bd_index))
    agree = float((seg.cpu() == seg_r).float().mean())
    assert agree >= 0.9999, agree


def test_config3_model_on_a_20k_tile_vs_oracle():
    """The configs[3] model (in=128, hidden=128, heads=4, 3 layers, k=20 neighbours) on a 20k-transcript tile."""
    ts, x, edges, pos, bat = synth_batch(20_000, 200, seed=3, k=20)
    ref, prod = make_models(ts.n_genes, ts.bd_x.shape[1], 128, 128, 128, 1, 4, seed=5)
    check_forward_backward(ref, prod, x, edges, pos, bat, "r2_grad_parity_cfg3_20k", grad_scale=1e-3, noise_trials=2)


def test_factored_first_layer_equals_dense_form(monkeypatch):
    """The gene-embedding half of the first layer consumed as (ids, table) -- a table lookup in the GEMM epilogue and a
    segment sum in the backward -- against the same model run with the concatenated [N, 2*in] input
    (SEGGER_B200_FACTOR=0): same math, different summation order."""
    ts, x, edges, pos, bat = synth_batch(6000, 60, seed=31)
    ref, prod = make_models(ts.n_genes, ts.bd_x.shape[1], 128, 64, 64, 0, 2, seed=2)
    prod.eval()
    args = (to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat))
    gen = torch.Generator().manual_seed(0)
    g = {"tx": torch.randn(6000, 64, generator=gen).cuda(), "bd": torch.randn(60, 64, generator=gen).cuda()}
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("SEGGER_B200_FACTOR", flag)
        prod.zero_grad()
        out = prod(*args)
        sum((out[k] * g[k]).sum() for k in g).backward()
        res[flag] = ({k: v.detach().clone() for k, v in out.items()},
                     {n: p.grad.clone() for n, p in prod.named_parameters() if p.grad is not None})
        res[flag] += (prod.conv_layers[0].attention_weights[TT].clone(),)     # first layer: factored input rebuilt
    for k in ("tx", "bd"):
        assert rel_err(res["1"][0][k], res["0"][0][k]) < 1e-5, k
    assert res["0"][1].keys() == res["1"][1].keys()
    assert float((res["1"][2] - res["0"][2]).abs().max()) < 1e-5
    # both forms are fp32 roundings of the same (partly ill-conditioned) gradients: measured against the fp64 oracle the
    # factored form must be as close as the dense one
    import copy
    r64 = copy.deepcopy(ref).double().eval()
    o64 = r64({"tx": x["tx"], "bd": x["bd"].double()}, edges, {k: v.double() for k, v in pos.items()}, bat)
    sum((o64[k] * g[k].cpu().double()).sum() for k in g).backward()
    for n, p in r64.named_parameters():
        e_dense, e_fact = rel_err(res["0"][1][n], p.grad), rel_err(res["1"][1][n], p.grad)
        assert e_fact <= 1.5 * e_dense + 1e-5, (n, e_fact, e_dense)
    # frozen pretrained embedding: no table gradient is computed or returned
    prod.lin_first["tx"].weight.requires_grad_(False)
    prod.zero_grad()
    out = prod(*args)
    sum((out[k] * g[k]).sum() for k in g).backward()
    assert prod.lin_first["tx"].weight.grad is None
    assert rel_err(prod.pos_emb.mlp[0].weight.grad, res["1"][1]["pos_emb.mlp.0.weight"]) < 1e-6


def test_side_stream_csr_builds_equal_inline_builds(monkeypatch):
    """Graphs of >= 2M edges get their CSRs built on a side stream while the input stage runs (ops.csr_build_overlapped).
    Forced here on a small graph: forward, gradients and the IndexError of a malformed edge list must equal the inline
    path bit for bit, under autograd and under inference_mode."""
    ts, x, edges, pos, bat = synth_batch(5000, 50, seed=21)
    _, prod = make_models(ts.n_genes, ts.bd_x.shape[1], 32, 64, 64, 0, 2, seed=0)
    prod.eval()                                            # (dropout off: the two runs must be comparable)
    args = (to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat))
    res = {}
    for name, min_edges in (("inline", 1 << 60), ("side", 0)):
        monkeypatch.setattr(ops, "_CSR_OVERLAP_MIN_EDGES", min_edges)
        ops.CSR_CACHE.clear()
        prod.zero_grad(set_to_none=True)
        out = prod(*args)
        (out["tx"].square().sum() + out["bd"].sum()).backward()
        torch.cuda.synchronize()
        res[name] = (out["tx"].detach().clone(), out["bd"].detach().clone(),
                     {n: p.grad.clone() for n, p in prod.named_parameters() if p.grad is not None})
    assert torch.equal(res["inline"][0], res["side"][0]) and torch.equal(res["inline"][1], res["side"][1])
    assert res["inline"][2].keys() == res["side"][2].keys() and len(res["side"][2]) > 10
    assert all(torch.equal(res["inline"][2][n], res["side"][2][n]) for n in res["side"][2])
    # repeated new graphs on the side stream (allocator reuse across streams), then a malformed one
    for _ in range(3):
        ops.CSR_CACHE.clear()
        with torch.inference_mode():
            again = prod(*(to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat)))["tx"]
        assert torch.equal(again, res["side"][0])
    bad = {k: v.clone() for k, v in edges.items()}
    bad[TT][1, 7] = 5000
    ops.CSR_CACHE.clear()
    with pytest.raises(IndexError), torch.no_grad():
        prod(to_dev(x), to_dev(bad), to_dev(pos), to_dev(bat))


def test_loss_bookkeeping_on_the_side_stream_equals_inline(monkeypatch):
    """LitISTEncoder.get_losses resolves masks / labels and samples the triplets on a side stream while the forward
    runs: same draws, same losses, same gradients as the inline order (SEGGER_B200_LOSS_OVERLAP=0)."""
    ts, x, edges, pos, bat = synth_batch(4000, 40, seed=5)
    g = torch.Generator().manual_seed(6)

    def sim(c):
        a = torch.rand(c, c, generator=g) * 2 - 1
        return ((a + a.t()) / 2).contiguous()

    s_tx, s_bd = sim(8), sim(4)
    b = HeteroBatch()
    for k in ("tx", "bd"):
        b[k]["x"], b[k]["pos"], b[k]["batch"] = x[k], pos[k], bat[k]
    b["tx"]["mask"] = torch.rand(4000, generator=g) < 0.8
    b["tx"]["cluster"] = torch.randint(0, 8, (4000,), generator=g)
    b["bd"]["mask"] = torch.rand(40, generator=g) < 0.9
    b["bd"]["cluster"] = torch.randint(-1, 4, (40,), generator=g)
    b[TT]["edge_index"], b[TB]["edge_index"] = edges[TT], edges[TB]
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("SEGGER_B200_LOSS_OVERLAP", flag)
        torch.manual_seed(0)
        lit = LitISTEncoder(ts.n_genes, in_channels=32, hidden_channels=32, out_channels=32, n_mid_layers=0).cuda().eval()
        lit.setup_losses(s_tx.clone(), s_bd.clone())
        lit.set_epoch(2, 10)
        bc = b.cuda()
        with torch.no_grad():
            lit.forward(bc)                                 # lazy parameters
        torch.manual_seed(33)
        ops.CSR_CACHE.clear()
        losses = lit.get_losses(bc)
        losses[3].backward()
        torch.cuda.synchronize()
        res[flag] = ([float(v.detach()) if torch.is_tensor(v) else float(v) for v in losses],
                     {n: p.grad.clone() for n, p in lit.model.named_parameters() if p.grad is not None})
    assert res["0"][0] == res["1"][0]
    assert res["0"][1].keys() == res["1"][1].keys()
    assert all(torch.equal(res["0"][1][n], res["1"][1][n]) for n in res["0"][1])
