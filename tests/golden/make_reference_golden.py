"""Golden fixtures produced by the REFERENCE'S OWN CODE (dpeerlab/segger, /root/reference), run in the
build container through oracle/reference_import.py (the reference's files executed unmodified; only the
third-party classes they import are stand-ins, see oracle/pyg_stub.py).

    python tests/golden/make_reference_golden.py          # writes tests/golden/ref_*.pt / ref_*.npz

The fixtures travel to the GPU box (the reference tree does not).  They pin:
  ref_posemb.pt    sinusoidal_embedding / Positional2dEmbedder            models/ist_encoder.py:22-79     (a4)
  ref_encoder_*.pt ISTEncoder.__init__/forward (+ autograd gradients)     models/ist_encoder.py:214-333   (a5-a9, a11)
  ref_losses.pt    FastTripletSelector / TripletLoss / MetricLoss         models/triplet_loss.py:8-204    (N1)
  ref_lit.pt       LitISTEncoder.predict_step / get_losses / _scheduled_weights  models/lightning_model.py:136-298 (a10, N1)
  ref_knn.npz      kdtree_neighbors / knn_to_edge_index / setup_transcripts_graph  data/utils/neighbors.py:54-180 (a1-a3)
Random draws the reference makes (torch.rand / torch.randint) are recorded into the fixture so that the
product can replay the very same numbers.
"""
from __future__ import annotations

import contextlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import reference_import  # noqa: E402
from segger_b200.hetero import HeteroBatch  # noqa: E402  (plain container, no kernels)

TT = ("tx", "neighbors", "tx")
TB = ("tx", "belongs", "bd")
PRED = ("tx", "neighbors", "bd")


@contextlib.contextmanager
def record_random(log: list):
    """Record every torch.rand / torch.randint result drawn inside the block (in call order)."""
    rand0, randint0 = torch.rand, torch.randint

    def rand(*a, **k):
        t = rand0(*a, **k)
        log.append(("rand", t.clone()))
        return t

    def randint(*a, **k):
        t = randint0(*a, **k)
        log.append(("randint", t.clone()))
        return t

    torch.rand, torch.randint = rand, randint
    try:
        yield log
    finally:
        torch.rand, torch.randint = rand0, randint0


def small_graph(N, M, n_genes, bd_in, gen, k=5, dist=8.0, side=50.0, tiles=2):
    ref = reference_import.load()
    x = {"tx": torch.randint(0, n_genes, (N,), generator=gen, dtype=torch.int32),
         "bd": torch.randn(M, bd_in, generator=gen)}
    pos = {"tx": torch.rand(N, 2, generator=gen) * side, "bd": torch.rand(M, 2, generator=gen) * side}
    per = -(-N // tiles)
    batch = {"tx": torch.arange(N) // per, "bd": (torch.arange(M) * tiles) // M}
    ett, _ = ref.neighbors.kdtree_neighbors(pos["tx"].numpy(), k, dist)          # the reference's own graph builder
    src = torch.arange(0, N, 3)
    etb = torch.stack([src, src % M])
    # candidate edges: every transcript gets 0-3 candidate boundaries, int32 as the reference emits them
    cand = []
    for t in range(N):
        for c in range(int(t % 4)):
            cand.append((t, (t * 7 + c * 3) % M))
    epred = torch.tensor(cand, dtype=torch.int32).t().contiguous()
    return x, pos, batch, {TT: ett, TB: etb, PRED: epred}


def state_of(model):
    """Materialised entries of a state_dict (the dead bd-contains-tx conv stays uninitialised, Appendix B.1)."""
    from torch.nn.parameter import UninitializedParameter
    sd, lazy = {}, []
    for k, v in model.state_dict().items():
        if isinstance(v, UninitializedParameter):
            lazy.append(k)
        elif "bd___contains___tx" in k:
            continue            # att / biases of the dead conv: allocated but never initialised (garbage memory)
        else:
            sd[k] = v.detach().clone()
    return sd, lazy


def make_posemb():
    ref = reference_import.load().ist_encoder
    g = torch.Generator().manual_seed(11)
    torch.manual_seed(11)
    emb = ref.Positional2dEmbedder(32)
    pos = torch.rand(70, 2, generator=g) * 300 + 1000
    batch = torch.arange(70) // 24
    x = torch.rand(40, generator=g)
    out = dict(
        mlp_state={k: v.clone() for k, v in emb.state_dict().items()},
        pos=pos, batch=batch,
        out_batched=emb(pos, batch).detach(), out_global=emb(pos, None).detach(),
        sin_x=x, sin_256=ref.sinusoidal_embedding(x, 256, 10000), sin_7=ref.sinusoidal_embedding(x, 7, 1000),
        embed_2d=ref.Positional2dEmbedder.embed(torch.rand(5, 2, generator=g), 256),
    )
    torch.save(out, os.path.join(HERE, "ref_posemb.pt"))


def make_encoder(tag, hp, N, M, bd_in, seed):
    ref = reference_import.load().ist_encoder
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    model = ref.ISTEncoder(**hp).eval()
    x, pos, batch, edges = small_graph(N, M, hp["n_genes"], bd_in, g)
    out = model(x, edges, pos, batch)                       # materialises the lazy parameters
    with torch.no_grad():                                   # non-trivial att / bias, as the tests of the product want
        for n, p in model.named_parameters():
            if not isinstance(p, torch.nn.parameter.UninitializedParameter) and p.dim() == 1:
                p.uniform_(-0.2, 0.2, generator=g)
    model.zero_grad()
    out = model(x, edges, pos, batch)
    gout = {k: torch.randn(v.shape, generator=g) for k, v in out.items()}
    sum((out[k] * gout[k]).sum() for k in out).backward()
    sd, lazy = state_of(model)
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()
             if not isinstance(p, torch.nn.parameter.UninitializedParameter) and p.grad is not None}
    hook = model.conv_layers[0]._attn_weights               # what the reference's forward hook stored (junk, B.2)
    torch.save(dict(hparams=hp, bd_in=bd_in, state_dict=sd, lazy_keys=lazy, x=x, pos=pos, batch=batch, edges=edges,
                    out={k: v.detach() for k, v in out.items()}, grad_out=gout, grads=grads,
                    hook_value_shape=tuple(hook[TT].shape)),
               os.path.join(HERE, f"ref_encoder_{tag}.pt"))


def make_losses():
    ref = reference_import.load().triplet_loss
    g = torch.Generator().manual_seed(4321)
    C, N, D = 7, 150, 16
    a = torch.rand(C, C, generator=g) * 2 - 1
    sim = ((a + a.t()) / 2).contiguous()
    labels = torch.randint(0, C, (N,), generator=g)
    labels[labels == 3] = 2                                  # cluster 3 absent: exercises the `present` remapping
    emb = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    torch.manual_seed(99)
    sel = ref.FastTripletSelector(sim.clone())
    log = []
    with record_random(log):
        pos, neg, dp, dn = sel.sample_triplets(labels)
    uniforms = [t for _, t in log]
    assert len(uniforms) == 4
    out = dict(similarity=sim, labels=labels, emb=emb, uniforms=uniforms, positives=pos, negatives=neg, dists_pos=dp,
               dists_neg=dn, sorted_idx=sel._sorted_idx.clone(), cdf_pos=sel._cdf_pos.clone(),
               cdf_neg=sel._cdf_neg.clone(), present=sel._present.clone())
    for name, make in (("triplet", lambda: ref.TripletLoss(sim.clone(), margin=0.3)),
                       ("metric", lambda: ref.MetricLoss(sim.clone()))):
        e = emb.clone().requires_grad_()
        log = []
        with record_random(log):
            loss = make().forward(e, labels)
        loss.backward()
        out[f"{name}_uniforms"] = [t for _, t in log]
        out[f"{name}_loss"] = loss.detach()
        out[f"{name}_grad"] = e.grad.clone()
    out["empty_triplet"] = ref.TripletLoss(sim.clone(), margin=0.3).forward(emb[:0], labels[:0])
    torch.save(out, os.path.join(HERE, "ref_losses.pt"))


def make_lit():
    ref = reference_import.load()
    g = torch.Generator().manual_seed(77)
    torch.manual_seed(77)
    hp = dict(n_genes=30, in_channels=32, hidden_channels=64, out_channels=64, n_mid_layers=0, n_heads=2)
    N, M, bd_in = 300, 10, 12
    x, pos, bat, edges = small_graph(N, M, hp["n_genes"], bd_in, g, side=40.0)
    b = HeteroBatch()
    b["tx"]["x"], b["tx"]["pos"], b["tx"]["batch"] = x["tx"], pos["tx"], bat["tx"]
    b["bd"]["x"], b["bd"]["pos"], b["bd"]["batch"] = x["bd"], pos["bd"], bat["bd"]
    b["tx"]["index"] = torch.randperm(N, generator=g) + 1000
    b["bd"]["index"] = (torch.randperm(M, generator=g) + 50).to(torch.int32)
    b["tx"]["predict_mask"] = torch.rand(N, generator=g) < 0.7
    b["tx"]["mask"] = torch.rand(N, generator=g) < 0.8
    b["bd"]["mask"] = torch.rand(M, generator=g) < 0.9
    b["tx"]["cluster"] = torch.randint(0, 6, (N,), generator=g)
    bdc = torch.randint(0, 4, (M,), generator=g)
    bdc[1] = -1                                              # unclustered boundary: filtered by `cluster >= 0`
    b["bd"]["cluster"] = bdc
    for et, ei in edges.items():
        b[et]["edge_index"] = ei
    a = torch.rand(6, 6, generator=g) * 2 - 1
    tx_sim = ((a + a.t()) / 2).contiguous()
    a = torch.rand(4, 4, generator=g) * 2 - 1
    bd_sim = ((a + a.t()) / 2).contiguous()

    out = dict(hparams=hp, bd_in=bd_in, batch={k: dict(v) for k, v in b._nodes.items()},
               edges={k: v for k, v in edges.items()}, tx_similarity=tx_sim, bd_similarity=bd_sim)
    results = {}
    for kind in ("triplet", "bce"):
        torch.manual_seed(5)
        lit = ref.lightning_model.LitISTEncoder(sg_loss_type=kind, **hp).eval()
        lit.trainer.max_epochs, lit.current_epoch = 10, 4
        lit.forward(b)                                       # materialise lazy parameters
        # LitISTEncoder.setup, loss part (:108-124) -- the datamodule type check and gene-embedding branch need a
        # Trainer; the loss objects are built exactly as :109-118 does
        lit.loss_tx = ref.triplet_loss.TripletLoss(tx_sim.clone(), margin=lit._tx_margin)
        lit.loss_bd = ref.triplet_loss.MetricLoss(bd_sim.clone())
        lit.loss_sg = (torch.nn.TripletMarginLoss(margin=lit._sg_margin) if kind == "triplet"
                       else torch.nn.BCEWithLogitsLoss())
        lit.zero_grad()
        log = []
        with record_random(log):
            l_tx, l_bd, l_sg, loss = lit.get_losses(b)
        loss.backward()
        orders = dict(tx=lit.loss_tx.selector._sorted_idx.clone(), bd=lit.loss_bd.selector._sorted_idx.clone())
        sd, lazy = state_of(lit)
        grads = {n: p.grad.detach().clone() for n, p in lit.named_parameters()
                 if not isinstance(p, torch.nn.parameter.UninitializedParameter) and p.grad is not None}
        with torch.no_grad():
            p_none = lit.predict_step(b, 0)
            p_thr = lit.predict_step(b, 0, min_similarity=0.5)
            emb = lit.forward(b)
        results[kind] = dict(state_dict=sd, lazy_keys=lazy, draws=log, sorted_idx=orders, loss_tx=l_tx.detach(), loss_bd=l_bd.detach(),
                             loss_sg=l_sg.detach(), loss=loss.detach(), grads=grads, predict=p_none, predict_thr=p_thr,
                             emb={k: v.clone() for k, v in emb.items()})
    out["runs"] = results
    lit = ref.lightning_model.LitISTEncoder(**hp)
    sched = {}
    for max_ep, cur in ((10, 0), (10, 4), (10, 9), (10, 30), (1, 0), (2, 1)):
        lit.trainer.max_epochs, lit.current_epoch = max_ep, cur
        sched[(max_ep, cur)] = (lit._scheduled_weights(lit._w_start.clone(), lit._w_end.clone()),
                                lit._scheduled_weights(lit._w_start.clone(), lit._w_end.clone(), normalize=False))
    out["schedule"] = sched
    torch.save(out, os.path.join(HERE, "ref_lit.pt"))


def make_knn():
    ref = reference_import.load()
    nb = ref.neighbors
    rng = np.random.default_rng(2024)
    pts = rng.uniform(0, 60, (900, 2)).astype(np.float32)
    pts[10] = pts[11]                                         # a coincident pair
    qry = rng.uniform(0, 60, (70, 2)).astype(np.float32)
    e_self, _ = nb.kdtree_neighbors(pts, 5, 5.0)
    e_chunk, _ = nb.kdtree_neighbors(pts, 5, 5.0, chunk_size=256)
    assert torch.equal(e_self, e_chunk)
    e_k20, _ = nb.kdtree_neighbors(pts, 20, 4.0)
    e_qry, _ = nb.kdtree_neighbors(pts, 4, 6.0, query=qry)
    f = ref.fields.TrainingTranscriptFields()
    frame = reference_import.Frame({f.x: pts[:, 0], f.y: pts[:, 1]})
    e_setup = nb.setup_transcripts_graph(frame, 5, 5.0)
    table = torch.from_numpy(rng.integers(0, 41, (40, 6)))    # 40 = padding value (N)
    coo, ptr = nb.knn_to_edge_index(table)
    coo7, ptr7 = nb.knn_to_edge_index(table, padding_value=7)
    np.savez(os.path.join(HERE, "ref_knn.npz"), points=pts, query=qry, e_self=e_self.numpy(), e_k20=e_k20.numpy(),
             e_qry=e_qry.numpy(), e_setup=e_setup.numpy(), x_col=f.x, y_col=f.y, table=table.numpy(), coo=coo.numpy(),
             ptr=ptr.numpy(), coo7=coo7.numpy(), ptr7=ptr7.numpy())


def main():
    make_posemb()
    make_encoder("generic", dict(n_genes=20, in_channels=16, hidden_channels=8, out_channels=8, n_mid_layers=1,
                                 n_heads=2), N=120, M=9, bd_in=12, seed=7)
    make_encoder("quad", dict(n_genes=30, in_channels=32, hidden_channels=64, out_channels=64, n_mid_layers=0,
                              n_heads=2), N=400, M=12, bd_in=12, seed=8)
    make_losses()
    make_lit()
    make_knn()
    print("reference-derived golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
