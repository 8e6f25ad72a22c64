"""GPU parity of the training losses against the CPU oracle on identical inputs: sampler indices
bit-exact given the same uniforms, losses and gradients within 1e-5 of the scale (fp32)."""
import pytest
import torch

from oracle import triplet_loss_ref as R
from segger_b200 import triplet_loss as TL
from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _similarity(C, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.rand(C, C, generator=g) * 2 - 1
    return ((a + a.t()) / 2).contiguous()


@pytest.mark.parametrize("N,C,missing", [(5000, 12, 0), (20000, 40, 7), (64, 5, 2), (1, 3, 0)])
def test_sampler_bit_exact_given_uniforms(N, C, missing):
    """FastTripletSelector.sample_triplets (triplet_loss.py:88-125): same clusters, same members, same distances as
    the oracle for injected uniforms; clusters absent from the batch are never sampled."""
    g = torch.Generator().manual_seed(N + C)
    sim = _similarity(C, 1)
    labels = torch.randint(0, C - missing, (N,), generator=g)
    uni = [torch.rand(N, generator=g) for _ in range(4)]
    ref = R.FastTripletSelectorRef(sim.clone()).sample_triplets(labels, uni)
    got = TL.FastTripletSelector(sim.clone()).sample_triplets(labels.cuda(), [u.cuda() for u in uni])
    assert torch.equal(got[0].cpu(), ref[0]) and torch.equal(got[1].cpu(), ref[1])
    assert torch.equal(got[2].cpu(), ref[2]) and torch.equal(got[3].cpu(), ref[3])
    assert int(labels[got[0].cpu()].max()) < C - missing


def test_sampler_draws_four_uniform_vectors_from_the_torch_generator():
    sim = _similarity(9, 3)
    labels = torch.randint(0, 9, (3000,), generator=torch.Generator().manual_seed(0)).cuda()
    sel = TL.FastTripletSelector(sim.clone())
    torch.manual_seed(11)
    a = sel.sample_triplets(labels)
    torch.manual_seed(11)
    uni = [torch.rand(3000, device="cuda") for _ in range(4)]
    b = sel.sample_triplets(labels, uni)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("N,D,margin", [(4000, 64, 0.3), (777, 20, 1.0), (33, 128, 0.05)])
def test_triplet_and_metric_loss_vs_oracle(N, D, margin):
    g = torch.Generator().manual_seed(N)
    C = 10
    sim = _similarity(C, 2)
    labels = torch.randint(0, C, (N,), generator=g)
    uni = [torch.rand(N, generator=g) for _ in range(4)]
    emb = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    pos, neg, dp, dn = R.FastTripletSelectorRef(sim.clone()).sample_triplets(labels, uni)
    # --- TripletLoss
    e_ref = emb.clone().double().requires_grad_()
    l_ref = R.triplet_loss_ref(e_ref, pos, neg, margin)
    l_ref.backward()
    e = emb.clone().cuda().requires_grad_()
    l = TL.triplet_margin(e, e, e, None, pos.cuda(), neg.cuda(), margin)
    (l * 3.0).backward()
    assert abs(float(l.detach()) - float(l_ref.detach())) < 1e-5 * max(1.0, abs(float(l_ref.detach())))
    assert rel_err(e.grad / 3.0, e_ref.grad) < 1e-5
    # --- MetricLoss
    e_ref = emb.clone().double().requires_grad_()
    m_ref = R.metric_loss_ref(e_ref, pos, neg, dp, dn)
    m_ref.backward()
    e = emb.clone().cuda().requires_grad_()
    m = (TL.cosine_mse(e, e, None, pos.cuda(), (1 - dp).cuda()) + TL.cosine_mse(e, e, None, neg.cuda(), (1 - dn).cuda()))
    m.backward()
    assert abs(float(m.detach()) - float(m_ref.detach())) < 1e-5 * max(1.0, abs(float(m_ref.detach())))
    assert rel_err(e.grad, e_ref.grad) < 1e-5
    # deterministic
    e2 = emb.clone().cuda().requires_grad_()
    TL.triplet_margin(e2, e2, e2, None, pos.cuda(), neg.cuda(), margin).backward()
    e3 = emb.clone().cuda().requires_grad_()
    TL.triplet_margin(e3, e3, e3, None, pos.cuda(), neg.cuda(), margin).backward()
    assert torch.equal(e2.grad, e3.grad)


def test_loss_modules_follow_the_reference_contract():
    """TripletLoss / MetricLoss .forward(embeddings, labels): empty labels -> 0.; a seeded call equals the functional
    form fed with the selector's own samples."""
    sim = _similarity(6, 4)
    lt, lm = TL.TripletLoss(sim.clone(), margin=0.3), TL.MetricLoss(sim.clone())
    emb = torch.randn(500, 32, generator=torch.Generator().manual_seed(1)).cuda()
    labels = torch.randint(0, 6, (500,), generator=torch.Generator().manual_seed(2)).cuda()
    assert lt.forward(emb[:0], labels[:0]) == 0. and lm.forward(emb[:0], labels[:0]) == 0.
    torch.manual_seed(5)
    a = lt.forward(emb, labels)
    torch.manual_seed(5)
    pos, neg, _, _ = lt.selector.sample_triplets(labels)
    assert torch.equal(a, TL.triplet_margin(emb, emb, emb, None, pos, neg, 0.3))
    with pytest.raises(NotImplementedError):
        TL.TripletLoss(sim.clone(), margin=0.3, p=1.0)
    with pytest.raises(ValueError):
        TL.TripletLoss(sim.clone(), margin=0.0)          # torch.nn.TripletMarginLoss rejects margin <= 0


@pytest.mark.parametrize("kind", ["triplet", "bce"])
def test_segmentation_loss_vs_oracle(kind):
    g = torch.Generator().manual_seed(9)
    n_tx, n_bd, D, E = 6000, 80, 64, 2500
    tx = torch.nn.functional.normalize(torch.randn(n_tx, D, generator=g), dim=-1)
    bd = torch.nn.functional.normalize(torch.randn(n_bd, D, generator=g), dim=-1)
    ei = torch.stack([torch.randperm(n_tx, generator=g)[:E], torch.randint(0, n_bd, (E,), generator=g)])
    dst_neg = (ei[1] + torch.randint(1, n_bd, (E,), generator=g)) % n_bd
    tx_r, bd_r = tx.clone().double().requires_grad_(), bd.clone().double().requires_grad_()
    l_ref = R.segmentation_loss_ref(tx_r, bd_r, ei, dst_neg, kind, 0.4)
    l_ref.backward()
    tx_c, bd_c = tx.clone().cuda().requires_grad_(), bd.clone().cuda().requires_grad_()
    l = TL.segmentation_loss(tx_c, bd_c, ei.cuda(), kind, 0.4, dst_neg.cuda())
    l.backward()
    assert abs(float(l.detach()) - float(l_ref.detach())) < 1e-5 * max(1.0, abs(float(l_ref.detach())))
    assert rel_err(tx_c.grad, tx_r.grad) < 1e-5 and rel_err(bd_c.grad, bd_r.grad) < 1e-5
    # one boundary only: zero loss that still carries a graph (lightning_model.py:171-174)
    z = TL.segmentation_loss(tx_c, bd_c[:1], ei.cuda(), kind, 0.4)
    assert float(z.detach()) == 0.0 and z.requires_grad


@pytest.mark.parametrize("kind", ["triplet", "bce"])
def test_get_losses_assembly_vs_oracle(kind):
    """LitISTEncoder.get_losses (lightning_model.py:151-211): masks, the three losses and the scheduled weights,
    against the oracle assembled from the same embeddings, the same samples and the same negatives."""
    from segger_b200.hetero import HeteroBatch
    from segger_b200.lightning_model import LitISTEncoder
    from tests.util import synth_batch
    from oracle.ist_encoder_ref import TB, TT
    ts, x, edges, pos, bat = synth_batch(3000, 30, seed=3)
    torch.manual_seed(0)
    lit = LitISTEncoder(ts.n_genes, in_channels=32, hidden_channels=32, out_channels=32, n_mid_layers=0,
                        sg_loss_type=kind).cuda().eval()
    g = torch.Generator().manual_seed(4)
    sim_tx, sim_bd = _similarity(8, 5), _similarity(4, 6)
    lit.setup_losses(sim_tx.clone(), sim_bd.clone())
    lit.set_epoch(3, 10)
    b = HeteroBatch()
    for k in ("tx", "bd"):
        b[k]["x"], b[k]["pos"], b[k]["batch"] = x[k], pos[k], bat[k]
    b["tx"]["mask"] = torch.rand(3000, generator=g) < 0.7
    b["tx"]["cluster"] = torch.randint(0, 8, (3000,), generator=g)
    b["bd"]["mask"] = torch.rand(30, generator=g) < 0.9
    b["bd"]["cluster"] = torch.randint(-1, 4, (30,), generator=g)
    b[TT]["edge_index"], b[TB]["edge_index"] = edges[TT], edges[TB]
    bc = b.cuda()
    with torch.no_grad():
        lit.forward(bc)              # materialise the lazy (-1) parameters first: their init draws random numbers
    torch.manual_seed(21)
    l_tx, l_bd, l_sg, loss = lit.get_losses(bc)
    loss.backward()
    # oracle: same embeddings (detached product output), same random draws replayed in the reference's order
    with torch.no_grad():
        emb = lit.forward(bc)
    e_tx, e_bd = emb["tx"].cpu().double().requires_grad_(), emb["bd"].cpu().double().requires_grad_()
    tx_mask = b["tx"]["mask"]
    bd_mask = b["bd"]["mask"] & (b["bd"]["cluster"] >= 0)
    torch.manual_seed(21)
    n_t, n_b = int(tx_mask.sum()), int(bd_mask.sum())
    u_t = [torch.rand(n_t, device="cuda").cpu() for _ in range(4)]
    u_b = [torch.rand(n_b, device="cuda").cpu() for _ in range(4)]
    E = edges[TB].size(1)
    dst_neg = ((edges[TB][1].cuda() + torch.randint(1, 30, (E,), device="cuda")) % 30).cpu()
    p, n, _, _ = R.FastTripletSelectorRef(sim_tx.clone()).sample_triplets(b["tx"]["cluster"][tx_mask], u_t)
    r_tx = R.triplet_loss_ref(e_tx[tx_mask], p, n, 0.3)
    p, n, dp, dn = R.FastTripletSelectorRef(sim_bd.clone()).sample_triplets(b["bd"]["cluster"][bd_mask], u_b)
    r_bd = R.metric_loss_ref(e_bd[bd_mask], p, n, dp, dn)
    r_sg = R.segmentation_loss_ref(e_tx, e_bd, edges[TB], dst_neg, kind, 0.4)
    w = R.scheduled_weights_ref(torch.tensor([1., 1., 0.]), torch.tensor([1., 1., .5]), 3, 10)
    r = w[0] * r_tx + w[1] * r_bd + w[2] * r_sg
    for got, ref in ((l_tx, r_tx), (l_bd, r_bd), (l_sg, r_sg), (loss, r)):
        assert abs(float(got.detach()) - float(ref.detach())) < 2e-5 * max(1.0, abs(float(ref.detach())))
    assert all(p_.grad is not None and bool(torch.isfinite(p_.grad).all()) for n_, p_ in lit.model.named_parameters()
               if "contains" not in n_)


def test_losses_against_committed_golden_fixture():
    """tests/golden/losses_small.pt (oracle outputs frozen by tests/golden/make_golden.py): sampler bit-exact, the four
    losses and the summed embedding gradient within fp32 rounding."""
    import os
    gd = torch.load(os.path.join(os.path.dirname(__file__), "golden", "losses_small.pt"))
    sel = TL.FastTripletSelector(gd["similarity"].clone())
    pos, neg, dp, dn = sel.sample_triplets(gd["labels"].cuda(), [u.cuda() for u in gd["uniforms"]])
    assert torch.equal(pos.cpu(), gd["positives"]) and torch.equal(neg.cpu(), gd["negatives"])
    assert torch.equal(dp.cpu(), gd["dists_pos"]) and torch.equal(dn.cpu(), gd["dists_neg"])
    e, bd = gd["emb"].cuda().requires_grad_(), gd["bd"].cuda()
    ei, dst_neg = gd["edge_index"].cuda(), gd["dst_neg"].cuda()
    l_t = TL.triplet_margin(e, e, e, None, pos, neg, 0.3)
    l_m = TL.cosine_mse(e, e, None, pos, 1 - dp) + TL.cosine_mse(e, e, None, neg, 1 - dn)
    l_s = TL.segmentation_loss(e, bd, ei, "triplet", 0.4, dst_neg)
    l_b = TL.segmentation_loss(e, bd, ei, "bce", 0.4, dst_neg)
    (l_t + l_m + l_s + l_b).backward()
    for got, key in ((l_t, "loss_triplet"), (l_m, "loss_metric"), (l_s, "loss_seg_triplet"), (l_b, "loss_seg_bce")):
        assert abs(float(got.detach()) - float(gd[key])) < 2e-6
    assert rel_err(e.grad, gd["grad"]) < 1e-5


@pytest.mark.parametrize("N,D", [(50_000, 64), (3001, 32), (129, 128)])
def test_triplet_row_owner_backward_equals_per_triplet_backward(N, D, monkeypatch):
    """TripletLoss's call shape (one matrix three times, anchors = all rows) takes the one-pass row-owner backward
    (sgb_triplet_self_bwd over CSRs of the sampled indices); it must equal the per-triplet rows + segment sums of the
    general path to rounding, with heavily repeated positives / negatives and rows nobody sampled."""
    g = torch.Generator().manual_seed(N + D)
    emb = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    hot = max(2, N // 50)                                   # most triplets point at a few rows
    pos = torch.randint(0, hot, (N,), generator=g)
    neg = torch.randint(0, N, (N,), generator=g)
    neg[: N // 3] = N - 1 - torch.randint(0, hot, (N // 3,), generator=g)
    grads = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("SEGGER_B200_LOSS_FUSED", flag)
        e = emb.clone().cuda().requires_grad_()
        loss = TL.triplet_margin(e, e, e, None, pos.cuda(), neg.cuda(), 0.4)
        (loss * 2.0).backward()
        grads[flag] = (float(loss.detach()), e.grad.clone())
    assert grads["0"][0] == grads["1"][0]
    assert rel_err(grads["1"][1], grads["0"][1]) < 2e-6
    e_ref = emb.clone().double().requires_grad_()
    R.triplet_loss_ref(e_ref, pos, neg, 0.4).backward()
    assert rel_err(grads["1"][1] / 2.0, e_ref.grad) < 1e-5


def test_triplet_row_owner_backward_heavy_rows():
    """A row that thousands of triplets sampled (a tiny cluster beside large ones): the compensated left fold of the
    row-owner backward stays within 1e-5 of float64."""
    g = torch.Generator().manual_seed(4)
    N, D = 20_000, 64
    emb = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    pos = torch.randint(0, 3, (N,), generator=g)            # three rows share all the positives
    neg = N - 1 - torch.randint(0, 2, (N,), generator=g)    # two rows share all the negatives
    e = emb.clone().cuda().requires_grad_()
    TL.triplet_margin(e, e, e, None, pos.cuda(), neg.cuda(), 0.4).backward()
    e_ref = emb.clone().double().requires_grad_()
    R.triplet_loss_ref(e_ref, pos, neg, 0.4).backward()
    assert rel_err(e.grad, e_ref.grad) < 1e-5
    assert rel_err(e.grad[:3], e_ref.grad[:3]) < 1e-5 and rel_err(e.grad[-2:], e_ref.grad[-2:]) < 1e-5
