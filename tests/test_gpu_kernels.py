"""GPU parity tests, kernel level: every check goes through the C ABI (ctypes) and compares with the
CPU oracle on identical seeded inputs.  Tolerance for fp32 work: 1e-4 relative to the tensor's scale
(BASELINE.json north_star); integer/index work is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import pyg_ref
from segger_b200 import ops
from segger_b200._lib import ACT_GELU, ACT_NONE, ACT_SILU
from tests.util import random_graph, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


# ---------------------------------------------------------------------------------------------- CSR
@pytest.mark.parametrize("E,n_src,n_dst,dtype", [
    (0, 5, 7, torch.int64), (1, 1, 1, torch.int64), (37, 10, 12, torch.int32),
    (5000, 300, 200, torch.int64), (70000, 70000, 1000, torch.int32), (300000, 5000, 100000, torch.int64)])
def test_csr_build_bit_exact(E, n_src, n_dst, dtype):
    ei = random_graph(n_src, n_dst, E, seed=E + 1, dtype=dtype) if E else torch.zeros(2, 0, dtype=dtype)
    csr = ops.build_csr(ei.cuda(), n_src, n_dst, transpose=True)
    src, dst = ei[0].long(), ei[1].long()
    order = torch.argsort(dst, stable=True)
    assert torch.equal(csr.eid.cpu().long(), order)
    assert torch.equal(csr.col.cpu().long(), src[order])
    rp = torch.zeros(n_dst + 1, dtype=torch.long)
    rp[1:] = torch.bincount(dst, minlength=n_dst).cumsum(0)
    assert torch.equal(csr.rowptr.cpu().long(), rp)
    torder = torch.argsort(src, stable=True)
    trp = torch.zeros(n_src + 1, dtype=torch.long)
    trp[1:] = torch.bincount(src, minlength=n_src).cumsum(0)
    assert torch.equal(csr.t_rowptr.cpu().long(), trp)
    assert torch.equal(csr.t_dst.cpu().long(), dst[torder])
    inv = torch.empty(E, dtype=torch.long)
    inv[order] = torch.arange(E)
    assert torch.equal(csr.t_pos.cpu().long(), inv[torder])
    assert int(csr.status[0]) == 0


@pytest.mark.parametrize("sorted_row", [0, 1])
def test_csr_build_presorted_inputs_take_the_copy_path(sorted_row):
    """kNN / setup_heterodata emit src-major edge lists: the sort detects non-decreasing keys on the device and
    copies instead of sorting; the result must equal the stable sort's."""
    ei = random_graph(3000, 2500, 40000, seed=9)
    order = torch.argsort(ei[sorted_row], stable=True)
    ei = ei[:, order].contiguous()
    csr = ops.build_csr(ei.cuda(), 3000, 2500, transpose=True)
    src, dst = ei[0], ei[1]
    o = torch.argsort(dst, stable=True)
    assert torch.equal(csr.eid.cpu().long(), o) and torch.equal(csr.col.cpu().long(), src[o])
    to = torch.argsort(src, stable=True)
    inv = torch.empty(ei.size(1), dtype=torch.long)
    inv[o] = torch.arange(ei.size(1))
    assert torch.equal(csr.t_dst.cpu().long(), dst[to]) and torch.equal(csr.t_pos.cpu().long(), inv[to])
    rp = torch.zeros(2501, dtype=torch.long); rp[1:] = torch.bincount(dst, minlength=2500).cumsum(0)
    trp = torch.zeros(3001, dtype=torch.long); trp[1:] = torch.bincount(src, minlength=3000).cumsum(0)
    assert torch.equal(csr.rowptr.cpu().long(), rp) and torch.equal(csr.t_rowptr.cpu().long(), trp)


def test_csr_strided_view_and_range_flag():
    ei = random_graph(50, 60, 400, seed=3)
    eit = ei.t().contiguous().cuda().t()          # [2,E] view with strides (1, 2)
    assert not eit.is_contiguous()
    a = ops.build_csr(eit, 50, 60)
    b = ops.build_csr(ei.cuda(), 50, 60)
    for f in ("rowptr", "col", "eid", "t_rowptr", "t_dst", "t_pos"):
        assert torch.equal(getattr(a, f), getattr(b, f))
    bad = ei.clone(); bad[1, 5] = 999
    assert int(ops.build_csr(bad.cuda(), 50, 60).status[0]) == 1


# ---------------------------------------------------------------------------------------------- GATv2
def _gat_case(n_src, n_dst, E, H, C, seed, bipartite=True):
    g = torch.Generator().manual_seed(seed)
    F = H * C
    x_l = torch.randn(n_src, F, generator=g)
    x_r = torch.randn(n_dst, F, generator=g) if bipartite else torch.randn(n_src, F, generator=g)
    att = torch.randn(1, H, C, generator=g) * 0.3
    bias = torch.randn(F, generator=g) * 0.1
    ei = random_graph(n_src, x_r.size(0), E, seed=seed) if E else torch.zeros(2, 0, dtype=torch.long)
    return x_l, x_r, att, bias, ei


SHAPES = [  # (H, C): vector path combos + generic path
    (2, 64), (1, 128), (4, 32), (2, 128), (4, 64), (8, 32), (1, 256), (3, 128), (4, 128), (8, 64), (2, 256),
    (1, 512), (3, 32), (1, 16), (2, 50), (5, 7), (1, 200)]


@pytest.mark.parametrize("H,C", SHAPES)
def test_gatv2_fwd_bwd_vs_oracle(H, C):
    n_src, n_dst, E = 257, 131, 1900
    x_l, x_r, att, bias, ei = _gat_case(n_src, n_dst, E, H, C, seed=H * 1000 + C)
    xl_r, xr_r, att_r, b_r = (t.clone().requires_grad_() for t in (x_l, x_r, att, bias))
    ref = pyg_ref.gatv2_aggregate(xl_r.view(n_src, H, C), xr_r.view(n_dst, H, C), ei, att_r, b_r)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(7))
    (ref * w).sum().backward()

    csr = ops.build_csr(ei.cuda(), n_src, n_dst)
    xl_g, xr_g, att_g, b_g = (t.clone().cuda().requires_grad_() for t in (x_l, x_r, att, bias))
    out = ops.GATv2AggregateFn.apply(xl_g, xr_g, att_g, b_g, csr, H, C, 0.2, 0.0, False, 0, False)
    (out * w.cuda()).sum().backward()
    assert rel_err(out, ref) < TOL
    assert rel_err(xl_g.grad, xl_r.grad) < TOL
    assert rel_err(xr_g.grad, xr_r.grad) < TOL
    assert rel_err(att_g.grad, att_r.grad) < TOL
    assert rel_err(b_g.grad, b_r.grad) < TOL
    # isolated destination rows: out == bias exactly
    deg = torch.bincount(ei[1], minlength=n_dst)
    assert torch.equal(out.detach().cpu()[deg == 0], bias.expand(int((deg == 0).sum()), -1))


def test_gatv2_fused_gelu_and_strided_slices():
    H, C, N, E = 2, 64, 500, 3000
    F = H * C
    x_l, _, att, bias, ei = _gat_case(N, N, E, H, C, seed=5, bipartite=False)
    g = torch.Generator().manual_seed(9)
    y = torch.randn(N, 3 * F, generator=g)                       # concatenated projection buffer
    yr = y.clone().requires_grad_()
    ref = torch.nn.functional.gelu(pyg_ref.gatv2_aggregate(
        yr[:, :F].reshape(N, H, C), yr[:, F:2 * F].reshape(N, H, C), ei, att, bias))
    w = torch.randn(N, F, generator=g)
    (ref * w).sum().backward()
    csr = ops.build_csr(ei.cuda(), N, N)
    yg = y.cuda()
    out, h, smax, sden = ops.gatv2_fwd(yg[:, :F], yg[:, F:2 * F], att.cuda(), bias.cuda(), csr, H, C, 0.2, 0.0,
                                       False, 0, True)
    assert rel_err(h, ref) < TOL
    G = torch.zeros(N, 3 * F, device="cuda")
    ops.gatv2_bwd(yg[:, :F], yg[:, F:2 * F], att.cuda(), bias.cuda(), out, w.cuda(), True, csr, H, C, 0.2, 0.0,
                  False, 0, smax, sden, grad_x_l=G[:, :F], grad_x_r=G[:, F:2 * F])
    assert rel_err(G[:, :2 * F], yr.grad[:, :2 * F]) < TOL
    assert float(G[:, 2 * F:].abs().max()) == 0.0


@pytest.mark.parametrize("H,C", [(2, 64), (4, 128)])
def test_gatv2_large_graph_multi_chunk(H, C):
    """Enough rows that every warp of the persistent dst pass owns several row chunks."""
    n_src, n_dst, E = 30_011, 40_003, 160_000
    x_l, x_r, att, bias, ei = _gat_case(n_src, n_dst, E, H, C, seed=77)
    xl_r, xr_r, att_r, b_r = (t.clone().requires_grad_() for t in (x_l, x_r, att, bias))
    ref = torch.nn.functional.gelu(
        pyg_ref.gatv2_aggregate(xl_r.view(n_src, H, C), xr_r.view(n_dst, H, C), ei, att_r, b_r))
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(7))
    (ref * w).sum().backward()
    csr = ops.build_csr(ei.cuda(), n_src, n_dst)
    xl_g, xr_g, att_g, b_g = (t.clone().cuda().requires_grad_() for t in (x_l, x_r, att, bias))
    out = ops.GATv2AggregateFn.apply(xl_g, xr_g, att_g, b_g, csr, H, C, 0.2, 0.0, False, 0, True)
    (out * w.cuda()).sum().backward()
    assert rel_err(out, ref) < TOL
    assert rel_err(xl_g.grad, xl_r.grad) < TOL
    assert rel_err(xr_g.grad, xr_r.grad) < TOL
    assert rel_err(att_g.grad, att_r.grad) < TOL
    assert rel_err(b_g.grad, b_r.grad) < TOL


def test_gatv2_edge_cases_empty_and_single():
    H, C = 2, 64
    F = H * C
    # E = 0: output is the bias, gradients are zero
    x_l, x_r = torch.randn(5, F), torch.randn(4, F)
    att, bias = torch.randn(1, H, C), torch.randn(F)
    csr = ops.build_csr(torch.zeros(2, 0, dtype=torch.long).cuda(), 5, 4)
    xl_g, xr_g = x_l.cuda().requires_grad_(), x_r.cuda().requires_grad_()
    out = ops.GATv2AggregateFn.apply(xl_g, xr_g, att.cuda(), bias.cuda(), csr, H, C, 0.2, 0.0, False, 0, False)
    assert torch.equal(out.detach().cpu(), bias.expand(4, -1))
    out.sum().backward()
    assert float(xl_g.grad.abs().max()) == 0.0 and float(xr_g.grad.abs().max()) == 0.0
    # a hub: one destination with a long row (exercises the chunk loop), self loops, int32 indices
    n = 300
    ei = torch.stack([torch.arange(n), torch.zeros(n, dtype=torch.long)]).int()
    x = torch.randn(n, F)
    ref = pyg_ref.gatv2_aggregate(x.view(n, H, C), x.view(n, H, C), ei.long(), att, bias)
    csr = ops.build_csr(ei.cuda(), n, n)
    got, _, _, _ = ops.gatv2_fwd(x.cuda(), x.cuda(), att.cuda(), bias.cuda(), csr, H, C, 0.2, 0.0, False, 0, False)
    assert rel_err(got, ref) < TOL


@pytest.mark.parametrize("H,C", [(2, 64), (3, 32)])
def test_gatv2_dropout_replays_injected_mask(H, C):
    n_src, n_dst, E, p = 200, 150, 2500, 0.2
    x_l, x_r, att, bias, ei = _gat_case(n_src, n_dst, E, H, C, seed=11)
    seed = 123456789
    keep = ops.dropout_keep_mask(seed, E, H, p, "cuda").cpu()
    assert abs(float(keep.float().mean()) - (1 - p)) < 0.02          # drop rate
    xl_r, xr_r = x_l.clone().requires_grad_(), x_r.clone().requires_grad_()
    ref = pyg_ref.gatv2_aggregate(xl_r.view(n_src, H, C), xr_r.view(n_dst, H, C), ei, att, bias,
                                  dropout_p=p, training=True, keep_mask=keep)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3))
    (ref * w).sum().backward()
    csr = ops.build_csr(ei.cuda(), n_src, n_dst)
    xl_g, xr_g = x_l.cuda().requires_grad_(), x_r.cuda().requires_grad_()
    out = ops.GATv2AggregateFn.apply(xl_g, xr_g, att.cuda(), bias.cuda(), csr, H, C, 0.2, p, True, seed, False)
    (out * w.cuda()).sum().backward()
    assert rel_err(out, ref) < TOL
    assert rel_err(xl_g.grad, xl_r.grad) < TOL
    assert rel_err(xr_g.grad, xr_r.grad) < TOL


def test_gatv2_deterministic_and_permutation_invariant():
    H, C, N, E = 2, 64, 400, 5000
    x_l, x_r, att, bias, ei = _gat_case(N, N, E, H, C, seed=21, bipartite=False)
    args = (x_l.cuda(), x_r.cuda(), att.cuda(), bias.cuda())
    csr = ops.build_csr(ei.cuda(), N, N)
    a = ops.gatv2_fwd(*args, csr, H, C, 0.2, 0.0, False, 0, False)[0]
    b = ops.gatv2_fwd(*args, csr, H, C, 0.2, 0.0, False, 0, False)[0]
    assert torch.equal(a, b)                                          # bit-reproducible
    perm = torch.randperm(E, generator=torch.Generator().manual_seed(1))
    c = ops.gatv2_fwd(*args, ops.build_csr(ei[:, perm].cuda(), N, N), H, C, 0.2, 0.0, False, 0, False)[0]
    assert rel_err(c, a) < 1e-5                                       # only summation order changes
    alpha = ops.gatv2_alpha(args[0], args[1], args[2], csr, H, C, 0.2, *ops.gatv2_fwd(
        *args, csr, H, C, 0.2, 0.0, False, 0, False)[2:])
    _, alpha_ref = pyg_ref.gatv2_aggregate(x_l.view(N, H, C), x_r.view(N, H, C), ei, att, bias, return_alpha=True)
    assert rel_err(alpha, alpha_ref) < TOL


# ---------------------------------------------------------------------------------------------- linear
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (7, 5, 3), (300, 384, 256), (1000, 64, 128), (129, 130, 131),
                                   (4096, 8, 256), (50, 128, 50)])
@pytest.mark.parametrize("act", [ACT_NONE, ACT_GELU, ACT_SILU])
def test_linear_fwd_bwd_vs_torch(M, N, K, act):
    g = torch.Generator().manual_seed(M * 7 + N)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    xr, wr, br = (t.clone().requires_grad_() for t in (x, w, b))
    y = torch.nn.functional.linear(xr, wr, br)
    y = {ACT_NONE: y, ACT_GELU: torch.nn.functional.gelu(y), ACT_SILU: torch.nn.functional.silu(y)}[act]
    gy = torch.randn(M, N, generator=g)
    (y * gy).sum().backward()
    xg, wg, bg = (t.clone().cuda().requires_grad_() for t in (x, w, b))
    yg = ops.linear(xg, wg, bg, act)
    (yg * gy.cuda()).sum().backward()
    assert rel_err(yg, y) < TOL
    assert rel_err(xg.grad, xr.grad) < TOL
    assert rel_err(wg.grad, wr.grad) < TOL
    assert rel_err(bg.grad, br.grad) < TOL


def test_linear_wgrad_large_m_splitk_deterministic():
    M, N, K = 200_000, 384, 128
    g = torch.Generator().manual_seed(1)
    dy, x = torch.randn(M, N, generator=g), torch.randn(M, K, generator=g)
    dw, db = ops.linear_wgrad(dy.cuda(), x.cuda())
    dw2, db2 = ops.linear_wgrad(dy.cuda(), x.cuda())
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
    ref = dy.double().t() @ x.double()
    assert rel_err(dw, ref) < TOL
    assert rel_err(db, dy.double().sum(0)) < TOL


# ---------------------------------------------------------------------------------------------- scoring
@pytest.mark.parametrize("D", [64, 32, 128, 256, 20])
def test_score_argmax_vs_oracle(D):
    from oracle.ist_encoder_ref import predict_scores_ref
    g = torch.Generator().manual_seed(D)
    n_tx, n_bd, E = 3000, 200, 5000
    tx = torch.nn.functional.normalize(torch.randn(n_tx, D, generator=g), dim=-1)
    bd = torch.nn.functional.normalize(torch.randn(n_bd, D, generator=g), dim=-1)
    src = torch.randint(0, n_tx - 500, (E,), generator=g)           # last 500 tx have no candidate
    dst = torch.randint(0, n_bd, (E,), generator=g)
    src[src == 7] = 8
    src[10:14] = 7; dst[10:14] = 3                                   # exact ties -> lowest edge id
    ei = torch.stack([src, dst]).int()
    bd_index = torch.randperm(n_bd, generator=g).int() + 1000
    for ms in (None, 0.05):
        seg_r, sim_r, idx_r = predict_scores_ref(tx, bd, ei, bd_index, ms)
        sim, idx, seg = ops.score_argmax(tx.cuda(), bd.cuda(), ei.cuda(), bd_index.cuda(), ms)
        assert rel_err(sim, sim_r) < TOL
        agree = (idx.cpu() == idx_r).float().mean()
        assert agree >= 0.9999, agree
        assert (seg.cpu() == seg_r).float().mean() >= 0.9999
        assert int(idx[7]) == 10
        assert torch.equal(idx.cpu()[-500:], torch.full((500,), E)) and torch.equal(seg.cpu()[-500:], torch.full((500,), -1))
        assert float(sim[-500:].abs().max()) == 0.0


def test_score_argmax_no_edges():
    sim, idx, seg = ops.score_argmax(torch.randn(10, 64).cuda(), torch.randn(3, 64).cuda(),
                                     torch.zeros(2, 0, dtype=torch.int32).cuda(), torch.arange(3).int().cuda())
    assert float(sim.abs().max()) == 0.0 and torch.equal(idx.cpu(), torch.zeros(10, dtype=torch.long))
    assert torch.equal(seg.cpu(), torch.full((10,), -1))


# ---------------------------------------------------------------------------------------------- posfreq
@pytest.mark.parametrize("dim,with_batch", [(256, True), (256, False), (64, True), (10, True), (7, False)])
def test_posfreq_vs_oracle(dim, with_batch):
    """sinusoid features of per-tile normalised coordinates (ist_encoder.py:22-31,57-79): the vectorised
    kernel (half % 4 == 0) and the scalar fallback (odd / small dims) against the oracle's restatement."""
    from oracle.ist_encoder_ref import sinusoidal_embedding
    g = torch.Generator().manual_seed(dim)
    N, nb = 3001, 4
    pos = torch.rand(N, 2, generator=g) * 300.0
    batch = torch.sort(torch.randint(0, nb, (N,), generator=g)).values if with_batch else None
    if with_batch:
        mins, maxs = torch.zeros(nb, 2), torch.zeros(nb, 2)
        for b in range(nb):
            mins[b], maxs[b] = pos[batch == b].min(0).values, pos[batch == b].max(0).values
        pn = (pos - mins[batch]) / (maxs[batch] - mins[batch] + 1e-8)
    else:
        pn = pos - pos.min(0).values
        pn = pn / pn.max(0).values
    ref = sinusoidal_embedding(pn.flatten(), dim, max_period=10000).reshape(N, 2, dim).permute(1, 0, 2)
    got = ops.posfreq(pos.cuda(), batch.cuda() if with_batch else None, nb if with_batch else 1, dim,
                      ops.sinusoid_freqs(dim, 10000, "cuda"))
    assert got.shape == (2, N, dim)
    assert float((got.cpu() - ref).abs().max()) < 2e-6


@pytest.mark.parametrize("with_batch", [True, False])
def test_poscheb_lowrank_reproduces_sinusoid_features(with_batch):
    """Chebyshev basis @ constant coefficient matrix == the 256 sinusoid columns (ist_encoder.py:22-31) to fp32
    rounding: the identity the fused input stage relies on to contract over 12 columns instead of 256."""
    g = torch.Generator().manual_seed(11)
    N, nb = 5003, 3
    pos = torch.rand(N, 2, generator=g) * 1000.0
    batch = torch.sort(torch.randint(0, nb, (N,), generator=g)).values if with_batch else None
    freqs = ops.sinusoid_freqs(256, 10000, "cuda")
    feat = ops.posfreq(pos.cuda(), batch.cuda() if with_batch else None, nb if with_batch else 1, 256, freqs)
    T = ops.poscheb(pos.cuda(), batch.cuda() if with_batch else None, nb if with_batch else 1)
    M = ops.cheb_feature_matrix(freqs)
    assert T.shape == (2 * N, ops.CHEB_DEG) and M.shape == (ops.CHEB_DEG, 256)
    low = (T.double() @ M.double()).view(2, N, 256)
    assert float((low - feat.double()).abs().max()) < 2e-6
    with pytest.raises(ValueError):
        ops.cheb_feature_matrix(freqs * 40.0)          # frequencies far above 1: series not converged at deg 12


@pytest.mark.parametrize("D,dtype", [(128, torch.int32), (16, torch.int64), (6, torch.int32)])
def test_embedding_gather_gelu_vs_torch(D, dtype):
    """Input stage of the transcript branch without positional features: GELU(Embedding(ids))
    (ist_encoder.py:312-320); vectorised kernel for D % 4 == 0, scalar fallback otherwise; backward =
    deterministic segment sum of the incoming gradient times GELU'."""
    g = torch.Generator().manual_seed(D)
    n_rows, N = 37, 4001
    table = torch.randn(n_rows, D, generator=g, dtype=torch.float64).float().requires_grad_()
    ids = torch.randint(0, n_rows, (N,), generator=g).to(dtype)
    ref = torch.nn.functional.gelu(table.double()[ids.long()])
    go = torch.randn(N, D, generator=g)
    ref.backward(go.double())
    t_cuda = table.detach().cuda().requires_grad_()
    h = ops.InputStageFn.apply(ids.cuda(), t_cuda, None, None, None, None, None, None, True, True)
    assert rel_err(h, ref) < 1e-6
    h.backward(go.cuda())
    assert rel_err(t_cuda.grad, table.grad) < 1e-5


@pytest.mark.parametrize("M,N,sliced", [(1000, 64, False), (999, 7, False), (500, 64, True)])
@pytest.mark.parametrize("act", [ACT_GELU, ACT_SILU])
def test_act_fwd_bwd_vs_torch(M, N, sliced, act):
    """erf-GELU / SiLU (ist_encoder.py:46,320,325) forward and derivative: 128-bit kernel for aligned
    shapes (also on column slices of a wider buffer), scalar fallback otherwise."""
    g = torch.Generator().manual_seed(M + N)
    big = torch.randn(M, 3 * N if sliced else N, generator=g) * 2
    x = big[:, N:2 * N] if sliced else big
    dy = torch.randn(M, N, generator=g)
    xd = x.double().requires_grad_()
    f = torch.nn.functional.gelu if act == ACT_GELU else torch.nn.functional.silu
    ref = f(xd)
    ref.backward(dy.double())
    xc = big.cuda()[:, N:2 * N] if sliced else big.cuda()
    assert rel_err(ops.act_fwd(xc, act), ref) < 1e-6
    assert rel_err(ops.act_bwd(dy.cuda(), xc, act), xd.grad) < 1e-5


@pytest.mark.parametrize("D", [64, 32, 128, 20])
def test_output_stage_linear_normalize_vs_torch(D):
    """lin_last + F.normalize(dim=-1, eps=1e-12) (ist_encoder.py:328-332), forward and backward, with an
    all-zero row (norm clamped at eps); sub-warp 128-bit kernels for D in {32, 64, 128}, scalar otherwise."""
    g = torch.Generator().manual_seed(D)
    M, K = 777, 48
    h = torch.randn(M, K, generator=g)
    h[5] = 0.0
    w, b = torch.randn(D, K, generator=g) / 7, torch.zeros(D)
    go = torch.randn(M, D, generator=g)
    go[5] = 0.0          # the clamped row's gradient is g / eps = 1e12 * g: keep it out of the max-norm comparison
    hd, wd, bd = h.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
    ref = torch.nn.functional.normalize(hd @ wd.t() + bd, dim=-1)
    ref.backward(go.double())
    hc, wc, bc = h.cuda().requires_grad_(), w.cuda().requires_grad_(), b.cuda().requires_grad_()
    out = ops.OutputStageFn.apply(hc, wc, bc, True)
    assert rel_err(out, ref) < 1e-5
    assert float(out[5].abs().max()) == 0.0
    out.backward(go.cuda())
    assert rel_err(hc.grad, hd.grad) < 1e-4 and rel_err(wc.grad, wd.grad) < 1e-4


@pytest.mark.parametrize("H,C", [(2, 64), (3, 32)])
def test_gatv2_fwd_activated_output_only(H, C):
    """Inference writes only GELU(out) (out = NULL in the C ABI): bit-identical to the two-output call."""
    F = H * C
    g = torch.Generator().manual_seed(H)
    ei = random_graph(900, 700, 6000, seed=21).cuda()
    x_l, x_r = torch.randn(900, F, generator=g).cuda(), torch.randn(700, F, generator=g).cuda()
    att, bias = (torch.randn(F, generator=g) * 0.3).cuda(), (torch.randn(F, generator=g) * 0.1).cuda()
    csr = ops.build_csr(ei, 900, 700, transpose=False)
    pre, act, m1, s1 = ops.gatv2_fwd(x_l, x_r, att, bias, csr, H, C, 0.2, 0.0, False, 0, True)
    none, act2, m2, s2 = ops.gatv2_fwd(x_l, x_r, att, bias, csr, H, C, 0.2, 0.0, False, 0, True, want_pre=False)
    assert none is None and torch.equal(act, act2) and torch.equal(m1, m2) and torch.equal(s1, s2)
    assert rel_err(act, torch.nn.functional.gelu(pre.double())) < 1e-6


@pytest.mark.parametrize("H,C", [(2, 64), (3, 32), (4, 128)])
def test_gatv2_bwd_one_source_per_edge_form(H, C):
    """tx-belongs-bd with one virtual source per edge (EdgeCSR.per_edge_sources): the dst pass writes grad_x_l itself
    (NULL transposed CSR in the C ABI); must equal the two-pass backward on the same virtual graph (to rounding: the
    one-pass form adds the rounded product where the two-pass form uses an FMA)."""
    F = H * C
    g = torch.Generator().manual_seed(C)
    n_src, n_dst, E = 5000, 300, 2100
    src = torch.sort(torch.randperm(n_src, generator=g)[:E]).values
    ei = torch.stack([src, torch.randint(0, n_dst - 20, (E,), generator=g)]).cuda()
    csr = ops.build_csr(ei, n_src, n_dst)
    assert csr.sources_unique_increasing()
    v = csr.per_edge_sources()
    assert v.one_source_per_edge and v.n_src == E
    x_l, x_r = torch.randn(E, F, generator=g).cuda(), torch.randn(n_dst, F, generator=g).cuda()
    att, bias = (torch.randn(F, generator=g) * 0.3).cuda(), (torch.randn(F, generator=g) * 0.1).cuda()
    go = torch.randn(n_dst, F, generator=g).cuda()
    out, _, smax, sden = ops.gatv2_fwd(x_l, x_r, att, bias, v, H, C, 0.2, 0.2, True, 5, True)
    a = ops.gatv2_bwd(x_l, x_r, att, bias, out, go, True, v, H, C, 0.2, 0.2, True, 5, smax, sden)
    two_pass = ops.EdgeCSR(v.rowptr, v.col, v.eid, v.t_rowptr, v.t_dst, v.t_pos, v.n_src, v.n_dst, v.E, v.status)
    b = ops.gatv2_bwd(x_l, x_r, att, bias, out, go, True, two_pass, H, C, 0.2, 0.2, True, 5, smax, sden)
    for u, w in zip(a, b):
        assert rel_err(u, w) < 1e-6
    assert torch.equal(a[2], b[2])                     # grad_att: same products, same order
    a2 = ops.gatv2_bwd(x_l, x_r, att, bias, out, go, True, v, H, C, 0.2, 0.2, True, 5, smax, sden)
    assert all(torch.equal(u, w) for u, w in zip(a, a2))


@pytest.mark.parametrize("H,C", [(2, 64), (1, 128), (4, 32), (3, 32), (4, 128), (1, 512)])
@pytest.mark.parametrize("direct", [False, True])
def test_gatv2_bwd_saved_logits_equal_recomputed(H, C, direct):
    """The sub-warp forward leaves the raw logits [E, H] (dst-CSR order); a backward handed that buffer reads them back
    instead of recomputing att . lrelu(x_l[j] + x_r[i]) and applies att once per row.  Must equal the recomputing
    backward to rounding -- two-pass form and one-source-per-edge form, dropout on, fused GELU' -- and must match the
    logits the oracle computes."""
    F = H * C
    g = torch.Generator().manual_seed(10 * C + H)
    if direct:
        n_src, n_dst, E = 5000, 300, 2100
        src = torch.sort(torch.randperm(n_src, generator=g)[:E]).values
        ei = torch.stack([src, torch.randint(0, n_dst - 20, (E,), generator=g)])
        csr = ops.build_csr(ei.cuda(), n_src, n_dst).per_edge_sources()
        n_l = E
    else:
        n_src, n_dst, E = 700, 401, 4000
        ei = random_graph(n_src, n_dst, E, seed=C)
        csr = ops.build_csr(ei.cuda(), n_src, n_dst)
        n_l = n_src
    x_l, x_r = torch.randn(n_l, F, generator=g).cuda(), torch.randn(n_dst, F, generator=g).cuda()
    att, bias = (torch.randn(F, generator=g) * 0.3).cuda(), (torch.randn(F, generator=g) * 0.1).cuda()
    go = torch.randn(n_dst, F, generator=g).cuda()
    out, act, smax, sden, lg = ops.gatv2_fwd(x_l, x_r, att, bias, csr, H, C, 0.2, 0.2, True, 5, True, want_logits=True)
    assert lg is not None and tuple(lg.shape) == (csr.E, H)
    out0, act0, smax0, sden0 = ops.gatv2_fwd(x_l, x_r, att, bias, csr, H, C, 0.2, 0.2, True, 5, True)
    assert torch.equal(out, out0) and torch.equal(act, act0) and torch.equal(smax, smax0) and torch.equal(sden, sden0)
    # logits vs a direct evaluation (dst-CSR order: source = csr.col, destination = row of the position)
    dst_of = torch.repeat_interleave(torch.arange(n_dst, device="cuda"), (csr.rowptr[1:] - csr.rowptr[:-1]).long())
    z = torch.nn.functional.leaky_relu(x_l[csr.col.long()] + x_r[dst_of], 0.2).view(-1, H, C)
    ref_lg = (z * att.view(1, H, C)).sum(-1)
    assert rel_err(lg, ref_lg) < 1e-5
    a = ops.gatv2_bwd(x_l, x_r, att, bias, out, go, True, csr, H, C, 0.2, 0.2, True, 5, smax, sden, e_logit=lg)
    b = ops.gatv2_bwd(x_l, x_r, att, bias, out, go, True, csr, H, C, 0.2, 0.2, True, 5, smax, sden)
    for u, w in zip(a, b):
        assert rel_err(u, w) < 2e-6
    a2 = ops.gatv2_bwd(x_l, x_r, att, bias, out, go, True, csr, H, C, 0.2, 0.2, True, 5, smax, sden, e_logit=lg)
    assert all(torch.equal(u, w) for u, w in zip(a, a2))          # deterministic


def test_gatv2_logits_not_offered_outside_the_subwarp_kernels(monkeypatch):
    """Shapes / modes the sub-warp kernels do not cover return no logit buffer (the backward then recomputes)."""
    H, C = 5, 7
    x_l, x_r, att, bias, ei = _gat_case(100, 80, 600, H, C, seed=3)
    csr = ops.build_csr(ei.cuda(), 100, 80)
    r = ops.gatv2_fwd(x_l.cuda(), x_r.cuda(), att.cuda(), bias.cuda(), csr, H, C, 0.2, 0.0, False, 0, False, want_logits=True)
    assert len(r) == 5 and r[4] is None
    monkeypatch.setenv("SEGGER_B200_GAT_LOGITS", "0")
    x_l, x_r, att, bias, ei = _gat_case(100, 80, 600, 2, 64, seed=3)
    csr = ops.build_csr(ei.cuda(), 100, 80)
    r = ops.gatv2_fwd(x_l.cuda(), x_r.cuda(), att.cuda(), bias.cuda(), csr, 2, 64, 0.2, 0.0, False, 0, False, want_logits=True)
    assert r[4] is None
