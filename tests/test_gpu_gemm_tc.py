"""tcgen05 3xTF32 GEMM parity: forward / dgrad / wgrad against float64 torch, and against the exact
fp32 SIMT back end.  Tolerance 1e-5 relative to the output scale (3xTF32 carries ~2^-21)."""
import os

import pytest
import torch

from segger_b200 import ops
from segger_b200._lib import ACT_GELU, ACT_NONE, ACT_SILU
from tests.util import rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-5

SHAPES = [  # (M, N, K)
    (128, 64, 32), (128, 128, 32), (128, 192, 32), (128, 256, 32), (128, 256, 64),
    (300, 384, 256), (1000, 64, 128), (4096, 128, 256), (5000, 384, 128), (777, 100, 36), (129, 72, 40),
    (20000, 64, 256), (20000, 256, 384), (333, 512, 128), (64, 16, 8),
    (50000, 64, 12), (64, 12, 256),      # low-rank positional front end: K = 12 basis columns / N = 12 reduction rows
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_forward(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    ref = x.double() @ w.double().t() + b.double()
    y, _ = ops.linear_fwd(x.cuda(), w.cuda(), b.cuda(), exact=False)
    assert rel_err(y, ref) < TOL
    y2, a2 = ops.linear_fwd(x.cuda(), w.cuda(), b.cuda(), ACT_GELU, exact=False)
    assert rel_err(a2, torch.nn.functional.gelu(ref)) < TOL


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_dgrad(M, N, K):
    g = torch.Generator().manual_seed(M + 2 * N + K)
    dy, w = torch.randn(M, N, generator=g), torch.randn(N, K, generator=g)
    ref = dy.double() @ w.double()
    dx = ops.linear_dgrad(dy.cuda(), w.cuda())
    assert rel_err(dx, ref) < TOL
    pre = torch.randn(M, K, generator=g)
    base = torch.randn(M, K, generator=g)
    dx2 = base.clone().cuda()
    ops.linear_dgrad(dy.cuda(), w.cuda(), dx=dx2, accumulate=True)
    assert rel_err(dx2, ref + base.double()) < TOL
    dx3 = ops.linear_dgrad(dy.cuda(), w.cuda(), act=ACT_SILU, act_pre=pre.cuda())
    s = torch.sigmoid(pre.double())
    assert rel_err(dx3, ref * (s * (1 + pre.double() * (1 - s)))) < TOL


@pytest.mark.parametrize("M,N,K", SHAPES + [(200000, 384, 256), (100000, 64, 64)])
def test_tc_wgrad(M, N, K):
    g = torch.Generator().manual_seed(M + N + 3 * K)
    dy, x = torch.randn(M, N, generator=g), torch.randn(M, K, generator=g)
    ref = dy.double().t() @ x.double()
    dw, db = ops.linear_wgrad(dy.cuda(), x.cuda())
    # long reductions accumulate fp32 rounding (in TMEM and across the split-K partials)
    assert rel_err(dw, ref) < (TOL if M <= 20000 else 1e-4)
    assert rel_err(db, dy.double().sum(0)) < 1e-4
    dw2, _ = ops.linear_wgrad(dy.cuda(), x.cuda())
    assert torch.equal(dw, dw2)                       # deterministic split-K


def test_exact_flag_selects_round_to_nearest_fp32_path():
    """exact=2 must give the fp32 SIMT result, exact=True/False the split-TF32 tensor-core one; both within
    fp32 rounding of the fp64 product."""
    g = torch.Generator().manual_seed(5)
    x, w = torch.randn(4096, 256, generator=g), torch.randn(384, 256, generator=g)
    ref = x.double() @ w.double().t()
    ys, _ = ops.linear_fwd(x.cuda(), w.cuda(), None, exact=2)
    ye, _ = ops.linear_fwd(x.cuda(), w.cuda(), None, exact=True)
    yt, _ = ops.linear_fwd(x.cuda(), w.cuda(), None, exact=False)
    assert rel_err(ys, ref) < 1.5e-6
    assert rel_err(ye, ref) < 1.5e-6
    assert rel_err(yt, ref) < 1.5e-6
    assert not torch.equal(ys, yt)


def test_tc_strided_operands_and_views():
    g = torch.Generator().manual_seed(0)
    big = torch.randn(3000, 3 * 128, generator=g).cuda()
    w = torch.randn(64, 128, generator=g).cuda()
    x = big[:, 128:256]                                # column slice, ld = 384
    y, _ = ops.linear_fwd(x, w, None, exact=False)
    assert rel_err(y, x.double() @ w.double().t()) < TOL
    out = torch.zeros(3000, 256, device="cuda")
    ops.linear_fwd(x, w, None, y=out[:, 64:128], exact=False)
    assert rel_err(out[:, 64:128], y) == 0 and float(out[:, :64].abs().max()) == 0 and float(out[:, 128:].abs().max()) == 0


@pytest.mark.parametrize("M,N,K", [(50000, 64, 12), (8191, 100, 16), (4096, 256, 4), (20001, 32, 8)])
def test_skinny_k_forward_and_wgrad(M, N, K):
    """Reduction depth K <= 16 (the low-rank positional front end): streaming fp32 kernels, forward with fused
    SiLU and weight + bias gradient in one deterministic pass."""
    g = torch.Generator().manual_seed(M + N + K)
    x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    ref = x.double() @ w.double().t() + b.double()
    y, a = ops.linear_fwd(x.cuda(), w.cuda(), b.cuda(), ACT_SILU)
    assert rel_err(y, ref) < 2e-6
    assert rel_err(a, torch.nn.functional.silu(ref)) < 2e-6
    dy = torch.randn(M, N, generator=g)
    dw, db = ops.linear_wgrad(dy.cuda(), x.cuda())
    assert rel_err(dw, dy.double().t() @ x.double()) < 1e-5
    assert rel_err(db, dy.double().sum(0)) < 1e-5
    dw2, db2 = ops.linear_wgrad(dy.cuda(), x.cuda())
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
