"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
from __future__ import annotations

import numpy as np
import torch

from oracle.ist_encoder_ref import ISTEncoderRef, TB, TT
from segger_b200.ist_encoder import ISTEncoder

PRED = ("tx", "neighbors", "bd")


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max(|b|_inf, tiny): error relative to the tensor's scale (fp32, 1e-4 bar)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def random_graph(n_src, n_dst, E, seed, isolated_frac=0.1, dup=True, dtype=torch.int64):
    g = torch.Generator().manual_seed(seed)
    n_live = max(1, int(n_dst * (1 - isolated_frac)))
    src = torch.randint(0, n_src, (E,), generator=g)
    dst = torch.randint(0, n_live, (E,), generator=g)
    if dup and E >= 4:
        src[1], dst[1] = src[0], dst[0]          # duplicate edge
    return torch.stack([src, dst]).to(dtype)


def make_models(n_genes, bd_in, in_channels, hidden, out, n_mid, heads, seed=0, device="cuda", **kw):
    """Oracle + product ISTEncoder with identical weights (oracle state_dict copied into product)."""
    torch.manual_seed(seed)
    ref = ISTEncoderRef(n_genes, bd_in, in_channels, hidden, out, n_mid, heads, **kw)
    # att / bias get non-trivial values so the tests see them
    with torch.no_grad():
        for p in ref.parameters():
            if p.dim() == 1:
                p.uniform_(-0.2, 0.2)
    prod = ISTEncoder(n_genes, in_channels, hidden, out, n_mid, heads,
                      normalize_embeddings=kw.get("normalize_embeddings", True),
                      use_positional_embeddings=kw.get("use_positional_embeddings", True))
    missing, unexpected = prod.load_state_dict(ref.state_dict(), strict=False)
    assert not unexpected, unexpected
    assert all("bd___contains___tx" in k for k in missing), missing
    return ref, prod.to(device)


def synth_batch(n_tx, n_cells, seed=0, k=5, dist=5.0, train_edges=True):
    """(SynthTileSet, cpu dicts) with the scipy oracle's tx-neighbors-tx graph."""
    from oracle.neighbors_ref import kdtree_neighbors
    from segger_b200.synth import drop_cross_tile_edges, synth
    ts = synth(n_tx, n_cells, seed=seed)
    ei, _ = kdtree_neighbors(ts.tx_pos, k, dist)
    ei = ei.numpy()
    if train_edges:
        ei = drop_cross_tile_edges(ei, ts.tx_tile, ts.tx_tile)
    x = {"tx": torch.from_numpy(ts.tx_gene), "bd": torch.from_numpy(ts.bd_x)}
    pos = {"tx": torch.from_numpy(ts.tx_pos), "bd": torch.from_numpy(ts.bd_pos)}
    bat = {"tx": torch.from_numpy(ts.tx_tile), "bd": torch.from_numpy(ts.bd_tile)}
    edges = {TT: torch.from_numpy(np.ascontiguousarray(ei)), TB: torch.from_numpy(ts.edge_tb),
             PRED: torch.from_numpy(ts.edge_pred)}
    return ts, x, edges, pos, bat


def to_dev(d, device="cuda"):
    return {k: v.to(device) for k, v in d.items()}


class replay_random:
    """Replay recorded ``torch.rand`` / ``torch.randint`` results (tests/golden/make_reference_golden.py records the
    draws the reference made) so the product consumes the very same numbers, in the same order, on its own device."""

    def __init__(self, draws):
        self.draws = list(draws)

    def __enter__(self):
        self._rand, self._randint = torch.rand, torch.randint

        def take(kind, kwargs):
            assert self.draws, f"the product drew more random tensors than the reference ({kind})"
            k, t = self.draws.pop(0)
            assert k == kind, (k, kind)
            dev = kwargs.get("device")
            return t.to(dev) if dev is not None else t.clone()

        torch.rand = lambda *a, **k: take("rand", k)
        torch.randint = lambda *a, **k: take("randint", k)
        return self

    def __exit__(self, *exc):
        torch.rand, torch.randint = self._rand, self._randint
        return False


# ---------------------------------------------------------------------------------------------------------------------
# Forward / backward parity of the whole encoder against the oracle (fp32 bar 1e-4 relative, BASELINE north_star).
#
# Some reference gradients are themselves ill-conditioned in fp32: LeakyReLU has a kink, and when some z = x_l[j] +
# x_r[i] of a tx-neighbors-tx conv lies within an ulp of 0, ANY fp32 implementation may put it on the other side than
# the fp64 run, which moves one output channel of that conv's lin_l / lin_r gradient -- and, diluted, every gradient
# upstream of it -- by far more than 1e-4.  Such a gradient is not pinned by the fp32 reference itself.  The rule:
#   * every tensor is compared with an fp64 copy of the oracle; error < 1e-4 -> fine (the overwhelming majority);
#   * a tensor that exceeds 1e-4 must be ON THE ALLOW-LIST below and stay under min(its ceiling, 3 x the fp32 oracle's
#     own conditioning noise), where the noise is the larger of (fp32 oracle vs fp64 oracle) and (fp32 oracle with every
#     weight moved by half an ulp vs fp64 oracle);
#   * a tensor whose oracle noise is < 2.5e-5 must ALSO hold the flat 1e-4 against the fp32 oracle.
# The allow-list is exactly the parameters at or upstream of the tx-neighbors-tx attention logits; measured tables
# (which tensors needed it, how large) are in profiles/r2_grad_parity.json.
RELAXED = {
    "<tx___neighbors___tx>.lin_r.weight": 5e-2, "<tx___neighbors___tx>.lin_r.bias": 5e-2,
    "<tx___neighbors___tx>.lin_l.weight": 5e-3, "<tx___neighbors___tx>.lin_l.bias": 5e-3,
    "<tx___neighbors___tx>.att": 5e-3, "<tx___neighbors___tx>.bias": 5e-3,
    "<tx___belongs___bd>.": 5e-3,            # layers above the first see the tt conv's output through tx features
    "lin_first.": 5e-3, "pos_emb.": 5e-3,
}
PARITY_TOL = 1e-4


def _ceiling(name):
    c = [v for k, v in RELAXED.items() if k in name]
    return max(c) if c else None


def _loss(out, g):
    return sum((out[k] * g[k]).sum() for k in ("tx", "bd"))


def check_forward_backward(ref, prod, x, edges, pos, bat, tag, grad_scale=1.0, noise_trials=3):
    """Outputs within 1e-4 of the fp32 oracle; gradients by the rule above.  Writes gpurun_out/<tag>.json."""
    import copy
    import json
    import os
    ref.eval(); prod.eval()
    out_r = ref(x, edges, pos, bat)
    gen = torch.Generator().manual_seed(0)
    g = {k: torch.randn(v.shape, generator=gen) * grad_scale for k, v in out_r.items()}
    _loss(out_r, g).backward()
    r64 = copy.deepcopy(ref).double()
    r64.zero_grad()
    x64 = {"tx": x["tx"], "bd": x["bd"].double()}
    _loss(r64(x64, edges, {k: v.double() for k, v in pos.items()}, bat), {k: v.double() for k, v in g.items()}).backward()
    grads_64 = {n: p.grad for n, p in r64.named_parameters()}
    out_p = prod(to_dev(x), to_dev(edges), to_dev(pos), to_dev(bat))
    _loss(out_p, to_dev(g)).backward()
    for k in ("tx", "bd"):
        assert out_p[k].shape == out_r[k].shape
        assert rel_err(out_p[k], out_r[k]) < PARITY_TOL, k
    ref_grads = {n: p.grad for n, p in ref.named_parameters()}
    gen = torch.Generator().manual_seed(1234)
    ulp_noise = {}
    for _ in range(noise_trials):
        r = copy.deepcopy(ref)
        r.zero_grad()
        with torch.no_grad():
            for p in r.parameters():
                p.mul_(1 + 2.0 ** -24 * torch.randn(p.shape, generator=gen))
        _loss(r(x, edges, pos, bat), g).backward()
        for n, p in r.named_parameters():
            ulp_noise[n] = max(ulp_noise.get(n, 0.0), rel_err(p.grad, grads_64[n]))
    table, relaxed, checked = {}, {}, 0
    for n, p in prod.named_parameters():
        if "bd___contains___tx" in n:
            continue
        assert p.grad is not None, n
        noise = max(rel_err(ref_grads[n], grads_64[n]), ulp_noise[n])
        e64, e32 = rel_err(p.grad, grads_64[n]), rel_err(p.grad, ref_grads[n])
        table[n] = {"err_vs_fp64": e64, "err_vs_fp32_oracle": e32, "oracle_noise": noise}
        if e64 >= PARITY_TOL:
            relaxed[n] = e64
            ceil = _ceiling(n)
            assert ceil is not None, (n, e64, "not on the RELAXED allow-list")
            assert e64 < min(ceil, max(PARITY_TOL, 3 * noise)), (n, e64, noise, ceil)
        if noise < PARITY_TOL / 4:
            assert e32 < PARITY_TOL, n
        checked += 1
    assert checked == len(ref_grads)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", tag + ".json"), "w") as f:
        json.dump({"tag": tag, "tolerance": PARITY_TOL, "n_tensors": checked, "relaxed": relaxed, "table": table}, f, indent=1)
    print(f"{tag}: {len(relaxed)} of {checked} gradient tensors used the relaxed bar: {relaxed}")
    return out_r, out_p
