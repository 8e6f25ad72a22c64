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
