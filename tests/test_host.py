"""CPU tests of the host side (no GPU): C-ABI library exports, header/ctypes agreement, drop-in module
structure (state-dict keys, lazy parameters, tuple-key ModuleDict), containers, and the fail-loud
behaviour of the product path without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from segger_b200 import _lib
from segger_b200.hetero import HeteroBatch, from_synth
from segger_b200.ist_encoder import BT, TB, TT, ISTEncoder, Positional2dEmbedder, SkipGAT
from segger_b200.nn import GATv2Conv, HeteroConv, HeteroDictLinear, Linear, ModuleDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "segger_b200.h")).read()
    return re.findall(r"SGB_API\s+[\w\s\*]+?\b(sgb_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol(lib):
    names = _header_functions()
    assert len(names) >= 30
    raw = ctypes.CDLL(str(_lib.lib_path()))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/segger_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.sgb_version() == 202


def test_ctypes_prototypes_match_header_arity():
    src = open(os.path.join(ROOT, "include", "segger_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, (_, args) in _lib._PROTOS.items():
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(args), f"{name}: header has {n} parameters, ctypes binding {len(args)}"


def test_abi_argument_validation_without_gpu(lib):
    # pure host-side argument checks return error codes (no CUDA call is reached)
    assert lib.sgb_csr_build(None, 3, 0, 1, 0, 1, 1, None, None, None, None, None, None, None, None, 0, None) == -1
    assert b"idx_bytes" in lib.sgb_last_error()
    assert lib.sgb_csr_build(None, 8, 0, 1, 1 << 31, 1, 1, None, None, None, None, None, None, None, None, 0, None) == -3
    assert lib.sgb_dropout_mask(0, 10, 2, 1.5, None, None) == -1
    assert lib.sgb_csr_workspace_bytes(1000) > 5 * 4000
    assert lib.sgb_gatv2_bwd_workspace_bytes(100, 1000, 2, 64) >= 2 * 1000 * 2 * 4


def test_product_path_fails_loudly_on_cpu_tensors():
    from segger_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.build_csr(torch.zeros(2, 3, dtype=torch.long), 4, 4)
    m = ISTEncoder(10, in_channels=16, hidden_channels=8, out_channels=8, n_mid_layers=0, n_heads=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m({"tx": torch.zeros(3, dtype=torch.int32), "bd": torch.zeros(2, 4)}, {}, {}, {})


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setenv("SEGGER_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU or eager fallback"):
        _lib.load()


def test_moduledict_tuple_keys_mangled_like_pyg():
    md = ModuleDict({TT: torch.nn.Identity(), "tx": torch.nn.Identity(), "type": torch.nn.Identity()})
    assert list(torch.nn.ModuleDict.keys(md)) == ["<tx___neighbors___tx>", "tx", "<type>"]
    assert TT in md and "tx" in md and ("a", "b", "c") not in md
    assert md.keys() == [TT, "tx", "type"]
    # __iter__ yields the INTERNAL string keys -> the reference's attention kwarg is a no-op (Appendix B.2)
    assert [k for k in md] == ["<tx___neighbors___tx>", "tx", "<type>"]
    assert all(not isinstance(k, tuple) for k in md)


def test_istencoder_state_dict_keys_and_hparams():
    m = ISTEncoder(n_genes=50, in_channels=128, hidden_channels=64, out_channels=64, n_mid_layers=2, n_heads=2)
    assert m.hparams == dict(n_genes=50, in_channels=128, hidden_channels=64, out_channels=64, n_mid_layers=2,
                             n_heads=2, normalize_embeddings=True, use_positional_embeddings=True)
    keys = list(m.state_dict().keys())
    expect = ["lin_first.tx.weight", "lin_first.bd.weight", "lin_first.bd.bias", "pos_emb.mlp.0.weight",
              "pos_emb.mlp.0.bias", "pos_emb.mlp.2.weight", "pos_emb.mlp.2.bias"]
    for layer in range(4):
        for et in ("<tx___neighbors___tx>", "<tx___belongs___bd>", "<bd___contains___tx>"):
            for leaf in ("att", "bias", "lin_l.weight", "lin_l.bias", "lin_r.weight", "lin_r.bias"):
                expect.append(f"conv_layers.{layer}.conv.convs.{et}.{leaf}")
    expect += ["lin_last.lins.tx.weight", "lin_last.lins.tx.bias", "lin_last.lins.bd.weight", "lin_last.lins.bd.bias"]
    assert sorted(keys) == sorted(expect)
    assert len(m.conv_layers) == 4 and m.pos_emb.dim == 64 and m.lin_first["tx"].weight.shape == (50, 128)
    tt = m.conv_layers[0].conv.convs[TT]
    assert tt.att.shape == (1, 2, 64) and tt.bias.shape == (128,) and tt.dropout == 0.2 and not tt.add_self_loops
    assert tt.negative_slope == 0.2 and not tt.initialized


def test_lazy_parameters_round_trip_through_state_dict():
    from torch.nn.parameter import UninitializedParameter
    a = ISTEncoder(20, in_channels=16, hidden_channels=8, out_channels=8, n_mid_layers=0, n_heads=2)
    sd = a.state_dict()
    assert isinstance(sd["lin_first.bd.weight"], UninitializedParameter)
    b = ISTEncoder(20, in_channels=16, hidden_channels=8, out_channels=8, n_mid_layers=0, n_heads=2)
    b.load_state_dict(sd)                                    # lazy -> lazy, strict
    assert isinstance(b.lin_first["bd"].weight, UninitializedParameter)
    # a materialised checkpoint (e.g. trained with PyG) loads into a fresh lazy model
    from oracle.ist_encoder_ref import ISTEncoderRef
    ref = ISTEncoderRef(20, 12, 16, 8, 8, 0, 2)
    missing, unexpected = b.load_state_dict(ref.state_dict(), strict=False)
    assert not unexpected and all("bd___contains___tx" in k for k in missing)
    assert b.lin_first["bd"].weight.shape == (16, 12) and b.lin_first["bd"].in_channels == 12
    assert torch.equal(b.conv_layers[1].conv.convs[TB].lin_r.weight, ref.conv_layers[1].conv.convs["<tx___belongs___bd>"].lin_r.weight)
    assert not b.conv_layers[0].conv.convs[BT].initialized  # the dead conv stays unmaterialised (Appendix B.1)
    c = ISTEncoder(20, in_channels=16, hidden_channels=8, out_channels=8, n_mid_layers=0, n_heads=2)
    c.load_state_dict(b.state_dict())                        # mixed lazy / materialised, strict
    assert torch.equal(c.lin_last.lins["tx"].weight, b.lin_last.lins["tx"].weight)


def test_unsupported_gatv2_options_raise():
    for kw in (dict(concat=False), dict(edge_dim=4), dict(share_weights=True), dict(residual=True)):
        with pytest.raises(NotImplementedError):
            GATv2Conv((-1, -1), 8, heads=2, **kw)
    with pytest.raises(ValueError):
        HeteroConv({TT: torch.nn.Identity()}, aggr="median")
    h = HeteroDictLinear(-1, 8, types=("tx", "bd"))
    assert list(h.state_dict().keys()) == ["lins.tx.weight", "lins.tx.bias", "lins.bd.weight", "lins.bd.bias"]
    assert Linear(4, 3, bias=False).bias is None


def test_hetero_batch_protocol_and_synth_layout():
    from segger_b200.synth import drop_cross_tile_edges, synth
    ts = synth(4000, 40, seed=3, nodes_per_tile=1000)
    assert ts.tx_pos.dtype == np.float32 and ts.tx_gene.dtype == np.int32 and ts.edge_pred.dtype == np.int32
    assert ts.n_tiles == 4 and np.all(np.diff(ts.tx_tile) >= 0)              # tile-major node order
    assert ts.edge_tb.shape[1] == 40 * 40 and np.all(ts.tx_compartment[ts.edge_tb[0]] == 2)
    assert np.all(ts.tx_cell[ts.edge_tb[0]] == ts.edge_tb[1])
    d = np.linalg.norm(ts.tx_pos[ts.edge_pred[0]] - ts.bd_pos[ts.edge_pred[1]], axis=1)
    assert d.max() < 6.5 * 1.05 and np.bincount(ts.edge_pred[0], minlength=4000).max() <= 3
    ei = np.stack([np.arange(3999), np.arange(1, 4000)])
    kept = drop_cross_tile_edges(ei, ts.tx_tile, ts.tx_tile)
    assert kept.shape[1] == 3999 - 3
    b = from_synth(ts, torch.from_numpy(ei))
    assert set(b.x_dict) == {"tx", "bd"} and set(b.edge_index_dict) == {TT, TB, ("tx", "neighbors", "bd")}
    assert b["tx"].num_nodes == 4000 and b["bd"].num_nodes == 40 and b[TT].edge_index.shape == (2, 3999)
    assert b.nbytes() > 0 and set(b.batch_dict) == {"tx", "bd"}


def test_synth_is_deterministic():
    from segger_b200.synth import synth
    a, b = synth(3000, 30, seed=5), synth(3000, 30, seed=5)
    assert np.array_equal(a.tx_pos, b.tx_pos) and np.array_equal(a.edge_pred, b.edge_pred)
    assert not np.array_equal(a.tx_pos, synth(3000, 30, seed=6).tx_pos)


def test_chebyshev_coefficient_matrix_reproduces_oracle_sinusoid():
    """Host-side constant of the low-rank positional front end: T(2p-1) @ M equals the oracle's
    sinusoidal_embedding (ist_encoder.py:22-31, max_period 10000, dim 256) for p in [0,1] to fp32 rounding."""
    from oracle.ist_encoder_ref import sinusoidal_embedding
    from segger_b200 import ops
    freqs = ops.sinusoid_freqs(256, 10000, "cpu")
    M = ops.cheb_feature_matrix(freqs)
    assert M.shape == (ops.CHEB_DEG, 256) and M.dtype == torch.float32
    p = torch.cat([torch.linspace(0, 1, 4001, dtype=torch.float64), torch.tensor([0.0, 1.0, 0.5], dtype=torch.float64)])
    t = 2 * p - 1
    T = torch.stack([torch.cos(n * torch.acos(t.clamp(-1, 1))) for n in range(ops.CHEB_DEG)], 1)
    ref = sinusoidal_embedding(p.float(), 256, max_period=10000).double()
    assert float((T @ M.double() - ref).abs().max()) < 3e-7
    with pytest.raises(ValueError):
        ops.cheb_feature_matrix(freqs * 50.0)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours): stdout is one JSON line with the contract's
    keys; nothing else may reach stdout (C libraries write their banners to fd 1)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0


def test_geometry_host_helpers_pack_and_grid():
    """pack_rings / PackedPolygons grid (host side of the point-in-polygon join): offsets, tolerated empty rings, a grid
    that covers every polygon box with cells of about one polygon extent."""
    from segger_b200.geometry import PackedPolygons, pack_rings
    rng = np.random.default_rng(0)
    ang = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    centers = rng.uniform(0, 500, (300, 2))
    rings = [np.stack([x + 7 * np.cos(ang), y + 7 * np.sin(ang)], 1) for x, y in centers]
    rings.insert(5, np.zeros((0, 2)))
    verts, off = pack_rings(rings)
    assert verts.shape == (300 * 16, 2) and off[0] == 0 and off[-1] == 300 * 16 and off[6] == off[5]
    p = PackedPolygons(verts, off)
    xmin, ymin, cell, nx, ny = p.grid
    assert p.n_poly == 301 and 13.0 < cell < 15.0
    assert xmin <= verts[:, 0].min() and ymin <= verts[:, 1].min()
    assert xmin + nx * cell > verts[:, 0].max() and ymin + ny * cell > verts[:, 1].max()
    assert nx * ny < 2 ** 26
    empty = PackedPolygons(*pack_rings([]))
    assert empty.grid is None and empty.n_poly == 0


def test_loss_modules_reject_unsupported_options_without_gpu():
    from segger_b200.triplet_loss import MetricLoss, TripletLoss
    sim = torch.eye(4)
    with pytest.raises(ValueError):
        TripletLoss(sim.clone(), margin=0.0)
    with pytest.raises(NotImplementedError):
        TripletLoss(sim.clone(), margin=0.3, swap=True)
    lt, lm = TripletLoss(sim.clone(), margin=0.3), MetricLoss(sim.clone())
    assert lt.forward(torch.zeros(0, 8), torch.zeros(0, dtype=torch.long)) == 0.
    assert lm.forward(torch.zeros(0, 8), torch.zeros(0, dtype=torch.long)) == 0.
    with pytest.raises(Exception):                      # CPU tensors: the product path has no CPU fallback
        lt.forward(torch.randn(10, 8), torch.randint(0, 4, (10,)))


def test_setup_gene_embedding_and_datamodule_contract():
    """LitISTEncoder.setup (lightning_model.py:86-125): pretrained gene embedding honours update_gene_embedding, and a
    data module without the similarity matrices is an error at setup time, not an AttributeError in get_losses."""
    from types import SimpleNamespace
    from segger_b200.lightning_model import LitISTEncoder
    w = torch.arange(12, dtype=torch.float64).reshape(4, 3)
    for update in (True, False):
        lit = LitISTEncoder(4, in_channels=3, n_mid_layers=0, update_gene_embedding=update)
        lit.setup_gene_embedding(w.numpy())
        emb = lit.model.lin_first["tx"]
        assert isinstance(emb, torch.nn.Embedding) and emb.weight.dtype == torch.float32
        assert torch.equal(emb.weight.detach(), w.float()) and emb.weight.requires_grad is update
    lit = LitISTEncoder(4, in_channels=3, n_mid_layers=0, update_gene_embedding=False)
    sim = torch.eye(3)
    lit.trainer = SimpleNamespace(datamodule=SimpleNamespace(gene_embedding=w, tx_similarity=sim.clone(),
                                                             bd_similarity=sim.clone()), max_epochs=2)
    lit.setup("fit")
    assert not lit.model.lin_first["tx"].weight.requires_grad and hasattr(lit, "loss_tx") and hasattr(lit, "loss_bd")
    lit.trainer = SimpleNamespace(datamodule=SimpleNamespace(tx_similarity=sim.clone()), max_epochs=2)
    with pytest.raises(TypeError):
        lit.setup("fit")
    lit.trainer = SimpleNamespace(datamodule=None, max_epochs=2)
    with pytest.raises(TypeError):
        lit.setup("fit")


def test_graph_setup_drop_ins_accept_frames_and_validate_modes():
    from segger_b200 import neighbors as nb
    xy = np.random.default_rng(0).uniform(0, 9, (7, 2)).astype(np.float32)
    assert np.array_equal(nb._xy({"x": xy[:, 0], "y": xy[:, 1]}), xy)
    import pandas as pd
    assert np.array_equal(nb._xy(pd.DataFrame({"x": xy[:, 0], "y": xy[:, 1], "g": 0})), xy)
    assert nb._xy(xy) is xy
    with pytest.raises(ValueError):
        nb.setup_prediction_graph(xy, (np.zeros((0, 2)), np.zeros(1, dtype=np.int64)), 3, 0.0, mode="tile")
    with pytest.raises(ValueError):          # packed outlines are already buffered
        nb.setup_prediction_graph(xy, (np.zeros((0, 2)), np.zeros(1, dtype=np.int64)), 3, 0.1, mode="cell")
