"""Minimal stand-in for the slice of ``torch_geometric.data.HeteroData`` / ``Batch`` that the hot
path touches (``batch.x_dict``, ``batch.edge_index_dict``, ``batch.pos_dict``, ``batch.batch_dict``,
``batch['tx']['index']``, ``batch['tx','neighbors','bd'].edge_index``, ``batch['tx'].num_nodes``;
see /root/reference/src/segger/models/lightning_model.py:127-134,263-298).  A real PyG ``Batch``
satisfies the same protocol, so the product modules accept either.
"""
from __future__ import annotations

from typing import Dict, Tuple, Union

import torch
from torch import Tensor


class _Store(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    @property
    def num_nodes(self) -> int:
        for key in ("x", "pos", "index"):
            if key in self:
                return int(self[key].size(0))
        raise AttributeError("num_nodes")


class HeteroBatch:
    """Node stores keyed by type ('tx', 'bd'), edge stores keyed by (src, rel, dst) tuples."""

    def __init__(self):
        self._nodes: Dict[str, _Store] = {}
        self._edges: Dict[Tuple[str, str, str], _Store] = {}

    def __getitem__(self, key: Union[str, Tuple[str, str, str]]) -> _Store:
        if isinstance(key, tuple):
            return self._edges.setdefault(key, _Store())
        return self._nodes.setdefault(key, _Store())

    def _collect(self, name: str) -> Dict:
        out = {k: s[name] for k, s in self._nodes.items() if name in s}
        out.update({k: s[name] for k, s in self._edges.items() if name in s})
        return out

    @property
    def x_dict(self): return {k: s["x"] for k, s in self._nodes.items() if "x" in s}
    @property
    def pos_dict(self): return {k: s["pos"] for k, s in self._nodes.items() if "pos" in s}
    @property
    def batch_dict(self): return {k: s["batch"] for k, s in self._nodes.items() if "batch" in s}
    @property
    def edge_index_dict(self): return {k: s["edge_index"] for k, s in self._edges.items() if "edge_index" in s}
    @property
    def num_graphs(self) -> int:
        """Number of tiles in the batch (PyG ``Batch.num_graphs``): set by the collate, else 1."""
        return int(getattr(self, "_num_graphs", 1))

    @property
    def node_types(self): return list(self._nodes)
    @property
    def edge_types(self): return list(self._edges)

    def to(self, device, non_blocking: bool = False) -> "HeteroBatch":
        out = HeteroBatch()
        if hasattr(self, "_num_graphs"):
            out._num_graphs = self._num_graphs
        for k, s in self._nodes.items():
            for n, v in s.items():
                out[k][n] = v.to(device, non_blocking=non_blocking) if isinstance(v, Tensor) else v
        for k, s in self._edges.items():
            for n, v in s.items():
                out[k][n] = v.to(device, non_blocking=non_blocking) if isinstance(v, Tensor) else v
        return out

    def cuda(self): return self.to("cuda")

    def pin_memory(self) -> "HeteroBatch":
        out = HeteroBatch()
        for k, s in list(self._nodes.items()) + list(self._edges.items()):
            for n, v in s.items():
                out[k][n] = v.pin_memory() if isinstance(v, Tensor) else v
        return out

    def nbytes(self) -> int:
        tot = 0
        for s in list(self._nodes.values()) + list(self._edges.values()):
            for v in s.values():
                if isinstance(v, Tensor):
                    tot += v.numel() * v.element_size()
        return tot


def from_synth(ts, edge_tt: Tensor, train_edges: bool = False) -> HeteroBatch:
    """Wrap a ``segger_b200.synth.SynthTileSet`` (+ a tx-neighbors-tx edge list) as a HeteroBatch
    laid out like ``setup_heterodata`` (data/utils/heterodata.py:114-162) after batching."""
    import numpy as np
    b = HeteroBatch()
    b["tx"]["x"] = torch.from_numpy(ts.tx_gene)
    b["tx"]["pos"] = torch.from_numpy(ts.tx_pos)
    b["tx"]["batch"] = torch.from_numpy(ts.tx_tile)
    b["tx"]["index"] = torch.from_numpy(ts.tx_index)
    b["tx"]["predict_mask"] = torch.ones(ts.tx_gene.shape[0], dtype=torch.bool)
    b["bd"]["x"] = torch.from_numpy(ts.bd_x)
    b["bd"]["pos"] = torch.from_numpy(ts.bd_pos)
    b["bd"]["batch"] = torch.from_numpy(ts.bd_tile)
    b["bd"]["index"] = torch.from_numpy(ts.bd_index)
    b["tx", "neighbors", "tx"]["edge_index"] = edge_tt
    b["tx", "belongs", "bd"]["edge_index"] = torch.from_numpy(ts.edge_tb)
    b["tx", "neighbors", "bd"]["edge_index"] = torch.from_numpy(ts.edge_pred)
    return b
