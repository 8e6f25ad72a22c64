"""B200-native drop-in for the hot-path methods of ``segger.models.lightning_model.LitISTEncoder``
(/root/reference/src/segger/models/lightning_model.py): ``forward`` (:127-134), ``predict_step``
(:263-298) and ``configure_optimizers`` (:300-303).  The losses (``get_losses``, :151-213) are
callers of the hot path and stay out of scope.

Lightning is optional: when it is importable the class derives from ``LightningModule`` so it can
be handed to a ``Trainer``; otherwise it is a plain ``torch.nn.Module`` with the same methods.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .ist_encoder import ISTEncoder

try:  # pragma: no cover - lightning is not part of the build image
    from lightning import LightningModule as _Base
except Exception:  # noqa: BLE001
    _Base = torch.nn.Module

PRED = ("tx", "neighbors", "bd")


class LitISTEncoder(_Base):
    def __init__(self, n_genes: int, in_channels: int, hidden_channels: int = 64, out_channels: int = 64,
                 n_mid_layers: int = 2, n_heads: int = 2, learning_rate: float = 1e-3,
                 sg_loss_type: str = "triplet", tx_margin: float = 0.3, sg_margin: float = 0.4,
                 tx_weight_start: float = 1., tx_weight_end: float = 1., bd_weight_start: float = 1.,
                 bd_weight_end: float = 1., sg_weight_start: float = 0., sg_weight_end: float = 0.5,
                 update_gene_embedding: bool = True, use_positional_embeddings: bool = True,
                 normalize_embeddings: bool = True):
        super().__init__()
        if hasattr(self, "save_hyperparameters"):
            self.save_hyperparameters()
        self.model = ISTEncoder(
            n_genes=n_genes, in_channels=in_channels, hidden_channels=hidden_channels,
            out_channels=out_channels, n_mid_layers=n_mid_layers, n_heads=n_heads,
            normalize_embeddings=normalize_embeddings, use_positional_embeddings=use_positional_embeddings)
        self.learning_rate = learning_rate
        self._sg_loss_type = sg_loss_type
        self._tx_margin = tx_margin
        self._sg_margin = sg_margin
        self._w_start = torch.tensor([tx_weight_start, bd_weight_start, sg_weight_start])
        self._w_end = torch.tensor([tx_weight_end, bd_weight_end, sg_weight_end])
        self._freeze_gene_embedding = not update_gene_embedding

    def forward(self, batch) -> dict:
        """lightning_model.py:127-134."""
        return self.model(batch.x_dict, batch.edge_index_dict, batch.pos_dict, batch.batch_dict)

    def predict_step(self, batch, batch_idx: int = 0, min_similarity: Optional[float] = None):
        """lightning_model.py:263-298: embeddings -> cosine similarity over tx-neighbors-bd candidate
        edges -> per-transcript max / arg-max -> cell id (or -1); returns CPU tensors
        (tx.index, seg_idx, max_sim, gene id) restricted to ``predict_mask``."""
        embeddings = self.forward(batch)
        edge_index = batch[PRED].edge_index
        max_sim, max_idx, seg_idx = ops.score_argmax(
            embeddings["tx"], embeddings["bd"], edge_index, batch["bd"]["index"], min_similarity)
        src_idx = batch["tx"]["index"]
        gen_idx = batch["tx"]["x"]
        mask = batch["tx"]["predict_mask"]
        # To cpu, else gpu is held until end of predict loop (reference comment, :296).  Same four CPU tensors as
        # the reference's `x[mask].cpu()` x 4, produced with one mask scan, four gathers and four asynchronous
        # copies into pinned host memory behind a single stream synchronisation (instead of four
        # nonzero + sync + pageable-copy round trips).
        n_keep = int(mask.sum())
        idx = None if n_keep == mask.numel() else torch.nonzero(mask).squeeze(1)
        outs = []
        for t in (src_idx, seg_idx, max_sim, gen_idx):
            sel = t if idx is None else t.index_select(0, idx)
            host = torch.empty(sel.shape, dtype=sel.dtype, device="cpu", pin_memory=True)
            host.copy_(sel, non_blocking=True)
            outs.append(host)
        torch.cuda.current_stream(max_sim.device).synchronize()
        return tuple(outs)

    def configure_optimizers(self) -> torch.optim.Optimizer:
        return torch.optim.Adam(self.parameters(), lr=self.learning_rate)
