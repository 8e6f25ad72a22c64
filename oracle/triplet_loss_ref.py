"""CPU restatement of segger's training losses -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/segger/models/triplet_loss.py (FastTripletSelector :8-125, TripletLoss :128-160,
MetricLoss :163-204) and the loss assembly of LitISTEncoder.get_losses / _scheduled_weights
(/root/reference/src/segger/models/lightning_model.py:136-211) with plain torch CPU ops.

Two documented deviations, both where the reference's own result is unspecified:
  * the reference draws its four uniform vectors with torch.rand on the embedding's device (:93,:99,:105,:112);
    here they can be injected (``uniforms=``) so that product and oracle sample from identical numbers;
  * the reference orders the members of a cluster with torch.argsort(labels) (:41), which is not stable; here the
    sort is stable (members in index order), the product does the same.  Which member a uniform number selects
    inside a cluster is therefore pinned only up to that ordering -- the distribution is identical.
PARITY UNPINNED in the sense of the task statement: the reference ships no tests or golden vectors for this path
and cannot be imported here (torch_geometric / lightning absent); the restatement is line-by-line torch.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
from torch import Tensor
from torch.nn import functional as F


class FastTripletSelectorRef:
    """triplet_loss.py:8-125."""

    def __init__(self, cluster_similarity: Tensor):
        cluster_similarity = cluster_similarity.clone()
        cluster_similarity.fill_diagonal_(1)
        self.similarity = cluster_similarity.clamp_min(1e-8)            # :21-23
        self.dissimilarity = (-cluster_similarity).clamp_min(1e-8)      # :24

    def build_index(self, labels: Tensor, sorted_idx: Optional[Tensor] = None):
        """:27-86.  Returns the tuple the sampler reads.  ``sorted_idx``: inject the member order the reference's
        (unstable) ``torch.argsort(labels)`` produced (:41), to compare everything downstream bit for bit."""
        C = self.similarity.size(0)
        counts = torch.bincount(labels, minlength=C).to(torch.long)
        offsets = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])[:-1]
        if sorted_idx is None:
            sorted_idx = torch.argsort(labels, stable=True)
        present = torch.nonzero(counts > 0, as_tuple=False).flatten()
        diss = self.dissimilarity[present][:, present]
        cdf_neg = torch.cumsum(diss / diss.sum(dim=1, keepdim=True), dim=1)
        cdf_neg[:, -1] = 1.0
        sim = self.similarity[present][:, present]
        cdf_pos = torch.cumsum(sim / sim.sum(dim=1, keepdim=True), dim=1)
        cdf_pos[:, -1] = 1.0
        present_idx = -torch.ones(C, dtype=torch.long)
        present_idx[present] = torch.arange(present.numel())
        return counts, offsets, sorted_idx, present, cdf_pos, cdf_neg, present_idx

    def sample_triplets(self, labels: Tensor, uniforms: Optional[Sequence[Tensor]] = None,
                        sorted_idx: Optional[Tensor] = None):
        """:88-125 -> (positives, negatives, dists_pos, dists_neg)."""
        counts, offsets, sorted_idx, present, cdf_pos, cdf_neg, present_idx = self.build_index(labels, sorted_idx)
        N = labels.numel()
        u_pos, u2, u_neg, u3 = uniforms if uniforms is not None else [torch.rand(N) for _ in range(4)]
        pres_idx = present_idx[labels]
        pos_pres = torch.searchsorted(cdf_pos[pres_idx], u_pos.unsqueeze(-1)).squeeze(-1)
        pos_clust = present[pos_pres]
        pos_pos = (u2 * counts[pos_clust].float()).floor().to(torch.long)
        positives = sorted_idx[offsets[pos_clust] + pos_pos]
        neg_pres = torch.searchsorted(cdf_neg[pres_idx], u_neg.unsqueeze(-1)).squeeze(-1)
        neg_clust = present[neg_pres]
        neg_pos = (u3 * counts[neg_clust].float()).floor().to(torch.long)
        negatives = sorted_idx[offsets[neg_clust] + neg_pos]
        dists = 1.0 - self.similarity
        return positives, negatives, dists[labels, labels[positives]], dists[labels, labels[negatives]]


def triplet_margin_ref(anchor: Tensor, positive: Tensor, negative: Tensor, margin: float) -> Tensor:
    """torch.nn.TripletMarginLoss(margin) defaults (p=2, eps=1e-6, swap=False, mean): triplet_loss.py:128-160 and
    lightning_model.py:116,181-186."""
    return F.triplet_margin_loss(anchor, positive, negative, margin=margin)


def triplet_loss_ref(embeddings: Tensor, positives: Tensor, negatives: Tensor, margin: float) -> Tensor:
    """TripletLoss.forward after sampling (:154-160)."""
    if embeddings.size(0) == 0:
        return embeddings.new_zeros(())
    return triplet_margin_ref(embeddings, embeddings[positives], embeddings[negatives], margin)


def metric_loss_ref(embeddings: Tensor, positives: Tensor, negatives: Tensor, dists_pos: Tensor,
                    dists_neg: Tensor) -> Tensor:
    """MetricLoss.forward after sampling (:193-204)."""
    if embeddings.size(0) == 0:
        return embeddings.new_zeros(())
    cos_pos = torch.cosine_similarity(embeddings, embeddings[positives])
    cos_neg = torch.cosine_similarity(embeddings, embeddings[negatives])
    return (F.mse_loss(cos_pos, 1 - dists_pos.to(cos_pos.dtype), reduction="mean")
            + F.mse_loss(cos_neg, 1 - dists_neg.to(cos_neg.dtype), reduction="mean"))


def segmentation_loss_ref(emb_tx: Tensor, emb_bd: Tensor, edge_index: Tensor, dst_neg: Tensor, kind: str,
                          margin: float) -> Tensor:
    """lightning_model.py:163-205 with the random negative destinations injected (:176-178 draws them with
    torch.randint)."""
    src_pos, dst_pos = edge_index[0].long(), edge_index[1].long()
    if emb_bd.size(0) <= 1:
        return emb_bd.new_zeros(())
    if kind == "triplet":
        return triplet_margin_ref(emb_tx[src_pos], emb_bd[dst_pos], emb_bd[dst_neg], margin)
    src = torch.cat([src_pos, src_pos])
    dst = torch.cat([dst_pos, dst_neg])
    logits = (emb_tx[src] * emb_bd[dst]).sum(dim=-1)
    labels = torch.cat([torch.ones(src_pos.numel()), torch.zeros(src_pos.numel())]).to(logits.dtype)
    return F.binary_cross_entropy_with_logits(logits, labels)


def scheduled_weights_ref(w_start: Tensor, w_end: Tensor, current_epoch: int, max_epochs: int,
                          normalize: bool = True) -> Tensor:
    """lightning_model.py:136-149 (cosine ramp)."""
    m = max(1, max_epochs - 1)
    t = min(current_epoch, m) / m
    alpha = 0.5 * (1.0 + math.cos(math.pi * t))
    w = w_end + (w_start - w_end) * alpha
    if normalize:
        w = w / (w.sum() + 1e-8)
    return w
