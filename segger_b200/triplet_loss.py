"""B200-native drop-in for ``segger.models.triplet_loss``
(/root/reference/src/segger/models/triplet_loss.py): ``FastTripletSelector``, ``TripletLoss``, ``MetricLoss`` with
the reference's constructor signatures and ``forward(embeddings, labels)`` contract, plus the two segmentation
losses ``LitISTEncoder.get_losses`` applies to tx-belongs-bd edges (lightning_model.py:163-205).

Sampling, distances, hinge / MSE / BCE and their means run in libsegger_b200 kernels (``sgb_loss.cu``); the
backward writes per-item row gradients and segment-sums them by target row with the deterministic embedding
backward (no atomics).  The four uniform vectors are drawn with ``torch.rand`` exactly where the reference draws
them, so ``torch.manual_seed`` governs sampling as before.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib, ops
from ._lib import check, ptr, require_cuda, stream_ptr


def _i64(t: Tensor) -> Tensor:
    return t.to(torch.int64).contiguous()


def _segment_rows(g: Tensor, idx: Optional[Tensor], n_rows: int) -> Tensor:
    """sum_k g[k] into row idx[k] of an [n_rows, D] tensor (deterministic); identity when idx is None."""
    if idx is None:
        return g
    lib = _lib.load()
    T, D = g.shape
    out = torch.empty(n_rows, D, dtype=torch.float32, device=g.device)
    ws = ops._ws(lib.sgb_embedding_bwd_workspace_bytes(T, D, n_rows), g.device)
    check(lib.sgb_embedding_bwd(ptr(g), D, ptr(idx), 8, T, D, n_rows, None, 0, ptr(out), ptr(ws), ws.numel(),
                                stream_ptr(g.device)), "segment_rows")
    ops._count(10)
    return out


class _TripletMarginFn(torch.autograd.Function):
    """mean(max(margin + |a - p + eps| - |a - n + eps|, 0)) over rows gathered from three tables."""

    @staticmethod
    def forward(ctx, ta, tp, tn, ia, ip, in_, margin, eps):
        require_cuda(ta, tp, tn)
        ta, tp, tn = ops._rowmajor(ta), ops._rowmajor(tp), ops._rowmajor(tn)
        T = ia.numel() if ia is not None else ta.size(0)
        D = ta.size(1)
        dev = ta.device
        lib = _lib.load()
        d_ap = torch.empty(T, dtype=torch.float32, device=dev)
        d_an = torch.empty(T, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = ops._ws(lib.sgb_loss_workspace_bytes(T), dev)
        check(lib.sgb_triplet_margin_fwd(ptr(ta), ops._ld(ta), ptr(ia), ptr(tp), ops._ld(tp), ptr(ip), ptr(tn), ops._ld(tn),
                                         ptr(in_), T, D, float(margin), float(eps), ptr(d_ap), ptr(d_an), ptr(loss),
                                         ptr(ws), ws.numel(), stream_ptr(dev)), "triplet_margin_fwd")
        ops._count(2)
        ctx.cfg = (T, D, float(margin), float(eps))
        # TripletLoss's own call shape: one embedding matrix three times, anchors = all its rows -> row-owner backward
        ctx.shared = (ia is None and ip is not None and in_ is not None and T == ta.size(0) and T > 0
                      and ta.data_ptr() == tp.data_ptr() == tn.data_ptr() and ta.shape == tp.shape == tn.shape
                      and ta.stride() == tp.stride() == tn.stride() and ta.data_ptr() % 16 == 0 and ops._ld(ta) % 4 == 0
                      and bool(lib.sgb_triplet_self_bwd_supported(D)) and os.environ.get("SEGGER_B200_LOSS_FUSED", "1") != "0")
        ctx.save_for_backward(ta, tp, tn, ia, ip, in_, d_ap, d_an)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        ta, tp, tn, ia, ip, in_, d_ap, d_an = ctx.saved_tensors
        T, D, margin, eps = ctx.cfg
        dev = ta.device
        if T == 0:
            return torch.zeros_like(ta), torch.zeros_like(tp), torch.zeros_like(tn), None, None, None, None, None
        if ctx.shared:
            # d loss / d emb[r] in ONE pass over the rows: CSRs of the sampled positives / negatives (which triplets
            # point at row r; two key sorts) replace three [T, D] per-triplet gradient tensors, two sorts + segment sums
            # of them and the two [T, D] additions autograd would make of the three results
            ar = torch.arange(T, dtype=torch.int64, device=dev)
            csr_p = ops.build_csr(torch.stack([ar, ip]), T, T, transpose=False)
            csr_n = ops.build_csr(torch.stack([ar, in_]), T, T, transpose=False)
            gout = torch.empty(T, D, dtype=torch.float32, device=dev)
            g = dloss.to(torch.float32).contiguous()
            check(_lib.load().sgb_triplet_self_bwd(ptr(ta), ops._ld(ta), ptr(ip), ptr(in_), T, D, margin, eps, ptr(d_ap),
                                                   ptr(d_an), ptr(g), ptr(csr_p.rowptr), ptr(csr_p.col), ptr(csr_n.rowptr),
                                                   ptr(csr_n.col), ptr(gout), D, stream_ptr(dev)), "triplet_self_bwd")
            ops._count(1)
            return gout, None, None, None, None, None, None, None
        ga = torch.empty(T, D, dtype=torch.float32, device=dev)
        gp, gn = torch.empty_like(ga), torch.empty_like(ga)
        g = dloss.to(torch.float32).contiguous()
        check(_lib.load().sgb_triplet_margin_bwd(ptr(ta), ops._ld(ta), ptr(ia), ptr(tp), ops._ld(tp), ptr(ip), ptr(tn),
                                                 ops._ld(tn), ptr(in_), T, D, margin, eps, ptr(d_ap), ptr(d_an), ptr(g),
                                                 ptr(ga), ptr(gp), ptr(gn), stream_ptr(dev)), "triplet_margin_bwd")
        ops._count(1)
        return (_segment_rows(ga, ia, ta.size(0)), _segment_rows(gp, ip, tp.size(0)), _segment_rows(gn, in_, tn.size(0)),
                None, None, None, None, None)


class _PairLossFn(torch.autograd.Function):
    """mode 0: mse(cosine(a, b), target); mode 1: BCE-with-logits(a . b, target); mean over gathered row pairs."""

    @staticmethod
    def forward(ctx, ta, tb, ia, ib, target, mode, eps):
        require_cuda(ta, tb, target)
        ta, tb = ops._rowmajor(ta), ops._rowmajor(tb)
        target = target.to(torch.float32).contiguous()
        T, D = target.numel(), ta.size(1)
        dev = ta.device
        lib = _lib.load()
        val = torch.empty(T, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = ops._ws(lib.sgb_loss_workspace_bytes(T), dev)
        check(lib.sgb_pair_loss_fwd(ptr(ta), ops._ld(ta), ptr(ia), ptr(tb), ops._ld(tb), ptr(ib), ptr(target), T, D, int(mode),
                                    float(eps), ptr(val), ptr(loss), ptr(ws), ws.numel(), stream_ptr(dev)), "pair_loss_fwd")
        ops._count(2)
        ctx.cfg = (T, D, int(mode), float(eps))
        ctx.save_for_backward(ta, tb, ia, ib, target, val)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        ta, tb, ia, ib, target, val = ctx.saved_tensors
        T, D, mode, eps = ctx.cfg
        dev = ta.device
        if T == 0:
            return torch.zeros_like(ta), torch.zeros_like(tb), None, None, None, None, None
        gA = torch.empty(T, D, dtype=torch.float32, device=dev)
        gB = torch.empty_like(gA)
        g = dloss.to(torch.float32).contiguous()
        check(_lib.load().sgb_pair_loss_bwd(ptr(ta), ops._ld(ta), ptr(ia), ptr(tb), ops._ld(tb), ptr(ib), ptr(target), T, D,
                                            mode, eps, ptr(val), ptr(g), ptr(gA), ptr(gB), stream_ptr(dev)), "pair_loss_bwd")
        ops._count(1)
        return _segment_rows(gA, ia, ta.size(0)), _segment_rows(gB, ib, tb.size(0)), None, None, None, None, None


def triplet_margin(ta: Tensor, tp: Tensor, tn: Tensor, ia: Optional[Tensor], ip: Optional[Tensor], in_: Optional[Tensor],
                   margin: float, eps: float = 1e-6) -> Tensor:
    """TripletMarginLoss(margin)(ta[ia], tp[ip], tn[in_]) without materialising the gathered operands."""
    return _TripletMarginFn.apply(ta, tp, tn, None if ia is None else _i64(ia), None if ip is None else _i64(ip),
                                  None if in_ is None else _i64(in_), margin, eps)


def cosine_mse(ta: Tensor, tb: Tensor, ia: Optional[Tensor], ib: Optional[Tensor], target: Tensor, eps: float = 1e-8) -> Tensor:
    return _PairLossFn.apply(ta, tb, None if ia is None else _i64(ia), None if ib is None else _i64(ib), target, 0, eps)


def dot_bce(ta: Tensor, tb: Tensor, ia: Optional[Tensor], ib: Optional[Tensor], target: Tensor) -> Tensor:
    return _PairLossFn.apply(ta, tb, None if ia is None else _i64(ia), None if ib is None else _i64(ib), target, 1, 0.0)


class FastTripletSelector:
    """triplet_loss.py:8-125.  ``sample_triplets(labels)`` -> (positives, negatives, dists_pos, dists_neg)."""

    @torch.no_grad()
    def __init__(self, cluster_similarity: Tensor):
        _min_sampling_prob = 1e-8
        cluster_similarity.fill_diagonal_(1)                               # in place, as the reference (:22)
        self.similarity = cluster_similarity.clamp_min(_min_sampling_prob)
        self.dissimilarity = (-cluster_similarity).clamp_min(_min_sampling_prob)
        self._index_built = False

    @torch.no_grad()
    def _build_index(self, labels: Tensor, sorted_idx: Optional[Tensor] = None) -> None:
        """:27-86.  Members of a cluster are listed in index order (stable sort through the CSR builder); the
        reference's ``torch.argsort(labels)`` (:41) is not stable, so which member a uniform number selects inside a
        cluster is pinned only up to that order -- ``sorted_idx`` injects a given order (parity tests)."""
        C = self.similarity.size(0)
        device = labels.device
        N = labels.numel()
        labels = _i64(labels)
        # stable sort by cluster = destination-sorted CSR of the "edges" (i -> labels[i])
        ei = torch.stack([torch.arange(N, device=device, dtype=torch.int64), labels])
        csr = ops.build_csr(ei, max(N, 1), C, transpose=False)
        offsets_all = csr.rowptr.to(torch.int64)
        counts = offsets_all[1:] - offsets_all[:-1]
        present = torch.nonzero(counts > 0, as_tuple=False).flatten()
        sim = self.similarity.to(device=device, dtype=torch.float32)
        diss = self.dissimilarity.to(device=device, dtype=torch.float32)
        diss_pres = diss[present][:, present]
        cdf_neg = torch.cumsum(diss_pres / diss_pres.sum(dim=1, keepdim=True), dim=1)
        cdf_neg[:, -1] = 1.0
        sim_pres = sim[present][:, present]
        cdf_pos = torch.cumsum(sim_pres / sim_pres.sum(dim=1, keepdim=True), dim=1)
        cdf_pos[:, -1] = 1.0
        present_idx = -torch.ones(C, dtype=torch.long, device=device)
        present_idx[present] = torch.arange(present.numel(), device=device)
        self._counts = counts.contiguous()
        self._offsets = offsets_all[:-1].contiguous()
        self._sorted_idx = csr.eid.to(torch.int64) if sorted_idx is None else _i64(sorted_idx).to(device)
        self._present = present.contiguous()
        self._cdf_neg = cdf_neg.contiguous()
        self._cdf_pos = cdf_pos.contiguous()
        self._present_idx = present_idx
        self._sim_dev = sim.contiguous()
        self._index_built = True

    @torch.no_grad()
    def sample_triplets(self, labels: Tensor, uniforms=None, sorted_idx=None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        require_cuda(labels)
        if sorted_idx is None:
            sorted_idx = getattr(self, "_inject_sorted_idx", None)      # parity-test hook (see _build_index)
        self._build_index(labels, sorted_idx)
        device = labels.device
        N = labels.numel()
        labels = _i64(labels)
        if uniforms is None:   # same four draws, same order, as triplet_loss.py:93,99,105,112
            u_pos = torch.rand(N, device=device)
            u2 = torch.rand(N, device=device)
            u_neg = torch.rand(N, device=device)
            u3 = torch.rand(N, device=device)
        else:
            u_pos, u2, u_neg, u3 = [u.to(device=device, dtype=torch.float32).contiguous() for u in uniforms]
        positives = torch.empty(N, dtype=torch.int64, device=device)
        negatives = torch.empty(N, dtype=torch.int64, device=device)
        dists_pos = torch.empty(N, dtype=torch.float32, device=device)
        dists_neg = torch.empty(N, dtype=torch.float32, device=device)
        C, P = self._sim_dev.size(0), self._present.numel()
        if N > 0:
            check(_lib.load().sgb_triplet_sample(ptr(labels), N, C, P, ptr(self._present_idx), ptr(self._present),
                                                 ptr(self._counts), ptr(self._offsets), ptr(self._sorted_idx),
                                                 ptr(self._cdf_pos), ptr(self._cdf_neg), ptr(self._sim_dev), ptr(u_pos),
                                                 ptr(u2), ptr(u_neg), ptr(u3), ptr(positives), ptr(negatives),
                                                 ptr(dists_pos), ptr(dists_neg), stream_ptr(device)), "triplet_sample")
            ops._count(1)
        return positives, negatives, dists_pos, dists_neg


class TripletLoss(torch.nn.Module):
    """triplet_loss.py:128-160: TripletMarginLoss(margin) on triplets sampled by FastTripletSelector."""

    def __init__(self, cluster_similarity: Tensor, margin: float = 1.0, **kwargs) -> None:
        super().__init__()
        unsupported = {k: v for k, v in kwargs.items() if (k, v) not in (("p", 2.0), ("p", 2), ("swap", False),
                                                                           ("reduction", "mean"))
                       and k != "eps"}
        if unsupported:
            raise NotImplementedError(f"TripletLoss: unsupported TripletMarginLoss options {unsupported}")
        if margin <= 0:
            raise ValueError(f"TripletMarginLoss: expected margin to be greater than 0, got {margin} instead")
        self.margin = margin
        self.eps = float(kwargs.get("eps", 1e-6))
        self.selector = FastTripletSelector(cluster_similarity)

    def forward(self, embeddings: Tensor, labels: Tensor, triplets=None):
        """``triplets`` = a ``selector.sample_triplets(labels)`` result drawn ahead of time (LitISTEncoder.get_losses
        samples on a side stream while the forward pass runs); sampling depends on the labels only."""
        if labels.numel() == 0:
            return 0.
        positives, negatives, _, _ = self.selector.sample_triplets(labels) if triplets is None else triplets
        return triplet_margin(embeddings, embeddings, embeddings, None, positives, negatives, self.margin, self.eps)


class MetricLoss:
    """triplet_loss.py:163-204: MSE between cosine similarities and (1 - cluster distance) of sampled pairs."""

    def __init__(self, cluster_similarity: Tensor) -> None:
        self.selector = FastTripletSelector(cluster_similarity)

    def forward(self, embeddings: Tensor, labels: Tensor, triplets=None):
        if labels.numel() == 0:
            return 0.
        positives, negatives, dists_pos, dists_neg = self.selector.sample_triplets(labels) if triplets is None else triplets
        return (cosine_mse(embeddings, embeddings, None, positives, 1 - dists_pos)
                + cosine_mse(embeddings, embeddings, None, negatives, 1 - dists_neg))


def segmentation_loss(emb_tx: Tensor, emb_bd: Tensor, edge_index: Tensor, kind: str = "triplet", margin: float = 0.4,
                      dst_neg: Optional[Tensor] = None) -> Tensor:
    """lightning_model.py:163-205: positive pairs = tx-belongs-bd edges, negatives = a random other boundary
    (``dst_neg``, drawn like the reference with torch.randint when not given)."""
    src_pos, dst_pos = edge_index[0].long(), edge_index[1].long()
    num_bd = emb_bd.size(0)
    N = src_pos.size(0)
    if num_bd <= 1:
        return torch.tensor(0.0, device=emb_bd.device, requires_grad=True)
    if dst_neg is None:
        dst_neg = (dst_pos + torch.randint(1, num_bd, (N,), device=dst_pos.device)) % num_bd
    if kind == "triplet":
        return triplet_margin(emb_tx, emb_bd, emb_bd, src_pos, dst_pos, dst_neg, margin)
    if kind != "bce":
        raise ValueError(f"Unrecognized segmentation loss: '{kind}'. Acceptable values are 'triplet' and 'bce'.")
    src = torch.cat([src_pos, src_pos])
    dst = torch.cat([dst_pos, dst_neg])
    labels = torch.cat([torch.ones(N, device=emb_tx.device), torch.zeros(N, device=emb_tx.device)])
    return dot_bce(emb_tx, emb_bd, src, dst, labels)
