"""kNN graph construction oracle -- the reference's own scipy call, restated.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PINNED: `kdtree_neighbors`
issues the same ``scipy.spatial.KDTree(points, leafsize=100).query(...)`` call
as /root/reference/src/segger/data/utils/neighbors.py:139-150 (scipy is a
third-party dependency of the reference: pixi.lock pins 1.17.1; 1.18.1 here,
same cKDTree algorithm) and `knn_to_edge_index` follows :54-92 1:1.

Tie contract: scipy returns equal-distance neighbours
in heap order; the product orders by (d^2, index).  `canonical_knn_table`
re-derives the (d^2, idx)-ordered table by brute force in float64 for rows the
test harness flags as tie-affected.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.spatial import KDTree


def knn_to_edge_index(neighbor_table: torch.Tensor, padding_value=None):
    """neighbors.py:54-92 (minus the gc / empty_cache housekeeping)."""
    N, K = neighbor_table.shape
    if padding_value is None:
        padding_value = N
    valid = neighbor_table != padding_value
    flat = valid.view(-1).nonzero(as_tuple=False).squeeze(1)
    col = neighbor_table.reshape(-1)[flat]
    row = flat // K
    edge_index = torch.stack([row, col])
    deg = valid.sum(dim=1)
    index_ptr = torch.cat((torch.zeros(1, dtype=torch.long), deg.cumsum(0)))
    return edge_index, index_ptr


def kdtree_table(points: np.ndarray, max_k: int, max_dist: float, query=None, workers: int = -1):
    """The raw scipy call of neighbors.py:139-150 -> (distances, indices) padded with tree.n / inf."""
    q = points if query is None else query
    tree = KDTree(points, leafsize=100)
    return tree.query(q, k=max_k, distance_upper_bound=max_dist, workers=workers)


def kdtree_neighbors(points: np.ndarray, max_k: int, max_dist: float,
                     chunk_size: int = 2_000_000, query=None, workers: int = -1):
    """neighbors.py:122-163."""
    q = points if query is None else query
    N = q.shape[0]
    tree = KDTree(points, leafsize=100)
    edge_indices = []
    for i in range(0, N, chunk_size):
        _, indices = tree.query(q[i:i + chunk_size], k=max_k, distance_upper_bound=max_dist,
                                workers=workers)
        indices = torch.from_numpy(indices.copy())
        edge_index, _ = knn_to_edge_index(indices, padding_value=N)
        edge_index[0] += i
        edge_indices.append(edge_index)
    return torch.cat(edge_indices, dim=1), None


def brute_force_knn_table(points: np.ndarray, max_k: int, max_dist: float, query=None,
                          rows=None) -> np.ndarray:
    """(d^2, idx)-ordered kNN table by exhaustive float64 search (pure numpy; small inputs / few rows).

    d^2 = fl(fl(dx*dx) + fl(dy*dy)) in float64 of the (float32 or float64) coordinates -- the same
    arithmetic cKDTree performs -- accepted iff d^2 < max_dist^2 (strict, Appendix A.5).
    """
    P = np.asarray(points, dtype=np.float64)
    Q = P if query is None else np.asarray(query, dtype=np.float64)
    rows = np.arange(Q.shape[0]) if rows is None else np.asarray(rows)
    n = P.shape[0]
    out = np.full((rows.shape[0], max_k), n, dtype=np.int64)
    r2 = np.float64(max_dist) * np.float64(max_dist)
    for o, qi in enumerate(rows):
        dx = P[:, 0] - Q[qi, 0]
        dy = P[:, 1] - Q[qi, 1]
        d2 = dx * dx + dy * dy
        cand = np.nonzero(d2 < r2)[0]
        order = np.lexsort((cand, d2[cand]))[:max_k]
        sel = cand[order]
        out[o, :sel.shape[0]] = sel
    return out


def canonical_knn_table(points: np.ndarray, max_k: int, max_dist: float, query=None,
                        workers: int = -1):
    """scipy result re-ordered to the product's (d^2, idx) contract (Appendix A.5).

    scipy is queried with k+1 so that a tie straddling the k-th slot is detectable
    (d2[k-1] == d2[k]); only those rows are re-derived by brute force.  All other rows are the
    scipy rows, stably re-sorted by (d^2, idx).  Returns (table int64 [Nq, k] padded with n,
    n_tie_rows, n_bruteforced_rows, scipy_raw_idx [Nq,k]).
    """
    P = np.asarray(points, dtype=np.float64)
    Q = P if query is None else np.asarray(query, dtype=np.float64)
    n = P.shape[0]
    _, idx1 = kdtree_table(points, max_k + 1, max_dist, query=query, workers=workers)
    idx1 = idx1.reshape(Q.shape[0], max_k + 1)
    safe = np.where(idx1 == n, 0, idx1)
    dx = P[safe, 0] - Q[:, None, 0]
    dy = P[safe, 1] - Q[:, None, 1]
    d2 = dx * dx + dy * dy
    d2 = np.where(idx1 == n, np.inf, d2)
    # canonical order: primary d2, secondary idx (padding has d2=inf, idx=n -> stays last)
    order = np.lexsort((idx1, d2), axis=1)
    idx_s = np.take_along_axis(idx1, order, axis=1)
    d2_s = np.take_along_axis(d2, order, axis=1)
    internal = ((np.diff(d2_s[:, :max_k], axis=1) == 0) & np.isfinite(d2_s[:, 1:max_k])).any(axis=1)
    boundary = np.isfinite(d2_s[:, max_k]) & (d2_s[:, max_k - 1] == d2_s[:, max_k])
    table = idx_s[:, :max_k].copy()
    rows = np.nonzero(boundary)[0]
    if rows.size:
        table[rows] = brute_force_knn_table(points, max_k, max_dist, query=query, rows=rows)
    _, raw = kdtree_table(points, max_k, max_dist, query=query, workers=workers)
    return table, int((internal | boundary).sum()), int(rows.size), raw.reshape(Q.shape[0], max_k)
