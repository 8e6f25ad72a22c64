"""B200-native drop-in for the graph-construction functions of ``segger.data.utils.neighbors``
(/root/reference/src/segger/data/utils/neighbors.py): ``kdtree_neighbors`` (:122-163),
``knn_to_edge_index`` (:54-92), ``edge_index_to_knn`` is index plumbing and not on the hot path.

Same signatures and return types (CPU ``torch.int64`` edge_index, as scipy + torch produce in the
reference) -- the search itself runs on the GPU (uniform-grid kNN kernel, float64 distances,
strict radius, rows ordered by (d^2, index)).  Pass ``device_output=True`` to keep results on the
device and skip the D2H copy.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib, ops
from ._lib import KnnPlan, check, ptr, stream_ptr


def _as_device_points(a, device) -> Tuple[Tensor, bool]:
    if isinstance(a, np.ndarray):
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    elif isinstance(a, Tensor):
        t = a if a.dtype in (torch.float32, torch.float64) else a.double()
    else:
        t = torch.as_tensor(np.asarray(a, dtype=np.float64))
    if t.dim() != 2 or t.size(1) != 2:
        raise ValueError(f"points must be [N, 2], got {tuple(t.shape)}")
    t = t.to(device, non_blocking=True).contiguous()
    return t, t.dtype == torch.float64


def knn_table(points, max_k: int, max_dist: float, query=None, device=None) -> Tuple[Tensor, Tensor]:
    """Padded neighbour table [Nq, k] int64 (pad = n_points) and valid counts [Nq] int32, on device.
    The GPU equivalent of ``KDTree(points, leafsize=100).query(q, k, distance_upper_bound)[1]``."""
    device = torch.device(device if device is not None else "cuda")
    pts, f64 = _as_device_points(points, device)
    qry = None
    if query is not None:
        qry, qf64 = _as_device_points(query, device)
        if qf64 != f64:
            pts, qry, f64 = pts.double(), qry.double(), True
    lib = _lib.load()
    n, nq = pts.size(0), (qry.size(0) if qry is not None else pts.size(0))
    plan = KnnPlan()
    box = torch.empty(8, dtype=torch.float64, device=device)
    check(lib.sgb_knn2d_plan(ptr(pts), int(f64), n, ptr(qry), nq, int(max_k), float(max_dist), C.byref(plan),
                             ptr(box), stream_ptr(device)), "knn2d_plan")
    table = torch.empty(nq, max_k, dtype=torch.int64, device=device)
    count = torch.empty(nq, dtype=torch.int32, device=device)
    ws = torch.empty(max(int(lib.sgb_knn2d_workspace_bytes(C.byref(plan))), 16), dtype=torch.uint8, device=device)
    check(lib.sgb_knn2d(C.byref(plan), ptr(pts), int(f64), ptr(qry), ptr(table), ptr(count), ptr(ws), ws.numel(),
                        stream_ptr(device)), "knn2d")
    ops._count(14)
    return table, count


def _table_to_coo(table: Tensor, count: Optional[Tensor], padding_value: int, row_offset: int = 0):
    device = table.device
    lib = _lib.load()
    n, k = table.shape
    table = table.contiguous()
    if count is None:
        count = torch.empty(n, dtype=torch.int32, device=device)
        check(lib.sgb_knn_count_valid(ptr(table), n, k, int(padding_value), ptr(count), stream_ptr(device)),
              "knn_count_valid")
    index_ptr = torch.empty(n + 1, dtype=torch.int64, device=device)
    ws = torch.empty(max(int(lib.sgb_knn_coo_workspace_bytes(n)), 16), dtype=torch.uint8, device=device)
    n_edges = C.c_int64(0)
    check(lib.sgb_knn_count_edges(ptr(count), n, ptr(index_ptr), C.byref(n_edges), ptr(ws), ws.numel(),
                                  stream_ptr(device)), "knn_count_edges")
    E = int(n_edges.value)
    edge_index = torch.empty(2, E, dtype=torch.int64, device=device)
    check(lib.sgb_knn_table_to_coo(ptr(table), ptr(index_ptr), n, k, int(padding_value), int(row_offset), E,
                                   ptr(edge_index), stream_ptr(device)), "knn_table_to_coo")
    ops._count(6)
    return edge_index, index_ptr


def knn_to_edge_index(neighbor_table: Tensor, padding_value=None) -> Tuple[Tensor, Tensor]:
    """neighbors.py:54-92: dense padded neighbour table -> (edge_index [2,E], index_ptr [N+1]).
    Output lives on the table's device (a CPU table is staged through the GPU)."""
    N, K = neighbor_table.shape
    if padding_value is None:
        padding_value = N
    src_device = neighbor_table.device
    t = neighbor_table.to(torch.int64)
    if not t.is_cuda:
        t = t.cuda()
    with torch.no_grad():
        edge_index, index_ptr = _table_to_coo(t, None, int(padding_value))
    return edge_index.to(src_device), index_ptr.to(src_device)


def kdtree_neighbors(points: np.ndarray, max_k: int, max_dist: float, chunk_size: int = 2_000_000,
                     query: np.ndarray | None = None, device_output: bool = False, device=None):
    """neighbors.py:122-163: kNN (k = max_k, strict radius max_dist) of every query among ``points``
    as a COO edge list, row 0 = query index, row 1 = neighbour index; returns ``(edge_index, None)``.

    ``chunk_size`` is accepted for signature compatibility; the whole query set is one launch.
    """
    q = points if query is None else query
    N = int(q.shape[0])
    with torch.no_grad():
        table, count = knn_table(points, max_k, max_dist, query=query, device=device)
        # the reference pads with the *tree* size but drops entries equal to the *query* count N
        # (neighbors.py:154); identical when query is None.  Reproduce that literally.
        n_points = int(points.shape[0])
        if N == n_points:
            edge_index, _ = _table_to_coo(table, count, N)
        else:
            edge_index, _ = _table_to_coo(table, None, N)
    if not device_output:
        edge_index = edge_index.cpu()
    return edge_index, None


def setup_transcripts_graph_xy(xy: np.ndarray, max_k: int, max_dist: float) -> Tensor:
    """neighbors.py:166-180 with the polars frame already reduced to its [N,2] coordinate array."""
    edge_index, _ = kdtree_neighbors(points=xy, max_k=max_k, max_dist=max_dist)
    return edge_index
