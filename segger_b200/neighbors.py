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


def _xy(tx) -> np.ndarray:
    """[N, 2] coordinates of a transcript table: a frame with 'x' / 'y' columns (polars / pandas: the reference's
    ``tx[[tx_fields.x, tx_fields.y]].to_numpy()``, neighbors.py:174, TrainingTranscriptFields.x/.y = 'x'/'y'), a
    mapping of columns, or an [N, 2] array / tensor."""
    if isinstance(tx, (np.ndarray, Tensor)):
        return tx
    if isinstance(tx, dict):
        return np.stack([np.asarray(tx["x"]), np.asarray(tx["y"])], axis=1)
    sub = tx[["x", "y"]]
    return sub.to_numpy() if hasattr(sub, "to_numpy") else np.asarray(sub)


def setup_transcripts_graph(tx, max_k: int, max_dist: float) -> Tensor:
    """neighbors.py:166-180: the tx-neighbors-tx edge_index of a transcript table."""
    edge_index, _ = kdtree_neighbors(points=_xy(tx), max_k=max_k, max_dist=max_dist)
    return edge_index


def setup_transcripts_graph_xy(xy: np.ndarray, max_k: int, max_dist: float) -> Tensor:
    """Round-1 name of ``setup_transcripts_graph`` for an [N, 2] array; kept as an alias."""
    return setup_transcripts_graph(xy, max_k, max_dist)


def _knn_unbounded(points, query, max_k: int) -> Tensor:
    """k nearest neighbours without a radius cap on the radius-capped grid kernel: start from the radius that holds
    ~4k points at the mean density, double it for the queries that found fewer than k until all are full (or the
    radius covers the whole point set)."""
    pts = np.asarray(points, dtype=np.float64)
    n = pts.shape[0]
    k = min(max_k, n)
    span = np.maximum(pts.max(0) - pts.min(0), 1e-9)
    r = float(np.sqrt(4.0 * max(k, 1) * span[0] * span[1] / (np.pi * max(n, 1)))) + 1e-9
    diag = float(np.hypot(*(np.maximum(pts.max(0), np.asarray(query).max(0)) - np.minimum(pts.min(0), np.asarray(query).min(0))))) + 1.0
    while True:
        table, count = knn_table(points, max_k, r, query=query)
        if r > diag or int((count < k).sum()) == 0:
            return table
        r *= 2.0


def setup_prediction_graph(tx, bd, max_k: int, buffer_ratio: float, mode: str = "cell", device=None) -> Tensor:
    """neighbors.py:200-238: the tx-neighbors-bd candidate edges.

    ``mode`` 'cell' / 'nucleus': transcripts strictly inside the outlines of that boundary type, each grown by
    ``sqrt(area / pi) * buffer_ratio`` -- the point-in-polygon join runs on the GPU (``segger_b200.geometry``);
    returns int32 CPU ``[2, E]`` = (transcript row, boundary row) like the reference.  ``bd`` is either a
    GeoDataFrame (buffered on the host with shapely exactly as the reference does, :229-231) or outlines that
    are buffered already: a ``geometry.PackedPolygons`` or a ``(verts, ring_off)`` pair (``buffer_ratio`` must
    then be 0 or None).
    ``mode`` 'uniform': k nearest transcripts of every boundary centroid.  The reference calls
    ``kdtree_neighbors(points, query, max_k)`` without the mandatory ``max_dist`` there and raises TypeError
    (SURVEY Appendix B.4); the evident intent -- no radius cap -- is what this does.  ``bd`` may be the
    GeoDataFrame or an [M, 2] centroid array; rows are (boundary row, transcript row) as that call returns them.
    """
    from .geometry import PackedPolygons, pack_rings, points_in_polygons
    points = _xy(tx)
    if mode == "uniform":
        if hasattr(bd, "geometry"):
            query = bd.geometry.centroid.get_coordinates().values
        else:
            query = bd.cpu().numpy() if isinstance(bd, Tensor) else np.asarray(bd)
        pts = points.cpu().numpy() if isinstance(points, Tensor) else np.asarray(points)
        table = _knn_unbounded(pts, query, max_k)
        with torch.no_grad():
            edge_index, _ = _table_to_coo(table, None, int(pts.shape[0]))
        return edge_index.cpu()
    if mode not in ("cell", "nucleus"):
        raise ValueError(f"setup_prediction_graph: unknown mode '{mode}' (expected 'nucleus', 'cell' or 'uniform')")
    if isinstance(bd, PackedPolygons):
        polys = bd
    elif isinstance(bd, (tuple, list)) and len(bd) == 2:
        polys = PackedPolygons(*bd)
    elif hasattr(bd, "geometry"):
        boundary_type = "cell" if mode == "cell" else "nucleus"          # StandardBoundaryFields (io/fields.py:121-123)
        polygons = bd[bd["boundary_type"] == boundary_type].geometry
        buffer_dists = np.sqrt(polygons.area / np.pi) * buffer_ratio
        polygons = polygons.buffer(buffer_dists).reset_index(drop=True)
        polys = PackedPolygons(*pack_rings([np.asarray(g.exterior.coords) for g in polygons]))
        buffer_ratio = 0
    else:
        raise TypeError("setup_prediction_graph: `bd` must be a GeoDataFrame, a PackedPolygons or (verts, ring_off)")
    if buffer_ratio:
        raise ValueError("setup_prediction_graph: packed outlines are taken as already buffered (pass buffer_ratio=0); "
                         "buffering is host geometry (shapely), as in the reference")
    return points_in_polygons(points, polys, device=device, device_output=False)
