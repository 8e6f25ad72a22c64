"""B200-native point-in-polygon join for the prediction graph:
``points_in_polygons(points, polygons, predicate='contains')`` of /root/reference/src/segger/geometry/query.py:21-100
and the shape modes of ``setup_prediction_graph`` (/root/reference/src/segger/data/utils/neighbors.py:226-238), on
``sgb_pip_count`` / ``sgb_pip_fill`` instead of cuSpatial's quadtree join.  Polygons are passed as packed rings
(``verts`` [V,2] float64 + ``ring_off`` [M+1]); buffering the outlines stays host geometry as in the reference."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib, ops
from ._lib import check, ptr, stream_ptr


def pack_rings(rings) -> Tuple[np.ndarray, np.ndarray]:
    """list of [n_i, 2] arrays -> (verts [V,2] float64, ring_off [M+1] int64)."""
    off = np.zeros(len(rings) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in rings])
    verts = np.concatenate([np.asarray(r, dtype=np.float64).reshape(-1, 2) for r in rings]) if rings else np.zeros((0, 2))
    return np.ascontiguousarray(verts, dtype=np.float64), off


def _grid(verts: np.ndarray, ring_off: np.ndarray):
    """Uniform grid over the polygons' bounding box, cell ~ the median polygon extent (host side: the rings come from
    host geometry anyway); capped at 2^26 cells."""
    sizes = np.diff(ring_off)
    starts = ring_off[:-1][sizes > 0]
    # per-ring boxes with reduceat (rings are contiguous runs; empty rings are skipped), then the union box
    mx = np.maximum.reduceat(verts, starts, axis=0)
    mn = np.minimum.reduceat(verts, starts, axis=0)
    ext = mx - mn
    lo = np.array([mn[:, 0].min(), mn[:, 1].min()])
    hi = np.array([mx[:, 0].max(), mx[:, 1].max()])
    cell = float(max(np.median(ext.max(1)), 1e-9))
    span = np.maximum(hi - lo, cell)
    while (np.floor(span[0] / cell) + 1) * (np.floor(span[1] / cell) + 1) >= 2 ** 26:
        cell *= 2.0
    nx, ny = int(np.floor(span[0] / cell)) + 1, int(np.floor(span[1] / cell)) + 1
    return float(lo[0]), float(lo[1]), cell, nx, ny


class PackedPolygons:
    """Rings + their grid + device copies, built once per boundary set (the outlines of a dataset do not change
    between prediction tiles)."""

    def __init__(self, verts: np.ndarray, ring_off: np.ndarray):
        self.verts = np.ascontiguousarray(verts, dtype=np.float64).reshape(-1, 2)
        self.ring_off = np.ascontiguousarray(ring_off, dtype=np.int64)
        self.n_poly = len(self.ring_off) - 1
        self.grid = _grid(self.verts, self.ring_off) if self.n_poly > 0 and self.verts.shape[0] > 0 else None
        self._dev = {}

    def on(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = (torch.from_numpy(self.verts).to(device), torch.from_numpy(self.ring_off).to(device))
        return self._dev[key]


def points_in_polygons(points, verts, ring_off=None, device=None, device_output: bool = False) -> Tensor:
    """-> int32 [2, E] (index_query = point, index_match = polygon), strict containment, point-major order.
    ``points``: [N, 2] numpy / tensor (float32 or float64; device tensors are used in place); polygons either as
    (``verts``, ``ring_off``) arrays or as a ``PackedPolygons`` (pass it as ``verts``)."""
    device = torch.device(device if device is not None else "cuda")
    polys = verts if isinstance(verts, PackedPolygons) else PackedPolygons(verts, ring_off)
    n_poly = polys.n_poly
    if isinstance(points, np.ndarray):
        pts = torch.from_numpy(np.ascontiguousarray(points if points.dtype in (np.float32, np.float64)
                                                    else points.astype(np.float64)))
    else:
        pts = points if points.dtype in (torch.float32, torch.float64) else points.double()
    pts = pts.to(device).contiguous()
    n = pts.size(0)
    empty = torch.zeros(2, 0, dtype=torch.int32, device=device if device_output else "cpu")
    if n == 0 or polys.grid is None:
        return empty
    xmin, ymin, cell, nx, ny = polys.grid
    lib = _lib.load()
    d_verts, d_off = polys.on(device)
    ws = torch.empty(max(int(lib.sgb_pip_workspace_bytes(n, n_poly, nx, ny)), 16), dtype=torch.uint8, device=device)
    total = torch.zeros(1, dtype=torch.int32, device=device)
    f64 = int(pts.dtype == torch.float64)
    check(lib.sgb_pip_count(ptr(pts), f64, n, ptr(d_verts), ptr(d_off), n_poly, xmin, ymin, cell, nx, ny, ptr(total), ptr(ws),
                            ws.numel(), stream_ptr(device)), "pip_count")
    E = int(total.item())
    ops._count(12)
    if E == 0:
        return empty
    out = torch.empty(2, E, dtype=torch.int32, device=device)
    scratch = torch.empty(int(lib.sgb_pip_fill_scratch_bytes(E)), dtype=torch.uint8, device=device)
    check(lib.sgb_pip_fill(ptr(d_verts), ptr(d_off), n, n_poly, xmin, ymin, cell, nx, ny, E, ptr(out), E, ptr(ws), ws.numel(),
                           ptr(scratch), scratch.numel(), stream_ptr(device)), "pip_fill")
    ops._count(12)
    return out if device_output else out.cpu()


def setup_prediction_graph(points, verts, ring_off=None, device=None) -> Tensor:
    """Round-1 entry point for already-buffered outlines; the drop-in with the reference's signature is
    ``segger_b200.neighbors.setup_prediction_graph(tx, bd, max_k, buffer_ratio, mode)``."""
    return points_in_polygons(points, verts, ring_off, device=device, device_output=False)
