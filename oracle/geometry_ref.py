"""CPU restatement of the prediction-graph spatial join -- TEST INFRASTRUCTURE ONLY.

The reference finds "which transcripts lie strictly inside which buffered cell outline" with cuSpatial
(/root/reference/src/segger/geometry/query.py:21-100: quadtree_point_in_polygon behind
points_in_polygons(predicate='contains'), called from setup_prediction_graph,
/root/reference/src/segger/data/utils/neighbors.py:226-238).  cuSpatial (pinned 25.04 in pixi.lock) is absent from
/root/reference and from this image, so this file restates its published algorithm -- the even-odd (crossing number)
point-in-polygon test over polygon rings -- in numpy float64.  PARITY UNPINNED: the reference ships no test or fixture
for this path; behaviour for points exactly on an edge is documented by cuSpatial as "not contained" for interior
crossings but is implementation-defined at vertices; here (and in the kernel, which performs the same fp64 operations in
the same order) the half-open rule "an edge counts iff exactly one endpoint is strictly above the point's y" decides.
Polygon buffering (geopandas .buffer, neighbors.py:229-230) is host geometry in the reference and is not restated:
inputs are the already-buffered rings.
"""
from __future__ import annotations

import numpy as np


def points_in_polygons_ref(points: np.ndarray, verts: np.ndarray, ring_off: np.ndarray) -> np.ndarray:
    """-> int32 [2, E]: row 0 = point index, row 1 = polygon index; point-major, polygons ascending per point
    (the reference's pair order is whatever the quadtree join emits; the set of pairs is the contract)."""
    pts = np.asarray(points, dtype=np.float64)
    verts = np.asarray(verts, dtype=np.float64)
    out_p, out_g = [], []
    for g in range(len(ring_off) - 1):
        ring = verts[ring_off[g]:ring_off[g + 1]]
        if len(ring) >= 2 and ring[0, 0] == ring[-1, 0] and ring[0, 1] == ring[-1, 1]:
            ring = ring[:-1]
        if len(ring) < 3:
            continue
        lo, hi = ring.min(0), ring.max(0)
        cand = np.nonzero((pts[:, 0] >= lo[0]) & (pts[:, 0] <= hi[0]) & (pts[:, 1] >= lo[1]) & (pts[:, 1] <= hi[1]))[0]
        if cand.size == 0:
            continue
        px, py = pts[cand, 0], pts[cand, 1]
        inside = np.zeros(cand.size, dtype=bool)
        a = ring[-1]
        for b in ring:
            straddle = (a[1] > py) != (b[1] > py)
            with np.errstate(divide="ignore", invalid="ignore"):
                t = ((py - a[1]) * (b[0] - a[0])) / (b[1] - a[1])     # same operation order as pip_inside
                cross = straddle & (px < a[0] + t)
            inside ^= cross
            a = b
        hit = cand[inside]
        out_p.append(hit)
        out_g.append(np.full(hit.size, g, dtype=np.int64))
    if not out_p:
        return np.zeros((2, 0), dtype=np.int32)
    p, g = np.concatenate(out_p), np.concatenate(out_g)
    order = np.lexsort((g, p))
    return np.stack([p[order], g[order]]).astype(np.int32)
