"""GPU parity of the point-in-polygon join against the numpy oracle: bit-exact pair lists."""
import numpy as np
import pytest
import torch

from oracle.geometry_ref import points_in_polygons_ref
from segger_b200.geometry import PackedPolygons, pack_rings, points_in_polygons, setup_prediction_graph
from segger_b200.synth import synth

pytestmark = pytest.mark.gpu


def _ngons(centers, radius, n=16, jitter=None, rng=None, close=False):
    rings = []
    for i, c in enumerate(centers):
        ang = np.linspace(0, 2 * np.pi, n, endpoint=False)
        r = radius if jitter is None else radius * (1 + jitter * rng.uniform(-1, 1, n))
        ring = np.stack([c[0] + r * np.cos(ang), c[1] + r * np.sin(ang)], 1)
        if close and i % 2 == 0:
            ring = np.concatenate([ring, ring[:1]])
        rings.append(ring)
    return rings


@pytest.mark.parametrize("n_pts,n_poly,dtype", [(20000, 200, np.float32), (50000, 37, np.float64), (300, 1, np.float32)])
def test_points_in_polygons_bit_exact_vs_oracle(n_pts, n_poly, dtype):
    rng = np.random.default_rng(n_pts)
    side = 14.0 * np.sqrt(n_poly) + 20
    centers = rng.uniform(10, side - 10, (n_poly, 2))
    verts, off = pack_rings(_ngons(centers, 6.8, 16, jitter=0.3, rng=rng, close=True))      # overlapping, non-convex-ish rings
    pts = rng.uniform(-5, side + 5, (n_pts, 2)).astype(dtype)                                # some outside the grid
    ref = points_in_polygons_ref(pts, verts, off)
    got = points_in_polygons(pts, verts, off)
    assert got.dtype == torch.int32 and not got.is_cuda
    assert np.array_equal(got.numpy(), ref)
    assert ref.shape[1] > 0
    # deterministic (slot order inside a polygon comes from atomics; the final order must not)
    assert torch.equal(points_in_polygons(torch.from_numpy(pts).cuda(), verts, off, device_output=True).cpu(), got)


def test_points_in_polygons_edge_cases():
    sq = np.array([[0.0, 0.0], [4.0, 0.0], [4.0, 4.0], [0.0, 4.0]])
    lshape = np.array([[10.0, 0.0], [16.0, 0.0], [16.0, 2.0], [12.0, 2.0], [12.0, 6.0], [10.0, 6.0]])   # concave
    verts, off = pack_rings([sq, lshape, np.zeros((0, 2)), np.array([[1.0, 1.0], [2.0, 2.0]])])       # + empty, degenerate
    pts = np.array([[2.0, 2.0], [4.0, 2.0], [0.0, 0.0], [2.0, 4.0], [14.0, 4.0], [11.0, 5.0], [15.0, 1.0], [100.0, 100.0],
                    [2.0, 0.0], [0.0, 2.0]], dtype=np.float64)
    ref = points_in_polygons_ref(pts, verts, off)
    got = points_in_polygons(pts, verts, off).numpy()
    assert np.array_equal(got, ref)
    pairs = set(map(tuple, got.T))
    assert (0, 0) in pairs and (5, 1) in pairs and (6, 1) in pairs and (4, 1) not in pairs and (7, 0) not in pairs
    assert points_in_polygons(np.zeros((0, 2)), verts, off).shape == (2, 0)
    assert points_in_polygons(pts, *pack_rings([])).shape == (2, 0)


def test_prediction_graph_on_synthetic_tile_matches_oracle_and_feeds_scoring():
    """Buffered 16-gon cell outlines of a synthetic tile (SURVEY 8d generator): candidate edges == oracle; every
    transcript generated inside a cell is a candidate of that cell."""
    ts = synth(50_000, 500, seed=3)
    r_buf = 6.5 * 1.05
    verts, off = pack_rings(_ngons(ts.bd_pos.astype(np.float64), r_buf, 16))
    ref = points_in_polygons_ref(ts.tx_pos, verts, off)
    got = setup_prediction_graph(ts.tx_pos, verts, off)
    assert np.array_equal(got.numpy(), ref)
    polys = PackedPolygons(verts, off)                         # packed once, reused across calls / tiles
    half = ts.tx_pos[:25_000]
    assert np.array_equal(points_in_polygons(half, polys).numpy(), points_in_polygons_ref(half, verts, off))
    assert np.array_equal(points_in_polygons(ts.tx_pos, polys).numpy(), ref)
    own = ts.tx_cell >= 0
    d = np.linalg.norm(ts.tx_pos[own].astype(np.float64) - ts.bd_pos[ts.tx_cell[own]].astype(np.float64), axis=1)
    deep = np.nonzero(own)[0][d < r_buf * np.cos(np.pi / 16) - 1e-6]          # inside the inscribed circle
    pairs = set(map(tuple, got.numpy().T))
    assert all((int(i), int(ts.tx_cell[i])) in pairs for i in deep[::50])


def test_points_in_polygons_against_committed_golden_fixture():
    import os
    gd = np.load(os.path.join(os.path.dirname(__file__), "golden", "pip_small.npz"))
    got = points_in_polygons(gd["points"], gd["verts"], gd["ring_off"])
    assert np.array_equal(got.numpy(), gd["pairs"])
