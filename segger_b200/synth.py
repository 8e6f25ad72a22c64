"""Synthetic Xenium-shaped inputs (SURVEY.md section 8d generator).

Cells sit on a jittered square grid (pitch 14 um, jitter U(-3,3)); each cell emits 100
transcripts: 40 uniform in the nucleus disc (r=3.5), 50 in the cytoplasm annulus (3.5..6.5) and
10 background transcripts uniform over the whole slide.  Coordinates are float32 micrometres.
Node stores mirror what ``setup_heterodata`` builds
(/root/reference/src/segger/data/utils/heterodata.py:114-162): ``tx.x`` int32 gene id,
``tx.pos`` float32 [N,2], ``bd.x`` float32 [M,128], ``bd.pos`` float32 [M,2], ``bd.index`` int32,
edge stores ``tx-belongs-bd`` (int64, nuclear transcripts -> own cell) and ``tx-neighbors-bd``
(int32, cells whose 1.05x-buffered outline contains the transcript, at most 3 per transcript).
The ``tx-neighbors-tx`` kNN graph is NOT built here -- that is the product's own
``kdtree_neighbors`` (or the scipy oracle in tests).

Nodes are ordered tile-major (square tiles of <= ``nodes_per_tile`` transcripts), the layout
``PartitionDataset`` produces (/root/reference/src/segger/data/partition/dataset.py:463-506), and
``tile`` ids double as the ``batch`` vector the positional embedder normalises over.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

PITCH = 14.0
JITTER = 3.0
R_NUC = 3.5
R_CELL = 6.5
TX_PER_CELL = 100
N_NUC, N_CYT, N_BG = 40, 50, 10
BUFFER_RATIO = 0.05
PRED_MAX_K = 3


@dataclass
class SynthTileSet:
    tx_pos: np.ndarray        # float32 [N,2]
    tx_gene: np.ndarray       # int32 [N]
    tx_tile: np.ndarray       # int64 [N]   (batch vector)
    tx_cell: np.ndarray       # int64 [N]   owning cell or -1 (background)
    tx_compartment: np.ndarray  # int8 [N]  2 nucleus, 1 cytoplasm, 0 background
    tx_index: np.ndarray      # int64 [N]   original row index
    bd_x: np.ndarray          # float32 [M,bd_dim]
    bd_pos: np.ndarray        # float32 [M,2]
    bd_tile: np.ndarray       # int64 [M]
    bd_index: np.ndarray      # int32 [M]   cell encoding
    edge_tb: np.ndarray       # int64 [2,E_tb]
    edge_pred: np.ndarray     # int32 [2,E_pred]
    n_genes: int
    n_tiles: int
    side: float


def _disc(rng, n, r0, r1):
    """Uniform samples in the annulus r0..r1 (area-uniform)."""
    r = np.sqrt(rng.uniform(r0 * r0, r1 * r1, n))
    t = rng.uniform(0.0, 2.0 * math.pi, n)
    return r * np.cos(t), r * np.sin(t)


def synth(n_tx: int, n_cells: int, seed: int = 0, n_genes: int = 500, bd_dim: int = 128,
          nodes_per_tile: int = 50_000, pred_edges: bool = True) -> SynthTileSet:
    """``pred_edges=False`` skips the host-side tx-neighbors-bd candidate list (a scipy kd-tree query): large
    benchmarks build it with the product's point-in-polygon join instead (segger_b200.geometry)."""
    rng = np.random.default_rng(seed)
    g = int(math.ceil(math.sqrt(n_cells)))
    side = g * PITCH
    cid = np.arange(n_cells)
    cx = (cid % g + 0.5) * PITCH + rng.uniform(-JITTER, JITTER, n_cells)
    cy = (cid // g + 0.5) * PITCH + rng.uniform(-JITTER, JITTER, n_cells)

    per = max(1, n_tx // n_cells)
    n_nuc = int(round(per * N_NUC / TX_PER_CELL))
    n_cyt = int(round(per * N_CYT / TX_PER_CELL))
    n_bg_total = n_tx - n_cells * (n_nuc + n_cyt)
    assert n_bg_total >= 0

    # per-cell gene programme: 10 cell types, Dirichlet(0.1) profiles
    ctype = rng.integers(0, 10, n_cells)
    prof = rng.dirichlet(np.full(n_genes, 0.1), 10)
    cdf = np.cumsum(prof, axis=1)

    def genes_for(cells):
        # number of CDF entries below u = searchsorted(side='left'), per cell type (same values as the dense compare)
        u = rng.uniform(0, 1, cells.shape[0])
        t = ctype[cells]
        out = np.empty(cells.shape[0], np.int64)
        for k in range(cdf.shape[0]):
            m = t == k
            out[m] = np.searchsorted(cdf[k], u[m], side="left")
        return out.clip(0, n_genes - 1).astype(np.int32)

    owner_n = np.repeat(cid, n_nuc)
    dx, dy = _disc(rng, owner_n.shape[0], 0.0, R_NUC)
    xn, yn = cx[owner_n] + dx, cy[owner_n] + dy
    owner_c = np.repeat(cid, n_cyt)
    dx, dy = _disc(rng, owner_c.shape[0], R_NUC, R_CELL)
    xc, yc = cx[owner_c] + dx, cy[owner_c] + dy
    xb = rng.uniform(0, side, n_bg_total)
    yb = rng.uniform(0, side, n_bg_total)

    x = np.concatenate([xn, xc, xb])
    y = np.concatenate([yn, yc, yb])
    cell = np.concatenate([owner_n, owner_c, np.full(n_bg_total, -1)])
    comp = np.concatenate([np.full(owner_n.shape[0], 2, np.int8), np.full(owner_c.shape[0], 1, np.int8),
                           np.zeros(n_bg_total, np.int8)])
    # genes in chunks (keeps the [n, n_genes] compare small)
    gene = np.empty(n_tx, np.int32)
    own_all = np.where(cell >= 0, cell, rng.integers(0, n_cells, n_tx))
    for s in range(0, n_tx, 200_000):
        gene[s:s + 200_000] = genes_for(own_all[s:s + 200_000])

    # square tiles with <= nodes_per_tile transcripts on average
    t = max(1, int(math.ceil(math.sqrt(n_tx / nodes_per_tile))))
    tw = side / t
    tile = (np.clip((y // tw).astype(np.int64), 0, t - 1) * t
            + np.clip((x // tw).astype(np.int64), 0, t - 1))
    # tile-major, then owner cell (spatially coherent inside a tile), stable
    order = np.lexsort((np.where(cell >= 0, cell, n_cells), tile))
    x, y, cell, comp, gene, tile = x[order], y[order], cell[order], comp[order], gene[order], tile[order]
    tx_pos = np.stack([x, y], 1).astype(np.float32)

    bd_pos = np.stack([cx, cy], 1).astype(np.float32)
    bd_tile = (np.clip((cy // tw).astype(np.int64), 0, t - 1) * t
               + np.clip((cx // tw).astype(np.int64), 0, t - 1))
    bd_x = rng.standard_normal((n_cells, bd_dim)).astype(np.float32)

    # tx-belongs-bd: nuclear transcripts -> own cell (segmentation_graph_mode="nucleus")
    nuc = np.nonzero(comp == 2)[0]
    edge_tb = np.stack([nuc, cell[nuc]]).astype(np.int64)

    # tx-neighbors-bd: buffered-outline containment, <= PRED_MAX_K nearest cells
    if pred_edges:
        from scipy.spatial import cKDTree  # generator-side only (host preprocessing)
        r_buf = R_CELL * (1.0 + BUFFER_RATIO)
        _, nn = cKDTree(bd_pos.astype(np.float64)).query(
            tx_pos.astype(np.float64), k=PRED_MAX_K, distance_upper_bound=r_buf, workers=-1)
        valid = nn != n_cells
        src = np.repeat(np.arange(n_tx), PRED_MAX_K).reshape(n_tx, PRED_MAX_K)[valid]
        edge_pred = np.stack([src, nn[valid]]).astype(np.int32)
    else:
        edge_pred = np.zeros((2, 0), np.int32)

    return SynthTileSet(
        tx_pos=tx_pos, tx_gene=gene, tx_tile=tile, tx_cell=cell, tx_compartment=comp,
        tx_index=order.astype(np.int64), bd_x=bd_x, bd_pos=bd_pos, bd_tile=bd_tile,
        bd_index=np.arange(n_cells, dtype=np.int32), edge_tb=edge_tb, edge_pred=edge_pred,
        n_genes=n_genes, n_tiles=t * t, side=side,
    )


def drop_cross_tile_edges(edge_index: np.ndarray, tile_src: np.ndarray, tile_dst: np.ndarray):
    """Training tiles silently drop inter-tile edges
    (/root/reference/src/segger/data/partition/dataset.py:483-494)."""
    keep = tile_src[edge_index[0]] == tile_dst[edge_index[1]]
    return edge_index[:, keep]
