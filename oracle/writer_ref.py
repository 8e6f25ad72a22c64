"""CPU restatement of segger's writer post-processing -- TEST INFRASTRUCTURE ONLY.

Follows ISTSegmentationWriter.assign_transcripts_to_cells (/root/reference/src/segger/data/writer.py:132-253) with
numpy instead of polars, and restates the two scikit-image functions it calls -- scikit-image (pixi.lock: 0.25.x) is a
third-party dependency absent from /root/reference and from this image, so ``threshold_yen`` / ``threshold_li`` below
follow the published implementations (skimage/filters/thresholding.py) -- PARITY UNPINNED for those two.

Two stated deviations from a literal run of the reference:
  * polars' ``sort`` is not stable and ``unique(keep='first')`` after it picks an unspecified row among exact
    (row_index, similarity) ties; here ties go to the lowest cell encoding (the product does the same);
  * the reference down-samples genes with more than 10 M assigned transcripts with a polars sampler (writer.py:225-227);
    here (and in the product) every value is used.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def threshold_yen(arr: np.ndarray, nbins: int = 256) -> float:
    """skimage.filters.threshold_yen on a 1-D float array."""
    arr = np.asarray(arr)
    counts, edges = np.histogram(arr.astype(np.float64), bins=nbins, range=(float(arr.min()), float(arr.max())))
    centers = (edges[:-1] + edges[1:]) / 2.0
    pmf = counts.astype(np.float64) / counts.sum()
    P1 = np.cumsum(pmf)
    P1_sq = np.cumsum(pmf ** 2)
    P2_sq = np.cumsum(pmf[::-1] ** 2)[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        crit = np.log(((P1_sq[:-1] * P2_sq[1:]) ** -1) * (P1[:-1] * (1.0 - P1[:-1])) ** 2)
    return float(centers[int(np.argmax(crit))])


class NotConverged(Exception):
    pass


def threshold_li(arr: np.ndarray, max_iter: int = 250) -> float:
    """skimage.filters.threshold_li with the iteration cap of data/utils/threshold.py:3-11 (StopIteration after more
    than ``max_iter`` callbacks; the first callback reports the initial guess)."""
    image = np.asarray(arr, dtype=np.float32)
    image = image[np.isfinite(image)]
    if np.all(image == image.flat[0]):
        return float(image.flat[0])
    image_min = np.min(image)
    image = image - image_min                                    # float32, as skimage does in place
    tolerance = float(np.min(np.diff(np.unique(image))) / 2)
    w = image.astype(np.float64)
    t_next = float(np.mean(w))
    t_curr = -2 * tolerance
    n_iter = 1                                                   # callback on the initial guess
    while abs(t_next - t_curr) > tolerance:
        t_curr = t_next
        fore = w > t_curr
        mean_fore, mean_back = float(np.mean(w[fore])), float(np.mean(w[~fore]))
        if mean_back == 0:
            break
        t_next = (mean_back - mean_fore) / (np.log(mean_back) - np.log(mean_fore))
        n_iter += 1
        if n_iter > max_iter:
            raise NotConverged
    return float(t_next + np.float64(image_min))


def dedupe_ref(row_index: np.ndarray, seg: np.ndarray, sim: np.ndarray, gene: np.ndarray):
    """writer.py:199-203: one row per row_index, highest similarity first (ties: lowest cell encoding)."""
    order = np.lexsort((seg, -sim.astype(np.float64), row_index))
    r = row_index[order]
    first = np.ones(r.shape[0], dtype=bool)
    first[1:] = r[1:] != r[:-1]
    keep = order[first]
    return row_index[keep], seg[keep], sim[keep], gene[keep]


def gene_thresholds_ref(gene: np.ndarray, seg: np.ndarray, sim: np.ndarray, max_iter: int = 250):
    """writer.py:209-246 -> {gene: (threshold, converged)} over genes with at least one assigned transcript."""
    out, failed = {}, []
    assigned = seg >= 0
    for g in np.unique(gene[assigned]):
        arr = sim[assigned & (gene == g)]
        try:
            tye = threshold_yen(arr) if arr.max() > arr.min() else float(arr[0])
            tli = threshold_li(arr, max_iter)
            out[int(g)] = (min(tye, tli), True)
        except NotConverged:
            failed.append(int(g))
    if failed:
        glob = float(np.quantile([t for t, _ in out.values()], 0.5))
        for g in failed:
            out[g] = (glob, False)
    return out


def assign_transcripts_to_cells_ref(predictions: Sequence[Sequence[np.ndarray]], cell_ids: Optional[Sequence] = None):
    """writer.py:152-253 -> dict of columns: row_index, segger_cell_id, segger_similarity, similarity_threshold, converged."""
    row = np.concatenate([np.asarray(b[0]) for b in predictions]).astype(np.int64)
    seg = np.concatenate([np.asarray(b[1]) for b in predictions]).astype(np.int64)
    sim = np.concatenate([np.asarray(b[2]) for b in predictions]).astype(np.float32)
    gene = np.concatenate([np.asarray(b[3]) for b in predictions]).astype(np.int64)
    row, seg, sim, gene = dedupe_ref(row, seg, sim, gene)
    thr = gene_thresholds_ref(gene, seg, sim)
    t = np.array([thr[int(g)][0] if int(g) in thr else np.nan for g in gene])
    c = np.array([thr[int(g)][1] if int(g) in thr else False for g in gene])
    has = np.array([int(g) in thr for g in gene])
    cell = np.full(row.shape[0], None, dtype=object)
    ok = seg >= 0
    cell[ok] = [cell_ids[i] for i in seg[ok]] if cell_ids is not None else seg[ok]
    return {"row_index": row, "segger_cell_id": cell, "segger_similarity": sim, "similarity_threshold": t, "converged": c,
            "has_threshold": has}
