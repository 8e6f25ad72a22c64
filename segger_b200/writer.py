"""B200-native post-processing of the predictions: the part of
``ISTSegmentationWriter.write_on_epoch_end`` / ``assign_transcripts_to_cells``
(/root/reference/src/segger/data/writer.py:46-253) that touches every transcript -- concatenate the per-batch
predictions, keep the best prediction per transcript, compute the per-gene similarity thresholds min(Yen, Li) and join
them back -- on ``sgb_writer.cu`` kernels instead of polars + scikit-image on the host.  The on-disk contract is the
reference's ``segger_segmentation.parquet`` (columns ``row_index``, ``segger_cell_id``, ``segger_similarity``,
``similarity_threshold``, ``converged``; read back by ``segger export``, cli/export.py:60-69).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib, ops
from ._lib import check, ptr, require_cuda, stream_ptr


def _bits(n: int) -> int:
    return max(1, int(n).bit_length())


def dedupe_predictions(row_index: Tensor, seg_idx: Tensor, max_sim: Tensor, gene: Tensor, max_row: Optional[int] = None
                       ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """writer.py:199-203 on the device: one row per transcript (highest similarity; exact ties -> lowest cell
    encoding), ascending in ``row_index``.  ``max_row``: an upper bound of the row indices (sizes the radix sort;
    read back from the data when not given)."""
    require_cuda(row_index, seg_idx, max_sim, gene)
    row = row_index.to(torch.int64).contiguous()
    seg = seg_idx.to(torch.int64).contiguous()
    sim = max_sim.to(torch.float32).contiguous()
    n, dev = row.numel(), row.device
    if n == 0:
        return row, seg, sim, gene
    if max_row is None:
        max_row = int(row.max())
    lib = _lib.load()
    order = torch.empty(n, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = ops._ws(lib.sgb_dedupe_workspace_bytes(n), dev)
    check(lib.sgb_dedupe_max(ptr(row), ptr(seg), ptr(sim), n, _bits(max_row), ptr(order), ptr(count), ptr(ws), ws.numel(),
                             stream_ptr(dev)), "dedupe_max")
    ops._count(14)
    k = int(count.item())
    from .tiles import gather_rows
    sel = order[:k]
    return gather_rows(row, sel), gather_rows(seg, sel), gather_rows(sim, sel), gather_rows(gene.contiguous(), sel)


def gene_thresholds(gene: Tensor, seg_idx: Tensor, max_sim: Tensor, n_genes: int, max_iter: int = 250):
    """writer.py:209-246 on the device -> (threshold float64 [n_genes] (NaN: no assigned transcript of that gene),
    converged bool [n_genes], count int32 [n_genes]).  threshold = min(Yen, Li); genes whose Li iteration needs more
    than ``max_iter`` callbacks get the median of the converged thresholds and ``converged = False``."""
    require_cuda(gene, seg_idx, max_sim)
    if gene.dtype not in (torch.int32, torch.int64):
        gene = gene.to(torch.int64)
    gene = gene.contiguous()
    seg = seg_idx.to(torch.int64).contiguous()
    sim = max_sim.to(torch.float32).contiguous()
    n, dev = gene.numel(), gene.device
    lib = _lib.load()
    yen = torch.empty(n_genes, dtype=torch.float64, device=dev)
    li = torch.empty(n_genes, dtype=torch.float64, device=dev)
    iters = torch.empty(n_genes, dtype=torch.int32, device=dev)
    counts = torch.empty(n_genes, dtype=torch.int32, device=dev)
    ws = ops._ws(lib.sgb_gene_threshold_workspace_bytes(n, n_genes), dev)
    check(lib.sgb_gene_thresholds(ptr(gene), gene.element_size(), ptr(seg), ptr(sim), n, n_genes, int(max_iter), ptr(yen),
                                  ptr(li), ptr(iters), ptr(counts), ptr(ws), ws.numel(), stream_ptr(dev)), "gene_thresholds")
    ops._count(24)
    present = counts > 0
    converged = present & (iters >= 0)
    thr = torch.minimum(yen, li)
    failed = present & ~converged
    if bool(failed.any()):
        ok = thr[converged]
        # np.quantile(., 0.5) of the converged thresholds (writer.py:243): linear interpolation = plain median average
        glob = torch.quantile(ok, 0.5) if ok.numel() else torch.tensor(float("nan"), dtype=torch.float64, device=dev)
        thr = torch.where(failed, glob, thr)
    thr = torch.where(present, thr, torch.full_like(thr, float("nan")))
    return thr, converged, counts


def assign_transcripts_to_cells(predictions: Sequence[Sequence[Tensor]], cell_ids: Optional[Sequence] = None,
                                n_genes: Optional[int] = None, device=None) -> dict:
    """``ISTSegmentationWriter.assign_transcripts_to_cells`` (writer.py:132-253).  ``predictions``: the per-batch
    4-tuples ``predict_step`` returns (row index, cell encoding or -1, similarity, gene id).  ``cell_ids[i]`` = id of
    the boundary with encoding i (the ``obs`` join, :178-196).  Returns numpy columns ``row_index``,
    ``segger_cell_id`` (None where unassigned), ``segger_similarity``, ``similarity_threshold`` (NaN where the gene
    has no assigned transcript), ``converged``."""
    device = torch.device(device if device is not None else "cuda")
    row = torch.cat([b[0].to(device, non_blocking=True).to(torch.int64) for b in predictions])
    seg = torch.cat([b[1].to(device, non_blocking=True).to(torch.int64) for b in predictions])
    sim = torch.cat([b[2].to(device, non_blocking=True).to(torch.float32) for b in predictions])
    gene = torch.cat([b[3].to(device, non_blocking=True).to(torch.int64) for b in predictions])
    row, seg, sim, gene = dedupe_predictions(row, seg, sim, gene)
    if n_genes is None:
        n_genes = int(gene.max()) + 1 if gene.numel() else 1
    thr, conv, counts = gene_thresholds(gene, seg, sim, n_genes)
    g = gene.clamp(0, n_genes - 1)
    col_thr = thr.index_select(0, g)            # the left join on the gene id (writer.py:249-252): index plumbing
    col_conv = conv.index_select(0, g)
    seg_h = seg.cpu().numpy()
    cell = np.full(seg_h.shape[0], None, dtype=object)
    ok = seg_h >= 0
    if cell_ids is not None:
        ids = np.asarray(cell_ids, dtype=object)
        cell[ok] = ids[seg_h[ok]]
    else:
        cell[ok] = seg_h[ok]
    return {"row_index": row.cpu().numpy(), "segger_cell_id": cell, "segger_similarity": sim.cpu().numpy(),
            "similarity_threshold": col_thr.cpu().numpy(), "converged": col_conv.cpu().numpy()}


def write_segmentation(columns: dict, path) -> None:
    """``segger_segmentation.parquet`` with the reference's column contract (writer.py:99-107, cli/export.py:60-69)."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    cell = columns["segger_cell_id"]
    table = pa.table({
        "row_index": pa.array(columns["row_index"]),
        "segger_cell_id": pa.array([None if c is None else c for c in cell]),
        "segger_similarity": pa.array(columns["segger_similarity"]),
        "similarity_threshold": pa.array(columns["similarity_threshold"], from_pandas=True),
        "converged": pa.array(columns["converged"]),
    })
    pq.write_table(table, str(path))
