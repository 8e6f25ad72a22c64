"""CPU restatement of segger's tile slicing / batching -- TEST INFRASTRUCTURE ONLY.

Follows, with plain torch CPU indexing on dictionaries of tensors (the PyG ``HeteroData`` container itself is a
third-party class, absent here -- PARITY UNPINNED for its ``subgraph`` / collate, restated from PyG 2.7 semantics):

  partition_ref      PartitionDataset._get_permutation / _permute_edge_store   data/partition/dataset.py:375-506
                     (``torch.argsort`` there is unstable; stable here, as in the product)
  get_tile_ref       PartitionDataset.__getitem__                               data/partition/dataset.py:512-579
  collate_ref        torch_geometric Batch.from_data_list of such tiles (node stores concatenated, edge_index shifted by the
                     cumulative node counts of its source / destination types, ``batch`` vector per node type)
  subset_ref         TilePredictDataset._subset                                 data/tile_dataset.py:218-246
                     + HeteroData.subgraph = bipartite_subgraph(relabel_nodes=True) per edge type
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch

EdgeType = Tuple[str, str, str]


def partition_ref(nodes: Dict[str, Dict[str, torch.Tensor]], edges: Dict[EdgeType, torch.Tensor],
                  labels: Dict[str, torch.Tensor], n_tiles: int):
    """-> (permuted node stores, permuted edge_index per type, node_indptr, edge_indptr, node_perm)."""
    perm, inv, indptr, out_nodes = {}, {}, {}, {}
    for nt, store in nodes.items():
        lab = labels[nt].long()
        p = torch.argsort(lab, stable=True)                                        # :385 (stable here)
        sizes = torch.bincount(lab[p], minlength=n_tiles)                          # :389-392
        indptr[nt] = torch.cat((torch.tensor([0]), torch.cumsum(sizes, 0)))        # :398-401
        perm[nt] = p
        iv = torch.empty_like(p)
        iv[p] = torch.arange(p.numel())                                            # :449-458
        inv[nt] = iv
        out_nodes[nt] = {k: v[p] for k, v in store.items()}
    out_edges, e_indptr = {}, {}
    for et, ei in edges.items():
        src, _, dst = et
        ei = ei.long()
        new = torch.stack([inv[src][ei[0]], inv[dst][ei[1]]])                      # :459-462
        ls = labels[src].long()[perm[src]][new[0]]                                 # :476
        ld = labels[dst].long()[perm[dst]][new[1]]
        order = torch.argsort(ls, stable=True)                                     # :478
        ls, ld = ls[order], ld[order]
        mask = ls == ld                                                            # :483
        out_edges[et] = new[:, order][:, mask]                                     # :488
        sizes = torch.bincount(ls[mask], minlength=n_tiles)                        # :497-500
        e_indptr[et] = torch.cat((torch.tensor([0]), torch.cumsum(sizes, 0)))
    return out_nodes, out_edges, indptr, e_indptr, perm


def get_tile_ref(nodes, edges, indptr, e_indptr, index: int):
    """:512-579 -> (node stores of the tile, tile-local edge_index per type)."""
    n = {nt: {k: v[indptr[nt][index]:indptr[nt][index + 1]] for k, v in store.items()} for nt, store in nodes.items()}
    e = {}
    for et, ei in edges.items():
        src, _, dst = et
        part = ei[:, e_indptr[et][index]:e_indptr[et][index + 1]].clone()
        part[0] -= indptr[src][index]
        part[1] -= indptr[dst][index]
        e[et] = part
    return n, e


def collate_ref(tiles: Sequence[Tuple[dict, dict]]):
    """PyG ``Batch.from_data_list``: concatenate node stores, shift edges, build ``batch``."""
    node_types = list(tiles[0][0].keys())
    out_n = {nt: {} for nt in node_types}
    offs = {nt: [0] for nt in node_types}
    for nt in node_types:
        for k in tiles[0][0][nt]:
            out_n[nt][k] = torch.cat([t[0][nt][k] for t in tiles])
        sizes = [next(iter(t[0][nt].values())).shape[0] for t in tiles]
        for s in sizes:
            offs[nt].append(offs[nt][-1] + s)
        out_n[nt]["batch"] = torch.cat([torch.full((s,), i, dtype=torch.long) for i, s in enumerate(sizes)]) if sizes else torch.zeros(0, dtype=torch.long)
    out_e = {}
    for et in tiles[0][1]:
        src, _, dst = et
        parts = []
        for i, t in enumerate(tiles):
            ei = t[1][et].clone()
            ei[0] += offs[src][i]
            ei[1] += offs[dst][i]
            parts.append(ei)
        out_e[et] = torch.cat(parts, dim=1)
    return out_n, out_e


def subset_ref(nodes, edges, bounds, margin: float):
    """tile_dataset.py:218-246 -> (node stores incl. predict_mask, relabelled edge_index per type, kept edge ids)."""
    x0, y0, x1, y1 = bounds
    outer = (x0 - margin, y0 - margin, x1 + margin, y1 + margin)     # shapely box.buffer(margin).bounds
    inner = (x0, y0, x1, y1)
    subset, out_n = {}, {}
    for nt, store in nodes.items():
        pos = store["pos"]
        keep = ((pos[:, 0] >= outer[0]) & (pos[:, 0] < outer[2]) & (pos[:, 1] >= outer[1]) & (pos[:, 1] < outer[3]))
        sub = keep.nonzero().squeeze(1)                                # :232-238 (_chunked_nonzero)
        subset[nt] = sub
        out_n[nt] = {k: v[sub] for k, v in store.items()}
        ps = pos[sub]
        out_n[nt]["predict_mask"] = ((ps[:, 0] >= inner[0]) & (ps[:, 0] <= inner[2]) &
                                     (ps[:, 1] >= inner[1]) & (ps[:, 1] <= inner[3]))     # :239-244
    out_e, kept = {}, {}
    for et, ei in edges.items():                                      # HeteroData.subgraph -> bipartite_subgraph
        src, _, dst = et
        ei = ei.long()
        n_s, n_d = nodes[src]["pos"].shape[0], nodes[dst]["pos"].shape[0]
        ms, md = torch.zeros(n_s, dtype=torch.bool), torch.zeros(n_d, dtype=torch.bool)
        ms[subset[src]] = True
        md[subset[dst]] = True
        emask = ms[ei[0]] & md[ei[1]]
        rs, rd = torch.full((n_s,), -1, dtype=torch.long), torch.full((n_d,), -1, dtype=torch.long)
        rs[subset[src]] = torch.arange(subset[src].numel())
        rd[subset[dst]] = torch.arange(subset[dst].numel())
        sel = ei[:, emask]
        out_e[et] = torch.stack([rs[sel[0]], rd[sel[1]]])
        kept[et] = emask.nonzero().squeeze(1)
    return out_n, out_e, kept
