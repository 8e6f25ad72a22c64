"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL over
NVLink/NVSwitch on the B200 box, gloo on CPU for tests).

* Inference shards by segger's own spatial tiles: tiles (+ halo) are independent
  (/root/reference/src/segger/data/tile_dataset.py:204-246), so ranks get disjoint tile sets, no
  collective in the loop.  ``assign_tiles`` is the same best-fit-decreasing packing
  ``PartitionSampler`` uses for batches (/root/reference/src/segger/data/partition/sampler.py:11-82).
* Training is data-parallel: each rank runs forward/backward on its own tile batch and the only
  exchange step is ONE all-reduce of a flat fp32 gradient buffer (< 4 MB: latency-bound, so a single
  bucket) -- what Lightning's implicit DDP would do for the reference
  (/root/reference/src/segger/cli/segment.py:400-405), minus the bucketing machinery.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist
from torch.nn.parameter import UninitializedParameter


def assign_tiles(tile_sizes: Sequence[int], world_size: int) -> List[List[int]]:
    """Best-fit-decreasing assignment of tiles (by edge/node count) to ranks; deterministic.
    Returns ``world_size`` lists of tile ids, loads balanced within one max tile."""
    order = sorted(range(len(tile_sizes)), key=lambda i: (-int(tile_sizes[i]), i))
    loads = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        out[r].append(i)
        loads[r] += int(tile_sizes[i])
    for lst in out:
        lst.sort()
    return out


def trainable_parameters(module: torch.nn.Module) -> List[torch.nn.Parameter]:
    """Materialised parameters only (the dead bd-contains-tx conv stays lazy)."""
    return [p for p in module.parameters() if not isinstance(p, UninitializedParameter) and p.requires_grad]


class FlatGradAllReduce:
    """Gradients of all parameters live in ONE flat fp32 buffer (``p.grad`` are views into it), so a
    step's exchange is a single ``all_reduce`` (sum, then / world_size)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = list(params)
        if not self.params:
            raise ValueError("FlatGradAllReduce: no parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def reduce(self) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(dist.get_world_size())

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4


def gather_predictions(src_idx: torch.Tensor, seg_idx: torch.Tensor, max_sim: torch.Tensor):
    """End-of-predict gather of per-rank result shards to rank 0 followed by the writer's dedupe:
    keep the max-similarity row per transcript (/root/reference/src/segger/data/writer.py:199-203);
    exact-similarity ties -> lower cell id (order-independent).  CPU tensors in, CPU tensors out
    (None on ranks != 0)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        parts = [None] * dist.get_world_size()
        dist.all_gather_object(parts, (src_idx, seg_idx, max_sim))
        if dist.get_rank() != 0:
            return None
        src_idx = torch.cat([p[0] for p in parts])
        seg_idx = torch.cat([p[1] for p in parts])
        max_sim = torch.cat([p[2] for p in parts])
    # sort by (row, -sim, seg) and keep the first of every row
    order = torch.argsort(seg_idx, stable=True)
    order = order[torch.argsort(-max_sim[order], stable=True)]
    order = order[torch.argsort(src_idx[order], stable=True)]
    s, g, m = src_idx[order], seg_idx[order], max_sim[order]
    first = torch.ones_like(s, dtype=torch.bool)
    first[1:] = s[1:] != s[:-1]
    return s[first], g[first], m[first]
