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
    """Gradients of all parameters live in ONE flat fp32 buffer (``p.grad`` are views into it, so the backward
    kernels write it directly) and a step's exchange is ``n_buckets`` all-reduces (sum, then / world size) of
    contiguous slices of it.

    Overlap: the buffer is laid out in REVERSE parameter order, i.e. roughly in the order the backward produces the
    gradients; a post-accumulate hook counts a bucket's parameters and, the moment the last one is written, launches
    that bucket's all-reduce on a side stream (behind an event on the compute stream) -- the output projection and
    the last layer are on the wire while the earlier layers' backward still runs.  ``finish()`` (before the
    optimiser step) joins the side stream, reduces whatever was not launched by a hook and averages.  The payload is
    ~1.3 MB for the default model: latency-bound, so the buckets are few (default 2)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], n_buckets: int = 2, overlap: bool = True):
        self.params = list(params)
        if not self.params:
            raise ValueError("FlatGradAllReduce: no parameters")
        dev = self.params[0].device
        order = list(reversed(self.params))
        n = sum(p.numel() for p in order)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        n_buckets = max(1, min(int(n_buckets), len(order)))
        target = n / n_buckets
        self.buckets: List[List[int]] = [[0, 0, 0]]          # [start, end, n_params]
        self._bucket_of = {}
        off = 0
        for p in order:
            if off - self.buckets[-1][0] >= target and len(self.buckets) < n_buckets:
                self.buckets.append([off, off, 0])
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            b = self.buckets[-1]
            b[1], b[2] = off, b[2] + 1
            self._bucket_of[id(p)] = len(self.buckets) - 1
        self._seen = [0] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        import os
        self.overlap = bool(overlap) and dev.type == "cuda" and os.environ.get("SEGGER_B200_ALLREDUCE_OVERLAP", "1") != "0"
        self._stream = torch.cuda.Stream(dev) if self.overlap else None
        if self.overlap:
            for p in self.params:
                p.register_post_accumulate_grad_hook(self._hook)

    def _active(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _hook(self, p) -> None:
        b = self._bucket_of[id(p)]
        self._seen[b] += 1
        if self._seen[b] == self.buckets[b][2] and self._work[b] is None and self._active():
            self._launch(b)

    def _launch(self, b: int) -> None:
        s, e, _ = self.buckets[b]
        if self._stream is None:
            self._work[b] = dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, async_op=True)
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.flat.device))
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ev)
            self._work[b] = dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, async_op=True)

    def zero(self) -> None:
        self.flat.zero_()
        self._seen = [0] * len(self.buckets)
        self._work = [None] * len(self.buckets)

    def reduce(self) -> None:
        """Join the bucket all-reduces (launching those no hook has launched) and average."""
        if not self._active():
            return
        for b in range(len(self.buckets)):
            if self._work[b] is None:
                self._launch(b)
        for w in self._work:
            w.wait()                       # NCCL: makes the current stream wait for the collective
        if self._stream is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._stream)
        self.flat.div_(dist.get_world_size())

    finish = reduce

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4


def gather_rows_fixed_width(tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """All-gather of per-rank row sets of different lengths as FIXED-WIDTH tensors (no pickling through the host):
    the lengths are exchanged first, every rank pads to the longest, one ``all_gather`` per tensor, the padding is
    cut off again.  Works on NCCL (device tensors stay on the device) and gloo.  Returns the concatenation over ranks
    of every input tensor, on every rank."""
    world = dist.get_world_size()
    n = tensors[0].size(0)
    dev = tensors[0].device
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n], dtype=torch.int64, device=dev))
    counts = [int(c) for c in torch.cat(counts).tolist()]
    width = max(counts) if counts else 0
    out = []
    for t in tensors:
        pad = torch.zeros((width,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[:n] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        out.append(torch.cat([p[:c] for p, c in zip(parts, counts)]))
    return out


def gather_predictions(src_idx: torch.Tensor, seg_idx: torch.Tensor, max_sim: torch.Tensor, gene: torch.Tensor = None):
    """End-of-predict gather of the per-rank result shards followed by the writer's de-duplication: one row per
    transcript, the one with the highest similarity (/root/reference/src/segger/data/writer.py:199-203); exact
    similarity ties -> lower cell id (order-independent).  Device tensors stay on the device (fixed-width NCCL
    all-gather + ``sgb_dedupe_max``); CPU tensors (gloo, host tooling) are sorted with torch.  Every rank takes part
    in the gather; the result is returned on rank 0 (``None`` elsewhere)."""
    cols = [src_idx, seg_idx, max_sim] + ([gene] if gene is not None else [])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        cols = gather_rows_fixed_width(cols)
        if dist.get_rank() != 0:
            return None
    src_idx, seg_idx, max_sim = cols[:3]
    if src_idx.is_cuda:
        from .writer import dedupe_predictions
        g = cols[3] if gene is not None else torch.zeros_like(src_idx, dtype=torch.int32)
        s, c, m, gg = dedupe_predictions(src_idx, seg_idx, max_sim, g)
        return (s, c, m, gg) if gene is not None else (s, c, m)
    # sort by (row, -sim, seg) and keep the first of every row
    order = torch.argsort(seg_idx, stable=True)
    order = order[torch.argsort(-max_sim[order], stable=True)]
    order = order[torch.argsort(src_idx[order], stable=True)]
    first = torch.ones(order.numel(), dtype=torch.bool)
    s = src_idx[order]
    first[1:] = s[1:] != s[:-1]
    keep = order[first]
    res = (src_idx[keep], seg_idx[keep], max_sim[keep])
    return res + ((cols[3][keep],) if gene is not None else ())
