"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: tile assignment, the flat-gradient
all-reduce that makes the replicas agree with the single-process sum, and the end-of-predict gather
+ max-similarity dedupe.  The device kernels themselves are covered by the -m gpu tests."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from segger_b200.distributed import FlatGradAllReduce, assign_tiles, gather_predictions, trainable_parameters


def test_assign_tiles_balanced_and_deterministic():
    sizes = [50, 10, 40, 30, 20, 45, 5, 25]
    a = assign_tiles(sizes, 3)
    assert a == assign_tiles(sizes, 3)
    assert sorted(i for r in a for i in r) == list(range(8))
    loads = [sum(sizes[i] for i in r) for r in a]
    assert max(loads) - min(loads) <= max(sizes)
    assert assign_tiles(sizes, 1) == [list(range(8))]
    assert assign_tiles([], 2) == [[], []]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                    # identical replicas
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.GELU(), torch.nn.Linear(5, 3))
        flat = FlatGradAllReduce(trainable_parameters(model))
        g = torch.Generator().manual_seed(100 + rank)           # each rank: its own tile batch
        x, y = torch.randn(16, 6, generator=g), torch.randn(16, 3, generator=g)
        flat.zero()
        ((model(x) - y) ** 2).mean().backward()
        local = flat.flat.clone()
        flat.reduce()
        # reference: average of the per-rank gradients
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        ok_grad = torch.allclose(flat.flat, torch.stack(gathered).mean(0), atol=1e-7)
        views_alias = all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in model.parameters())
        # predictions: rank r owns transcripts [r*4, r*4+6) -> halo overlap of 2 with the neighbour
        src = torch.arange(rank * 4, rank * 4 + 6)
        seg = torch.full((6,), rank)
        sim = torch.full((6,), 0.5 + 0.1 * rank)
        sim[:2] += 0.3 * (1 - rank)
        res = gather_predictions(src, seg, sim)
        if rank == 0:
            s, gseg, m = res
            out.put((ok_grad, views_alias, s.tolist(), gseg.tolist(), [round(float(v), 3) for v in m]))
        else:
            assert res is None
            out.put((ok_grad, views_alias))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_grad_allreduce_and_prediction_gather_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=100) for _ in range(2)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    full = [r for r in results if len(r) == 5][0]
    assert all(r[0] and r[1] for r in results)
    ok, alias, s, seg, m = full
    assert s == list(range(10))                               # every transcript once after the dedupe
    # overlap rows 4,5: rank 0 has sim 0.5, rank 1 has 0.6 -> rank 1 wins
    assert seg == [0, 0, 0, 0, 1, 1, 1, 1, 1, 1]
    assert m[4] == 0.6 and m[0] == 0.8


def test_gather_predictions_single_process_dedupe_ties():
    src = torch.tensor([3, 1, 3, 2, 1])
    seg = torch.tensor([7, 5, 4, 9, 6])
    sim = torch.tensor([0.5, 0.9, 0.5, 0.1, 0.2])
    s, g, m = gather_predictions(src, seg, sim)
    assert s.tolist() == [1, 2, 3] and g.tolist() == [5, 9, 4]   # exact tie on row 3 -> lower cell id
