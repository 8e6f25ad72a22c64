"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one process per GPU, NCCL).

Checks, on every rank, that
  * the all-reduced gradients of a data-parallel step equal the single-process gradients of the concatenated batch
    (tiles share no edges, the loss is a sum over nodes) within 1e-6 relative;
  * tile-sharded inference gives, after the device all-gather + de-duplication, exactly the assignments a single
    process computes over all tiles.
Prints one line "MULTI_GPU_OK <world>" on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from segger_b200 import ops, tiles  # noqa: E402
from segger_b200.distributed import FlatGradAllReduce, assign_tiles, gather_predictions, trainable_parameters  # noqa: E402
from segger_b200.hetero import HeteroBatch  # noqa: E402
from segger_b200.ist_encoder import ISTEncoder  # noqa: E402
from segger_b200.lightning_model import LitISTEncoder  # noqa: E402
from segger_b200.neighbors import kdtree_neighbors  # noqa: E402
from segger_b200.synth import synth  # noqa: E402

TT, TB, PRED = ("tx", "neighbors", "tx"), ("tx", "belongs", "bd"), ("tx", "neighbors", "bd")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ts = synth(60_000, 600, seed=3, nodes_per_tile=8_000)
    ei, _ = kdtree_neighbors(ts.tx_pos, 5, 5.0, device_output=True, device=dev)
    b = HeteroBatch()
    b["tx"]["x"], b["tx"]["pos"], b["tx"]["index"] = (torch.from_numpy(a).to(dev) for a in (ts.tx_gene, ts.tx_pos, ts.tx_index))
    b["bd"]["x"], b["bd"]["pos"], b["bd"]["index"] = (torch.from_numpy(a).to(dev) for a in (ts.bd_x, ts.bd_pos, ts.bd_index))
    b[TT]["edge_index"], b[TB]["edge_index"], b[PRED]["edge_index"] = ei, torch.from_numpy(ts.edge_tb).to(dev), torch.from_numpy(ts.edge_pred).to(dev)

    # ---- training: all-reduced gradients == single-process gradients over all tiles ----------------------------------
    part = tiles.TilePartition(b, {"tx": torch.from_numpy(ts.tx_tile).to(dev), "bd": torch.from_numpy(ts.bd_tile).to(dev)}, ts.n_tiles)
    torch.manual_seed(0)
    model = ISTEncoder(ts.n_genes, 32, 64, 64, 0, 2).to(dev).eval()
    every = part.collate(list(range(ts.n_tiles)))
    with torch.no_grad():
        model(every.x_dict, {TT: every[TT]["edge_index"], TB: every[TB]["edge_index"]}, every.pos_dict, every.batch_dict)
    params = trainable_parameters(model)
    for p in params:
        dist.broadcast(p.data, 0)
    gen = torch.Generator().manual_seed(1)
    w_tx = torch.randn(60_000, 64, generator=gen).to(dev)
    w_bd = torch.randn(600, 64, generator=gen).to(dev)

    def grads(tile_ids, flat):
        flat.zero()
        batch = part.collate(tile_ids)
        out = model(batch.x_dict, {TT: batch[TT]["edge_index"], TB: batch[TB]["edge_index"]}, batch.pos_dict, batch.batch_dict)
        # weights addressed by ORIGINAL node identity so that every split of the tiles computes the same total loss
        loss = (out["tx"] * w_tx[batch["tx"]["index"]]).sum() + (out["bd"] * w_bd[batch["bd"]["index"].long()]).sum()
        loss.backward()

    flat = FlatGradAllReduce(params)
    grads(list(range(ts.n_tiles)), flat)
    single = flat.flat.clone()                      # what one process computes over all tiles (no reduce yet)
    mine = assign_tiles(part.weights("edge"), world)[rank]
    grads(mine, flat)
    flat.reduce()
    summed = flat.flat * world                      # reduce() averages; the data-parallel SUM is what must match
    err = float((summed - single).abs().max() / single.abs().max())
    assert err < 1e-6, f"rank {rank}: all-reduced gradient differs from the single-process sum by {err:.2e}"

    # ---- inference: tile-sharded assignments identical to the single-process result ----------------------------------
    torch.manual_seed(0)
    lit = LitISTEncoder(ts.n_genes, in_channels=32, n_mid_layers=0).to(dev).eval()
    g = int(np.ceil(np.sqrt(ts.n_tiles)))
    lo, hi = ts.tx_pos.min(0) - 1e-3, ts.tx_pos.max(0) + 1e-3
    boxes = tiles.square_tiles(float(lo[0]), float(lo[1]), float(hi[0]), float(hi[1]), g, g)
    ds = tiles.TilePredictSet(b, boxes, margin=20.0, grid=(g, g))
    with torch.no_grad():
        lit.predict_step(ds[0], 0, device_output=True)          # materialise lazy parameters identically everywhere
    for p in lit.parameters():
        if not isinstance(p, torch.nn.parameter.UninitializedParameter):
            dist.broadcast(p.data, 0)

    def predict(tile_ids):
        shards = []
        for t in tile_ids:
            ops.CSR_CACHE.clear()
            with torch.no_grad():
                shards.append(lit.predict_step(ds[t], 0, device_output=True))
        if not shards:
            return [torch.zeros(0, dtype=dt, device=dev) for dt in (torch.int64, torch.int64, torch.float32, torch.int32)]
        return [torch.cat([s[i] for s in shards]) for i in range(4)]

    counts = [1] * len(boxes)
    mine = assign_tiles(counts, world)[rank]
    res = gather_predictions(*predict(mine))
    if rank == 0:
        from segger_b200.writer import dedupe_predictions
        want = dedupe_predictions(*predict(list(range(len(boxes)))))
        assert torch.equal(res[0], want[0]) and res[0].numel() == 60_000
        assert torch.equal(res[1], want[1]), "tile-sharded assignments differ from the single-process result"
        assert torch.equal(res[2], want[2]) and torch.equal(res[3], want[3])
        print(f"MULTI_GPU_OK {world} grad_err={err:.2e}", flush=True)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
