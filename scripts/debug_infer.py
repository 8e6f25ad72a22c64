"""Where does the cfg-3 inference leg spend its time? (run on the GPU box)"""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from segger_b200 import ops
from segger_b200.geometry import PackedPolygons, points_in_polygons
from segger_b200.hetero import HeteroBatch
from segger_b200.lightning_model import LitISTEncoder
from segger_b200.neighbors import kdtree_neighbors
from segger_b200.synth import synth
from segger_b200 import tiles
from segger_b200.tiles import TilePredictSet, square_tiles
TT, TB, PRED = bench.TT, bench.TB, bench.PRED
dev = torch.device("cuda")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
ts = synth(N, N // 100, seed=0, pred_edges=False)
pos = torch.from_numpy(ts.tx_pos).to(dev)
ei, _ = kdtree_neighbors(pos, 5, 5.0, device_output=True, device=dev)
ang = np.linspace(0, 2 * np.pi, 16, endpoint=False)
c = ts.bd_pos.astype(np.float64)
verts = np.stack([c[:, None, 0] + 6.5 * 1.05 * np.cos(ang)[None], c[:, None, 1] + 6.5 * 1.05 * np.sin(ang)[None]], -1).reshape(-1, 2)
ep = points_in_polygons(pos, PackedPolygons(verts, np.arange(N // 100 + 1, dtype=np.int64) * 16), device=dev, device_output=True)
b = HeteroBatch()
b["tx"]["x"], b["tx"]["pos"], b["tx"]["index"] = torch.from_numpy(ts.tx_gene).to(dev), pos, torch.from_numpy(ts.tx_index).to(dev)
b["bd"]["x"], b["bd"]["pos"], b["bd"]["index"] = torch.from_numpy(ts.bd_x).to(dev), torch.from_numpy(ts.bd_pos).to(dev), torch.from_numpy(ts.bd_index).to(dev)
b[TT]["edge_index"], b[TB]["edge_index"], b[PRED]["edge_index"] = ei, torch.from_numpy(ts.edge_tb).to(dev), ep
nt = max(1, int(math.ceil(math.sqrt(N / 50_000))))
lo, hi = ts.tx_pos.min(0) - 1e-3, ts.tx_pos.max(0) + 1e-3
boxes = square_tiles(float(lo[0]), float(lo[1]), float(hi[0]), float(hi[1]), nt, nt)
torch.manual_seed(0)
lit = LitISTEncoder(ts.n_genes, in_channels=128, n_mid_layers=0).to(dev).eval()
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

ds = TilePredictSet(b, boxes, margin=20.0, grid=(nt, nt))
print("cut one tile (indexed): %.2f ms" % timeit(lambda: ds[nt + 1], 20))
ds0 = TilePredictSet(b, boxes, margin=20.0)
print("cut one tile (full scan): %.2f ms" % timeit(lambda: ds0[nt + 1], 5))
tl = [ds[i] for i in range(16)]
print("concat 16 tiles: %.2f ms" % timeit(lambda: tiles.collate_tiles(tl), 5))
bb = tiles.collate_tiles(tl)
n16 = int(bb["tx"]["x"].size(0))
def pred():
    ops.CSR_CACHE.clear()
    with torch.no_grad():
        lit.predict_step(bb, 0, device_output=True)
t = timeit(pred, 10)
print("predict_step on a 16-tile batch (%d tx, %d tt edges): %.2f ms -> %.1f M tx/s" % (n16, bb[TT]["edge_index"].size(1), t, n16 / t / 1e3))
def fwd():
    ops.CSR_CACHE.clear()
    with torch.no_grad():
        lit(bb)
print("  forward only: %.2f ms" % timeit(fwd, 10))
# the same number of transcripts as one contiguous region (no tile structure): first n16 nodes of the sorted data
import torch.cuda.profiler
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pred(); torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:14]
for r in rows:
    print("   %8.1f us x%3d  %s" % (r.device_time_total, r.count, r.key[:90]))
