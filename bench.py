#!/usr/bin/env python
"""Benchmark of the segger hot path on B200 (contract: see the task statement / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg4]

A *step* = one training pass of the hot path over one synthetic tile batch: CSR build, ISTEncoder
forward (input stage, hetero GATv2 layers, output projection, normalise), a linear synthetic loss,
the fused deterministic backward, the data-parallel gradient all-reduce (N > 1) and an Adam step.
Workload (N=1): BASELINE.json configs[1] -- 1M transcripts / 10k cells, k=5, 2-layer hetero GATv2,
hidden=64, heads=2.  Under torchrun every rank owns its own 1M-transcript tile set (weak scaling).

One JSON line is printed by rank 0.  `value` = GATv2 edge-layers/s (sum over layers and live edge
types of E, divided by step time) with inputs resident in HBM; `e2e` = the same with pinned-host
inputs copied H2D (prefetched one step ahead on a copy stream) and the loss read back D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_tx, n_cells, k, in_channels, hidden, out, n_mid_layers, heads)
    "cfg1": (50_000, 500, 5, 128, 64, 64, 0, 2),
    "cfg2": (1_000_000, 10_000, 5, 128, 64, 64, 0, 2),
    "cfg4": (1_000_000, 10_000, 20, 128, 128, 128, 1, 4),
}
WORKLOADS_K = {}          # n_tx -> k of the workload being run (filled in main)
METRIC = "gatv2_fwd_bwd_edge_layers_per_sec"
UNIT = "edge-layers/s"
TT = ("tx", "neighbors", "tx")
TB = ("tx", "belongs", "bd")
PRED = ("tx", "neighbors", "bd")


_OUT_FD = None


def capture_stdout():
    """Route file descriptor 1 to stderr for the whole run (NCCL prints its version banner on stdout from C) and keep
    the real stdout for the single JSON line."""
    global _OUT_FD
    if _OUT_FD is None:
        sys.stdout.flush()
        _OUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _OUT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_OUT_FD, data)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------------
def build_workload(name: str, seed: int, device):
    """Synthetic tile set -> pinned host tensors (what a DataLoader would hand over) + sizes."""
    from segger_b200.neighbors import kdtree_neighbors
    from segger_b200.synth import synth
    n_tx, n_cells, k, *_ = WORKLOADS[name]
    ts = synth(n_tx, n_cells, seed=seed)
    # graph construction with the product's GPU kNN, then the training-tile rule: drop cross-tile edges
    ei, _ = kdtree_neighbors(ts.tx_pos, k, 5.0, device_output=True, device=device)
    tile = torch.from_numpy(ts.tx_tile).to(device)
    keep = tile[ei[0]] == tile[ei[1]]
    ei_train = ei[:, keep].contiguous()
    host = {
        "tx_x": torch.from_numpy(ts.tx_gene), "tx_pos": torch.from_numpy(ts.tx_pos),
        "tx_batch": torch.from_numpy(ts.tx_tile), "bd_x": torch.from_numpy(ts.bd_x),
        "bd_pos": torch.from_numpy(ts.bd_pos), "bd_batch": torch.from_numpy(ts.bd_tile),
        "e_tt": ei_train.cpu(), "e_tb": torch.from_numpy(ts.edge_tb), "e_pred": torch.from_numpy(ts.edge_pred),
        "e_tt_full": ei.cpu(), "bd_index": torch.from_numpy(ts.bd_index), "tx_index": torch.from_numpy(ts.tx_index),
    }
    host = {k_: v.pin_memory() for k_, v in host.items()}
    return ts, host


def to_device(host, device, keys):
    return {k: host[k].to(device, non_blocking=True) for k in keys}


TRAIN_KEYS = ("tx_x", "tx_pos", "tx_batch", "bd_x", "bd_pos", "bd_batch", "e_tt", "e_tb")
PRED_KEYS = ("tx_x", "tx_pos", "tx_batch", "bd_x", "bd_pos", "bd_batch", "e_tt_full", "e_tb", "e_pred", "bd_index",
             "tx_index")


def model_inputs(d, tt_key="e_tt"):
    x = {"tx": d["tx_x"], "bd": d["bd_x"]}
    pos = {"tx": d["tx_pos"], "bd": d["bd_pos"]}
    bat = {"tx": d["tx_batch"], "bd": d["bd_batch"]}
    edges = {TT: d[tt_key], TB: d["e_tb"]}
    if "e_pred" in d:
        edges[PRED] = d["e_pred"]
    return x, edges, pos, bat


def gat_bytes(n_src_touched, n_dst, n_src, E, H, C):
    """Algorithmic (compulsory) bytes of one GATv2Conv, SURVEY.md section 8d / BASELINE.md section 3."""
    F = H * C
    fwd = 4 * F * (n_src_touched + n_dst) + 4 * F * n_dst + 4 * E + 4 * (n_dst + 1) + 8 * H * n_dst
    bwd = 8 * F * (n_src_touched + n_dst) + 4 * F * n_dst + 8 * H * n_dst + 16 * E + 4 * (n_dst + 1) + 4 * (n_src + 1)
    return fwd, bwd


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle (CPU restatement of the reference path) on host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(workload: str, seed: int = 0):
    """One bounded sample of the workload on the CPU: a single <=50k-transcript tile (the reference's
    own tile size, tiling_nodes_per_tile=50_000) with the workload's model, forward+backward."""
    from oracle import neighbors_ref
    from oracle.ist_encoder_ref import ISTEncoderRef
    from segger_b200.synth import synth
    _, _, k, in_c, hid, out_c, n_mid, heads = WORKLOADS[workload]
    torch.set_num_threads(os.cpu_count() or 1)
    ts = synth(50_000, 500, seed=seed)
    ei, _ = neighbors_ref.kdtree_neighbors(ts.tx_pos, k, 5.0)
    x = {"tx": torch.from_numpy(ts.tx_gene), "bd": torch.from_numpy(ts.bd_x)}
    pos = {"tx": torch.from_numpy(ts.tx_pos), "bd": torch.from_numpy(ts.bd_pos)}
    bat = {"tx": torch.from_numpy(ts.tx_tile), "bd": torch.from_numpy(ts.bd_tile)}
    edges = {TT: ei, TB: torch.from_numpy(ts.edge_tb)}
    torch.manual_seed(0)
    model = ISTEncoderRef(ts.n_genes, ts.bd_x.shape[1], in_c, hid, out_c, n_mid, heads).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(1)
    t_tx, t_bd = torch.randn(50_000, out_c, generator=g), torch.randn(500, out_c, generator=g)
    n_layers = n_mid + 2
    edge_layers = n_layers * (ei.size(1) + ts.edge_tb.shape[1])

    def step():
        opt.zero_grad(set_to_none=True)
        out = model(x, edges, pos, bat)
        loss = (out["tx"] * t_tx).sum() / 50_000 + (out["bd"] * t_bd).sum() / 500
        loss.backward()
        opt.step()
        return float(loss.detach())

    sample = (f"one 50k-transcript / 500-cell tile of the workload's model (k={k}, {n_layers} layers, "
              f"hidden={hid}, heads={heads}), fwd+bwd+Adam, plain-torch CPU restatement of the PyG path "
              f"(oracle/), {edge_layers} edge-layers per step")
    return step, edge_layers, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, edge_layers, sample = cpu_reference_step_factory(args.workload)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    v = edge_layers / dt
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": max(1, min(args.warmup, 2)), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS_DESC[args.workload], "timing": "host wall clock, CPU only"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


WORKLOADS_DESC = {
    "cfg1": "BASELINE configs[0]: 50k transcripts / 500 cells single tile, kNN k=5, 2-layer hetero GATv2 hidden=64 heads=2, training step",
    "cfg2": "BASELINE configs[1]: 1M transcripts / 10k cells tile set (25 tiles), kNN k=5, 2-layer hetero GATv2 hidden=64 heads=2, training step",
    "cfg4": "BASELINE configs[3]: 1M transcripts / 10k cells, kNN k=20, 3-layer hetero GATv2 hidden=128 heads=4, training step",
}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from segger_b200 import ops
    from segger_b200.distributed import FlatGradAllReduce, trainable_parameters
    from segger_b200.ist_encoder import ISTEncoder
    from segger_b200.lightning_model import LitISTEncoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback; use --impl reference)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set: keep stdout to the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    n_tx, n_cells, k, in_c, hid, out_c, n_mid, heads = WORKLOADS[args.workload]
    n_layers = n_mid + 2
    WORKLOADS_K[n_tx] = k

    ts, host = build_workload(args.workload, seed=rank, device=device)
    torch.manual_seed(0)
    lit = LitISTEncoder(ts.n_genes, in_channels=in_c, hidden_channels=hid, out_channels=out_c, n_mid_layers=n_mid,
                        n_heads=heads).to(device)
    model = lit.model.train()
    dev_in = to_device(host, device, TRAIN_KEYS)
    with torch.no_grad():      # materialise lazy parameters (first call), as Lightning's dry run would
        model(*model_inputs(dev_in))
    params = trainable_parameters(model)
    if world > 1:              # identical replicas
        for p in params:
            dist.broadcast(p.data, 0)
    flat = FlatGradAllReduce(params)
    opt = torch.optim.Adam(params, lr=1e-3, fused=True)
    g = torch.Generator(device="cpu").manual_seed(1)
    t_tx = torch.randn(n_tx, out_c, generator=g).to(device)
    t_bd = torch.randn(n_cells, out_c, generator=g).to(device)
    E_tt, E_tb = host["e_tt"].size(1), host["e_tb"].size(1)
    edge_layers = n_layers * (E_tt + E_tb)

    def train_step(d):
        ops.CSR_CACHE.clear()          # every training batch is a new graph: the CSR build is in the step
        flat.zero()
        out = model(*model_inputs(d))
        loss = (out["tx"] * t_tx).sum() / n_tx + (out["bd"] * t_bd).sum() / n_cells
        loss.backward()
        flat.reduce()
        opt.step()
        return loss

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident measurement -----------------------------------------------------------
    for _ in range(W):
        train_step(dev_in)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.LAUNCHES
    ms_total = timed(lambda: train_step(dev_in), K)
    launches = (ops.LAUNCHES - l0)
    clocks = sampler.stop() if rank == 0 else {}
    ms_step = ms_total / K
    value = world * edge_layers / (ms_step * 1e-3)

    # ---- end to end: pinned host inputs -> H2D every step, loss -> D2H every step --------------
    h2d = sum(host[k_].numel() * host[k_].element_size() for k_ in TRAIN_KEYS)

    # Input pipeline of the end-to-end arm: what a DataLoader with pinned memory does -- the H2D copy of batch
    # i+1 is enqueued on a copy stream while batch i computes.  Every step still copies its own inputs inside
    # the timed region (K copies for K steps; the first one is not overlapped) and reads its loss back.
    copy_stream = torch.cuda.Stream(device)
    compute_stream = torch.cuda.current_stream(device)
    # two preallocated device input sets (double buffering): no allocator traffic across streams inside the loop
    bufs = [{k_: torch.empty_like(host[k_], device=device) for k_ in TRAIN_KEYS} for _ in range(2)]
    done = [None, None]

    def fetch(slot):
        with torch.cuda.stream(copy_stream):
            if done[slot] is not None:
                copy_stream.wait_event(done[slot])       # the step that last read this buffer set has finished
            for k_ in TRAIN_KEYS:
                bufs[slot][k_].copy_(host[k_], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_run(steps):
        ev = fetch(0)
        for i in range(steps):
            slot = i % 2
            compute_stream.wait_event(ev)
            if i + 1 < steps:
                ev = fetch((i + 1) % 2)
            loss = train_step(bufs[slot])
            float(loss.item())             # D2H read of the step's result
            done[slot] = torch.cuda.Event()
            done[slot].record(compute_stream)

    e2e_run(2)
    ms_e2e = timed(lambda: e2e_run(K), 1) / K
    e2e_value = world * edge_layers / (ms_e2e * 1e-3)

    # ---- the same step with segger's own losses instead of the linear synthetic loss -----------
    loss_step = None
    try:
        loss_step = losses_step_time(lit, dev_in, n_tx, n_cells, device, flat, opt, timed, K, ms_step)
    except Exception as e:  # noqa: BLE001 -- secondary measurement: never take the headline line down with it
        loss_step = {"error": f"{type(e).__name__}: {e}"}

    # ---- roofline of the dominant kernels, timed alone with CUDA events, L2 flushed between ------
    hbm_peak, peak_src = peaks()
    roof = kernel_rooflines(model, dev_in, n_tx, n_cells, heads, hid, device, hbm_peak, peak_src)

    # ---- segmentation throughput (predict_step over the full graph incl. cross-tile edges) -------
    seg = None
    if rank == 0 or world > 1:
        seg = segmentation_throughput(lit, host, ts, device, world, timed)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        step, el, sample = cpu_reference_step_factory(args.workload)
        step()
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter(); step(); best = min(best, time.perf_counter() - t0)
        cpu = {"value": el / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": sample + "; best of 3 after 1 warm-up"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": WORKLOADS_DESC[args.workload],
            "per_gpu": {"n_tx": n_tx, "n_cells": n_cells, "E_tt": E_tt, "E_tb": E_tb, "layers": n_layers,
                        "edge_layers_per_step": edge_layers, "tiles": ts.n_tiles},
            "step": "CSR build + ISTEncoder fwd + linear synthetic loss + fused bwd + grad all-reduce + fused Adam",
            "l2": "inputs larger than L2 (per-layer activations 0.5-1.5 GB vs 126 MB L2); kernel-alone timings flush L2",
            "parallelism": f"dp{world} (one tile set per rank, one flat-gradient all-reduce of {flat.nbytes} B per step)",
            "timing": "CUDA events on the launching stream, barrier+synchronize on both sides, max over ranks",
            "roofline": "`roofline` = the slowest fused message-passing launch group (HBM-bound; the kernels BASELINE.json's "
                        "metric names); `roofline_kernels` adds the other one and the largest projection GEMM against the "
                        "tensor pipe",
        },
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": roof["dominant"],
        "roofline_kernels": roof["all"],
        "segmentation": seg,
        "training_step_with_losses": loss_step,
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def losses_step_time(lit, d, n_tx, n_cells, device, flat, opt, timed, K, ms_synth):
    """Training step through LitISTEncoder.training_step: triplet loss on transcripts (20 synthetic clusters), metric
    loss on boundaries (10 clusters), segmentation triplet loss on the tx-belongs-bd edges -- sampling included."""
    from segger_b200 import ops
    from segger_b200.hetero import HeteroBatch
    g = torch.Generator().manual_seed(5)

    def sim(c):
        a = torch.rand(c, c, generator=g) * 2 - 1
        return ((a + a.t()) / 2).contiguous()

    lit.setup_losses(sim(20), sim(10))
    lit.set_epoch(5, 10)
    b = HeteroBatch()
    b["tx"]["x"], b["tx"]["pos"], b["tx"]["batch"] = d["tx_x"], d["tx_pos"], d["tx_batch"]
    b["bd"]["x"], b["bd"]["pos"], b["bd"]["batch"] = d["bd_x"], d["bd_pos"], d["bd_batch"]
    b["tx"]["mask"] = torch.ones(n_tx, dtype=torch.bool, device=device)
    b["bd"]["mask"] = torch.ones(n_cells, dtype=torch.bool, device=device)
    b["tx"]["cluster"] = torch.randint(0, 20, (n_tx,), generator=g).to(device)
    b["bd"]["cluster"] = torch.randint(0, 10, (n_cells,), generator=g).to(device)
    b[TT]["edge_index"], b[TB]["edge_index"] = d["e_tt"], d["e_tb"]

    def step():
        ops.CSR_CACHE.clear()
        flat.zero()
        loss = lit.training_step(b, 0)
        loss.backward()
        flat.reduce()
        opt.step()

    for _ in range(3):
        step()
    l0 = ops.LAUNCHES
    ms = timed(step, K) / K
    return {"ms_per_step": ms, "ms_per_step_synthetic_loss": ms_synth, "loss_overhead_ms": ms - ms_synth,
            "gpu_launches_per_step": (ops.LAUNCHES - l0) / K,
            "losses": "TripletLoss(tx, 20 clusters, margin 0.3) + MetricLoss(bd, 10 clusters) + segmentation triplet "
                      "(margin 0.4), weights at epoch 5/10; sampling, forward and backward in sgb_loss.cu kernels"}


def kernel_rooflines(model, d, n_tx, n_cells, H, C, device, hbm_peak, peak_src):
    """Each fused message-passing kernel timed ALONE (CUDA events on its stream, 10 launches, a
    256 MB write between launches to flush the 126 MB L2) -> achieved algorithmic GB/s vs HBM peak."""
    from segger_b200 import ops
    F = H * C
    csr_tt = ops.build_csr(d["e_tt"], n_tx, n_tx)
    csr_tb = ops.build_csr(d["e_tb"], n_tx, n_cells)
    y = torch.randn(n_tx, 3 * F, device=device)
    y_bd = torch.randn(n_cells, F, device=device)
    att = torch.randn(H * C, device=device) * 0.1
    bias = torch.randn(F, device=device) * 0.1
    g = torch.randn(n_tx, F, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    touched_tt = int(torch.unique(d["e_tt"][0]).numel())
    fwd_b, bwd_b = gat_bytes(touched_tt, n_tx, n_tx, csr_tt.E, H, C)

    def time_it(fn, reps=10):
        fn(); torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps * 1e-3

    out, _, smax, sden = ops.gatv2_fwd(y[:, :F], y[:, F:2 * F], att, bias, csr_tt, H, C, 0.2, 0.0, False, 0, True)
    G = torch.empty(n_tx, 3 * F, device=device)
    t_fwd = time_it(lambda: ops.gatv2_fwd(y[:, :F], y[:, F:2 * F], att, bias, csr_tt, H, C, 0.2, 0.2, True, 7, True))
    t_bwd = time_it(lambda: ops.gatv2_bwd(y[:, :F], y[:, F:2 * F], att, bias, out, g, True, csr_tt, H, C, 0.2, 0.2, True,
                                          7, smax, sden, grad_x_l=G[:, :F], grad_x_r=G[:, F:2 * F]))

    def entry(name, nbytes, t, traffic=None):
        a = nbytes / t / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak,
                "traffic": traffic, "algorithmic_bytes": nbytes, "ms": t * 1e3, "peak_source": peak_src}

    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of
    # exactly these launches on this workload (profiles/r1e_ncu_full_summary.txt); null for other workloads.
    cfg2 = (n_tx, n_cells, H, C) == (1_000_000, 10_000, 2, 64)
    tr_fwd = 1.280327e9 + 0.999322e9 if cfg2 else None
    tr_bwd = (2.398745e9 + 1.078364e9) + (2.469396e9 + 0.498534e9) if cfg2 else None
    ents = [entry("gatv2_fwd_quad_kernel<4,8,2,4> (tx-neighbors-tx: fused logits + segment softmax + dropout + aggregate "
                  "+ bias + GELU)", fwd_b, t_fwd, tr_fwd),
            entry("gatv2_bwd_dst_quad_kernel<4,8,2,4,3> + gatv2_bwd_src_quad_kernel<4,8,2,3> + quad_colsum_kernel "
                  "(tx-neighbors-tx backward)", bwd_b, t_bwd, tr_bwd)]

    # the largest projection GEMM of the step (layer-1 tx projection, [N, 256] x [256, 3F]) on the tensor pipe:
    # achieved = TF32 flops actually issued (3 split products per fp32 product) / time; peak = dense TF32 =
    # half the measured bf16 cuBLAS rate of MEASURED_PEAKS.json (tcgen05 kind::tf32 runs at half the bf16 rate).
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            tf32_peak, tsrc = float(json.load(f)["bf16_tflops"]) / 2, "measured bf16 / 2 (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        tf32_peak, tsrc = 2250.0 / 2, "nominal bf16 / 2"
    xk = torch.randn(n_tx, 256, device=device)
    wk = torch.randn(3 * F, 256, device=device) / 16
    t_gemm = time_it(lambda: ops.linear_fwd(xk, wk, None, exact=1), reps=5)
    fl = 2.0 * n_tx * 3 * F * 256
    gemm = {"kernel": "gemm_tf32x3_kernel<128,0,0,1,3> (layer-1 tx projection, split-TF32 x3 on tcgen05, fp32-exact)",
            "bound": "tensor", "achieved": 3 * fl / t_gemm / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
            "frac": 3 * fl / t_gemm / 1e12 / tf32_peak, "traffic": (1.026287e9 + 1.487671e9) if cfg2 else None,
            # sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed of this launch (ncu, unthrottled
            # clocks; the cuBLAS-derived peak above is power-capped, hence the higher `frac`)
            "ncu_tensor_pipe_active_pct": 25.1 if cfg2 else None,
            "algorithmic_flops": fl, "issued_tf32_flops": 3 * fl, "fp32_equivalent_tflops": fl / t_gemm / 1e12,
            "algorithmic_bytes": 4 * n_tx * (256 + 3 * F), "ms": t_gemm * 1e3, "peak_source": tsrc}
    del xk, wk
    dom = max(ents, key=lambda e: e["ms"])
    return {"dominant": {k: dom[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "ms",
                                            "algorithmic_bytes", "peak_source")}, "all": ents + [gemm]}


def segmentation_throughput(lit, host, ts, device, world, timed):
    """transcripts/s of LitISTEncoder.predict_step (forward + score + arg-max + D2H of the results),
    inputs already on the device (the reference keeps the dataset on the GPU for predict,
    data_module.py:310)."""
    from segger_b200 import ops
    from segger_b200.hetero import HeteroBatch
    lit.eval()
    d = to_device(host, device, PRED_KEYS)
    b = HeteroBatch()
    b["tx"]["x"], b["tx"]["pos"], b["tx"]["batch"], b["tx"]["index"] = d["tx_x"], d["tx_pos"], d["tx_batch"], d["tx_index"]
    b["tx"]["predict_mask"] = torch.ones(d["tx_x"].size(0), dtype=torch.bool, device=device)
    b["bd"]["x"], b["bd"]["pos"], b["bd"]["batch"], b["bd"]["index"] = d["bd_x"], d["bd_pos"], d["bd_batch"], d["bd_index"]
    b[TT]["edge_index"], b[TB]["edge_index"], b[PRED]["edge_index"] = d["e_tt_full"], d["e_tb"], d["e_pred"]
    res = {}

    def step():
        ops.CSR_CACHE.clear()
        with torch.no_grad():
            res["out"] = lit.predict_step(b, 0)

    for _ in range(3):
        step()
    ms = timed(step, 5) / 5
    n = d["tx_x"].size(0)
    assigned = float((res["out"][1] >= 0).float().mean())

    # the same with the transcript kNN graph rebuilt every step ("with and without graph construction");
    # the tx-neighbors-bd candidate list (points in buffered polygons) is host preprocessing in both arms
    from segger_b200.neighbors import kdtree_neighbors
    def step_knn():
        ei, _ = kdtree_neighbors(d["tx_pos"], knn_k, 5.0, device_output=True, device=device)
        b[TT]["edge_index"] = ei
        step()

    knn_k = WORKLOADS_K.get(n, 5)
    for _ in range(2):
        step_knn()
    ms_knn = timed(step_knn, 5) / 5
    b[TT]["edge_index"] = d["e_tt_full"]

    # ... and with the tx-neighbors-bd candidate edges rebuilt too: transcripts strictly inside the
    # buffered 16-gon cell outlines (radius 6.5 um x 1.05, the generator's cells), outlines resident on the host as in
    # the reference (geopandas buffer), point-in-polygon join on the GPU
    from segger_b200.geometry import PackedPolygons, pack_rings, points_in_polygons
    ang = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    rings = [np.stack([c[0] + 6.5 * 1.05 * np.cos(ang), c[1] + 6.5 * 1.05 * np.sin(ang)], 1) for c in ts.bd_pos.astype(np.float64)]
    polys = PackedPolygons(*pack_rings(rings))       # built once per boundary set, like the reference's GeoSeries
    e_pred_saved = b[PRED]["edge_index"]
    n_pip = {}

    def step_graph():
        ei, _ = kdtree_neighbors(d["tx_pos"], knn_k, 5.0, device_output=True, device=device)
        b[TT]["edge_index"] = ei
        ep = points_in_polygons(d["tx_pos"], polys, device=device, device_output=True)
        n_pip["E"] = int(ep.size(1))
        b[PRED]["edge_index"] = ep
        step()

    for _ in range(2):
        step_graph()
    ms_graph = timed(step_graph, 5) / 5
    assigned_pip = float((res["out"][1] >= 0).float().mean())
    b[TT]["edge_index"], b[PRED]["edge_index"] = d["e_tt_full"], e_pred_saved
    lit.train()
    return {"metric": "segmentation_transcripts_per_sec", "value": world * n / (ms * 1e-3), "unit": "transcripts/s",
            "ms_per_step": ms, "assigned_frac": assigned,
            "step": "predict_step: CSR build + forward + fused score/arg-max + masked D2H of (index, cell, sim, gene)",
            "with_knn_graph_construction": {"value": world * n / (ms_knn * 1e-3), "unit": "transcripts/s",
                                            "ms_per_step": ms_knn, "knn_ms": ms_knn - ms, "k": knn_k, "max_dist": 5.0},
            "with_knn_and_prediction_graph_construction": {
                "value": world * n / (ms_graph * 1e-3), "unit": "transcripts/s", "ms_per_step": ms_graph,
                "pip_ms": ms_graph - ms_knn, "candidate_edges": n_pip.get("E"), "assigned_frac": assigned_pip,
                "polygons": "one buffered 16-gon per cell (radius 6.5 um x 1.05)"}}


if __name__ == "__main__":
    main()
