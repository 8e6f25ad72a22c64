#!/usr/bin/env python
"""Benchmark of the segger hot path on B200 (contract: the task statement / DESIGN.md section 6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch_cuda] [--workload cfg2|cfg1|cfg4]

A *step* = one training pass of the hot path over one synthetic tile batch: CSR build, ISTEncoder forward (input
stage, hetero GATv2 layers, output projection, normalise), a linear synthetic loss, the fused deterministic backward,
the data-parallel gradient all-reduce (N > 1) and an Adam step.
Workload (N=1): BASELINE.json configs[1] -- 1M transcripts / 10k cells, k=5, 2-layer hetero GATv2, hidden=64, heads=2.

One JSON line is printed by rank 0:
  value / ms_per_step   GATv2 edge-layers/s with inputs resident in HBM; under torchrun every rank owns its own 1M tile
                        set ("scaling": "weak")
  e2e                   the same with pinned-host inputs copied H2D every step and the loss read back
  strong_scaling        ONE 1M tile set split tile-wise over the ranks (segger_b200.distributed.assign_tiles), each rank
                        collates and trains on its own tiles, gradients all-reduced
  roofline / roofline_kernels / mp_only   the fused message-passing kernels timed alone against the measured HBM peak
  segmentation          predict_step transcripts/s (resident and host->device->host), its roofline and CPU baseline
  inference_cfg3        BASELINE configs[2]: `segger segment`-style inference over 20M transcripts: tiles + 20 um halo
                        packed onto the ranks, per-rank predict_step, device all-gather + de-duplication, all timed
  cpu_baseline          the CPU oracle on the box's host cores (N=1 only)
--impl reference times the CPU restatement of the reference path; --impl torch_cuda (informational) runs the same
restatement's ATen ops on the GPU (what PyG dispatches to), the same-box comparator for the hand-written pipeline.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import tempfile
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
WORKLOADS_DESC = {
    "cfg1": "BASELINE configs[0]: 50k transcripts / 500 cells single tile, kNN k=5, 2-layer hetero GATv2 hidden=64 heads=2, training step",
    "cfg2": "BASELINE configs[1]: 1M transcripts / 10k cells tile set (25 tiles), kNN k=5, 2-layer hetero GATv2 hidden=64 heads=2, training step",
    "cfg4": "BASELINE configs[3]: 1M transcripts / 10k cells, kNN k=20, 3-layer hetero GATv2 hidden=128 heads=4, training step",
}
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
    """(hbm GB/s, bf16 TF/s burst, source) from the driver-written MEASURED_PEAKS.json, else the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel_key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel_key` from the committed `ncu --set full`
    capture (profiles/ncu_traffic.json, written by scripts/ncu_traffic.py); None when the capture predates the current
    kernel source (sha1 of the .cu file differs) or the kernel was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            db = json.load(f)
        ent = db["kernels"][kernel_key]
        with open(os.path.join(ROOT, ent["source"]), "rb") as f:
            if hashlib.sha1(f.read()).hexdigest() != ent["source_sha1"]:
                return None
        return ent
    except Exception:  # noqa: BLE001
        return None


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

    def wait_first_sample(self, timeout: float = 15.0):
        """Block until nvidia-smi has printed its first line: its start-up (NVML initialisation, ~1 s on a fresh box)
        holds driver locks that stall kernel launches, which must not overlap the timed steps."""
        t0 = time.time()
        while self.proc is not None and time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                pass
            time.sleep(0.05)

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
def oracle_step_factory(workload: str, seed: int = 0, device="cpu"):
    """One bounded sample of the workload: a single <=50k-transcript tile (the reference's own tile size,
    tiling_nodes_per_tile=50_000) with the workload's model, forward+backward+Adam, on `device`."""
    from oracle import neighbors_ref
    from oracle.ist_encoder_ref import ISTEncoderRef
    from segger_b200.synth import synth
    _, _, k, in_c, hid, out_c, n_mid, heads = WORKLOADS[workload]
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    ts = synth(50_000, 500, seed=seed)
    ei, _ = neighbors_ref.kdtree_neighbors(ts.tx_pos, k, 5.0)
    mv = lambda t: t.to(device)
    x = {"tx": mv(torch.from_numpy(ts.tx_gene)), "bd": mv(torch.from_numpy(ts.bd_x))}
    pos = {"tx": mv(torch.from_numpy(ts.tx_pos)), "bd": mv(torch.from_numpy(ts.bd_pos))}
    bat = {"tx": mv(torch.from_numpy(ts.tx_tile)), "bd": mv(torch.from_numpy(ts.bd_tile))}
    edges = {TT: mv(ei), TB: mv(torch.from_numpy(ts.edge_tb))}
    torch.manual_seed(0)
    model = ISTEncoderRef(ts.n_genes, ts.bd_x.shape[1], in_c, hid, out_c, n_mid, heads).to(device).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(1)
    t_tx, t_bd = mv(torch.randn(50_000, out_c, generator=g)), mv(torch.randn(500, out_c, generator=g))
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
              f"hidden={hid}, heads={heads}), fwd+bwd+Adam, plain-torch restatement of the PyG path "
              f"(oracle/), {edge_layers} edge-layers per step")
    return step, edge_layers, sample


def base_config(workload):
    return {"workload": WORKLOADS_DESC[workload]}


def run_reference(args):
    """--impl reference: the reference's CPU path (its restatement, oracle/ -- the reference itself cannot be imported
    on the GPU box) on all host cores; every step is one 50k-transcript tile of the workload (bounded sample)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    step, edge_layers, sample = oracle_step_factory(args.workload)
    W, K = max(1, args.warmup), max(1, args.steps)
    for _ in range(W):
        step()
    t0 = time.perf_counter()
    for _ in range(K):
        step()
    dt = (time.perf_counter() - t0) / K
    v = edge_layers / dt
    cores = torch.get_num_threads()
    cfg = base_config(args.workload)
    cfg.update(timing="host wall clock, CPU only", sample=sample)
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def run_torch_cuda(args):
    """--impl torch_cuda (informational): the oracle's ATen ops (index_select / scatter_reduce / index_add / mm -- what
    PyG dispatches to) on the same GPU, cuBLAS fp32, on one 50k-transcript tile like the CPU arm."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    assert torch.cuda.is_available()
    torch.backends.cuda.matmul.allow_tf32 = False
    step, edge_layers, sample = oracle_step_factory(args.workload, device="cuda")
    W, K = max(3, args.warmup), max(1, args.steps)
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) / K * 1e-3
    cfg = base_config(args.workload)
    cfg.update(timing="CUDA events", sample=sample.replace("plain-torch", "plain-torch on cuda:0 (eager ATen + cuBLAS fp32)"))
    emit({"impl": "torch_cuda", "metric": METRIC, "value": edge_layers / dt, "unit": UNIT, "n_gpus": 1, "steps": K,
          "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f32", "data": "synthetic", "config": cfg, "gpu_launches": 0})


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_cuda"])
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-step", action="store_true",
                    help="profiling aid: the headline step legs only (the launch list under ncu), no secondary legs")
    ap.add_argument("--infer-tx", type=int, default=20_000_000, help="transcripts of the configs[2] inference leg (0: skip)")
    ap.add_argument("--max-edges-per-batch", type=int, default=8_000_000,
                    help="edges per predict batch of the inference leg (the reference's default is 1M: also reported)")
    args = ap.parse_args()
    capture_stdout()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "torch_cuda":
        return run_torch_cuda(args)

    import torch.distributed as dist
    from segger_b200 import ops
    from segger_b200.distributed import FlatGradAllReduce, trainable_parameters
    from segger_b200.lightning_model import LitISTEncoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback; use --impl reference)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    n_tx, n_cells, k, in_c, hid, out_c, n_mid, heads = WORKLOADS[args.workload]
    n_layers = n_mid + 2

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
    # the clock sampler (an nvidia-smi child process) is started BEFORE the warm-up and has printed its first sample
    # before the warm-up begins: its start-up stalls kernel launches for up to a second on a fresh box, which must not
    # land inside the timed steps; it keeps sampling (every 100 ms) through them
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    for _ in range(W):
        train_step(dev_in)
    l0 = ops.LAUNCHES
    ms_total = timed(lambda: train_step(dev_in), K)
    launches = (ops.LAUNCHES - l0)
    clocks = sampler.stop() if rank == 0 else {}
    ms_step = ms_total / K
    value = world * edge_layers / (ms_step * 1e-3)

    # ---- end to end: pinned host inputs -> H2D every step, loss -> D2H every step --------------
    h2d = sum(host[k_].numel() * host[k_].element_size() for k_ in TRAIN_KEYS)
    copy_stream = torch.cuda.Stream(device)
    compute_stream = torch.cuda.current_stream(device)
    # two preallocated device input sets (double buffering, what a pinned-memory DataLoader + prefetch does)
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

    def guarded(fn, *a, **kw):
        """Secondary measurements never take the headline line down with them."""
        try:
            return fn(*a, **kw)
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            return {"error": f"{type(e).__name__}: {e}"}

    if args.only_step:
        if rank == 0:
            emit({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                  "ms_per_step": ms_step, "e2e_ms_per_step": ms_e2e, "gpu_launches": launches, "only_step": True})
        if world > 1:
            dist.destroy_process_group()
        return
    loss_step = guarded(losses_step_time, lit, dev_in, n_tx, n_cells, device, flat, opt, timed, K, ms_step)
    hbm_peak, bf16_peak, peak_src = peaks()
    roof = guarded(kernel_rooflines, dev_in, n_tx, n_cells, heads, hid, n_layers, device, hbm_peak, bf16_peak, peak_src,
                   args.workload)
    strong = guarded(strong_scaling_leg, args.workload, lit, flat, opt, device, world, rank, timed, K, W, t_tx.size(1))
    seg = guarded(segmentation_throughput, lit, host, ts, device, world, rank, timed, hbm_peak, peak_src, k,
                  not args.no_cpu_baseline)
    small = guarded(small_tile_graph_leg, device, timed, K) if world == 1 else None
    infer = None
    if args.infer_tx > 0 and args.workload == "cfg2":
        infer = guarded(inference_cfg3_leg, lit, args.infer_tx, args.max_edges_per_batch, device, world, rank, timed, hbm_peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        step, el, sample = oracle_step_factory(args.workload)
        step()
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter(); step(); best = min(best, time.perf_counter() - t0)
        cpu = {"value": el / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": sample + "; best of 3 after 1 warm-up"}

    cfg = base_config(args.workload)
    cfg.update({
        "per_gpu": {"n_tx": n_tx, "n_cells": n_cells, "E_tt": E_tt, "E_tb": E_tb, "layers": n_layers,
                    "edge_layers_per_step": edge_layers, "tiles": ts.n_tiles},
        "step": "CSR build + ISTEncoder fwd + linear synthetic loss + fused bwd + grad all-reduce + fused Adam",
        "l2": "inputs larger than L2 (per-layer activations 0.5-1.5 GB vs 126 MB L2); kernel-alone timings flush L2",
        "parallelism": f"dp{world} (one tile set per rank; flat gradient of {flat.nbytes} B all-reduced in "
                       f"{len(flat.buckets)} buckets on a side stream, overlapped with the backward)",
        "timing": "CUDA events on the launching stream, barrier+synchronize on both sides, max over ranks",
        "roofline": "`roofline` = the slowest fused message-passing launch group (HBM-bound; the kernels BASELINE.json's "
                    "metric names); `mp_only` = all message-passing launches of a step; `roofline_kernels` adds the "
                    "largest projection GEMM against the tensor pipe",
    })
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        "roofline": roof.get("dominant") if isinstance(roof, dict) else None,
        "roofline_kernels": roof.get("all") if isinstance(roof, dict) else roof,
        "mp_only": roof.get("mp_only") if isinstance(roof, dict) else None,
        "strong_scaling": strong,
        "segmentation": seg,
        "inference_cfg3": infer,
        "small_tile_cfg1": small,
        "training_step_with_losses": loss_step,
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def small_tile_graph_leg(device, timed, K):
    """BASELINE configs[0] (one 50k-transcript / 500-cell tile, the reference's own tile size): the training step eager
    (launch-bound: ~300 launches of a few microseconds) and replayed from a CUDA graph (segger_b200.graphs).  The
    tile's CSRs are built once -- tiles keep their edges across epochs -- in BOTH variants, so the two differ only in
    how the launches reach the GPU."""
    from segger_b200 import graphs, ops
    from segger_b200.lightning_model import LitISTEncoder
    n_tx, n_cells, k, in_c, hid, out_c, n_mid, heads = WORKLOADS["cfg1"]
    ts, host = build_workload("cfg1", seed=0, device=device)
    torch.manual_seed(0)
    lit = LitISTEncoder(ts.n_genes, in_channels=in_c, hidden_channels=hid, out_channels=out_c, n_mid_layers=n_mid,
                        n_heads=heads).to(device)
    lit.train()
    model = lit.model
    d = to_device(host, device, TRAIN_KEYS)
    inputs = model_inputs(d)
    with torch.no_grad():
        model(*inputs)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3, fused=True, capturable=True)
    g = torch.Generator(device="cpu").manual_seed(1)
    t_tx = torch.randn(n_tx, out_c, generator=g).to(device)
    t_bd = torch.randn(n_cells, out_c, generator=g).to(device)
    edge_layers = (n_mid + 2) * (host["e_tt"].size(1) + host["e_tb"].size(1))

    def forward_loss():
        out = model(*inputs)
        return (out["tx"] * t_tx).sum() / n_tx + (out["bd"] * t_bd).sum() / n_cells

    def eager():
        opt.zero_grad(set_to_none=True)
        forward_loss().backward()
        opt.step()

    for _ in range(3):
        eager()
    l0 = ops.LAUNCHES
    reps = max(20, K)
    ms_eager = timed(eager, reps) / reps
    launches_eager = (ops.LAUNCHES - l0) / reps
    step = graphs.graphed_train_step(forward_loss, opt, warmup=3)
    for _ in range(3):
        step.replay()
    ms_graph = timed(step.replay, reps) / reps
    loss = float(step.out.item())
    return {"workload": WORKLOADS_DESC["cfg1"], "edge_layers_per_step": edge_layers,
            "eager": {"ms_per_step": ms_eager, "value": edge_layers / (ms_eager * 1e-3), "gpu_launches_per_step": launches_eager},
            "cuda_graph": {"ms_per_step": ms_graph, "value": edge_layers / (ms_graph * 1e-3),
                           "gpu_launches_in_graph": step.launches, "loss_finite": bool(np.isfinite(loss))},
            "unit": UNIT, "steps": reps,
            "what": "forward + synthetic linear loss + backward + fused capturable Adam on one resident tile; CSRs of the "
                    "tile built once (static across epochs) in both variants; dropout seeds from a device word in the "
                    "graph (fresh mask every replay)"}


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


def _time_alone(fn, flush, reps=10):
    """One launch group timed ALONE with CUDA events on its stream, a 256 MB write before every launch (flushes the
    126 MB L2)."""
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e-3


def kernel_rooflines(d, n_tx, n_cells, H, C, n_layers, device, hbm_peak, bf16_peak, peak_src, workload):
    """The fused message-passing kernels (both edge types, forward and backward) and the largest projection GEMM, each
    timed alone -> achieved algorithmic GB/s against the measured HBM peak / useful TFLOP/s against the tensor peak."""
    from segger_b200 import _lib, ops
    F = H * C
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    csr_tt = ops.build_csr(d["e_tt"], n_tx, n_tx)
    csr_tb = ops.build_csr(d["e_tb"], n_tx, n_cells)
    quad = bool(_lib.load().sgb_gatv2_quad_supported(H, C))
    att = torch.randn(F, device=device) * 0.1
    bias = torch.randn(F, device=device) * 0.1
    y = torch.randn(n_tx, 2 * F, device=device)
    touched_tt = int(torch.unique(d["e_tt"][0]).numel())
    touched_tb = int(torch.unique(d["e_tb"][0]).numel())

    def entry(name, nbytes, t, key=None):
        a = nbytes / t / 1e9
        tr = measured_traffic(key) if key else None
        return {"kernel": name, "bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak,
                "traffic": (tr["dram_read"] + tr["dram_write"]) if tr else None,
                "traffic_source": tr["captured"] if tr else None,
                "algorithmic_bytes": nbytes, "ms": t * 1e3, "peak_source": peak_src}

    # tx-neighbors-tx conv
    fwd_b, bwd_b = gat_bytes(touched_tt, n_tx, n_tx, csr_tt.E, H, C)
    g = torch.randn(n_tx, F, device=device)
    # the launches a training step makes: the forward also leaves the raw logits [E, H] the backward's dst pass reads back
    # (want_logits: 4H bytes per edge each way, NOT counted in the algorithmic bytes below)
    out, _, smax, sden, lg = ops.gatv2_fwd(y[:, :F], y[:, F:], att, bias, csr_tt, H, C, 0.2, 0.2, True, 7, True, want_logits=True)
    G = torch.empty(n_tx, 2 * F, device=device)
    t_fwd = _time_alone(lambda: ops.gatv2_fwd(y[:, :F], y[:, F:], att, bias, csr_tt, H, C, 0.2, 0.2, True, 7, True,
                                              want_logits=True), flush)
    t_bwd = _time_alone(lambda: ops.gatv2_bwd(y[:, :F], y[:, F:], att, bias, out, g, True, csr_tt, H, C, 0.2, 0.2, True,
                                              7, smax, sden, grad_x_l=G[:, :F], grad_x_r=G[:, F:], e_logit=lg), flush)
    path = "sub-warp (quad) kernels" if quad else "row-per-warp kernels"
    ents = [entry(f"gatv2 forward, tx-neighbors-tx, H={H} C={C} ({path}: fused logits + segment softmax + dropout + "
                  "aggregate + bias + GELU)", fwd_b, t_fwd, f"gatv2_fwd_tt_{workload}"),
            entry(f"gatv2 backward, tx-neighbors-tx, H={H} C={C} ({path}: saved-logit dst pass + src pass + column sums)", bwd_b, t_bwd,
                  f"gatv2_bwd_tt_{workload}")]
    # tx-belongs-bd conv on one virtual source per edge (what SkipGATLayerFn runs when the belongs list is unique)
    vcsr = csr_tb.per_edge_sources() if csr_tb.sources_unique_increasing() else csr_tb
    n_src_b = vcsr.n_src
    xl_b = torch.randn(n_src_b, F, device=device)
    xr_b = torch.randn(n_cells, F, device=device)
    gb = torch.randn(n_cells, F, device=device)
    fb, bb = gat_bytes(touched_tb, n_cells, n_src_b, vcsr.E, H, C)
    out_b, _, smax_b, sden_b, lg_b = ops.gatv2_fwd(xl_b, xr_b, att, bias, vcsr, H, C, 0.2, 0.2, True, 7, True, want_logits=True)
    t_fwd_b = _time_alone(lambda: ops.gatv2_fwd(xl_b, xr_b, att, bias, vcsr, H, C, 0.2, 0.2, True, 7, True, want_logits=True), flush)
    t_bwd_b = _time_alone(lambda: ops.gatv2_bwd(xl_b, xr_b, att, bias, out_b, gb, True, vcsr, H, C, 0.2, 0.2, True, 7,
                                                smax_b, sden_b, e_logit=lg_b), flush)
    ents += [entry("gatv2 forward, tx-belongs-bd", fb, t_fwd_b), entry("gatv2 backward, tx-belongs-bd", bb, t_bwd_b)]
    mp_t = n_layers * (t_fwd + t_bwd + t_fwd_b + t_bwd_b)
    mp_bytes = n_layers * (fwd_b + bwd_b + fb + bb)
    mp_el = n_layers * (csr_tt.E + csr_tb.E)
    mp_only = {"edge_layers_per_s": mp_el / mp_t, "ms_per_step": mp_t * 1e3, "algorithmic_bytes": mp_bytes,
               "achieved_gbs": mp_bytes / mp_t / 1e9, "frac": mp_bytes / mp_t / 1e9 / hbm_peak, "peak": hbm_peak,
               "what": f"every fused message-passing launch of one training step ({n_layers} layers x (tt, tb) x (fwd, bwd)), "
                       "each timed alone with L2 flushed"}

    # the largest projection GEMM of the step ([N, K] x [K, 2F], K = the first layer's dense width) on the tensor pipe
    Kdim = 128 if workload != "cfg4" else 128
    xk = torch.randn(n_tx, Kdim, device=device)
    wk = torch.randn(2 * F, Kdim, device=device) / 16
    t_gemm = _time_alone(lambda: ops.linear_fwd(xk, wk, None, exact=1), flush, reps=5)
    fl = 2.0 * n_tx * 2 * F * Kdim
    tf32_peak = bf16_peak / 2
    gbytes = 4 * n_tx * (Kdim + 2 * F)
    tr = measured_traffic(f"gemm_fwd_{workload}")
    gemm = {"kernel": f"split-TF32 tcgen05 GEMM, forward projection [{n_tx} x {Kdim}] x [{Kdim} x {2 * F}] (3 TF32 products per "
                      "fp32 product, fp32-exact)",
            "bound": "tensor", "achieved": fl / t_gemm / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
            "frac": fl / t_gemm / 1e12 / tf32_peak,
            "frac_note": "USEFUL fp32-equivalent flops / (measured bf16 cuBLAS peak / 2 = dense TF32 rate); the kernel issues 3x as many TF32 flops",
            "issued_tf32_tflops": 3 * fl / t_gemm / 1e12, "hbm_gbs": gbytes / t_gemm / 1e9, "hbm_frac": gbytes / t_gemm / 1e9 / hbm_peak,
            "traffic": (tr["dram_read"] + tr["dram_write"]) if tr else None,
            "ncu_tensor_pipe_active_pct": tr.get("tensor_pipe_pct") if tr else None,
            "algorithmic_flops": fl, "algorithmic_bytes": gbytes, "ms": t_gemm * 1e3, "peak_source": peak_src}
    dom = max(ents[:2], key=lambda e: e["ms"])
    return {"dominant": {k: dom[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "ms",
                                            "algorithmic_bytes", "peak_source")}, "all": ents + [gemm], "mp_only": mp_only}


def strong_scaling_leg(workload, lit, flat, opt, device, world, rank, timed, K, W, out_c):
    """ONE tile set (seed 0 on every rank) partitioned on the device by its tiles (TilePartition = PartitionDataset),
    the tiles dealt to the ranks by best-fit-decreasing on their edge counts; a step = collate the rank's tiles into a
    batch + CSR build + fwd + bwd + gradient all-reduce + Adam.  value = edge-layers of the WHOLE set / max-over-ranks
    time: the fixed-total-work counterpart of the headline number."""
    from segger_b200 import ops
    from segger_b200.distributed import assign_tiles
    from segger_b200.hetero import HeteroBatch
    from segger_b200.neighbors import kdtree_neighbors
    from segger_b200.synth import synth
    from segger_b200.tiles import TilePartition
    n_tx, n_cells, k, in_c, hid, _, n_mid, heads = WORKLOADS[workload]
    ts = synth(n_tx, n_cells, seed=0)
    ei, _ = kdtree_neighbors(ts.tx_pos, k, 5.0, device_output=True, device=device)
    b = HeteroBatch()
    b["tx"]["x"], b["tx"]["pos"] = torch.from_numpy(ts.tx_gene).to(device), torch.from_numpy(ts.tx_pos).to(device)
    b["bd"]["x"], b["bd"]["pos"] = torch.from_numpy(ts.bd_x).to(device), torch.from_numpy(ts.bd_pos).to(device)
    b[TT]["edge_index"], b[TB]["edge_index"] = ei, torch.from_numpy(ts.edge_tb).to(device)
    part = TilePartition(b, {"tx": torch.from_numpy(ts.tx_tile).to(device), "bd": torch.from_numpy(ts.bd_tile).to(device)},
                         ts.n_tiles)
    w = part.weights("edge")
    mine = assign_tiles(w, world)[rank]
    total_edges = sum(w)
    model = lit.model.train()
    g = torch.Generator(device="cpu").manual_seed(2)
    n_mine_tx = sum(part.node_sizes["tx"][t] for t in mine)
    n_mine_bd = sum(part.node_sizes["bd"][t] for t in mine)
    t_tx = torch.randn(max(n_mine_tx, 1), out_c, generator=g).to(device)
    t_bd = torch.randn(max(n_mine_bd, 1), out_c, generator=g).to(device)

    def step():
        ops.CSR_CACHE.clear()
        flat.zero()
        if mine:
            batch = part.collate(mine)
            out = model(batch.x_dict, {TT: batch[TT]["edge_index"], TB: batch[TB]["edge_index"]}, batch.pos_dict, batch.batch_dict)
            loss = (out["tx"] * t_tx).sum() / n_tx + (out["bd"] * t_bd).sum() / n_cells
            loss.backward()
        flat.reduce()
        opt.step()

    for _ in range(W):
        step()
    ms = timed(step, K) / K
    el = (n_mid + 2) * total_edges
    return {"value": el / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "scaling": "strong", "n_gpus": world,
            "tiles": ts.n_tiles, "tiles_this_rank": len(mine), "edges_total": total_edges,
            "rank0_edge_share": sum(w[t] for t in mine) / max(total_edges, 1),
            "what": "one 1M-transcript tile set split tile-wise over the ranks (assign_tiles), device-side collate inside the "
                    "timed step, same model and optimiser as the headline",
            "limiter": "per-rank work shrinks with N while ~0.3k launches per step and the tile imbalance (25 tiles over N "
                       "ranks) stay: at N=8 a rank holds 3-4 tiles (~125k transcripts, ~3 ms of kernels)"}


def segmentation_throughput(lit, host, ts, device, world, rank, timed, hbm_peak, peak_src, knn_k, want_cpu):
    """transcripts/s of LitISTEncoder.predict_step (forward + score + arg-max + compaction + D2H of the results):
    `value` with the batch resident on the device (the reference keeps the dataset on the GPU for predict,
    data_module.py:310), `e2e` with the batch copied host->device inside every step."""
    from segger_b200 import ops
    from segger_b200.hetero import HeteroBatch
    lit.eval()
    d = to_device(host, device, PRED_KEYS)

    def make_batch(d):
        b = HeteroBatch()
        b["tx"]["x"], b["tx"]["pos"], b["tx"]["batch"], b["tx"]["index"] = d["tx_x"], d["tx_pos"], d["tx_batch"], d["tx_index"]
        b["tx"]["predict_mask"] = torch.ones(d["tx_x"].size(0), dtype=torch.bool, device=device)
        b["bd"]["x"], b["bd"]["pos"], b["bd"]["batch"], b["bd"]["index"] = d["bd_x"], d["bd_pos"], d["bd_batch"], d["bd_index"]
        b[TT]["edge_index"], b[TB]["edge_index"], b[PRED]["edge_index"] = d["e_tt_full"], d["e_tb"], d["e_pred"]
        return b

    b = make_batch(d)
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
    d2h = sum(t.numel() * t.element_size() for t in res["out"])

    # end to end: host (pinned) -> device copy of the whole batch inside every step
    h2d = sum(host[k_].numel() * host[k_].element_size() for k_ in PRED_KEYS)
    # two device input sets, the next batch copied on a copy stream while the current one is processed (what a
    # pinned-memory DataLoader with prefetch does; same scheme as the training leg): every timed step issues the copy
    # of one full batch and consumes one
    copy_stream = torch.cuda.Stream(device)
    compute_stream = torch.cuda.current_stream(device)
    dbufs = [{k_: torch.empty_like(host[k_], device=device) for k_ in PRED_KEYS} for _ in range(2)]
    used = [None, None]

    def fetch(slot):
        with torch.cuda.stream(copy_stream):
            if used[slot] is not None:
                copy_stream.wait_event(used[slot])       # the step that last read this buffer set has finished
            for k_ in PRED_KEYS:
                dbufs[slot][k_].copy_(host[k_], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    state = {"slot": 0, "ready": fetch(0)}

    def step_e2e():
        slot = state["slot"]
        compute_stream.wait_event(state["ready"])
        state["ready"] = fetch(slot ^ 1)                 # next batch: H2D overlaps this batch's predict_step
        bb = make_batch(dbufs[slot])
        ops.CSR_CACHE.clear()
        with torch.no_grad():
            res["out"] = lit.predict_step(bb, 0)
        used[slot] = torch.cuda.Event()
        used[slot].record(compute_stream)
        state["slot"] = slot ^ 1

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, 5) / 5
    copy_stream.synchronize()

    # with the transcript kNN graph (and the tx-neighbors-bd candidates: points in buffered polygons) rebuilt
    from segger_b200.geometry import PackedPolygons, pack_rings, points_in_polygons
    from segger_b200.neighbors import kdtree_neighbors
    ang = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    rings = [np.stack([c[0] + 6.5 * 1.05 * np.cos(ang), c[1] + 6.5 * 1.05 * np.sin(ang)], 1) for c in ts.bd_pos.astype(np.float64)]
    polys = PackedPolygons(*pack_rings(rings))
    saved = (b[TT]["edge_index"], b[PRED]["edge_index"])
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
    b[TT]["edge_index"], b[PRED]["edge_index"] = saved

    # roofline of the scoring kernel (the part BASELINE's metric names besides the message passing), timed alone
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    D = lit.model.hparams["out_channels"]
    e_tx = torch.nn.functional.normalize(torch.randn(n, D, device=device), dim=-1)
    e_bd = torch.nn.functional.normalize(torch.randn(d["bd_x"].size(0), D, device=device), dim=-1)
    csr = ops.candidate_csr(d["e_pred"], n, e_bd.size(0))
    t_score = _time_alone(lambda: ops.score_argmax(e_tx, e_bd, d["e_pred"], d["bd_index"], csr=csr), flush)
    score_bytes = 4 * D * (n + e_bd.size(0)) + 8 * d["e_pred"].size(1) + 12 * n
    # end-to-end byte model of SURVEY 8d: MP fwd + score + projections + input stage + lin_last
    model_bytes = None
    if (n, e_bd.size(0)) == (1_000_000, 10_000):
        model_bytes = 10.9e9
    roof = {"bound": "hbm", "kernel": "score_argmax (cosine similarity over candidate edges + per-transcript arg-max + cell lookup)",
            "achieved": score_bytes / t_score / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": score_bytes / t_score / 1e9 / hbm_peak,
            "algorithmic_bytes": score_bytes, "ms": t_score * 1e3, "traffic": None, "peak_source": peak_src,
            "whole_step_vs_byte_model": (model_bytes / (ms * 1e-3) / 1e9 / hbm_peak) if model_bytes else None,
            "byte_model": "SURVEY 8d end-to-end inference model, 10.9 GB per 1M transcripts" if model_bytes else None}

    cpu = None
    if want_cpu and rank == 0 and world == 1:
        cpu = segmentation_cpu_baseline(knn_k)
    lit.train()
    return {"metric": "segmentation_transcripts_per_sec", "value": world * n / (ms * 1e-3), "unit": "transcripts/s",
            "ms_per_step": ms, "assigned_frac": assigned,
            "step": "predict_step: CSR build + forward + fused score/arg-max + device compaction + D2H of (index, cell, sim, gene)",
            "e2e": {"value": world * n / (ms_e2e * 1e-3), "unit": "transcripts/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "how": "pinned host batch -> device on a copy stream one batch ahead (double-buffered), results read back every step"},
            "roofline": roof, "cpu_baseline": cpu,
            "with_graph_construction": {"value": world * n / (ms_graph * 1e-3), "unit": "transcripts/s", "ms_per_step": ms_graph,
                                        "graph_ms": ms_graph - ms, "k": knn_k, "max_dist": 5.0, "candidate_edges": n_pip.get("E"),
                                        "what": "kNN graph (sgb_knn2d) and tx-neighbors-bd candidates (point-in-polygon join, one "
                                                "buffered 16-gon per cell) rebuilt every step"}}


def segmentation_cpu_baseline(knn_k):
    """The oracle's predict path (eval forward + cosine scoring + scatter-max) on one 50k-transcript tile, host cores."""
    from oracle import neighbors_ref
    from oracle.ist_encoder_ref import ISTEncoderRef, predict_scores_ref
    from segger_b200.synth import synth
    torch.set_num_threads(os.cpu_count() or 1)
    ts = synth(50_000, 500, seed=0)
    ei, _ = neighbors_ref.kdtree_neighbors(ts.tx_pos, knn_k, 5.0)
    x = {"tx": torch.from_numpy(ts.tx_gene), "bd": torch.from_numpy(ts.bd_x)}
    pos = {"tx": torch.from_numpy(ts.tx_pos), "bd": torch.from_numpy(ts.bd_pos)}
    bat = {"tx": torch.from_numpy(ts.tx_tile), "bd": torch.from_numpy(ts.bd_tile)}
    edges = {TT: ei, TB: torch.from_numpy(ts.edge_tb)}
    torch.manual_seed(0)
    model = ISTEncoderRef(ts.n_genes, ts.bd_x.shape[1], 128, 64, 64, 0, 2).eval()

    def step():
        with torch.no_grad():
            emb = model(x, edges, pos, bat)
            predict_scores_ref(emb["tx"], emb["bd"], torch.from_numpy(ts.edge_pred), torch.from_numpy(ts.bd_index))

    step()
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter(); step(); best = min(best, time.perf_counter() - t0)
    return {"value": 50_000 / best, "unit": "transcripts/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "one 50k-transcript / 500-cell tile, eval forward + cosine scoring + scatter-max (oracle/), best of 3"}


def inference_cfg3_leg(lit, n_tx_total, max_edges, device, world, rank, timed, hbm_peak):
    """BASELINE configs[2]: inference over `n_tx_total` synthetic transcripts tiled across the ranks.  Every rank holds the
    whole graph on its GPU (as the reference does, data_module.py:310), takes the tiles best-fit-decreasing assigns to it,
    cuts each tile + 20 um halo out on the device (TilePredictSet), batches tiles up to `max_edges` edges, runs
    predict_step with device-side results, and at the end all ranks all-gather fixed-width result tensors and the
    de-duplication (max similarity per transcript) runs on the device.  Everything from the first tile cut to the
    de-duplicated table is inside the timed region; value = transcripts / max-over-ranks time."""
    import torch.distributed as dist
    from segger_b200 import ops
    from segger_b200.distributed import assign_tiles, gather_rows_fixed_width
    from segger_b200.geometry import PackedPolygons, pack_rings, points_in_polygons
    from segger_b200.hetero import HeteroBatch
    from segger_b200.neighbors import kdtree_neighbors
    from segger_b200.synth import synth
    from segger_b200.tiles import TilePredictSet, slice_slots, square_tiles
    from segger_b200.writer import dedupe_predictions
    n_cells = n_tx_total // 100
    t0 = time.perf_counter()
    ts = synth(n_tx_total, n_cells, seed=0, pred_edges=False)
    t_synth = time.perf_counter() - t0
    pos = torch.from_numpy(ts.tx_pos).to(device)
    ei, _ = kdtree_neighbors(pos, 5, 5.0, device_output=True, device=device)
    ang = np.linspace(0, 2 * np.pi, 16, endpoint=False)
    c = ts.bd_pos.astype(np.float64)
    verts = np.stack([c[:, None, 0] + 6.5 * 1.05 * np.cos(ang)[None], c[:, None, 1] + 6.5 * 1.05 * np.sin(ang)[None]], -1).reshape(-1, 2)
    off = np.arange(n_cells + 1, dtype=np.int64) * 16
    ep = points_in_polygons(pos, PackedPolygons(verts, off), device=device, device_output=True)
    b = HeteroBatch()
    b["tx"]["x"], b["tx"]["pos"] = torch.from_numpy(ts.tx_gene).to(device), pos
    b["tx"]["index"] = torch.from_numpy(ts.tx_index).to(device)
    b["bd"]["x"], b["bd"]["pos"] = torch.from_numpy(ts.bd_x).to(device), torch.from_numpy(ts.bd_pos).to(device)
    b["bd"]["index"] = torch.from_numpy(ts.bd_index).to(device)
    b[TT]["edge_index"], b[TB]["edge_index"], b[PRED]["edge_index"] = ei, torch.from_numpy(ts.edge_tb).to(device), ep
    # square tiles of <= ~50k transcripts (the reference's tiling_nodes_per_tile), 20 um halo (prediction margin)
    nt = max(1, int(math.ceil(math.sqrt(n_tx_total / 50_000))))
    lo = ts.tx_pos.min(0) - 1e-3
    hi = ts.tx_pos.max(0) + 1e-3
    boxes = square_tiles(float(lo[0]), float(lo[1]), float(hi[0]), float(hi[1]), nt, nt)
    ix = np.clip(((ts.tx_pos[:, 0] - lo[0]) / ((hi[0] - lo[0]) / nt)).astype(np.int64), 0, nt - 1)
    iy = np.clip(((ts.tx_pos[:, 1] - lo[1]) / ((hi[1] - lo[1]) / nt)).astype(np.int64), 0, nt - 1)
    tile_tx = np.bincount(iy * nt + ix, minlength=nt * nt)
    mine = assign_tiles(tile_tx.tolist(), world)[rank]
    ds = TilePredictSet(b, boxes, margin=20.0, grid=(nt, nt))
    lit.eval()
    stats = {}

    # tiles are cut in GROUPS (TilePredictSet.cut: all tiles of a group in a handful of launches, collated on the device);
    # a group is sized to fill one batch of `max_edges` edges from the mean edges per tile incl. halo, then split by the
    # exact per-tile counts the cut returns, so no batch exceeds max_edges (a batch never spans two groups)
    E_all = sum(int(b[et]["edge_index"].size(1)) for et in (TT, TB, PRED))
    w_tile = float(hi[0] - lo[0]) / nt
    est_edges = E_all / (nt * nt) * ((w_tile + 40.0) / w_tile) ** 2

    def run():
        ops.CSR_CACHE.clear()
        shards, n_batches = [], 0
        G = max(1, min(64, int(max_edges // est_edges)))
        for g0 in range(0, len(mine), G):
            ids = mine[g0:g0 + G]
            group, info = ds.cut(ids)
            per_slot = [sum(info["edges"][et][s_] for et in (TT, TB, PRED)) for s_ in range(len(ids))]
            a, acc = 0, 0
            cuts = []
            for s_, e in enumerate(per_slot):
                if s_ > a and acc + e > max_edges:
                    cuts.append((a, s_)); a, acc = s_, 0
                acc += e
            cuts.append((a, len(ids)))
            for (a, b_) in cuts:
                batch = slice_slots(group, info, a, b_)
                with torch.no_grad():
                    shards.append(lit.predict_step(batch, 0, device_output=True))
                ops.CSR_CACHE.clear()
                n_batches += 1
        if shards:
            cols = [torch.cat([s[i] for s in shards]) for i in range(4)]
        else:
            cols = [torch.zeros(0, dtype=dt, device=device) for dt in (torch.int64, torch.int64, torch.float32, torch.int32)]
        if world > 1:
            cols = gather_rows_fixed_width(cols)
        row, seg, sim, gene = dedupe_predictions(*cols, max_row=n_tx_total)
        stats.update(rows=int(row.numel()), assigned=float((seg >= 0).float().mean()), batches=n_batches,
                     predicted_before_dedupe=int(cols[0].numel()))

    run()
    ms = timed(run, 1)
    ref_batching = None
    if max_edges != 1_000_000:          # the reference's own batch size (max_edges_per_batch = 1M, data_module.py:158)
        big, max_edges = max_edges, 1_000_000
        run()
        ms_ref = timed(run, 1)
        ref_batching = {"max_edges_per_batch": 1_000_000, "ms": ms_ref, "value": n_tx_total / (ms_ref * 1e-3),
                        "batches_this_rank": stats.get("batches")}
        max_edges = big
        run()
    # kernel-only byte model of SURVEY 8d: 4.48 GB per 1M transcripts (MP forward + score), 10.9 GB end to end
    return {"metric": "segmentation_transcripts_per_sec", "value": n_tx_total / (ms * 1e-3), "unit": "transcripts/s",
            "ms": ms, "n_tx": n_tx_total, "n_cells": n_cells, "n_gpus": world, "tiles": nt * nt, "tiles_this_rank": len(mine),
            "halo_um": 20.0, "max_edges_per_batch": max_edges, "batches_this_rank": stats.get("batches"),
            "rows_after_dedupe": stats.get("rows"), "rows_before_dedupe": stats.get("predicted_before_dedupe"),
            "assigned_frac": stats.get("assigned"), "complete": stats.get("rows") == n_tx_total,
            "frac_of_byte_model": (10.9e9 * n_tx_total / 1e6) / (ms * 1e-3) / 1e9 / (hbm_peak * world),
            "host_synth_s": t_synth, "with_reference_batch_size": ref_batching,
            "what": "grouped tile cut (+20 um halo; TilePredictSet.cut, collated on the device) + predict_step per batch + device all-gather of fixed-width result "
                    "tensors + device de-duplication, all timed; graph construction (kNN, point-in-polygon) is outside"}


if __name__ == "__main__":
    main()
