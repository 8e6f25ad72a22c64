""".ncu-rep of scripts/ncu_targets_r2.py -> profiles/ncu_traffic.json (the measured DRAM traffic bench.py reports as
`roofline.traffic`, keyed by kernel and guarded by the sha1 of the kernel's source file) + a per-launch text summary.

    python scripts/ncu_traffic.py gpurun_out/r2_top.ncu-rep r2x"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "r2")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    i = col.get(name)
    if i is None or i >= len(r) or r[i] == "":
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3,
             "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u)
    return v * scale if scale else v


launches = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    launches.append(dict(
        name=name, dur=val(r, "gpu__time_duration.sum"), rd=val(r, "dram__bytes_read.sum"), wr=val(r, "dram__bytes_write.sum"),
        dram_pct=val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        tensor_pct=val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
        or val(r, "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
        issue_pct=val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        occ_pct=val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), regs=val(r, "launch__registers_per_thread"),
        inst=val(r, "smsp__inst_executed.sum"), l2hit=val(r, "lts__t_sector_hit_rate.pct")))


def sha(path):
    with open(os.path.join(ROOT, path), "rb") as f:
        return hashlib.sha1(f.read()).hexdigest()


def first(pred, n=1, start=0):
    got = [l for l in launches[start:] if pred(l["name"])]
    return got[:n]


def entry(ls, source):
    return {"dram_read": sum(l["rd"] or 0 for l in ls), "dram_write": sum(l["wr"] or 0 for l in ls), "ms": sum(l["dur"] or 0 for l in ls) * 1e3,
            "launches": [l["name"][:80] for l in ls], "source": source, "source_sha1": sha(source), "captured": tag,
            "tensor_pipe_pct": max((l["tensor_pct"] or 0) for l in ls) if ls else None}


quad, tc = "segger_b200/csrc/sgb_gatv2_quad.cu", "segger_b200/csrc/sgb_linear_tc.cu"
gemms = [l for l in launches if "gemm_tf32x3" in l["name"]]
db = {"captured": tag, "command": "ncu --set full --clock-control none python scripts/ncu_targets_r2.py", "kernels": {
    "gatv2_fwd_tt_cfg2": entry(first(lambda n: "gatv2_fwd_quad" in n), quad),
    "gatv2_bwd_tt_cfg2": entry(first(lambda n: "gatv2_bwd_dst_quad" in n or "gatv2_bwd_dst_lg_quad" in n) + first(lambda n: "gatv2_bwd_src_quad" in n), quad),
    "gemm_fwd_cfg2": entry(gemms[:1], tc), "gemm_dgrad_cfg2": entry(gemms[1:2], tc), "gemm_wgrad_cfg2": entry(gemms[2:3], tc)}}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
    json.dump(db, f, indent=1)
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.txt"), "w") as f:
    for l in launches:
        f.write(f"{l['name'][:100]}\n   dur={(l['dur'] or 0) * 1e3:.3f} ms  dram rd={(l['rd'] or 0) / 1e9:.3f} GB wr={(l['wr'] or 0) / 1e9:.3f} GB  "
                f"dram%={l['dram_pct']}  tensor%={l['tensor_pct']}  issue%={l['issue_pct']}  occ%={l['occ_pct']}  regs={l['regs']}  "
                f"inst={l['inst']}  l2hit%={l['l2hit']}\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.txt")).read())
