"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of one step.
usage: python scripts/launch_breakdown.py file.csv [step_index_from_end]"""
import collections, csv, re, sys
path = sys.argv[1]
back = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with open(path) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
names = [r["Kernel Name"] for r in rows]
starts = [i for i, n in enumerate(names) if "csr_convert" in n]
# a training step builds two CSRs (tt, tb): steps start at every other csr_convert launch
steps = starts[::2]
b = steps[-back - 1]
e = steps[-back]
agg, tot = collections.OrderedDict(), 0.0
for r in rows[b:e]:
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("sgb::<unnamed>::", "")
    t = float(r["Metric Value"].replace(",", ""))
    t = t / 1e3 if r["Metric Unit"] == "ns" else (t * 1e3 if r["Metric Unit"] == "ms" else t)
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
print(f"one step = launches [{b},{e}) : {e-b} launches, {tot/1e3:.2f} ms of kernel time (cold-cache, serialised)")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  x{c:3d}  {n[:100]}")
