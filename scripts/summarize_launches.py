"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name for the
LAST full step of the run (steps are delimited by the first library kernel of a step: knn_kernel<16> on
the level-0 cloud)."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((int(r["ID"]), r["Kernel Name"], ns))
print(f"{len(rows)} launches captured, total {sum(r[2] for r in rows)/1e6:.3f} ms")

def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("b200::", "").replace("at::native::", "")
    return name[:110]

# split into steps: a step starts at a knn_kernel launch preceded (anywhere since the previous start) by
# an optimizer kernel; simpler: starts = indices of the 1st knn_kernel after >= 200 other launches
starts, last = [], -10**9
for i, (_, name, _) in enumerate(rows):
    if "knn_kernel" in name and i - last > 400:
        starts.append(i)
        last = i
    elif "knn_kernel" in name:
        last = last  # same step
print("step starts at launch indices", starts)
if "--all" in sys.argv:  # the capture is already exactly one step (B200_NCU_RANGE=1 / --profile-from-start off)
    seg = rows
elif len(starts) >= 2:
    seg = rows[starts[-2]:starts[-1]]
else:
    seg = rows
agg = collections.OrderedDict()
for _, name, ns in seg:
    k = short(name)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(v[1] for v in agg.values())
print(f"one step: {len(seg)} launches, {tot/1e6:.3f} ms of kernel time")
ours = sum(v[1] for k, v in agg.items() if "_kernel" in k and ("lfa" in k or "knn" in k or "linear" in k or "affine" in k or "bn_finalize" in k or "rows" in k or "interp" in k or "moments" in k or "tc_" in k or "tn_skinny" in k or "grid_" in k or "fold" in k or "adam" in k or "increment" in k))
print(f"libb200randla kernels: {ours/1e6:.3f} ms ({100*ours/tot:.1f} %), other (torch) kernels: {(tot-ours)/1e6:.3f} ms")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{ns/1e3:10.1f} us {100*ns/tot:5.1f}%  n={n:4d}  {k}")
