"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:  python tools/launch_shares.py in.csv out.csv
Also prints, for the headline kernel, the launches grouped by grid size (the timed step of bench.py is the 8800-CTA launch)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
grids = collections.defaultdict(list)
for r in rows:
    name, grid, val, unit = r[4], r[8], float(r[-1].replace(",", "")), r[-2]
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
    a = agg.setdefault(name[:80], [0, 0.0])
    a[0] += 1
    a[1] += us
    if "iou_tile_kernel<1, 1" in name:
        grids[(name.split("(")[0].replace("void glenet::", ""), grid)].append(us)
# the FP32-peak microkernel bench.py runs once (tools/cuda/ffma_peak.cu, ~0.5 s sustained) is listed but kept out of the shares
total = sum(a[1] for k, a in agg.items() if not k.startswith("ffma_kernel"))
with open(sys.argv[2], "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_us", "share"])
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, round(t, 1), "(measurement microkernel)" if k.startswith("ffma_kernel") else round(t / total, 3)])
for g, v in grids.items():
    print(f"{g[0]} grid {g[1]}: {len(v)} launches, mean {sum(v) / len(v):.1f} us")
