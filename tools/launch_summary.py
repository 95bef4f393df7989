"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and share per kernel (template arguments kept).
usage: python tools/launch_summary.py launches.csv [--one-step] [header comment ...] > profiles/rN_bench_launch_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ci = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) != len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    val = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    name = r[ci["Kernel Name"]]
    name = re.sub(r"\(.*$", "", name).replace("void ", "")
    rows.append((name, us))
if "--one-step" in sys.argv:
    # the launch list covers several training steps: find the period of the kernel-name sequence and keep exactly one step
    sys.argv.remove("--one-step")
    names = [n for n, _ in rows]
    period = None
    for P in range(50, len(names) // 2):
        if all(names[i] == names[i + P] for i in range(len(names) - P)):
            period = P
            break
    if period is None:
        raise SystemExit("no periodic step structure found in the launch list")
    rows = rows[:period]
    sys.argv.append(f"exactly one step: {period} launches (period of the kernel-name sequence over {len(names)} captured launches)")
tot = sum(u for _, u in rows)
agg = defaultdict(lambda: [0.0, 0])
for n, u in rows:
    agg[n][0] += u
    agg[n][1] += 1
for c in sys.argv[2:]:
    print("#", c)
print(f"# launches {len(rows)} total_us {tot:.0f}")
for n, (u, k) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{u:10.1f} us {100 * u / tot:5.1f}%  n={k:4d}  {n[:150]}")
