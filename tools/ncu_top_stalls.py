"""Print the instructions with the most warp-stall samples from `ncu --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
data = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ci = {name: i for i, name in enumerate(hdr)}


def num(r, c):
    try:
        return int(r[ci[c]])
    except Exception:
        return 0


tot = sum(num(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
stallcols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: sum(num(r, c) for r in data) for c in stallcols}
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:n]:
    st = sorted([(num(r, c), c) for c in stallcols], reverse=True)[:2]
    print(str(num(r, "# Samples")).rjust(7), r[ci["Instructions Executed"]].rjust(10), r[ci["Source"]].strip()[:90].ljust(90), st)
