"""Summarise an .ncu-rep (read here, no GPU): one line per profiled launch with the metrics the roofline needs."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}


def g(r, k, default=""):
    return r[ci[k]] if k in ci else default


print("# kernel | duration us | tensor-pipe % of peak (active) | DRAM read MB | DRAM write MB | DRAM throughput % | L2 throughput % | "
      "issue-slot busy % | regs | achieved warps/SM %")
for r in rows[2:]:
    name = g(r, "Kernel Name").replace("void ", "")[:70]
    print(" | ".join([
        name,
        f"{float(g(r, 'gpu__time_duration.sum', '0')):.1f}",
        f"{float(g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', '0') or 0):.1f}",
        f"{float(g(r, 'dram__bytes_read.sum', '0')):.1f} {units[ci['dram__bytes_read.sum']]}",
        f"{float(g(r, 'dram__bytes_write.sum', '0')):.1f} {units[ci['dram__bytes_write.sum']]}",
        f"{float(g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', '0')):.1f}",
        f"{float(g(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed', '0')):.1f}",
        f"{float(g(r, 'sm__inst_issued.avg.pct_of_peak_sustained_active', '0') or 0):.1f}",
        g(r, "launch__registers_per_thread"),
        f"{float(g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active', '0')):.1f}",
    ]))
