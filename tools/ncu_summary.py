"""Summarise an .ncu-rep (development tool): key metrics + sample distribution by opcode / region."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
for i, h in enumerate(hdr):
    if h in want:
        print("%-70s %s %s" % (h, vals[i], units[i]))
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        v = float(vals[i])
        if v > 0.05:
            print("  stall %-40s %.2f" % (h.split("stalled_")[1].split("_per_issue")[0], v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


tot = sum(f(r, "# Samples") for r in data) or 1
ti = sum(f(r, "Instructions Executed") for r in data) or 1
op = defaultdict(lambda: [0, 0])
for r in data:
    s = r[ix["Source"]].split()
    o = s[0] if s else "?"
    if o.startswith("@") and len(s) > 1:
        o = s[1]
    op[o][0] += f(r, "# Samples")
    op[o][1] += f(r, "Instructions Executed")
print("samples %d, instructions %d" % (tot, ti))
for o, (s, n) in sorted(op.items(), key=lambda x: -x[1][0])[:14]:
    print("  %-20s samples %5.1f%%   inst %5.1f%% (%d)" % (o, 100 * s / tot, 100 * n / ti, n))
# regions by execution count
cnt = defaultdict(lambda: [0, 0, 0])
for r in data:
    n = int(f(r, "Instructions Executed"))
    cnt[n][0] += f(r, "# Samples")
    cnt[n][1] += n
    cnt[n][2] += 1
print("regions by per-line execution count (top by samples):")
for n, (s, t, l) in sorted(cnt.items(), key=lambda x: -x[1][0])[:10]:
    print("  exec/line %9d  lines %4d  samples %5.1f%%  inst %5.1f%%" % (n, l, 100 * s / tot, 100 * t / ti))
