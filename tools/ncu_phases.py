"""Phase breakdown (samples = time, instructions) of nn_fwd_mma_kernel from an .ncu-rep.
Development tool; line ranges below must follow nn_mma.cuh / nn_search.cuh / nn_tiles.cuh.
usage: ncu_phases.py report.ncu-rep cubin kernel-substring warps_total"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1:4]
nwarps = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
R = [  # (file, first, last, phase), first match along [innermost, parent] wins
    ("nn_mma.cuh", 72, 90, "stage_bfrag"), ("nn_mma.cuh", 93, 124, "A fragments"), ("nn_mma.cuh", 135, 173, "mma_scan"),
    ("nn_mma.cuh", 46, 51, "mma_scan"), ("ga_common.cuh", 48, 53, "mma_scan"),
    ("nn_search.cuh", 45, 66, "refine: tile scan"), ("nn_search.cuh", 69, 79, "refine: eval"),
    ("nn_search.cuh", 86, 121, "refine: warp_exact_scan"), ("nn_mma.cuh", 179, 219, "refine: other"),
    ("nn_tiles.cuh", 21, 67, "stage_targets"), ("nn_tiles.cuh", 70, 76, "refine: tile scan"),
    ("ga_common.cuh", 37, 46, "refine: tile scan"), ("ga_common.cuh", 56, 64, "refine: eval"),
    ("nn_search.cuh", 257, 262, "output"),
]


def fwd_phase(l):
    if l <= 64: return "setup"
    if l <= 88: return "query state init"
    if l <= 100: return "chunk loop"
    if l == 101: return "stage_targets"
    if l == 102: return "stage_bfrag"
    if l <= 105: return "zero+sync"
    if l <= 108: return "mma_scan"
    if l <= 131: return "publish"
    if l <= 143: return "read lists"
    if l == 144: return "refine: other"
    return "output"


def phase(chain):
    for f, l in chain:
        for rf, a, b, p in R:
            if f == rf and a <= l <= b:
                return p
    for f, l in chain:
        if f == "nn_distance_fwd_mma.cu":
            return fwd_phase(l)
    return "?" + str(chain[:2])


dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
lines, cur, infn = [], None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = [(m.group(1).split("/")[-1], int(m.group(2)))] + [(f.split("/")[-1], int(l)) for f, l in re.findall(
            r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0])
ts = ti = 0
for r, l in zip(data, lines):
    s = float(r[ix["# Samples"]] or 0)
    n = float(r[ix["Instructions Executed"]] or 0)
    k = phase(l) if l else "none"
    agg[k][0] += s
    agg[k][1] += n
    ts += s
    ti += n
for k, (s, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-28s samples %5.1f%%  inst %5.1f%%  inst/warp %7.0f" % (k, 100 * s / ts, 100 * n / ti, n / nwarps))
