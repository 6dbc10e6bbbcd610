"""Attribute the samples / executed instructions of an .ncu-rep to CUDA source lines
(development tool).  usage: ncu_lines.py report.ncu-rep cubin kernel-substring [top]
The SASS page of the report is joined, instruction by instruction, with `nvdisasm -g` of the
cubin (cuobjdump -xelf all libga_b200.so) -- both list the function's instructions in order."""
import csv
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines = []          # (file:line [inlined chain]) per instruction of the wanted function
cur = None
infn = False
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        infn = kname in m.group(1)
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = "%s:%s" % (m.group(1).split("/")[-1], m.group(2))
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        if inl:
            cur += " <- " + " <- ".join("%s:%s" % (f.split("/")[-1], l) for f, l in inl)
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
if len(data) != len(lines):
    print("warning: %d SASS rows in the report vs %d in the cubin" % (len(data), len(lines)))
agg = defaultdict(lambda: [0.0, 0.0])
ts = ti = 0.0
for r, l in zip(data, lines):
    s, n = float(r[ix["# Samples"]] or 0), float(r[ix["Instructions Executed"]] or 0)
    key = l.split(" <- ")[0] if "--full" not in sys.argv else l
    agg[key][0] += s
    agg[key][1] += n
    ts += s
    ti += n
print("samples %d, instructions %d" % (ts, ti))
for k, (s, n) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("  %-34s samples %5.1f%%  inst %5.1f%%" % (k, 100 * s / ts, 100 * n / ti))
