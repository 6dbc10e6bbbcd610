"""Turn gpurun_out/ artefacts into the tracked summaries under profiles/ (development tool)."""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
for rep in sorted(os.listdir(G)):
    if rep.startswith(tag + "_") and rep.endswith(".ncu-rep"):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, rep)],
                             capture_output=True, text=True).stdout
        with open(os.path.join(P, rep.replace(".ncu-rep", "_ncu_summary.txt")), "w") as f:
            f.write("# ncu --set full --clock-control none --import-source on, summarised by tools/ncu_summary.py\n" + out)
        raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) >= 3:
            h, v = rows[0], rows[2]
            d = dict(zip(h, v))
            kn = d.get("Kernel Name", "").split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip()
            if kn.startswith("nn_fwd") and "_fwd" in rep:
                units = dict(zip(h, rows[1]))
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

                def mb(x):
                    return float(d[x]) * scale.get(units.get(x, "byte"), 1.0) if d.get(x) else 0.0
                tp = os.path.join(P, "traffic.json")
                tr = json.load(open(tp)) if os.path.exists(tp) else {}
                tr[kn + "_b50_dram_bytes_per_launch"] = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
                tr[kn + "_source"] = rep
                tr["note"] = "dram__bytes_read.sum + dram__bytes_write.sum, bytes per launch"
                json.dump(tr, open(tp, "w"), indent=1)
for name in ("tune.json", "microbench.txt", "loopbench.txt", "bench_%s.json" % tag):
    src = os.path.join(G, name)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, (tag + "_" + name) if not name.startswith("bench_") else name))
ll = os.path.join(G, "launches_bench_%s.csv" % tag)
if os.path.exists(ll):
    rows = [r for r in csv.reader(open(ll)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"]
    rows = rows[hdr[0]:] if hdr else rows
    with open(os.path.join(P, "%s_launches_bench.csv" % tag), "w") as f:
        w = csv.writer(f); w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration_ns"])
        tot = {}
        for r in rows[1:]:
            k = r[4].split("(")[0]; w.writerow([r[0], k, r[7] if len(r) > 7 else "", r[8] if len(r) > 8 else "", r[-1]])
            try: tot[k] = tot.get(k, 0.0) + float(r[-1])
            except ValueError: pass
    s = sum(tot.values()) or 1
    with open(os.path.join(P, "%s_launches_bench_shares.txt" % tag), "w") as f:
        f.write("# share of summed gpu__time_duration per kernel (ncu launch list; cold-cache, serialised: compare shares)\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write("%-60s %10.1f us  %5.1f%%\n" % (k, v / 1e3, 100 * v / s))
print(sorted(os.listdir(P)))
