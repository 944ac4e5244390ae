#!/usr/bin/env python
"""Turn the ncu artefacts of a GPU visit (gpurun_out/) into the tracked summaries under profiles/.
usage: ncu_summary.py <round tag, e.g. r01>"""
import csv, json, os, re, subprocess, sys
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
# 1. launch list: per-kernel totals and shares
ll = os.path.join(G, "launches.csv")
if os.path.exists(ll):
    lines = [l for l in open(ll) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = {}
    order = []
    for r in rows:
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else (v * 1e6 if u == "s" else v))
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")[:90]
        if n not in agg:
            agg[n] = [0, 0.0]; order.append(n)
        agg[n][0] += 1; agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, "%s_launch_list.txt" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# command: see scripts/gpu_check.sh; %d launches, %.1f us total\n" % (len(rows), tot))
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%6.2f%%  %10.1f us  %4d x  %s\n" % (100 * t / tot, t, c, n))
    import shutil
    shutil.copy(ll, os.path.join(P, "%s_launches.csv" % tag))
# 2. full capture of the K1 tile kernel
rep = os.path.join(G, "k1_tile_full.ncu-rep")
if not os.path.exists(rep):
    rep = os.path.join(G, "k1_tile.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, "k1_tile_kernel", "30", "k1_tile_kernelILi8"],
                         capture_output=True, text=True).stdout
    with open(os.path.join(P, "%s_k1_tile_ncu_summary.txt" % tag), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:k1_tile (report: %s)\n" % os.path.basename(rep))
        f.write(out)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines())); ix = {k: i for i, k in enumerate(rr[0])}
    for r in rr[2:]:
        if "k1_tile_kernel" in r[ix["Kernel Name"]]:
            def val(k):
                v = float(r[ix[k]].replace(",", "")); u = rr[1][ix[k]]
                return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1)
            meta = {}
            mp = os.path.join(G, "k1_tile_full.meta.json")
            if os.path.exists(mp):
                meta = json.load(open(mp))
            json.dump({"kernel": "k1_tile_kernel", "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                       "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
                       "records": meta.get("records"), "source": os.path.basename(rep), "round": tag},
                      open(os.path.join(P, "k1_traffic.json"), "w"), indent=1)
            break
print(os.listdir(P))
