#!/usr/bin/env python
"""Per-source-line instruction counts / stall samples of one kernel from an ncu report (no GPU needed).
usage: ncu_lines.py report.ncu-rep <kernel substring> [top]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
sect = sys.argv[4] if len(sys.argv) > 4 else kern      # mangled-name substring selecting ONE instantiation in the cubin
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s/phaser_b200/_phz.so >/dev/null 2>&1 && nvdisasm --print-line-info *.cubin > all.sass" % (tmp, ROOT), shell=True, check=True)
lines = open(os.path.join(tmp, "all.sass")).read().split("\n")
start = next(i for i, l in enumerate(lines) if ".section" in l and ".text." in l and sect in l)
end = next(i for i in range(start + 1, len(lines)) if ".section" in lines[i])
cur = None; a2l = {}
for l in lines[start + 1:end]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        a2l[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h = rr[0]; ix = {k: i for i, k in enumerate(h)}
for r in rr[2:]:
    if kern in r[ix["Kernel Name"]]:
        for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                  "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
                  "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
                  "smsp__thread_inst_executed_per_inst_executed.ratio"]:
            print(k, "=", r[ix[k]], rr[1][ix[k]])
        st = [(float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else 0, k) for k, i in ix.items() if "issue_stalled" in k and "per_issue_active" in k]
        print("stalls:", ", ".join("%s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, k in sorted(st, reverse=True)[:7]))
        break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {k: i for i, k in enumerate(hdr)}
base = None; agg = collections.Counter(); smp = collections.Counter(); tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    base = a if base is None else base
    ie = int(r[idx["Instructions Executed"]]); s = int(r[idx["# Samples"]])
    ln = a2l.get(a - base)
    agg[ln] += ie; smp[ln] += s; tot += ie
ts = max(1, sum(smp.values()))
srcs = {}
print("total warp-instructions", tot, "samples", ts)
for ln, v in agg.most_common(top):
    t = ""
    if ln:
        p = os.path.join(ROOT, "phaser_b200", "csrc", ln[0])
        if os.path.exists(p):
            srcs.setdefault(p, open(p).read().split("\n"))
            t = srcs[p][ln[1] - 1].strip()[:95]
    print("%-28s inst %5.1f%%  stall-samples %5.1f%%  %s" % ("%s:%d" % ln if ln else "?", 100 * v / tot, 100 * smp[ln] / ts, t))
