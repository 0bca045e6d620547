"""Prints a compact summary of the JSON lines and launch lists a GPU call left under gpurun_out/ (development helper)."""
import collections
import csv
import glob
import json
import os
import re
import sys

d0 = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
if os.path.exists(os.path.join(d0, "pytest_gpu.txt")):
    print(open(os.path.join(d0, "pytest_gpu.txt")).read()[-1500:])
for f in sorted(glob.glob(os.path.join(d0, "bench_*.json"))):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-1200:])
        continue
    print("=====", os.path.basename(f), "value %.3e  ms %.1f  e2e %.3e  rows %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("rows_per_step"), d.get("gpu_launches")))
    if "stage_ms" in d:
        print(" stage", {k: round(v, 1) for k, v in d["stage_ms"].items()})
    r = d.get("roofline")
    if not r:
        continue
    print(" roofline frac %.3f alone %s kernel_ms %.2f bytes %.3e" % (r["frac"], r.get("alone", {}).get("frac"), r["kernel_ms_per_step"], r["algorithmic_bytes_per_step"]), r.get("random_sector_ceiling", ""))
    if "debug" in d:
        g = d["debug"]
        print(" kernel_ms", {k: round(v, 1) for k, v in g["kernel_ms"].items()}, "wfa jobs/not-reg/general", g["wfa_jobs"], g["wfa_fallback_first"], g["wfa_general_jobs"], "per_round", g["wfa_per_round"], "surv", g["probe_survivors"])
    if d.get("cpu_baseline"):
        print(" cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:70])
for f in sorted(glob.glob(os.path.join(d0, "launches_*.csv"))):
    rows = list(csv.reader(open(f)))
    hdr, data = None, []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        n = re.sub(r"\(.*", "", d["Kernel Name"])[:60]
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = d["Metric Unit"]
        ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        agg[n][0] += 1
        agg[n][1] += ms
    tot = sum(v[1] for v in agg.values())
    print("-----", os.path.basename(f), "launches", len(data), "total ms %.1f" % tot)
    for n, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[:18]:
        print("%-62s %5d %9.3f ms %5.1f%%" % (n, c, ms, 100 * ms / tot))
