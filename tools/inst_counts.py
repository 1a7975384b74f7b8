#!/usr/bin/env python3
"""Fold an ncu launch list with instruction counters into profiles/r2_inst_counts.json (read by bench.py's int_issue roofline).

  ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/inst_<config>.csv python tools/prof_stage.py --config <config> --reads 100000 --iters 2
  python tools/inst_counts.py <config> <reads> gpurun_out/inst_<config>.csv
"""
import csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
name, reads, path = sys.argv[1], int(sys.argv[2]), sys.argv[3]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
# keep the LAST step only (a step starts with k_chunk_count; the first step of a context may re-run the scan to grow its buffers)
starts = [int(r["ID"]) for r in rows if "k_chunk_count" in r["Kernel Name"]]
if starts:
    rows = [r for r in rows if int(r["ID"]) >= max(starts)]
kern = {}
for r in rows:
    k = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    k = re.sub(r"^void ", "", k)
    d = kern.setdefault(k, dict(thread_inst=0, warp_inst=0, time_us=0.0, launches=set()))
    v = float(r["Metric Value"].replace(",", ""))
    m = r["Metric Name"]
    if m == "smsp__thread_inst_executed.sum": d["thread_inst"] += int(v)
    elif m == "smsp__inst_executed.sum": d["warp_inst"] += int(v)
    elif m == "gpu__time_duration.sum": d["time_us"] += v / (1000.0 if r["Metric Unit"] in ("nsecond", "ns") else 1.0)
    d["launches"].add(r["ID"])
for d in kern.values():
    d["launches"] = len(d["launches"])
out_p = os.path.join(ROOT, "profiles", "r2_inst_counts.json")
allc = json.load(open(out_p)) if os.path.exists(out_p) else {}
allc[name] = dict(reads=reads, read_len=10000, kernels=kern,
                  note="one step (all kernel launches of one bb_annotate_device call, library kernels such as cub included) under ncu; "
                       "thread_inst = smsp__thread_inst_executed.sum (lane-operations), warp_inst = smsp__inst_executed.sum; times are ncu's "
                       "cold-cache serialised launch times -- compare shares, not absolutes")
json.dump(allc, open(out_p, "w"), indent=1, sort_keys=True)
tot = sum(d["thread_inst"] for d in kern.values())
for k, d in sorted(kern.items(), key=lambda kv: -kv[1]["thread_inst"]):
    print(f"{k[:70]:70s} launches {d['launches']:3d} lane-ops {d['thread_inst']:14d} ({100.0 * d['thread_inst'] / tot:5.1f} %)  time {d['time_us']:9.1f} us")
print("lane-ops per base:", tot / (reads * 10000.0))
