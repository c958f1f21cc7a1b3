#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list.
usage: python profiles/summarize_launches.py launches.csv [bench.json]  (prints a markdown table; with a bench line the
share of every tick kernel inside the tick is printed beside bench.py's own CUDA-event share)"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = next(r for r in rows if r[0] == "ID")
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows:
    if not r[0].isdigit():
        continue
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("msim::<unnamed>::", "").replace("unnamed>::", "").strip()
    v = float(r[ix["Metric Value"]])
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r[ix["Metric Unit"]], v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
TICK = {"query_kernel<1, 0, 1>": "query", "cell_scatter_kernel": "cell_scatter", "move_kernel<1, 0>": "move", "scan_tiles_kernel": "cell_scan",
        "scan_tile_sums_kernel": "cell_scan", "fold_counters_kernel": "query"}
bench = None
if len(sys.argv) > 2:
    bench = {k["name"]: k for k in json.loads(open(sys.argv[2]).read().splitlines()[-1])["kernels"]}
tot = sum(a[1] for a in agg.values())
tick_tot = sum(t / c for k, (c, t) in agg.items() if k in TICK)
print("| kernel | launches | total us | avg us | share of all launches | share of one tick (avg us / sum of tick kernels) | bench.py share (CUDA events) |\n|---|---|---|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tick_share = f"{100 * (t / c) / tick_tot:.1f}%" if k in TICK else ""
    b = ""
    if bench and k in TICK and TICK[k] in bench and not k.startswith(("scan_tile_sums", "fold")):
        tick_kernels = [v for n, v in bench.items() if n in ("query", "cell_scatter", "move", "cell_scan")]
        b = f"{100 * bench[TICK[k]]['avg_us'] / sum(v['avg_us'] for v in tick_kernels):.1f}%"
    print(f"| {k} | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% | {tick_share} | {b} |")
