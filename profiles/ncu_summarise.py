#!/usr/bin/env python
"""Summaries of the round's ncu captures for profiles/ (read with `ncu -i ... --page raw --csv` here, no GPU needed).
usage: python profiles/ncu_summarise.py <launches.csv> <rep> [<rep> ...]  ->  markdown on stdout, traffic JSON on stderr"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict

WANT = OrderedDict([
    ("gpu__time_duration.sum", "duration us"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("dram__bytes_read.sum", "dram read MB"),
    ("dram__bytes_write.sum", "dram write MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
])


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.split("::")[-1].replace("unnamed>", "").strip()


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[1:]:
        k = short(r[ix["Kernel Name"]])
        t = float(r[ix["Metric Value"]].replace(",", "")) / 1e3
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    return agg


def full(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": short(vals[hdr.index("Kernel Name")])}
    for m, label in WANT.items():
        if m in hdr:
            i = hdr.index(m)
            v = vals[i].replace(",", "")
            try:
                v = float(v)
                if units[i] in ("byte", "Byte", "bytes"):
                    v /= 1e6
                elif units[i] in ("ns", "nsecond"):
                    v /= 1e3
                elif units[i] == "Kbyte":
                    v /= 1e3
                elif units[i] == "Gbyte":
                    v *= 1e3
            except ValueError:
                pass
            d[label] = v
    return d


def main():
    la = launches(sys.argv[1])
    tick_prefixes = ("query_tiles_kernel", "cell_scatter_slots_kernel", "move_kernel<1", "scan_cells_kernel", "fold_stripes_kernel")
    tick = [k for k in la if k.startswith(tick_prefixes)]
    tick_sum = sum(la[k][1] / la[k][0] for k in tick)
    print("| kernel | launches | total us | avg us | share of one tick (avg / sum of the main-stream tick kernels) |")
    print("|---|---|---|---|---|")
    for k, (c, t) in sorted(la.items(), key=lambda kv: -kv[1][1]):
        share = f"{100 * (t / c) / tick_sum:.1f} %" if k in tick else ""
        print(f"| {k} | {c} | {t:.1f} | {t / c:.1f} | {share} |")
    print()
    reps = [full(r) for r in sys.argv[2:]]
    labels = ["kernel"] + list(WANT.values())
    print("| " + " | ".join(labels) + " |")
    print("|" + "---|" * len(labels))
    for d in reps:
        cells = []
        for l in labels:
            v = d.get(l, "")
            cells.append(f"{v:.1f}" if isinstance(v, float) and l not in ("warp instructions",) else (f"{v:,.0f}" if isinstance(v, float) else str(v)))
        print("| " + " | ".join(cells) + " |")
    traffic = {d["kernel"]: (d.get("dram read MB", 0) + d.get("dram write MB", 0)) * 1e6 for d in reps}
    issue = {d["kernel"]: d.get("issue active %") for d in reps}
    sys.stderr.write(json.dumps({"traffic": traffic, "issue": issue}) + "\n")


if __name__ == "__main__":
    main()
