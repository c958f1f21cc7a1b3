#!/usr/bin/env python
"""One markdown row per captured launch of an `ncu --set full` report.  usage: python profiles/summarize_ncu.py report.ncu-rep"""
import csv
import io
import re
import subprocess
import sys

M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
     "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__grid_size"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(M)], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
print("| kernel | grid | time us | dram read MB | dram write MB | dram % of peak | issue active % | warps active % | regs | warp inst | threads/inst | L1 hit % | L2 hit % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")


def val(r, m, scale=1.0, fmt="%.1f"):
    if m not in ix or r[ix[m]] in ("", "n/a"):
        return "-"
    v = float(r[ix[m]].replace(",", ""))
    u = units[ix[m]]
    if u == "byte":
        v /= 1e6
    elif u == "Kbyte":
        v /= 1e3
    elif u == "Gbyte":
        v *= 1e3
    elif u == "ns":
        v /= 1e3
    elif u == "ms":
        v *= 1e3
    return fmt % (v * scale)


for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("msim::<unnamed>::", "").replace("unnamed>::", "").strip()
    print("| %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %s |" % (
        name, val(r, "launch__grid_size", fmt="%d"), val(r, M[0]), val(r, M[1]), val(r, M[2]), val(r, M[3]), val(r, M[4]), val(r, M[5]),
        val(r, M[6], fmt="%d"), val(r, M[7], fmt="%d"), val(r, M[8], fmt="%.2f"), val(r, M[9]), val(r, M[10])))
