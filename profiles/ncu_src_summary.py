#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: total stall samples per reason and the hottest SASS lines.
usage: ncu -i rep.ncu-rep --page source --csv -k regex:<kernel> -c 1 > src.csv; python ncu_src_summary.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(rows[0][1][:120])
hdr = rows[1]; data = [r for r in rows[2:] if len(r) >= len(rows[1])]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
total = 0
for r in data:
    for s in stalls:
        try: tot[s] += int(r[ix[s]])
        except ValueError: pass
    try: total += int(r[ix["# Samples"]])
    except ValueError: pass
print("total samples", total)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    print("  %-24s %7d  %5.1f%%" % (s, v, 100.0 * v / max(total, 1)))
def samples(r):
    try: return int(r[ix["# Samples"]])
    except ValueError: return 0
for r in sorted(data, key=samples, reverse=True)[:top]:
    why = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print("%6d %5.1f%%  %-70s %s" % (samples(r), 100.0 * samples(r) / max(total, 1), r[ix["Source"]].strip()[:70], " ".join("%s=%d" % (s[6:], v) for v, s in why if v)))
