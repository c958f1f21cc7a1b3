#!/usr/bin/env python
"""One table over several bench.py JSON lines (the knob sweep of profiles/run_r2_first.sh):
    python profiles/compare_bench.py gpurun_out/r2a_bench_*.json
columns: file, us per tick, entity-updates/s, every kernel's average launch time, move_only / flags_only / e2e extras.  The first file
is the reference the deltas are taken against (put the default run first)."""
import json
import os
import sys


def load(path):
    try:
        with open(path) as f:
            text = f.read().strip()
        return json.loads(text.splitlines()[-1]) if text else None
    except (OSError, ValueError):
        return None


def main():
    rows = [(os.path.basename(p), load(p)) for p in sys.argv[1:]]
    rows = [(n, d) for n, d in rows if d and "ms_per_step" in d]
    if not rows:
        print("no readable bench lines")
        return 1
    kernels = []
    for _, d in rows:
        for k in d.get("kernels") or []:
            if k["name"] not in kernels:
                kernels.append(k["name"])
    base = rows[0][1]["ms_per_step"] * 1e3
    head = ["run", "us/tick", "delta", "G upd/s"] + kernels + ["move_only us", "flags_only us", "e2e ms", "e2e piped ms", "experiments"]
    print("| " + " | ".join(head) + " |")
    print("|" + "---|" * len(head))
    for name, d in rows:
        us = d["ms_per_step"] * 1e3
        kt = {k["name"]: k["avg_us"] for k in d.get("kernels") or []}
        ex = (d.get("config") or {}).get("experiments") or {}
        cells = [name, f"{us:.1f}", f"{(us / base - 1) * 100:+.1f} %", f"{d['value'] / 1e9:.2f}"]
        cells += [f"{kt[k]:.1f}" if k in kt else "-" for k in kernels]
        mo, fo, e2e = d.get("move_only") or {}, d.get("flags_only") or {}, d.get("e2e") or {}
        cells.append(f"{mo['ms_per_step'] * 1e3:.1f}" if "ms_per_step" in mo else "-")
        cells.append(f"{fo['ms_per_step'] * 1e3:.1f}" if "ms_per_step" in fo else "-")
        cells.append(f"{e2e['ms_per_step']:.2f}" if "ms_per_step" in e2e else "-")
        cells.append(f"{e2e['pipelined']['ms_per_step']:.2f}" if "pipelined" in e2e else "-")
        cells.append(", ".join(f"{k}={v}" for k, v in ex.items() if v))
        print("| " + " | ".join(cells) + " |")
    return 0


if __name__ == "__main__":
    sys.exit(main())
