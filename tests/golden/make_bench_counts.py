#!/usr/bin/env python
"""Generator of tests/golden/bench_counts.json: unique in-range pairs and flagged entities of bench.py's population after a given
number of move passes, computed by the CPU ORACLE (oracle/msim_oracle.c) - test infrastructure, never the product path.

bench.py prints `move_passes_done`, `pairs_last_tick` and `flagged_last_tick` at the end of its timed region in every line and
compares them with the stored values (bench.check_counts): the same population at the same tick must give the same numbers on
1, 2, 4 and 8 GPUs.  The tick index at the end of the timed region follows from (pre-roll, warm-up W, steps K):
    move passes = pre-roll + 1 + max(3, W) + alignment + K,   alignment = ticks until the next tick is a re-sorting one
(bench.align_resort_phase; the first re-sort follows the first collision pass, then one every 32 collision passes).

usage: python tests/golden/make_bench_counts.py [--workload munich_10m_collisions] [--entities N] [--threads T]
Takes a few minutes at 10 M entities (one oracle move pass = 50 ms, one oracle collision pass = ~1 s on 8 cores)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import movement_sim_b200 as M  # noqa: E402  (host helpers only)
from oracle import oracle as O  # noqa: E402


def passes_done(preroll: int, warmup: int, steps: int) -> int:
    w = max(3, warmup)
    collide_before = 1 + w  # the untimed tick behind the pre-roll re-sorts (first collision pass after the upload)
    # re-sorts happen on collision passes 1, 33, 65, ...: tick until one has happened, then 31 more
    to_next = (bench.RESORT_EVERY - (collide_before - 1) % bench.RESORT_EVERY) % bench.RESORT_EVERY or bench.RESORT_EVERY
    return preroll + collide_before + to_next + (bench.RESORT_EVERY - 1) + steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="munich_10m_collisions")
    ap.add_argument("--entities", type=int, default=None)
    ap.add_argument("--preroll", type=int, default=256)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--combos", default="5:20,3:20,10:200,5:32,5:64,10:20,5:10,3:10,5:50,5:100,5:200,10:100,10:50")
    ap.add_argument("--passes", default=None, help="explicit move-pass counts instead of --combos (tests/test_gpu_configs45.py), e.g. 40,72")
    args = ap.parse_args()
    w, m = bench.build_workload(M, args.workload, args.entities)
    n = w["entities"]
    if args.passes:
        want = sorted({int(p) for p in args.passes.split(",")})
    else:
        want = sorted({passes_done(args.preroll, int(c.split(":")[0]), int(c.split(":")[1])) for c in args.combos.split(",")})
    e = np.ascontiguousarray(bench.build_population(M, m, n, w["box"])).view(O.ENTITY_DTYPE).copy()
    om = O.OracleMap(m.width, m.height, m.roads.view(O.ROAD_DTYPE), m.connections)
    O.move_pass(e, om, threads=args.threads)  # the init-only first dispatch
    out = {}
    done = 0
    for target in want:
        while done < target:
            O.move_pass(e, om, threads=args.threads)
            done += 1
        pairs = O.collide_pass(e, m.width, m.height, 10.0, threads=args.threads)
        flagged = int(O.collision_flags(e).sum())
        out[str(target)] = {"pairs": int(pairs), "flagged": flagged}
        print(target, out[str(target)], flush=True)
    path = bench.COUNTS_PATH
    stored = {}
    if os.path.exists(path):
        with open(path) as f:
            stored = json.load(f)
    stored.setdefault(args.workload, {}).setdefault(str(n), {}).update(out)
    with open(path, "w") as f:
        json.dump(stored, f, indent=1, sort_keys=True)
        f.write("\n")


if __name__ == "__main__":
    main()
