"""Where a pipelined e2e step spends its wall time: per-call host timing of upload / dispatch / snapshot_end / snapshot_begin at 10 M entities."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
import movement_sim_b200 as M

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
w, m = bench.build_workload(M, "munich_10m_collisions", n)
ents = bench.build_population(M, m, n, None)
stream = torch.cuda.Stream()
sim = M.Simulation(m, ents, radius=10.0, stream=stream.cuda_stream)
sim.dispatch(2)
sim.enqueue_ticks(64, False)
sim.enqueue_ticks(1, True)
pinned = torch.empty(n * 64, dtype=torch.uint8, pin_memory=True)
ptr = pinned.data_ptr()
sim.read_entities_ptr(ptr, n)
tick = 4
rows = []
for it in range(7):
    t = [time.perf_counter()]
    sim.upload_ptr(ptr, n); t.append(time.perf_counter())
    sim.dispatch(tick); t.append(time.perf_counter())
    sim.dispatch(tick + 1); t.append(time.perf_counter())
    tick += 2
    if it:
        sim.snapshot_end(copy=False)
    t.append(time.perf_counter())
    sim.snapshot_begin(); t.append(time.perf_counter())
    rows.append([round((b - a) * 1e3, 3) for a, b in zip(t, t[1:])])
sim.snapshot_end(copy=False)
print(json.dumps({"columns": ["upload", "dispatch_move", "dispatch_collide", "snapshot_end", "snapshot_begin"], "ms": rows}))
