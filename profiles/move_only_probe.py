#!/usr/bin/env python
"""Move-only ticks at 10 M entities: per-kernel CUDA-event times, (a) on a collisions-off handle, (b) on a collisions-on handle
whose storage is in cell order (what bench.py's `move_only` leg measures).  usage: python profiles/move_only_probe.py [entities]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import movement_sim_b200 as M  # noqa: E402
from bench import build_workload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
w, m = build_workload(M, "munich_10m_collisions", n)
ents = m.init_entities(n, seed=42)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    for label, flags, warm_collide in (("collisions-off handle", M.FLAG_NO_COLLISIONS, 0), ("collisions-on handle after 40 full ticks", 0, 40)):
        sim = M.Simulation(m, ents, radius=10.0, flags=flags, stream=stream.cuda_stream)
        sim.dispatch(2)
        sim.enqueue_ticks(256, False)
        if warm_collide:
            sim.enqueue_ticks(warm_collide, True)
        sim.enqueue_ticks(10, False)
        sim.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sim.enqueue_ticks(200, False)
        e1.record(stream)
        e1.synchronize()
        sim.profile_begin()
        sim.enqueue_ticks(100, False)
        kt = sim.profile_end()
        print(label, "us per move-only tick: %.1f" % (e0.elapsed_time(e1) / 200 * 1e3), {k: round(t / c * 1e3, 1) for k, (c, t) in kt.items()})
        sim.close()
