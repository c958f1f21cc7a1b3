#!/usr/bin/env python
"""Two (or more) bands of the Munich stand-in on ONE GPU, peer-memory exchange through local arenas: a single-process
stand-in for the per-GPU work of a multi-GPU tick that ncu can profile (ncu must not wrap a multi-rank command).
usage: python profiles/band_microbench.py [total_entities] [bands] [ticks]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import movement_sim_b200 as M  # noqa: E402
from bench import build_workload  # noqa: E402
from movement_sim_b200 import sharding as S  # noqa: E402

total = int(sys.argv[1]) if len(sys.argv) > 1 else 2_500_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 100
w, m = build_workload(M, "munich_10m_collisions", total)
hist, ncx, ncy = S.global_row_histogram(M, m, total, 42, 10.0)
splits = S.balanced_splits(hist, world)
max_row = int(hist.max())
cap_x = max(4096, 3 * max_row)
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    sims = []
    for r in range(world):
        ents, gids = S.collect_band(M, m, total, 42, 10.0, int(splits[r]), int(splits[r + 1]))
        sim = M.Simulation(m, ents, radius=10.0, stream=stream.cuda_stream, capacity=int(ents.shape[0] * 1.3) + 8 * cap_x + 1024)
        sim.shard_enable(gids, cap_x, cap_x)
        sim.dispatch(2)
        sims.append(sim)
    arenas = [s.shard_p2p_create()[1] for s in sims]
    for r, s in enumerate(sims):
        s.shard_p2p_connect_local(arenas[r - 1] if r > 0 else None, arenas[r + 1] if r + 1 < world else None)

    def tick(collide=True):
        for r, s in enumerate(sims):
            s.shard_p2p_move_pack(int(splits[r]), int(splits[r + 1]))
        for s in sims:
            s.shard_p2p_integrate()
            if collide:
                s.enqueue_collide()

    for _ in range(64):
        tick(False)
    for _ in range(10):
        tick()
    sims[0].sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(ticks):
        tick()
    e1.record(stream)
    e1.synchronize()
    print("us per tick (all %d bands on one GPU): %.1f" % (world, e0.elapsed_time(e1) / ticks * 1e3))
    sims[0].profile_begin()
    for _ in range(ticks):
        tick()
    kt = sims[0].profile_end()
    print({k: round(t / ticks * 1e3, 1) for k, (c, t) in sorted(kt.items(), key=lambda kv: -kv[1][1])})
    print("owned", [s.stats()["entity_count"] for s in sims], "pairs", sum(s.stats()["last_pair_count"] for s in sims))
    for s in sims:
        s.close()
