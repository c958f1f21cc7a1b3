"""Regenerates tests/golden/oracle_digests.json: SHA-256 digests and sample entities of oracle runs on
the BASELINE configs that fit a CPU test (config 1 in full; a small Munich-style city with collisions).
The reference is GLSL and cannot run here, so these vectors pin the ORACLE (and through the GPU parity
tests, the CUDA path) against regressions; they are not reference outputs (see oracle/msim_oracle.h).

Run: python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _digest(e):
    return hashlib.sha256(np.ascontiguousarray(e).tobytes()).hexdigest()


def compute_digests():
    import movement_sim_b200 as M
    from oracle import oracle as O

    out = {}
    # config 1: test_map.json, 10 k entities, seed 42, 1000 move passes, collisions off
    m = M.Map.load_json(os.path.join(HERE, "test_map.json"))
    om = O.OracleMap(m.width, m.height, m.roads.view(O.ROAD_DTYPE), m.connections)
    e = np.ascontiguousarray(m.init_entities(10_000, seed=42)).view(O.ENTITY_DTYPE).copy()
    out["config1_initial"] = _digest(e)
    O.move_pass(e, om)
    totals = {"arrivals": 0, "rng_draws": 0, "uturns": 0, "oob_reads": 0}
    for t in range(1000):
        st = O.move_pass(e, om)
        for k in totals:
            totals[k] += st[k]
        if t + 1 in (1, 100, 1000):
            out[f"config1_after_{t + 1}_moves"] = _digest(e)
    out["config1_counters"] = totals
    out["config1_entity_0"] = {k: np.asarray(e[0][k]).tolist() for k in ("pos", "target", "dir", "road", "rng")}
    # Munich-style city, collisions on, 40 sim ticks
    c = M.Map.city(2200.0, 1600.0, 35.0, 0.3, 0.12, 7)
    oc = O.OracleMap(c.width, c.height, c.roads.view(O.ROAD_DTYPE), c.connections)
    out["city_map"] = {"roads": int(c.roads.shape[0]), "connections": int(c.connections.shape[0]),
                       "roads_sha256": hashlib.sha256(c.roads.tobytes()).hexdigest(),
                       "connections_sha256": hashlib.sha256(c.connections.tobytes()).hexdigest()}
    e = np.ascontiguousarray(c.init_entities(30_000, seed=42)).view(O.ENTITY_DTYPE).copy()
    pairs = []
    for tick in range(2, 2 + 2 * 40):
        p, _ = O.dispatch(e, oc, 10.0, tick)
        if tick % 2 == 1:
            pairs.append(p)
    out["city_after_40_ticks"] = _digest(e)
    out["city_pairs_first_last"] = [pairs[0], pairs[1], pairs[-1]]
    out["city_flagged"] = int(O.collision_flags(e).sum())
    return out


if __name__ == "__main__":
    d = compute_digests()
    with open(os.path.join(HERE, "oracle_digests.json"), "w") as f:
        json.dump(d, f, indent=1, sort_keys=True)
        f.write("\n")
    print(json.dumps(d, indent=1, sort_keys=True))
