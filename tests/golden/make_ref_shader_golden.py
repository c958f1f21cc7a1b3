#!/usr/bin/env python
"""Digests of the entity state produced by the reference's OWN movement code — random_move.comp compiled for the CPU
(oracle/_ref/libref_shader_move.so, built by oracle/Makefile from /root/reference) — for the cases
tests/test_oracle_vs_ref_shader.py::test_oracle_matches_committed_shader_digests replays with the oracle.  They keep the
oracle pinned to the shader text on a box that has neither /root/reference nor oracle/_ref.

    python tests/golden/make_ref_shader_golden.py      # in the build container, after `make -C oracle`
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import movement_sim_b200 as M  # noqa: E402  (host helpers only: map + seeded entity init)
from oracle import oracle as O  # noqa: E402
from test_oracle_vs_ref_shader import digest_run  # noqa: E402


def main():
    assert O.ref_shader_available(), "build oracle/_ref first (make -C oracle)"
    test_map = M.Map.load_json(os.path.join(HERE, "test_map.json"))
    small_city = M.Map.city(2200.0, 1600.0, 35.0, 0.3, 0.12, 7)  # tests/conftest.py::small_city
    cases = {"test_map_10k_seed42": (test_map, test_map.init_entities(10_000, seed=42), 1001, 77),
             "small_city_20k_seed9": (small_city, small_city.init_entities(20_000, seed=9), 300, 50)}
    out = {"generated_by": "tests/golden/make_ref_shader_golden.py with oracle/_ref/libref_shader_move.so "
                           "(random_move.comp:5-24,725-750,778-852 compiled as C++, IEEE RNE, no FMA)"}
    for name, (m, ents, passes, every) in cases.items():
        out[name] = {"passes": passes, "every": every, "sha256": digest_run(O, O.ref_shader_move_pass, m, ents, passes, every)}
        print(name, len(out[name]["sha256"]), "digests")
    # the whole shader (main() with quadtree insert / update / collision walk): dispatch ticks 2, 3, 4, ... on a world one metre larger
    # than the map (the shader never terminates for an entity on the map's maximum coordinate)
    import hashlib

    assert O.ref_shader_full_available()
    om = O.OracleMap(small_city.width + 1.0, small_city.height + 1.0, small_city.roads.view(O.ROAD_DTYPE), small_city.connections)
    e = O.np.ascontiguousarray(small_city.init_entities(6000, seed=5)).view(O.ENTITY_DTYPE).copy()
    sim = O.RefShaderSim(e, om, radius=10.0)
    digests = []
    for tick in range(2, 2 + 80):
        sim.dispatch(tick)
        if tick % 8 == 1:
            digests.append(hashlib.sha256(e.tobytes()).hexdigest())
    out["small_city_6k_seed5_full_shader"] = {"dispatches": 80, "every": 8, "radius": 10.0, "world_pad": 1.0, "sha256": digests,
                                              "debug_data": sim.debug.tolist()}
    print("small_city_6k_seed5_full_shader", len(digests), "digests, debugData", sim.debug[:2].tolist())
    with open(os.path.join(HERE, "ref_shader_digests.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
