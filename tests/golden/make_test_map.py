"""Regenerates tests/golden/test_map.json: the reference's 4-road fixture (/root/reference/test_map.json:
four roads from the corners (5|95, 5|95) meeting at (50, 50) in a 100 x 100 m world, five connection
entries per road [own, own, others...]) rebuilt from its defining numbers, in the loader's schema.
Run once: python tests/golden/make_test_map.py"""
import json
import os

corners = [(5, 5), (95, 5), (5, 95), (95, 95)]
others = {0: [1, 2, 3], 1: [0, 2, 3], 2: [1, 0, 3], 3: [1, 2, 0]}
roads, conns = [], []
for r, (x, y) in enumerate(corners):
    base = len(conns)
    conns += [r, r] + others[r]
    roads.append({
        "start": {"lat": 0, "long": 0, "distLat": x, "distLong": y},
        "end": {"lat": 0, "long": 0, "distLat": 50, "distLong": 50},
        "connIndexStart": base, "connCountStart": 1, "connIndexEnd": base + 1, "connCountEnd": 4,
    })
doc = {"minDistLat": 0, "maxDistLat": 100, "minDistLong": 0, "maxDistLong": 100, "roads": roads, "connectionRoadIndexList": conns}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_map.json"), "w") as f:
    json.dump(doc, f)
    f.write("\n")
