#!/usr/bin/env python
"""Golden vectors for the native map generator (movement-sim_b200/csrc/host_mapgen.cpp), produced by running the
reference's own script /root/reference/map/generate_map.py on small synthetic GeoJSON inputs.

Run in the build container (the GPU box has no /root/reference):   python tests/golden/make_mapgen_golden.py

How the script is run unmodified:
  * it opens "map/munich.geojson" relative to the working directory and writes "munich.json" there -> a scratch
    directory holds both;
  * it imports two PyPI packages that are not installed here.  Stand-ins are put on sys.path for the run:
    `geojson.load` = json.load, and `haversine.haversine` = the published formula of the `haversine` package
    (2 * 6371.0088 km * asin(sqrt(sin^2(dlat/2) + cos(lat1) cos(lat2) sin^2(dlng/2))), degrees in, km out).  The
    distances in the fixtures therefore pin the native code to that formula evaluated by CPython, not to the
    package binary.
What is recorded per case (tests/golden/mapgen_<case>.json):
  * "pieces": the road pieces in the order the script's `list(map.roads)` presented them to remove_not_connected
    (Python set order, recovered by calling the script's own build_map() a second time in the same process);
    the native generator takes its input in exactly this order, so MSIM_MAPGEN_EXACT_TRAVERSAL has to reproduce
    the script's traversal including the roads it loses by deleting from the list it iterates;
  * "roads": the script's Road objects sorted by the `index` build_road_connections gave them;
  * "connections": connectionRoadIndexList; "bounds": min/max distances; "ref": the reference point.
"""
import json
import math
import os
import random
import runpy
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SCRIPT = "/root/reference/map/generate_map.py"


def haversine_standin(p1, p2):
    lat1, lng1 = p1
    lat2, lng2 = p2
    lat1, lng1, lat2, lng2 = map(math.radians, (lat1, lng1, lat2, lng2))
    lat = lat2 - lat1
    lng = lng2 - lng1
    d = math.sin(lat * 0.5) ** 2 + math.cos(lat1) * math.cos(lat2) * math.sin(lng * 0.5) ** 2
    return 2 * 6371.0088 * math.asin(math.sqrt(d))


def run_reference(features):
    geo = types.ModuleType("geojson")
    geo.load = json.load
    hav = types.ModuleType("haversine")
    hav.haversine = haversine_standin
    saved = {k: sys.modules.get(k) for k in ("geojson", "haversine")}
    sys.modules["geojson"], sys.modules["haversine"] = geo, hav
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "map"))
        with open(os.path.join(tmp, "map", "munich.geojson"), "w") as f:
            json.dump({"type": "FeatureCollection", "features": features}, f)
        os.chdir(tmp)
        try:
            out = sys.stdout
            sys.stdout = open(os.devnull, "w")
            try:
                g = runpy.run_path(SCRIPT)
            finally:
                sys.stdout.close()
                sys.stdout = out
        finally:
            os.chdir(cwd)
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
    result = g["map"]
    # the order remove_not_connected saw: same elements inserted in the same order into a fresh set, same process
    again = g["build_map"](features)
    order = [[r.start.lat, r.start.long, r.end.lat, r.end.long] for r in list(again.roads)]
    roads = sorted(result.roads, key=lambda r: r.index)
    assert [r.index for r in roads] == list(range(len(roads)))
    return {
        "pieces": order,
        "roads": [
            {
                "start": [r.start.lat, r.start.long, r.start.distLat, r.start.distLong],
                "end": [r.end.lat, r.end.long, r.end.distLat, r.end.distLong],
                "connIndexStart": r.connIndexStart, "connCountStart": r.connCountStart,
                "connIndexEnd": r.connIndexEnd, "connCountEnd": r.connCountEnd,
            }
            for r in roads
        ],
        "connections": list(result.connectionRoadIndexList),
        "bounds": [result.minDistLat, result.maxDistLat, result.minDistLong, result.maxDistLong],
    }


def line(points, **props):
    return {"type": "Feature", "properties": props, "geometry": {"type": "LineString", "coordinates": [list(p) for p in points]}}


def case_small():
    """Hand-made: a block of streets with a multi-point way, a dead end, a way drawn towards the junction (gets flipped),
    an island that is not connected, a zero-length piece, a one-point way, and non-LineString features."""
    a, b, c, d = (11.50, 48.10), (11.51, 48.10), (11.51, 48.11), (11.50, 48.11)
    e, f = (11.52, 48.10), (11.505, 48.105)
    return [
        line([a, b, c, d, a], name="ring"),
        line([e, b], name="drawn towards the junction"),
        line([b, f], name="dead end"),
        line([f, f, c], name="holds a zero-length piece"),
        line([(11.60, 48.20), (11.61, 48.20), (11.61, 48.21)], name="island"),
        line([(11.7, 48.3)], name="one point"),
        {"type": "Feature", "properties": {}, "geometry": {"type": "Point", "coordinates": [11.5, 48.1]}},
        {"type": "Feature", "properties": {}, "geometry": {"type": "Polygon", "coordinates": [[list(a), list(b), list(c), list(a)]]}},
    ]


def case_city(seed, nx, ny, ways):
    """Random walks over a jittered lattice of junctions: many ways share junctions, so one traversal step finds several
    roads at once and the script's delete-while-iterating loop loses some of them."""
    rnd = random.Random(seed)
    node = {(i, j): (round(11.4 + 0.002 * i + rnd.uniform(-4e-4, 4e-4), 7), round(48.0 + 0.0015 * j + rnd.uniform(-3e-4, 3e-4), 7))
            for i in range(nx) for j in range(ny)}
    used = set()
    feats = []
    for _ in range(ways):
        i, j = rnd.randrange(nx), rnd.randrange(ny)
        pts = [node[(i, j)]]
        for _ in range(rnd.randrange(1, 7)):
            di, dj = rnd.choice([(1, 0), (-1, 0), (0, 1), (0, -1)])
            ni, nj = i + di, j + dj
            if not (0 <= ni < nx and 0 <= nj < ny):
                continue
            piece = (node[(i, j)], node[(ni, nj)])
            # the script asserts that no piece occurs twice (:180) and, after re-orienting, that no road equals one
            # already registered at its start (:200): a street drawn once in each direction aborts it
            if piece in used or piece[::-1] in used:
                break
            used.add(piece)
            pts.append(node[(ni, nj)])
            i, j = ni, nj
        if len(pts) >= 2:
            feats.append(line(pts))
    return feats


CASES = {
    "small": case_small(),
    "city_a": case_city(7, 9, 7, 70),
    "city_b": case_city(8, 14, 11, 260),
}


def main():
    import subprocess

    if len(sys.argv) == 3:  # child: one case under the PYTHONHASHSEED the parent chose
        name, out_path = sys.argv[1], sys.argv[2]
        gold = run_reference(CASES[name])
        gold["features"] = CASES[name]
        gold["python_hash_seed"] = os.environ.get("PYTHONHASHSEED")
        gold["generated_by"] = "tests/golden/make_mapgen_golden.py running /root/reference/map/generate_map.py"
        with open(out_path, "w") as f:
            json.dump(gold, f)
        return
    # Python set order depends on the process's string-hash seed, and with it the road the script starts from and the
    # roads it loses: every seed gives a different, equally valid vector.  Take the first seed whose traversal keeps a
    # good part of the input (a start road that ends in a dead end keeps one road and tests nothing).
    for name in CASES:
        path = os.path.join(HERE, f"mapgen_{name}.json")
        for seed in range(1, 40):
            env = dict(os.environ, PYTHONHASHSEED=str(seed))
            subprocess.run([sys.executable, os.path.abspath(__file__), name, path], check=True, env=env)
            gold = json.load(open(path))
            if len(gold["roads"]) * 10 >= len(gold["pieces"]) * 6:
                break
        print(f"{name}: hash seed {seed}: {len(gold['pieces'])} pieces -> {len(gold['roads'])} roads, "
              f"{len(gold['connections'])} connection entries -> {path}")


if __name__ == "__main__":
    main()
