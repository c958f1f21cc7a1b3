"""Map pipeline (include/msim_mapgen.h, csrc/host_mapgen.cpp) against golden vectors produced by the reference's own
generator script /root/reference/map/generate_map.py (tests/golden/make_mapgen_golden.py ran it; fixtures are
committed because the GPU box has no reference checkout), plus the binary map cache.  Host only: no GPU.

Parity bar: road geometry (binary32 distLat/distLong of both ends), orientation, road order (= the index the script
assigns), connection index / count of both ends and the min/max bounds are exact.  The order of the roads INSIDE one
coordinate's block of the connection table is Python set order in the script (arbitrary), so block contents are
compared as multisets."""
import json
import os
import struct

import numpy as np
import pytest

from conftest import ROOT

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["small", "city_a", "city_b"]


def load_case(name):
    with open(os.path.join(GOLDEN, f"mapgen_{name}.json")) as f:
        return json.load(f)


def write_ordered_geojson(gold, path):
    """One two-point LineString per road piece, in the order the script's traversal saw them."""
    feats = [{"type": "Feature", "properties": {}, "geometry": {"type": "LineString", "coordinates": [[p[0], p[1]], [p[2], p[3]]]}}
             for p in gold["pieces"]]
    with open(path, "w") as f:
        json.dump({"type": "FeatureCollection", "features": feats}, f)


def blocks_of(roads, connections, duplicate_end=True):
    """{(block start, count): sorted entries} for every coordinate block referenced by a road end."""
    out = {}
    for r in roads:
        for idx, cnt in ((r["connIndexStart"], r["connCountStart"]), (r["connIndexEnd"], r["connCountEnd"])):
            out[(idx, cnt)] = None
    starts = sorted(k[0] for k in out)
    ends = starts[1:] + [len(connections)]
    by_start = dict(zip(starts, ends))
    return {k: sorted(connections[k[0]:by_start[k[0]]]) for k in out}


def as_dicts(m):
    return [{"connIndexStart": int(r["start_index"]), "connCountStart": int(r["start_count"]), "connIndexEnd": int(r["end_index"]),
             "connCountEnd": int(r["end_count"])} for r in m.roads]


@pytest.mark.parametrize("name", CASES)
def test_generator_matches_reference_script(msim, tmp_path, name):
    gold = load_case(name)
    path = str(tmp_path / "ordered.geojson")
    write_ordered_geojson(gold, path)
    m, st = msim.Map.from_geojson(path, msim.MAPGEN_EXACT_TRAVERSAL, with_stats=True)
    want = gold["roads"]
    assert m.roads.shape[0] == len(want) == st["connected"]
    # geometry, orientation and order: binary32(distLat), binary32(distLong) of both ends (Map.cpp:121-122)
    want_start = np.array([[r["start"][2], r["start"][3]] for r in want], dtype=np.float64).astype(np.float32)
    want_end = np.array([[r["end"][2], r["end"][3]] for r in want], dtype=np.float64).astype(np.float32)
    assert np.array_equal(m.roads["start_pos"], want_start)
    assert np.array_equal(m.roads["end_pos"], want_end)
    # the distances themselves in binary64 (haversine restatement): bounds are the max / min over them
    assert [st["min_dist_lat"], st["max_dist_lat"], st["min_dist_long"], st["max_dist_long"]] == gold["bounds"]
    assert m.width == float(np.float32(gold["bounds"][1])) and m.height == float(np.float32(gold["bounds"][3]))
    # connection table: same block starts / counts per road end, same length, block contents as multisets
    got = as_dicts(m)
    for g, w in zip(got, want):
        assert g == {k: w[k] for k in g}
    assert m.connections.shape[0] == len(gold["connections"])
    assert blocks_of(got, m.connections.tolist()) == blocks_of(want, gold["connections"])
    # a road appears once in the block of its start and twice in the block of its end (generate_map.py:244-258)
    for i, r in enumerate(got):
        s = m.connections[r["connIndexStart"]:].tolist()
        assert i in s


def test_small_case_statistics(msim, tmp_path):
    gold = load_case("small")
    path = str(tmp_path / "small.geojson")
    with open(path, "w") as f:
        json.dump({"type": "FeatureCollection", "features": gold["features"]}, f)
    m, st = msim.Map.from_geojson(path, with_stats=True)
    assert st["features"] == 8 and st["line_strings"] == 6
    assert st["road_pieces"] == 10 and st["skipped_zero"] == 1 and st["skipped_duplicate"] == 0
    # file order starts on the ring: everything but the two island pieces hangs together
    assert st["connected"] == 7 and m.roads.shape[0] == 7
    assert st["ref_lat"] == 11.5 and st["ref_long"] == 48.1
    # every road index in the table is valid and every block is as long as count + number of END-matching roads
    assert m.connections.max() < 7
    no_dup = msim.Map.from_geojson(path, msim.MAPGEN_NO_DUPLICATE_END)
    assert no_dup.connections.shape[0] == 2 * 7 and m.connections.shape[0] == 3 * 7
    assert np.array_equal(no_dup.roads["start_pos"], m.roads["start_pos"])


@pytest.mark.parametrize("name", ["city_a", "city_b"])
def test_plain_traversal_keeps_a_superset(msim, tmp_path, name):
    """Default traversal = the script's without the roads it loses to delete-while-iterating: same start road, same
    orientation rule; every road the exact emulation keeps is kept, in both modes every kept road is connected."""
    gold = load_case(name)
    path = str(tmp_path / "ordered.geojson")
    write_ordered_geojson(gold, path)
    exact = msim.Map.from_geojson(path, msim.MAPGEN_EXACT_TRAVERSAL)
    plain = msim.Map.from_geojson(path)

    def undirected(m):
        return {frozenset((tuple(a), tuple(b))) for a, b in zip(m.roads["start_pos"].tolist(), m.roads["end_pos"].tolist())}

    assert undirected(exact) <= undirected(plain)
    assert plain.roads.shape[0] <= len(gold["pieces"])
    # first road identical (the traversal's seed), connection indices in range
    assert np.array_equal(plain.roads[0:1]["start_pos"], exact.roads[0:1]["start_pos"])
    assert plain.connections.max() < plain.roads.shape[0]


def test_haversine_restatement(msim):
    L = msim.lib()
    # along a meridian the great-circle distance is R * dlat: one degree = 111.195 km with R = 6371.0088 km
    assert abs(L.msim_haversine_m(0.0, 0.0, 1.0, 0.0) - 6371008.8 * np.pi / 180.0) < 1e-6
    assert L.msim_haversine_m(48.1, 11.5, 48.1, 11.5) == 0.0
    # Lyon - Paris, the package's README example: 392.2172595594006 km
    assert abs(L.msim_haversine_m(45.7597, 4.8422, 48.8567, 2.3508) - 392217.2595594006) < 1e-6


def test_geojson_errors(msim, tmp_path):
    with pytest.raises(msim.MsimError) as ei:
        msim.Map.from_geojson(str(tmp_path / "missing.geojson"))
    assert ei.value.status == msim.MSIM_ERR_IO and "File does not exist" in ei.value.message
    bad = tmp_path / "bad.geojson"
    bad.write_text('{"type": "FeatureCollection"}')
    with pytest.raises(msim.MsimError) as ei:
        msim.Map.from_geojson(str(bad))
    assert ei.value.status == msim.MSIM_ERR_PARSE and "'features' field missing" in ei.value.message
    bad.write_text('{"features": [{"type": "Feature", "geometry": {"type": "Point", "coordinates": [1, 2]}}]}')
    with pytest.raises(msim.MsimError) as ei:
        msim.Map.from_geojson(str(bad))
    assert ei.value.status == msim.MSIM_ERR_PARSE and "no road" in ei.value.message
    bad.write_text('{"features": [{"geometry": {"type": "LineString", "coordinates": [[1, 2], [3')
    with pytest.raises(msim.MsimError) as ei:
        msim.Map.from_geojson(str(bad))
    assert ei.value.status == msim.MSIM_ERR_PARSE
    with pytest.raises(msim.MsimError) as ei:
        msim.Map.from_geojson(str(bad), flags=1 << 9)
    assert ei.value.status == msim.MSIM_ERR_INVALID


def same_map(a, b):
    return (a.width == b.width and a.height == b.height and a.roads.tobytes() == b.roads.tobytes()
            and a.connections.tobytes() == b.connections.tobytes())


def test_binary_cache_round_trip(msim, tmp_path, test_map):
    city = msim.Map.city(1500.0, 1000.0, 35.0, 0.3, 0.12, 3)
    for m in (test_map, city):
        path = str(tmp_path / "map.msimmap")
        m.save_binary(path)
        assert os.path.getsize(path) == 40 + 32 * m.roads.shape[0] + 4 * m.connections.shape[0] + 8
        assert same_map(msim.Map.load(path), m)            # detected by magic
    # msim_map_load falls through to the reference's JSON schema for everything else, and to GeoJSON by suffix
    assert same_map(msim.Map.load(os.path.join(GOLDEN, "test_map.json")), msim.Map.load_json(os.path.join(GOLDEN, "test_map.json")))
    gold = load_case("small")
    gj = str(tmp_path / "small.geojson")
    with open(gj, "w") as f:
        json.dump({"type": "FeatureCollection", "features": gold["features"]}, f)
    assert same_map(msim.Map.load(gj), msim.Map.from_geojson(gj))


def test_binary_cache_rejects_damage(msim, tmp_path, test_map):
    path = str(tmp_path / "map.msimmap")
    test_map.save_binary(path)
    blob = bytearray(open(path, "rb").read())

    def expect(status, text, data):
        p = str(tmp_path / "damaged.msimmap")
        with open(p, "wb") as f:
            f.write(data)
        h = __import__("ctypes").c_void_p()
        rc = msim.lib().msim_map_load_binary(p.encode(), __import__("ctypes").byref(h))
        assert rc == status and text in msim.lib().msim_map_last_error().decode()

    flipped = bytearray(blob)
    flipped[60] ^= 0x40  # inside the road records
    expect(msim.MSIM_ERR_PARSE, "checksum mismatch", flipped)
    expect(msim.MSIM_ERR_PARSE, "truncated", blob[:-9])
    expect(msim.MSIM_ERR_PARSE, "bad magic", b"NOTAMAP!" + blob[8:])
    version2 = bytearray(blob)
    version2[8:12] = struct.pack("<I", 2)
    expect(msim.MSIM_ERR_UNSUPPORTED, "version 2", version2)
    huge = bytearray(blob)
    huge[24:32] = struct.pack("<Q", 1 << 40)  # road count far beyond the file size: rejected before any allocation
    expect(msim.MSIM_ERR_PARSE, "truncated", huge)


def test_headless_runner_loads_a_map_cache(msim, tmp_path):
    """The drop-in sim::Map::load_from_file goes through msim_map_load: the runner accepts a binary cache.  Without a
    GPU it then stops at msim_create (no CPU fallback), which is as far as this box can go."""
    import subprocess

    import torch

    runner = os.path.join(ROOT, "movement-sim_b200", "msim_headless")
    if not os.path.exists(runner):
        pytest.skip("msim_headless has not been built")
    city = msim.Map.city(1500.0, 1000.0, 35.0, 0.3, 0.12, 3)
    path = str(tmp_path / "city.msimmap")
    city.save_binary(path)
    r = subprocess.run([runner, "--headless", "--map", path, "--entities", "64", "--ticks", "2"], capture_output=True, text=True, timeout=120)
    assert f"Found {city.roads.shape[0]} roads with {city.connections.shape[0]} connections" in r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert r.returncode != 0 and "no CPU fallback" in r.stdout + r.stderr
    else:
        assert r.returncode == 0
