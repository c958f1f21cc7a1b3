"""bench.py's CPU-side pieces: the reference arm (--impl reference) prints the contract's JSON line, and the workload
builders of BASELINE configs 3-5 produce what the GPU arm will be given.  No GPU needed."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)


def run_bench(*argv, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *argv], capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"stdout must carry exactly one JSON line, got {len(lines)}"
    return json.loads(lines[0])


def check_reference_line(line, workload):
    assert line["impl"] == "reference" and line["metric"] == "entity-updates/sec" and line["unit"] == "entity-updates/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["config"]["workload"] == workload
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "entity-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_config1():
    line = run_bench("--impl", "reference", "--workload", "test_map_10k_nocollisions", "--steps", "3", "--warmup", "1", "--preroll", "4")
    check_reference_line(line, "test_map_10k_nocollisions")
    assert line["config"]["collisions"] is False and line["config"]["entities_sampled"] == 10_000
    from oracle import oracle as O

    # movement is the reference shader's own code compiled for the CPU when oracle/_ref was built, else the oracle port
    assert line["cpu_baseline"]["kind"] == ("reference" if O.ref_shader_available() else "port")
    assert ("libref_shader_move.so" in line["cpu_baseline"]["sample"]) == O.ref_shader_available()


def test_reference_arm_dense_crowd_sample():
    line = run_bench("--impl", "reference", "--workload", "munich_50m_dense", "--steps", "1", "--warmup", "1", "--preroll", "8", "--ref-sample", "3000")
    check_reference_line(line, "munich_50m_dense")
    assert line["config"]["collisions"] is True and line["config"]["entities_sampled"] == 3000
    assert "central box" in line["config"]["map"]
    from oracle import oracle as O

    if O.ref_shader_full_available():  # the whole compiled shader, one thread, beside the multi-threaded arm
        whole = line["cpu_baseline"]["whole_shader_1_thread"]
        assert whole["cores"] == 1 and whole["value"] > 0


def test_reference_arm_other_ranks_print_nothing():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"], capture_output=True, text=True,
                       timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_builders(msim, monkeypatch):
    import bench

    monkeypatch.setenv("MSIM_BENCH_GRID_SIDE", "33")
    w, m = bench.build_workload(msim, "grid4096_100m_collisions", None)
    assert m.roads.shape[0] == 2 * 33 * 32 and m.width == 640.0 and m.height == 640.0
    assert w["entities"] == 100_000_000 and w["collisions"] and w["box"] is None and w["survey_bytes"] == 140.0
    # lattice coordinates are multiples of 20 m: exact in binary32
    assert np.all(np.mod(m.roads["start_pos"], 20.0) == 0)

    w, m = bench.build_workload(msim, "munich_50m_dense", 5000)
    assert w["entities"] == 5000 and w["survey_bytes"] == 124.0
    x0, y0, x1, y1 = w["box"].tolist()
    assert abs((x1 - x0) / m.width - 0.1) < 1e-3 and abs((y1 - y0) / m.height - 0.1) < 1e-3
    ents = m.init_entities(5000, seed=42, box=w["box"])
    for end in ("start_pos", "end_pos"):  # every chosen road lies inside the central box with both ends
        p = m.roads[end][ents["road_index"]]
        assert np.all((p[:, 0] >= x0) & (p[:, 0] <= x1) & (p[:, 1] >= y0) & (p[:, 1] <= y1))
    assert len(np.unique(ents["road_index"])) > 1000  # ~1 % of 701 590 roads are eligible

    w, m = bench.build_workload(msim, "munich_10m_collisions", None)
    assert w["entities"] == 10_000_000 and w["box"] is None and w["survey_bytes"] == 124.0


def test_reference_arm_runs_the_gpu_arms_configuration():
    """same population, same pre-roll, structured same_config / sampled-count fields; a budget that one step crosses stops the timing"""
    line = run_bench("--impl", "reference", "--workload", "munich_10m_collisions", "--entities", "4000", "--steps", "5", "--warmup", "3", "--preroll", "4",
                     "--ref-budget-s", "0", "--no-whole-shader")
    check_reference_line(line, "munich_10m_collisions")
    c = line["config"]
    assert c["entities"] == 4000 and c["entities_sampled"] == 4000 and c["same_config"] is True
    assert c["steps_asked"] == 5 and c["steps_timed"] == 1 and line["steps"] == 1
    assert c["move_passes_done"] == 4 + 1 + 3 + 1 and c["entity_seed"] == 42 and c["preroll_move_passes"] == 4
    assert isinstance(c["flagged_last_tick"], int) or c["flagged_last_tick"] is None


def test_input_preparation_never_maps_the_cuda_library():
    """bench.py --impl reference prepares its inputs through libmsim_host.so (the reference arm loaded libmsim_cuda.so)"""
    code = ("import sys; sys.path.insert(0, %r); import bench, movement_sim_b200 as M; w, m = bench.build_workload(M, 'munich_10m_collisions', 2000); "
            "e = bench.build_population(M, m, 2000, w['box']); maps = open('/proc/self/maps').read(); "
            "assert 'libmsim_host.so' in maps and 'libmsim_cuda.so' not in maps, maps; print(e.shape[0])") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == "2000"


def test_population_is_the_sharded_generators(msim):
    """N = 1 and N > 1 bench the same draws: bench.build_population == concatenated sharding.generate_population chunks"""
    import bench
    from movement_sim_b200 import sharding

    m = msim.Map.load_json(os.path.join(ROOT, "tests", "golden", "test_map.json"))
    old = sharding.CHUNK
    sharding.CHUNK = 700
    try:
        a = bench.build_population(msim, m, 2000)
        b = np.concatenate([e for _, e in sharding.generate_population(msim, m, 2000, 42)])
    finally:
        sharding.CHUNK = old
    assert a.shape[0] == 2000 and a.tobytes() == b.tobytes()


def test_resort_alignment_matches_the_golden_generators_formula():
    """align_resort_phase leaves the NEXT collision pass a re-sorting one, and tests/golden/make_bench_counts.py predicts the tick index"""
    import importlib.util

    import bench

    spec = importlib.util.spec_from_file_location("make_bench_counts", os.path.join(ROOT, "tests", "golden", "make_bench_counts.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)

    class Fake:  # api.cu's cadence: since_reorder = every at upload; a collision pass re-sorts when since_reorder reaches `every`
        def __init__(self):
            self.since, self.reorders, self.moves = bench.RESORT_EVERY, 0, 0

        def tick(self):
            self.moves += 1
            self.since += 1
            if self.since >= bench.RESORT_EVERY:
                self.reorders += 1
                self.since = 0

        def stats(self):
            return {"reorders": self.reorders}

    for warmup in (0, 3, 5, 10, 31, 40):
        for steps in (1, 20, 33, 200):
            f = Fake()
            f.since = bench.RESORT_EVERY - 1  # (the first collision pass behind an upload re-sorts)
            preroll = 7
            f.moves = preroll
            f.tick()
            for _ in range(max(3, warmup)):
                f.tick()
            bench.align_resort_phase(f.stats, f.tick)
            before = f.reorders
            f.tick()
            assert f.reorders == before + 1, "the first timed tick must re-sort"
            for _ in range(steps - 1):
                f.tick()
            assert f.reorders - before == 1 + (steps - 1) // bench.RESORT_EVERY
            assert f.moves == gen.passes_done(preroll, warmup, steps)


def test_check_counts(tmp_path, monkeypatch):
    import bench

    p = tmp_path / "counts.json"
    p.write_text(json.dumps({"w": {"100": {"340": {"pairs": 7, "flagged": 5}}}}))
    monkeypatch.setattr(bench, "COUNTS_PATH", str(p))
    assert bench.check_counts("w", 100, 340, 7, 5)["status"] == "ok"
    assert bench.check_counts("w", 100, 340, 8, 5)["status"] == "mismatch"
    assert bench.check_counts("w", 100, 341, 7, 5)["status"] == "no stored value"
    assert bench.check_counts("w", 100, 340, None, None)["status"].startswith("not applicable")
    stored = json.load(open(os.path.join(ROOT, "tests", "golden", "bench_counts.json")))
    assert "340" in stored["munich_10m_collisions"]["10000000"]  # the driver's --steps 20 --warmup 5
