"""Dry run of bench.py's GPU arm (run_b200) on the CPU: torch.cuda and the library's Simulation are replaced by stand-ins (the
Simulation computes with the oracle, events return a made-up clock), so that every statement that assembles the JSON line - including the
blocks added after the last GPU run (colours-only extra, experiment flags, CPU baseline with the compiled shader, forked whole-shader
timing, optional pipelined e2e) - executes once before a GPU box has to.  The numbers mean nothing; the keys and the control flow do."""
import ctypes
import json
import os
import sys
import types

import numpy as np
import pytest

from conftest import ROOT, oracle_map, to_oracle_entities

sys.path.insert(0, ROOT)


class FakeEvent:
    clock = 0.0

    def __init__(self, enable_timing=True):
        self.t = None

    def record(self, stream=None):
        FakeEvent.clock += 0.37
        self.t = FakeEvent.clock

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return other.t - self.t


class FakeTensor:
    def __init__(self, n, itemsize=1):
        self.a = np.zeros(n * itemsize, dtype=np.uint8)

    def data_ptr(self):
        return self.a.ctypes.data

    def numpy(self):
        return self.a

    def fill_(self, v):
        return self


def fake_torch():
    t = types.ModuleType("torch")
    t.uint8, t.float32 = "uint8", "float32"
    t.empty = lambda n, dtype=None, pin_memory=False, device=None: FakeTensor(n, 4 if dtype == "float32" else 1)
    cuda = types.SimpleNamespace()
    cuda.set_device = lambda d: None
    cuda.Stream = lambda: types.SimpleNamespace(cuda_stream=0)
    cuda.Event = FakeEvent
    cuda.synchronize = lambda: None
    cuda.get_device_properties = lambda d: types.SimpleNamespace(uuid=None)

    class stream_ctx:
        def __init__(self, s):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    cuda.stream = stream_ctx
    t.cuda = cuda
    return t


def make_fake_simulation(M, O):
    class Sim:
        def __init__(self, m, entities, radius=10.0, device=0, flags=0, stream=None, **_):
            self.om = oracle_map(O, m)
            self.e = to_oracle_entities(O, entities)
            self.radius, self.flags, self.count = float(radius), flags, self.e.shape[0]
            self.launches, self.pairs, self.profile = 0, 0, None
            self.snap = None
            self.moves = self.collides = self.reorders = 0
            self.since = 32  # api.cu: the first collision pass behind an upload re-sorts, then every 32nd

        def _tick(self, collide):
            O.move_pass(self.e, self.om, threads=4)
            self.moves += 1
            self.launches += 2
            if collide:
                self.collides += 1
                self.since += 1
                if self.since >= 32:
                    self.reorders, self.since = self.reorders + 1, 0
            if self.profile is not None:
                for k in ("move", "arrive"):
                    c, ms = self.profile.get(k, (0, 0.0))
                    self.profile[k] = (c + 1, ms + 0.05)
            if collide:
                self.pairs = O.collide_pass(self.e, self.om.world_w, self.om.world_h, self.radius, threads=4)
                self.launches += 5
                if self.profile is not None:
                    for k, cost in (("cell_scan", 0.03), ("cell_scatter", 0.1), ("query", 0.25)):
                        c, ms = self.profile.get(k, (0, 0.0))
                        self.profile[k] = (c + 1, ms + cost)

        def dispatch(self, tick):
            if not self.e["initialized"].all() or tick % 2 == 0:
                O.move_pass(self.e, self.om, threads=4)
            else:
                self.pairs = O.collide_pass(self.e, self.om.world_w, self.om.world_h, self.radius, threads=4)

        def enqueue_ticks(self, k, collide):
            for _ in range(k):
                self._tick(collide)

        def join(self):
            pass

        def sync(self):
            pass

        def stats(self):
            return {"kernel_launches": self.launches, "last_pair_count": self.pairs, "last_flagged_count": int(O.collision_flags(self.e).sum()),
                    "entity_count": self.count, "grid_cells_x": 1, "grid_cells_y": 1, "move_passes": self.moves, "collide_passes": self.collides,
                    "reorders": self.reorders}

        def profile_begin(self):
            self.profile = {}

        def profile_end(self):
            out, self.profile = self.profile, None
            return out

        def read_entities_ptr(self, ptr, n):
            ctypes.memmove(ptr, self.e.ctypes.data, n * 64)

        def upload_ptr(self, ptr, n):
            ctypes.memmove(self.e.ctypes.data, ptr, n * 64)

        def read_positions_ptr(self, ptr, n):
            ctypes.memmove(ptr, np.ascontiguousarray(self.e["pos"][:n]).ctypes.data, n * 8)

        def read_collision_flags_ptr(self, ptr, n):
            ctypes.memmove(ptr, np.ascontiguousarray(O.collision_flags(self.e)[:n]).ctypes.data, n)

        def snapshot_begin(self):
            self.snap = self.e.copy()

        def snapshot_end(self, copy=True):
            return self.snap

        def close(self):
            pass

    return Sim


@pytest.mark.timeout(1800)
@pytest.mark.parametrize("backend,argv", [
    ("stand-in", ["--workload", "munich_10m_collisions", "--entities", "6000"]),
    ("stand-in", ["--workload", "munich_1m_nocollisions", "--entities", "5000"]),
    ("stand-in", ["--workload", "munich_10m_collisions", "--entities", "4000", "--no-e2e-variants", "--no-flags-only"]),
    # against the real api.cu + kernels under the emulator: the bench's map has 4.8 M grid cells, i.e. 2 x 1165 scan CTAs per tick, which costs
    # the emulator minutes - all three configurations passed that way by hand; in the suite only with MSIM_TEST_SLOW=1
    pytest.param("emulated library", ["--workload", "munich_10m_collisions", "--entities", "1200"],
                 marks=pytest.mark.skipif(os.environ.get("MSIM_TEST_SLOW") != "1", reason="minutes under the emulator: set MSIM_TEST_SLOW=1")),
])
def test_gpu_arm_assembles_its_json_line(msim, orc, monkeypatch, capfd, argv, backend):
    import bench

    monkeypatch.setitem(sys.modules, "torch", fake_torch())
    if backend == "stand-in":
        monkeypatch.setattr(msim, "Simulation", make_fake_simulation(msim, orc))
    else:  # the real api.cu and kernels under the host SIMT emulator (tests/cuda_emu): bench's calls meet the code they will meet on a GPU
        from conftest import load_library_under_emulator

        monkeypatch.setitem(sys.modules, "movement_sim_b200", load_library_under_emulator())
    monkeypatch.setattr(bench, "ClockSampler", lambda uuid: types.SimpleNamespace(stop=lambda: {"sm_mhz": None, "reasons": ["dry run"]}))
    monkeypatch.setattr(bench, "_RESULT_FD", None)
    monkeypatch.delenv("MSIM_BENCH_RESULT_FD", raising=False)
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "3", "--warmup", "3", "--preroll", "4", "--cpu-sample", "1000", "--cpu-steps", "1", *argv])
    monkeypatch.setattr(bench, "quiet_stdout", lambda: None)
    monkeypatch.setattr(bench.time, "time", iter(range(10_000)).__next__)  # the "keep the GPU busy for 1.2 s" loop ends at once
    assert bench.main() == 0
    out = capfd.readouterr().out.strip().splitlines()
    line = json.loads(out[-1])
    collisions = "nocollisions" not in argv[1]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "roofline", "tick", "kernels", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["metric"] == "entity-updates/sec" and line["n_gpus"] == 1 and line["steps"] == 3 and line["vs_baseline"] is None
    assert line["config"]["workload"] == argv[1]
    assert line["roofline"]["bound"] == "hbm"
    if backend == "stand-in":
        assert line["roofline"]["kernel"] == ("query" if collisions else "move")
    else:  # real per-kernel names from msim_profile_end (times are the emulated runtime's constant)
        names = {k["name"] for k in line["kernels"]}
        assert "move" in names and (("query" in names and "cell_scatter" in names) if collisions else "query" not in names)
        assert line["gpu_launches"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == line["config"]["entities"] * 64 == line["e2e"]["d2h_bytes_per_step"]
    variants = line["e2e"]["variants"]
    assert ("pipelined_full_aos" in variants) == ("--no-e2e-variants" not in argv)
    if "--no-e2e-variants" not in argv:
        assert variants["full_aos_up_render_state_down"]["d2h_bytes_per_step"] == line["config"]["entities"] * (9 if collisions else 8)
        assert line["e2e"]["variant"] in ("blocking_full_aos", "pipelined_full_aos")
    c = line["config"]
    assert c["move_passes_done"] > 0 and c["counts_check"]["status"] in ("no stored value", "not applicable (collisions off)")
    if collisions:
        assert c["pairs_last_tick"] is not None and c["flagged_last_tick"] is not None and c["resort"]["in_timed_region"] == 1
    assert (line["move_only"] is not None) == collisions
    if collisions and "--no-flags-only" not in argv:
        assert line["flags_only"]["ms_per_step"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] > 0 and cb["port_value"] > 0
    if collisions and orc.ref_shader_full_available():
        assert cb["whole_shader_1_thread"]["value"] > 0


@pytest.mark.timeout(600)
@pytest.mark.parametrize("argv", [["--workload", "munich_10m_collisions", "--entities", "5000"],
                                  ["--workload", "munich_50m_dense", "--entities", "3000"],
                                  ["--workload", "munich_1m_nocollisions", "--entities", "4000"],
                                  ["--workload", "munich_10m_collisions", "--entities", "2500", "--scaling", "weak"]])
def test_multi_gpu_arm_assembles_its_json_line(tmp_path, argv):
    """bench.py --gpus 2 (sharding.bench_main) as two gloo ranks on the CPU: stand-ins for torch.cuda and the CUDA engine (the
    oracle-backed engine of the sharding tests), everything else - workload builder, partition histogram, population build on the rank's
    host threads, orchestration, collectives, JSON assembly - is the code the GPU run executes (tests/bench_shard_dry_worker.py)."""
    import socket
    import subprocess

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    full = ["--steps", "3", "--warmup", "3", "--preroll", "3", "--e2e-steps", "1", *argv]
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "bench_shard_dry_worker.py"), str(r), "2", str(port),
                               str(tmp_path / f"out{r}.json"), json.dumps(full)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=500) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-3000:]
    r0 = json.load(open(tmp_path / "out0.json"))
    r1 = json.load(open(tmp_path / "out1.json"))
    assert r0["rc"] == 0 and r1["rc"] == 0 and len(r0["lines"]) == 1 and r1["lines"] == []  # rank 0 alone prints
    line = r0["lines"][0]
    collisions = "nocollisions" not in argv[1]
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "roofline", "tick", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    weak = "weak" in argv
    assert line["n_gpus"] == 2 and line["scaling"] == ("weak" if weak else "strong") and line["config"]["workload"] == argv[1]
    assert sum(line["config"]["owned_per_rank"]) == line["config"]["entities_total"] == int(argv[3]) * (2 if weak else 1)
    assert line["tick"]["survey_bytes_per_entity_update"] == (124.0 if collisions else 24.0)
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    if collisions:
        assert line["config"]["exchange"] == "collective" and line["config"]["pairs_last_tick"] is not None
        assert line["config"]["flagged_last_tick"] is not None and line["config"]["move_passes_done"] > 0
        assert line["config"]["counts_check"]["status"] == "no stored value"
        assert line["roofline"]["kernel"] == "query"
    if "dense" in argv[1]:
        assert "central box" in line["config"]["map"]
