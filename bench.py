#!/usr/bin/env python
"""bench.py — entity-updates/sec of the per-tick entity update on B200, with the HBM roofline of the
dominant kernel and a CPU baseline timed in the same run.  Contract: see DESIGN.md §Measurement.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

A "step" is one sim tick = one move pass + one collision pass over the whole resident population
(Simulator::sim_tick, /root/reference/src/sim/Simulator.cpp:213-241), except for the collisions-off
workloads where it is one move pass.  Default workload: BASELINE.json configs[2], the configuration
its metric is quoted on ("Munich map ... 10M entities ... collisions on ... 1/2/4/8 B200").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (entities per GPU, collisions, map)
    "munich_10m_collisions": dict(entities=10_000_000, collisions=True, map="city"),
    "munich_1m_nocollisions": dict(entities=1_000_000, collisions=False, map="city"),
    "test_map_10k_nocollisions": dict(entities=10_000, collisions=False, map="test_map"),
    # BASELINE configs[3] / configs[4]: sized for --gpus 8 (ONE population split over the GPUs; they also fit one B200)
    "grid4096_100m_collisions": dict(entities=100_000_000, collisions=True, map="grid4096", survey_bytes=140.0),  # 26-bit keys: k = 4 passes
    "munich_50m_dense": dict(entities=50_000_000, collisions=True, map="city", box_frac=(0.45, 0.45, 0.55, 0.55)),
}
METRIC = "entity-updates/sec"
SURVEY_BYTES = {True: 124.0, False: 24.0}  # B_coll (23-bit keys, 3 passes) / B_move
# algorithmic HBM bytes per entity per launch of each kernel
KERNEL_BYTES = {
    "move": 24.0,  # pos R8 + target R8 + pos W8 (the per-cell population leaves as reductions into an L2-resident table; sharded handles add a 4 B key)
    "cell_count": 8.0,  # pos R8 (only when no counting move pass preceded the collision pass)
    "cell_scatter": 20.0,  # pos R8 + sorted pos W8 + slot W4 (the key is recomputed from the position)
    "keygen": 12.0,
    "histogram": 4.0,
    "sort_pass0": 12.0,  # key R4 + pair W8
    "sort_pass1": 16.0,
    "sort_pass2": 16.0,
    "sort_pass3": 16.0,
    "build_cells": 24.0,  # pair R8 + pos gather R8 + sorted pos W8
    "query": 9.0,  # sorted pos R8 + flag W1 (neighbour reads are cache hits)
    "fold_counts": 1.0,  # flag R1: the flagged count is taken from the stored flags (collide.cu fold_counts_kernel)
    "reorder": 98.0,  # every 32nd tick: idx R4 + flag R1 + (prev pos 8, target 8, road 4, rng 16, id 4) gathered and written + flag W1 + slot map W4
}


_RESULT_FD = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to
    fd 1 when NCCL_DEBUG is set), so everything else is sent to stderr and only emit() reaches the real stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)
        os.environ["MSIM_BENCH_RESULT_FD"] = str(_RESULT_FD)  # sharding.py imports this file a second time as module "bench"


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    fd = _RESULT_FD if _RESULT_FD is not None else (int(os.environ["MSIM_BENCH_RESULT_FD"]) if "MSIM_BENCH_RESULT_FD" in os.environ else None)
    if fd is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(fd, data)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(workload: str):
    """dram bytes per launch from the committed ncu --set full capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(workload, {})
    return {}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 100 ms while the GPU is under the bench load."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid: str | None):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        cmd = ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"]
        if uuid:
            cmd += ["-i", uuid]
        try:
            self.proc = subprocess.Popen(cmd, stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower() == "active":
                    reasons.add(name)
        self.tmp.close()
        os.unlink(self.tmp.name)
        # "under load": samples in the upper half of the power range seen
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(reasons), "samples": 0}
        thr = (max(power) + min(power)) / 2 if max(power) > min(power) + 50 else min(power)
        loaded = [c for c, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "samples_under_load": len(loaded), "power_w_max": max(power)}


def build_workload(M, name: str, entities_override: int | None, seed: int = 42):
    w = dict(WORKLOADS[name])
    if entities_override:
        w["entities"] = entities_override
    if w["map"] == "city":
        m = M.Map.city()  # 29007.4609 x 16463.7656 m Munich stand-in, seed 2022
        w["map_desc"] = f"synthetic Munich stand-in {m.width:.1f}x{m.height:.1f} m, {m.roads.shape[0]} roads (munich.json absent from the reference checkout)"
    elif w["map"] == "grid4096":
        side = int(os.environ.get("MSIM_BENCH_GRID_SIDE", "4096"))  # smaller lattices for dry runs
        m = M.Map.grid(side, side, 20.0)  # nodes at (20 i, 20 j): every coordinate exact in binary32
        w["map_desc"] = f"synthetic {side}x{side} lattice, 20 m spacing, {m.width:.0f}x{m.height:.0f} m, {m.roads.shape[0]} roads"
    else:
        m = M.Map.load_json(os.path.join(ROOT, "tests", "golden", "test_map.json"))
        w["map_desc"] = "test_map.json (4 roads)"
    w.setdefault("survey_bytes", SURVEY_BYTES[w["collisions"]])
    w["box"] = None
    if w.get("box_frac"):  # dense crowd: initial roads drawn only from roads with both ends inside the central box
        fx0, fy0, fx1, fy1 = w["box_frac"]
        w["box"] = np.array([fx0 * m.width, fy0 * m.height, fx1 * m.width, fy1 * m.height], dtype=np.float32)
        w["map_desc"] += f"; entities start on roads inside the central box {w['box'].tolist()}"
    return w, m


def build_population(M, m, total: int, box=None, seed: int = 42):
    """The seeded global population every arm runs (1, 2, 4, 8 GPUs and the CPU reference): 1 Mi-entity chunks with seeds seed + 1000 * chunk
    (movement-sim_b200/sharding.py generate_population), so that pair and flag counts can be compared across arms."""
    from movement_sim_b200 import sharding

    threads = max(1, min(8, os.cpu_count() or 1))
    parts = sharding._map_chunks(lambda c: m.init_entities(c[2], seed=seed + 1000 * c[0], box=box), total, threads)
    return np.concatenate(parts) if parts else m.init_entities(0, seed=seed, box=box)


COUNTS_PATH = os.path.join(ROOT, "tests", "golden", "bench_counts.json")


def check_counts(workload: str, entities: int, move_passes: int, pairs, flagged) -> dict:
    """Self-check of a bench line: unique pairs and flagged entities of the last timed tick against the values stored for the same
    population after the same number of move passes (tests/golden/bench_counts.json: measured at N = 1 and reproduced by the CPU
    oracle, tests/golden/make_bench_counts.py).  Every N must report the same numbers; a mismatch fails the run."""
    if pairs is None:
        return {"status": "not applicable (collisions off)"}
    try:
        with open(COUNTS_PATH) as f:
            stored = json.load(f).get(workload, {}).get(str(entities), {}).get(str(move_passes))
    except (OSError, ValueError):
        stored = None
    if not stored:
        return {"status": "no stored value", "key": [workload, entities, move_passes]}
    ok = int(stored["pairs"]) == int(pairs) and int(stored["flagged"]) == int(flagged)
    return {"status": "ok" if ok else "mismatch", "expected": stored, "key": [workload, entities, move_passes]}


RESORT_EVERY = 32  # collision passes between two re-sorts of the resident state into cell order (api.cu reorder_every)


def align_resort_phase(stats, step, limit: int = 2 * RESORT_EVERY):
    """Ticks (untimed) until the NEXT collision pass is one that re-sorts the storage: the timed K steps that follow then hold
    1 + (K - 1) // RESORT_EVERY re-sorts, never fewer than their share K / RESORT_EVERY.  Returns the ticks spent."""
    start = stats()["reorders"]
    spent = 0
    while stats()["reorders"] == start and spent < limit:
        step()
        spent += 1
    if stats()["reorders"] == start:
        return spent  # re-sorting is off (MSIM_FLAG_NO_REORDER, onesweep): nothing to align
    for _ in range(RESORT_EVERY - 1):
        step()
    return spent + RESORT_EVERY - 1


# --------------------------------------------------------------------------------------------------
# CPU: the reference's own CPU path (oracle/_ref quadtree) + the oracle port
# --------------------------------------------------------------------------------------------------
def cpu_step(O, e, omap, radius, threads, collisions, ref_move, ref_tree, large=False):
    """One sim tick on the host with as much of the reference's OWN code as compiles here (oracle/_ref):
    movement = the shader's update_direction / move / new_target / next compiled for the CPU (random_move.comp:725-852,
    libref_shader_move.so) on all host threads by static entity ranges, else the oracle port;
    neighbour structure = the reference's CPU quadtree (shader_validation/src/main.cpp, libref_quadtree.so): rebuilt by
    quad_tree_insert (single-threaded: its multi-threaded insert and its incremental quad_tree_update trip the harness's
    own lock assertions) and queried by quad_tree_check_collisions on all host threads, else the oracle's
    cell grid."""
    if ref_move:
        O.ref_shader_move_pass(e, omap, threads=threads)
    else:
        O.move_pass(e, omap, threads=threads)
    if not collisions:
        return
    if ref_tree:
        q = O.RefQuadTree(omap.world_w, omap.world_h, radius, 10, large=large)
        q.insert(e["pos"], 1)
        flags, _ = q.collide(threads)
        return int(flags.sum())
    else:
        O.collide_pass(e, omap.world_w, omap.world_h, radius, threads=threads)


def time_cpu(O, ents_aos, omap, radius, collisions, steps, warmup, ref_move, ref_tree, threads, large=False, budget_s=None):
    """Returns (entity-updates/s, seconds per step, steps done, flagged entities of the last step or None).  budget_s: stop after the
    step that crosses it (at least one step is always timed)."""
    e = np.ascontiguousarray(ents_aos).view(O.ENTITY_DTYPE).copy()
    e["initialized"] = 1
    for _ in range(warmup):
        cpu_step(O, e, omap, radius, threads, collisions, ref_move, ref_tree, large)
    t0 = time.perf_counter()
    done, flagged = 0, None
    for _ in range(steps):
        flagged = cpu_step(O, e, omap, radius, threads, collisions, ref_move, ref_tree, large)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return e.shape[0] * done / dt, dt / done, done, flagged


def time_whole_shader(O, ents_aos, m, radius, collisions, sample=100_000, ticks=2, limit_s=90.0):
    """The reference's ENTIRE shader compiled for the CPU (oracle/_ref/libref_shader_full.so), one host thread, dispatch by dispatch:
    reported beside the multi-threaded arm.  The world is one metre larger than the map (the shader never terminates for an
    entity on the map's maximum coordinate).  Runs in a forked child under a time limit: the reference's lock protocol is known to leave
    locks behind, and a bench must never hang on the code it is being compared with (the child touches no CUDA state)."""
    if not (collisions and O.ref_shader_full_available()):
        return None
    import select

    # copied BEFORE the fork: the caller's buffer may be CUDA-pinned memory, which a forked child does not inherit
    e = np.ascontiguousarray(ents_aos[:sample]).view(O.ENTITY_DTYPE).copy()
    e["initialized"] = 0

    def work():
        om = O.OracleMap(m.width + 1.0, m.height + 1.0, m.roads.view(O.ROAD_DTYPE), m.connections)
        sim = O.RefShaderSim(e, om, radius=radius)
        try:
            sim.dispatch(2)  # initialise + quad_tree_insert (untimed, like the GPU arm's first dispatch)
            sim.dispatch(3)
            t0 = time.perf_counter()
            for k in range(ticks):
                sim.dispatch(4 + 2 * k)
                sim.dispatch(5 + 2 * k)
            dt = time.perf_counter() - t0
        except O.RefShaderDeadlock as ex:
            return {"error": str(ex)}
        return {"value": e.shape[0] * ticks / dt, "unit": "entity-updates/s", "cores": 1, "sample": f"{e.shape[0]} entities, {ticks} sim ticks",
                "what": "random_move.comp compiled as C++: main() with quad_tree_update + quad_tree_check_collisions, invocations in index order"}

    rd, wr = os.pipe()
    pid = os.fork()
    if pid == 0:  # child
        code = 1
        try:
            os.close(rd)
            os.write(wr, json.dumps(work()).encode())
            code = 0
        finally:
            os._exit(code)
    os.close(wr)
    out = b""
    deadline = time.monotonic() + limit_s
    while True:
        left = deadline - time.monotonic()
        ready, _, _ = select.select([rd], [], [], max(0.0, left))
        if not ready:
            os.kill(pid, 9)
            result = {"error": f"no result within {limit_s:.0f} s"}
            break
        chunk = os.read(rd, 65536)
        if not chunk:
            result = json.loads(out.decode()) if out else {"error": "the child exited without a result"}
            break
        out += chunk
    os.close(rd)
    os.waitpid(pid, 0)
    return result


def cpu_arm_description(ref_move, ref_tree, collisions):
    """kind + wording of the CPU arm: "reference" only when every part of the step is the reference's own code."""
    move = ("movement = the reference shader's own move / new_target / RNG code compiled for the CPU (oracle/_ref/libref_shader_move.so), all threads"
            if ref_move else "movement = oracle port, all threads")
    if not collisions:
        return ("reference" if ref_move else "port"), move
    tree = ("neighbour structure = reference quadtree (oracle/_ref/libref_quadtree.so): insert on 1 thread, collision walk on all threads"
            if ref_tree else "neighbour structure = oracle cell grid, all threads")
    return ("reference" if (ref_move and ref_tree) else "port"), move + "; " + tree


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores, on the GPU arm's configuration:
    the same seeded population at its full size, the same pre-roll.  The reference's collision walk is quadratic in the leaf
    population (depth-8 quadtree, shader_validation/src/main.cpp), so one full-size sim tick costs minutes of CPU time at 10 M
    entities: the run times as many of the K steps as fit --ref-budget-s (at least one) and says how many it did."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import movement_sim_b200 as M  # host-only helpers (libmsim_host.so): this arm never loads the CUDA library
    from oracle import oracle as O

    w, m = build_workload(M, args.workload, args.entities)
    threads = os.cpu_count() or 1
    collisions = w["collisions"]
    total = w["entities"] * (args.gpus if args.scaling == "weak" else 1)
    ref_move = O.ref_shader_available()
    sample = min(total, args.ref_sample) if args.ref_sample else total
    large = collisions and sample > (int(O.ref().ref_capacity()) if O.ref_available() else 0) and O.ref_large_available()
    if large:
        sample = min(sample, int(O.ref(large=True).ref_capacity()))
    ref_tree = collisions and (large or (O.ref_available() and sample <= int(O.ref().ref_capacity())))
    if collisions and not ref_tree and O.ref_available():  # no large build on this box: the largest sample the harness holds
        sample = min(sample, int(O.ref().ref_capacity()))
        ref_tree = True
    omap = O.OracleMap(m.width, m.height, m.roads.view(O.ROAD_DTYPE), m.connections)
    ents = build_population(M, m, total, w["box"])[:sample]
    e = np.ascontiguousarray(ents).view(O.ENTITY_DTYPE).copy()
    del ents
    O.move_pass(e, omap, threads=threads)  # init dispatch
    for _ in range(args.preroll + 1):  # the GPU arm's pre-roll and its one untimed collision tick's move pass
        O.move_pass(e, omap, threads=threads)
    # warm-up: the GPU arm's W warm-up ticks as move passes (same population state when timing starts); the reference rebuilds its
    # tree from nothing every tick, so a warm-up collision pass would only warm caches - at minutes per pass it is left out and says so
    warm_moves = max(3, args.warmup)
    for _ in range(warm_moves):
        O.move_pass(e, omap, threads=threads)
    steps_asked = max(1, args.steps)
    value, sec, steps, flagged = time_cpu(O, e, omap, 10.0, collisions, steps_asked, 0, ref_move, ref_tree, threads, large=large,
                                          budget_s=args.ref_budget_s)
    kind, how = cpu_arm_description(ref_move, ref_tree, collisions)
    if large:
        how = how.replace("libref_quadtree.so", "libref_quadtree_large.so")
    sample_desc = (f"{sample} of {total} entities of the same seeded population, {steps} of {steps_asked} sim ticks timed (budget {args.ref_budget_s:.0f} s, "
                   f"{sec:.1f} s per tick) after {args.preroll} pre-roll + {warm_moves + 1} warm-up move passes; {how}")
    whole = None
    if not args.no_whole_shader:
        try:
            whole = time_whole_shader(O, e, m, 10.0, collisions, limit_s=45.0)
        except Exception as ex:  # an extra beside the baseline: never allowed to cost the line
            whole = {"error": repr(ex)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "entity-updates/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": max(3, args.warmup), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32+u32",
        "data": "synthetic",
        "config": {"workload": args.workload, "entities": total, "entities_sampled": sample, "same_config": sample == total, "collisions": collisions,
                   "collision_radius_m": 10.0, "map": w["map_desc"], "entity_seed": 42, "preroll_move_passes": args.preroll,
                   "steps_asked": steps_asked, "steps_timed": steps, "warmup_as_move_passes": warm_moves,
                   "flagged_last_tick": flagged, "move_passes_done": args.preroll + 1 + warm_moves + steps},
        "cpu_baseline": {"value": value, "unit": "entity-updates/s", "cores": threads, "kind": kind, "sample": sample_desc, "whole_shader_1_thread": whole},
        "e2e": {"value": value, "unit": "entity-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch

    import movement_sim_b200 as M

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one process per GPU)")
    M.lib()  # fail loudly without the CUDA library
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from movement_sim_b200 import sharding  # noqa: F401  (multi-GPU path)

        return sharding.bench_main(args, M, rank, world, local_rank)

    w, m = build_workload(M, args.workload, args.entities)
    n = w["entities"]
    collisions = w["collisions"]
    peak, peak_src = load_peaks()
    stream = torch.cuda.Stream()
    ents = build_population(M, m, n, w["box"])  # the same seeded population at every N (sharding.generate_population)
    if args.presort:  # experiment: spatially coherent storage order (sort the host array by cell row, then x)
        rows, _, _ = M.grid_rows(m.width, m.height, 10.0, ents["pos"])
        order = np.lexsort((ents["pos"][:, 0], rows))
        ents = np.ascontiguousarray(ents[order])
    flags = (0 if collisions else M.FLAG_NO_COLLISIONS) | (M.FLAG_SORT_COUNTING if args.counting_sort else 0) | (M.FLAG_NO_REORDER if args.no_reorder else 0) | (
        M.FLAG_SORT_ONESWEEP if args.onesweep else 0)
    sim = M.Simulation(m, ents, radius=10.0, device=local_rank, flags=flags, stream=stream.cuda_stream)
    sim.dispatch(2)  # the reference's first dispatch: initialise only
    sim.enqueue_ticks(args.preroll, False)  # disperse the population along the roads (untimed)
    sim.enqueue_ticks(1, collisions)
    sim.sync()

    props = torch.cuda.get_device_properties(local_rank)
    uuid = getattr(props, "uuid", None)
    sampler = ClockSampler(f"GPU-{uuid}" if uuid and not str(uuid).startswith("GPU-") else (str(uuid) if uuid else None))

    small = n * (40 if collisions else 24) < 200e6  # hot state would sit in the 126 MB L2
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda") if small else None

    def timed_steps(k):
        """k steps; returns device ms (CUDA events on the launching stream)."""
        if flush_buf is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(k):
                sim.enqueue_ticks(1, collisions)
            sim.join()  # the library's side stream (pass B of the last tick) is inside the timed region too
            e1.record(stream)
            e1.synchronize()
            return e0.elapsed_time(e1)
        total = 0.0
        for _ in range(k):
            with torch.cuda.stream(stream):
                flush_buf.fill_(1)  # evict the working set from L2 between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sim.enqueue_ticks(1, collisions)
            sim.join()
            e1.record(stream)
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total

    timed_steps(max(3, args.warmup))
    aligned = align_resort_phase(sim.stats, lambda: sim.enqueue_ticks(1, collisions)) if collisions else 0
    sim.sync()
    st0 = sim.stats()
    launches0 = st0["kernel_launches"]
    torch.cuda.synchronize()
    ms = timed_steps(args.steps)
    torch.cuda.synchronize()
    sim.sync()
    st1 = sim.stats()  # state at the END of the timed region: what the line's self-check fields describe
    launches = st1["kernel_launches"] - launches0
    value = n * args.steps / (ms * 1e-3)
    check = check_counts(args.workload, n, st1["move_passes"], st1["last_pair_count"] if collisions else None, st1["last_flagged_count"] if collisions else None)

    # per-kernel device time over the same K steps (events around every launch; separate pass so the
    # event records do not sit inside the throughput measurement)
    sim.profile_begin()
    timed_steps(args.steps)
    kt = sim.profile_end()
    total_kernel_ms = sum(t for _, t in kt.values())
    traffic = load_traffic(args.workload)
    kernels = []
    for name, (cnt, tms) in sorted(kt.items(), key=lambda kv: -kv[1][1]):
        bpe = KERNEL_BYTES.get(name)
        entry = {"name": name, "launches": cnt, "avg_us": tms / cnt * 1e3, "share": tms / total_kernel_ms if total_kernel_ms else None}
        if bpe:
            gbs = bpe * n / (tms / cnt * 1e-3) / 1e9
            entry.update({"alg_bytes_per_entity": bpe, "achieved_gbs": gbs, "frac": gbs / peak})
        kernels.append(entry)
    dom = next((k for k in kernels if "achieved_gbs" in k), None)
    roofline = None
    if dom:
        roofline = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": dom["frac"], "traffic": traffic.get(dom["name"]), "peak_source": peak_src,
                    "alg_bytes_per_launch": dom["alg_bytes_per_entity"] * n, "avg_launch_us": dom["avg_us"], "share_of_step": dom["share"]}
    tick_gbs = w["survey_bytes"] * n * args.steps / (ms * 1e-3) / 1e9
    if roofline and traffic.get("_issue_active_pct", {}).get(roofline["kernel"]) is not None:
        roofline["issue_active_pct_ncu"] = traffic["_issue_active_pct"][roofline["kernel"]]
        roofline["note"] = ("this kernel is instruction-issue-bound, not HBM-bound (ncu smsp__issue_active, profiles/r2_ncu_full.md); "
                            "its HBM fraction is reported because the contract asks for the dominant kernel")
    # the streaming move pass alone (collisions-off dispatch on the same resident population): the HBM-bound kernel of the path
    move_only = None
    if collisions:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sim.enqueue_ticks(3, False)
        e0.record(stream)
        sim.enqueue_ticks(args.steps, False)
        sim.join()
        e1.record(stream)
        e1.synchronize()
        mo_ms = e0.elapsed_time(e1) / args.steps
        mo_gbs = 24.0 * n / (mo_ms * 1e-3) / 1e9
        move_only = {"what": "move pass only (pass A streaming + pass B arrivals), 24 B per entity-update", "ms_per_step": mo_ms,
                     "value": n / (mo_ms * 1e-3), "achieved_gbs": mo_gbs, "frac_of_measured_peak": mo_gbs / peak}

    # the same tick with MSIM_FLAG_NO_PAIR_COUNT: colours only, the query stops at an entity's first neighbour.  The colours are all the
    # reference's host ever looks at (it discards debugData, Simulator.cpp:273); `value` above keeps the exact unique-pair count switched on
    flags_only = None
    if collisions and not args.no_flags_only:
        try:
            sim2 = M.Simulation(m, ents, radius=10.0, device=local_rank, flags=flags | M.FLAG_NO_PAIR_COUNT, stream=stream.cuda_stream)
            sim2.dispatch(2)
            sim2.enqueue_ticks(args.preroll, False)
            sim2.enqueue_ticks(5, True)
            sim2.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sim2.enqueue_ticks(args.steps, True)
            sim2.join()
            e1.record(stream)
            e1.synchronize()
            fo_ms = e0.elapsed_time(e1) / args.steps
            flags_only = {"what": "sim tick with MSIM_FLAG_NO_PAIR_COUNT (collision colours exact, no pair count)", "ms_per_step": fo_ms,
                          "value": n / (fo_ms * 1e-3), "flagged_last_tick": sim2.stats()["last_flagged_count"]}
            sim2.close()
        except Exception as ex:  # an extra, never the headline: report and go on
            flags_only = {"error": repr(ex)}

    # keep the GPU under the same load long enough for nvidia-smi to sample it (untimed)
    t_end = time.time() + 1.2
    while time.time() < t_end:
        timed_steps(max(1, min(args.steps, 50)))
    clocks = sampler.stop()

    # ---- e2e: through the C ABI with HOST (pinned) buffers: upload AoS + sim tick + readback AoS ----
    pinned = torch.empty(n * 64, dtype=torch.uint8, pin_memory=True)
    ptr = pinned.data_ptr()
    sim.read_entities_ptr(ptr, n)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    tick = 4

    def e2e_step():
        nonlocal tick
        sim.upload_ptr(ptr, n)
        sim.dispatch(tick)
        if collisions:
            sim.dispatch(tick + 1)
        tick += 2
        sim.read_entities_ptr(ptr, n)

    e2e_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record(stream)
    e1.synchronize()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = max(e0.elapsed_time(e1), e2e_wall * 1e3)
    e2e = {"value": n * e2e_steps / (e2e_ms * 1e-3), "unit": "entity-updates/s", "h2d_bytes_per_step": n * 64, "d2h_bytes_per_step": n * 64,
           "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
           "what": "msim_upload_entities(pinned AoS) + msim_dispatch(move)" + (" + msim_dispatch(collide)" if collisions else "") + " + msim_read_entities(pinned AoS), blocking calls"}
    variants = {"blocking_full_aos": dict(e2e)}
    if not args.no_e2e_variants:
        # (1) the same bytes, pipelined: the readback of step i leaves through the copy stream (msim_snapshot_begin/end, library-owned pinned
        # buffers) while step i+1 uploads and computes - PCIe is full duplex, so the two 64 B x N transfers overlap instead of adding up
        def piped_step(first):
            nonlocal tick
            sim.upload_ptr(ptr, n)
            sim.dispatch(tick)
            if collisions:
                sim.dispatch(tick + 1)
            tick += 2
            if not first:
                sim.snapshot_end(copy=False)  # result of the previous step
            sim.snapshot_begin()

        piped_step(True)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            piped_step(False)
        sim.snapshot_end(copy=False)
        wall = time.perf_counter() - t0
        variants["pipelined_full_aos"] = {
            "value": n * e2e_steps / wall, "unit": "entity-updates/s", "ms_per_step": wall * 1e3 / e2e_steps, "h2d_bytes_per_step": n * 64, "d2h_bytes_per_step": n * 64,
            "steps": e2e_steps, "what": "same bytes per step; msim_snapshot_begin/end instead of msim_read_entities (D2H of step i overlaps H2D + compute of step i+1), wall clock"}
        # (2) what the renderer consumes per frame (EntityGlObject.cpp:9-12: position + colour): 8 B position + 1 B collision flag per entity
        pos_host = torch.empty(n * 2, dtype=torch.float32, pin_memory=True)
        flag_host = torch.empty(n, dtype=torch.uint8, pin_memory=True)

        def render_step():
            nonlocal tick
            sim.upload_ptr(ptr, n)
            sim.dispatch(tick)
            if collisions:
                sim.dispatch(tick + 1)
            tick += 2
            sim.read_positions_ptr(pos_host.data_ptr(), n)
            if collisions:
                sim.read_collision_flags_ptr(flag_host.data_ptr(), n)

        render_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            render_step()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        variants["full_aos_up_render_state_down"] = {
            "value": n * e2e_steps / wall, "unit": "entity-updates/s", "ms_per_step": wall * 1e3 / e2e_steps, "h2d_bytes_per_step": n * 64,
            "d2h_bytes_per_step": n * (9 if collisions else 8), "steps": e2e_steps,
            "what": "msim_upload_entities(pinned AoS) + sim tick + msim_read_positions" + (" + msim_read_collision_flags" if collisions else "") + " (pinned), blocking calls"}
        # the headline keeps the FULL 64-byte AoS in both directions (what sim::Simulator::get_entities hands the UI): the faster of the two
        best = min(("blocking_full_aos", "pipelined_full_aos"), key=lambda k: variants[k]["ms_per_step"])
        e2e = dict(variants[best])
        e2e["variant"] = best
    e2e["variants"] = variants

    # ---- CPU baseline on this box's host cores (bounded sample of the same workload) ----
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O

        threads = os.cpu_count() or 1
        ref_move = O.ref_shader_available()
        ref_tree = O.ref_available() and collisions
        sample = min(n, args.cpu_sample, int(O.ref().ref_capacity()) if ref_tree else 1 << 62)
        host = np.frombuffer(pinned.numpy(), dtype=M.ENTITY_DTYPE)[:sample]
        omap = O.OracleMap(m.width, m.height, m.roads.view(O.ROAD_DTYPE), m.connections)
        v, sec, _, _ = time_cpu(O, host, omap, 10.0, collisions, args.cpu_steps, 1, ref_move, ref_tree, threads)
        v_port, _, _, _ = time_cpu(O, host, omap, 10.0, collisions, args.cpu_steps, 1, False, False, threads)
        kind, how = cpu_arm_description(ref_move, ref_tree, collisions)
        cpu = {"value": v, "unit": "entity-updates/s", "cores": threads, "kind": kind, "entities_sampled": int(sample), "same_config": int(sample) == n,
               "sample": f"first {sample} entities of the resident population, {args.cpu_steps} sim ticks, {sec:.3f} s per tick; {how}",
               "port_value": v_port}
        try:  # an extra beside the baseline: never allowed to cost the line
            cpu["whole_shader_1_thread"] = time_whole_shader(O, host, m, 10.0, collisions)
        except Exception as ex:
            cpu["whole_shader_1_thread"] = {"error": repr(ex)}

    line = {
        "metric": METRIC, "value": value, "unit": "entity-updates/s", "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
        "config": {"workload": args.workload, "entities": n, "collisions": collisions, "collision_radius_m": 10.0, "map": w["map_desc"],
                   "entity_seed": 42, "preroll_move_passes": args.preroll, "pair_count": collisions,
                   "move_passes_done": st1["move_passes"], "pairs_last_tick": st1["last_pair_count"] if collisions else None,
                   "flagged_last_tick": st1["last_flagged_count"] if collisions else None, "counts_check": check,
                   "resort": {"every_collision_passes": RESORT_EVERY, "in_timed_region": st1["reorders"] - st0["reorders"],
                              "alignment_ticks_untimed": aligned,
                              "note": "the timed region starts on a re-sorting tick: it holds 1 + (K - 1) // 32 re-sorts, never fewer than its share"},
                   "l2": ("flushed between timed steps (512 MiB fill)" if small else "inputs larger than L2 (no flush)"),
                   "experiments": {k: v for k, v in os.environ.items() if k.startswith("MSIM_") and k != "MSIM_BENCH_RESULT_FD"},
                   "grid": sim.stats()},
        "roofline": roofline,
        "tick": {"survey_bytes_per_entity_update": w["survey_bytes"], "achieved_gbs": tick_gbs, "frac_of_measured_peak": tick_gbs / peak,
                 "frac_of_nominal_8tbs": tick_gbs / 8000.0, "kernel_time_ms_per_step": total_kernel_ms / args.steps},
        "kernels": kernels,
        "move_only": move_only,
        "flags_only": flags_only,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    sim.close()
    emit(line)
    return 3 if check.get("status") == "mismatch" else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="munich_10m_collisions")
    ap.add_argument("--entities", type=int, default=None, help="override the per-GPU entity count")
    ap.add_argument("--preroll", type=int, default=256, help="untimed move passes that disperse the population before timing")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--ref-sample", type=int, default=0, help="--impl reference: entities of the population to run (0 = all of them, the GPU arm's configuration)")
    ap.add_argument("--ref-budget-s", type=float, default=100.0, help="--impl reference: stop timing after the step that crosses this many seconds")
    ap.add_argument("--no-whole-shader", action="store_true", help="skip the single-thread whole-shader extra of the CPU arms")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flags-only", action="store_true", help="skip the extra colours-only measurement")
    ap.add_argument("--scaling", choices=["strong", "weak"], default="strong",
                    help="N > 1: strong = the workload's population split over N GPUs (BASELINE config 3), weak = that population per GPU")
    ap.add_argument("--exchange", choices=["p2p", "collective"], default="p2p",
                    help="N > 1 with collisions: p2p = move kernel stores into the neighbours' buffers over peer memory; collective = NCCL all_to_all per tick")
    ap.add_argument("--presort", action="store_true", help="experiment: upload the entities in cell order")
    ap.add_argument("--counting-sort", action="store_true", help="force the single-digit counting sort")
    ap.add_argument("--onesweep", action="store_true", help="force the multi-pass onesweep radix sort")
    ap.add_argument("--no-reorder", action="store_true", help="keep the state in upload order (onesweep rebuild)")
    ap.add_argument("--no-e2e-variants", action="store_true", help="only the blocking full-AoS round trip (skips the pipelined and render-state variants)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
