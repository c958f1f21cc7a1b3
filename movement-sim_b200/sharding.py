"""Host-side logic of the multi-GPU path: one process per GPU, torch.distributed for the
plumbing.  The reference is single-device (/root/reference/src/sim/Simulator.cpp:52); this module makes
N msim handles behave like one.

  collisions off  entities are independent -> contiguous entity ranges, no data-path collective at all
  collisions on   spatial bands of whole cell rows, road graph replicated.  Per tick each rank enqueues
                  move+pack, integrate, collide and never waits for its GPU.  Two transports:
                    p2p         (default on GPUs) the library's move / exchange kernels store leavers and halo
                                straight into the neighbours' receive buffers over NVLink peer memory and
                                synchronise through flag words; this module only exchanges the CUDA IPC
                                handles once (msim_shard.h, "peer-memory exchange")
                    collective  each rank sends ONE fixed-size device buffer (migrant records + boundary-row
                                halo, packed by the library's kernels) to the rank below and one to the rank
                                above with a single all_to_all_single (send/recv pairs over gloo)
                  Every REBALANCE_EVERY ticks a (rows x u32) all-reduce of the per-row histogram re-chooses
                  the split rows; boundaries then walk one row per tick towards the target, which turns
                  re-balancing into ordinary migration.

The compute object is abstract (`engine`): the product engine is CudaShardEngine (C ABI ->
hand-written kernels); the CPU tests drive the same orchestration over gloo with an oracle-backed
engine (tests/shard_oracle_engine.py)."""
from __future__ import annotations

import json
import os
import time

import numpy as np

REBALANCE_EVERY = 64
CHUNK = 1 << 20  # entities are generated in seeded chunks so that every rank can rebuild the global population


# --------------------------------------------------------------------------------------------------
# partitioning (pure numpy; unit-tested on CPU)
# --------------------------------------------------------------------------------------------------
def balanced_splits(row_hist: np.ndarray, world: int) -> np.ndarray:
    """Split rows [0, R) into `world` contiguous non-empty bands holding ~equal entity counts.
    Returns int64 array s of length world+1 with s[0] = 0, s[-1] = R; band r = rows [s[r], s[r+1])."""
    rows = int(row_hist.shape[0])
    if world > rows:
        raise ValueError(f"{world} ranks but only {rows} cell rows")
    cum = np.concatenate([[0], np.cumsum(row_hist.astype(np.int64))])
    total = int(cum[-1])
    s = np.zeros(world + 1, dtype=np.int64)
    s[-1] = rows
    for r in range(1, world):
        want = total * r / world
        k = int(np.searchsorted(cum, want, side="left"))
        # the boundary before row k or before row k-1, whichever is closer to the ideal count
        if k > 0 and abs(cum[k - 1] - want) <= abs(cum[min(k, rows)] - want):
            k -= 1
        s[r] = k
    for r in range(1, world):  # strictly increasing, every band at least one row
        s[r] = max(s[r], s[r - 1] + 1)
    for r in range(world - 1, 0, -1):
        s[r] = min(s[r], s[r + 1] - 1)
    return s


def step_towards(splits: np.ndarray, target: np.ndarray) -> np.ndarray:
    """Move every interior boundary at most one row towards `target`, keeping bands non-empty."""
    out = splits.copy()
    for r in range(1, len(splits) - 1):
        out[r] += np.sign(target[r] - splits[r])
    for r in range(1, len(out) - 1):
        out[r] = max(out[r], out[r - 1] + 1)
    for r in range(len(out) - 2, 0, -1):
        out[r] = min(out[r], out[r + 1] - 1)
    return out


def entity_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Collisions off: contiguous index range of `rank` (no collective ever needed)."""
    return total * rank // world, total * (rank + 1) // world


def generate_population(M, m, total: int, seed: int, box=None):
    """Yields (first_gid, entities) chunks of the seeded global population; identical on every rank."""
    done, chunk = 0, 0
    while done < total:
        k = min(CHUNK, total - done)
        yield done, m.init_entities(k, seed=seed + 1000 * chunk, box=box)
        done += k
        chunk += 1


def _chunks(total: int):
    """(chunk number, first gid, count) of the seeded chunks generate_population yields."""
    out, done, chunk = [], 0, 0
    while done < total:
        k = min(CHUNK, total - done)
        out.append((chunk, done, k))
        done += k
        chunk += 1
    return out


def _map_chunks(fn, total: int, threads: int):
    """fn over the chunks, in chunk order; the chunks carry their own seeds, so they can be generated side by side (the C generators
    run without the GIL).  Building a 100 M-entity population costs 0.25 s of host time per million entities on one core."""
    chunks = _chunks(total)
    if threads <= 1 or len(chunks) <= 1:
        return [fn(c) for c in chunks]
    from concurrent.futures import ThreadPoolExecutor

    with ThreadPoolExecutor(max_workers=threads) as pool:
        return list(pool.map(fn, chunks))


def host_threads(world: int) -> int:
    """Host threads one rank may use while the population is being built (all ranks of a node do it at the same time)."""
    return max(1, min(8, (os.cpu_count() or 1) // max(1, world)))


def global_row_histogram(M, m, total: int, seed: int, radius: float, box=None, threads: int = 1):
    """Entities per cell row of the seeded global population.  Every entity starts on the first point of its road, so the road-index
    stream of the generator is all that is needed (msim_entities_init_roads: one draw per entity instead of ten and no 64-byte records)."""
    def one(c):
        chunk, _, k = c
        idx = m.init_road_indices(k, seed=seed + 1000 * chunk, box=box)
        rows, ncx, ncy = M.grid_rows(m.width, m.height, radius, m.roads["start_pos"][idx])
        return np.bincount(rows, minlength=ncy).astype(np.int64), ncx, ncy

    parts = _map_chunks(one, total, threads)
    if not parts:
        _, ncx, ncy = M.grid_rows(m.width, m.height, radius, np.zeros((0, 2), dtype=np.float32))
        return np.zeros(ncy, dtype=np.int64), ncx, ncy
    return sum(p[0] for p in parts), parts[0][1], parts[0][2]


def collect_band(M, m, total: int, seed: int, radius: float, row_lo: int, row_hi: int, box=None, threads: int = 1):
    """This rank's slice of the global population: entities whose start row is in [row_lo, row_hi)."""
    def one(c):
        chunk, first, k = c
        ents = m.init_entities(k, seed=seed + 1000 * chunk, box=box)
        rows, _, _ = M.grid_rows(m.width, m.height, radius, ents["pos"])
        keep = (rows >= row_lo) & (rows < row_hi)
        return ents[keep], (first + np.nonzero(keep)[0]).astype(np.uint32)

    parts = _map_chunks(one, total, threads)
    if not parts:
        return m.init_entities(0, seed=seed, box=box), np.zeros(0, dtype=np.uint32)
    return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])


# --------------------------------------------------------------------------------------------------
# orchestration
# --------------------------------------------------------------------------------------------------
class ShardedSimulation:
    """Drives one rank's engine through sharded sim ticks.  `dist` is torch.distributed (nccl on GPUs,
    gloo in the CPU tests); buffers are torch tensors on the engine's device."""

    def __init__(self, engine, rank: int, world: int, splits: np.ndarray, rows: int, dist, torch, device, migrant_capacity: int,
                 halo_capacity: int, buffer_bytes: int, rebalance_every: int = REBALANCE_EVERY, exchange: str = "collective"):
        self.engine, self.rank, self.world, self.dist, self.torch = engine, rank, world, dist, torch
        self.splits = np.asarray(splits, dtype=np.int64).copy()
        self.target = self.splits.copy()
        self.rows = rows
        self.rebalance_every = rebalance_every
        self.ticks = 0
        self.device = device
        self.has_down, self.has_up = rank > 0, rank + 1 < world
        # one send and one receive arena, [down | up]: lets NCCL move both directions with ONE collective call
        # (all_to_all_single with zero-sized splits for non-neighbours) instead of four point-to-point ops
        self.send_arena = torch.zeros(2 * buffer_bytes, dtype=torch.uint8, device=device)
        self.recv_arena = torch.zeros(2 * buffer_bytes, dtype=torch.uint8, device=device)
        self.send_down = self.send_arena[:buffer_bytes] if self.has_down else None
        self.send_up = self.send_arena[buffer_bytes:] if self.has_up else None
        self.recv_down = self.recv_arena[:buffer_bytes] if self.has_down else None
        self.recv_up = self.recv_arena[buffer_bytes:] if self.has_up else None
        self.splits_bytes = [0] * world
        if self.has_down:
            self.splits_bytes[rank - 1] = buffer_bytes
        if self.has_up:
            self.splits_bytes[rank + 1] = buffer_bytes
        self.use_all_to_all = world > 1 and str(device).startswith("cuda")
        self.migrant_capacity, self.halo_capacity = migrant_capacity, halo_capacity
        self.exchanged_bytes = 0
        self.phase_us = {}
        self.fused_move_pack = hasattr(engine, "move_pack")  # one kernel moves and packs (msim_shard_move_pack)
        # exchange = "p2p": the move kernel stores into the neighbours' receive buffers over NVLink peer memory and the
        # integrate kernel waits on a flag: no collective call per tick.  "collective": one all_to_all_single per tick.
        self.exchange = exchange if world > 1 else "collective"
        if self.exchange == "p2p":
            self.exchange = "p2p" if self._connect_p2p() else "collective"

    def _connect_p2p(self) -> bool:
        """Exchange the CUDA IPC handles of the receive arenas and open the neighbours'.  All ranks agree on the outcome."""
        ok, handle = 1, b""
        try:
            handle = self.engine.p2p_create()
        except Exception as e:  # e.g. IPC not permitted in this container
            ok, self.p2p_error = 0, str(e)
        handles = [None] * self.world
        self.dist.all_gather_object(handles, (ok, handle))
        if all(h[0] for h in handles):
            try:
                self.engine.p2p_connect(handles[self.rank - 1][1] if self.has_down else None, handles[self.rank + 1][1] if self.has_up else None)
            except Exception as e:
                ok, self.p2p_error = 0, str(e)
        flags = [None] * self.world
        self.dist.all_gather_object(flags, ok)
        return all(flags)

    @property
    def band(self):
        return int(self.splits[self.rank]), int(self.splits[self.rank + 1])

    def _exchange(self):
        d = self.dist
        if self.use_all_to_all:
            lo = 0 if self.has_down else self.send_arena.numel() // 2
            hi = self.send_arena.numel() if self.has_up else self.send_arena.numel() // 2
            d.all_to_all_single(self.recv_arena[lo:hi], self.send_arena[lo:hi], self.splits_bytes, self.splits_bytes)
            self.exchanged_bytes += hi - lo
            return
        ops = []
        if self.has_up:
            ops += [d.P2POp(d.isend, self.send_up, self.rank + 1), d.P2POp(d.irecv, self.recv_up, self.rank + 1)]
        if self.has_down:
            ops += [d.P2POp(d.isend, self.send_down, self.rank - 1), d.P2POp(d.irecv, self.recv_down, self.rank - 1)]
        if ops:
            for req in d.batch_isend_irecv(ops):
                req.wait()
            self.exchanged_bytes += sum(op.tensor.numel() for op in ops) // 2

    def tick(self, collide: bool = True):
        """One sim tick: move, migrate/halo exchange, (collision pass)."""
        if not np.array_equal(self.splits, self.target):
            self.splits = step_towards(self.splits, self.target)
        lo, hi = self.band
        if self.exchange == "p2p":
            t0 = time.perf_counter()
            self.engine.p2p_move_pack(lo, hi)
            t2 = time.perf_counter()
            self.engine.p2p_integrate()
            t3 = time.perf_counter()
            if collide:
                self.engine.collide()
            t4 = time.perf_counter()
            for k, v in (("enqueue_move_pack", t2 - t0), ("exchange_enqueue", 0.0), ("integrate", t3 - t2), ("enqueue_collide", t4 - t3)):
                self.phase_us[k] = 0.9 * self.phase_us.get(k, v * 1e6) + 0.1 * v * 1e6
            self.ticks += 1
            if self.rebalance_every and self.world > 1 and self.ticks % self.rebalance_every == 0:
                self.rebalance()
            return
        t0 = time.perf_counter()
        if self.fused_move_pack:
            self.engine.move_pack(lo, hi, self.send_down, self.send_up)
        else:
            self.engine.move()
            self.engine.pack(lo, hi, self.send_down, self.send_up)
        t1 = time.perf_counter()
        self._exchange()
        t2 = time.perf_counter()
        self.engine.integrate(self.recv_down, self.recv_up)
        t3 = time.perf_counter()
        if collide:
            self.engine.collide()
        t4 = time.perf_counter()
        # host-side wall time per phase (integrate contains the tick's only host<->device round trip)
        for k, v in (("enqueue_move_pack", t1 - t0), ("exchange_enqueue", t2 - t1), ("integrate", t3 - t2), ("enqueue_collide", t4 - t3)):
            self.phase_us[k] = 0.9 * self.phase_us.get(k, v * 1e6) + 0.1 * v * 1e6
        self.ticks += 1
        if self.rebalance_every and self.world > 1 and self.ticks % self.rebalance_every == 0:
            self.rebalance()

    def rebalance(self):
        """All-reduce the per-row histogram and re-choose the split rows (dense crowds, config 5)."""
        hist = self.torch.from_numpy(self.engine.row_histogram(self.rows).astype(np.int64)).to(self.device)
        self.dist.all_reduce(hist)
        self.target = balanced_splits(hist.cpu().numpy(), self.world)

    def global_sum(self, value: int) -> int:
        t = self.torch.tensor([int(value)], dtype=self.torch.int64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t)
        return int(t.item())


class CudaShardEngine:
    """The product engine: one msim handle (C ABI, hand-written sm_100a kernels)."""

    def __init__(self, M, sim, asynchronous: bool = True):
        self.M, self.sim, self.asynchronous = M, sim, asynchronous

    @staticmethod
    def _ptr(t):
        return None if t is None else t.data_ptr()

    def move(self):
        self.sim.enqueue_move()

    def pack(self, lo, hi, send_down, send_up):
        self.sim.shard_pack(lo, hi, self._ptr(send_down), self._ptr(send_up))

    def move_pack(self, lo, hi, send_down, send_up):
        self.sim.shard_move_pack(lo, hi, self._ptr(send_down), self._ptr(send_up))

    def p2p_create(self) -> bytes:
        return self.sim.shard_p2p_create()[0]

    def p2p_connect(self, down_handle, up_handle):
        self.sim.shard_p2p_connect(down_handle, up_handle)

    def p2p_move_pack(self, lo, hi):
        self.sim.shard_p2p_move_pack(lo, hi)

    def p2p_integrate(self):
        self.sim.shard_p2p_integrate()

    def integrate(self, recv_down, recv_up):
        if self.asynchronous:  # device-side integrate: nothing to wait for, the host keeps enqueueing
            return self.sim.shard_integrate_async(self._ptr(recv_down), self._ptr(recv_up))
        return self.sim.shard_integrate(self._ptr(recv_down), self._ptr(recv_up))

    def collide(self):
        self.sim.enqueue_collide()

    def row_histogram(self, rows):
        return self.sim.shard_row_histogram(rows)

    def stats(self):
        return self.sim.stats()

    def read_owned(self):
        self.sim.shard_counts()
        return self.sim.read_entities(), self.sim.shard_read_gids()


def make_cuda_shard(M, m, total: int, seed: int, radius: float, rank: int, world: int, dist, torch, local_rank: int, stream, box=None,
                    rebalance_every: int = REBALANCE_EVERY, exchange: str = "p2p"):
    """Builds this rank's band of the seeded global population on its GPU."""
    hist, ncx, ncy = global_row_histogram(M, m, total, seed, radius, box, threads=host_threads(world))
    splits = balanced_splits(hist, world)
    lo, hi = int(splits[rank]), int(splits[rank + 1])
    ents, gids = collect_band(M, m, total, seed, radius, lo, hi, box, threads=host_threads(world))
    max_row = int(hist.max())
    migrant_capacity = max(4096, 3 * max_row)
    halo_capacity = max(4096, 3 * max_row)
    capacity = int(ents.shape[0] * 1.3) + 4 * (migrant_capacity + halo_capacity) + 1024
    sim = M.Simulation(m, ents, radius=radius, device=local_rank, stream=stream.cuda_stream, capacity=capacity)
    sim.shard_enable(gids, migrant_capacity, halo_capacity)
    engine = CudaShardEngine(M, sim)
    device = torch.device("cuda", local_rank)
    sh = ShardedSimulation(engine, rank, world, splits, ncy, dist, torch, device, migrant_capacity, halo_capacity,
                           M.shard_buffer_bytes(migrant_capacity, halo_capacity), rebalance_every, exchange)
    return sh, sim


# --------------------------------------------------------------------------------------------------
# bench.py --gpus N (N > 1): launched by torch.distributed.run, one rank per GPU
# --------------------------------------------------------------------------------------------------
def bench_main(args, M, rank: int, world: int, local_rank: int) -> int:
    import torch
    import torch.distributed as dist

    from bench import KERNEL_BYTES, METRIC, RESORT_EVERY, ClockSampler, align_resort_phase, build_workload, check_counts, emit, load_peaks

    w, m = build_workload(M, args.workload, args.entities)
    collisions = w["collisions"]
    weak = getattr(args, "scaling", "strong") == "weak"
    # strong (default): BASELINE config 3 names ONE population (10 M) on 1/2/4/8 GPUs.  weak: 10 M per GPU on the same map.
    total = w["entities"] * world if weak else w["entities"]
    per_gpu = total // world
    peak, peak_src = load_peaks()
    stream = torch.cuda.Stream()
    device = torch.device("cuda", local_rank)
    sampler = None
    with torch.cuda.stream(stream):
        if collisions:
            sh, sim = make_cuda_shard(M, m, total, 42, 10.0, rank, world, dist, torch, local_rank, stream, box=w["box"],
                                      exchange=getattr(args, "exchange", "p2p"))
            sim.dispatch(2)  # the reference's first dispatch: initialise only
            step = lambda: sh.tick(True)
            for _ in range(args.preroll):
                sh.tick(False)
        else:
            lo, hi = entity_range(total, rank, world)
            parts = [e for _, e in generate_population(M, m, total, 42, w["box"])]
            ents = np.concatenate(parts)[lo:hi]
            sim = M.Simulation(m, ents, radius=10.0, device=local_rank, flags=M.FLAG_NO_COLLISIONS, stream=stream.cuda_stream)
            sh = None
            sim.dispatch(2)
            sim.enqueue_ticks(args.preroll, False)
            step = lambda: sim.enqueue_ticks(1, False)
        step()  # (the single-GPU arm's one untimed tick behind the pre-roll: every N is at the same tick when timing starts)
        for _ in range(max(3, args.warmup)):
            step()
        aligned = align_resort_phase(sim.stats, step) if collisions else 0  # same cadence on every rank: same number of ticks
        sim.sync()
        st0 = sim.stats()
        if rank == 0:
            props = torch.cuda.get_device_properties(local_rank)
            uuid = getattr(props, "uuid", None)
            sampler = ClockSampler(f"GPU-{uuid}" if uuid and not str(uuid).startswith("GPU-") else (str(uuid) if uuid else None))
        launches0 = sim.stats()["kernel_launches"]
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        sim.join()  # the library's side / push streams (pass B, the exchange of the last tick) are inside the timed region too
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        dist.barrier()
        ms_local = e0.elapsed_time(e1)
        t = torch.tensor([ms_local], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # device time, max over ranks
        ms = float(t.item())
        st1 = sim.stats()  # state at the END of the timed region
        launches = st1["kernel_launches"] - launches0
        owned = torch.tensor([st1["entity_count"]], dtype=torch.int64, device=device)
        gathered = [torch.zeros_like(owned) for _ in range(world)]
        dist.all_gather(gathered, owned)
        per_rank = [int(g.item()) for g in gathered]
        pairs = flagged = None
        counts_by_rank = [None] * world
        dist.all_gather_object(counts_by_rank, [int(st1["last_pair_count"]), int(st1["last_flagged_count"])])
        if collisions:
            pairs, flagged = sh.global_sum(st1["last_pair_count"]), sh.global_sum(st1["last_flagged_count"])
        check = check_counts(args.workload, total, st1["move_passes"], pairs, flagged)
        # per-kernel device time on every rank over a second pass of the same K steps (events around every launch); rank 0's feed the roofline
        sim.profile_begin()
        ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ep0.record(stream)
        for _ in range(args.steps):
            step()
        ep1.record(stream)
        ep1.synchronize()
        kt = sim.profile_end()
        kernels = {name: round(t / args.steps * 1e3, 1) for name, (cnt, t) in sorted(kt.items(), key=lambda kv: -kv[1][1])}
        kernels["_step_us_with_events"] = round(ep0.elapsed_time(ep1) / args.steps * 1e3, 1)
        kernels_by_rank = [None] * world
        dist.all_gather_object(kernels_by_rank, kernels)
        # keep the load up for the clock sampler (untimed).  A FIXED number of steps derived from the all-reduced
        # time: every rank must enqueue the same number of exchanges (a wall-clock loop would not)
        for _ in range(max(10, min(5000, int(1000.0 / max(ms / args.steps, 0.01))))):
            step()
        sim.sync()
        clocks = sampler.stop() if sampler else None

        # e2e per rank: host (pinned) AoS up, one tick, AoS back; max over ranks
        n_local = sim.stats()["entity_count"]
        pinned = torch.empty(int(n_local * 1.2 + 4096) * 64, dtype=torch.uint8, pin_memory=True)
        ptr = pinned.data_ptr()
        sim.read_entities_ptr(ptr, n_local)
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        h2d = d2h = 0

        def e2e_step():
            nonlocal n_local, h2d, d2h
            sim.upload_ptr(ptr, n_local)
            h2d += n_local * 64
            step()
            sim.sync()
            n_local = sim.stats()["entity_count"]
            sim.read_entities_ptr(ptr, n_local)
            d2h += n_local * 64

        e2e_step()
        h2d = d2h = 0
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        bytes_t = torch.tensor([h2d, d2h], dtype=torch.int64, device=device)
        dist.all_reduce(bytes_t)
    value = total * args.steps / (ms * 1e-3)
    tick_gbs = w["survey_bytes"] * value / 1e9
    roofline = None
    if rank == 0 and kernels:
        # dominant kernel on rank 0 (events around every launch, second pass of the same K steps): algorithmic bytes of ONE launch
        # = bytes per entity (bench.KERNEL_BYTES) x the entities rank 0 owns, over that kernel's average launch time
        step_us = sum(v for k, v in kernels.items() if not k.startswith("_") and k != "arrive")  # pass B runs beside the query
        for name, us in kernels.items():
            bpe = KERNEL_BYTES.get(name)
            if name.startswith("_") or not bpe:
                continue
            if name == "move" and collisions:
                bpe = 28.0  # sharded handles also write the 4 B cell key (the exchange kernels work on keys)
            elif name == "cell_scatter" and collisions:
                bpe = 24.0  # ... and the scatter reads it
            n0 = per_rank[0]
            gbs = bpe * n0 / (us * 1e-6) / 1e9
            roofline = {"bound": "hbm", "kernel": name, "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                        "traffic": None, "peak_source": peak_src, "alg_bytes_per_launch": bpe * n0, "avg_launch_us": us,
                        "share_of_step": us / step_us if step_us else None, "rank": 0,
                        "note": ("query is instruction-issue-bound, not HBM-bound (profiles/): its HBM fraction is reported because the contract asks "
                                 "for the dominant kernel; see tick.frac_of_measured_peak for the whole step") if name == "query" else None}
            break
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "entity-updates/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
            "config": {"workload": args.workload, "entities_per_gpu": per_gpu, "entities_total": total, "collisions": collisions,
                       "collision_radius_m": 10.0, "map": w["map_desc"], "entity_seed": 42, "preroll_move_passes": args.preroll, "entities": total,
                       "parallelism": (f"{world} spatial bands of cell rows; halo + migrants per tick "
                                       + ("stored by the move kernel into the neighbours' buffers over NVLink peer memory, flag-synchronised (no collective call)"
                                          if sh.exchange == "p2p" else "through one NCCL all_to_all_single")
                                       + f"; row-histogram all-reduce every {REBALANCE_EVERY} ticks"
                                       if collisions else f"{world} entity ranges, no collective"),
                       "exchange": (sh.exchange if sh else None), "p2p_fallback_reason": getattr(sh, "p2p_error", None),
                       "owned_per_rank": per_rank, "l2": ("inputs larger than L2 (no flush)" if per_gpu * 40 > 200e6 else "per-GPU working set may sit in L2 (strong scaling of a fixed population)"),
                       "phase_us_rank0": getattr(sh, "phase_us", None), "kernel_us_per_step_rank0": kernels, "kernel_us_per_step_by_rank": kernels_by_rank,
                       "exchange_buffer_bytes": (M.shard_buffer_bytes(sh.migrant_capacity, sh.halo_capacity) if sh else 0),
                       "move_passes_done": st1["move_passes"], "pairs_last_tick": pairs, "flagged_last_tick": flagged, "counts_check": check,
                       "pairs_flagged_by_rank": counts_by_rank, "splits": [int(x) for x in sh.splits] if sh else None,
                       "resort": {"every_collision_passes": RESORT_EVERY, "in_timed_region": st1["reorders"] - st0["reorders"],
                                  "alignment_ticks_untimed": aligned,
                                  "note": "the timed region starts on a re-sorting tick: it holds 1 + (K - 1) // 32 re-sorts, never fewer than its share"}},
            "roofline": roofline,
            "tick": {"survey_bytes_per_entity_update": w["survey_bytes"], "achieved_gbs": tick_gbs, "frac_of_measured_peak": tick_gbs / (peak * world),
                     "frac_of_nominal_8tbs": tick_gbs / (8000.0 * world), "peak_source": peak_src},
            "cpu_baseline": None,
            "e2e": {"value": total * e2e_steps / float(dt.item()), "unit": "entity-updates/s", "h2d_bytes_per_step": int(bytes_t[0].item()) // e2e_steps,
                    "d2h_bytes_per_step": int(bytes_t[1].item()) // e2e_steps, "steps": e2e_steps,
                    "what": "per rank: msim_upload_entities(pinned AoS) + one sharded sim tick + msim_read_entities(pinned AoS); wall clock, max over ranks"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        emit(line)
    dist.barrier()
    sim.close()
    dist.destroy_process_group()
    return 3 if check.get("status") == "mismatch" else 0
