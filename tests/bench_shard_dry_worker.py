"""Worker of tests/test_bench_gpu_arm_dry_run.py::test_multi_gpu_arm_assembles_its_json_line: one rank of bench.py's N > 1 arm
(movement-sim_b200/sharding.py: bench_main) over gloo on the CPU.  torch.cuda is replaced by stand-ins, the CUDA engine by the oracle-backed
engine of the sharding tests; the population build, the partition, the orchestration, the collectives and the assembly of the JSON line are
the real code."""
import contextlib
import ctypes
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


class FakeEvent:
    clock = 0.0

    def __init__(self, enable_timing=True):
        self.t = None

    def record(self, stream=None):
        FakeEvent.clock += 410.0  # a long made-up step keeps the untimed "hold the load for the clock sampler" loop at its minimum of 10 steps
        self.t = FakeEvent.clock

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return other.t - self.t


def run(rank, world, port, out_path, argv):
    import torch
    import torch.distributed as dist

    import bench
    import movement_sim_b200 as M
    from movement_sim_b200 import sharding as S
    from oracle import oracle as O
    from shard_oracle_engine import OracleShardEngine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    # ---- stand-ins for the CUDA-only pieces --------------------------------------------------------------------------------
    cpu = torch.device("cpu")
    real_empty = torch.empty
    torch.device = lambda *a, **k: cpu
    torch.empty = lambda *a, pin_memory=False, **k: real_empty(*a, **k)
    torch.cuda.Stream = lambda: types.SimpleNamespace(cuda_stream=0)
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.Event = FakeEvent
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.get_device_properties = lambda d: types.SimpleNamespace(uuid=None)
    bench.ClockSampler = lambda uuid: types.SimpleNamespace(stop=lambda: {"sm_mhz": None, "reasons": ["dry run"]})
    lines = []
    bench.emit = lines.append

    class FakeSim:
        def __init__(self, engine):
            self.engine, self.launches, self.profiling = engine, 0, False
            # the library's counters the bench reads: move passes, and the re-sort cadence of api.cu (first collision pass, then every 32nd)
            self.moves, self.reorders, self.since = 0, 0, 32
            for name in ("move", "move_pack"):
                if hasattr(engine, name):
                    setattr(engine, name, self._counted(getattr(engine, name), "moves"))
            if hasattr(engine, "collide"):
                engine.collide = self._counted(engine.collide, "collides")

        def _counted(self, fn, what):
            def wrapped(*a, **k):
                if what == "moves":
                    self.moves += 1
                else:
                    self.since += 1
                    if self.since >= 32:
                        self.reorders, self.since = self.reorders + 1, 0
                return fn(*a, **k)
            return wrapped

        def dispatch(self, tick):
            self.engine.move()  # the init-only first dispatch

        def join(self):
            pass

        def sync(self):
            pass

        def stats(self):
            self.launches += 11
            return dict(self.engine.stats(), kernel_launches=self.launches, move_passes=self.moves, reorders=self.reorders)

        def profile_begin(self):
            self.profiling = True

        def profile_end(self):
            return {"query": (3, 0.14), "move": (3, 0.06), "shard": (3, 0.09), "cell_scatter": (3, 0.04)}

        def read_entities_ptr(self, ptr, n):
            ctypes.memmove(ptr, self.engine.e.ctypes.data, n * 64)

        def upload_ptr(self, ptr, n):
            ctypes.memmove(self.engine.e.ctypes.data, ptr, n * 64)

        def close(self):
            pass

    def make_shard(M_, m, total, seed, radius, rank_, world_, dist_, torch_, local_rank, stream, box=None, rebalance_every=S.REBALANCE_EVERY,
                   exchange="p2p"):
        # the real partition and population build, with the thread count bench uses
        hist, ncx, ncy = S.global_row_histogram(M_, m, total, seed, radius, box, threads=S.host_threads(world_))
        splits = S.balanced_splits(hist, world_)
        ents, gids = S.collect_band(M_, m, total, seed, radius, int(splits[rank_]), int(splits[rank_ + 1]), box, threads=S.host_threads(world_))
        cap = max(4096, 3 * int(hist.max()))
        engine = OracleShardEngine(O, M_, m, ents, gids, radius, cap, cap)
        sh = S.ShardedSimulation(engine, rank_, world_, splits, ncy, dist_, torch_, cpu, cap, cap, M_.shard_buffer_bytes(cap, cap), rebalance_every,
                                 "collective")
        return sh, FakeSim(engine)

    S.make_cuda_shard = make_shard
    entity_range = S.entity_range

    class RangeSim(FakeSim):  # collisions off: plain entity ranges, no exchange
        def __init__(self, m, ents, **_):
            e = types.SimpleNamespace(e=np.ascontiguousarray(ents).view(O.ENTITY_DTYPE).copy(), om=O.OracleMap(m.width, m.height, m.roads.view(O.ROAD_DTYPE), m.connections))
            e.move = lambda: O.move_pass(e.e, e.om)
            e.stats = lambda: {"entity_count": e.e.shape[0], "last_pair_count": 0, "last_flagged_count": 0}
            super().__init__(e)

        def enqueue_ticks(self, k, collide):
            for _ in range(k):
                self.engine.move()

    M.Simulation = RangeSim
    assert entity_range(10, 0, 2) == (0, 5)

    sys.argv = ["bench.py", "--gpus", str(world), *argv]
    import argparse  # noqa: F401  (bench.main parses sys.argv)

    real_run_b200 = bench.run_b200

    def run_b200(args):  # bench.run_b200 minus the NCCL rendez-vous and the CUDA device selection
        return S.bench_main(args, M, rank, world, rank)

    bench.run_b200 = run_b200
    bench.quiet_stdout = lambda: None
    real_destroy = dist.destroy_process_group
    rc = bench.main()
    with open(out_path, "w") as f:
        json.dump({"rc": rc, "lines": lines}, f)
    del real_run_b200, real_destroy


if __name__ == "__main__":
    run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], json.loads(sys.argv[5]))
