"""Worker of the sharded-simulation tests: runs on every rank (gloo + oracle engine on CPU, nccl + CUDA
engine on GPUs), simulates `ticks` sharded sim ticks and writes this rank's owned entities + gids +
pair counts to `outdir`.  The parent test compares the union with an unsharded oracle run."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def make_map(M, cfg):
    if cfg["map"] == "city":
        return M.Map.city(*cfg["city"])
    return M.Map.load_json(os.path.join(ROOT, "tests", "golden", "test_map.json"))


def run(rank, world, port, backend, outdir, cfg):
    import torch
    import torch.distributed as dist

    import movement_sim_b200 as M
    from movement_sim_b200 import sharding as S

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if backend == "nccl":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    m = make_map(M, cfg)
    total, seed, radius = cfg["entities"], cfg["seed"], cfg["radius"]
    box = cfg.get("box")
    hist, ncx, ncy = S.global_row_histogram(M, m, total, seed, radius, box)
    splits = S.balanced_splits(hist, world)
    if cfg.get("skew_splits"):  # start from a deliberately bad partition so that re-balancing has work to do
        splits = np.linspace(0, ncy, world + 1).astype(np.int64)
    lo, hi = int(splits[rank]), int(splits[rank + 1])
    ents, gids = S.collect_band(M, m, total, seed, radius, lo, hi, box)
    mig_cap = hal_cap = cfg.get("capacity", 1 << 15)
    nbytes = M.shard_buffer_bytes(mig_cap, hal_cap)
    if backend == "nccl":
        stream = torch.cuda.Stream()
        ctx = torch.cuda.stream(stream)
        ctx.__enter__()
        sim = M.Simulation(m, ents, radius=radius, device=rank, stream=stream.cuda_stream, capacity=total + 4 * (mig_cap + hal_cap))
        sim.shard_enable(gids, mig_cap, hal_cap)
        sim.dispatch(2)
        engine = S.CudaShardEngine(M, sim)
        device = torch.device("cuda", rank)
    else:
        from oracle import oracle as O
        from shard_oracle_engine import OracleShardEngine

        engine = OracleShardEngine(O, M, m, ents, gids, radius, mig_cap, hal_cap)
        engine.move()  # the init-only first dispatch
        device = torch.device("cpu")
    sh = S.ShardedSimulation(engine, rank, world, splits, ncy, dist, torch, device, mig_cap, hal_cap, nbytes, cfg.get("rebalance_every", 8),
                             exchange=cfg.get("exchange", "collective"))
    pairs, owned_hist, split_hist = [], [], []
    for t in range(cfg["ticks"]):
        sh.tick(True)
        if backend == "nccl":
            sim.sync()
        st = engine.stats()
        pairs.append(sh.global_sum(st["last_pair_count"]))
        owned_hist.append(int(st["entity_count"]))
        split_hist.append([int(v) for v in sh.splits])
    e, g = engine.read_owned()
    np.save(os.path.join(outdir, f"ents_{rank}.npy"), np.ascontiguousarray(e).view(np.uint8))
    np.save(os.path.join(outdir, f"gids_{rank}.npy"), g)
    with open(os.path.join(outdir, f"meta_{rank}.json"), "w") as f:
        json.dump({"pairs": pairs, "owned": owned_hist, "splits": split_hist, "exchanged_bytes": sh.exchanged_bytes, "exchange": sh.exchange,
                   "p2p_error": getattr(sh, "p2p_error", None)}, f)
    dist.barrier()
    if backend == "nccl":
        sim.close()
    dist.destroy_process_group()


if __name__ == "__main__":  # torchrun / mp.spawn entry for the GPU test
    cfg = json.loads(sys.argv[1])
    run(int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["MASTER_PORT"]), sys.argv[2], sys.argv[3], cfg)
