"""Race hunt for the sharded tick: the sums of pairs / flagged entities over EVERY collision pass must be the same on 1 GPU and on N.

    python profiles/stress_shard.py --ticks 3000 --every 50                         (N = 1: writes gpurun_out/stress_n1.json)
    python -m torch.distributed.run --nproc-per-node N ... profiles/stress_shard.py --ticks 3000 --every 50 --modes p2p,p2p,collective

Every `--every` ticks the running totals (msim_stats.total_pair_count / total_flagged_count, summed over ranks) are recorded; rank 0
compares them with the N = 1 file and with the first repetition of the same N and prints the first window that differs, per rank.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ticks", type=int, default=2000)
    ap.add_argument("--every", type=int, default=50)
    ap.add_argument("--preroll", type=int, default=280)
    ap.add_argument("--workload", default="munich_10m_collisions")
    ap.add_argument("--entities", type=int, default=None)
    ap.add_argument("--modes", default="p2p,p2p")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    import torch

    import movement_sim_b200 as M
    from bench import build_population, build_workload
    from movement_sim_b200 import sharding as S

    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    M.lib()
    torch.cuda.set_device(local_rank)
    w, m = build_workload(M, args.workload, args.entities)
    total = w["entities"]
    stream = torch.cuda.Stream()
    os.makedirs("gpurun_out", exist_ok=True)
    n1_path = f"gpurun_out/stress_n1_{args.workload}_{total}.json"

    if world == 1:
        ents = build_population(M, m, total, w["box"])
        sim = M.Simulation(m, ents, radius=10.0, device=local_rank, stream=stream.cuda_stream)
        sim.dispatch(2)
        sim.enqueue_ticks(args.preroll, False)
        seq = []
        t0 = time.perf_counter()
        for i in range(args.ticks):
            sim.enqueue_ticks(1, True)
            if (i + 1) % args.every == 0:
                sim.sync()
                st = sim.stats()
                seq.append([i + 1, st["total_pair_count"], st["total_flagged_count"]])
        print(f"N=1: {args.ticks} ticks in {time.perf_counter() - t0:.2f} s; last {seq[-1]}", flush=True)
        json.dump(seq, open(n1_path, "w"))
        return 0

    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)
    hist, ncx, ncy = S.global_row_histogram(M, m, total, 42, 10.0, w["box"], threads=S.host_threads(world))
    splits = S.balanced_splits(hist, world)
    lo, hi = int(splits[rank]), int(splits[rank + 1])
    ents, gids = S.collect_band(M, m, total, 42, 10.0, lo, hi, w["box"], threads=S.host_threads(world))
    max_row = int(hist.max())
    mig_cap = halo_cap = max(4096, 3 * max_row)
    capacity = int(ents.shape[0] * 1.3) + 4 * (mig_cap + halo_cap) + 1024
    ref = json.load(open(n1_path)) if os.path.exists(n1_path) else None
    first = None
    for rep, mode in enumerate(args.modes.split(",")):
        with torch.cuda.stream(stream):
            sim = M.Simulation(m, ents, radius=10.0, device=local_rank, stream=stream.cuda_stream, capacity=capacity)
            sim.shard_enable(gids, mig_cap, halo_cap)
            engine = S.CudaShardEngine(M, sim)
            sh = S.ShardedSimulation(engine, rank, world, splits, ncy, dist, torch, device, mig_cap, halo_cap, M.shard_buffer_bytes(mig_cap, halo_cap),
                                     S.REBALANCE_EVERY, mode)
            sim.dispatch(2)
            for _ in range(args.preroll):
                sh.tick(False)
            seq = []
            t0 = time.perf_counter()
            for i in range(args.ticks):
                sh.tick(True)
                if (i + 1) % args.every == 0:
                    sim.sync()
                    st = sim.stats()
                    mine = torch.tensor([st["total_pair_count"], st["total_flagged_count"], st["entity_count"]], dtype=torch.int64, device=device)
                    allr = [torch.zeros_like(mine) for _ in range(world)]
                    dist.all_gather(allr, mine)
                    per = [[int(x) for x in t.tolist()] for t in allr]
                    seq.append([i + 1, sum(p[0] for p in per), sum(p[1] for p in per), per])
            dt = time.perf_counter() - t0
            torch.cuda.synchronize()
            dist.barrier()
        if rank == 0:
            msg = f"N={world} rep {rep} mode {sh.exchange}{args.tag}: {args.ticks} ticks in {dt:.2f} s; last {seq[-1][:3]}"
            if ref:
                bad = [(a[0], a[1] - b[1], a[2] - b[2]) for a, b in zip(seq, ref) if a[1] != b[1] or a[2] != b[2]]
                msg += f" | vs N=1: {'OK' if not bad else 'FIRST MISMATCH (tick, d_pairs, d_flagged) ' + str(bad[0]) + ' windows differing ' + str(len(bad))}"
                if bad:  # how the difference grows: one event or many
                    msg += " | growth " + str([b for b in bad[:: max(1, len(bad) // 8)]][:10])
            if first is not None:
                badr = [(a[0], [[x - y for x, y in zip(pa, pb)] for pa, pb in zip(a[3], b[3])]) for a, b in zip(seq, first) if a[3] != b[3]]
                msg += f" | vs rep 0 per rank: {'same' if not badr else badr[0]}"
            print(msg, flush=True)
            json.dump(seq, open(f"gpurun_out/stress_n{world}_rep{rep}_{sh.exchange}{args.tag}.json", "w"))
        if first is None:
            first = seq
        del sh, engine, sim
        torch.cuda.synchronize()
        dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
