"""Race hunt on ONE GPU: an unsharded handle, a second unsharded handle and W sharded bands (peer-memory exchange between handles of this
process) step through the same ticks; after every tick the flagged / pair counts must agree.  At the first ticks that differ the
per-entity flags are read back and the entities whose colour differs are printed with their cell row, the band splits and the truth
(numpy brute force over the whole population)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ticks", type=int, default=1500)
    ap.add_argument("--preroll", type=int, default=280)
    ap.add_argument("--workload", default="munich_10m_collisions")
    ap.add_argument("--entities", type=int, default=None)
    ap.add_argument("--bands", type=int, default=4)
    ap.add_argument("--events", type=int, default=4)
    ap.add_argument("--force", type=int, default=-1, help="treat this tick as a mismatch (exercises the dump)")
    args = ap.parse_args()
    import torch

    import movement_sim_b200 as M
    from bench import build_population, build_workload
    from movement_sim_b200 import sharding as S

    M.lib()
    w, m = build_workload(M, args.workload, args.entities)
    total, radius, box, world = w["entities"], 10.0, w["box"], args.bands
    stream = torch.cuda.Stream()
    ents_all = build_population(M, m, total, box)
    hist, ncx, ncy = S.global_row_histogram(M, m, total, 42, radius, box)
    splits = S.balanced_splits(hist, world)
    cap = max(4096, 3 * int(hist.max()))
    with torch.cuda.stream(stream):
        A = M.Simulation(m, ents_all, radius=radius, stream=stream.cuda_stream)
        A2 = M.Simulation(m, ents_all, radius=radius, stream=stream.cuda_stream)
        sims = []
        for r in range(world):
            e, g = S.collect_band(M, m, total, 42, radius, int(splits[r]), int(splits[r + 1]), box)
            s = M.Simulation(m, e, radius=radius, stream=stream.cuda_stream, capacity=int(e.shape[0] * 1.3) + 8 * cap + 1024)
            s.shard_enable(g, cap, cap)
            sims.append(s)
        arenas = [s.shard_p2p_create()[1] for s in sims]
        for r, s in enumerate(sims):
            s.shard_p2p_connect_local(arenas[r - 1] if r > 0 else None, arenas[r + 1] if r + 1 < world else None)
        for s in [A, A2] + sims:
            s.dispatch(2)

        def tick(collide):
            A.enqueue_ticks(1, collide)
            A2.enqueue_ticks(1, collide)
            for r, s in enumerate(sims):
                s.shard_p2p_move_pack(int(splits[r]), int(splits[r + 1]))
            for s in sims:
                s.shard_p2p_integrate()
                if collide:
                    s.enqueue_collide()

        for _ in range(args.preroll):
            tick(False)
        events = 0
        n_a2 = n_b = 0
        for t in range(args.ticks):
            tick(True)
            sa, sa2 = A.stats(), A2.stats()
            sb = [s.stats() for s in sims]
            fa, fa2, fb = sa["last_flagged_count"], sa2["last_flagged_count"], sum(x["last_flagged_count"] for x in sb)
            pa, pa2, pb = sa["last_pair_count"], sa2["last_pair_count"], sum(x["last_pair_count"] for x in sb)
            if (fa, pa) != (fa2, pa2):
                n_a2 += 1
            if (fa, pa) != (fb, pb):
                n_b += 1
            if (fa, pa) == (fa2, pa2) == (fb, pb) and t != args.force:
                continue
            print(f"tick {t}: unsharded {fa}/{pa}  unsharded#2 {fa2}/{pa2}  bands {fb}/{pb}  per band flagged {[x['last_flagged_count'] for x in sb]}", flush=True)
            if events >= args.events:
                continue
            events += 1
            flags_a, flags_a2 = A.read_collision_flags(), A2.read_collision_flags()
            pos = A.read_entities()["pos"].astype(np.float64)
            rows, _, _ = M.grid_rows(m.width, m.height, radius, pos.astype(np.float32))
            flags_b = np.zeros(total, dtype=np.uint8)
            owner = np.full(total, -1, dtype=np.int32)
            for r, s in enumerate(sims):
                g = s.shard_read_gids()
                flags_b[g] = s.read_collision_flags()[: g.shape[0]]
                owner[g] = r
            print(f"   flags read back: unsharded {np.unique(flags_a, return_counts=True)} unsharded#2 {np.unique(flags_a2, return_counts=True)} bands {np.unique(flags_b, return_counts=True)}", flush=True)
            for name, other in (("unsharded#2", flags_a2), ("bands", flags_b)):
                bad = np.nonzero(flags_a != other)[0]
                for gid in (bad[:6] if t != args.force else [0, 1]):
                    d = np.hypot(pos[:, 0] - pos[gid, 0], pos[:, 1] - pos[gid, 1])
                    near = np.nonzero((d < radius * 1.001) & (np.arange(total) != gid))[0]
                    print(f"   {name}: gid {gid} pos {pos[gid].tolist()} row {int(rows[gid])} owner band {int(owner[gid])} splits {splits.tolist()} "
                          f"flag unsharded {int(flags_a[gid])} other {int(other[gid])}; neighbours within r (float64): "
                          f"{[(int(k), round(float(d[k]), 4), int(rows[k]), int(owner[k])) for k in near[:6]]}", flush=True)
        print(f"done: {args.ticks} ticks, unsharded#2 differed on {n_a2}, bands differed on {n_b}", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
