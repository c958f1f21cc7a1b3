"""TEST INFRASTRUCTURE: an oracle-backed stand-in for CudaShardEngine so that the multi-GPU host logic
(movement-sim_b200/sharding.py: banding, exchange protocol, re-balancing) can run on CPU over gloo.
Same contract as the CUDA engine; its own wire format (the orchestration treats buffers as bytes)."""
import numpy as np

WIRE = np.dtype([("ent", "V64"), ("gid", "<u4"), ("pad", "<u4")])  # 72 bytes, like the CUDA record


class OracleShardEngine:
    def __init__(self, O, M, m, ents, gids, radius, migrant_capacity, halo_capacity):
        self.O, self.M, self.m, self.radius = O, M, m, float(radius)
        self.omap = O.OracleMap(m.width, m.height, m.roads.view(O.ROAD_DTYPE), m.connections)
        self.e = np.ascontiguousarray(ents).view(O.ENTITY_DTYPE).copy()
        self.gids = np.asarray(gids, dtype=np.uint32).copy()
        self.mig_cap, self.halo_cap = migrant_capacity, halo_capacity
        self.ghosts = np.zeros((0, 2), dtype=np.float32)
        self.local_ghosts = np.zeros((0, 2), dtype=np.float32)
        self.last_pairs = 0
        self.lo = self.hi = 0

    def _rows(self, xy):
        return self.M.grid_rows(self.m.width, self.m.height, self.radius, xy)[0].astype(np.int64)

    def move(self):
        self.O.move_pass(self.e, self.omap)

    def _write(self, tensor, ents, gids, halo):
        assert len(ents) <= self.mig_cap and len(halo) <= self.halo_cap, "exchange capacity exceeded"
        buf = tensor.numpy()
        buf[:8] = np.array([len(ents), len(halo)], dtype=np.uint32).view(np.uint8)
        rec = np.zeros(len(ents), dtype=WIRE)
        rec["ent"] = ents.view("V64")
        rec["gid"] = gids
        o = 32
        buf[o : o + rec.nbytes] = rec.view(np.uint8)
        o = 32 + self.mig_cap * 72
        h = np.ascontiguousarray(halo, dtype=np.float32)
        buf[o : o + h.nbytes] = h.view(np.uint8).ravel()

    def _read(self, tensor):
        buf = tensor.numpy()
        n_mig, n_halo = (int(v) for v in buf[:8].view(np.uint32))
        rec = buf[32 : 32 + n_mig * 72].view(WIRE)
        ents = rec["ent"].view(self.O.ENTITY_DTYPE).copy() if n_mig else np.zeros(0, dtype=self.O.ENTITY_DTYPE)
        o = 32 + self.mig_cap * 72
        halo = buf[o : o + n_halo * 8].view(np.float32).reshape(-1, 2).copy()
        return ents, rec["gid"].copy(), halo

    def pack(self, lo, hi, send_down, send_up):
        self.lo, self.hi = lo, hi
        rows = self._rows(self.e["pos"])
        down = (rows < lo) if send_down is not None else np.zeros(len(rows), bool)
        up = (rows >= hi) if send_up is not None else np.zeros(len(rows), bool)
        stay = ~(down | up)
        if send_down is not None:
            self._write(send_down, self.e[down], self.gids[down], self.e["pos"][stay & (rows == lo)])
        if send_up is not None:
            self._write(send_up, self.e[up], self.gids[up], self.e["pos"][stay & (rows == hi - 1)])
        self.local_ghosts = self.e["pos"][down | up].copy()
        self.e, self.gids = self.e[stay].copy(), self.gids[stay].copy()

    def integrate(self, recv_down, recv_up):
        ghosts = [self.local_ghosts]
        for t in (recv_down, recv_up):
            if t is not None:
                ents, gids, halo = self._read(t)
                self.e = np.concatenate([self.e, ents])
                self.gids = np.concatenate([self.gids, gids])
                ghosts.append(halo)
        self.ghosts = np.concatenate(ghosts) if ghosts else np.zeros((0, 2), np.float32)
        return len(self.e), len(self.ghosts)

    def _pairs(self, xy):
        a = np.zeros(len(xy), dtype=self.O.ENTITY_DTYPE)
        a["pos"] = xy
        a["initialized"] = 1
        return self.O.collide_pass(a, self.m.width, self.m.height, self.radius), a

    def collide(self):
        n = len(self.e)
        p_all, both = self._pairs(np.concatenate([self.e["pos"], self.ghosts]))
        self.e["color"] = both["color"][:n]  # flags of the owned entities, ghosts included as neighbours
        p_own, _ = self._pairs(self.e["pos"])
        # a cross-band pair is counted by the rank owning its higher-row member: here, ghosts from BELOW
        low = self.ghosts[self._rows(self.ghosts) < self.lo] if len(self.ghosts) else self.ghosts
        p_low_mix, _ = self._pairs(np.concatenate([self.e["pos"], low]))
        p_low_only, _ = self._pairs(low)
        self.last_pairs = p_own + (p_low_mix - p_own - p_low_only)

    def row_histogram(self, rows):
        return np.bincount(self._rows(self.e["pos"]), minlength=rows).astype(np.uint32)

    def stats(self):
        return {"entity_count": len(self.e), "last_pair_count": self.last_pairs,
                "last_flagged_count": int(self.O.collision_flags(self.e).sum())}

    def read_owned(self):
        return self.e.copy(), self.gids.copy()
