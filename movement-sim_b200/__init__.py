"""movement_sim_b200 — Python host binding of libmsim_cuda.so (the C ABI in include/msim.h).

This is plumbing for tests and bench.py: every compute call goes straight through the C ABI into the
hand-written sm_100a kernels under csrc/.  There is no CPU fallback and no other backend: if the
shared library is missing, importing this module raises; if no B200-class GPU is present, creating a
Simulation raises MsimError(MSIM_ERR_CUDA).

The directory is called ``movement-sim_b200`` (not a Python identifier); ``import movement_sim_b200``
works through the one-line shim module at the repository root.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmsim_cuda.so")

# ---- data model (byte-identical to include/msim.h) ------------------------------------------------
ENTITY_DTYPE = np.dtype(
    [
        ("color", "<f4", (4,)),
        ("rand_state", "<u4", (4,)),
        ("pos", "<f4", (2,)),
        ("target", "<f4", (2,)),
        ("direction", "<f4", (2,)),
        ("road_index", "<u4"),
        ("initialized", "<u4"),
    ]
)
ROAD_DTYPE = np.dtype(
    [
        ("start_pos", "<f4", (2,)),
        ("start_index", "<u4"),
        ("start_count", "<u4"),
        ("end_pos", "<f4", (2,)),
        ("end_index", "<u4"),
        ("end_count", "<u4"),
    ]
)
QUADTREE_NODE_DTYPE = np.dtype(
    [
        ("acquire_lock", "<i4"),
        ("write_lock", "<i4"),
        ("reader_lock", "<i4"),
        ("offset_x", "<f4"),
        ("offset_y", "<f4"),
        ("width", "<f4"),
        ("height", "<f4"),
        ("content_type", "<u4"),
        ("entity_count", "<u4"),
        ("first", "<u4"),
        ("prev_node_index", "<u4"),
        ("next_tl", "<u4"),
        ("next_tr", "<u4"),
        ("next_bl", "<u4"),
        ("next_br", "<u4"),
        ("padding", "<u4"),
    ]
)
assert ENTITY_DTYPE.itemsize == 64 and ROAD_DTYPE.itemsize == 32 and QUADTREE_NODE_DTYPE.itemsize == 64

MSIM_ABI_VERSION = 2
(MSIM_OK, MSIM_ERR_INVALID, MSIM_ERR_CUDA, MSIM_ERR_OOM, MSIM_ERR_UNSUPPORTED, MSIM_ERR_IO, MSIM_ERR_PARSE,
 MSIM_ERR_CAPACITY, MSIM_ERR_INTERNAL) = range(9)
FLAG_NO_COLLISIONS = 1 << 0
FLAG_NO_PAIR_COUNT = 1 << 1
FLAG_NO_QUADTREE = 1 << 2
FLAG_SORT_COUNTING = 1 << 3
FLAG_NO_REORDER = 1 << 4
FLAG_SORT_ONESWEEP = 1 << 5

# every symbol include/msim.h declares (tests/test_abi.py checks the library exports exactly these)
ABI_SYMBOLS = [
    "msim_create", "msim_destroy", "msim_last_error", "msim_status_string", "msim_upload_entities",
    "msim_dispatch", "msim_enqueue_move", "msim_enqueue_collide", "msim_enqueue_ticks", "msim_sync",
    "msim_set_stream", "msim_read_entities", "msim_read_positions", "msim_read_collision_flags",
    "msim_read_quadtree_nodes", "msim_read_debug", "msim_snapshot_begin", "msim_snapshot_poll", "msim_snapshot_end", "msim_get_stats", "msim_get_device_view",
    "msim_profile_begin", "msim_profile_end",
    "msim_map_load_json", "msim_map_save_json", "msim_map_generate_city", "msim_map_generate_grid",
    "msim_map_free", "msim_map_width", "msim_map_height", "msim_map_road_count",
    "msim_map_connection_count", "msim_map_roads", "msim_map_connections", "msim_map_last_error",
    "msim_entities_init", "msim_calc_node_count", "msim_abi_version", "msim_quadtree_from_positions", "msim_entities_init_roads",
    # include/msim_shard.h
    "msim_shard_buffer_bytes", "msim_shard_enable", "msim_shard_pack", "msim_shard_move_pack", "msim_shard_p2p_create", "msim_shard_p2p_connect",
    "msim_shard_p2p_connect_local", "msim_shard_p2p_move_pack", "msim_shard_p2p_integrate", "msim_shard_integrate", "msim_shard_integrate_async",
    "msim_shard_counts", "msim_shard_read_gids",
    "msim_shard_row_histogram", "msim_grid_rows", "msim_grid_params",
    # include/msim_mapgen.h
    "msim_map_from_geojson", "msim_map_save_binary", "msim_map_load_binary", "msim_map_load", "msim_map_from_arrays", "msim_haversine_m",
]
MAPGEN_EXACT_TRAVERSAL = 1 << 0
MAPGEN_NO_DUPLICATE_END = 1 << 1


class MapgenStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("features", "line_strings", "road_pieces", "skipped_zero", "skipped_duplicate", "connected",
                                          "coordinates")] + [
        (n, C.c_double) for n in ("min_dist_lat", "max_dist_lat", "min_dist_long", "max_dist_long", "ref_lat", "ref_long")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class MsimError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"msim status {status}: {message}")
        self.status = status
        self.message = message


class PushConsts(C.Structure):
    _pack_ = 1
    _fields_ = [
        ("world_size_x", C.c_float),
        ("world_size_y", C.c_float),
        ("node_count", C.c_uint32),
        ("max_depth", C.c_uint32),
        ("entity_node_cap", C.c_uint32),
        ("collision_radius", C.c_float),
        ("tick", C.c_uint32),
    ]


assert C.sizeof(PushConsts) == 28


class _Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("device", C.c_int32),
        ("flags", C.c_uint32),
        ("reserved0", C.c_uint32),
        ("world_w", C.c_float),
        ("world_h", C.c_float),
        ("collision_radius", C.c_float),
        ("quadtree_max_depth", C.c_uint32),
        ("quadtree_node_cap", C.c_uint32),
        ("reserved1", C.c_uint32),
        ("roads", C.c_void_p),
        ("road_count", C.c_uint64),
        ("connections", C.c_void_p),
        ("connection_count", C.c_uint64),
        ("entities", C.c_void_p),
        ("entity_count", C.c_uint64),
        ("entity_capacity", C.c_uint64),
        ("cuda_stream", C.c_void_p),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("entity_count", C.c_uint64),
        ("move_passes", C.c_uint64),
        ("collide_passes", C.c_uint64),
        ("last_pair_count", C.c_uint64),
        ("total_pair_count", C.c_uint64),
        ("last_flagged_count", C.c_uint64),
        ("kernel_launches", C.c_uint64),
        ("grid_cells_x", C.c_uint32),
        ("grid_cells_y", C.c_uint32),
        ("key_bits", C.c_uint32),
        ("sort_passes", C.c_uint32),
        ("cell_size", C.c_float),
        ("reorders", C.c_uint32),
        ("total_flagged_count", C.c_uint64),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class DeviceView(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("target", C.c_void_p), ("road", C.c_void_p), ("rng", C.c_void_p), ("count", C.c_uint64),
                ("ext_id", C.c_void_p)]


class KernelTime(C.Structure):
    _fields_ = [("name", C.c_char * 24), ("launches", C.c_uint64), ("total_ms", C.c_double)]


_lib = None
_host = None
HOST_LIB_PATH = os.path.join(HERE, "libmsim_host.so")
# the C-ABI entry points that never touch a GPU (csrc/host_map.cpp, csrc/host_mapgen.cpp): map loading / generation and entity initialisation
HOST_SYMBOLS = ["msim_map_load_json", "msim_map_save_json", "msim_map_generate_city", "msim_map_generate_grid", "msim_map_free", "msim_map_width", "msim_map_height", "msim_map_road_count", "msim_map_connection_count", "msim_map_roads", "msim_map_connections", "msim_map_last_error", "msim_entities_init", "msim_entities_init_roads", "msim_calc_node_count", "msim_abi_version", "msim_map_from_geojson", "msim_map_save_binary", "msim_map_load_binary", "msim_map_load", "msim_map_from_arrays", "msim_haversine_m"]


def _signatures():
    vp, u64, u32, i32, f32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_float
    return {
        "msim_create": (i32, [C.POINTER(_Config), C.POINTER(vp)]),
        "msim_destroy": (None, [vp]),
        "msim_last_error": (C.c_char_p, [vp]),
        "msim_status_string": (C.c_char_p, [i32]),
        "msim_upload_entities": (i32, [vp, vp, u64]),
        "msim_dispatch": (i32, [vp, C.POINTER(PushConsts)]),
        "msim_enqueue_move": (i32, [vp]),
        "msim_enqueue_collide": (i32, [vp]),
        "msim_enqueue_ticks": (i32, [vp, u32, i32]),
        "msim_sync": (i32, [vp]),
        "msim_set_stream": (i32, [vp, vp]),
        "msim_read_entities": (i32, [vp, vp, u64]),
        "msim_read_positions": (i32, [vp, vp, u64]),
        "msim_read_collision_flags": (i32, [vp, vp, u64]),
        "msim_read_quadtree_nodes": (i32, [vp, vp, u64, C.POINTER(u64)]),
        "msim_read_debug": (i32, [vp, vp]),
        "msim_snapshot_begin": (i32, [vp]),
        "msim_snapshot_poll": (i32, [vp, C.POINTER(i32)]),
        "msim_snapshot_end": (i32, [vp, C.POINTER(vp), C.POINTER(u64)]),
        "msim_get_stats": (i32, [vp, C.POINTER(Stats)]),
        "msim_get_device_view": (i32, [vp, C.POINTER(DeviceView)]),
        "msim_profile_begin": (i32, [vp]),
        "msim_profile_end": (i32, [vp, C.POINTER(KernelTime), u32, C.POINTER(u32)]),
        "msim_shard_buffer_bytes": (u64, [u32, u32]),
        "msim_shard_enable": (i32, [vp, vp, u64, u32, u32]),
        "msim_shard_pack": (i32, [vp, u32, u32, vp, vp]),
        "msim_shard_move_pack": (i32, [vp, u32, u32, vp, vp]),
        "msim_shard_p2p_create": (i32, [vp, vp, C.POINTER(vp)]),
        "msim_shard_p2p_connect": (i32, [vp, vp, vp]),
        "msim_shard_p2p_connect_local": (i32, [vp, vp, vp]),
        "msim_shard_p2p_move_pack": (i32, [vp, u32, u32]),
        "msim_shard_p2p_integrate": (i32, [vp]),
        "msim_shard_integrate": (i32, [vp, vp, vp, C.POINTER(u64), C.POINTER(u64)]),
        "msim_shard_integrate_async": (i32, [vp, vp, vp]),
        "msim_shard_counts": (i32, [vp, C.POINTER(u64), C.POINTER(u64)]),
        "msim_shard_read_gids": (i32, [vp, vp, u64]),
        "msim_shard_row_histogram": (i32, [vp, vp, u32]),
        "msim_grid_rows": (i32, [f32, f32, f32, vp, u64, vp, C.POINTER(u32), C.POINTER(u32)]),
        "msim_grid_params": (i32, [f32, f32, f32, C.POINTER(f32), C.POINTER(f32), C.POINTER(u32), C.POINTER(u32)]),
        "msim_map_load_json": (i32, [C.c_char_p, C.POINTER(vp)]),
        "msim_map_save_json": (i32, [vp, C.c_char_p]),
        "msim_map_generate_city": (i32, [f32, f32, f32, f32, f32, u64, C.POINTER(vp)]),
        "msim_map_generate_grid": (i32, [u32, u32, f32, C.POINTER(vp)]),
        "msim_map_free": (None, [vp]),
        "msim_map_width": (f32, [vp]),
        "msim_map_height": (f32, [vp]),
        "msim_map_road_count": (u64, [vp]),
        "msim_map_connection_count": (u64, [vp]),
        "msim_map_roads": (vp, [vp]),
        "msim_map_connections": (vp, [vp]),
        "msim_map_last_error": (C.c_char_p, []),
        "msim_entities_init": (i32, [vp, u64, u64, u64, vp, vp]),
        "msim_entities_init_roads": (i32, [vp, u64, u64, u64, vp, vp]),
        "msim_calc_node_count": (u64, [u32]),
        "msim_quadtree_from_positions": (i32, [vp, u64, f32, f32, u32, u32, vp, u64, C.POINTER(u64)]),
        "msim_abi_version": (u32, []),
        "msim_map_from_geojson": (i32, [C.c_char_p, u32, C.POINTER(vp), C.POINTER(MapgenStats)]),
        "msim_map_save_binary": (i32, [vp, C.c_char_p]),
        "msim_map_load_binary": (i32, [C.c_char_p, C.POINTER(vp)]),
        "msim_map_load": (i32, [C.c_char_p, C.POINTER(vp)]),
        "msim_map_from_arrays": (i32, [f32, f32, vp, u64, vp, u64, C.POINTER(vp)]),
        "msim_haversine_m": (C.c_double, [C.c_double] * 4),
    }


def _bind(L, names):
    sigs = _signatures()
    for name in names:
        fn = getattr(L, name)
        fn.restype, fn.argtypes = sigs[name]
    return L


def lib():
    """Loads libmsim_cuda.so; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C {HERE}` (or __graft_entry__.build()). "
            "movement_sim_b200 has no CPU or alternative backend."
        )
    _lib = _bind(C.CDLL(LIB_PATH), list(_signatures()))
    return _lib


def host_lib():
    """The host-only part of the C ABI (HOST_SYMBOLS) from libmsim_host.so: the same objects libmsim_cuda.so links, without the CUDA
    runtime, so that a process that only prepares inputs (bench.py --impl reference) never maps the CUDA library."""
    global _host
    if _host is not None:
        return _host
    if not os.path.exists(HOST_LIB_PATH):
        raise ImportError(f"{HOST_LIB_PATH} is missing: build it with `make -C {HERE}` (or __graft_entry__.build()).")
    _host = _bind(C.CDLL(HOST_LIB_PATH), HOST_SYMBOLS)
    return _host


# ---- map -------------------------------------------------------------------------------------
class Map:
    """Road graph: `roads` (ROAD_DTYPE) + flat `connections` (uint32) + world size.
    Mirrors sim::Map (/root/reference/src/sim/Map.hpp:27-47) minus the render-only roadPieces."""

    def __init__(self, width: float, height: float, roads: np.ndarray, connections: np.ndarray):
        self.width = float(np.float32(width))
        self.height = float(np.float32(height))
        self.roads = np.ascontiguousarray(roads, dtype=ROAD_DTYPE)
        self.connections = np.ascontiguousarray(connections, dtype=np.uint32)

    @classmethod
    def _take(cls, handle) -> "Map":
        L = host_lib()
        try:
            nr, nc = L.msim_map_road_count(handle), L.msim_map_connection_count(handle)
            roads = np.empty(nr, dtype=ROAD_DTYPE)
            conns = np.empty(nc, dtype=np.uint32)
            if nr:
                C.memmove(roads.ctypes.data, L.msim_map_roads(handle), nr * 32)
            if nc:
                C.memmove(conns.ctypes.data, L.msim_map_connections(handle), nc * 4)
            return cls(L.msim_map_width(handle), L.msim_map_height(handle), roads, conns)
        finally:
            L.msim_map_free(handle)

    @classmethod
    def load_json(cls, path: str) -> "Map":
        """Map::load_from_file (/root/reference/src/sim/Map.cpp:28-150)."""
        L = host_lib()
        h = C.c_void_p()
        rc = L.msim_map_load_json(os.fsencode(path), C.byref(h))
        if rc != MSIM_OK:
            raise MsimError(rc, L.msim_map_last_error().decode())
        return cls._take(h)

    @classmethod
    def load(cls, path: str) -> "Map":
        """Any supported map file: binary cache, GeoJSON export, or the reference's map JSON (msim_map_load)."""
        L = host_lib()
        h = C.c_void_p()
        rc = L.msim_map_load(os.fsencode(path), C.byref(h))
        if rc != MSIM_OK:
            raise MsimError(rc, L.msim_map_last_error().decode())
        return cls._take(h)

    @classmethod
    def from_geojson(cls, path: str, flags: int = 0, with_stats: bool = False):
        """/root/reference/map/generate_map.py as a native pipeline (include/msim_mapgen.h)."""
        L = host_lib()
        h = C.c_void_p()
        st = MapgenStats()
        rc = L.msim_map_from_geojson(os.fsencode(path), flags, C.byref(h), C.byref(st))
        if rc != MSIM_OK:
            raise MsimError(rc, L.msim_map_last_error().decode())
        m = cls._take(h)
        return (m, st.as_dict()) if with_stats else m

    def save_binary(self, path: str) -> None:
        """Binary map cache (msim_map_save_binary)."""
        L = host_lib()
        h = C.c_void_p()
        rc = L.msim_map_from_arrays(self.width, self.height, self.roads.ctypes.data, self.roads.shape[0], self.connections.ctypes.data,
                                    self.connections.shape[0], C.byref(h))
        if rc != MSIM_OK:
            raise MsimError(rc, L.msim_map_last_error().decode())
        try:
            rc = L.msim_map_save_binary(h, os.fsencode(path))
            if rc != MSIM_OK:
                raise MsimError(rc, L.msim_map_last_error().decode())
        finally:
            L.msim_map_free(h)

    @classmethod
    def city(cls, world_w: float = 29007.4609, world_h: float = 16463.7656, spacing: float = 35.0, jitter: float = 0.3,
             drop_prob: float = 0.12, seed: int = 2022) -> "Map":
        """Seeded stand-in for the missing munich.json (world size from
        /root/reference/shader_validation/src/main.cpp:155)."""
        L = host_lib()
        h = C.c_void_p()
        rc = L.msim_map_generate_city(world_w, world_h, spacing, jitter, drop_prob, seed, C.byref(h))
        if rc != MSIM_OK:
            raise MsimError(rc, L.msim_map_last_error().decode())
        return cls._take(h)

    @classmethod
    def grid(cls, nx: int, ny: int, spacing: float = 20.0) -> "Map":
        L = host_lib()
        h = C.c_void_p()
        rc = L.msim_map_generate_grid(nx, ny, spacing, C.byref(h))
        if rc != MSIM_OK:
            raise MsimError(rc, L.msim_map_last_error().decode())
        return cls._take(h)

    def init_entities(self, count: int, seed: int = 42, box=None) -> np.ndarray:
        """Simulator::add_entities (/root/reference/src/sim/Simulator.cpp:114-129), seeded."""
        L = host_lib()
        out = np.zeros(count, dtype=ENTITY_DTYPE)
        boxp = None
        if box is not None:
            box = np.ascontiguousarray(box, dtype=np.float32)
            assert box.shape == (4,)
            boxp = box.ctypes.data
        rc = L.msim_entities_init(self.roads.ctypes.data, self.roads.shape[0], count, seed, boxp, out.ctypes.data)
        if rc != MSIM_OK:
            raise MsimError(rc, L.msim_map_last_error().decode())
        return out


def _init_road_indices(self, count: int, seed: int = 42, box=None) -> np.ndarray:
    """road_index of every entity init_entities(count, seed, box) would create, without creating them (msim_entities_init_roads)."""
    out = np.zeros(count, dtype=np.uint32)
    boxp = None
    if box is not None:
        box = np.ascontiguousarray(box, dtype=np.float32)
        assert box.shape == (4,)
        boxp = box.ctypes.data
    rc = host_lib().msim_entities_init_roads(self.roads.ctypes.data, self.roads.shape[0], count, seed, boxp, out.ctypes.data)
    if rc != MSIM_OK:
        raise MsimError(rc, host_lib().msim_map_last_error().decode())
    return out


Map.init_road_indices = _init_road_indices


def grid_rows(world_w: float, world_h: float, radius: float, xy: np.ndarray):
    """Cell row of every position, exactly as the device computes it; returns (rows, cells_x, cells_y)."""
    xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
    rows = np.empty(xy.shape[0], dtype=np.uint32)
    cx, cy = C.c_uint32(), C.c_uint32()
    rc = lib().msim_grid_rows(world_w, world_h, radius, xy.ctypes.data, xy.shape[0], rows.ctypes.data, C.byref(cx), C.byref(cy))
    if rc != MSIM_OK:
        raise MsimError(rc, "msim_grid_rows")
    return rows, cx.value, cy.value


def quadtree_from_positions(xy: np.ndarray, world_w: float, world_h: float, max_depth: int = 8, node_cap: int = 10) -> np.ndarray:
    """Display quadtree of the given positions, built on the host by the code behind msim_read_quadtree_nodes."""
    xy = np.ascontiguousarray(xy, dtype=np.float32).reshape(-1, 2)
    cap = calc_node_count(max_depth)
    out = np.zeros(cap, dtype=QUADTREE_NODE_DTYPE)
    n = C.c_uint64()
    rc = lib().msim_quadtree_from_positions(xy.ctypes.data, xy.shape[0], world_w, world_h, max_depth, node_cap, out.ctypes.data, cap, C.byref(n))
    if rc != MSIM_OK:
        raise MsimError(rc, "msim_quadtree_from_positions")
    return out[: n.value]


def grid_params(world_w: float, world_h: float, radius: float) -> dict:
    """The device's neighbour grid for this world and radius: inv_cell, hit_threshold (binary32), cells_x, cells_y."""
    inv, thr, cx, cy = C.c_float(), C.c_float(), C.c_uint32(), C.c_uint32()
    rc = lib().msim_grid_params(world_w, world_h, radius, C.byref(inv), C.byref(thr), C.byref(cx), C.byref(cy))
    if rc != MSIM_OK:
        raise MsimError(rc, "msim_grid_params")
    return {"inv_cell": np.float32(inv.value), "hit_threshold": np.float32(thr.value), "cells_x": cx.value, "cells_y": cy.value}


def shard_buffer_bytes(migrant_capacity: int, halo_capacity: int) -> int:
    return int(lib().msim_shard_buffer_bytes(migrant_capacity, halo_capacity))


def calc_node_count(depth: int) -> int:
    return int(host_lib().msim_calc_node_count(depth))


# ---- simulation handle -----------------------------------------------------------------------
class Simulation:
    """One msim_handle: the resident entity population of one GPU."""

    def __init__(self, m: Map, entities: np.ndarray, radius: float = 10.0, device: int = 0, flags: int = 0,
                 stream: int | None = None, capacity: int = 0, quadtree_depth: int = 8, quadtree_cap: int = 10):
        L = lib()
        ents = np.ascontiguousarray(entities, dtype=ENTITY_DTYPE)
        self.map = m
        self.radius = float(np.float32(radius))
        self._cfg = _Config(
            MSIM_ABI_VERSION, device, flags, 0, m.width, m.height, self.radius, quadtree_depth, quadtree_cap, 0,
            m.roads.ctypes.data, m.roads.shape[0], m.connections.ctypes.data if m.connections.size else None, m.connections.shape[0],
            ents.ctypes.data if ents.size else None, ents.shape[0], capacity, stream,
        )
        self._h = C.c_void_p()
        rc = L.msim_create(C.byref(self._cfg), C.byref(self._h))
        if rc != MSIM_OK:
            raise MsimError(rc, L.msim_last_error(None).decode())
        self.count = ents.shape[0]
        self._sharded_async = False
        self.quadtree_depth, self.quadtree_cap = quadtree_depth, quadtree_cap

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().msim_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != MSIM_OK:
            raise MsimError(rc, lib().msim_last_error(self._h).decode())

    # -- dispatch
    def push_consts(self, tick: int) -> PushConsts:
        """The reference's push constants (/root/reference/src/sim/Simulator.cpp:94-101)."""
        return PushConsts(self.map.width, self.map.height, calc_node_count(self.quadtree_depth), self.quadtree_depth,
                          self.quadtree_cap, self.radius, tick)

    def dispatch(self, tick: int):
        pc = self.push_consts(tick)
        self._check(lib().msim_dispatch(self._h, C.byref(pc)))

    def enqueue_move(self):
        self._check(lib().msim_enqueue_move(self._h))

    def enqueue_collide(self):
        self._check(lib().msim_enqueue_collide(self._h))

    def enqueue_ticks(self, sim_ticks: int, collisions: bool):
        self._check(lib().msim_enqueue_ticks(self._h, sim_ticks, 1 if collisions else 0))

    def sync(self):
        self._check(lib().msim_sync(self._h))

    def set_stream(self, stream: int | None):
        self._check(lib().msim_set_stream(self._h, stream))

    # -- transfers
    def upload(self, entities: np.ndarray):
        ents = np.ascontiguousarray(entities, dtype=ENTITY_DTYPE)
        self._check(lib().msim_upload_entities(self._h, ents.ctypes.data, ents.shape[0]))
        self.count = ents.shape[0]

    def upload_ptr(self, ptr: int, count: int):
        self._check(lib().msim_upload_entities(self._h, ptr, count))
        self.count = count

    def read_entities(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None and self._sharded_async:
            self.shard_counts()
        if out is None:
            out = np.empty(self.count, dtype=ENTITY_DTYPE)
        assert out.dtype == ENTITY_DTYPE and out.flags.c_contiguous
        self._check(lib().msim_read_entities(self._h, out.ctypes.data, out.shape[0]))
        return out

    def read_entities_ptr(self, ptr: int, count: int):
        self._check(lib().msim_read_entities(self._h, ptr, count))

    def snapshot_begin(self):
        """Asynchronous readback: image of the current state, copied to pinned host memory while later ticks run."""
        self._check(lib().msim_snapshot_begin(self._h))

    def snapshot_ready(self) -> bool:
        ready = C.c_int(0)
        self._check(lib().msim_snapshot_poll(self._h, C.byref(ready)))
        return bool(ready.value)

    def snapshot_end(self, copy: bool = True) -> np.ndarray:
        """Entities of the last snapshot_begin.  copy=False returns a view of the library's pinned buffer (valid until the
        begin after the next one)."""
        ptr, count = C.c_void_p(), C.c_uint64()
        self._check(lib().msim_snapshot_end(self._h, C.byref(ptr), C.byref(count)))
        if count.value == 0:
            return np.empty(0, dtype=ENTITY_DTYPE)
        view = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count.value * 64,)).view(ENTITY_DTYPE)
        return view.copy() if copy else view

    def read_positions(self) -> np.ndarray:
        out = np.empty((self.count, 2), dtype=np.float32)
        self._check(lib().msim_read_positions(self._h, out.ctypes.data, self.count))
        return out

    def read_positions_ptr(self, ptr: int, count: int):
        self._check(lib().msim_read_positions(self._h, ptr, count))

    def read_collision_flags_ptr(self, ptr: int, count: int):
        self._check(lib().msim_read_collision_flags(self._h, ptr, count))

    def read_collision_flags(self) -> np.ndarray:
        out = np.empty(self.count, dtype=np.uint8)
        self._check(lib().msim_read_collision_flags(self._h, out.ctypes.data, self.count))
        return out

    def read_quadtree_nodes(self) -> np.ndarray:
        cap = calc_node_count(self.quadtree_depth)
        out = np.zeros(cap, dtype=QUADTREE_NODE_DTYPE)
        n = C.c_uint64()
        self._check(lib().msim_read_quadtree_nodes(self._h, out.ctypes.data, cap, C.byref(n)))
        return out[: n.value]

    def read_debug(self) -> np.ndarray:
        out = np.zeros(10, dtype=np.uint32)
        self._check(lib().msim_read_debug(self._h, out.ctypes.data))
        return out

    def stats(self) -> dict:
        st = Stats()
        self._check(lib().msim_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    # -- multi-GPU sharding (include/msim_shard.h)
    def shard_enable(self, gids: np.ndarray, migrant_capacity: int, halo_capacity: int):
        gids = np.ascontiguousarray(gids, dtype=np.uint32)
        self._check(lib().msim_shard_enable(self._h, gids.ctypes.data, gids.shape[0], migrant_capacity, halo_capacity))

    def shard_pack(self, row_lo: int, row_hi: int, send_down_ptr: int | None, send_up_ptr: int | None):
        self._check(lib().msim_shard_pack(self._h, row_lo, row_hi, send_down_ptr, send_up_ptr))

    def shard_move_pack(self, row_lo: int, row_hi: int, send_down_ptr: int | None, send_up_ptr: int | None):
        """Move pass and shard pack fused into one kernel (msim_shard.h)."""
        self._check(lib().msim_shard_move_pack(self._h, row_lo, row_hi, send_down_ptr, send_up_ptr))

    def shard_integrate(self, recv_down_ptr: int | None, recv_up_ptr: int | None):
        owned, ghosts = C.c_uint64(), C.c_uint64()
        self._check(lib().msim_shard_integrate(self._h, recv_down_ptr, recv_up_ptr, C.byref(owned), C.byref(ghosts)))
        self.count = owned.value
        return owned.value, ghosts.value

    # -- peer-memory exchange (msim_shard.h): no collective call per tick
    def shard_p2p_create(self) -> tuple[bytes, int]:
        """Returns (CUDA IPC handle of this handle's receive arena, its device pointer)."""
        buf = C.create_string_buffer(64)
        arena = C.c_void_p()
        self._check(lib().msim_shard_p2p_create(self._h, buf, C.byref(arena)))
        return bytes(buf.raw), int(arena.value or 0)

    def shard_p2p_connect(self, down_handle: bytes | None, up_handle: bytes | None):
        self._check(lib().msim_shard_p2p_connect(self._h, down_handle, up_handle))

    def shard_p2p_connect_local(self, down_arena: int | None, up_arena: int | None):
        self._check(lib().msim_shard_p2p_connect_local(self._h, down_arena, up_arena))

    def shard_p2p_move_pack(self, row_lo: int, row_hi: int):
        self._check(lib().msim_shard_p2p_move_pack(self._h, row_lo, row_hi))

    def shard_p2p_integrate(self):
        self._sharded_async = True
        self._check(lib().msim_shard_p2p_integrate(self._h))

    def shard_integrate_async(self, recv_down_ptr: int | None, recv_up_ptr: int | None):
        self._sharded_async = True
        self._check(lib().msim_shard_integrate_async(self._h, recv_down_ptr, recv_up_ptr))

    def shard_counts(self):
        owned, ghosts = C.c_uint64(), C.c_uint64()
        self._check(lib().msim_shard_counts(self._h, C.byref(owned), C.byref(ghosts)))
        self.count = owned.value
        return owned.value, ghosts.value

    def shard_read_gids(self) -> np.ndarray:
        self.shard_counts()
        out = np.empty(self.count, dtype=np.uint32)
        self._check(lib().msim_shard_read_gids(self._h, out.ctypes.data, self.count))
        return out

    def shard_row_histogram(self, rows: int) -> np.ndarray:
        out = np.zeros(rows, dtype=np.uint32)
        self._check(lib().msim_shard_row_histogram(self._h, out.ctypes.data, rows))
        return out

    def profile_begin(self):
        self._check(lib().msim_profile_begin(self._h))

    def profile_end(self) -> dict:
        """{kernel name: (launches, total_ms)} since profile_begin()."""
        buf = (KernelTime * 32)()
        n = C.c_uint32()
        self._check(lib().msim_profile_end(self._h, buf, 32, C.byref(n)))
        return {buf[i].name.decode(): (int(buf[i].launches), float(buf[i].total_ms)) for i in range(n.value)}

    def device_view(self) -> DeviceView:
        v = DeviceView()
        self._check(lib().msim_get_device_view(self._h, C.byref(v)))
        return v

    def join(self):
        """Orders the handle's stream behind everything enqueued so far, the library's internal streams included (pass B, an overlapped move
        phase, the pipelined rebuild); no host wait.  An event recorded on the handle's stream after this call marks the end of ALL that work."""
        self.device_view()
