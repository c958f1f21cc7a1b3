"""ctypes binding of the CPU ORACLE (oracle/libmsim_oracle.so) and of the compiled reference harness
(oracle/_ref/libref_quadtree.so).

TEST INFRASTRUCTURE ONLY - see oracle/msim_oracle.h.  Importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# 64-byte AoS entity, byte-identical to sim::Entity (/root/reference/src/sim/Entity.hpp:33-46)
ENTITY_DTYPE = np.dtype(
    [
        ("color", "<f4", (4,)),
        ("rng", "<u4", (4,)),
        ("pos", "<f4", (2,)),
        ("target", "<f4", (2,)),
        ("dir", "<f4", (2,)),
        ("road", "<u4"),
        ("initialized", "<u4"),
    ]
)
# 32-byte road: two {vec2 pos, uint connectedIndex, uint connectedCount} (/root/reference/src/sim/Map.hpp:12-25)
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
assert ENTITY_DTYPE.itemsize == 64 and ROAD_DTYPE.itemsize == 32


class _Map(C.Structure):
    _fields_ = [
        ("world_w", C.c_float),
        ("world_h", C.c_float),
        ("roads", C.c_void_p),
        ("road_count", C.c_size_t),
        ("connections", C.c_void_p),
        ("connection_count", C.c_size_t),
    ]


class MoveStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("moved", "arrivals", "rng_draws", "uturns", "oob_reads", "initialised")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(ref: bool = True) -> None:
    """(Re)build the oracle; also oracle/_ref when /root/reference is present."""
    subprocess.check_call(["make", "-C", HERE, "libmsim_oracle.so"] + (["ref"] if ref else []), stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "libmsim_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.orc_xorshift128.restype = C.c_uint32
        L.orc_xorshift128.argtypes = [C.c_void_p]
        L.orc_next_float.restype = C.c_float
        L.orc_next_float.argtypes = [C.c_void_p]
        L.orc_next_range.restype = C.c_uint32
        L.orc_next_range.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_move_pass.restype = None
        L.orc_move_pass.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_Map), C.POINTER(MoveStats)]
        L.orc_move_pass_mt.restype = None
        L.orc_move_pass_mt.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_Map), C.c_int, C.POINTER(MoveStats)]
        L.orc_collide_pass_grid.restype = C.c_uint64
        L.orc_collide_pass_grid.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_float]
        L.orc_collide_pass_grid_mt.restype = C.c_uint64
        L.orc_collide_pass_grid_mt.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_int]
        L.orc_collide_pass_brute.restype = C.c_uint64
        L.orc_collide_pass_brute.argtypes = [C.c_void_p, C.c_size_t, C.c_float]
        L.orc_in_range.restype = C.c_int
        L.orc_in_range.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        L.orc_dispatch.restype = C.c_uint64
        L.orc_dispatch.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(_Map), C.c_float, C.c_uint32, C.POINTER(MoveStats)]
        L.orc_calc_node_count.restype = C.c_size_t
        L.orc_calc_node_count.argtypes = [C.c_size_t]
        _lib = L
    return _lib


def _check_entities(e: np.ndarray) -> None:
    assert e.dtype == ENTITY_DTYPE and e.flags.c_contiguous and e.flags.writeable


class OracleMap:
    """Holds the road / connection tables (numpy-owned) and the C view of them."""

    def __init__(self, world_w: float, world_h: float, roads: np.ndarray, connections: np.ndarray):
        self.roads = np.ascontiguousarray(roads, dtype=ROAD_DTYPE)
        self.connections = np.ascontiguousarray(connections, dtype=np.uint32)
        self.world_w = float(np.float32(world_w))
        self.world_h = float(np.float32(world_h))
        self.c = _Map(
            self.world_w,
            self.world_h,
            self.roads.ctypes.data,
            self.roads.shape[0],
            self.connections.ctypes.data,
            self.connections.shape[0],
        )


def xorshift128(state: np.ndarray) -> int:
    assert state.dtype == np.uint32 and state.shape == (4,)
    return int(lib().orc_xorshift128(state.ctypes.data))


def next_float(state: np.ndarray) -> np.float32:
    return np.float32(lib().orc_next_float(state.ctypes.data))


def next_range(state: np.ndarray, lo: int, hi: int) -> int:
    return int(lib().orc_next_range(state.ctypes.data, lo, hi))


def move_pass(e: np.ndarray, m: OracleMap, threads: int = 1) -> dict:
    _check_entities(e)
    st = MoveStats()
    if threads <= 1:
        lib().orc_move_pass(e.ctypes.data, e.shape[0], C.byref(m.c), C.byref(st))
    else:
        lib().orc_move_pass_mt(e.ctypes.data, e.shape[0], C.byref(m.c), threads, C.byref(st))
    return st.as_dict()


def collide_pass(e: np.ndarray, world_w: float, world_h: float, radius: float, threads: int = 1) -> int:
    _check_entities(e)
    return int(lib().orc_collide_pass_grid_mt(e.ctypes.data, e.shape[0], world_w, world_h, radius, threads))


def collide_pass_brute(e: np.ndarray, radius: float) -> int:
    _check_entities(e)
    return int(lib().orc_collide_pass_brute(e.ctypes.data, e.shape[0], radius))


def dispatch(e: np.ndarray, m: OracleMap, radius: float, tick: int) -> tuple[int, dict]:
    _check_entities(e)
    st = MoveStats()
    pairs = int(lib().orc_dispatch(e.ctypes.data, e.shape[0], C.byref(m.c), radius, tick, C.byref(st)))
    return pairs, st.as_dict()


def in_range(a, b, r: float) -> bool:
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    return bool(lib().orc_in_range(a.ctypes.data, b.ctypes.data, r))


def calc_node_count(depth: int) -> int:
    return int(lib().orc_calc_node_count(depth))


BLUE = np.array([0, 0, 1, 1], dtype=np.float32)
GREEN = np.array([0, 1, 0, 1], dtype=np.float32)


def collision_flags(e: np.ndarray) -> np.ndarray:
    """1 where colour == blue (0,0,1,1), 0 where green; asserts nothing else is present."""
    blue = (e["color"] == BLUE).all(axis=1)
    return blue.astype(np.uint8)


# --------------------------------------------------------------------------------------------
# oracle/_ref: the reference's own CPU quadtree (shader_validation/src/main.cpp) compiled here
# --------------------------------------------------------------------------------------------
REF_LIB = os.path.join(HERE, "_ref", "libref_quadtree.so")
REF_LIB_LARGE = os.path.join(HERE, "_ref", "libref_quadtree_large.so")  # same sources, capacity for BASELINE config 3 at full size
REF_KAT = os.path.join(HERE, "_ref", "ref_kat")
_ref = None
_ref_large = None


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


def ref_large_available() -> bool:
    return os.path.exists(REF_LIB_LARGE)


def ref(large: bool = False):
    """The compiled reference harness.  large=True: the build with room for 10 M entities (bench.py's reference arm)."""
    global _ref, _ref_large
    if large:
        if _ref_large is None:
            _ref_large = _bind_ref(C.CDLL(REF_LIB_LARGE))
        return _ref_large
    if _ref is None:
        _ref = _bind_ref(C.CDLL(REF_LIB))
    return _ref


def _bind_ref(R):
    if True:
        R.ref_capacity.restype = C.c_size_t
        R.ref_node_count.restype = C.c_size_t
        R.ref_reset.restype = None
        R.ref_reset.argtypes = [C.c_float, C.c_float, C.c_uint32, C.c_float]
        R.ref_insert_all.restype = C.c_int
        R.ref_insert_all.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        R.ref_update_all.restype = C.c_int
        R.ref_update_all.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        R.ref_collide_all.restype = C.c_uint64
        R.ref_collide_all.argtypes = [C.c_size_t, C.c_int, C.c_void_p]
        R.ref_get_positions.restype = None
        R.ref_get_positions.argtypes = [C.c_void_p, C.c_size_t]
        R.ref_count_entities_in_tree.restype = C.c_size_t
        R.ref_dump_tree.restype = C.c_size_t
        R.ref_dump_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    return R


class RefQuadTree:
    """Drives the compiled reference quadtree: insert -> (update)* -> collide, as the shader's
    dispatches do (random_move.comp:863-877)."""

    def __init__(self, world_w: float, world_h: float, radius: float = 10.0, node_cap: int = 10, large: bool = False):
        self.R = ref(large)
        self.R.ref_reset(world_w, world_h, node_cap, radius)
        self.n = 0

    def insert(self, xy: np.ndarray, threads: int = 1) -> None:
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        self.n = xy.shape[0]
        if self.R.ref_insert_all(xy.ctypes.data, self.n, threads) != 0:
            raise ValueError("reference harness capacity exceeded")

    def update(self, xy: np.ndarray, threads: int = 1) -> None:
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        assert xy.shape[0] == self.n
        self.R.ref_update_all(xy.ctypes.data, self.n, threads)

    def collide(self, threads: int = 1) -> tuple[np.ndarray, int]:
        flags = np.zeros(self.n, dtype=np.uint8)
        cnt = int(self.R.ref_collide_all(self.n, threads, flags.ctypes.data))
        return flags, cnt

    def count_in_tree(self) -> int:
        return int(self.R.ref_count_entities_in_tree())

    def dump_tree(self):
        """Depth-first (TL, TR, BL, BR) list of the reference tree's nodes: (rects[n,4], types[n], counts[n])."""
        cap = int(self.R.ref_node_count())
        rects = np.zeros((cap, 4), dtype=np.float32)
        types = np.zeros(cap, dtype=np.uint32)
        counts = np.zeros(cap, dtype=np.uint32)
        n = int(self.R.ref_dump_tree(rects.ctypes.data, types.ctypes.data, counts.ctypes.data, cap))
        return rects[:n], types[:n], counts[:n]


# --------------------------------------------------------------------------------------------
# The movement half of the reference's compute shader, compiled from its own text (oracle/Makefile,
# oracle/ref_shader_prelude.inc + random_move.comp:5-24,725-750,778-852 + oracle/ref_shader_driver.inc)
# --------------------------------------------------------------------------------------------
REF_SHADER_LIB = os.path.join(HERE, "_ref", "libref_shader_move.so")
_ref_shader = None


def ref_shader_available() -> bool:
    return os.path.exists(REF_SHADER_LIB)


def ref_shader():
    global _ref_shader
    if _ref_shader is None:
        R = C.CDLL(REF_SHADER_LIB)
        R.ref_shader_move_pass.restype = None
        R.ref_shader_move_pass.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        for name in ("ref_shader_next", "ref_shader_next_range"):
            getattr(R, name).restype = C.c_uint32
        R.ref_shader_next.argtypes = [C.c_void_p]
        R.ref_shader_next_range.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        R.ref_shader_next_float.restype = C.c_float
        R.ref_shader_next_float.argtypes = [C.c_void_p]
        R.ref_shader_speed.restype = C.c_float
        R.ref_shader_move_pass_mt.restype = None
        R.ref_shader_move_pass_mt.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        _ref_shader = R
    return _ref_shader


def ref_shader_move_pass(e: np.ndarray, m: OracleMap, threads: int = 1) -> None:
    """One even-tick dispatch of the shader's main() (random_move.comp:860-873) over all entities, executed by the
    shader's own update_direction / move / new_target / next compiled for the CPU."""
    _check_entities(e)
    padded = getattr(m, "_padded_connections", None)
    if padded is None:  # App. B1: the shader reads one entry past the table; canonical value 0
        padded = np.concatenate([m.connections, np.zeros(1, dtype=np.uint32)])
        m._padded_connections = padded
    if threads > 1:
        ref_shader().ref_shader_move_pass_mt(e.ctypes.data, e.shape[0], m.roads.ctypes.data, padded.ctypes.data, threads)
    else:
        ref_shader().ref_shader_move_pass(e.ctypes.data, 0, e.shape[0], m.roads.ctypes.data, padded.ctypes.data)


# --------------------------------------------------------------------------------------------
# The WHOLE shader compiled from its own text (oracle/Makefile: libref_shader_full.so): main() with its three branches -
# initialise + quad_tree_insert, even tick move + quad_tree_update, odd tick colour + quad_tree_check_collisions -
# executed invocation by invocation on one host thread.
# --------------------------------------------------------------------------------------------
REF_SHADER_FULL_LIB = os.path.join(HERE, "_ref", "libref_shader_full.so")
NODE_DTYPE = np.dtype([("acquireLock", "<i4"), ("writeLock", "<i4"), ("readerLock", "<i4"), ("offsetX", "<f4"), ("offsetY", "<f4"), ("width", "<f4"),
                       ("height", "<f4"), ("contentType", "<u4"), ("entityCount", "<u4"), ("first", "<u4"), ("prevNodeIndex", "<u4"), ("nextTL", "<u4"),
                       ("nextTR", "<u4"), ("nextBL", "<u4"), ("nextBR", "<u4"), ("padding", "<u4")])
TREE_ENTITY_DTYPE = np.dtype([("nodeIndex", "<u4"), ("typeNext", "<u4"), ("next", "<u4"), ("typePrev", "<u4"), ("prev", "<u4")])
assert NODE_DTYPE.itemsize == 64 and TREE_ENTITY_DTYPE.itemsize == 20
_ref_shader_full = None


def ref_shader_full_available() -> bool:
    return os.path.exists(REF_SHADER_FULL_LIB)


class RefShaderDeadlock(RuntimeError):
    """The shader's lock protocol left a lock behind (a reference bug in quad_tree_update's remove / merge path, also tripped by
    the author's own CPU harness: DESIGN.md)."""


class RefShaderSim:
    """The reference's Simulator::init + sim_tick around the compiled shader: buffers as Simulator.cpp:58-89 creates them, push
    constants as :94-101, one ref_shader_full_dispatch per tick number (:220-235).  Single-threaded; a handful of thousand
    entities per second of patience (every insert / update walks the tree from the root under its lock protocol)."""

    def __init__(self, e: np.ndarray, m: OracleMap, radius: float = 10.0, max_depth: int = 8, node_cap: int = 10):
        global _ref_shader_full
        if _ref_shader_full is None:
            R = C.CDLL(REF_SHADER_FULL_LIB)
            R.ref_shader_full_bind.restype = None
            R.ref_shader_full_bind.argtypes = [C.c_void_p] * 7
            R.ref_shader_full_push_consts.restype = None
            R.ref_shader_full_push_consts.argtypes = [C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32]
            R.ref_shader_full_dispatch.restype = C.c_int64
            R.ref_shader_full_dispatch.argtypes = [C.c_uint64]
            _ref_shader_full = R
        self.R = _ref_shader_full
        _check_entities(e)
        self.e, self.m, self.radius, self.max_depth, self.node_cap = e, m, float(radius), int(max_depth), int(node_cap)
        self.connections = np.concatenate([m.connections, np.zeros(1, dtype=np.uint32)])  # App. B1
        self.nodes = np.zeros(calc_node_count(max_depth), dtype=NODE_DTYPE)
        self.nodes[0]["width"], self.nodes[0]["height"], self.nodes[0]["contentType"] = m.world_w, m.world_h, 2  # init_node_zero
        self.tree_entities = np.zeros(max(1, e.shape[0]), dtype=TREE_ENTITY_DTYPE)
        self.used = np.zeros(self.nodes.shape[0] + 2, dtype=np.uint32)
        self.used[1] = 2  # Simulator.cpp:81
        self.debug = np.zeros(10, dtype=np.uint32)

    def dispatch(self, tick: int) -> None:
        self.R.ref_shader_full_bind(self.e.ctypes.data, self.connections.ctypes.data, self.m.roads.ctypes.data, self.nodes.ctypes.data,
                                    self.tree_entities.ctypes.data, self.used.ctypes.data, self.debug.ctypes.data)
        self.R.ref_shader_full_push_consts(self.m.world_w, self.m.world_h, self.nodes.shape[0], self.max_depth, self.node_cap, self.radius, tick)
        stuck = int(self.R.ref_shader_full_dispatch(self.e.shape[0]))
        if stuck >= 0:
            raise RefShaderDeadlock(f"tick {tick}: invocation {stuck} found a quadtree lock that no running invocation holds")


def run_ref_kat() -> subprocess.CompletedProcess:
    """Runs the reference's own (disabled) known-answer test in a subprocess (it asserts)."""
    return subprocess.run([REF_KAT], capture_output=True, text=True, timeout=120)
