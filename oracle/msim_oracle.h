/*
 * msim_oracle.h — CPU ORACLE for the movement-sim per-tick entity update.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker or the timed CPU baseline.  The product path (movement-sim_b200/) never links
 * or calls it and fails loudly when its CUDA library is missing.
 *
 * What it restates (citations relative to /root/reference):
 *   - src/sim/shader/random_move.comp:725-746   xorshift128, next_float, next(state,min,max)
 *   - src/sim/shader/random_move.comp:778-852   new_target, update_direction, move
 *   - src/sim/shader/random_move.comp:860-879   main(): init / even tick move / odd tick collide
 *   - src/sim/shader/random_move.comp:545-562   collision colour + in_range predicate
 *   - src/sim/Simulator.cpp:220-235             tick sequencing (2 dispatches per sim tick)
 * Float rules: IEEE-754 binary32, round-to-nearest-even, no FMA contraction (build with
 * -ffp-contract=off), correctly rounded sqrt and divide.
 *
 * PARITY PIN STATUS.  The reference holds no golden vector for movement / road choice / RNG
 * (SURVEY.md §8c), so the pin is the reference's own source: oracle/Makefile compiles the movement
 * functions of random_move.comp (:5-24, :725-750, :778-852) as C++ into oracle/_ref/libref_shader_move.so
 * (GLSL prelude + driver of ours around the streamed shader text) and this oracle must equal it bit for
 * bit on every field after every pass (tests/test_oracle_vs_ref_shader.py: config 1 in full, a street
 * graph population, RNG known answers, degenerate distances); digests the compiled shader produced are
 * committed for boxes without oracle/_ref (tests/golden/ref_shader_digests.json).  What stays open is
 * the float freedom of a real Vulkan driver: both sides use IEEE RNE without FMA.
 * The whole shader - main() and the lock-based quadtree included - is compiled the same way into
 * oracle/_ref/libref_shader_full.so; run dispatch by dispatch on one host thread it equals this oracle on all
 * 64 bytes of every entity, colours included (same test file).
 * The collision half is pinned a second time against the author's C++ harness: the reference's CPU
 * restatement (shader_validation/src/main.cpp) is compiled into oracle/_ref and (a) its disabled
 * known-answer test run_collision_detection_test_1 passes, (b) its flagged set equals this oracle's
 * on seeded point clouds (tests/test_oracle_vs_ref.py).  calc_node_count(1,2,3,4,8) = 1,5,21,85,21845
 * (src/sim/Simulator.cpp:72-76) is checked too.
 */
#ifndef MSIM_ORACLE_H
#define MSIM_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 64-byte AoS entity — byte-identical to sim::Entity (src/sim/Entity.hpp:33-46) and the shader's
 * EntityDescriptor (random_move.comp:5-13). */
typedef struct orc_entity {
    float color[4];     /*  0 */
    uint32_t rng[4];    /* 16  x,y,z,w */
    float pos[2];       /* 32 */
    float target[2];    /* 40 */
    float dir[2];       /* 48 */
    uint32_t road;      /* 56 */
    uint32_t initialized; /* 60 */
} orc_entity;

/* 16-byte coordinate / 32-byte road — sim::Coordinate / sim::Road (src/sim/Map.hpp:12-25). */
typedef struct orc_coord {
    float pos[2];
    uint32_t conn_index;
    uint32_t conn_count;
} orc_coord;

typedef struct orc_road {
    orc_coord start;
    orc_coord end;
} orc_road;

typedef struct orc_map {
    float world_w, world_h;
    const orc_road* roads;
    size_t road_count;
    const uint32_t* connections;
    size_t connection_count;
} orc_map;

/* per-pass counters the tests and the roofline text use (p_arr, p_rng of SURVEY §8d) */
typedef struct orc_move_stats {
    uint64_t moved;        /* entities that took the move branch */
    uint64_t arrivals;     /* dist <= SPEED */
    uint64_t rng_draws;    /* arrivals with connectedCount > 2 */
    uint64_t uturns;       /* arrivals with connectedCount <= 1 */
    uint64_t oob_reads;    /* connections[] index >= connection_count (App. B1) */
    uint64_t initialised;  /* entities that took the init branch */
} orc_move_stats;

/* random_move.comp:725-736 */
uint32_t orc_xorshift128(uint32_t s[4]);
/* random_move.comp:738-741 — returned as raw float bits-compatible value */
float orc_next_float(uint32_t s[4]);
/* random_move.comp:743-746 */
uint32_t orc_next_range(uint32_t s[4], uint32_t lo, uint32_t hi);

/* One even-tick dispatch over entities [0,n): init branch for uninitialised entities, otherwise
 * update_direction + move (+ new_target).  random_move.comp:860-873. */
void orc_move_pass(orc_entity* e, size_t n, const orc_map* map, orc_move_stats* stats);
/* same, statically split over `threads` pthreads (the timed CPU baseline) */
void orc_move_pass_mt(orc_entity* e, size_t n, const orc_map* map, int threads, orc_move_stats* stats);

/* One odd-tick dispatch, deterministic intent of random_move.comp:875-877 (App. B4): initialised
 * entities become green (0,1,0,1), then blue (0,0,1,1) iff another initialised entity is in range.
 * Uninitialised entities take the init branch.  Returns the number of unique in-range pairs. */
uint64_t orc_collide_pass_grid(orc_entity* e, size_t n, float world_w, float world_h, float radius);
uint64_t orc_collide_pass_grid_mt(orc_entity* e, size_t n, float world_w, float world_h, float radius, int threads);
/* O(n^2) cross-check of the grid version */
uint64_t orc_collide_pass_brute(orc_entity* e, size_t n, float radius);

/* random_move.comp:551-562 */
int orc_in_range(const float a[2], const float b[2], float max_distance);

/* The dispatch as Kompute issues it: branches on tick parity (random_move.comp:869). */
uint64_t orc_dispatch(orc_entity* e, size_t n, const orc_map* map, float radius, uint32_t tick, orc_move_stats* stats);

/* gpu_quad_tree::calc_node_count (src/sim/GpuQuadTree.cpp:11-17) */
size_t orc_calc_node_count(size_t max_depth);

#ifdef __cplusplus
}
#endif
#endif
