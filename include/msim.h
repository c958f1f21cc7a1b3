/*
 * msim.h — C ABI of libmsim_cuda.so: the B200 (sm_100a) replacement for movement-sim's per-tick
 * entity update.
 *
 * The reference has no FFI layer; its de-facto operator API is the handful of Kompute calls that
 * sim::Simulator makes.  Each entry point below names the reference call it replaces
 * (paths relative to the reference checkout).  Plain pointers and sizes only; no C++/torch types; no
 * exceptions cross this boundary (int status, 0 = MSIM_OK, message via msim_last_error()).
 *
 * Ownership: the caller owns every pointer passed in or out; the library copies and retains nothing
 * after return.  Threading: thread-compatible — one caller at a time per handle; create on thread A
 * and step on thread B is fine (every call binds the handle's device first), mirroring
 * src/sim/Simulator.cpp:44-108 (init on main thread) vs :185-211 (worker thread).
 *
 * There is NO CPU fallback: every compute entry point returns MSIM_ERR_CUDA when no sm_100 device is
 * usable.
 */
#ifndef MSIM_H
#define MSIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSIM_ABI_VERSION 2u /* 2: msim_stats grew by total_flagged_count */

/* ---- status codes -------------------------------------------------------------------------- */
enum {
    MSIM_OK = 0,
    MSIM_ERR_INVALID = 1,     /* bad argument / malformed map (e.g. connection index >= road count) */
    MSIM_ERR_CUDA = 2,        /* CUDA runtime error or no usable device */
    MSIM_ERR_OOM = 3,         /* host or device allocation failed */
    MSIM_ERR_UNSUPPORTED = 4, /* valid request this build does not implement */
    MSIM_ERR_IO = 5,          /* file missing / unreadable (Map.cpp:30-38 returns nullptr) */
    MSIM_ERR_PARSE = 6,       /* JSON schema error (Map.cpp:43-117 throws std::runtime_error) */
    MSIM_ERR_CAPACITY = 7,    /* shard ran out of entity / exchange capacity */
    MSIM_ERR_INTERNAL = 8     /* device-side watchdog tripped (never expected) */
};

/* ---- data model: byte-identical PODs ------------------------------------------------------- */

/* sim::Entity, 64 B (src/sim/Entity.hpp:33-46) == shader EntityDescriptor (random_move.comp:5-13).
 * This is the readback contract of the UI: colour @0, position @32
 * (src/ui/widgets/opengl/EntityGlObject.cpp:9-12,67,72). */
typedef struct msim_entity {
    float color[4];         /*  0 */
    uint32_t rand_state[4]; /* 16: x,y,z,w of the xorshift128 state */
    float pos[2];           /* 32 */
    float target[2];        /* 40 */
    float direction[2];     /* 48 */
    uint32_t road_index;    /* 56 */
    uint32_t initialized;   /* 60 */
} msim_entity;

/* sim::Coordinate 16 B / sim::Road 32 B (src/sim/Map.hpp:12-25; shader :15-24) */
typedef struct msim_coordinate {
    float pos[2];
    uint32_t connected_index;
    uint32_t connected_count;
} msim_coordinate;

typedef struct msim_road {
    msim_coordinate start;
    msim_coordinate end;
} msim_road;

/* sim::gpu_quad_tree::Node, 64 B (src/sim/GpuQuadTree.hpp:14-41; shader :54-76) */
typedef struct msim_quadtree_node {
    int32_t acquire_lock, write_lock, reader_lock;
    float offset_x, offset_y, width, height;
    uint32_t content_type; /* 0 invalid, 1 node, 2 entity */
    uint32_t entity_count;
    uint32_t first;
    uint32_t prev_node_index;
    uint32_t next_tl, next_tr, next_bl, next_br;
    uint32_t padding;
} msim_quadtree_node;

/* sim::PushConsts, 28 B packed (src/sim/PushConsts.hpp:9-20; shader :26-37) */
typedef struct msim_push_consts {
    float world_size_x;
    float world_size_y;
    uint32_t node_count;
    uint32_t max_depth;
    uint32_t entity_node_cap;
    float collision_radius;
    uint32_t tick;
} msim_push_consts;

/* ---- handle -------------------------------------------------------------------------------- */
typedef struct msim_handle msim_handle;

enum {
    MSIM_FLAG_NO_COLLISIONS = 1u << 0, /* even-tick dispatches do not emit cell keys; odd ticks are rejected */
    MSIM_FLAG_NO_PAIR_COUNT = 1u << 1, /* collision query stops at the first neighbour (flags only) */
    MSIM_FLAG_NO_QUADTREE   = 1u << 2, /* msim_read_quadtree_nodes returns the root only */
    MSIM_FLAG_SORT_COUNTING = 1u << 3, /* always rebuild the neighbour structure with the single-digit (counting) radix sort */
    MSIM_FLAG_NO_REORDER    = 1u << 4, /* keep the resident state in upload order (default: re-sorted into cell order every
                                          32 collision passes, with the counting sort as the rebuild) */
    MSIM_FLAG_SORT_ONESWEEP = 1u << 5  /* always rebuild with the multi-pass onesweep radix sort */
};

typedef struct msim_config {
    uint32_t abi_version; /* MSIM_ABI_VERSION */
    int32_t device;       /* CUDA ordinal */
    uint32_t flags;       /* MSIM_FLAG_* */
    uint32_t reserved0;

    float world_w, world_h; /* Map::width/height = maxDistLat/maxDistLong (Map.cpp:45-52, Simulator.cpp:95-96) */
    float collision_radius; /* sim::COLLISION_RADIUS (Simulator.hpp:43) */
    uint32_t quadtree_max_depth;  /* sim::QUAD_TREE_MAX_DEPTH, display tree only (Simulator.hpp:37) */
    uint32_t quadtree_node_cap;   /* sim::QUAD_TREE_ENTITY_NODE_CAP, display tree only (Simulator.hpp:38) */
    uint32_t reserved1;

    const msim_road* roads;       /* tensorRoads (Simulator.cpp:64) */
    uint64_t road_count;
    const uint32_t* connections;  /* tensorConnections (Simulator.cpp:65) */
    uint64_t connection_count;
    const msim_entity* entities;  /* tensorEntities (Simulator.cpp:61); may be NULL when entity_count == 0 */
    uint64_t entity_count;
    uint64_t entity_capacity;     /* >= entity_count; 0 means entity_count. Head-room for migrants when sharded */

    void* cuda_stream; /* cudaStream_t to enqueue on, or NULL for a library-owned stream */
} msim_config;

/* Replaces kp::Manager() + 7x mgr->tensor(...) + mgr->algorithm(...) (Simulator.cpp:52-103) and the
 * one-off OpTensorSyncDevice upload (Simulator.cpp:191-192): validates the map, allocates the SoA
 * state in HBM, uploads and transposes the AoS entities on the device. */
int msim_create(const msim_config* cfg, msim_handle** out);
/* Implicit shared_ptr release in the reference. */
void msim_destroy(msim_handle* h);
/* Last error text of this handle (or of the failed msim_create when h == NULL); never NULL. */
const char* msim_last_error(const msim_handle* h);
const char* msim_status_string(int status);

/* Re-upload entity state (same count or fewer than capacity): OpTensorSyncDevice on tensorEntities. */
int msim_upload_entities(msim_handle* h, const msim_entity* src, uint64_t count);

/* calcSeq->eval<OpAlgoDispatch>(algo, pushConsts) (Simulator.cpp:224,235): one blocking dispatch of
 * the shader's main() (random_move.comp:860-879) over all entities —
 *   uninitialised entities: initialise only;   tick even: move pass;   tick odd: collision pass.
 * Returns after the device finished, preserving the timing semantics of Simulator.cpp:223-225. */
int msim_dispatch(msim_handle* h, const msim_push_consts* pc);

/* Asynchronous building blocks of the same dispatch (enqueue only; pair with msim_sync). */
int msim_enqueue_move(msim_handle* h);
int msim_enqueue_collide(msim_handle* h);
/* `sim_ticks` x (move pass [+ collision pass]) == that many Simulator::sim_tick calls
 * (Simulator.cpp:213-241) without the host round trips: everything is enqueued, nothing is waited for. */
int msim_enqueue_ticks(msim_handle* h, uint32_t sim_ticks, int with_collisions);
int msim_sync(msim_handle* h);
int msim_set_stream(msim_handle* h, void* cuda_stream);

/* OpTensorSyncLocal({tensorEntities}) + tensor->vector<Entity>() (Simulator.cpp:197,250,262):
 * packs SoA -> 64-byte AoS on the device and copies `count` entities to dst (host memory). */
int msim_read_entities(msim_handle* h, msim_entity* dst, uint64_t count);
/* Asynchronous readback: what replaces the reference's blocking per-tick evalAsync/evalAwait +
 * tensor->vector<Entity>() copy (Simulator.cpp:250-262) for consumers that can take a pointer.
 *   begin  packs the CURRENT state (everything enqueued so far) SoA -> AoS into a device image on the handle's stream and
 *          starts its copy into library-owned PINNED host memory on a separate copy stream; returns at once, ticks enqueued
 *          afterwards run while the copy engine drains the image;
 *   poll   *ready = 1 once the copy has landed (never blocks);
 *   end    waits for it and hands out the pinned buffer: valid until the begin AFTER the next one (two host buffers
 *          alternate, so a consumer may keep reading one snapshot while the next is being filled) or msim_destroy.
 * One snapshot may be in flight per handle: begin while one is pending first waits (on the device) for its copy. */
int msim_snapshot_begin(msim_handle* h);
int msim_snapshot_poll(msim_handle* h, int* ready);
int msim_snapshot_end(msim_handle* h, const msim_entity** entities, uint64_t* count);
/* Rendering fast path: positions only (8 B per entity) and collision flags (1 B per entity). */
int msim_read_positions(msim_handle* h, float* dst_xy, uint64_t count);
int msim_read_collision_flags(msim_handle* h, uint8_t* dst, uint64_t count);
/* OpTensorSyncLocal({tensorQuadTreeNodes}) (Simulator.cpp:198,255,267): display quadtree rebuilt from
 * the current positions. *count receives the number of nodes written (<= cap). */
int msim_read_quadtree_nodes(msim_handle* h, msim_quadtree_node* dst, uint64_t cap, uint64_t* count);
/* tensorDebugData (Simulator.cpp:86-89,273): [0] = cumulative initialisations (the quad_tree_insert
 * calls of the first dispatch, shader :311; the shader also counts the re-inserts of quad_tree_update
 * there, which depend on the tree's shape and are not reproduced), [1] = cumulative UNIQUE in-range
 * pairs (documented deviation from the reference's over-count), [2..9] = 0. */
int msim_read_debug(msim_handle* h, uint32_t dst[10]);

typedef struct msim_stats {
    uint64_t entity_count;
    uint64_t move_passes;
    uint64_t collide_passes;
    uint64_t last_pair_count;    /* unique pairs found by the last collision pass */
    uint64_t total_pair_count;
    uint64_t last_flagged_count; /* entities coloured blue by the last collision pass */
    uint64_t kernel_launches;    /* kernels launched by this handle so far */
    uint32_t grid_cells_x, grid_cells_y;
    uint32_t key_bits, sort_passes;
    float cell_size;
    uint32_t reorders;           /* physical re-sorts of the resident state so far */
    uint64_t total_flagged_count; /* last_flagged_count summed over every collision pass so far */
} msim_stats;
int msim_get_stats(msim_handle* h, msim_stats* out);

/* Per-kernel device time between msim_profile_begin and msim_profile_end, measured with CUDA events
 * on the handle's stream around every launch (bench.py's live roofline).  Profiling adds two event
 * records per launch, so throughput numbers are taken with profiling off. */
typedef struct msim_kernel_time {
    char name[24];
    uint64_t launches;
    double total_ms;
} msim_kernel_time;
int msim_profile_begin(msim_handle* h);
int msim_profile_end(msim_handle* h, msim_kernel_time* out, uint32_t cap, uint32_t* count);

/* Device pointers of the resident SoA state, for zero-copy consumers (CUDA-GL interop, torch). */
typedef struct msim_device_view {
    void* pos;       /* float2[count], current */
    void* target;    /* float2[count] */
    void* road;      /* uint32[count] */
    void* rng;       /* uint4[count] */
    uint64_t count;
    void* ext_id;    /* uint32[count]: external entity id of each storage slot, or NULL while storage is in upload order */
} msim_device_view;
int msim_get_device_view(msim_handle* h, msim_device_view* out);

/* ---- host-side helpers (no GPU needed) ----------------------------------------------------- */
typedef struct msim_map msim_map;

/* Map::load_from_file (src/sim/Map.cpp:28-150), same schema, same zero-length-road skip (:124-128). */
int msim_map_load_json(const char* path, msim_map** out);
int msim_map_save_json(const msim_map* m, const char* path);
/* Seeded synthetic stand-in for the missing munich.json: jittered-grid street graph emitted with the
 * connection-table layout of map/generate_map.py:234-258. */
int msim_map_generate_city(float world_w, float world_h, float spacing, float jitter, float drop_prob,
                           uint64_t seed, msim_map** out);
/* nx x ny lattice, nodes at (spacing*i, spacing*j) (BASELINE config 4). */
int msim_map_generate_grid(uint32_t nx, uint32_t ny, float spacing, msim_map** out);
void msim_map_free(msim_map* m);
float msim_map_width(const msim_map* m);
float msim_map_height(const msim_map* m);
uint64_t msim_map_road_count(const msim_map* m);
uint64_t msim_map_connection_count(const msim_map* m);
const msim_road* msim_map_roads(const msim_map* m);
const uint32_t* msim_map_connections(const msim_map* m);
const char* msim_map_last_error(void);

/* Simulator::add_entities (src/sim/Simulator.cpp:114-129) made reproducible: three std::mt19937
 * seeded seed, seed+1, seed+2 for road index / colour / RNG state (Map.cpp:152-157,
 * Entity.cpp:44-60).  When box != NULL ({x0,y0,x1,y1}) roads are drawn only from those with both
 * ends inside the box (dense-crowd config 5). */
int msim_entities_init(const msim_road* roads, uint64_t road_count, uint64_t count, uint64_t seed,
                       const float* box, msim_entity* out);

/* The road-index stream of msim_entities_init alone: road_index_out[i] = the road entity i of that seeded population starts on (its position
 * is that road's start point).  Lets a sharded host find out where a population lives without building it. */
int msim_entities_init_roads(const msim_road* roads, uint64_t road_count, uint64_t count, uint64_t seed, const float* box,
                             uint32_t* road_index_out);

/* Host-only twin of msim_read_quadtree_nodes (no GPU): the display quadtree of caller-owned positions (xy = count x {x, y}), built by the
 * same code from a leaf histogram taken on the host. */
int msim_quadtree_from_positions(const float* xy, uint64_t count_in, float world_w, float world_h, uint32_t max_depth, uint32_t node_cap,
                                 msim_quadtree_node* dst, uint64_t cap, uint64_t* count);

/* gpu_quad_tree::calc_node_count (src/sim/GpuQuadTree.cpp:11-17) */
uint64_t msim_calc_node_count(uint32_t max_depth);
uint32_t msim_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MSIM_H */
