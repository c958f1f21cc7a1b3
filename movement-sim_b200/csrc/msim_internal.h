// msim_internal.h — internal declarations shared by the CUDA translation units of libmsim_cuda.so.
//
// HBM layout (structure of arrays, one element per entity, capacity `cap`):
//   pos[2]   float2   ping-pong position buffers (move reads pos[cur], writes pos[cur^1])
//   target   float2   current waypoint
//   road     u32      current road index              (touched only on arrival)
//   rng      uint4    xorshift128 state (x,y,z,w)     (touched only on arrival at a junction of >2 roads)
//   color0   float4   colour as uploaded              (cold: readback only)
//   dir0     float2   direction as uploaded           (cold: readback before the first move pass)
//   arrived  u32 bitmask, written by the move pass; lets readback reconstruct `direction` exactly
//            without storing 8 B per entity per tick (see pack.cu)
// Neighbour structure, rebuilt every collision pass:
//   keys     u32      cell key (row-major) emitted by the move pass or by keygen
//   sort_a/b u64      (key << 32 | entity index) pairs, ping-pong for the LSD radix passes
//   sorted_pos float2 positions in cell order
//   cell_range uint2  per cell {first, ~end} into the sorted order (0xFF-filled = empty)
//   flag_sorted u8    collision flag per sorted slot (scattered back lazily at readback)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/msim.h"

namespace msim {

// ---- neighbour grid geometry -----------------------------------------------------------------
struct GridParams {
    float inv_cell;      // 1 / cell edge
    float hit_threshold; // d2 < hit_threshold  <=>  sqrtf(d2) < radius  (exact, see api.cu)
    float radius;
    int ncx, ncy;        // cells per axis
    uint32_t ncells;
};

// ---- radix sort constants --------------------------------------------------------------------
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 pairs per tile
constexpr int MAX_SORT_PASSES = 4;

struct SortWorkspace {
    uint32_t* hist;          // [MAX_SORT_PASSES][RADIX] global digit histograms
    uint32_t* tile_counter;  // [MAX_SORT_PASSES] dynamic tile ids
    uint32_t* tile_state;    // [MAX_SORT_PASSES][tiles][RADIX] decoupled look-back words
    uint32_t* error_flag;    // set by the look-back watchdog
    void* zero_base;         // start of the region that must be zeroed before every sort
    size_t zero_bytes;
    uint32_t tiles_cap;
};

// ---- device counters -------------------------------------------------------------------------
struct Counters {
    unsigned long long pairs_last;
    unsigned long long pairs_total;
    unsigned long long flagged_last;
    unsigned long long flagged_total;  // summed over every collision pass (cross-checks between runs on different GPU counts)
    unsigned int error_flag;
    unsigned int pad;
};

// ---- fused move + shard pack (move.cu, shard.cu): what the move kernel needs to classify the entities it has
// just moved against the band [lo_key, hi_key) of cell keys and to write leavers / halo straight into the
// exchange buffers (local send buffers, or the neighbour's receive buffers over NVLink peer memory)
struct ShardMoveArgs {
    uint32_t lo_key, hi_key;  // first key of the band's first row, first key past its last row
    uint32_t ncx;
    void* buf_down;           // NULL: no neighbour below (nobody leaves that way, no halo)
    void* buf_up;
    uint32_t mig_cap, halo_cap, holes_cap;
    uint32_t* holes;
    float2* local_ghosts;
    uint32_t* ctr;
    const uint4* rng;
    const float4* color0;
    const uint32_t* road;
    const uint32_t* gid;
    uint32_t* error_word;      // |= 2 when the hole list overflows
    uint32_t* peer_flag_down;  // (stand-alone use of shard_publish on remote buffers) raised once everything of this tick is in the buffers
    uint32_t* peer_flag_up;
    uint32_t signal_value;
    uint32_t publish;  // 1: a one-thread kernel behind the move kernel writes the headers of buf_down / buf_up (local send buffers of the collective exchange)
};

// ---- per-kernel CUDA-event timing (bench.py's live roofline; off unless msim_profile_begin) -----
enum KernelId {
    K_MOVE = 0, K_ARRIVE, K_KEYGEN, K_HISTOGRAM, K_SORT_PASS0, K_SORT_PASS1, K_SORT_PASS2, K_SORT_PASS3, K_BUILD_CELLS, K_QUERY,
    K_SCATTER_FLAGS, K_PACK, K_UNPACK, K_MEMSET, K_MISC, K_SHARD, K_CELL_COUNT, K_CELL_SCAN, K_CELL_SCATTER, K_REORDER, K_FOLD, K_COUNT
};

struct Profiler {
    bool enabled{false};
    std::vector<cudaEvent_t> pool;   // reusable events
    std::vector<int> ids;            // kernel id per (start, stop) pair in use
    size_t used{0};                  // events handed out this session

    void begin(cudaStream_t s, int id) {
        if (!enabled) return;
        if (used + 2 > pool.size()) {
            pool.resize(used + 2);
            cudaEventCreate(&pool[used]);
            cudaEventCreate(&pool[used + 1]);
        }
        ids.push_back(id);
        cudaEventRecord(pool[used], s);
    }
    void end(cudaStream_t s) {
        if (!enabled) return;
        cudaEventRecord(pool[used + 1], s);
        used += 2;
    }
};

// ---- launch tuning read once from the environment (api.cu)
struct Tuning {
    int arrive_beside_ctas_per_sm{1};  // MSIM_ARRIVE_BESIDE_CTAS=0..8: when pass B rides beside the query it is launched as a strided grid of that many
                                  // CTAs per SM, so that it trickles through the whole query on a fraction of the warp slots (it is latency-bound and
                                  // only has to finish before the next move); 0 = full grid.  Measured at 10 M entities: tick 361 / 334 / 339 us for 0 / 1 / 2
    int csort_max_cells_log2{27};     // MSIM_CSORT_MAX_CELLS_LOG2: 25 .. 27 (default: every grid the library accepts); grids with more cells take the onesweep
                                  // rebuild.  BASELINE config 4 (8159 x 8159 cells = 2^25.99) keeps the counting sort: two 266 MB tables per GPU, of
                                  // which a band-sharded handle scans and touches only its own rows
    bool overlap_ticks{true};         // MSIM_OVERLAP_TICKS=0: the move phase of tick t+1 waits for the query of tick t (it runs beside it by default)
    int move_beside_ctas_per_sm{2};   // MSIM_MOVE_BESIDE_CTAS=1..8: CTAs per SM of the move kernel while it shares the SMs with a query (8 = the stand-alone grid).  Tick at 10 M entities: 304 / 306 / 313 / 317 / 315 us for 2 / 3 / 4 / 6 / 8, 327 us without the overlap
    // band-sharded handles have their own pair of knobs (MSIM_SHARD_ARRIVE_BESIDE_CTAS 0 = full grid .. 8, MSIM_SHARD_MOVE_BESIDE_CTAS 1 .. 8).
    // Tick at 2 GPUs, 10 M entities in total: 209 / 210 / 218 / 226 us for (1, 2) / (2, 4) / (4, 8) / (full, 8), 233 us without the overlap
    int shard_arrive_beside_ctas_per_sm{1};
    int shard_move_beside_ctas_per_sm{2};
    bool shard_arrive_early{false};   // MSIM_SHARD_ARRIVE_EARLY=1: pass B of a band-sharded tick follows the exchange on the side stream instead of waiting for the
                                  // scatter (the tick's dependency loop move -> exchange -> scan -> scatter -> pass B -> move loses its last link).  Measured
                                  // at 2 GPUs, 5 M entities each: 245 us per tick against 210 us - beside the memory-bound scan and scatter pass B costs more
                                  // than the shorter loop saves; off by default
    int scatter_beside_ctas_per_sm{8};  // MSIM_SCATTER_BESIDE_CTAS=1..8: CTAs per SM of the scatter in the pipelined rebuild (it runs beside the previous tick's query)
    bool pipeline_build{true};        // MSIM_PIPELINE_BUILD=0: scan + scatter of tick t+1 wait for the query of tick t on the main stream (by default they follow the
                                  // move phase on the side stream, into the other of two {sorted positions, prefix table} sets; unsharded handles)
    bool l2_persist_roads{false};        // MSIM_L2_PERSIST_ROADS=1: road table as a persisting L2 access-policy window on the handle's streams (api.cu)
};
const Tuning& tuning();

// ---- kernel launchers (each returns the number of kernels it launched) ------------------------
// move.cu
// pass A (streaming) and pass B (next waypoint of the arrived entities) of one move dispatch
// Device-resident counts (asynchronous sharded ticks): when `n_dev` is non-NULL a kernel takes its element
// count from *n_dev (written by an earlier kernel on the stream) and the host-side `n` is only an upper
// bound used to size the grid.
int launch_move(cudaStream_t s, int sm_count, uint32_t n, const float2* pos_in, float2* pos_out, const float2* target, uint32_t* arrived,
                uint32_t* keys /* nullable */, const GridParams& grid, uint32_t* hist /* nullable: fused digit histograms */,
                int hist_passes, uint32_t* cell_count /* nullable: fused per-cell population */, Profiler* prof,
                const uint32_t* n_dev = nullptr, const struct ShardMoveArgs* shard = nullptr, int ctas_per_sm = 0 /* 0: the stand-alone grid, 8 per SM */);
// `beside`: the pass runs on the side stream next to the issue-bound query (see Tuning::arrive_beside_ctas_per_sm)
int launch_arrive(cudaStream_t s, uint32_t n, float2* target, uint32_t* road, uint4* rng, const uint32_t* arrived, const msim_road* roads,
                  const uint32_t* connections, uint64_t connection_count, Profiler* prof, const uint32_t* n_dev = nullptr, bool beside = false,
                  int beside_ctas_per_sm = -1 /* -1: Tuning::arrive_beside_ctas_per_sm */);
int launch_keygen(cudaStream_t s, uint32_t n, const float2* pos, uint32_t* keys, const GridParams& grid, Profiler* prof);

// sort.cu
size_t sort_workspace_bytes(uint32_t capacity);
void sort_workspace_bind(SortWorkspace& ws, void* base, uint32_t capacity);
// sorts (key, index) by key; result lands in *result (either buf_a or buf_b)
int launch_sort(cudaStream_t s, uint32_t n, const uint32_t* keys, uint64_t* buf_a, uint64_t* buf_b, int key_bits,
                const SortWorkspace& ws, uint64_t** result, bool hist_ready, Profiler* prof, const uint32_t* n_dev = nullptr);
// zeroes histograms / tickets / look-back words; call before a move pass that fuses the histogram
void sort_prepare(cudaStream_t s, uint32_t n, int key_bits, const SortWorkspace& ws, Profiler* prof);

// csort.cu — single-digit radix (counting) sort over the whole cell key
uint32_t csort_tiles(uint32_t cells);
void csort_clear(cudaStream_t s, uint32_t* cell_count, uint32_t cells, Profiler* prof);
void csort_band(uint32_t cells, int ncx, uint32_t row_lo, uint32_t row_hi, int ncy, uint32_t* c0, uint32_t* c1);
int launch_cell_count(cudaStream_t s, uint32_t n, const uint32_t* keys, uint32_t* cell_count, uint32_t c0, uint32_t c1, Profiler* prof,
                      const uint32_t* n_dev = nullptr);
int launch_cell_count_pos(cudaStream_t s, int sm_count, uint32_t n, const float2* pos, uint32_t* cell_count, const GridParams& grid, Profiler* prof);
size_t csort_scan_scratch_words(uint32_t cells);
int launch_cell_scan(cudaStream_t s, uint32_t* cell_count, uint32_t cells, uint32_t* scratch, uint32_t scratch_tiles, uint32_t epoch, uint32_t* cell_start,
                     uint32_t* error_flag, Profiler* prof);
// slots come from atomics on the scanned table `cursor`, which ends up shifted by one cell.  `keys` non-NULL (sharded handles):
// keys are read instead of recomputed, entities whose key is outside [c0, c1) get the slot CSORT_SKIP, the count may live in n_dev
int launch_cell_scatter_slots(cudaStream_t s, int sm_count, uint32_t n, const float2* pos, const uint32_t* keys, uint32_t* cursor, float2* sorted_pos,
                              uint32_t* slot_of_entity, const GridParams& grid, uint32_t c0, uint32_t c1, Profiler* prof, const uint32_t* n_dev = nullptr,
                              int ctas_per_sm = 8 /* fewer while the kernel shares the SMs with a query */);
int launch_invert_slots(cudaStream_t s, uint32_t n, const uint32_t* slot_of_entity, uint32_t* sorted_idx, Profiler* prof, const uint32_t* n_dev = nullptr);
int launch_gather_flags(cudaStream_t s, uint32_t n, const uint32_t* slot_of_entity, const uint8_t* flag_sorted, uint8_t* flag_entity, Profiler* prof);

// collide.cu
int launch_build_cells(cudaStream_t s, uint32_t n, const uint64_t* sorted, const float2* pos, float2* sorted_pos, uint32_t* sorted_idx,
                       uint2* cell_range, const GridParams& grid, Counters* counters, Profiler* prof, const uint32_t* n_dev = nullptr);
// n_owned < n: slots whose entity index (sorted_idx) is >= n_owned are ghosts (neighbours only).
// Cell directory: cell_range ({first, ~end} per cell) built from the onesweep order; the counting sort has its own query (collide_tiles.cu).
int launch_query(cudaStream_t s, uint32_t n, uint32_t n_owned, const uint32_t* sorted_idx, const float2* sorted_pos, const uint2* cell_range,
                 uint8_t* flag_sorted, const GridParams& grid, bool count_pairs, Counters* counters, unsigned long long* stripes, Profiler* prof,
                 const uint32_t* n_dev = nullptr, const uint32_t* n_owned_dev = nullptr);
size_t query_stripe_bytes();
// totals of one query: flagged entities counted from the stored flags, pair stripes folded, Counters written (collide.cu)
int launch_fold_counts(cudaStream_t s, uint32_t n, const uint32_t* n_dev, const uint8_t* flag_sorted, unsigned long long* stripes, Counters* counters, Profiler* prof);
// collide_tiles.cu — the query over the counting sort's prefix table: aligned candidate groups, striped counters folded by a one-CTA
// kernel behind it.  ghosts (sharded handles): slots whose cell row is outside [row_lo, row_hi) are neighbours only; n_dev: slot count
int launch_query_tiles(cudaStream_t s, uint32_t n, const float2* sorted_pos, const uint32_t* tab, uint8_t* flag_sorted, const GridParams& grid, bool count_pairs,
                       Counters* counters, unsigned long long* stripes, Profiler* prof, const uint32_t* n_dev = nullptr, bool ghosts = false,
                       uint32_t row_lo = 0, uint32_t row_hi = 0);
int launch_scatter_flags(cudaStream_t s, uint32_t n, uint32_t n_owned, const uint32_t* sorted_idx, const uint8_t* flag_sorted, uint8_t* flag_entity, Profiler* prof);

// pack.cu
struct PackArgs {
    const float2* pos_cur;
    const float2* pos_prev;
    const float2* target;
    const uint32_t* road;
    const uint4* rng;
    const float4* color0;
    const float2* dir0;
    const uint32_t* arrived;
    const uint8_t* flag_entity;  // nullable: no collision pass since upload -> colour as uploaded
    const uint8_t* init_mask;    // nullable
    const uint32_t* slot_of;     // nullable: external id -> storage slot (cell-ordered storage)
    uint32_t initialized_all;    // value of `initialized` when init_mask == nullptr
    uint32_t has_moved;          // 0: direction = dir0
};
int launch_pack(cudaStream_t s, uint32_t first, uint32_t count, const PackArgs& a, msim_entity* dst, Profiler* prof);
int launch_unpack(cudaStream_t s, uint32_t first, uint32_t count, const msim_entity* src, float2* pos, float2* target,
                  uint32_t* road, uint4* rng, float4* color0, float2* dir0, uint8_t* init_mask, unsigned int* uninit_count, Profiler* prof);
int launch_max_road(cudaStream_t s, uint32_t n, const uint32_t* road, unsigned int* out_max);

// cell-ordered storage: permute the resident state into the order of the last collision pass
struct ReorderArrays {
    const float2* pos_prev;  float2* pos_prev_new;
    const float2* target;    float2* target_new;
    const uint32_t* road;    uint32_t* road_new;
    const uint4* rng;        uint4* rng_new;
    const uint32_t* ext_id;  uint32_t* ext_id_new;   // ext_id may be NULL (identity)
    uint32_t* slot_of;
    const uint32_t* arrived; uint32_t* arrived_new;
    uint8_t* flag_entity;
    // sharded handles (ext_id = gid, slot_of = NULL): the owned entities are the middle run of the sorted order,
    // between the ghost rows below and above the band; it starts at *first_owned (an entry of the prefix table),
    // its length is *n_owned_dev, and the current positions are copied out of sorted_pos instead of swapped in
    const uint32_t* first_owned;
    const uint32_t* n_owned_dev;
    const float2* sorted_pos;
    float2* pos_cur_new;
    uint32_t* error_word;  // |= 16 when a slot of the middle run turns out to be a ghost
};
int launch_reorder(cudaStream_t s, uint32_t n, const uint32_t* sorted_idx, const uint8_t* flag_sorted, const ReorderArrays& a, Profiler* prof);
int launch_reorder_sharded(cudaStream_t s, uint32_t n_upper, uint32_t arrived_words, const uint32_t* sorted_idx, const uint8_t* flag_sorted,
                           const ReorderArrays& a, Profiler* prof);
int launch_gather_pos(cudaStream_t s, uint32_t n, const uint32_t* slot_of, const float2* pos, float2* out);
int launch_gather_flag(cudaStream_t s, uint32_t n, const uint32_t* slot_of, const uint8_t* flag, uint8_t* out);

// quadtree.cu — histogram over the finest display-quadtree cells (levels = maxDepth - 1, <= 8)
int launch_leaf_histogram(cudaStream_t s, int sm_count, uint32_t n, const float2* pos, float world_w, float world_h, int levels, uint32_t* hist);

// ---- multi-GPU sharding (shard.cu) -------------------------------------------------------------
struct ShardHeader {
    uint32_t n_migrants;
    uint32_t n_halo;
    uint32_t overflow;
    uint32_t pad[5];
};
constexpr uint32_t MIGRANT_BYTES = 72;  // pos, pos_prev, target, rng, color0, road, gid, arrived bit
// counters of one pack (device, cleared per tick).  The fused move + pack kernels count leavers / halo entries HERE, in local
// memory, and publish the totals into the (possibly remote) buffer headers once: the exchange buffers only ever see plain
// stores, never an atomic round trip over NVLink
enum { SHARD_CTR_HOLES = 0, SHARD_CTR_LOCAL_GHOSTS = 1, SHARD_CTR_MIG_DOWN = 2, SHARD_CTR_MIG_UP = 3, SHARD_CTR_HALO_DOWN = 4, SHARD_CTR_HALO_UP = 5,
       SHARD_CTR_COMPACTED = 7,  // merged exchange kernel: CTA 0 has emptied the tail the ghosts are about to overwrite
       SHARD_CTR_COUNT = 8 };

struct ShardArrays {
    float2* pos_cur;
    float2* pos_prev;
    float2* target;
    uint32_t* road;
    uint4* rng;
    float4* color0;
    uint32_t* gid;
    uint32_t* keys;
    uint32_t* arrived;
    // per-cell population maintained through the exchange (NULL cell_count: a count kernel runs later instead): the fused
    // move + pack kernel has counted the entities that stay, so whoever is placed or appended as a ghost afterwards is
    // counted on arrival and the per-cell counters are complete when the integrate is done
    uint32_t* cell_count;
    uint32_t c0, c1;  // counted cell range; keys outside it stay out of the order
};
constexpr uint32_t CSORT_SKIP = 0xffffffffu;  // slot of an entity that is left out of the cell order

int launch_shard_reset(cudaStream_t s, void* buf_down, void* buf_up, uint32_t* ctr);
int launch_shard_pack(cudaStream_t s, const ShardArrays& a, uint32_t n, int ncx, uint32_t row_lo, uint32_t row_hi, void* buf_down, void* buf_up,
                      uint32_t mig_cap, uint32_t halo_cap, uint32_t* holes, uint32_t holes_cap, float2* local_ghosts, uint32_t* ctr, Profiler* prof,
                      const uint32_t* n_dev = nullptr, bool reset = true);
int launch_shard_signal(cudaStream_t s, uint32_t* flag_down, uint32_t* flag_up, uint32_t value);
void launch_shard_stamp(cudaStream_t s, unsigned long long* trace, int slot);  // MSIM_SHARD_TRACE=1: %globaltimer into trace[slot]
// peer-memory exchange, sender side: local send buffers -> the neighbours' receive buffers, headers, flags (on a stream of its own)
int launch_shard_push(cudaStream_t s, const void* send_down, const void* send_up, void* peer_down, void* peer_up, uint32_t* flag_down, uint32_t* flag_up,
                      uint32_t signal_value, uint32_t mig_cap, uint32_t halo_cap, uint32_t holes_cap, uint32_t* ctr, uint32_t* error_word, uint32_t* ticket,
                      bool counts_in_headers);
// device-side integrate (asynchronous sharded tick): placement, tail compaction and the new counts without a host round trip
// Two sets of DEV_COUNT_WORDS words that alternate from one integrate to the next (shard.cu IntegrateArgs), then the sticky error word.
enum { DEV_N_OWNED = 0, DEV_N_GHOST = 1, DEV_N_TOTAL = 2, DEV_HALO_DOWN = 4, DEV_HALO_UP = 5, DEV_COUNT_WORDS = 8, DEV_SHARD_ERROR = 2 * DEV_COUNT_WORDS,
       DEV_ALLOC_WORDS = 3 * DEV_COUNT_WORDS };
struct ShardCounts {
    const uint32_t* in;    // set the previous integrate (or the host) wrote
    uint32_t* out;         // the other set
    uint32_t* error_word;  // sticky error bits
};
// device-side wait of the peer-memory exchange: the integrate kernel spins until both flag words reach `expected`
struct ShardWait {
    const uint32_t* flag_down;  // NULL: nothing to wait for on that side
    const uint32_t* flag_up;
    uint32_t expected;
    unsigned long long timeout_ns;
};
int launch_shard_integrate_device(cudaStream_t s, const ShardArrays& a, const ShardCounts& counts, const void* sent_down, const void* sent_up,
                                  const void* recv_down, const void* recv_up, const uint32_t* holes, const uint32_t* ctr, const float2* local_ghosts,
                                  uint32_t mig_cap, uint32_t halo_cap, uint32_t holes_cap, uint32_t entity_cap, uint32_t* scratch_bits, uint2* scratch_moves,
                                  const GridParams& grid, Profiler* prof, const ShardWait* wait = nullptr);
int launch_shard_exchange(cudaStream_t s, const ShardArrays& a, const ShardCounts& counts, const void* recv_down, const void* recv_up,
                          const uint32_t* holes, const uint32_t* ctr, const float2* local_ghosts, uint32_t mig_cap, uint32_t halo_cap, uint32_t holes_cap,
                          uint32_t entity_cap, uint32_t* scratch_bits, uint2* scratch_moves, const GridParams& grid, Profiler* prof, const ShardWait& wait,
                          unsigned long long* trace = nullptr);
int launch_shard_place(cudaStream_t s, const ShardArrays& a, const void* recv_down, uint32_t n_down, const void* recv_up, uint32_t n_up,
                       const uint32_t* dst, const GridParams& grid, Profiler* prof);
int launch_shard_relocate(cudaStream_t s, const ShardArrays& a, const uint2* moves, uint32_t count, Profiler* prof);
int launch_shard_append_ghosts(cudaStream_t s, const ShardArrays& a, uint32_t first, const void* recv_down, uint32_t h_down, const void* recv_up,
                               uint32_t h_up, const float2* local_ghosts, uint32_t h_local, uint32_t mig_cap, const GridParams& grid, Profiler* prof);
int launch_shard_row_histogram(cudaStream_t s, const uint32_t* keys, uint32_t n, int ncx, uint32_t* rows, uint32_t nrows, Profiler* prof);

// ---- device helpers shared by several translation units ---------------------------------------
// (MSIM_HOST_EMU: tests/cuda_emu compiles the kernels for a host SIMT emulator; it has twins for the few inline-PTX helpers)
#if defined(__CUDACC__) || defined(MSIM_HOST_EMU)
__device__ __forceinline__ uint32_t cell_key_of(float2 p, const GridParams& g) {
    // single multiply per axis (no FMA involved), floor, clamp: monotone in each coordinate
    int cx = __float2int_rd(__fmul_rn(p.x, g.inv_cell));
    int cy = __float2int_rd(__fmul_rn(p.y, g.inv_cell));
    cx = min(max(cx, 0), g.ncx - 1);
    cy = min(max(cy, 0), g.ncy - 1);
    return static_cast<uint32_t>(cy) * static_cast<uint32_t>(g.ncx) + static_cast<uint32_t>(cx);
}

// exchange buffer accessors and the warp-aggregated list append shared by shard.cu and the fused move kernel
__device__ __forceinline__ ShardHeader* header_of(void* buf) { return static_cast<ShardHeader*>(buf); }
__device__ __forceinline__ uint2* records_of(void* buf) { return reinterpret_cast<uint2*>(static_cast<char*>(buf) + sizeof(ShardHeader)); }
__device__ __forceinline__ float2* halo_of(void* buf, uint32_t mig_cap) {
    return reinterpret_cast<float2*>(static_cast<char*>(buf) + sizeof(ShardHeader) + static_cast<size_t>(mig_cap) * MIGRANT_BYTES);
}
// returns this lane's slot in a list whose length lives at *counter (one atomic per warp); whole warp must call
__device__ __forceinline__ uint32_t warp_append(bool want, uint32_t* counter) {
    const uint32_t m = __ballot_sync(0xffffffffu, want);
    if (m == 0) return 0;
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (static_cast<int>(lane) == leader) base = atomicAdd(counter, static_cast<uint32_t>(__popc(m)));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

// system-scope flag accesses of the peer-memory exchange (the flag lives in another GPU's memory or is written by one)
#ifndef MSIM_HOST_EMU
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif

// ---- behind the fused move + pack kernel (move.cu shard_publish_kernel): one thread -----------------------------------------------
// Collective exchange: writes the list lengths into the headers of the send buffers the move kernel has filled (every counter is local).
// (The peer-memory exchange has its own sender, shard.cu shard_push_kernel, which writes the headers where they are read.)
#ifndef MSIM_HOST_EMU
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
#endif
__device__ __forceinline__ void shard_publish(const ShardMoveArgs& sh) {
    const uint32_t holes_total = __ldcg(sh.ctr + SHARD_CTR_HOLES);
    sh.ctr[SHARD_CTR_LOCAL_GHOSTS] = min(holes_total, sh.holes_cap);
    if (holes_total > sh.holes_cap) atomicOr(sh.error_word, 2u);
    if (sh.buf_down) {
        const uint32_t m = __ldcg(sh.ctr + SHARD_CTR_MIG_DOWN), hl = __ldcg(sh.ctr + SHARD_CTR_HALO_DOWN);
        ShardHeader* hd = header_of(sh.buf_down);
        hd->n_migrants = m;
        hd->n_halo = hl;
        hd->overflow = (m > sh.mig_cap || hl > sh.halo_cap) ? 1u : 0u;
    }
    if (sh.buf_up) {
        const uint32_t m = __ldcg(sh.ctr + SHARD_CTR_MIG_UP), hl = __ldcg(sh.ctr + SHARD_CTR_HALO_UP);
        ShardHeader* hd = header_of(sh.buf_up);
        hd->n_migrants = m;
        hd->n_halo = hl;
        hd->overflow = (m > sh.mig_cap || hl > sh.halo_cap) ? 1u : 0u;
    }
    __threadfence_system();  // headers (and everything the finished move kernel stored) before the flags
    if (sh.peer_flag_down) st_relaxed_sys(sh.peer_flag_down, sh.signal_value);
    if (sh.peer_flag_up) st_relaxed_sys(sh.peer_flag_up, sh.signal_value);
}

// bit position of entity e inside the `arrived` mask written by the move kernel: entities are
// processed as float4 pairs (2*lane, 2*lane+1) of a 64-entity warp chunk; each parity has its own word.
__device__ __forceinline__ uint32_t arrived_word(uint32_t e) { return (e >> 6) * 2u + (e & 1u); }
__device__ __forceinline__ uint32_t arrived_bit(uint32_t e) { return (e & 63u) >> 1; }
#endif

}  // namespace msim
