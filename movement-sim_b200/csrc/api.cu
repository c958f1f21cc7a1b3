// api.cu — the C ABI of libmsim_cuda.so (include/msim.h): handle lifetime, HBM allocation, the
// dispatch state machine of the reference's host driver, and the AoS<->SoA boundary copies.
//
// Reference behaviour mirrored here (/root/reference/src/sim/Simulator.cpp):
//   :52-103   buffer creation + push constants      -> msim_create
//   :191-192  one-off upload                        -> msim_create / msim_upload_entities
//   :220-235  two blocking dispatches per sim tick  -> msim_dispatch (tick parity, shader :860-879)
//   :248-273  readbacks                             -> msim_read_*
// No CPU fallback exists: without a usable sm_100 device every compute entry point fails.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/msim_shard.h"
#include "msim_internal.h"

using namespace msim;

namespace {
thread_local std::string g_create_error;

constexpr uint64_t MAX_ENTITIES_PER_HANDLE = 1ull << 30;  // look-back words carry 30-bit counts
constexpr uint32_t MAX_GRID_CELLS = 1u << 27;
constexpr uint32_t STAGE_ENTITIES = 1u << 20;  // 64 MiB AoS staging chunk
// counting sort: counter + prefix tables of 2 x 512 MiB at most (2^27 cells, the largest grid the library accepts) unless the limit is lowered
inline uint32_t csort_max_cells() { return 1u << msim::tuning().csort_max_cells_log2; }
}  // namespace

struct msim_handle {
    int device{0};
    int sm_count{148};
    cudaStream_t stream{nullptr};
    bool own_stream{false};
    cudaStream_t side{nullptr};  // pass B of a move runs here while the collision pass uses `stream`
    cudaEvent_t ev_moved{nullptr}, ev_arrived{nullptr};
    bool arrive_early{false};     // pass B of this tick was launched behind the exchange; the collision pass only has to order the next move behind its scatter
    bool side_pending{false};     // the side stream holds work the main stream has not been ordered behind yet (ev_arrived marks its end)
    bool arrive_deferred{false};  // pass B of the last move has not been launched yet (it will ride beside the query)
    // Overlapped ticks: the move phase of tick t+1 (pass B of tick t, move + pack, shard exchange) runs on the side stream while the
    // query of tick t runs on the main one - movement never reads collision results, and the query only reads what the scatter copied.
    cudaStream_t ms{nullptr};     // stream of the move phase in flight (side or main)
    bool main_touched{true};      // the main stream has touched entity state since the side stream was last ordered behind it
    uint32_t flags{0};

    uint32_t n{0};
    uint32_t cap{0};  // padded to a multiple of 64 entities
    float world_w{0}, world_h{0}, radius{0};
    uint32_t qt_depth{8}, qt_cap{10};

    // resident state
    float2* pos[2]{nullptr, nullptr};
    int cur{0};
    float2* target{nullptr};
    uint32_t* road{nullptr};
    uint4* rng{nullptr};
    float4* color0{nullptr};
    float2* dir0{nullptr};
    uint32_t* arrived{nullptr};
    msim_road* roads{nullptr};
    uint32_t* conn{nullptr};
    uint64_t road_count{0}, conn_count{0};

    // neighbour structure
    uint32_t* keys{nullptr};
    uint64_t* sort_a{nullptr};
    uint64_t* sort_b{nullptr};
    uint64_t* sorted{nullptr};
    float2* sorted_pos{nullptr};
    // pipelined rebuild (unsharded handles): a second {sorted positions, prefix table} set, so that scan + scatter of tick t+1 can run
    // (side stream) while the query of tick t (main stream) still reads the first; the pointers above always name the set in use
    float2* sorted_pos_alt{nullptr};
    uint32_t* cell_table_alt{nullptr};
    cudaEvent_t ev_built{nullptr}, ev_queried[2]{nullptr, nullptr};
    bool queried_recorded[2]{false, false};
    uint32_t build_set{0};
    uint2* cell_range{nullptr};   // onesweep path: {first, ~end} per cell
    uint32_t cell_capacity{0};
    uint32_t* cell_count{nullptr};  // counting-sort path: per-cell counters, prefix table, scan scratch, ranks
    uint32_t* cell_table{nullptr};  // allocation behind cell_start: cell_start = cell_table + 4, and cell_table[3] is a permanent 0
    uint32_t* cell_start{nullptr};
    uint32_t* tile_sums{nullptr};   // scratch of the single-pass scan (csort.cu ScanState)
    uint32_t scan_tiles_cap{0};
    uint32_t scan_epoch{0};
    uint32_t* rank{nullptr};
    uint32_t* sorted_idx{nullptr};
    bool use_csort{false};
    bool counts_valid{false};
    bool counts_dirty{true};   // the counter table is not all-zero (every scan leaves it zeroed: no per-tick memset in steady state)
    uint8_t* flag_sorted{nullptr};
    uint8_t* flag_entity{nullptr};
    void* sort_mem{nullptr};
    SortWorkspace ws{};
    GridParams grid{};
    int key_bits{1};

    Counters* counters{nullptr};
    unsigned long long* stripes{nullptr};  // striped per-query counters (collide.cu), followed by the query's completion ticket
    bool slots_valid{false};               // h->rank holds entity -> sorted slot of the last collision pass (tiles path)
    unsigned int* scratch{nullptr};  // [0] uninitialised count, [1] max road index
    uint32_t* leaf_hist{nullptr};    // display quadtree: entities per finest cell
    msim_entity* stage{nullptr};

    bool uninitialised{false};
    bool has_moved{false};
    bool keys_valid{false};
    bool hist_valid{false};
    bool collided{false};
    bool flags_scattered{false};
    uint64_t move_passes{0}, collide_passes{0}, launches{0}, initialised_total{0};
    uint64_t last_pairs{0}, total_pairs{0}, last_flagged{0}, total_flagged{0};

    // cell-ordered storage: periodic physical re-sort of the state (pack.cu), slot <-> external id maps
    bool reorder_enabled{false};
    bool perm_active{false};
    uint32_t reorder_every{32};
    uint32_t since_reorder{0};
    uint64_t reorders{0};
    float2* pos_spare{nullptr};
    float2* target_alt{nullptr};
    uint32_t* road_alt{nullptr};
    uint4* rng_alt{nullptr};
    uint32_t* arrived_alt{nullptr};
    uint32_t* ext_id{nullptr};
    uint32_t* ext_id_alt{nullptr};
    uint32_t* slot_of{nullptr};

    // multi-GPU sharding (msim_shard.h): band of cell rows per handle, ghosts behind the owned entities
    bool sharded{false};
    bool flags_stale{false};
    uint32_t n_ghost{0};
    uint32_t collide_total{0}, collide_owned{0};
    uint32_t mig_cap{0}, halo_cap{0}, holes_cap{0};
    uint32_t* gid{nullptr};
    uint32_t* holes{nullptr};
    float2* local_ghosts{nullptr};
    uint32_t* shard_ctr{nullptr};
    uint32_t* place_dst{nullptr};
    uint2* moves{nullptr};
    uint32_t* row_hist{nullptr};
    uint32_t row_hist_rows{0};  // rows the allocation holds (the grid can grow through push constants)
    uint32_t* host_stage{nullptr};  // pinned, two halves
    bool stage_flip{false};
    void* sent_down{nullptr};
    void* sent_up{nullptr};
    uint32_t* gid_alt{nullptr};
    // peer-memory exchange (msim_shard_p2p_*): our receive arena, the neighbours' arenas as peer pointers
    char* p2p_arena{nullptr};
    size_t p2p_buf_bytes{0};
    char* p2p_peer_down{nullptr};
    char* p2p_peer_up{nullptr};
    char* p2p_send{nullptr};             // local send buffers, index 2 * parity + side: the move kernel fills them, the push kernel copies them out
    uint32_t* p2p_push_ticket{nullptr};
    cudaStream_t push_stream{nullptr};   // the push kernel's own stream: its remote stores complete off the tick's critical path
    cudaEvent_t ev_packed_move{nullptr}, ev_pushed{nullptr};
    bool push_pending{false};
    bool p2p_down_ipc{false}, p2p_up_ipc{false};
    bool p2p_connected{false};
    uint32_t p2p_tick{0};
    unsigned long long p2p_timeout_ns{10000000000ull};
    uint32_t band_lo{0}, band_hi{0};  // cell rows that can hold this handle's keys after the last pack + integrate (ghost rows excluded)
    bool packed{false};
    bool count_fused{false};          // the last move + pack ranked the stayers: place / relocate / ghost kernels complete the per-cell counters
    bool awaiting_integrate{false};   // a fused move + pack has run: pass B stays deferred until the exchange has been integrated
    bool band_valid{false};           // false between a move pass and the integrate that follows it
    unsigned long long* shard_trace{nullptr};  // MSIM_SHARD_TRACE=1: phase times of the exchange kernel, printed at msim_destroy
    uint32_t* dev_counts{nullptr};  // device-resident {owned, ghosts, total, ...} in two alternating sets + sticky error bits: what asynchronous sharded ticks run on
    uint32_t counts_set{0};         // the set that holds the current counts (msim_internal.h ShardCounts)
    bool async_counts{false};       // host-side n / n_ghost are stale (upper bound = capacity) until the next refresh

    // asynchronous readback (msim_snapshot_*): device image of the AoS state, two pinned host buffers used alternately
    msim_entity* snap_dev{nullptr};
    msim_entity* snap_host[2]{nullptr, nullptr};
    size_t snap_cap{0};
    uint64_t snap_count{0};
    int snap_slot{0};
    bool snap_pending{false};
    cudaStream_t copy_stream{nullptr};
    cudaEvent_t ev_packed{nullptr}, ev_copied{nullptr};

    Profiler prof;
    std::string error;
};

namespace msim {
const Tuning& tuning() {
    static const Tuning t = [] {
        Tuning v;
        if (const char* e = std::getenv("MSIM_L2_PERSIST_ROADS")) v.l2_persist_roads = std::atoi(e) == 1;
        if (const char* e = std::getenv("MSIM_ARRIVE_BESIDE_CTAS")) {
            const int k = std::atoi(e);
            if (k >= 0 && k <= 8) v.arrive_beside_ctas_per_sm = k;
        }
        if (const char* e = std::getenv("MSIM_OVERLAP_TICKS")) v.overlap_ticks = std::atoi(e) != 0;
        if (const char* e = std::getenv("MSIM_PIPELINE_BUILD")) v.pipeline_build = std::atoi(e) != 0;
        if (const char* e = std::getenv("MSIM_SCATTER_BESIDE_CTAS")) {
            const int k = std::atoi(e);
            if (k >= 1 && k <= 8) v.scatter_beside_ctas_per_sm = k;
        }
        if (const char* e = std::getenv("MSIM_SHARD_ARRIVE_EARLY")) v.shard_arrive_early = std::atoi(e) != 0;
        if (const char* e = std::getenv("MSIM_MOVE_BESIDE_CTAS")) {
            const int k = std::atoi(e);
            if (k >= 1 && k <= 8) v.move_beside_ctas_per_sm = k;
        }
        if (const char* e = std::getenv("MSIM_SHARD_ARRIVE_BESIDE_CTAS")) {
            const int k = std::atoi(e);
            if (k >= 0 && k <= 8) v.shard_arrive_beside_ctas_per_sm = k;
        }
        if (const char* e = std::getenv("MSIM_SHARD_MOVE_BESIDE_CTAS")) {
            const int k = std::atoi(e);
            if (k >= 1 && k <= 8) v.shard_move_beside_ctas_per_sm = k;
        }
        if (const char* e = std::getenv("MSIM_CSORT_MAX_CELLS_LOG2")) {
            const int k = std::atoi(e);
            if (k >= 25 && k <= 27) v.csort_max_cells_log2 = k;
        }
        return v;
    }();
    return t;
}
}  // namespace msim

namespace {

int fail(msim_handle* h, int code, const std::string& msg) {
    if (h) h->error = msg;
    else g_create_error = msg;
    return code;
}

#define MSIM_CUDA(h, call)                                                                                         \
    do {                                                                                                           \
        cudaError_t err__ = (call);                                                                                \
        if (err__ != cudaSuccess) {                                                                                \
            return fail((h), err__ == cudaErrorMemoryAllocation ? MSIM_ERR_OOM : MSIM_ERR_CUDA,                    \
                        std::string(#call) + ": " + cudaGetErrorString(err__));                                    \
        }                                                                                                          \
    } while (0)

// smallest binary32 T with sqrtf(T) >= r; then (d2 < T) <=> (sqrtf(d2) < r) because sqrtf is monotone
// and correctly rounded on both host and device.
float exact_hit_threshold(float r) {
    if (!(r > 0.0f)) return 0.0f;
    float t = static_cast<float>(static_cast<double>(r) * static_cast<double>(r));
    while (t > 0.0f && std::sqrt(t) >= r) t = std::nextafter(t, 0.0f);
    while (std::sqrt(t) < r) t = std::nextafter(t, INFINITY);
    return t;
}

// Cell edge slightly above the radius: two points closer than r must land in adjacent cells even
// after the rounding of pos * inv_cell (error <= cells_per_axis * 2^-24 cell units).
void compute_grid(float world_w, float world_h, float radius, GridParams& g, int& key_bits) {
    g = GridParams{};
    g.radius = radius;
    g.hit_threshold = exact_hit_threshold(radius);
    double r = radius > 0.0f ? static_cast<double>(radius) : 1.0;
    double w = world_w > 0.0f ? world_w : 1.0, hh = world_h > 0.0f ? world_h : 1.0;
    double cell = r;
    for (int iter = 0; iter < 64; iter++) {
        const double axis = std::max(w, hh) / cell + 2.0;
        const double eps = std::max(1.0 / 1024.0, axis / 2097152.0);  // >= 4x the rounding bound
        const double c = cell * (1.0 + eps);
        const double ncx = std::floor(w / c) + 1.0, ncy = std::floor(hh / c) + 1.0;
        if (ncx * ncy <= static_cast<double>(MAX_GRID_CELLS)) {
            g.inv_cell = static_cast<float>(1.0 / c);
            g.ncx = static_cast<int>(ncx);
            g.ncy = static_cast<int>(ncy);
            break;
        }
        cell *= 2.0;
    }
    if (g.ncx < 1) g.ncx = 1;
    if (g.ncy < 1) g.ncy = 1;
    g.ncells = static_cast<uint32_t>(g.ncx) * static_cast<uint32_t>(g.ncy);
    int bits = 1;
    while (bits < 32 && (1ull << bits) < g.ncells) bits++;
    key_bits = bits;
}

void configure_grid(msim_handle* h) { compute_grid(h->world_w, h->world_h, h->radius, h->grid, h->key_bits); }

template <typename T>
cudaError_t dev_alloc(T** p, size_t count) {
    return cudaMalloc(reinterpret_cast<void**>(p), count ? count * sizeof(T) : sizeof(T));
}

void free_all(msim_handle* h) {
    cudaSetDevice(h->device);
    cudaFree(h->pos[0]); cudaFree(h->pos[1]); cudaFree(h->target); cudaFree(h->road); cudaFree(h->rng);
    cudaFree(h->color0); cudaFree(h->dir0); cudaFree(h->arrived); cudaFree(h->roads); cudaFree(h->conn);
    cudaFree(h->keys); cudaFree(h->sort_a); cudaFree(h->sort_b); cudaFree(h->sorted_pos); cudaFree(h->cell_range);
    cudaFree(h->cell_count); cudaFree(h->cell_table); cudaFree(h->tile_sums); cudaFree(h->rank); cudaFree(h->sorted_idx);
    cudaFree(h->flag_sorted); cudaFree(h->flag_entity); cudaFree(h->sort_mem); cudaFree(h->counters);
    cudaFree(h->scratch); cudaFree(h->stage); cudaFree(h->stripes); cudaFree(h->leaf_hist);
    cudaFree(h->pos_spare); cudaFree(h->target_alt); cudaFree(h->road_alt); cudaFree(h->rng_alt); cudaFree(h->arrived_alt);
    cudaFree(h->ext_id); cudaFree(h->ext_id_alt); cudaFree(h->slot_of);
    cudaFree(h->dev_counts); cudaFree(h->gid_alt);
    if (h->p2p_peer_down && h->p2p_down_ipc) cudaIpcCloseMemHandle(h->p2p_peer_down);
    if (h->p2p_peer_up && h->p2p_up_ipc) cudaIpcCloseMemHandle(h->p2p_peer_up);
    cudaFree(h->p2p_arena); cudaFree(h->p2p_send); cudaFree(h->p2p_push_ticket);
    if (h->push_stream) cudaStreamDestroy(h->push_stream);
    if (h->ev_packed_move) cudaEventDestroy(h->ev_packed_move);
    if (h->ev_pushed) cudaEventDestroy(h->ev_pushed);
    cudaFree(h->gid); cudaFree(h->holes); cudaFree(h->local_ghosts); cudaFree(h->shard_ctr); cudaFree(h->place_dst);
    cudaFree(h->moves); cudaFree(h->row_hist);
    if (h->host_stage) cudaFreeHost(h->host_stage);
    if (h->copy_stream) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamDestroy(h->copy_stream);
    }
    cudaFree(h->snap_dev);
    for (msim_entity* p : h->snap_host)
        if (p) cudaFreeHost(p);
    if (h->ev_packed) cudaEventDestroy(h->ev_packed);
    if (h->ev_copied) cudaEventDestroy(h->ev_copied);
    for (cudaEvent_t e : h->prof.pool) cudaEventDestroy(e);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_moved) cudaEventDestroy(h->ev_moved);
    if (h->ev_arrived) cudaEventDestroy(h->ev_arrived);
    cudaFree(h->sorted_pos_alt); cudaFree(h->cell_table_alt);
    if (h->ev_built) cudaEventDestroy(h->ev_built);
    for (cudaEvent_t e : h->ev_queried)
        if (e) cudaEventDestroy(e);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
}

// make the main stream wait for a pass B still running on the side stream
// element counts for launches: exact host values, or (asynchronous sharded ticks) the capacity as an upper
// bound for grid sizing plus device pointers the kernels read the exact counts from
inline uint32_t launch_owned(const msim_handle* h) { return h->async_counts ? h->cap : h->n; }
inline uint32_t launch_total(const msim_handle* h) { return h->async_counts ? h->cap : h->n + h->n_ghost; }
inline uint32_t* cur_counts(const msim_handle* h) { return h->dev_counts + h->counts_set * DEV_COUNT_WORDS; }
inline const uint32_t* dev_owned(const msim_handle* h) { return h->async_counts ? cur_counts(h) + DEV_N_OWNED : nullptr; }
inline const uint32_t* dev_total(const msim_handle* h) { return h->async_counts ? cur_counts(h) + DEV_N_TOTAL : nullptr; }
inline uint32_t* dev_error(const msim_handle* h) { return h->dev_counts + DEV_SHARD_ERROR; }
// counts for an integrate: reads the current set, writes the other one, which becomes current for everything enqueued behind it
inline ShardCounts flip_counts(msim_handle* h) {
    ShardCounts c{cur_counts(h), nullptr, dev_error(h)};
    h->counts_set ^= 1u;
    c.out = cur_counts(h);
    return c;
}

void launch_deferred_arrive(msim_handle* h, bool beside) {
    if (h->arrive_early && beside && h->side) {
        // pass B is already on the side stream, right behind the exchange (msim_shard_p2p_integrate).  What is left to do here, behind the
        // scatter: the next move pass (side stream) rewrites the keys the scatter has just read
        h->arrive_early = false;
        cudaEventRecord(h->ev_moved, h->stream);
        cudaStreamWaitEvent(h->side, h->ev_moved, 0);
        h->main_touched = false;
        cudaEventRecord(h->ev_arrived, h->side);
        h->side_pending = true;
        return;
    }
    h->arrive_early = false;
    if (!h->arrive_deferred) return;
    if (h->awaiting_integrate) return;  // migrants travel with their pre-arrival state: pass B must see the integrated population
    h->arrive_deferred = false;
    if (beside && h->side) {
        // pass B (dependent gathers, latency-bound, few issue slots) runs on the side stream from here on:
        // called right before the query, which is issue-bound and leaves the memory system idle
        cudaEventRecord(h->ev_moved, h->stream);
        cudaStreamWaitEvent(h->side, h->ev_moved, 0);
        h->main_touched = false;  // the side stream is now ordered behind everything the main stream has done to the state (up to the scatter)
        h->launches += launch_arrive(h->side, launch_owned(h), h->target, h->road, h->rng, h->arrived, h->roads, h->conn, h->conn_count, &h->prof,
                                     dev_owned(h), /*beside=*/true, h->sharded ? tuning().shard_arrive_beside_ctas_per_sm : -1);
        cudaEventRecord(h->ev_arrived, h->side);
        h->side_pending = true;
    } else {
        h->launches += launch_arrive(h->stream, launch_owned(h), h->target, h->road, h->rng, h->arrived, h->roads, h->conn, h->conn_count, &h->prof,
                                     dev_owned(h));
    }
}


// MSIM_L2_PERSIST_ROADS=1 (opt-in, not yet run on hardware): the road table (32 B per road, 22 MB for the Munich stand-in) is declared a
// persisting L2 access-policy window on both streams, so that pass B's dependent gathers (two 16-byte loads of the current road, two of the
// next one) hit L2 although every tick streams ~0.8 GB through it (ncu: 30 % L2 hit rate in arrive_kernel).  Best effort: a device or stream
// that refuses the attribute leaves everything as it was.
void apply_l2_window(msim_handle* h) {
    if (!tuning().l2_persist_roads || !h->roads || h->road_count == 0) return;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0 || prop.accessPolicyMaxWindowSize <= 0) {
        cudaGetLastError();
        return;
    }
    const size_t table = sizeof(msim_road) * static_cast<size_t>(h->road_count);
    const size_t carve = std::min<size_t>(table, static_cast<size_t>(prop.persistingL2CacheMaxSize));
    const size_t window = std::min<size_t>(table, static_cast<size_t>(prop.accessPolicyMaxWindowSize));
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = h->roads;
    attr.accessPolicyWindow.num_bytes = window;
    attr.accessPolicyWindow.hitRatio = window ? std::min(1.0f, static_cast<float>(static_cast<double>(carve) / static_cast<double>(window))) : 0.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (h->stream) cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    if (h->side) cudaStreamSetAttribute(h->side, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}

// main stream behind the side stream (pass B, an overlapped move phase), without launching anything
void join_side_work(msim_handle* h) {
    if (h->side_pending) {
        cudaStreamWaitEvent(h->stream, h->ev_arrived, 0);
        h->side_pending = false;
    }
}

// everything of the resident state is current and owned by the main stream: what readbacks, uploads and re-sorts start with
void join_side(msim_handle* h) {
    join_side_work(h);
    launch_deferred_arrive(h, false);  // (on the main stream, behind the move it belongs to)
    h->main_touched = true;
}

// Stream of the move phase that is about to be enqueued.  Overlapped (side stream): ordered behind the main stream only as far as
// needed - the collision pass recorded ev_moved right behind its scatter, so the query that follows it keeps running beside us.
cudaStream_t begin_move_phase(msim_handle* h, bool overlap) {
    if (!overlap) {
        join_side(h);  // the previous pass B must have rewritten the targets before they are read again
        h->ms = h->stream;
        return h->ms;
    }
    if (h->main_touched) {  // a readback, an upload, a re-sort ... since the side stream last waited for the main one
        cudaEventRecord(h->ev_moved, h->stream);
        cudaStreamWaitEvent(h->side, h->ev_moved, 0);
        h->main_touched = false;
    }
    if (h->arrive_deferred && !h->awaiting_integrate) {  // two moves in a row: pass B of the first one, in front of the second
        h->arrive_deferred = false;
        h->launches += launch_arrive(h->side, launch_owned(h), h->target, h->road, h->rng, h->arrived, h->roads, h->conn, h->conn_count, &h->prof, dev_owned(h));
    }
    h->ms = h->side;
    return h->ms;
}

void end_move_phase(msim_handle* h) {
    if (h->ms == h->side && h->side) {
        cudaEventRecord(h->ev_arrived, h->side);
        h->side_pending = true;
    }
}

int write_dev_counts(msim_handle* h) {
    if (!h->dev_counts) return MSIM_OK;
    uint32_t v[DEV_ALLOC_WORDS] = {0};  // (the sticky error word is cleared as well: the population is new)
    h->counts_set = 0;
    v[DEV_N_OWNED] = h->n;
    v[DEV_N_GHOST] = h->n_ghost;
    v[DEV_N_TOTAL] = h->n + h->n_ghost;
    MSIM_CUDA(h, cudaMemcpyAsync(h->dev_counts, v, sizeof(v), cudaMemcpyHostToDevice, h->stream));  // pageable source: staged before return
    return MSIM_OK;
}

// asynchronous sharded ticks: bring the host-side counts up to date (one small D2H + stream sync)
int refresh_counts(msim_handle* h) {
    if (!h->async_counts) return MSIM_OK;
    join_side_work(h);  // the exchange that wrote the counts may have run beside the last query
    uint32_t all[DEV_ALLOC_WORDS] = {0};
    MSIM_CUDA(h, cudaMemcpyAsync(all, h->dev_counts, sizeof(all), cudaMemcpyDeviceToHost, h->stream));
    MSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    const uint32_t* v = all + h->counts_set * DEV_COUNT_WORDS;
    const uint32_t errors = all[DEV_SHARD_ERROR];
    h->n = v[DEV_N_OWNED];
    h->n_ghost = v[DEV_N_GHOST];
    h->async_counts = false;
    if (!h->flags_stale) {  // counts only change in an integrate, which marks the flags stale: these are the last collision pass's
        h->collide_owned = h->n;
        h->collide_total = h->n + h->n_ghost;
    }
    if (errors & 32u) return fail(h, MSIM_ERR_INTERNAL, "peer-memory exchange: a neighbour never signalled its tick (timeout)");
    if (errors & 7u) return fail(h, MSIM_ERR_CAPACITY, "shard exchange overflow (migrant / halo / entity capacity) during asynchronous ticks");
    if (errors) return fail(h, MSIM_ERR_INTERNAL, "shard compaction bookkeeping mismatch");
    return MSIM_OK;
}

// The scan that consumes the per-cell counters also zeroes them (csort.cu), so between two collision passes the table is
// all zero and the next count needs no memset.  Only a count that was never scanned (a move pass that fused the count and
// was not followed by a collision pass, an upload in between, ...) leaves it dirty: then the whole table is cleared.
void prepare_counts(msim_handle* h, cudaStream_t s = nullptr) {
    if (h->counts_dirty) csort_clear(s ? s : h->stream, h->cell_count, h->cell_capacity, &h->prof);  // the whole allocation: an earlier, larger grid may have counted beyond ncells
    h->counts_dirty = true;  // about to be counted into
}

// exclusive prefix sum of the per-cell counters over cells [c0, c1) into cell_start (one kernel, csort.cu)
int scan_cells(msim_handle* h, uint32_t c0, uint32_t c1, cudaStream_t s = nullptr) {
    if (++h->scan_epoch == 0u) ++h->scan_epoch;
    return launch_cell_scan(s ? s : h->stream, h->cell_count + c0, c1 - c0, h->tile_sums, h->scan_tiles_cap, h->scan_epoch, h->cell_start + c0, &h->counters->error_flag,
                            &h->prof);
}

int ensure_cells(msim_handle* h) {
    // counting sort: on request, or by default whenever the storage is kept in cell order
    const bool want_counting = (h->flags & MSIM_FLAG_SORT_COUNTING) || (h->reorder_enabled && !(h->flags & MSIM_FLAG_SORT_ONESWEEP));
    h->use_csort = want_counting && h->grid.ncells <= csort_max_cells();
    if (h->grid.ncells <= h->cell_capacity && (h->use_csort ? h->cell_count != nullptr : h->cell_range != nullptr)) return MSIM_OK;
    cudaFree(h->cell_range); cudaFree(h->cell_count); cudaFree(h->cell_table); cudaFree(h->tile_sums); cudaFree(h->cell_table_alt);
    h->cell_range = nullptr; h->cell_count = nullptr; h->cell_table = nullptr; h->cell_start = nullptr; h->tile_sums = nullptr; h->cell_table_alt = nullptr;
    h->cell_capacity = 0;
    if (h->use_csort) {
        MSIM_CUDA(h, dev_alloc(&h->cell_count, static_cast<size_t>(h->grid.ncells) + 1));
        h->counts_dirty = true;
        // four words in front of the table: the tiles path reads the table one word earlier (csort.cu), that word stays 0
        MSIM_CUDA(h, dev_alloc(&h->cell_table, static_cast<size_t>(h->grid.ncells) + 8));
        MSIM_CUDA(h, cudaMemsetAsync(h->cell_table, 0, 4 * sizeof(uint32_t), h->stream));
        h->cell_start = h->cell_table + 4;
        h->scan_tiles_cap = csort_tiles(h->grid.ncells);
        MSIM_CUDA(h, dev_alloc(&h->tile_sums, csort_scan_scratch_words(h->grid.ncells)));
        MSIM_CUDA(h, cudaMemsetAsync(h->tile_sums, 0, csort_scan_scratch_words(h->grid.ncells) * sizeof(uint32_t), h->stream));
    } else {
        MSIM_CUDA(h, dev_alloc(&h->cell_range, h->grid.ncells));
    }
    h->cell_capacity = h->grid.ncells;
    h->counts_valid = false;
    return MSIM_OK;
}

int alloc_collision_buffers(msim_handle* h) {
    if (h->keys) return MSIM_OK;
    MSIM_CUDA(h, dev_alloc(&h->keys, h->cap));
    MSIM_CUDA(h, dev_alloc(&h->sort_a, h->cap));
    MSIM_CUDA(h, dev_alloc(&h->sort_b, h->cap));
    MSIM_CUDA(h, dev_alloc(&h->sorted_pos, h->cap));
    MSIM_CUDA(h, dev_alloc(&h->rank, h->cap));
    MSIM_CUDA(h, dev_alloc(&h->sorted_idx, h->cap));
    MSIM_CUDA(h, dev_alloc(&h->flag_sorted, h->cap));
    MSIM_CUDA(h, dev_alloc(&h->flag_entity, h->cap));
    MSIM_CUDA(h, cudaMemsetAsync(h->flag_entity, 0, h->cap, h->stream));
    const size_t ws_bytes = sort_workspace_bytes(h->cap);
    MSIM_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->stripes), query_stripe_bytes() + 128));
    MSIM_CUDA(h, cudaMemsetAsync(h->stripes, 0, query_stripe_bytes() + 128, h->stream));
    MSIM_CUDA(h, cudaMalloc(&h->sort_mem, ws_bytes));
    sort_workspace_bind(h->ws, h->sort_mem, h->cap);
    h->ws.error_flag = &h->counters->error_flag;
    return ensure_cells(h);
}

int upload(msim_handle* h, const msim_entity* src, uint64_t count) {
    if (count > h->cap) return fail(h, MSIM_ERR_CAPACITY, "msim_upload_entities: count exceeds entity_capacity");
    if (count && !src) return fail(h, MSIM_ERR_INVALID, "msim_upload_entities: null source");
    if (h->awaiting_integrate) h->arrive_deferred = false;  // pass B of a move whose exchange never completed: the population is being replaced
    join_side(h);
    MSIM_CUDA(h, cudaMemsetAsync(h->scratch, 0, 2 * sizeof(unsigned int), h->stream));
    for (uint64_t off = 0; off < count; off += STAGE_ENTITIES) {
        const uint32_t chunk = static_cast<uint32_t>(std::min<uint64_t>(STAGE_ENTITIES, count - off));
        MSIM_CUDA(h, cudaMemcpyAsync(h->stage, src + off, static_cast<size_t>(chunk) * sizeof(msim_entity), cudaMemcpyHostToDevice, h->stream));
        h->launches += launch_unpack(h->stream, static_cast<uint32_t>(off), chunk, h->stage, h->pos[0], h->target, h->road, h->rng, h->color0,
                                     h->dir0, nullptr, h->scratch, &h->prof);
    }
    h->launches += launch_max_road(h->stream, static_cast<uint32_t>(count), h->road, h->scratch + 1);
    unsigned int host_scratch[2] = {0, 0};
    MSIM_CUDA(h, cudaMemcpyAsync(host_scratch, h->scratch, sizeof(host_scratch), cudaMemcpyDeviceToHost, h->stream));
    MSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    MSIM_CUDA(h, cudaGetLastError());
    h->n = static_cast<uint32_t>(count);
    h->cur = 0;
    h->has_moved = false;
    h->keys_valid = false;
    h->hist_valid = false;
    h->counts_valid = false;
    h->collided = false;
    h->flags_scattered = false;
    h->slots_valid = false;
    h->n_ghost = 0;
    h->flags_stale = false;
    h->async_counts = false;
    h->awaiting_integrate = false;          // a pending exchange refers to the population that has just been replaced
    h->packed = false;
    h->band_valid = false;
    h->count_fused = false;
    {
        const int wrc = write_dev_counts(h);
        if (wrc != MSIM_OK) return wrc;
    }
    h->perm_active = false;                 // uploaded state is in external order again
    h->since_reorder = h->reorder_every;    // re-sort right after the first collision pass
    if (h->flag_entity) MSIM_CUDA(h, cudaMemsetAsync(h->flag_entity, 0, h->cap, h->stream));
    if (count && host_scratch[1] >= h->road_count) {
        h->n = 0;
        return fail(h, MSIM_ERR_INVALID, "entity road_index out of range (max " + std::to_string(host_scratch[1]) + ", roads " + std::to_string(h->road_count) + ")");
    }
    if (host_scratch[0] != 0 && host_scratch[0] != count) {
        h->n = 0;
        return fail(h, MSIM_ERR_UNSUPPORTED, "mixed initialized flags: the reference uploads every entity with initialized = 0 (Simulator.cpp:121-127); all-0 or all-1 is supported");
    }
    h->uninitialised = host_scratch[0] != 0;
    return MSIM_OK;
}

int check_device_errors(msim_handle* h) {
    join_side(h);
    {
        const int rc = refresh_counts(h);
        if (rc != MSIM_OK) return rc;
    }
    Counters c{};
    MSIM_CUDA(h, cudaMemcpyAsync(&c, h->counters, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    MSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    MSIM_CUDA(h, cudaGetLastError());
    h->last_pairs = h->n ? c.pairs_last : 0;
    h->total_pairs = c.pairs_total;
    h->last_flagged = h->n ? c.flagged_last : 0;
    h->total_flagged = c.flagged_total;
    if (c.error_flag) return fail(h, MSIM_ERR_INTERNAL, "look-back watchdog tripped (radix sort / cell scan): a tile never published its total");
    return MSIM_OK;
}

// first dispatch after an upload of uninitialised entities: the shader's init branch only
// (random_move.comp:863-867) — every entity is registered with the neighbour structure, nothing moves.
bool consume_init_dispatch(msim_handle* h) {
    if (!h->uninitialised) return false;
    h->uninitialised = false;
    h->initialised_total += h->n;
    return true;
}

// shard != NULL: called by move_pack_common, which has begun the move phase itself (`ms`)
int enqueue_move(msim_handle* h, bool want_keys, const ShardMoveArgs* shard = nullptr, cudaStream_t ms = nullptr) {
    if (h->awaiting_integrate) return fail(h, MSIM_ERR_INVALID, "move pass on a sharded handle whose last msim_shard_move_pack has not been integrated");
    if (consume_init_dispatch(h)) return MSIM_OK;
    const bool emit = want_keys && !(h->flags & MSIM_FLAG_NO_COLLISIONS);
    if (emit) {
        const int rc = alloc_collision_buffers(h);
        if (rc != MSIM_OK) return rc;
    }
    const int passes = (h->key_bits + RADIX_BITS - 1) / RADIX_BITS;
    // what the neighbour rebuild can take over from this pass: the counting sort's per-cell population, or the
    // onesweep digit histograms.  Sharded handles change their key set in the exchange that follows.
    const bool fuse_count = emit && h->use_csort && (!h->sharded || shard);
    const bool fuse_hist = emit && !h->use_csort && !h->sharded;
    const bool count_only = fuse_count && !h->sharded;  // no keys written: the scatter recomputes the key and takes the slot
    h->n_ghost = 0;
    h->count_fused = fuse_count && h->sharded;
    // a collision pass follows and needs only positions: the pass runs beside the previous tick's query (unsharded handles; sharded
    // ones decide in move_pack_common), and its own pass B will be launched beside the next query
    const bool defer_arrive = emit && h->side && (!h->sharded || shard);
    const bool beside = ms ? ms == h->side : (defer_arrive && !h->sharded && tuning().overlap_ticks);
    if (!ms) ms = begin_move_phase(h, beside);
    if (fuse_count) prepare_counts(h, ms);
    if (fuse_hist) sort_prepare(ms, h->n, h->key_bits, h->ws, &h->prof);
    h->launches += launch_move(ms, h->sm_count, launch_owned(h), h->pos[h->cur], h->pos[h->cur ^ 1], h->target, h->arrived,
                               emit && !count_only ? h->keys : nullptr, h->grid, fuse_hist ? h->ws.hist : nullptr,
                               passes > MAX_SORT_PASSES ? MAX_SORT_PASSES : passes, fuse_count ? h->cell_count : nullptr, &h->prof, dev_owned(h), shard,
                               beside ? (h->sharded ? tuning().shard_move_beside_ctas_per_sm : tuning().move_beside_ctas_per_sm) : 0);
    h->counts_valid = fuse_count;
    if (defer_arrive) {
        h->arrive_deferred = true;
    } else {
        h->launches += launch_arrive(ms, launch_owned(h), h->target, h->road, h->rng, h->arrived, h->roads, h->conn, h->conn_count, &h->prof,
                                     dev_owned(h));
    }
    if (!shard) end_move_phase(h);
    h->cur ^= 1;
    h->has_moved = true;
    h->band_valid = false;  // entities may have crossed the band's rows until the next pack + integrate
    h->keys_valid = emit && !count_only;
    h->hist_valid = fuse_hist;
    h->move_passes++;
    return MSIM_OK;
}

// Permute the resident state into the cell order of the collision pass that just ran.
int reorder_storage(msim_handle* h) {
    if (!h->pos_spare) {
        MSIM_CUDA(h, dev_alloc(&h->pos_spare, h->cap));
        MSIM_CUDA(h, dev_alloc(&h->target_alt, h->cap));
        MSIM_CUDA(h, dev_alloc(&h->road_alt, h->cap));
        MSIM_CUDA(h, dev_alloc(&h->rng_alt, h->cap));
        MSIM_CUDA(h, dev_alloc(&h->arrived_alt, h->cap / 32 + 2));
        MSIM_CUDA(h, dev_alloc(&h->ext_id, h->cap));
        MSIM_CUDA(h, dev_alloc(&h->ext_id_alt, h->cap));
        MSIM_CUDA(h, dev_alloc(&h->slot_of, h->cap));
        MSIM_CUDA(h, cudaMemsetAsync(h->pos_spare, 0, sizeof(float2) * h->cap, h->stream));
        MSIM_CUDA(h, cudaMemsetAsync(h->target_alt, 0, sizeof(float2) * h->cap, h->stream));
    }
    join_side(h);  // pass B of the last move rewrites target / road / rng
    if (h->sharded) {
        // gid is the id map; the owned entities are the run of the sorted order that starts at the band's first row
        if (!h->gid_alt) MSIM_CUDA(h, dev_alloc(&h->gid_alt, h->cap));
        ReorderArrays a{};
        a.pos_prev = h->pos[h->cur ^ 1];
        a.pos_prev_new = h->pos_spare;
        a.target = h->target;      a.target_new = h->target_alt;
        a.road = h->road;          a.road_new = h->road_alt;
        a.rng = h->rng;            a.rng_new = h->rng_alt;
        a.ext_id = h->gid;         a.ext_id_new = h->gid_alt;
        a.slot_of = nullptr;
        a.arrived = h->arrived;    a.arrived_new = h->arrived_alt;
        a.flag_entity = h->flag_entity;
        a.first_owned = h->cell_start - 1 + static_cast<size_t>(h->band_lo) * static_cast<size_t>(h->grid.ncx);  // the shifted table: run starts
        a.n_owned_dev = cur_counts(h) + DEV_N_OWNED;
        a.sorted_pos = h->sorted_pos;
        a.pos_cur_new = h->pos[h->cur];  // not read by the kernel: sorted_pos holds the same positions in the new order
        a.error_word = dev_error(h);
        h->launches += launch_invert_slots(h->stream, launch_total(h), h->rank, h->sorted_idx, &h->prof, dev_total(h));  // slot -> entity on demand
        h->launches += launch_reorder_sharded(h->stream, launch_owned(h), h->cap / 32 + 2, h->sorted_idx, h->flag_sorted, a, &h->prof);
        h->slots_valid = false;
        float2* old_prev = h->pos[h->cur ^ 1];
        h->pos[h->cur ^ 1] = h->pos_spare;
        h->pos_spare = old_prev;
        std::swap(h->target, h->target_alt);
        std::swap(h->road, h->road_alt);
        std::swap(h->rng, h->rng_alt);
        std::swap(h->arrived, h->arrived_alt);
        std::swap(h->gid, h->gid_alt);
        h->flags_scattered = true;
        h->keys_valid = false;
        h->hist_valid = false;
        h->counts_valid = false;
        h->since_reorder = 0;
        h->reorders++;
        h->main_touched = true;
        return MSIM_OK;
    }
    ReorderArrays a{};
    a.pos_prev = h->pos[h->cur ^ 1];
    a.pos_prev_new = h->pos_spare;
    a.target = h->target;      a.target_new = h->target_alt;
    a.road = h->road;          a.road_new = h->road_alt;
    a.rng = h->rng;            a.rng_new = h->rng_alt;
    a.ext_id = h->perm_active ? h->ext_id : nullptr;
    a.ext_id_new = h->ext_id_alt;
    a.slot_of = h->slot_of;
    a.arrived = h->arrived;    a.arrived_new = h->arrived_alt;
    a.flag_entity = h->flag_entity;
    if (h->slots_valid) h->launches += launch_invert_slots(h->stream, h->n, h->rank, h->sorted_idx, &h->prof);  // counting sort: slot -> entity on demand
    h->launches += launch_reorder(h->stream, h->n, h->sorted_idx, h->flag_sorted, a, &h->prof);
    h->slots_valid = false;
    // the sorted positions ARE the new current positions: swap buffers instead of copying
    float2* old_cur = h->pos[h->cur];
    float2* old_prev = h->pos[h->cur ^ 1];
    h->pos[h->cur] = h->sorted_pos;
    h->sorted_pos = old_cur;
    h->pos[h->cur ^ 1] = h->pos_spare;
    h->pos_spare = old_prev;
    std::swap(h->target, h->target_alt);
    std::swap(h->road, h->road_alt);
    std::swap(h->rng, h->rng_alt);
    std::swap(h->arrived, h->arrived_alt);
    std::swap(h->ext_id, h->ext_id_alt);
    h->perm_active = true;
    h->flags_scattered = true;   // flag_entity was written in slot order by the re-sort
    h->keys_valid = false;       // keys / ranks were indexed by the old slots
    h->hist_valid = false;
    h->counts_valid = false;
    h->since_reorder = 0;
    h->reorders++;
    h->main_touched = true;
    return MSIM_OK;
}

int finish_collide(msim_handle* h, uint32_t total) {
    h->collide_total = total;
    h->collide_owned = launch_owned(h);
    h->flags_stale = false;
    h->collided = true;
    h->flags_scattered = false;
    h->collide_passes++;
    h->since_reorder++;
    // a sharded handle finds its owned run through the counting sort's prefix table and only while every owned entity lies in the band
    if (h->reorder_enabled && h->has_moved && (h->n > 1 || h->async_counts) && h->since_reorder >= h->reorder_every &&
        (!h->sharded || (h->use_csort && h->band_valid)))
        return reorder_storage(h);
    return MSIM_OK;
}

// Unsharded handles, counting sort (the single-GPU default).  The main stream carries nothing but the query and its fold; everything else
// of a tick - pass B, the move pass, the scan of the per-cell counters, the scatter into cell order - lives on the side stream, so the
// rebuild of tick t+1 runs beside the query of tick t instead of behind it (the query is issue-bound, the rest is memory-bound).  The
// rebuild writes the other of two {sorted positions, prefix table} sets; a set is rewritten only when the query that read it is done
// (ev_queried), and a query starts when its set is built (ev_built).  The pointers h->sorted_pos / h->cell_start name the set of the
// collision pass enqueued last, which is what readbacks and the periodic re-sort work on (they join the streams first).
int enqueue_collide_pipelined(msim_handle* h, uint32_t total, bool count_pairs) {
    if (!h->sorted_pos_alt) {
        MSIM_CUDA(h, dev_alloc(&h->sorted_pos_alt, h->cap));
        MSIM_CUDA(h, cudaEventCreateWithFlags(&h->ev_built, cudaEventDisableTiming));
        MSIM_CUDA(h, cudaEventCreateWithFlags(&h->ev_queried[0], cudaEventDisableTiming));
        MSIM_CUDA(h, cudaEventCreateWithFlags(&h->ev_queried[1], cudaEventDisableTiming));
    }
    if (!h->cell_table_alt) {  // (freed with the first table when the grid outgrows it)
        MSIM_CUDA(h, dev_alloc(&h->cell_table_alt, static_cast<size_t>(h->cell_capacity) + 8));
        MSIM_CUDA(h, cudaMemsetAsync(h->cell_table_alt, 0, 4 * sizeof(uint32_t), h->stream));
        h->main_touched = true;
    }
    const cudaStream_t side = h->side;
    if (h->main_touched) {  // an upload, a readback, a re-sort ... since the side stream last waited for the main one
        MSIM_CUDA(h, cudaEventRecord(h->ev_moved, h->stream));
        MSIM_CUDA(h, cudaStreamWaitEvent(side, h->ev_moved, 0));
        h->main_touched = false;
    }
    // the other set: free once the query that read it (two collision passes ago) is done
    h->build_set ^= 1u;
    std::swap(h->sorted_pos, h->sorted_pos_alt);
    std::swap(h->cell_table, h->cell_table_alt);
    h->cell_start = h->cell_table + 4;
    if (h->queried_recorded[h->build_set]) MSIM_CUDA(h, cudaStreamWaitEvent(side, h->ev_queried[h->build_set], 0));
    if (!h->counts_valid) {  // no counting move pass in front of this dispatch: count now
        prepare_counts(h, side);
        h->launches += launch_cell_count_pos(side, h->sm_count, h->n, h->pos[h->cur], h->cell_count, h->grid, &h->prof);
    }
    h->counts_valid = false;
    h->launches += scan_cells(h, 0, h->grid.ncells, side);
    h->counts_dirty = false;  // the scan zeroed every counter it read
    h->launches += launch_cell_scatter_slots(side, h->sm_count, total, h->pos[h->cur], nullptr, h->cell_start, h->sorted_pos, h->rank, h->grid, 0, h->grid.ncells,
                                             &h->prof, nullptr, tuning().scatter_beside_ctas_per_sm);
    MSIM_CUDA(h, cudaEventRecord(h->ev_built, side));
    if (h->arrive_deferred && !h->awaiting_integrate) {  // pass B of the move in front of this pass: behind the scatter, beside the query
        h->arrive_deferred = false;
        h->launches += launch_arrive(side, launch_owned(h), h->target, h->road, h->rng, h->arrived, h->roads, h->conn, h->conn_count, &h->prof, nullptr,
                                     /*beside=*/true, -1);
    }
    MSIM_CUDA(h, cudaEventRecord(h->ev_arrived, side));
    h->side_pending = true;
    MSIM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_built, 0));
    h->launches += launch_query_tiles(h->stream, total, h->sorted_pos, h->cell_start - 1, h->flag_sorted, h->grid, count_pairs, h->counters, h->stripes, &h->prof,
                                      nullptr, false, 0, 0);
    MSIM_CUDA(h, cudaEventRecord(h->ev_queried[h->build_set], h->stream));
    h->queried_recorded[h->build_set] = true;
    h->slots_valid = true;
    return finish_collide(h, total);
}

int enqueue_collide(msim_handle* h) {
    if (h->flags & MSIM_FLAG_NO_COLLISIONS) return fail(h, MSIM_ERR_INVALID, "collision dispatch on a handle created with MSIM_FLAG_NO_COLLISIONS");
    if (consume_init_dispatch(h)) return MSIM_OK;
    int rc = alloc_collision_buffers(h);
    if (rc != MSIM_OK) return rc;
    join_side_work(h);  // an overlapped move phase (and the pass B in front of it) must be complete before the rebuild reads its results
    if (!h->keys_valid && (h->sharded || !h->use_csort)) {  // (the unsharded counting sort works on positions: no keys)
        rc = refresh_counts(h);  // (asynchronous sharded ticks) keygen is sized by the exact owned count
        if (rc != MSIM_OK) return rc;
        h->launches += launch_keygen(h->stream, h->n, h->pos[h->cur], h->keys, h->grid, &h->prof);
        h->keys_valid = true;
        h->hist_valid = false;
    }
    const uint32_t total = launch_total(h);  // ghosts (multi-GPU halo) sit behind the owned entities
    const bool count_pairs = !(h->flags & MSIM_FLAG_NO_PAIR_COUNT);
    if (h->use_csort && !h->sharded && h->side && tuning().overlap_ticks && tuning().pipeline_build) return enqueue_collide_pipelined(h, total, count_pairs);
    if (h->use_csort) {
        // sharded handles: the counter table is cleared / scanned over the band's cell range only, keys outside it
        // (a leaver that jumped two rows while the boundary moved: out of everybody's reach) are left out of the
        // order, and the number of sorted slots is the scan's grand total, read by the query from the table itself
        uint32_t c0 = 0, c1 = h->grid.ncells;
        const bool band = h->sharded && h->band_valid;
        if (band) csort_band(h->grid.ncells, h->grid.ncx, h->band_lo, h->band_hi, h->grid.ncy, &c0, &c1);
        if (!h->counts_valid) {  // no counting move pass in front of this dispatch, or the keys changed in a shard exchange: count now
            prepare_counts(h);
            if (h->sharded) h->launches += launch_cell_count(h->stream, total, h->keys, h->cell_count, c0, c1, &h->prof, dev_total(h));
            else h->launches += launch_cell_count_pos(h->stream, h->sm_count, h->n, h->pos[h->cur], h->cell_count, h->grid, &h->prof);
        }
        h->counts_valid = false;
        h->launches += scan_cells(h, c0, c1);
        h->counts_dirty = false;  // the scan zeroed every counter it read, and nothing was counted outside [c0, c1)
        // the scatter's atomics turn cell_start into the table of run ENDS; read one word earlier it is the table of run starts
        h->launches += launch_cell_scatter_slots(h->stream, h->sm_count, total, h->pos[h->cur], h->sharded ? h->keys : nullptr, h->cell_start, h->sorted_pos,
                                                 h->rank, h->grid, c0, c1, &h->prof, dev_total(h));
        launch_deferred_arrive(h, true);
        const uint32_t* tab = h->cell_start - 1;
        h->launches += launch_query_tiles(h->stream, total, h->sorted_pos, tab, h->flag_sorted, h->grid, count_pairs, h->counters, h->stripes, &h->prof,
                                          h->sharded ? tab + c1 : nullptr, band, h->band_lo, h->band_hi);
        h->slots_valid = true;
    } else {
        h->launches += launch_sort(h->stream, total, h->keys, h->sort_a, h->sort_b, h->key_bits, h->ws, &h->sorted,
                                   h->hist_valid && h->n_ghost == 0 && !h->async_counts, &h->prof, dev_total(h));
        h->hist_valid = false;  // the sort consumed the tickets and look-back words
        h->launches += launch_build_cells(h->stream, total, h->sorted, h->pos[h->cur], h->sorted_pos, h->sorted_idx, h->cell_range, h->grid, h->counters,
                                          &h->prof, dev_total(h));
        launch_deferred_arrive(h, true);
        h->launches += launch_query(h->stream, total, launch_owned(h), h->sorted_idx, h->sorted_pos, h->cell_range, h->flag_sorted, h->grid, count_pairs,
                                    h->counters, h->stripes, &h->prof, dev_total(h), dev_owned(h));
        h->slots_valid = false;
    }
    return finish_collide(h, total);
}

int bind(msim_handle* h) {
    if (!h) return MSIM_ERR_INVALID;
    MSIM_CUDA(h, cudaSetDevice(h->device));
    return MSIM_OK;
}

int materialise_flags(msim_handle* h) {
    if (h->collided && !h->flags_scattered) {
        if (h->async_counts) {  // the collision pass ran on device-resident counts: fetch them (they have not changed since)
            if (refresh_counts(h) != MSIM_OK) return MSIM_ERR_CAPACITY;
            h->collide_owned = h->n;
            h->collide_total = h->n + h->n_ghost;
        }
        if (h->slots_valid)
            h->launches += launch_gather_flags(h->stream, h->collide_owned, h->rank, h->flag_sorted, h->flag_entity, &h->prof);
        else
            h->launches += launch_scatter_flags(h->stream, h->collide_total, h->collide_owned, h->sorted_idx, h->flag_sorted, h->flag_entity, &h->prof);
        h->flags_scattered = true;
    }
    return MSIM_OK;
}

// arguments of the SoA -> AoS pack for a readback of the current state; completes a pending pass B and the lazy flag scatter first
PackArgs pack_args(msim_handle* h) {
    materialise_flags(h);
    join_side(h);
    PackArgs a{};
    a.pos_cur = h->pos[h->cur];
    a.pos_prev = h->pos[h->cur ^ 1];
    a.target = h->target;
    a.road = h->road;
    a.rng = h->rng;
    a.color0 = h->color0;
    a.dir0 = h->dir0;
    a.arrived = h->arrived;
    a.flag_entity = h->collided ? h->flag_entity : nullptr;
    a.init_mask = nullptr;
    a.slot_of = h->perm_active ? h->slot_of : nullptr;
    a.initialized_all = h->uninitialised ? 0u : 1u;
    a.has_moved = h->has_moved ? 1u : 0u;
    return a;
}

}  // namespace

extern "C" {

const char* msim_status_string(int status) {
    switch (status) {
        case MSIM_OK: return "ok";
        case MSIM_ERR_INVALID: return "invalid argument";
        case MSIM_ERR_CUDA: return "CUDA error";
        case MSIM_ERR_OOM: return "out of memory";
        case MSIM_ERR_UNSUPPORTED: return "unsupported";
        case MSIM_ERR_IO: return "I/O error";
        case MSIM_ERR_PARSE: return "parse error";
        case MSIM_ERR_CAPACITY: return "capacity exceeded";
        case MSIM_ERR_INTERNAL: return "internal error";
        default: return "unknown status";
    }
}

const char* msim_last_error(const msim_handle* h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int msim_create(const msim_config* cfg, msim_handle** out) {
    if (!out) return fail(nullptr, MSIM_ERR_INVALID, "msim_create: out is null");
    *out = nullptr;
    if (!cfg) return fail(nullptr, MSIM_ERR_INVALID, "msim_create: cfg is null");
    if (cfg->abi_version != MSIM_ABI_VERSION) return fail(nullptr, MSIM_ERR_INVALID, "msim_create: abi_version mismatch");
    if (!cfg->roads || cfg->road_count == 0) return fail(nullptr, MSIM_ERR_INVALID, "msim_create: the map has no roads");
    if (cfg->road_count > 0xffffffffull) return fail(nullptr, MSIM_ERR_INVALID, "msim_create: more than 2^32 roads");
    if (cfg->connection_count && !cfg->connections) return fail(nullptr, MSIM_ERR_INVALID, "msim_create: connections is null");
    if (cfg->entity_count && !cfg->entities) return fail(nullptr, MSIM_ERR_INVALID, "msim_create: entities is null");
    const uint64_t capacity = cfg->entity_capacity ? cfg->entity_capacity : cfg->entity_count;
    if (capacity < cfg->entity_count) return fail(nullptr, MSIM_ERR_INVALID, "msim_create: entity_capacity < entity_count");
    if (capacity > MAX_ENTITIES_PER_HANDLE) return fail(nullptr, MSIM_ERR_UNSUPPORTED, "msim_create: more than 2^30 entities per handle");
    if (!(cfg->world_w >= 0.0f) || !(cfg->world_h >= 0.0f)) return fail(nullptr, MSIM_ERR_INVALID, "msim_create: bad world size");
    for (uint64_t i = 0; i < cfg->connection_count; i++) {
        if (cfg->connections[i] >= cfg->road_count)
            return fail(nullptr, MSIM_ERR_INVALID, "msim_create: connections[" + std::to_string(i) + "] = " + std::to_string(cfg->connections[i]) + " is not a road index");
    }

    int device_count = 0;
    cudaError_t err = cudaGetDeviceCount(&device_count);
    if (err != cudaSuccess || device_count == 0)
        return fail(nullptr, MSIM_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(err) + " (libmsim_cuda has no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= device_count) return fail(nullptr, MSIM_ERR_INVALID, "msim_create: device ordinal out of range");
    cudaDeviceProp prop{};
    err = cudaGetDeviceProperties(&prop, cfg->device);
    if (err != cudaSuccess) return fail(nullptr, MSIM_ERR_CUDA, cudaGetErrorString(err));
    if (prop.major != 10) return fail(nullptr, MSIM_ERR_CUDA, std::string("device '") + prop.name + "' is not sm_100-class; this library ships sm_100a code only");

    msim_handle* h = new (std::nothrow) msim_handle();
    if (!h) return fail(nullptr, MSIM_ERR_OOM, "out of host memory");
    h->device = cfg->device;
    h->sm_count = prop.multiProcessorCount;
    h->flags = cfg->flags;
    h->world_w = cfg->world_w;
    h->world_h = cfg->world_h;
    h->radius = cfg->collision_radius;
    h->qt_depth = cfg->quadtree_max_depth ? cfg->quadtree_max_depth : 8;
    h->qt_cap = cfg->quadtree_node_cap ? cfg->quadtree_node_cap : 10;
    h->road_count = cfg->road_count;
    h->conn_count = cfg->connection_count;
    h->reorder_enabled = !(cfg->flags & (MSIM_FLAG_NO_COLLISIONS | MSIM_FLAG_NO_REORDER));
    if (const char* env = std::getenv("MSIM_REORDER_EVERY")) {
        const long v = std::atol(env);
        if (v > 0) h->reorder_every = static_cast<uint32_t>(v);
        else h->reorder_enabled = false;
    }
    h->since_reorder = h->reorder_every;
    h->cap = static_cast<uint32_t>((capacity + 63ull) & ~63ull);
    if (h->cap == 0) h->cap = 64;

    auto body = [&]() -> int {
        MSIM_CUDA(h, cudaSetDevice(h->device));
        if (cfg->cuda_stream) {
            h->stream = static_cast<cudaStream_t>(cfg->cuda_stream);
        } else {
            MSIM_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
            h->own_stream = true;
        }
        {
            // the side stream carries pass B, the move pass and (unsharded handles) the rebuild beside the issue-bound, ~80 k CTA collision
            // query: highest priority, so that its CTAs are placed as soon as query CTAs retire instead of queueing behind all of them
            // (measured with the pipelined rebuild: 288 us per tick, 336 us with the side stream at the LOWEST priority)
            int prio_low = 0, prio_high = 0;
            MSIM_CUDA(h, cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high));
            MSIM_CUDA(h, cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_high));
        }
        MSIM_CUDA(h, cudaEventCreateWithFlags(&h->ev_moved, cudaEventDisableTiming));
        MSIM_CUDA(h, cudaEventCreateWithFlags(&h->ev_arrived, cudaEventDisableTiming));
        MSIM_CUDA(h, dev_alloc(&h->pos[0], h->cap));
        MSIM_CUDA(h, dev_alloc(&h->pos[1], h->cap));
        MSIM_CUDA(h, dev_alloc(&h->target, h->cap));
        MSIM_CUDA(h, dev_alloc(&h->road, h->cap));
        MSIM_CUDA(h, dev_alloc(&h->rng, h->cap));
        MSIM_CUDA(h, dev_alloc(&h->color0, h->cap));
        MSIM_CUDA(h, dev_alloc(&h->dir0, h->cap));
        MSIM_CUDA(h, dev_alloc(&h->arrived, h->cap / 32 + 2));
        MSIM_CUDA(h, dev_alloc(&h->roads, h->road_count));
        MSIM_CUDA(h, dev_alloc(&h->conn, h->conn_count));
        MSIM_CUDA(h, dev_alloc(&h->counters, 1));
        MSIM_CUDA(h, dev_alloc(&h->scratch, 4));
        MSIM_CUDA(h, dev_alloc(&h->stage, std::min<uint64_t>(STAGE_ENTITIES, h->cap)));
        MSIM_CUDA(h, cudaMemsetAsync(h->counters, 0, sizeof(Counters), h->stream));
        MSIM_CUDA(h, cudaMemsetAsync(h->pos[0], 0, sizeof(float2) * h->cap, h->stream));
        MSIM_CUDA(h, cudaMemsetAsync(h->pos[1], 0, sizeof(float2) * h->cap, h->stream));
        MSIM_CUDA(h, cudaMemsetAsync(h->target, 0, sizeof(float2) * h->cap, h->stream));
        MSIM_CUDA(h, cudaMemsetAsync(h->arrived, 0, sizeof(uint32_t) * (h->cap / 32 + 2), h->stream));
        MSIM_CUDA(h, cudaMemcpyAsync(h->roads, cfg->roads, sizeof(msim_road) * h->road_count, cudaMemcpyHostToDevice, h->stream));
        if (h->conn_count)
            MSIM_CUDA(h, cudaMemcpyAsync(h->conn, cfg->connections, sizeof(uint32_t) * h->conn_count, cudaMemcpyHostToDevice, h->stream));
        configure_grid(h);
        apply_l2_window(h);
        return upload(h, cfg->entities, cfg->entity_count);
    };
    const int rc = body();
    if (rc != MSIM_OK) {
        g_create_error = h->error;
        free_all(h);
        delete h;
        return rc;
    }
    *out = h;
    return MSIM_OK;
}

void msim_destroy(msim_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->side) cudaStreamSynchronize(h->side);
    if (h->push_stream) cudaStreamSynchronize(h->push_stream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->shard_trace) {  // MSIM_SHARD_TRACE=1
        unsigned long long t[16] = {};
        if (cudaMemcpyAsync(t, h->shard_trace, sizeof(t), cudaMemcpyDeviceToHost, h->stream) == cudaSuccess && cudaStreamSynchronize(h->stream) == cudaSuccess && t[4]) {
            const double k = 1e-3 / static_cast<double>(t[4]);
            std::fprintf(stderr, "msim shard trace (device %d, %llu exchanges, CTA 0): gap behind the move kernel's stamp %.1f us, flag wait %.1f us, integrate %.1f us, ghosts %.1f us; start to last CTA done %.1f us, from there to a stamp kernel behind it %.1f us\n",
                         h->device, t[4], t[0] * k, t[1] * k, t[2] * k, t[3] * k, t[10] * k, t[11] * k);
        }
        cudaFree(h->shard_trace);
    }
    free_all(h);
    delete h;
}

int msim_upload_entities(msim_handle* h, const msim_entity* src, uint64_t count) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    return upload(h, src, count);
}

int msim_set_stream(msim_handle* h, void* cuda_stream) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    join_side(h);
    MSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->own_stream) {
        cudaStreamDestroy(h->stream);
        h->own_stream = false;
    }
    if (cuda_stream) {
        h->stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
        MSIM_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    apply_l2_window(h);
    return MSIM_OK;
}

int msim_enqueue_move(msim_handle* h) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    rc = enqueue_move(h, true);
    join_side_work(h);  // callers order their own work behind the handle's stream: the pass may have run beside it
    return rc;
}

int msim_enqueue_collide(msim_handle* h) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    return enqueue_collide(h);
}

int msim_enqueue_ticks(msim_handle* h, uint32_t sim_ticks, int with_collisions) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    for (uint32_t t = 0; t < sim_ticks; t++) {
        rc = enqueue_move(h, with_collisions != 0);
        if (rc != MSIM_OK) return rc;
        if (with_collisions) {
            rc = enqueue_collide(h);
            if (rc != MSIM_OK) return rc;
        }
    }
    return MSIM_OK;
}

int msim_sync(msim_handle* h) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    return check_device_errors(h);
}

int msim_dispatch(msim_handle* h, const msim_push_consts* pc) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!pc) return fail(h, MSIM_ERR_INVALID, "msim_dispatch: push constants are null");
    if (pc->world_size_x != h->world_w || pc->world_size_y != h->world_h || pc->collision_radius != h->radius) {
        // push constants are per-dispatch state in the reference: follow them.  Work enqueued asynchronously under the old grid (a move
        // phase or a rebuild still on the side stream) is completed first
        join_side(h);
        MSIM_CUDA(h, cudaStreamSynchronize(h->stream));
        h->world_w = pc->world_size_x;
        h->world_h = pc->world_size_y;
        h->radius = pc->collision_radius;
        configure_grid(h);
        h->keys_valid = false;
        h->hist_valid = false;
        h->counts_valid = false;
        h->counts_dirty = true;  // a count under the old grid may have left counters anywhere in the allocation
        if (h->keys) {
            rc = ensure_cells(h);
            if (rc != MSIM_OK) return rc;
        }
    }
    rc = (pc->tick % 2u == 0u) ? enqueue_move(h, true) : enqueue_collide(h);
    if (rc != MSIM_OK) return rc;
    return check_device_errors(h);
}

namespace {
// what every readback checks first: the host's entity count is current (asynchronous sharded ticks keep it on the device), the
// request fits it, and a sharded handle is not in the middle of an exchange
int readback_preconditions(msim_handle* h, const char* who, const void* dst, uint64_t count, bool needs_flags) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    join_side(h);  // a move phase that ran beside the last query, its pass B
    rc = refresh_counts(h);
    if (rc != MSIM_OK) return rc;
    if (count > h->n) return fail(h, MSIM_ERR_INVALID, std::string(who) + ": count exceeds the resident entity count");
    if (count && !dst) return fail(h, MSIM_ERR_INVALID, std::string(who) + ": dst is null");
    if (h->awaiting_integrate) return fail(h, MSIM_ERR_INVALID, std::string(who) + ": sharded handle is between msim_shard_move_pack and msim_shard_integrate");
    if (needs_flags && h->flags_stale) return fail(h, MSIM_ERR_INVALID, std::string(who) + ": sharded handle is between msim_shard_integrate and the collision pass");
    return MSIM_OK;
}
}  // namespace

int msim_read_entities(msim_handle* h, msim_entity* dst, uint64_t count) {
    int rc = readback_preconditions(h, "msim_read_entities", dst, count, true);
    if (rc != MSIM_OK) return rc;
    const PackArgs a = pack_args(h);
    for (uint64_t off = 0; off < count; off += STAGE_ENTITIES) {
        const uint32_t chunk = static_cast<uint32_t>(std::min<uint64_t>(STAGE_ENTITIES, count - off));
        h->launches += launch_pack(h->stream, static_cast<uint32_t>(off), chunk, a, h->stage, &h->prof);
        MSIM_CUDA(h, cudaMemcpyAsync(dst + off, h->stage, static_cast<size_t>(chunk) * sizeof(msim_entity), cudaMemcpyDeviceToHost, h->stream));
    }
    return check_device_errors(h);
}

/* ---- asynchronous readback: snapshot now, copy while later ticks run (include/msim.h) -------------------------- */
int msim_snapshot_begin(msim_handle* h) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    rc = refresh_counts(h);
    if (rc != MSIM_OK) return rc;
    if (h->awaiting_integrate) return fail(h, MSIM_ERR_INVALID, "msim_snapshot_begin: sharded handle is between msim_shard_move_pack and msim_shard_integrate");
    if (h->flags_stale) return fail(h, MSIM_ERR_INVALID, "msim_snapshot_begin: sharded handle is between msim_shard_integrate and the collision pass");
    if (h->snap_cap < h->n || !h->snap_dev) {  // (re)allocate for the resident population: device image + two pinned host buffers
        if (h->snap_pending) MSIM_CUDA(h, cudaEventSynchronize(h->ev_copied));
        h->snap_pending = false;
        cudaFree(h->snap_dev);
        h->snap_dev = nullptr;
        for (msim_entity*& p : h->snap_host) {
            if (p) cudaFreeHost(p);
            p = nullptr;
        }
        h->snap_cap = 0;
        const size_t cap = std::max<size_t>(h->n, 1);
        MSIM_CUDA(h, dev_alloc(&h->snap_dev, cap));
        for (msim_entity*& p : h->snap_host) MSIM_CUDA(h, cudaHostAlloc(reinterpret_cast<void**>(&p), cap * sizeof(msim_entity), cudaHostAllocDefault));
        h->snap_cap = cap;
        if (!h->copy_stream) MSIM_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        if (!h->ev_packed) MSIM_CUDA(h, cudaEventCreateWithFlags(&h->ev_packed, cudaEventDisableTiming));
        if (!h->ev_copied) MSIM_CUDA(h, cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming));
    }
    // the device image is reused: the previous copy must have left it before it is packed again
    if (h->snap_pending) MSIM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_copied, 0));
    const PackArgs a = pack_args(h);
    h->launches += launch_pack(h->stream, 0u, h->n, a, h->snap_dev, &h->prof);
    MSIM_CUDA(h, cudaEventRecord(h->ev_packed, h->stream));
    // from here on the main stream is free for the next ticks; the copy engine drains the image on its own stream
    h->snap_slot ^= 1;
    MSIM_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->ev_packed, 0));
    if (h->n) MSIM_CUDA(h, cudaMemcpyAsync(h->snap_host[h->snap_slot], h->snap_dev, static_cast<size_t>(h->n) * sizeof(msim_entity), cudaMemcpyDeviceToHost, h->copy_stream));
    MSIM_CUDA(h, cudaEventRecord(h->ev_copied, h->copy_stream));
    h->snap_count = h->n;
    h->snap_pending = true;
    return MSIM_OK;
}

int msim_snapshot_poll(msim_handle* h, int* ready) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!ready) return fail(h, MSIM_ERR_INVALID, "msim_snapshot_poll: ready is null");
    if (!h->snap_pending) return fail(h, MSIM_ERR_INVALID, "msim_snapshot_poll: no snapshot has been started");
    const cudaError_t q = cudaEventQuery(h->ev_copied);
    if (q != cudaSuccess && q != cudaErrorNotReady) MSIM_CUDA(h, q);
    *ready = q == cudaSuccess ? 1 : 0;
    return MSIM_OK;
}

int msim_snapshot_end(msim_handle* h, const msim_entity** entities, uint64_t* count) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!entities || !count) return fail(h, MSIM_ERR_INVALID, "msim_snapshot_end: null argument");
    if (!h->snap_pending) return fail(h, MSIM_ERR_INVALID, "msim_snapshot_end: no snapshot has been started");
    MSIM_CUDA(h, cudaEventSynchronize(h->ev_copied));
    h->snap_pending = false;
    *entities = h->snap_host[h->snap_slot];
    *count = h->snap_count;
    return MSIM_OK;
}

int msim_read_positions(msim_handle* h, float* dst_xy, uint64_t count) {
    int rc = readback_preconditions(h, "msim_read_positions", dst_xy, count, false);
    if (rc != MSIM_OK) return rc;
    const float2* src = h->pos[h->cur];
    if (count && h->perm_active) {  // storage is in cell order: gather into external order first
        float2* tmp = reinterpret_cast<float2*>(h->sort_a);
        h->launches += launch_gather_pos(h->stream, static_cast<uint32_t>(count), h->slot_of, h->pos[h->cur], tmp);
        src = tmp;
    }
    if (count) MSIM_CUDA(h, cudaMemcpyAsync(dst_xy, src, count * sizeof(float2), cudaMemcpyDeviceToHost, h->stream));
    return check_device_errors(h);
}

int msim_read_collision_flags(msim_handle* h, uint8_t* dst, uint64_t count) {
    int rc = readback_preconditions(h, "msim_read_collision_flags", dst, count, true);
    if (rc != MSIM_OK) return rc;
    if (!h->collided) {
        if (count) std::memset(dst, 0, count);
        return MSIM_OK;
    }
    rc = materialise_flags(h);
    if (rc != MSIM_OK) return rc;
    if (count) {  // external order, 0 / 1: one gather on the device instead of a pass over the result on the host
        uint8_t* tmp = reinterpret_cast<uint8_t*>(h->sort_b);
        h->launches += launch_gather_flag(h->stream, static_cast<uint32_t>(count), h->perm_active ? h->slot_of : nullptr, h->flag_entity, tmp);
        MSIM_CUDA(h, cudaMemcpyAsync(dst, tmp, count, cudaMemcpyDeviceToHost, h->stream));
    }
    return check_device_errors(h);
}

namespace {
// Emits the subtree of the node covering finest cells [x0, x0+size) x [y0, y0+size); returns its index.
struct QuadBuilder {
    const std::vector<std::vector<uint32_t>>* sums;  // sums[l][iy * side_l + ix]: entities in the level-l cell (l = 0: finest)
    int levels;
    uint32_t node_cap;
    msim_quadtree_node* out;
    uint64_t cap;
    uint64_t used;
    bool overflow;

    uint32_t emit(int level_up, uint32_t ix, uint32_t iy, float off_x, float off_y, float w, float hgt, uint32_t parent) {
        if (used >= cap) {
            overflow = true;
            return 0;
        }
        const uint32_t self = static_cast<uint32_t>(used++);
        msim_quadtree_node& nd = out[self];
        std::memset(&nd, 0, sizeof(nd));
        nd.offset_x = off_x;
        nd.offset_y = off_y;
        nd.width = w;
        nd.height = hgt;
        nd.prev_node_index = parent;
        const uint32_t side = 1u << (levels - level_up);
        const uint32_t count = (*sums)[level_up][static_cast<size_t>(iy) * side + ix];
        if (count > node_cap && level_up > 0) {  // random_move.comp:354: split unless there is room or maxDepth is reached
            nd.content_type = 1;  // NODE
            const float hw = w * 0.5f, hh = hgt * 0.5f;  // quad_tree_split_up_node, :285-295
            nd.next_tl = emit(level_up - 1, 2 * ix, 2 * iy, off_x, off_y, hw, hh, self);
            nd.next_tr = emit(level_up - 1, 2 * ix + 1, 2 * iy, off_x + hw, off_y, hw, hh, self);
            nd.next_bl = emit(level_up - 1, 2 * ix, 2 * iy + 1, off_x, off_y + hh, hw, hh, self);
            nd.next_br = emit(level_up - 1, 2 * ix + 1, 2 * iy + 1, off_x + hw, off_y + hh, hw, hh, self);
        } else {
            nd.content_type = 2;  // ENTITY leaf
            nd.entity_count = count;
        }
        return self;
    }
};
}  // namespace

int msim_read_quadtree_nodes(msim_handle* h, msim_quadtree_node* dst, uint64_t cap, uint64_t* count) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!dst || cap == 0 || !count) return fail(h, MSIM_ERR_INVALID, "msim_read_quadtree_nodes: bad arguments");
    const int depth = static_cast<int>(std::min<uint32_t>(std::max<uint32_t>(h->qt_depth, 1u), 8u));  // 4^7 finest cells fit one CTA's shared memory; the reference uses 8
    const int levels = depth - 1;  // the root is depth 1 (quad_tree_insert(index, 0, 1), :864)
    const bool root_only = (h->flags & MSIM_FLAG_NO_QUADTREE) || h->uninitialised || h->n == 0 || levels == 0;
    std::vector<std::vector<uint32_t>> sums(levels + 1);
    if (!root_only) {
        const size_t bins = static_cast<size_t>(1) << (2 * levels);
        if (!h->leaf_hist) MSIM_CUDA(h, dev_alloc(&h->leaf_hist, static_cast<size_t>(1) << 16));
        join_side(h);
        h->launches += launch_leaf_histogram(h->stream, h->sm_count, h->n, h->pos[h->cur], h->world_w, h->world_h, levels, h->leaf_hist);
        sums[0].resize(bins);
        MSIM_CUDA(h, cudaMemcpyAsync(sums[0].data(), h->leaf_hist, bins * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        rc = check_device_errors(h);
        if (rc != MSIM_OK) return rc;
        for (int l = 1; l <= levels; l++) {  // bottom-up sums
            const uint32_t side = 1u << (levels - l);
            sums[l].resize(static_cast<size_t>(side) * side);
            for (uint32_t y = 0; y < side; y++)
                for (uint32_t x = 0; x < side; x++) {
                    const std::vector<uint32_t>& f = sums[l - 1];
                    const size_t fs = static_cast<size_t>(side) * 2;
                    sums[l][static_cast<size_t>(y) * side + x] = f[(2 * y) * fs + 2 * x] + f[(2 * y) * fs + 2 * x + 1] + f[(2 * y + 1) * fs + 2 * x] + f[(2 * y + 1) * fs + 2 * x + 1];
                }
        }
    } else {
        sums.assign(1, std::vector<uint32_t>(1, (h->uninitialised || (h->flags & MSIM_FLAG_NO_QUADTREE)) ? 0u : h->n));
    }
    QuadBuilder b{&sums, root_only ? 0 : levels, h->qt_cap, dst, cap, 0, false};
    b.emit(root_only ? 0 : levels, 0, 0, 0.0f, 0.0f, h->world_w, h->world_h, 0);
    if (b.overflow) return fail(h, MSIM_ERR_CAPACITY, "msim_read_quadtree_nodes: node buffer too small (calc_node_count(maxDepth) entries suffice)");
    *count = b.used;
    return MSIM_OK;
}

int msim_read_debug(msim_handle* h, uint32_t dst[10]) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!dst) return fail(h, MSIM_ERR_INVALID, "msim_read_debug: dst is null");
    rc = check_device_errors(h);
    std::memset(dst, 0, 10 * sizeof(uint32_t));
    dst[0] = static_cast<uint32_t>(h->initialised_total);
    dst[1] = static_cast<uint32_t>(h->total_pairs);
    return rc;
}

int msim_get_stats(msim_handle* h, msim_stats* out) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!out) return fail(h, MSIM_ERR_INVALID, "msim_get_stats: out is null");
    rc = check_device_errors(h);
    std::memset(out, 0, sizeof(*out));
    out->entity_count = h->n;
    out->move_passes = h->move_passes;
    out->collide_passes = h->collide_passes;
    out->last_pair_count = h->last_pairs;
    out->total_pair_count = h->total_pairs;
    out->last_flagged_count = h->last_flagged;
    out->kernel_launches = h->launches;
    out->grid_cells_x = static_cast<uint32_t>(h->grid.ncx);
    out->grid_cells_y = static_cast<uint32_t>(h->grid.ncy);
    out->key_bits = static_cast<uint32_t>(h->key_bits);
    // passes over the keys per rebuild: one for the counting sort (known once the collision buffers exist), one per 8-bit digit for onesweep
    out->sort_passes = (h->keys && h->use_csort) ? 1u : static_cast<uint32_t>((h->key_bits + RADIX_BITS - 1) / RADIX_BITS);
    out->cell_size = h->grid.inv_cell > 0.0f ? 1.0f / h->grid.inv_cell : 0.0f;
    out->reorders = static_cast<uint32_t>(h->reorders);
    out->total_flagged_count = h->total_flagged;
    return rc;
}

int msim_profile_begin(msim_handle* h) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    h->prof.enabled = true;
    h->prof.used = 0;
    h->prof.ids.clear();
    return MSIM_OK;
}

int msim_profile_end(msim_handle* h, msim_kernel_time* out, uint32_t cap, uint32_t* count) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!out || !count) return fail(h, MSIM_ERR_INVALID, "msim_profile_end: null argument");
    static const char* const names[K_COUNT] = {"move", "arrive", "keygen", "histogram", "sort_pass0", "sort_pass1", "sort_pass2", "sort_pass3",
                                               "build_cells", "query", "scatter_flags", "pack", "unpack", "memset", "misc", "shard", "cell_count",
                                               "cell_scan", "cell_scatter", "reorder", "fold_counts"};
    h->prof.enabled = false;
    join_side(h);  // events were recorded on the library's own streams too
    MSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    if (h->push_stream) MSIM_CUDA(h, cudaStreamSynchronize(h->push_stream));
    double ms[K_COUNT] = {0};
    uint64_t launches[K_COUNT] = {0};
    for (size_t i = 0; i < h->prof.ids.size(); i++) {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, h->prof.pool[2 * i], h->prof.pool[2 * i + 1]) == cudaSuccess) {
            ms[h->prof.ids[i]] += t;
            launches[h->prof.ids[i]]++;
        }
    }
    uint32_t n = 0;
    for (int k = 0; k < K_COUNT && n < cap; k++) {
        if (!launches[k]) continue;
        std::memset(&out[n], 0, sizeof(out[n]));
        std::strncpy(out[n].name, names[k], sizeof(out[n].name) - 1);
        out[n].launches = launches[k];
        out[n].total_ms = ms[k];
        n++;
    }
    *count = n;
    h->prof.used = 0;
    h->prof.ids.clear();
    return MSIM_OK;
}

int msim_get_device_view(msim_handle* h, msim_device_view* out) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!out) return fail(h, MSIM_ERR_INVALID, "msim_get_device_view: out is null");
    join_side(h);
    out->pos = h->pos[h->cur];
    out->target = h->target;
    out->road = h->road;
    out->rng = h->rng;
    out->count = h->n;
    out->ext_id = h->perm_active ? h->ext_id : nullptr;
    return MSIM_OK;
}


// ------------------------------------------------------------------------------------------------
// multi-GPU sharding (include/msim_shard.h)
// ------------------------------------------------------------------------------------------------
namespace {
ShardArrays shard_arrays(msim_handle* h) {
    ShardArrays a{};
    a.pos_cur = h->pos[h->cur];
    a.pos_prev = h->pos[h->cur ^ 1];
    a.target = h->target;
    a.road = h->road;
    a.rng = h->rng;
    a.color0 = h->color0;
    a.gid = h->gid;
    a.keys = h->keys;
    a.arrived = h->arrived;
    if (h->count_fused) {
        a.cell_count = h->cell_count;
        csort_band(h->grid.ncells, h->grid.ncx, h->band_lo, h->band_hi, h->grid.ncy, &a.c0, &a.c1);
    }
    return a;
}

int ensure_keys(msim_handle* h) {
    int rc = alloc_collision_buffers(h);
    if (rc != MSIM_OK) return rc;
    join_side(h);
    if (!h->keys_valid) {
        rc = refresh_counts(h);
        if (rc != MSIM_OK) return rc;
        h->launches += launch_keygen(h->stream, h->n, h->pos[h->cur], h->keys, h->grid, &h->prof);
        h->keys_valid = true;
        h->hist_valid = false;
        h->counts_valid = false;
    }
    return MSIM_OK;
}
}  // namespace

uint64_t msim_shard_buffer_bytes(uint32_t migrant_capacity, uint32_t halo_capacity) {
    return sizeof(ShardHeader) + static_cast<uint64_t>(migrant_capacity) * MIGRANT_BYTES + static_cast<uint64_t>(halo_capacity) * sizeof(float2);
}

int msim_shard_enable(msim_handle* h, const uint32_t* gids, uint64_t count, uint32_t migrant_capacity, uint32_t halo_capacity) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (count != h->n) return fail(h, MSIM_ERR_INVALID, "msim_shard_enable: count differs from the resident entity count");
    if (count && !gids) return fail(h, MSIM_ERR_INVALID, "msim_shard_enable: gids is null");
    if (migrant_capacity == 0 || halo_capacity == 0) return fail(h, MSIM_ERR_INVALID, "msim_shard_enable: zero capacity");
    if (h->flags & MSIM_FLAG_NO_COLLISIONS) return fail(h, MSIM_ERR_INVALID, "msim_shard_enable: collisions-off handles shard by entity range and need no exchange");
    if (h->sharded) return fail(h, MSIM_ERR_INVALID, "msim_shard_enable: already enabled");
    rc = alloc_collision_buffers(h);
    if (rc != MSIM_OK) return rc;
    h->mig_cap = migrant_capacity;
    h->halo_cap = halo_capacity;
    h->holes_cap = 2u * migrant_capacity;
    MSIM_CUDA(h, dev_alloc(&h->gid, h->cap));
    MSIM_CUDA(h, dev_alloc(&h->holes, h->holes_cap));
    MSIM_CUDA(h, dev_alloc(&h->local_ghosts, h->holes_cap));
    MSIM_CUDA(h, dev_alloc(&h->shard_ctr, SHARD_CTR_COUNT));
    MSIM_CUDA(h, dev_alloc(&h->place_dst, h->holes_cap));
    MSIM_CUDA(h, dev_alloc(&h->moves, h->holes_cap));
    MSIM_CUDA(h, dev_alloc(&h->row_hist, static_cast<size_t>(h->grid.ncy)));
    h->row_hist_rows = static_cast<uint32_t>(h->grid.ncy);
    MSIM_CUDA(h, dev_alloc(&h->dev_counts, DEV_ALLOC_WORDS));
    rc = write_dev_counts(h);
    if (rc != MSIM_OK) return rc;
    MSIM_CUDA(h, cudaMallocHost(reinterpret_cast<void**>(&h->host_stage), 2 * (static_cast<size_t>(h->holes_cap) * 4 + 64) * sizeof(uint32_t)));
    if (count) MSIM_CUDA(h, cudaMemcpyAsync(h->gid, gids, count * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    MSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    h->sharded = true;
    h->band_valid = false;
    h->perm_active = false;  // gid is the id map of a sharded handle; a cell re-sort permutes it with the state
    return ensure_cells(h);
}

int msim_shard_pack(msim_handle* h, uint32_t row_lo, uint32_t row_hi, void* send_down, void* send_up) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!h->sharded) return fail(h, MSIM_ERR_INVALID, "msim_shard_pack: call msim_shard_enable first");
    if (row_lo >= row_hi || row_hi > static_cast<uint32_t>(h->grid.ncy)) return fail(h, MSIM_ERR_INVALID, "msim_shard_pack: bad row range");
    join_side(h);  // the records carry target / road / rng, which pass B of the move may still be writing
    rc = ensure_keys(h);
    if (rc != MSIM_OK) return rc;
    h->n_ghost = 0;
    h->launches += launch_shard_pack(h->stream, shard_arrays(h), launch_owned(h), h->grid.ncx, row_lo, row_hi, send_down, send_up, h->mig_cap, h->halo_cap,
                                     h->holes, h->holes_cap, h->local_ghosts, h->shard_ctr, &h->prof, dev_owned(h));
    h->sent_down = send_down;
    h->sent_up = send_up;
    // rows that can hold owned entities once the leavers are gone: without a neighbour nobody leaves on that side
    h->band_lo = send_down ? row_lo : 0u;
    h->band_hi = send_up ? row_hi : static_cast<uint32_t>(h->grid.ncy);
    h->packed = true;
    return MSIM_OK;
}

namespace {
// arena of one handle: [flag from below @0][flag from above @128][256: receive buffers, index 2 * parity + side]
// side 0 = written by the neighbour below, side 1 = written by the neighbour above
constexpr size_t P2P_FLAG_BYTES = 256;
inline uint32_t* p2p_flag(char* arena, int side) { return reinterpret_cast<uint32_t*>(arena + 128 * side); }
inline char* p2p_recv(const msim_handle* h, char* arena, uint32_t parity, int side) { return arena + P2P_FLAG_BYTES + (2u * parity + side) * h->p2p_buf_bytes; }

struct P2PSignal {
    uint32_t* flag_down;
    uint32_t* flag_up;
    uint32_t value;
    void* peer_down;  // the neighbours' receive buffers of this tick (the send buffers handed to move_pack_common are local)
    void* peer_up;
};

// sender side of the peer-memory exchange, on its own stream behind the pack that has just been enqueued on the main one
int enqueue_push(msim_handle* h, const void* send_down, const void* send_up, const P2PSignal& sig, bool counts_in_headers) {
    MSIM_CUDA(h, cudaEventRecord(h->ev_packed_move, h->ms));
    MSIM_CUDA(h, cudaStreamWaitEvent(h->push_stream, h->ev_packed_move, 0));
    h->launches += launch_shard_push(h->push_stream, send_down, send_up, sig.peer_down, sig.peer_up, sig.flag_down, sig.flag_up, sig.value, h->mig_cap, h->halo_cap,
                                     h->holes_cap, h->shard_ctr, dev_error(h), h->p2p_push_ticket, counts_in_headers);
    MSIM_CUDA(h, cudaEventRecord(h->ev_pushed, h->push_stream));
    h->push_pending = true;
    return MSIM_OK;
}

int move_pack_common(msim_handle* h, const char* who, uint32_t row_lo, uint32_t row_hi, void* send_down, void* send_up, const P2PSignal* sig) {
    if (!h->sharded) return fail(h, MSIM_ERR_INVALID, std::string(who) + ": call msim_shard_enable first");
    if (row_lo >= row_hi || row_hi > static_cast<uint32_t>(h->grid.ncy)) return fail(h, MSIM_ERR_INVALID, std::string(who) + ": bad row range");
    if (h->awaiting_integrate) return fail(h, MSIM_ERR_INVALID, std::string(who) + ": the previous exchange has not been integrated");
    int rc = MSIM_OK;
    const bool init_only = h->uninitialised;  // the reference's first dispatch moves nobody (random_move.comp:863-867)
    if (init_only) consume_init_dispatch(h);
    // peer-memory exchange: the whole move phase (pass B of the previous tick, move + pack, exchange) runs beside the previous tick's
    // query; the collective exchange stays on the main stream, where the caller enqueues its collective
    const bool beside = sig && !init_only && h->side && tuning().overlap_ticks && !(h->flags & MSIM_FLAG_NO_COLLISIONS);
    const cudaStream_t ms = begin_move_phase(h, beside);  // (pass B of the previous tick first: the records read target / road / rng)
    if (h->push_pending) {  // the previous tick's push reads the counters that are about to be cleared (it finished long ago)
        MSIM_CUDA(h, cudaStreamWaitEvent(ms, h->ev_pushed, 0));
        h->push_pending = false;
    }
    // the fused kernel counts in h->shard_ctr (a one-thread kernel behind it, or the push kernel, writes the buffer headers); the
    // stand-alone pack kernel (init-only dispatch below) counts in the headers of the (local) buffers, which the reset clears
    h->launches += launch_shard_reset(ms, init_only ? send_down : nullptr, init_only ? send_up : nullptr, h->shard_ctr);
    ShardMoveArgs sh{};
    sh.lo_key = row_lo * static_cast<uint32_t>(h->grid.ncx);
    sh.hi_key = row_hi * static_cast<uint32_t>(h->grid.ncx);
    sh.ncx = static_cast<uint32_t>(h->grid.ncx);
    sh.buf_down = send_down;
    sh.buf_up = send_up;
    sh.mig_cap = h->mig_cap;
    sh.halo_cap = h->halo_cap;
    sh.holes_cap = h->holes_cap;
    sh.holes = h->holes;
    sh.local_ghosts = h->local_ghosts;
    sh.ctr = h->shard_ctr;
    sh.rng = h->rng;
    sh.color0 = h->color0;
    sh.road = h->road;
    sh.gid = h->gid;
    sh.error_word = dev_error(h);
    sh.publish = sig ? 0u : 1u;  // peer-memory exchange: the push kernel writes the headers where they are read
    // rows that can hold owned entities once the leavers are gone: without a neighbour nobody leaves on that side
    h->band_lo = send_down ? row_lo : 0u;
    h->band_hi = send_up ? row_hi : static_cast<uint32_t>(h->grid.ncy);
    if (init_only) {
        h->count_fused = false;
        // nobody moves: pack the resident positions with the stand-alone kernel, then raise the flags
        rc = ensure_keys(h);
        if (rc != MSIM_OK) return rc;
        h->launches += launch_shard_pack(h->stream, shard_arrays(h), launch_owned(h), h->grid.ncx, row_lo, row_hi, send_down, send_up, h->mig_cap, h->halo_cap,
                                         h->holes, h->holes_cap, h->local_ghosts, h->shard_ctr, &h->prof, dev_owned(h), /*reset=*/false);
        if (sig) {
            rc = enqueue_push(h, send_down, send_up, *sig, /*counts_in_headers=*/true);
            if (rc != MSIM_OK) return rc;
        }
    } else {
        rc = enqueue_move(h, true, &sh, ms);
        if (rc != MSIM_OK) return rc;
        if (h->shard_trace) launch_shard_stamp(ms, h->shard_trace, 5);
        if (sig) {
            rc = enqueue_push(h, send_down, send_up, *sig, /*counts_in_headers=*/false);
            if (rc != MSIM_OK) return rc;
        }
        h->awaiting_integrate = true;  // pass B is deferred until the exchange has been integrated
    }
    h->n_ghost = 0;
    h->sent_down = sig ? nullptr : send_down;  // peer buffers: their owner checks the overflow flag
    h->sent_up = sig ? nullptr : send_up;
    h->packed = true;
    end_move_phase(h);
    return MSIM_OK;
}

int integrate_device_common(msim_handle* h, const void* recv_down, const void* recv_up, const ShardWait* wait, bool launch = true) {
    if (launch)
        h->launches += launch_shard_integrate_device(h->stream, shard_arrays(h), flip_counts(h), h->sent_down, h->sent_up, recv_down, recv_up, h->holes,
                                                     h->shard_ctr, h->local_ghosts, h->mig_cap, h->halo_cap, h->holes_cap, h->cap, h->place_dst, h->moves,
                                                     h->grid, &h->prof, wait);
    h->async_counts = true;  // from here on kernels take their counts from device memory; the host values are refreshed on demand
    h->band_valid = h->packed;
    h->packed = false;
    h->awaiting_integrate = false;
    h->keys_valid = true;
    h->hist_valid = false;
    h->counts_valid = h->count_fused;  // ranks and per-cell counters were completed by the placement kernels
    h->flags_stale = h->collided;
    return MSIM_OK;
}
}  // namespace

int msim_shard_move_pack(msim_handle* h, uint32_t row_lo, uint32_t row_hi, void* send_down, void* send_up) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    return move_pack_common(h, "msim_shard_move_pack", row_lo, row_hi, send_down, send_up, nullptr);
}

// ---- peer-memory exchange ---------------------------------------------------------------------------
int msim_shard_p2p_create(msim_handle* h, void* ipc_handle_out, void** arena_out) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!h->sharded) return fail(h, MSIM_ERR_INVALID, "msim_shard_p2p_create: call msim_shard_enable first");
    if (!h->p2p_arena) {
        h->p2p_buf_bytes = (msim_shard_buffer_bytes(h->mig_cap, h->halo_cap) + 255ull) & ~255ull;
        const size_t bytes = P2P_FLAG_BYTES + 4 * h->p2p_buf_bytes;
        MSIM_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->p2p_arena), bytes));
        MSIM_CUDA(h, cudaMemsetAsync(h->p2p_arena, 0, bytes, h->stream));
        MSIM_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->p2p_send), 4 * h->p2p_buf_bytes));
        MSIM_CUDA(h, cudaMemsetAsync(h->p2p_send, 0, 4 * h->p2p_buf_bytes, h->stream));
        MSIM_CUDA(h, dev_alloc(&h->p2p_push_ticket, 1));
        MSIM_CUDA(h, cudaMemsetAsync(h->p2p_push_ticket, 0, sizeof(uint32_t), h->stream));
        MSIM_CUDA(h, cudaStreamCreateWithFlags(&h->push_stream, cudaStreamNonBlocking));
        MSIM_CUDA(h, cudaEventCreateWithFlags(&h->ev_packed_move, cudaEventDisableTiming));
        MSIM_CUDA(h, cudaEventCreateWithFlags(&h->ev_pushed, cudaEventDisableTiming));
        MSIM_CUDA(h, cudaStreamSynchronize(h->stream));  // neighbours may write as soon as they hold the pointer
        if (const char* env = std::getenv("MSIM_P2P_TIMEOUT_MS")) {
            const long v = std::atol(env);
            if (v > 0) h->p2p_timeout_ns = static_cast<unsigned long long>(v) * 1000000ull;
        }
        if (const char* env = std::getenv("MSIM_SHARD_TRACE")) {
            if (std::atoi(env) == 1 && !h->shard_trace) {
                MSIM_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->shard_trace), 16 * sizeof(unsigned long long)));
                MSIM_CUDA(h, cudaMemsetAsync(h->shard_trace, 0, 16 * sizeof(unsigned long long), h->stream));
            }
        }
    }
    if (ipc_handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == MSIM_P2P_HANDLE_BYTES, "cudaIpcMemHandle_t size");
        cudaIpcMemHandle_t ipc{};
        MSIM_CUDA(h, cudaIpcGetMemHandle(&ipc, h->p2p_arena));
        std::memcpy(ipc_handle_out, &ipc, sizeof(ipc));
    }
    if (arena_out) *arena_out = h->p2p_arena;
    return MSIM_OK;
}

int msim_shard_p2p_connect(msim_handle* h, const void* down_ipc_handle, const void* up_ipc_handle) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!h->p2p_arena) return fail(h, MSIM_ERR_INVALID, "msim_shard_p2p_connect: call msim_shard_p2p_create first");
    if (h->p2p_connected) return fail(h, MSIM_ERR_INVALID, "msim_shard_p2p_connect: already connected");
    auto open = [&](const void* src, char** out) -> int {
        cudaIpcMemHandle_t ipc{};
        std::memcpy(&ipc, src, sizeof(ipc));
        void* p = nullptr;
        MSIM_CUDA(h, cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
        *out = static_cast<char*>(p);
        return MSIM_OK;
    };
    if (down_ipc_handle) {
        rc = open(down_ipc_handle, &h->p2p_peer_down);
        if (rc != MSIM_OK) return rc;
        h->p2p_down_ipc = true;
    }
    if (up_ipc_handle) {
        rc = open(up_ipc_handle, &h->p2p_peer_up);
        if (rc != MSIM_OK) return rc;
        h->p2p_up_ipc = true;
    }
    h->p2p_connected = true;
    return MSIM_OK;
}

int msim_shard_p2p_connect_local(msim_handle* h, void* down_arena, void* up_arena) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!h->p2p_arena) return fail(h, MSIM_ERR_INVALID, "msim_shard_p2p_connect_local: call msim_shard_p2p_create first");
    if (h->p2p_connected) return fail(h, MSIM_ERR_INVALID, "msim_shard_p2p_connect_local: already connected");
    h->p2p_peer_down = static_cast<char*>(down_arena);
    h->p2p_peer_up = static_cast<char*>(up_arena);
    h->p2p_connected = true;
    // (Neighbours driven by ONE process may share a stream: the move kernels raise the flags themselves, so as long as every band's
    // move + pack is enqueued before any band's integrate, no kernel waits for one queued behind it.)
    return MSIM_OK;
}

int msim_shard_p2p_move_pack(msim_handle* h, uint32_t row_lo, uint32_t row_hi) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!h->p2p_connected) return fail(h, MSIM_ERR_INVALID, "msim_shard_p2p_move_pack: call msim_shard_p2p_connect first");
    const uint32_t parity = h->p2p_tick & 1u;
    // we are the neighbour ABOVE the rank below us (its side 1) and the neighbour BELOW the rank above us (its side 0)
    P2PSignal sig{};
    sig.flag_down = h->p2p_peer_down ? p2p_flag(h->p2p_peer_down, 1) : nullptr;
    sig.flag_up = h->p2p_peer_up ? p2p_flag(h->p2p_peer_up, 0) : nullptr;
    sig.value = h->p2p_tick + 1u;
    sig.peer_down = h->p2p_peer_down ? p2p_recv(h, h->p2p_peer_down, parity, 1) : nullptr;
    sig.peer_up = h->p2p_peer_up ? p2p_recv(h, h->p2p_peer_up, parity, 0) : nullptr;
    // the pack fills local send buffers (two sets, by tick parity); the push kernel copies them into the neighbours' arenas
    return move_pack_common(h, "msim_shard_p2p_move_pack", row_lo, row_hi, h->p2p_peer_down ? h->p2p_send + (2u * parity) * h->p2p_buf_bytes : nullptr,
                            h->p2p_peer_up ? h->p2p_send + (2u * parity + 1u) * h->p2p_buf_bytes : nullptr, &sig);
}

int msim_shard_p2p_integrate(msim_handle* h) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!h->p2p_connected) return fail(h, MSIM_ERR_INVALID, "msim_shard_p2p_integrate: call msim_shard_p2p_connect first");
    if (!h->packed) return fail(h, MSIM_ERR_INVALID, "msim_shard_p2p_integrate: call msim_shard_p2p_move_pack first");
    const uint32_t parity = h->p2p_tick & 1u;
    void* recv_down = h->p2p_peer_down ? p2p_recv(h, h->p2p_arena, parity, 0) : nullptr;
    void* recv_up = h->p2p_peer_up ? p2p_recv(h, h->p2p_arena, parity, 1) : nullptr;
    ShardWait w{};
    w.flag_down = recv_down ? p2p_flag(h->p2p_arena, 0) : nullptr;
    w.flag_up = recv_up ? p2p_flag(h->p2p_arena, 1) : nullptr;
    w.expected = h->p2p_tick + 1u;
    w.timeout_ns = h->p2p_timeout_ns;
    h->p2p_tick++;
    const cudaStream_t ms = h->ms ? h->ms : h->stream;  // the stream the move + pack of this tick ran on
    h->launches += launch_shard_exchange(ms, shard_arrays(h), flip_counts(h), recv_down, recv_up, h->holes,
                                         h->shard_ctr, h->local_ghosts, h->mig_cap, h->halo_cap, h->holes_cap, h->cap, h->place_dst, h->moves, h->grid,
                                         &h->prof, w, h->shard_trace);
    if (h->shard_trace) launch_shard_stamp(ms, h->shard_trace, 9);
    rc = integrate_device_common(h, recv_down, recv_up, nullptr, /*launch=*/false);
    // pass B needs the integrated population and nothing of the rebuild: it follows the exchange at once (side stream), beside the scan and
    // the scatter of the collision pass on the main stream, instead of waiting for them
    if (rc == MSIM_OK && tuning().shard_arrive_early && ms == h->side && h->side && h->arrive_deferred && !h->awaiting_integrate) {
        h->arrive_deferred = false;
        h->launches += launch_arrive(ms, launch_owned(h), h->target, h->road, h->rng, h->arrived, h->roads, h->conn, h->conn_count, &h->prof, dev_owned(h),
                                     /*beside=*/true, tuning().shard_arrive_beside_ctas_per_sm);
        h->arrive_early = true;
    }
    end_move_phase(h);
    return rc;
}

int msim_shard_integrate(msim_handle* h, const void* recv_down, const void* recv_up, uint64_t* owned, uint64_t* ghosts) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!h->sharded) return fail(h, MSIM_ERR_INVALID, "msim_shard_integrate: call msim_shard_enable first");
    rc = refresh_counts(h);
    if (rc != MSIM_OK) return rc;
    // one host round trip: counters, the four headers and the hole list
    // two staging areas used alternately: the H2D copies of tick t may still be in flight when tick t+1 stages
    const size_t stage_words = static_cast<size_t>(h->holes_cap) * 4 + 64;
    uint32_t* hs = h->host_stage + (h->stage_flip ? stage_words : 0);
    h->stage_flip = !h->stage_flip;
    uint32_t* h_ctr = hs;            // [8]
    uint32_t* h_hdr = hs + 8;        // 4 x 8 words: sent_down, sent_up, recv_down, recv_up
    uint32_t* h_holes = hs + 64;     // [holes_cap]
    std::memset(hs, 0, 64 * sizeof(uint32_t));
    MSIM_CUDA(h, cudaMemcpyAsync(h_ctr, h->shard_ctr, SHARD_CTR_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    const void* hdrs[4] = {h->sent_down, h->sent_up, recv_down, recv_up};
    for (int i = 0; i < 4; i++)
        if (hdrs[i]) MSIM_CUDA(h, cudaMemcpyAsync(h_hdr + 8 * i, hdrs[i], sizeof(ShardHeader), cudaMemcpyDeviceToHost, h->stream));
    MSIM_CUDA(h, cudaMemcpyAsync(h_holes, h->holes, static_cast<size_t>(h->holes_cap) * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    MSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    const uint32_t k_out = h_ctr[SHARD_CTR_HOLES], g_local = h_ctr[SHARD_CTR_LOCAL_GHOSTS];
    for (int i = 0; i < 4; i++)
        if (h_hdr[8 * i + 2]) return fail(h, MSIM_ERR_CAPACITY, "shard exchange buffer overflow: raise migrant_capacity / halo_capacity");
    if (k_out > h->holes_cap) return fail(h, MSIM_ERR_CAPACITY, "more leavers than hole capacity");
    const uint32_t in_down = recv_down ? h_hdr[16] : 0, halo_down = recv_down ? h_hdr[17] : 0;
    const uint32_t in_up = recv_up ? h_hdr[24] : 0, halo_up = recv_up ? h_hdr[25] : 0;
    if (in_down > h->mig_cap || in_up > h->mig_cap || halo_down > h->halo_cap || halo_up > h->halo_cap)
        return fail(h, MSIM_ERR_CAPACITY, "received shard buffer exceeds the configured capacities");
    const uint32_t k_in = in_down + in_up;
    if (k_in > h->holes_cap) return fail(h, MSIM_ERR_CAPACITY, "more arrivals than placement capacity");
    const uint32_t n_old = h->n;
    const uint64_t n_new64 = static_cast<uint64_t>(n_old) + k_in - k_out;
    const uint32_t n_ghost = halo_down + halo_up + g_local;
    if (n_new64 + n_ghost > h->cap) return fail(h, MSIM_ERR_CAPACITY, "entity_capacity too small for arrivals + ghosts");
    const uint32_t n_new = static_cast<uint32_t>(n_new64);

    // placement: arrivals fill holes first, then append; left-over holes are closed from the tail
    uint32_t* h_dst = hs + 64 + h->holes_cap;                                 // [holes_cap]
    uint2* h_moves = reinterpret_cast<uint2*>(hs + 64 + 2 * h->holes_cap);  // [holes_cap] pairs
    for (uint32_t i = 0; i < k_in; i++) h_dst[i] = i < k_out ? h_holes[i] : n_old + (i - k_out);
    uint32_t n_moves = 0;
    if (k_out > k_in) {
        // holes still open: h_holes[k_in .. k_out).  Live entities in [n_new, n_old) move into holes < n_new.
        std::vector<uint32_t> open(h_holes + k_in, h_holes + k_out);
        std::sort(open.begin(), open.end());
        std::vector<uint32_t> low;
        size_t tail_holes_begin = std::lower_bound(open.begin(), open.end(), n_new) - open.begin();
        low.assign(open.begin(), open.begin() + tail_holes_begin);
        size_t hp = tail_holes_begin;  // walks the holes inside the tail
        for (uint32_t src = n_new; src < n_old && n_moves < low.size(); src++) {
            if (hp < open.size() && open[hp] == src) {
                hp++;
                continue;
            }
            h_moves[n_moves] = make_uint2(src, low[n_moves]);
            n_moves++;
        }
        if (n_moves != low.size()) return fail(h, MSIM_ERR_INTERNAL, "shard compaction bookkeeping mismatch");
    }
    if (k_in) MSIM_CUDA(h, cudaMemcpyAsync(h->place_dst, h_dst, k_in * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    if (n_moves) MSIM_CUDA(h, cudaMemcpyAsync(h->moves, h_moves, n_moves * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
    const ShardArrays a = shard_arrays(h);
    h->launches += launch_shard_place(h->stream, a, recv_down, in_down, recv_up, in_up, h->place_dst, h->grid, &h->prof);
    h->launches += launch_shard_relocate(h->stream, a, h->moves, n_moves, &h->prof);
    h->launches += launch_shard_append_ghosts(h->stream, a, n_new, recv_down, halo_down, recv_up, halo_up, h->local_ghosts, g_local, h->mig_cap,
                                              h->grid, &h->prof);
    h->n = n_new;
    h->n_ghost = n_ghost;
    rc = write_dev_counts(h);
    if (rc != MSIM_OK) return rc;
    h->band_valid = h->packed;
    h->packed = false;
    h->awaiting_integrate = false;
    h->keys_valid = true;
    h->hist_valid = false;
    h->counts_valid = h->count_fused;  // ranks and per-cell counters were completed by the placement kernels
    h->flags_stale = h->collided;
    if (owned) *owned = n_new;
    if (ghosts) *ghosts = n_ghost;
    return MSIM_OK;
}

int msim_shard_integrate_async(msim_handle* h, const void* recv_down, const void* recv_up) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!h->sharded) return fail(h, MSIM_ERR_INVALID, "msim_shard_integrate_async: call msim_shard_enable first");
    return integrate_device_common(h, recv_down, recv_up, nullptr);
}

int msim_shard_counts(msim_handle* h, uint64_t* owned, uint64_t* ghosts) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    rc = refresh_counts(h);
    if (rc != MSIM_OK) return rc;
    if (owned) *owned = h->n;
    if (ghosts) *ghosts = h->n_ghost;
    return MSIM_OK;
}

int msim_shard_read_gids(msim_handle* h, uint32_t* dst, uint64_t count) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!h->sharded) return fail(h, MSIM_ERR_INVALID, "msim_shard_read_gids: call msim_shard_enable first");
    rc = refresh_counts(h);
    if (rc != MSIM_OK) return rc;
    if (count > h->n || (count && !dst)) return fail(h, MSIM_ERR_INVALID, "msim_shard_read_gids: bad arguments");
    if (count) MSIM_CUDA(h, cudaMemcpyAsync(dst, h->gid, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    MSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    return MSIM_OK;
}

int msim_shard_row_histogram(msim_handle* h, uint32_t* dst, uint32_t rows) {
    int rc = bind(h);
    if (rc != MSIM_OK) return rc;
    if (!dst || rows != static_cast<uint32_t>(h->grid.ncy)) return fail(h, MSIM_ERR_INVALID, "msim_shard_row_histogram: rows must equal grid_cells_y");
    if (!h->row_hist || h->row_hist_rows < rows) {  // first use, or the grid has grown since (radius / world change through push constants)
        cudaFree(h->row_hist);
        h->row_hist = nullptr;
        h->row_hist_rows = 0;
        MSIM_CUDA(h, dev_alloc(&h->row_hist, static_cast<size_t>(rows)));
        h->row_hist_rows = rows;
    }
    rc = refresh_counts(h);
    if (rc != MSIM_OK) return rc;
    rc = ensure_keys(h);
    if (rc != MSIM_OK) return rc;
    h->launches += launch_shard_row_histogram(h->stream, h->keys, h->n, h->grid.ncx, h->row_hist, rows, &h->prof);
    MSIM_CUDA(h, cudaMemcpyAsync(dst, h->row_hist, static_cast<size_t>(rows) * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    MSIM_CUDA(h, cudaStreamSynchronize(h->stream));
    return MSIM_OK;
}

int msim_grid_rows(float world_w, float world_h, float radius, const float* xy, uint64_t count, uint32_t* rows_out, uint32_t* cells_x, uint32_t* cells_y) {
    GridParams g{};
    int bits = 0;
    compute_grid(world_w, world_h, radius, g, bits);
    if (cells_x) *cells_x = static_cast<uint32_t>(g.ncx);
    if (cells_y) *cells_y = static_cast<uint32_t>(g.ncy);
    if (count && (!xy || !rows_out)) return MSIM_ERR_INVALID;
    for (uint64_t i = 0; i < count; i++) {
        const float prod = xy[2 * i + 1] * g.inv_cell;  // the device's cell_key_of: one binary32 multiply, floor, clamp
        int cy = static_cast<int>(std::floor(prod));
        if (!(prod == prod)) cy = 0;  // NaN -> 0, like __float2int_rd
        cy = std::min(std::max(cy, 0), g.ncy - 1);
        rows_out[i] = static_cast<uint32_t>(cy);
    }
    return MSIM_OK;
}

int msim_grid_params(float world_w, float world_h, float radius, float* inv_cell, float* hit_threshold, uint32_t* cells_x, uint32_t* cells_y) {
    GridParams g{};
    int bits = 0;
    compute_grid(world_w, world_h, radius, g, bits);
    if (inv_cell) *inv_cell = g.inv_cell;
    if (hit_threshold) *hit_threshold = g.hit_threshold;
    if (cells_x) *cells_x = static_cast<uint32_t>(g.ncx);
    if (cells_y) *cells_y = static_cast<uint32_t>(g.ncy);
    return MSIM_OK;
}

// Host-only twin of msim_read_quadtree_nodes: the same tree from caller-owned positions, the leaf histogram taken on the host with the
// descent quadtree.cu uses on the device (compare with offset + width / 2 in binary32, halve, repeat: random_move.comp:319-341).  Lets the
// display quadtree be checked without a GPU (tests/test_quadtree_host.py: against the tree the reference's own shader code builds).
int msim_quadtree_from_positions(const float* xy, uint64_t count_in, float world_w, float world_h, uint32_t max_depth, uint32_t node_cap,
                                 msim_quadtree_node* dst, uint64_t cap, uint64_t* count) {
    if (!dst || cap == 0 || !count || (count_in && !xy)) return MSIM_ERR_INVALID;
    const int depth = static_cast<int>(std::min<uint32_t>(std::max<uint32_t>(max_depth ? max_depth : 8u, 1u), 8u));
    const int levels = depth - 1;
    const bool root_only = count_in == 0 || levels == 0;
    std::vector<std::vector<uint32_t>> sums(levels + 1);
    if (!root_only) {
        const uint32_t side0 = 1u << levels;
        sums[0].assign(static_cast<size_t>(side0) * side0, 0u);
        auto descend = [levels](float x, float extent) {
            float off = 0.0f, width = extent;
            uint32_t index = 0;
            for (int l = 0; l < levels; l++) {
                width = width * 0.5f;
                const float mid = off + width;
                index <<= 1;
                if (!(x < mid)) {
                    off = mid;
                    index |= 1u;
                }
            }
            return index;
        };
        for (uint64_t i = 0; i < count_in; i++) sums[0][static_cast<size_t>(descend(xy[2 * i + 1], world_h)) * side0 + descend(xy[2 * i], world_w)]++;
        for (int l = 1; l <= levels; l++) {
            const uint32_t side = 1u << (levels - l);
            sums[l].resize(static_cast<size_t>(side) * side);
            const std::vector<uint32_t>& f = sums[l - 1];
            const size_t fs = static_cast<size_t>(side) * 2;
            for (uint32_t y = 0; y < side; y++)
                for (uint32_t x = 0; x < side; x++)
                    sums[l][static_cast<size_t>(y) * side + x] = f[(2 * y) * fs + 2 * x] + f[(2 * y) * fs + 2 * x + 1] + f[(2 * y + 1) * fs + 2 * x] + f[(2 * y + 1) * fs + 2 * x + 1];
        }
    } else {
        sums.assign(1, std::vector<uint32_t>(1, static_cast<uint32_t>(count_in)));
    }
    QuadBuilder b{&sums, root_only ? 0 : levels, node_cap ? node_cap : 10u, dst, cap, 0, false};
    b.emit(root_only ? 0 : levels, 0, 0, 0.0f, 0.0f, world_w, world_h, 0);
    if (b.overflow) return MSIM_ERR_CAPACITY;
    *count = b.used;
    return MSIM_OK;
}

}  // extern "C"
