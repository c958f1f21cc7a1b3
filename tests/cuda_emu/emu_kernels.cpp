// tests/cuda_emu/emu_kernels.cpp - TEST INFRASTRUCTURE: the kernels of movement-sim_b200/csrc compiled for the host SIMT emulator
// (tests/cuda_emu/cuda_runtime.h) with C entry points that launch single kernels, for tests/test_kernels_under_emulator.py.  Built by
// tests/cuda_emu/build_emu_lib.py next to the whole-library build (libmsim_emu.so).
#include "cuda_runtime.h"

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
namespace cuda_emu {
thread_local Block* block = nullptr;
thread_local unsigned lane = 0, warp = 0;
thread_local void* dynamic_smem_ptr = nullptr;
}  // namespace cuda_emu

// the scratch copies build_emu_lib.py writes: the sources with <<<...>>> rewritten to cuda_emu::cfg(...)(...)
#include "_gen/movement-sim_b200/csrc/move.cpp"
#include "_gen/movement-sim_b200/csrc/collide_paired.cpp"
#include "_gen/movement-sim_b200/csrc/csort.cpp"

namespace msim {
const Tuning& tuning() {
    static const Tuning t;
    return t;
}
}  // namespace msim

using namespace msim;

extern "C" {

// move_kernel<keys = false, shard = false, FUSE = fuse>: one move pass over n entities (SoA), `blocks` CTAs of 256 threads
void emu_move(uint32_t n, const float* pos_in, float* pos_out, float* target, uint32_t* arrived, uint32_t* road, uint32_t* rng,
              const void* roads, const uint32_t* conn, uint64_t conn_count, int fuse, int consume, uint32_t blocks) {
    const ShardMoveArgs none{};
    FusedArrive fa{};
    fa.target = reinterpret_cast<float2*>(target);
    fa.road = road;
    fa.rng = reinterpret_cast<uint4*>(rng);
    fa.roads = static_cast<const uint4*>(roads);
    fa.conn = conn;
    fa.conn_count = conn_count;
    fa.consume = consume ? 1u : 0u;
    GridParams grid{};
    const float4* pin = reinterpret_cast<const float4*>(pos_in);
    float4* pout = reinterpret_cast<float4*>(pos_out);
    const float4* tgt = reinterpret_cast<const float4*>(target);
    if (fuse)
        cuda_emu::launch(move_kernel<false, false, true, 0>, blocks, MOVE_THREADS, n, static_cast<const uint32_t*>(nullptr), pin, pout, static_cast<const float4*>(nullptr),
                         arrived, static_cast<uint2*>(nullptr), grid, static_cast<uint32_t*>(nullptr), 0, static_cast<uint32_t*>(nullptr), static_cast<uint2*>(nullptr), none, fa);
    else
        cuda_emu::launch(move_kernel<false, false, false, 0>, blocks, MOVE_THREADS, n, static_cast<const uint32_t*>(nullptr), pin, pout, tgt, arrived,
                         static_cast<uint2*>(nullptr), grid, static_cast<uint32_t*>(nullptr), 0, static_cast<uint32_t*>(nullptr), static_cast<uint2*>(nullptr), none, fa);
}

// arrive_kernel<stride>: pass B; blocks = 0 means "as many as the words need"
void emu_arrive(uint32_t n, const uint32_t* arrived, float* target, uint32_t* road, uint32_t* rng, const void* roads, const uint32_t* conn,
                uint64_t conn_count, int stride, uint32_t blocks) {
    const uint32_t words = ((n + 63u) >> 6) << 1;
    const uint32_t need = (words + ARRIVE_THREADS - 1) / ARRIVE_THREADS;
    if (stride)
        cuda_emu::launch(arrive_kernel<true>, blocks ? blocks : need, ARRIVE_THREADS, n, static_cast<const uint32_t*>(nullptr), arrived, reinterpret_cast<float2*>(target), road,
                         reinterpret_cast<uint4*>(rng), static_cast<const uint4*>(roads), conn, conn_count);
    else
        cuda_emu::launch(arrive_kernel<false>, need, ARRIVE_THREADS, n, static_cast<const uint32_t*>(nullptr), arrived, reinterpret_cast<float2*>(target), road,
                         reinterpret_cast<uint4*>(rng), static_cast<const uint4*>(roads), conn, conn_count);
}

// query_paired_kernel over n sorted slots; stripes = 64 x 16 u64 (hits at [s * 16], pairs at [s * 16 + 1])
void emu_query_paired(uint32_t n, const float* sorted_pos, const uint32_t* cell_start, uint8_t* flag_sorted, float inv_cell, float hit_threshold,
                      float radius, int ncx, int ncy, unsigned long long* stripes) {
    GridParams grid{};
    grid.inv_cell = inv_cell;
    grid.hit_threshold = hit_threshold;
    grid.radius = radius;
    grid.ncx = ncx;
    grid.ncy = ncy;
    grid.ncells = static_cast<uint32_t>(ncx) * static_cast<uint32_t>(ncy);
    if (n == 0) return;
    cuda_emu::launch(query_paired_kernel, (n + 2 * PAIRED_THREADS - 1) / (2 * PAIRED_THREADS), PAIRED_THREADS, n, reinterpret_cast<const float2*>(sorted_pos), cell_start,
                     flag_sorted, grid, stripes);
}

uint32_t emu_query_window(void) { return QUERY_WINDOW; }

// move_kernel<keys = true, shard = false, FUSE = fuse> with the counting sort's rank fused in (cell keys, per-cell counters, rank per entity)
void emu_move_keys(uint32_t n, const float* pos_in, float* pos_out, float* target, uint32_t* arrived, uint32_t* road, uint32_t* rng, const void* roads,
                   const uint32_t* conn, uint64_t conn_count, int fuse, int consume, uint32_t blocks, uint32_t* keys, uint32_t* cell_count, uint32_t* rank,
                   float inv_cell, int ncx, int ncy) {
    const ShardMoveArgs none{};
    FusedArrive fa{};
    fa.target = reinterpret_cast<float2*>(target);
    fa.road = road;
    fa.rng = reinterpret_cast<uint4*>(rng);
    fa.roads = static_cast<const uint4*>(roads);
    fa.conn = conn;
    fa.conn_count = conn_count;
    fa.consume = consume ? 1u : 0u;
    GridParams grid{};
    grid.inv_cell = inv_cell;
    grid.ncx = ncx;
    grid.ncy = ncy;
    grid.ncells = static_cast<uint32_t>(ncx) * static_cast<uint32_t>(ncy);
    const float4* pin = reinterpret_cast<const float4*>(pos_in);
    float4* pout = reinterpret_cast<float4*>(pos_out);
    uint2* keys2 = reinterpret_cast<uint2*>(keys);
    uint2* rank2 = reinterpret_cast<uint2*>(rank);
    if (fuse)
        cuda_emu::launch(move_kernel<true, false, true, 0>, blocks, MOVE_THREADS, n, static_cast<const uint32_t*>(nullptr), pin, pout, static_cast<const float4*>(nullptr),
                         arrived, keys2, grid, static_cast<uint32_t*>(nullptr), 0, cell_count, rank2, none, fa);
    else
        cuda_emu::launch(move_kernel<true, false, false, 0>, blocks, MOVE_THREADS, n, static_cast<const uint32_t*>(nullptr), pin, pout,
                         reinterpret_cast<const float4*>(target), arrived, keys2, grid, static_cast<uint32_t*>(nullptr), 0, cell_count, rank2, none, fa);
}

// scan_tile_sums + scan_tiles<MINB> (the counters are zeroed as they are consumed) and cell_scatter: the counting sort behind the cell directory
void emu_scan_scatter(uint32_t n, uint32_t cells, uint32_t* cell_count, uint32_t* tile_sums, uint32_t* cell_start, const uint32_t* keys, const uint32_t* rank,
                      const float* pos, float* sorted_pos, uint32_t* sorted_idx, int scan_min_blocks_8, uint32_t scatter_blocks) {
    const uint32_t tiles = csort_tiles(cells);
    cuda_emu::launch(scan_tile_sums_kernel, tiles, SCAN_THREADS, static_cast<const uint32_t*>(cell_count), cells, tile_sums);
    if (scan_min_blocks_8)
        cuda_emu::launch(scan_tiles_kernel<8>, tiles, SCAN_THREADS, cell_count, cells, static_cast<const uint32_t*>(tile_sums), cell_start);
    else
        cuda_emu::launch(scan_tiles_kernel<0>, tiles, SCAN_THREADS, cell_count, cells, static_cast<const uint32_t*>(tile_sums), cell_start);
    cuda_emu::launch(cell_scatter_kernel, scatter_blocks, 256u, n, static_cast<const uint32_t*>(nullptr), reinterpret_cast<const uint2*>(keys),
                     reinterpret_cast<const uint2*>(rank), reinterpret_cast<const float4*>(pos), static_cast<const uint32_t*>(cell_start),
                     reinterpret_cast<float2*>(sorted_pos), sorted_idx);
}

}  // extern "C"
