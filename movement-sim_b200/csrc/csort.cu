// csort.cu — single-digit radix sort of the cell keys (a counting sort over the whole key): the
// opt-in (MSIM_FLAG_SORT_COUNTING) per-tick rebuild of the neighbour structure for populations whose
// storage order is spatially coherent; the per-cell counter table must fit the L2
// (Munich: 4.77 M cells x 4 B = 19 MB of B200's 126 MB).
//
// An LSD radix sort needs one pass per digit; with a table as large as the key space the digit IS the
// key and one pass suffices:
//   count    rank[e] = atomicAdd(&cell_count[key[e]], 1)   — fused into the move kernel (move.cu), or
//            cell_count_kernel below when the keys did not come from a move pass (sharded / keygen)
//   scan     cell_start = exclusive prefix sum of cell_count (two small kernels over the table)
//   scatter  slot = cell_start[key] + rank: sorted_pos[slot] = pos[e], sorted_idx[slot] = e
// Per entity that is key W4 + rank W4 in the move pass and key R4 + rank R4 + pos R8 + pos W8 + idx W4
// in the scatter = 36 B, against 44 B + 24 B for three onesweep passes plus the gather — and far fewer
// instructions (no ranking by warp votes).  The prefix table doubles as the cell directory: the run
// of cells x0..x1 of a row is [cell_start[row*ncx+x0], cell_start[row*ncx+x1+1]).
// The order of entities inside one cell is the arrival order of the atomics (not deterministic);
// nothing observable depends on it: flags and the unique-pair count are order-independent.
// Measured (profiles/r1_sort_paths.md): with entities in random index order the 20 M scattered 4/8-byte
// stores cost 484 us at 10 M entities against 3 x 90 us for the staged onesweep scatter, so onesweep
// (sort.cu) stays the default; with cell-ordered storage the scatter drops to 92 us and this path wins.
#include "msim_internal.h"

namespace msim {
namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096 cells per CTA

__global__ void __launch_bounds__(256)
cell_count_kernel(uint32_t n_host, const uint32_t* __restrict__ n_dev, const uint32_t* __restrict__ keys, uint32_t* __restrict__ cell_count,
                  uint32_t* __restrict__ rank, uint32_t c0, uint32_t c1) {
    const uint32_t n = n_dev ? *n_dev : n_host;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint32_t k = __ldcs(keys + e);
        // keys outside the counted cell range [c0, c1) (sharded handles only) stay out of the order
        rank[e] = (k - c0 < c1 - c0) ? atomicAdd(&cell_count[k], 1u) : CSORT_SKIP;
    }
}

__device__ __forceinline__ uint32_t block_reduce_sum(uint32_t v, uint32_t* s_warp) {
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) total += s_warp[w];
    __syncthreads();
    return total;
}

// phase 1: per-tile totals
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const uint32_t* __restrict__ counts, uint32_t cells, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint32_t c = base + i * SCAN_THREADS + threadIdx.x;
        if (c < cells) v += __ldcs(counts + c);
    }
    const uint32_t total = block_reduce_sum(v, s_warp);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// phase 2: exclusive scan inside each tile + the sum of the earlier tiles' totals.  Thread t owns SCAN_ITEMS consecutive cells.
// MINB = 8 caps the kernel at 32 registers (31 used, nothing spilled) so that 8 CTAs fit one SM: Munich's 1165 tiles are then ONE
// resident wave; with the 34 registers of MINB = 0 (the measured default) 6 CTAs fit and 277 tiles wait for a second wave.
template <int MINB>
__global__ void __launch_bounds__(SCAN_THREADS, MINB)
scan_tiles_kernel(uint32_t* __restrict__ counts, uint32_t cells, const uint32_t* __restrict__ tile_offsets /* per-tile TOTALS */,
                  uint32_t* __restrict__ starts) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32], s_before[SCAN_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t first = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
    // the counters are consumed here: they are zeroed on the way, so the next tick's count needs no memset of the table
    if (first + SCAN_ITEMS <= cells) {
        uint4* src = reinterpret_cast<uint4*>(counts + first);  // first is a multiple of 16 and the base 16-byte aligned
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; q++) {
            const uint4 c = __ldcs(src + q);
            v[4 * q] = c.x; v[4 * q + 1] = c.y; v[4 * q + 2] = c.z; v[4 * q + 3] = c.w;
            src[q] = make_uint4(0u, 0u, 0u, 0u);
        }
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            v[i] = 0u;
            if (first + i < cells) {
                v[i] = counts[first + i];
                counts[first + i] = 0u;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) sum += v[i];
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= static_cast<uint32_t>(d)) incl += up;
    }
    // offset of this tile = sum of the totals of the tiles before it (at most ~1.2 k words, L2-resident): cheaper than
    // a third kernel that scans the totals with one CTA between the two passes over the table
    uint32_t before = 0;
    for (uint32_t t = threadIdx.x; t < blockIdx.x; t += SCAN_THREADS) before += __ldcg(tile_offsets + t);
    before = __reduce_add_sync(0xffffffffu, before);
    if (lane == 31) s_warp[warp] = incl;
    if (lane == 0) s_before[warp] = before;
    __syncthreads();
    uint32_t run = incl - sum;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        run += s_before[w];
        if (static_cast<uint32_t>(w) < warp) run += s_warp[w];
    }
    // starts has cells + 1 entries: entry `cells` receives the grand total
    if (first + SCAN_ITEMS <= cells) {
        uint4* dst = reinterpret_cast<uint4*>(starts + first);
#pragma unroll
        for (int q = 0; q < SCAN_ITEMS / 4; q++) {
            uint4 o;
            o.x = run; run += v[4 * q];
            o.y = run; run += v[4 * q + 1];
            o.z = run; run += v[4 * q + 2];
            o.w = run; run += v[4 * q + 3];
            dst[q] = o;
        }
        if (first + SCAN_ITEMS == cells) starts[cells] = run;
    } else {
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            if (first + i <= cells) starts[first + i] = run;
            run += v[i];
        }
    }
}

// two entities per thread: 128-bit position loads, 64-bit key / rank loads.  Measured and rejected (profiles/r1_final.md):
// four entities per thread (+20 %: a warp's stores spread over 128 slots and coalesce less) and a software pipeline that
// issues the next pair's streaming loads before the dependent cell-start gathers (+3 %).
__global__ void __launch_bounds__(256)
cell_scatter_kernel(uint32_t n_host, const uint32_t* __restrict__ n_dev, const uint2* __restrict__ keys, const uint2* __restrict__ rank,
                    const float4* __restrict__ pos, const uint32_t* __restrict__ starts, float2* __restrict__ sorted_pos, uint32_t* __restrict__ sorted_idx) {
    const uint32_t n = n_dev ? *n_dev : n_host;
    const uint32_t pairs = (n + 1u) >> 1;
    for (uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x; pi < pairs; pi += gridDim.x * blockDim.x) {
        const uint2 k = __ldcs(keys + pi), r = __ldcs(rank + pi);
        const float4 p = __ldcs(pos + pi);
        const uint32_t e0 = pi * 2u, e1 = e0 + 1u;
        if (r.x != CSORT_SKIP) {
            const uint32_t s0 = __ldg(starts + k.x) + r.x;
            sorted_pos[s0] = make_float2(p.x, p.y);
            sorted_idx[s0] = e0;
        }
        if (e1 < n && r.y != CSORT_SKIP) {
            const uint32_t s1 = __ldg(starts + k.y) + r.y;
            sorted_pos[s1] = make_float2(p.z, p.w);
            sorted_idx[s1] = e1;
        }
    }
}


// ---- slot-returning scatter (single-GPU default) ------------------------------------------------------------------
// The move pass only COUNTED the cells (one RED per run, move.cu); here the run heads take their slots from the scanned
// table itself: base = atomicAdd(&cursor[key], run length), where cursor[c] starts out as start(c).  When the kernel is
// done cursor[c] = start(c) + count(c) = start(c + 1), so the same table, read one word earlier (the word in front of it is
// a permanent 0), is the prefix table the query needs: tab[c] = start(c), tab[ncells] = n.
// Per entity: position R8 (the key is recomputed: 6 instructions instead of 8 bytes), sorted position W8, slot W4
// (coalesced, entity order: the flag readback and the periodic re-sort find an entity's slot there; nothing scattered but
// the positions themselves).  One entity per lane per load: a warp's 32 consecutive entities form a few runs whose slots are
// consecutive, so the 8-byte stores of a run coalesce into whole sectors; four independent chunks per warp iteration keep
// four atomic round trips in flight per lane.
constexpr int SCATTER_CHUNKS = 4;

struct RunSlot {
    uint32_t base;     // head lanes: first slot of the run
    uint32_t my_head;  // lane of the head of this lane's run
};
__device__ __forceinline__ RunSlot run_slot_issue(uint32_t* __restrict__ cursor, uint32_t key, bool valid, uint32_t lane) {
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane == 0 || key != prev;
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const uint32_t valid_mask = __ballot_sync(0xffffffffu, valid);
    RunSlot r;
    r.my_head = 31u - __clz(heads & ((2u << lane) - 1u));
    r.base = 0;
    if (head && valid) {  // valid lanes precede invalid ones (tail of the array), so a run with a valid member has a valid head
        const uint32_t above = lane == 31u ? 0u : (heads >> (lane + 1u)) << (lane + 1u);
        const uint32_t next_head = above ? static_cast<uint32_t>(__ffs(above) - 1) : 32u;
        const uint32_t run = (next_head == 32u ? 0xffffffffu : ((1u << next_head) - 1u)) & ~((1u << lane) - 1u);
        r.base = atomicAdd(&cursor[key], static_cast<uint32_t>(__popc(run & valid_mask)));
    }
    return r;
}

__global__ void __launch_bounds__(256)
cell_scatter_slots_kernel(uint32_t n, const float2* __restrict__ pos, uint32_t* __restrict__ cursor, float2* __restrict__ sorted_pos,
                          uint32_t* __restrict__ slot_of_entity, GridParams grid) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps_total = (gridDim.x * blockDim.x) >> 5;
    constexpr uint32_t PER_WARP = 32u * SCATTER_CHUNKS;
    for (uint32_t base = warp_global * PER_WARP; base < n; base += warps_total * PER_WARP) {
        float2 p[SCATTER_CHUNKS];
        RunSlot r[SCATTER_CHUNKS];
#pragma unroll
        for (int k = 0; k < SCATTER_CHUNKS; k++) {
            const uint32_t e = base + k * 32u + lane;
            p[k] = e < n ? __ldcs(pos + e) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < SCATTER_CHUNKS; k++) {
            const uint32_t e = base + k * 32u + lane;
            r[k] = run_slot_issue(cursor, cell_key_of(p[k], grid), e < n, lane);
        }
#pragma unroll
        for (int k = 0; k < SCATTER_CHUNKS; k++) {
            const uint32_t e = base + k * 32u + lane;
            const uint32_t slot = __shfl_sync(0xffffffffu, r[k].base, r[k].my_head) + (lane - r[k].my_head);
            if (e < n) {
                sorted_pos[slot] = p[k];
                __stcs(slot_of_entity + e, slot);
            }
        }
    }
}

// count pass for a collision dispatch that was not preceded by a counting move pass (first dispatch after an upload, grid change)
__global__ void __launch_bounds__(256)
cell_count_pos_kernel(uint32_t n, const float2* __restrict__ pos, uint32_t* __restrict__ cell_count, GridParams grid) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
        atomicAdd(&cell_count[cell_key_of(__ldcs(pos + e), grid)], 1u);
}

// sorted slot -> entity, from the entity -> slot map the scatter wrote (only the periodic re-sort wants this direction)
__global__ void __launch_bounds__(256) invert_slots_kernel(uint32_t n, const uint32_t* __restrict__ slot_of_entity, uint32_t* __restrict__ sorted_idx) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) sorted_idx[__ldcs(slot_of_entity + e)] = e;
}

// collision flag per entity = flag of its sorted slot (+1: 1 = green, 2 = blue, 0 = "no collision pass yet")
__global__ void __launch_bounds__(256)
gather_flags_kernel(uint32_t n, const uint32_t* __restrict__ slot_of_entity, const uint8_t* __restrict__ flag_sorted, uint8_t* __restrict__ flag_entity) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) flag_entity[e] = flag_sorted[__ldcs(slot_of_entity + e)] + 1;
}

}  // namespace

uint32_t csort_tiles(uint32_t cells) { return (cells + 1u + SCAN_TILE - 1) / SCAN_TILE; }


void csort_clear(cudaStream_t s, uint32_t* cell_count, uint32_t cells, Profiler* prof) {
    prof->begin(s, K_MEMSET);
    cudaMemsetAsync(cell_count, 0, (static_cast<size_t>(cells) + 1) * sizeof(uint32_t), s);
    prof->end(s);
}

// A sharded handle only ever holds keys of its band's rows plus one ghost row on either side: the counter
// table is cleared and scanned over that cell range [c0, c1) only (c0 rounded down to the scan kernels'
// 16-byte vector alignment), so the per-tick table work shrinks with the band instead of staying
// grid-sized on every GPU.  cell_start is valid on [c0, c1]; nothing outside is read (collide.cu).
void csort_band(uint32_t cells, int ncx, uint32_t row_lo, uint32_t row_hi, int ncy, uint32_t* c0, uint32_t* c1) {
    const uint32_t r0 = row_lo > 0 ? row_lo - 1 : 0;
    const uint32_t r1 = row_hi + 1 < static_cast<uint32_t>(ncy) ? row_hi + 1 : static_cast<uint32_t>(ncy);
    *c0 = (r0 * static_cast<uint32_t>(ncx)) & ~3u;
    *c1 = r1 * static_cast<uint32_t>(ncx);
    if (*c1 > cells) *c1 = cells;
}

int launch_cell_count(cudaStream_t s, uint32_t n, const uint32_t* keys, uint32_t* cell_count, uint32_t* rank, uint32_t c0, uint32_t c1, Profiler* prof,
                      const uint32_t* n_dev) {
    if (n == 0) return 0;
    uint32_t blocks = (n + 255u) / 256u;
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    prof->begin(s, K_CELL_COUNT);
    cell_count_kernel<<<blocks, 256, 0, s>>>(n, n_dev, keys, cell_count, rank, c0, c1);
    prof->end(s);
    return 1;
}

int launch_cell_scan(cudaStream_t s, uint32_t* cell_count, uint32_t cells, uint32_t* tile_sums, uint32_t* cell_start, Profiler* prof) {
    const uint32_t tiles = csort_tiles(cells);
    prof->begin(s, K_CELL_SCAN);
    scan_tile_sums_kernel<<<tiles, SCAN_THREADS, 0, s>>>(cell_count, cells, tile_sums);
    scan_tiles_kernel<0><<<tiles, SCAN_THREADS, 0, s>>>(cell_count, cells, tile_sums, cell_start);
    prof->end(s);
    return 2;
}

int launch_cell_scatter(cudaStream_t s, uint32_t n, const uint32_t* keys, const uint32_t* rank, const float2* pos, const uint32_t* cell_start,
                        float2* sorted_pos, uint32_t* sorted_idx, Profiler* prof, const uint32_t* n_dev) {
    if (n == 0) return 0;
    const uint32_t pairs = (n + 1u) >> 1;
    uint32_t blocks = (pairs + 255u) / 256u;
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    prof->begin(s, K_CELL_SCATTER);
    cell_scatter_kernel<<<blocks, 256, 0, s>>>(n, n_dev, reinterpret_cast<const uint2*>(keys), reinterpret_cast<const uint2*>(rank),
                                               reinterpret_cast<const float4*>(pos), cell_start, sorted_pos, sorted_idx);
    prof->end(s);
    return 1;
}

}  // namespace msim

namespace msim {

int launch_cell_scatter_slots(cudaStream_t s, int sm_count, uint32_t n, const float2* pos, uint32_t* cursor, float2* sorted_pos, uint32_t* slot_of_entity,
                              const GridParams& grid, Profiler* prof) {
    if (n == 0) return 0;
    uint32_t blocks = (n + 256u * SCATTER_CHUNKS - 1u) / (256u * SCATTER_CHUNKS);
    const uint32_t resident = static_cast<uint32_t>(sm_count) * 8u;
    if (blocks > resident) blocks = resident;
    prof->begin(s, K_CELL_SCATTER);
    cell_scatter_slots_kernel<<<blocks, 256, 0, s>>>(n, pos, cursor, sorted_pos, slot_of_entity, grid);
    prof->end(s);
    return 1;
}

int launch_cell_count_pos(cudaStream_t s, int sm_count, uint32_t n, const float2* pos, uint32_t* cell_count, const GridParams& grid, Profiler* prof) {
    if (n == 0) return 0;
    uint32_t blocks = (n + 255u) / 256u;
    const uint32_t resident = static_cast<uint32_t>(sm_count) * 8u;
    if (blocks > resident) blocks = resident;
    prof->begin(s, K_CELL_COUNT);
    cell_count_pos_kernel<<<blocks, 256, 0, s>>>(n, pos, cell_count, grid);
    prof->end(s);
    return 1;
}

int launch_invert_slots(cudaStream_t s, uint32_t n, const uint32_t* slot_of_entity, uint32_t* sorted_idx, Profiler* prof) {
    if (n == 0) return 0;
    prof->begin(s, K_MISC);
    invert_slots_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(n, slot_of_entity, sorted_idx);
    prof->end(s);
    return 1;
}

int launch_gather_flags(cudaStream_t s, uint32_t n, const uint32_t* slot_of_entity, const uint8_t* flag_sorted, uint8_t* flag_entity, Profiler* prof) {
    if (n == 0) return 0;
    prof->begin(s, K_SCATTER_FLAGS);
    gather_flags_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(n, slot_of_entity, flag_sorted, flag_entity);
    prof->end(s);
    return 1;
}

}  // namespace msim
