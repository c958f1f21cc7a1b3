// csort.cu — single-digit radix sort of the cell keys (a counting sort over the whole key): the per-tick rebuild of the
// neighbour structure for populations whose storage order is spatially coherent; the per-cell counter table must fit the L2
// (Munich: 4.77 M cells x 4 B = 19 MB of B200's 126 MB).
//
// An LSD radix sort needs one pass per digit; with a table as large as the key space the digit IS the key and one pass suffices:
//   count    cell_count[key] += 1 per entity, as one RED per run of equal keys in adjacent lanes - fused into the move kernel
//            (move.cu), or cell_count_kernel below when the dispatch was not preceded by a counting move pass
//   scan     cell_start = exclusive prefix sum of cell_count (one kernel, one pass over the table)
//   scatter  slot = atomicAdd(&cell_start[key], run length): sorted_pos[slot] = pos[e], slot_of_entity[e] = slot
// When the scatter is done cell_start[c] = start(c) + count(c) = start(c + 1): the same table read one word earlier (the scan
// leaves a 0 in front of its first word) is the prefix table the query walks, tab[c] = first slot of cell c, and the cell
// directory: the run of cells x0..x1 of a row is [tab[row*ncx+x0], tab[row*ncx+x1+1]).
// Per entity: position R8 (+ key R4 on sharded handles; otherwise the key is recomputed: 6 instructions instead of 8 bytes),
// sorted position W8, slot W4 (coalesced, entity order: the flag readback and the periodic re-sort find an entity's slot there).
// The order of entities inside one cell is the arrival order of the atomics (not deterministic); nothing observable depends
// on it: flags and the unique-pair count are order-independent.
// Measured (profiles/r1_sort_paths.md): with entities in random index order the scattered stores cost 4-5 times as much, so
// the storage is kept in cell order (periodic re-sort, pack.cu); upload-ordered storage (MSIM_FLAG_NO_REORDER) takes the
// staged onesweep radix sort (sort.cu) unless MSIM_FLAG_SORT_COUNTING asks for this path.
#include <cstdlib>

#include "msim_internal.h"

namespace msim {
namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096 cells per CTA

// count pass for keys that did not come from a counting move pass (sharded handles after a host-side exchange)
__global__ void __launch_bounds__(256)
cell_count_kernel(uint32_t n_host, const uint32_t* __restrict__ n_dev, const uint32_t* __restrict__ keys, uint32_t* __restrict__ cell_count, uint32_t c0,
                  uint32_t c1) {
    const uint32_t n = n_dev ? *n_dev : n_host;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint32_t k = __ldcs(keys + e);
        // keys outside the counted cell range [c0, c1) (sharded handles only) stay out of the order
        if (k - c0 < c1 - c0) atomicAdd(&cell_count[k], 1u);
    }
}

// ---- single-pass scan ---------------------------------------------------------------------------------------------------------
// One kernel, one pass over the table: 19 MB in, 19 MB out, counters zeroed on the way.  A tile (4096 cells) publishes its total
// as soon as it has read its cells, and needs the sum of the totals before it.  Two things were measured and discarded first:
//   * the classic look-back (wait for a predecessor's inclusive prefix) degenerates when the whole table is one resident wave -
//     every tile walks all the way back (32-40 us);
//   * every tile summing ALL earlier totals (also what the round-1 two-kernel scan did) is 680 k relaxed loads on the same 73 L2
//     lines: the hot lines serialise and the kernel takes 25-37 us however its own accesses are laid out.
// So the totals are summed in two levels: tiles form groups of SCAN_GROUP; the last tile of a group to arrive (found with one
// 64-bit atomic that carries {arrivals, sum}) publishes the group's sum, and a tile reads the totals before it in its own group
// plus the sums of the groups before its own: at most 31 + T / 32 words instead of T.  Nothing waits for anything but totals.
// A status word is {epoch : 32 | value : 32}: words of an earlier launch (other epoch) read as "not there yet", so the arrays are
// never cleared.  Tiles take their index from a ticket, so a tile only ever waits for tiles that are already running.
#ifdef MSIM_HOST_EMU
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
#else
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

constexpr int SCAN_STEPS = SCAN_ITEMS / 4;                 // uint4 loads per thread
constexpr int SCAN_WARP_CELLS = 32 * SCAN_ITEMS;           // 512 consecutive cells per warp

struct ScanState {
    unsigned long long* tile_total;   // [tiles]   {epoch, total of the tile}
    unsigned long long* group_total;  // [groups]  {epoch, sum of the group's tiles}
    unsigned long long* group_acc;    // [groups]  {arrivals, running sum}; left at 0 by the last arrival
    uint32_t* ticket;
    uint32_t group;                   // tiles per group
};

__global__ void __launch_bounds__(SCAN_THREADS)
scan_cells_kernel(uint32_t* __restrict__ counts, uint32_t cells, uint32_t* __restrict__ starts, ScanState st, uint32_t epoch, uint32_t* __restrict__ error_flag) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    __shared__ uint32_t s_tile;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(st.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t warp_first = tile * SCAN_TILE + warp * SCAN_WARP_CELLS;
    uint4 c[SCAN_STEPS];
    // the counters are consumed here; they are zeroed when the tile writes its results (below), so the next tick's count needs no memset.
    // NOT right behind the load: a store to the address a load has just been issued for waits for that load in the LSU and holds up
    // the loads behind it - measured 36 us with the zeroing here against 20 us with it at the end
#pragma unroll
    for (int q = 0; q < SCAN_STEPS; q++) {
        const uint32_t first = warp_first + q * 128u + lane * 4u;  // multiple of 4: 16-byte aligned (the base is)
        if (first + 4u <= cells) {
            uint4* src = reinterpret_cast<uint4*>(counts + first);
            c[q] = __ldcs(src);
        } else {
            uint32_t t[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (first + i < cells) {
                    t[i] = counts[first + i];
                    counts[first + i] = 0u;  // (tail of the table: a handful of cells)
                }
            }
            c[q] = make_uint4(t[0], t[1], t[2], t[3]);
        }
    }
    // exclusive offset of this lane's four cells inside its step, and the step totals
    uint32_t excl[SCAN_STEPS], step_base[SCAN_STEPS];
    uint32_t warp_total = 0;
#pragma unroll
    for (int q = 0; q < SCAN_STEPS; q++) {
        const uint32_t mine = c[q].x + c[q].y + c[q].z + c[q].w;
        uint32_t incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= static_cast<uint32_t>(d)) incl += up;
        }
        excl[q] = incl - mine;
        step_base[q] = warp_total;
        warp_total += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) s_warp[warp] = warp_total;
    __syncthreads();
    uint32_t warp_before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        if (static_cast<uint32_t>(w) < warp) warp_before += s_warp[w];
        total += s_warp[w];
    }
    // publish the tile's total (and, as the group's last arrival, the group's sum); sum what lies before this tile
    const unsigned long long tag = static_cast<unsigned long long>(epoch) << 32;
    const uint32_t g = tile / st.group, in_group = tile - g * st.group;
    if (threadIdx.x == 0) {
        if (tile == 0u) starts[-1] = 0u;  // the word in front of the table: tab[first cell] = 0 for the readers of the shifted table
        st_relaxed_u64(st.tile_total + tile, tag | total);
        const uint32_t members = min(st.group, gridDim.x - g * st.group);
        const unsigned long long old = atomicAdd(st.group_acc + g, (1ull << 32) | total);
        if (static_cast<uint32_t>(old >> 32) + 1u == members) {
            st_relaxed_u64(st.group_total + g, tag | (static_cast<uint32_t>(old) + total));
            st.group_acc[g] = 0ull;  // nobody touches it again in this launch
        }
    }
    uint32_t acc = 0, spins = 0;
    for (uint32_t i = threadIdx.x; i < in_group + g; i += SCAN_THREADS) {
        const unsigned long long* word = i < in_group ? st.tile_total + (g * st.group + i) : st.group_total + (i - in_group);
        unsigned long long w = ld_relaxed_u64(word);
        while ((w >> 32) != epoch) {
            if (++spins > (1u << 22)) {  // never a hang: a tile that does not show up is reported
                atomicExch(error_flag, 1u);
                break;
            }
            w = ld_relaxed_u64(word);
        }
        acc += static_cast<uint32_t>(w);
    }
    acc = __reduce_add_sync(0xffffffffu, acc);
    __syncthreads();  // s_warp has been read by everybody
    if (lane == 0) s_warp[warp] = acc;
    if (threadIdx.x == 0 && tile == gridDim.x - 1u) *st.ticket = 0u;  // every ticket has been handed out: ready for the next launch
    __syncthreads();
    uint32_t before = warp_before;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) before += s_warp[w];
    // starts has cells + 1 entries: entry `cells` receives the grand total
#pragma unroll
    for (int q = 0; q < SCAN_STEPS; q++) {
        const uint32_t first = warp_first + q * 128u + lane * 4u;
        uint4 o;
        o.x = before + step_base[q] + excl[q];
        o.y = o.x + c[q].x;
        o.z = o.y + c[q].y;
        o.w = o.z + c[q].z;
        if (first + 4u <= cells) {
            *reinterpret_cast<uint4*>(starts + first) = o;
            *reinterpret_cast<uint4*>(counts + first) = make_uint4(0u, 0u, 0u, 0u);
            if (first + 4u == cells) starts[cells] = o.w + c[q].w;
        } else {
            const uint32_t t[5] = {o.x, o.y, o.z, o.w, o.w + c[q].w};
#pragma unroll
            for (int i = 0; i < 5; i++)
                if (first + i <= cells && (i < 4 || first + 4u > cells)) starts[first + i] = t[i];
        }
    }
}

// ---- slot-returning scatter -----------------------------------------------------------------------------------------------------
// The count pass only COUNTED the cells; here the run heads take their slots from the scanned table itself:
// base = atomicAdd(&cursor[key], run length), where cursor[c] starts out as start(c).  One entity per lane per load: a warp's 32
// consecutive entities form a few runs whose slots are consecutive, so the 8-byte stores of a run coalesce into whole sectors;
// eight independent chunks per warp iteration keep eight atomic round trips in flight per lane (the kernel is bound by those
// round trips: 79 % long-scoreboard stalls; 2 / 4 / 8 chunks = 85 / 82.6 / 76.3 us at 10 M entities).
// BAND (sharded handles): the element count lives in device memory, the keys are read (the exchange kernels maintain them), and
// keys outside the band's cell range [c0, c1) stay out of the order (slot CSORT_SKIP).
struct RunSlot {
    uint32_t base;     // head lanes: first slot of the run
    uint32_t my_head;  // lane of the head of this lane's run
};
__device__ __forceinline__ RunSlot run_slot_issue(uint32_t* __restrict__ cursor, uint32_t key, bool valid, uint32_t lane) {
    if (!valid) key = 0xffffffffu - lane;  // a key nobody shares: its own run, never issued
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane == 0 || key != prev;
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    RunSlot r;
    r.my_head = 31u - __clz(heads & ((2u << lane) - 1u));
    r.base = 0;
    if (head && valid) {
        const uint32_t above = (heads >> lane) >> 1;  // heads behind this lane, shifted down to bit 0
        const uint32_t len = above ? static_cast<uint32_t>(__ffs(above)) : 32u - lane;  // lanes up to the next head
        r.base = atomicAdd(&cursor[key], len);
    }
    return r;
}

template <int SCATTER_CHUNKS, bool BAND>
__global__ void __launch_bounds__(256)
cell_scatter_slots_kernel(uint32_t n_host, const uint32_t* __restrict__ n_dev, const float2* __restrict__ pos, const uint32_t* __restrict__ keys,
                          uint32_t* __restrict__ cursor, float2* __restrict__ sorted_pos, uint32_t* __restrict__ slot_of_entity, GridParams grid, uint32_t c0,
                          uint32_t c1) {
    const uint32_t n = BAND && n_dev ? *n_dev : n_host;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps_total = (gridDim.x * blockDim.x) >> 5;
    constexpr uint32_t PER_WARP = 32u * SCATTER_CHUNKS;
    for (uint32_t base = warp_global * PER_WARP; base < n; base += warps_total * PER_WARP) {
        float2 p[SCATTER_CHUNKS];
        uint32_t key[SCATTER_CHUNKS];
        RunSlot r[SCATTER_CHUNKS];
#pragma unroll
        for (int k = 0; k < SCATTER_CHUNKS; k++) {
            const uint32_t e = base + k * 32u + lane;
            p[k] = e < n ? __ldcs(pos + e) : make_float2(0.f, 0.f);
            if (BAND) key[k] = e < n ? __ldcs(keys + e) : 0u;
        }
#pragma unroll
        for (int k = 0; k < SCATTER_CHUNKS; k++) {
            const uint32_t e = base + k * 32u + lane;
            if (!BAND) key[k] = cell_key_of(p[k], grid);
            const bool valid = e < n && (!BAND || key[k] - c0 < c1 - c0);
            r[k] = run_slot_issue(cursor, key[k], valid, lane);
            if (!valid) r[k].my_head = 32u;  // marks the lane: nothing to store
        }
#pragma unroll
        for (int k = 0; k < SCATTER_CHUNKS; k++) {
            const uint32_t e = base + k * 32u + lane;
            const bool valid = r[k].my_head != 32u;
            const uint32_t head = valid ? r[k].my_head : lane;
            const uint32_t slot = __shfl_sync(0xffffffffu, r[k].base, head) + (lane - head);
            if (valid) {
                sorted_pos[slot] = p[k];
                __stcs(slot_of_entity + e, slot);
            } else if (BAND && e < n) {
                __stcs(slot_of_entity + e, CSORT_SKIP);
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
__global__ void __launch_bounds__(256)
invert_slots_kernel(uint32_t n_host, const uint32_t* __restrict__ n_dev, const uint32_t* __restrict__ slot_of_entity, uint32_t* __restrict__ sorted_idx) {
    const uint32_t n = n_dev ? *n_dev : n_host;
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint32_t slot = __ldcs(slot_of_entity + e);
    if (slot != CSORT_SKIP) sorted_idx[slot] = e;
}

// collision flag per entity = flag of its sorted slot (+1: 1 = green, 2 = blue, 0 = "no collision pass yet")
__global__ void __launch_bounds__(256)
gather_flags_kernel(uint32_t n, const uint32_t* __restrict__ slot_of_entity, const uint8_t* __restrict__ flag_sorted, uint8_t* __restrict__ flag_entity) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint32_t slot = __ldcs(slot_of_entity + e);
    flag_entity[e] = slot != CSORT_SKIP ? flag_sorted[slot] + 1 : 1;  // out of everybody's reach: green
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

int launch_cell_count(cudaStream_t s, uint32_t n, const uint32_t* keys, uint32_t* cell_count, uint32_t c0, uint32_t c1, Profiler* prof, const uint32_t* n_dev) {
    if (n == 0) return 0;
    uint32_t blocks = (n + 255u) / 256u;
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    prof->begin(s, K_CELL_COUNT);
    cell_count_kernel<<<blocks, 256, 0, s>>>(n, n_dev, keys, cell_count, c0, c1);
    prof->end(s);
    return 1;
}

// scan scratch: zero-initialised device memory of csort_scan_scratch_words(cells) 32-bit words; `epoch` differs from launch to launch, never 0
size_t csort_scan_scratch_words(uint32_t cells) {
    const size_t tiles = csort_tiles(cells);
    return 2 * tiles + 4 * (tiles / 32 + 2) + 4;
}

int launch_cell_scan(cudaStream_t s, uint32_t* cell_count, uint32_t cells, uint32_t* scratch, uint32_t scratch_tiles, uint32_t epoch, uint32_t* cell_start,
                     uint32_t* error_flag, Profiler* prof) {
    const uint32_t tiles = csort_tiles(cells);  // <= scratch_tiles (a band of a sharded handle scans fewer cells than the table has)
    ScanState st;
    st.group = scratch_tiles > 4096u ? 256u : 32u;
    st.tile_total = reinterpret_cast<unsigned long long*>(scratch);
    st.group_total = st.tile_total + scratch_tiles;
    st.group_acc = st.group_total + (scratch_tiles / 32u + 2u);
    st.ticket = reinterpret_cast<uint32_t*>(st.group_acc + (scratch_tiles / 32u + 2u));
    prof->begin(s, K_CELL_SCAN);
    scan_cells_kernel<<<tiles, SCAN_THREADS, 0, s>>>(cell_count, cells, cell_start, st, epoch, error_flag);
    prof->end(s);
    return 1;
}

int launch_cell_scatter_slots(cudaStream_t s, int sm_count, uint32_t n, const float2* pos, const uint32_t* keys, uint32_t* cursor, float2* sorted_pos,
                              uint32_t* slot_of_entity, const GridParams& grid, uint32_t c0, uint32_t c1, Profiler* prof, const uint32_t* n_dev, int ctas_per_sm) {
    if (n == 0) return 0;
    constexpr uint32_t chunks = 8u;
    uint32_t blocks = (n + 256u * chunks - 1u) / (256u * chunks);
    const uint32_t resident = static_cast<uint32_t>(sm_count) * static_cast<uint32_t>(ctas_per_sm > 0 && ctas_per_sm < 8 ? ctas_per_sm : 8);
    if (blocks > resident) blocks = resident;
    prof->begin(s, K_CELL_SCATTER);
    if (keys)
        cell_scatter_slots_kernel<chunks, true><<<blocks, 256, 0, s>>>(n, n_dev, pos, keys, cursor, sorted_pos, slot_of_entity, grid, c0, c1);
    else
        cell_scatter_slots_kernel<chunks, false><<<blocks, 256, 0, s>>>(n, nullptr, pos, nullptr, cursor, sorted_pos, slot_of_entity, grid, 0u, grid.ncells);
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

int launch_invert_slots(cudaStream_t s, uint32_t n, const uint32_t* slot_of_entity, uint32_t* sorted_idx, Profiler* prof, const uint32_t* n_dev) {
    if (n == 0) return 0;
    prof->begin(s, K_MISC);
    invert_slots_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(n, n_dev, slot_of_entity, sorted_idx);
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
