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
    if (tuning().scan_min_blocks == 8)
        scan_tiles_kernel<8><<<tiles, SCAN_THREADS, 0, s>>>(cell_count, cells, tile_sums, cell_start);
    else
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
