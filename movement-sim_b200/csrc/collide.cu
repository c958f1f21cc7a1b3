// collide.cu — the odd-tick dispatch: who has a neighbour within the collision radius?
//
// Observable contract taken from the reference shader (/root/reference/src/sim/shader/random_move.comp):
//   :875-877  every entity turns green (0,1,0,1), then
//   :545-547  both members of every in-range pair turn blue (0,0,1,1);
//   :551-562  in_range(a,b,r): |dx|<=r, |dy|<=r and sqrt(dx*dx+dy*dy) < r  (strict).
// The reference finds the pairs by walking a lock-protected quadtree (:564-719).  Here the pairs are
// found on a uniform cell grid (edge slightly above r) built from the radix-sorted cell keys:
//   build_cells : gathers positions into cell order and records, per cell, {first, ~end} of its run
//   query       : one thread per sorted entity scans the 3 cell rows around it; the three cells of a
//                 row are adjacent keys, hence ONE contiguous run of the sorted position array
// The colour is kept as a 1-byte flag per sorted slot and expanded to RGBA at readback (pack.cu).
// The predicate is evaluated without sqrt: hit_threshold is the exact binary32 bound T with
// (d2 < T) <=> (sqrtf(d2) < r), computed on the host (api.cu), so flags match the oracle bit for bit.
#include "msim_internal.h"

namespace msim {
namespace {

__global__ void __launch_bounds__(256)
build_cells_kernel(uint32_t n, const unsigned long long* __restrict__ sorted, const float2* __restrict__ pos,
                   float2* __restrict__ sorted_pos, uint2* __restrict__ cell_range, Counters* __restrict__ counters) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) {  // per-pass counters of the query that follows
        counters->pairs_last = 0;
        counters->flagged_last = 0;
    }
    const uint32_t lane = threadIdx.x & 31u;
    const bool in = j < n;
    unsigned long long kv = in ? __ldcs(sorted + j) : ~0ull;
    const uint32_t key = static_cast<uint32_t>(kv >> 32);
    uint32_t prev_key = __shfl_up_sync(0xffffffffu, key, 1);
    if (!in) return;
    if (lane == 0) prev_key = (j == 0) ? 0xffffffffu : static_cast<uint32_t>(sorted[j - 1] >> 32);
    const uint32_t idx = static_cast<uint32_t>(kv);
    sorted_pos[j] = pos[idx];  // 8-byte gather, coalesced store
    if (j == 0 || key != prev_key) {
        cell_range[key].x = j;                       // first slot of this cell
        if (j != 0) cell_range[prev_key].y = ~j;     // one past the last slot of the previous cell
    }
    if (j == n - 1) cell_range[key].y = ~n;
}

// squared distance with individually rounded operations, as the oracle computes it
__device__ __forceinline__ float dist2(float2 a, float2 b) {
    const float dx = __fsub_rn(a.x, b.x);
    const float dy = __fsub_rn(a.y, b.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// number of slots k in [a, b) with dist2(sorted_pos[k], p) < threshold; four loads in flight
__device__ __forceinline__ uint32_t count_in_range(const float2* __restrict__ sorted_pos, uint32_t a, uint32_t b, float2 p, float threshold) {
    uint32_t c = 0;
    uint32_t k = a;
    for (; k + 4 <= b; k += 4) {
        const float2 q0 = __ldg(sorted_pos + k), q1 = __ldg(sorted_pos + k + 1), q2 = __ldg(sorted_pos + k + 2), q3 = __ldg(sorted_pos + k + 3);
        c += (dist2(q0, p) < threshold) ? 1u : 0u;
        c += (dist2(q1, p) < threshold) ? 1u : 0u;
        c += (dist2(q2, p) < threshold) ? 1u : 0u;
        c += (dist2(q3, p) < threshold) ? 1u : 0u;
    }
    for (; k < b; k++) c += (dist2(__ldg(sorted_pos + k), p) < threshold) ? 1u : 0u;
    return c;
}

// true iff some slot k in [a, b) is within range; four independent loads per step, so the scan costs
// one memory latency per four candidates instead of one per candidate
__device__ __forceinline__ bool any_in_range(const float2* __restrict__ sorted_pos, uint32_t a, uint32_t b, float2 p, float threshold) {
    uint32_t k = a;
    for (; k + 4 <= b; k += 4) {
        const float2 q0 = __ldg(sorted_pos + k), q1 = __ldg(sorted_pos + k + 1), q2 = __ldg(sorted_pos + k + 2), q3 = __ldg(sorted_pos + k + 3);
        const bool h = (dist2(q0, p) < threshold) | (dist2(q1, p) < threshold) | (dist2(q2, p) < threshold) | (dist2(q3, p) < threshold);
        if (h) return true;
    }
    for (; k < b; k++)
        if (dist2(__ldg(sorted_pos + k), p) < threshold) return true;
    return false;
}

// the three cells x0..x1 of one grid row are adjacent keys = one contiguous run [lo, hi) of the sorted order
__device__ __forceinline__ void row_run(const uint2* __restrict__ cell_range, int ncx, int yy, int x0, int x1, uint32_t& lo, uint32_t& hi) {
    const uint2* row = cell_range + static_cast<size_t>(yy) * ncx;
    lo = 0xffffffffu;
    hi = 0u;
    for (int xx = x0; xx <= x1; xx++) {
        const uint2 r = __ldg(row + xx);  // empty cell = {0xffffffff, ~0xffffffff = 0}: neutral for min/max
        lo = min(lo, r.x);
        hi = max(hi, ~r.y);
    }
    if (lo > hi) lo = hi;  // all three empty
}

// One thread per sorted slot j.  Every unordered pair is examined from its HIGHER slot only: thread j
// counts its in-range partners among the slots below it (the whole row above, and its own row up to
// j) — that is the exact unique-pair count and already decides the flag for almost everybody.  Only
// when nothing was found below does it look at the slots above (rest of its row, row below), and
// there the first hit is enough.  Half the distance tests of a full 3x3 scan, same flags, same count.
// GHOSTS: slots whose entity index is >= n_owned belong to a neighbouring GPU (halo): they are
// candidates for everybody else but get no flag and count no pairs here — their owner does that.
template <bool COUNT_PAIRS, bool GHOSTS>
__global__ void __launch_bounds__(256)
query_kernel(uint32_t n, uint32_t n_owned, const unsigned long long* __restrict__ sorted, const float2* __restrict__ sorted_pos,
             const uint2* __restrict__ cell_range, uint8_t* __restrict__ flag_sorted, GridParams grid, Counters* __restrict__ counters) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t pairs = 0;
    bool hit = false;
    bool mine = j < n;
    if (GHOSTS && mine) {
        mine = static_cast<uint32_t>(__ldcs(sorted + j)) < n_owned;
        if (!mine) flag_sorted[j] = 0;
    }
    if (mine) {
        const float2 p = sorted_pos[j];
        const float thr = grid.hit_threshold;
        int cx = __float2int_rd(__fmul_rn(p.x, grid.inv_cell));
        int cy = __float2int_rd(__fmul_rn(p.y, grid.inv_cell));
        cx = min(max(cx, 0), grid.ncx - 1);
        cy = min(max(cy, 0), grid.ncy - 1);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, grid.ncx - 1);
        uint32_t lo, hi;
        // own row: slots below j, then (flags-only or nothing found) slots above j
        uint32_t own_lo, own_hi;
        row_run(cell_range, grid.ncx, cy, x0, x1, own_lo, own_hi);
        if (COUNT_PAIRS) {
            if (cy > 0) {
                row_run(cell_range, grid.ncx, cy - 1, x0, x1, lo, hi);
                pairs += count_in_range(sorted_pos, lo, hi, p, thr);
            }
            pairs += count_in_range(sorted_pos, own_lo, min(j, own_hi), p, thr);
            hit = pairs != 0;
        } else {
            hit = any_in_range(sorted_pos, own_lo, min(j, own_hi), p, thr);
            if (!hit && cy > 0) {
                row_run(cell_range, grid.ncx, cy - 1, x0, x1, lo, hi);
                hit = any_in_range(sorted_pos, lo, hi, p, thr);
            }
        }
        if (!hit) hit = any_in_range(sorted_pos, max(j + 1, own_lo), own_hi, p, thr);
        if (!hit && cy + 1 < grid.ncy) {
            row_run(cell_range, grid.ncx, cy + 1, x0, x1, lo, hi);
            hit = any_in_range(sorted_pos, lo, hi, p, thr);
        }
        flag_sorted[j] = hit ? 1 : 0;
    }
    // per-warp reduction, one atomic per warp per counter
    const uint32_t hits = __popc(__ballot_sync(0xffffffffu, hit));
    if (COUNT_PAIRS) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) pairs += __shfl_down_sync(0xffffffffu, pairs, d);
    }
    if (lane == 0) {
        if (hits) atomicAdd(&counters->flagged_last, static_cast<unsigned long long>(hits));
        if (COUNT_PAIRS && pairs) {
            atomicAdd(&counters->pairs_last, static_cast<unsigned long long>(pairs));
            atomicAdd(&counters->pairs_total, static_cast<unsigned long long>(pairs));
        }
    }
}

__global__ void __launch_bounds__(256)
scatter_flags_kernel(uint32_t n, uint32_t n_owned, const unsigned long long* __restrict__ sorted, const uint8_t* __restrict__ flag_sorted,
                     uint8_t* __restrict__ flag_entity) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t idx = static_cast<uint32_t>(__ldcs(sorted + j));
    if (idx >= n_owned) return;  // ghost
    flag_entity[idx] = flag_sorted[j] + 1;  // 1 = green, 2 = blue (0 = "no collision pass yet")
}

}  // namespace

int launch_build_cells(cudaStream_t s, uint32_t n, const uint64_t* sorted, const float2* pos, float2* sorted_pos, uint2* cell_range,
                       const GridParams& grid, Counters* counters, Profiler* prof) {
    prof->begin(s, K_MEMSET);
    cudaMemsetAsync(cell_range, 0xff, static_cast<size_t>(grid.ncells) * sizeof(uint2), s);
    prof->end(s);
    if (n == 0) return 0;
    prof->begin(s, K_BUILD_CELLS);
    build_cells_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(n, reinterpret_cast<const unsigned long long*>(sorted), pos, sorted_pos, cell_range, counters);
    prof->end(s);
    return 1;
}

int launch_query(cudaStream_t s, uint32_t n, uint32_t n_owned, const uint64_t* sorted, const float2* sorted_pos, const uint2* cell_range,
                 uint8_t* flag_sorted, const GridParams& grid, bool count_pairs, Counters* counters, Profiler* prof) {
    if (n == 0) return 0;
    const uint32_t blocks = (n + 255u) / 256u;
    const unsigned long long* so = reinterpret_cast<const unsigned long long*>(sorted);
    prof->begin(s, K_QUERY);
    if (n_owned < n) {
        if (count_pairs) query_kernel<true, true><<<blocks, 256, 0, s>>>(n, n_owned, so, sorted_pos, cell_range, flag_sorted, grid, counters);
        else query_kernel<false, true><<<blocks, 256, 0, s>>>(n, n_owned, so, sorted_pos, cell_range, flag_sorted, grid, counters);
    } else {
        if (count_pairs) query_kernel<true, false><<<blocks, 256, 0, s>>>(n, n_owned, so, sorted_pos, cell_range, flag_sorted, grid, counters);
        else query_kernel<false, false><<<blocks, 256, 0, s>>>(n, n_owned, so, sorted_pos, cell_range, flag_sorted, grid, counters);
    }
    prof->end(s);
    return 1;
}

int launch_scatter_flags(cudaStream_t s, uint32_t n, uint32_t n_owned, const uint64_t* sorted, const uint8_t* flag_sorted, uint8_t* flag_entity, Profiler* prof) {
    if (n == 0) return 0;
    prof->begin(s, K_SCATTER_FLAGS);
    scatter_flags_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(n, n_owned, reinterpret_cast<const unsigned long long*>(sorted), flag_sorted, flag_entity);
    prof->end(s);
    return 1;
}

}  // namespace msim
