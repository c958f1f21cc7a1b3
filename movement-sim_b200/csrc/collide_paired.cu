// collide_paired.cu - opt-in variant of the collision query (MSIM_QUERY_PAIRED=1): two adjacent slots per thread.  A translation unit of its
// own so that adding it leaves the machine code of the default query kernels untouched (profiles/sass_identity.py).
// NOT YET RUN ON HARDWARE (written after the round's GPU budget was spent): tests/test_zz_gpu_unverified.py.
#include "msim_internal.h"

namespace msim {
namespace {

#include "collide_common.cuh"

// ---- opt-in variant (MSIM_QUERY_PAIRED=1; unsharded handles, counting-sort directory, pair count on): TWO adjacent slots per thread ------
// Adjacent slots of the cell order nearly always share a cell or sit in neighbouring cells of one row, so their candidate runs nearly
// coincide: both entities are tested against every candidate of the union of their runs with ONE shared-memory load per candidate, half the
// loop overhead and half the per-thread set-up (cell look-ups, hull, reductions) per entity.  A candidate of the union that lies outside an
// entity's own 3 x 2 cells is at least one cell edge (> radius) away in x, so it can never be a hit: no range test per entity is needed, the
// counts stay the exact unique-pair counts.  Slot pairs that straddle a row boundary or an empty stretch of cells take the one-entity loops.
constexpr int PAIRED_THREADS = 128;  // 256 slots per CTA as in query_kernel: same windows, same shared memory

__device__ __forceinline__ void count_two_in_tile(const float2* __restrict__ tile, uint32_t a, uint32_t b, float2 p0, float2 p1, float threshold,
                                                  uint32_t& c0, uint32_t& c1) {
    uint32_t k = a;
    for (; k + 4 <= b; k += 4) {
        const float2 q0 = tile[k], q1 = tile[k + 1], q2 = tile[k + 2], q3 = tile[k + 3];
        c0 += (dist2(q0, p0) < threshold) ? 1u : 0u;
        c1 += (dist2(q0, p1) < threshold) ? 1u : 0u;
        c0 += (dist2(q1, p0) < threshold) ? 1u : 0u;
        c1 += (dist2(q1, p1) < threshold) ? 1u : 0u;
        c0 += (dist2(q2, p0) < threshold) ? 1u : 0u;
        c1 += (dist2(q2, p1) < threshold) ? 1u : 0u;
        c0 += (dist2(q3, p0) < threshold) ? 1u : 0u;
        c1 += (dist2(q3, p1) < threshold) ? 1u : 0u;
    }
    for (; k < b; k++) {
        const float2 q = tile[k];
        c0 += (dist2(q, p0) < threshold) ? 1u : 0u;
        c1 += (dist2(q, p1) < threshold) ? 1u : 0u;
    }
}

struct SlotRuns {
    int cx, cy, x0, x1;
    uint32_t own_lo, own_hi, ab_lo, ab_hi;  // ab_lo == ab_hi when there is no row above
};
__device__ __forceinline__ SlotRuns slot_runs(float2 p, const uint32_t* __restrict__ cell_start, const GridParams& grid) {
    SlotRuns r;
    r.cx = min(max(__float2int_rd(__fmul_rn(p.x, grid.inv_cell)), 0), grid.ncx - 1);
    r.cy = min(max(__float2int_rd(__fmul_rn(p.y, grid.inv_cell)), 0), grid.ncy - 1);
    r.x0 = max(r.cx - 1, 0);
    r.x1 = min(r.cx + 1, grid.ncx - 1);
    row_run<true>(nullptr, cell_start, grid.ncx, r.cy, r.x0, r.x1, r.own_lo, r.own_hi);
    if (r.cy > 0) row_run<true>(nullptr, cell_start, grid.ncx, r.cy - 1, r.x0, r.x1, r.ab_lo, r.ab_hi);
    else r.ab_lo = r.ab_hi = 0;
    return r;
}
// what is left for an entity that found nobody below it: the rest of its row and the row below, first hit wins (rare)
__device__ __forceinline__ bool look_above(const float2* __restrict__ sorted_pos, const uint32_t* __restrict__ cell_start, const GridParams& grid,
                                           const SlotRuns& r, uint32_t j, float2 p) {
    bool hit = any_in_range(sorted_pos, max(j + 1, r.own_lo), r.own_hi, p, grid.hit_threshold);
    if (!hit && r.cy + 1 < grid.ncy) {
        uint32_t lo, hi;
        row_run<true>(nullptr, cell_start, grid.ncx, r.cy + 1, r.x0, r.x1, lo, hi);
        hit = any_in_range(sorted_pos, lo, hi, p, grid.hit_threshold);
    }
    return hit;
}

__global__ void __launch_bounds__(PAIRED_THREADS)
query_paired_kernel(uint32_t n, const float2* __restrict__ sorted_pos, const uint32_t* __restrict__ cell_start, uint8_t* __restrict__ flag_sorted,
                    GridParams grid, unsigned long long* __restrict__ stripes) {
    __shared__ __align__(16) float2 s_above[QUERY_WINDOW];
    __shared__ __align__(16) float2 s_own[QUERY_WINDOW];
    __shared__ uint32_t s_red[3][PAIRED_THREADS / 32];
    __shared__ __align__(8) unsigned long long s_bar;

    const uint32_t block_base = blockIdx.x * (2u * PAIRED_THREADS);
    if (block_base >= n) return;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    const uint32_t j0 = block_base + 2u * threadIdx.x, j1 = j0 + 1u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool v0 = j0 < n, v1 = j1 < n;
    const uint32_t last = min(block_base + 2u * PAIRED_THREADS, n) - 1u;  // the CTA's last slot

    float2 p0 = make_float2(0.f, 0.f), p1 = make_float2(0.f, 0.f);
    SlotRuns r0{}, r1{};
    if (v0) {
        const float4 pp = *reinterpret_cast<const float4*>(sorted_pos + j0);  // j0 is even and sorted_pos holds a multiple of 64 slots
        p0 = make_float2(pp.x, pp.y);
        p1 = make_float2(pp.z, pp.w);
        r0 = slot_runs(p0, cell_start, grid);
        if (v1) r1 = slot_runs(p1, cell_start, grid);
    }
    // hull of the CTA's windows from its first and last slot (the bounds are non-decreasing in the slot index, see query_kernel)
    if (threadIdx.x == 0) {
        s_red[0][0] = r0.own_lo;
        s_red[1][0] = r0.cy > 0 ? r0.ab_lo : 0u;
    }
    if (v0 && (j0 == last || j1 == last)) {
        const SlotRuns& rl = (j1 == last) ? r1 : r0;
        s_red[2][0] = rl.cy > 0 ? rl.ab_hi : 0u;
    }
    __syncthreads();
    uint32_t w_own_lo = s_red[0][0], w_ab_lo = s_red[1][0], w_ab_hi = s_red[2][0];
    const uint32_t w_own_hi = last + 1u;
    if (w_ab_lo > w_ab_hi) w_ab_lo = w_ab_hi = 0;
    w_own_lo &= ~1u;
    w_ab_lo &= ~1u;
    const uint32_t own_slots = (w_own_hi - w_own_lo + 1u) & ~1u;
    const uint32_t ab_slots = (w_ab_hi - w_ab_lo + 1u) & ~1u;
    const bool tiled = own_slots <= QUERY_WINDOW && ab_slots <= QUERY_WINDOW;  // CTA-uniform
    if (tiled) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&s_bar, (own_slots + ab_slots) * static_cast<uint32_t>(sizeof(float2)));
            bulk_copy_g2s(s_own, sorted_pos + w_own_lo, own_slots * static_cast<uint32_t>(sizeof(float2)), &s_bar);
            if (ab_slots) bulk_copy_g2s(s_above, sorted_pos + w_ab_lo, ab_slots * static_cast<uint32_t>(sizeof(float2)), &s_bar);
        }
        mbar_wait(&s_bar, 0);
    }

    const float thr = grid.hit_threshold;
    uint32_t c0 = 0, c1 = 0;
    if (v0) {
        const bool together = tiled && v1 && r0.cy == r1.cy && static_cast<uint32_t>(r1.cx - r0.cx) <= 1u;
        if (together) {
            // union of the two entities' runs: [first bound of slot 0, last bound of slot 1) above, [first bound of slot 0, slot 0) in the own row
            if (r0.ab_lo < r1.ab_hi) count_two_in_tile(s_above, r0.ab_lo - w_ab_lo, r1.ab_hi - w_ab_lo, p0, p1, thr, c0, c1);
            count_two_in_tile(s_own, r0.own_lo - w_own_lo, j0 - w_own_lo, p0, p1, thr, c0, c1);
            c1 += (dist2(p0, p1) < thr) ? 1u : 0u;  // slot 0 is below slot 1
        } else if (tiled) {
            if (r0.ab_lo < r0.ab_hi) c0 += count_in_tile(s_above, r0.ab_lo - w_ab_lo, r0.ab_hi - w_ab_lo, p0, thr);
            c0 += count_in_tile(s_own, r0.own_lo - w_own_lo, j0 - w_own_lo, p0, thr);
            if (v1) {
                if (r1.ab_lo < r1.ab_hi) c1 += count_in_tile(s_above, r1.ab_lo - w_ab_lo, r1.ab_hi - w_ab_lo, p1, thr);
                c1 += count_in_tile(s_own, r1.own_lo - w_own_lo, j1 - w_own_lo, p1, thr);
            }
        } else {
            if (r0.ab_lo < r0.ab_hi) c0 += count_in_range(sorted_pos, r0.ab_lo, r0.ab_hi, p0, thr);
            c0 += count_in_range(sorted_pos, r0.own_lo, min(j0, r0.own_hi), p0, thr);
            if (v1) {
                if (r1.ab_lo < r1.ab_hi) c1 += count_in_range(sorted_pos, r1.ab_lo, r1.ab_hi, p1, thr);
                c1 += count_in_range(sorted_pos, r1.own_lo, min(j1, r1.own_hi), p1, thr);
            }
        }
    }
    bool hit0 = c0 != 0, hit1 = c1 != 0;
    if (v0 && !hit0) hit0 = look_above(sorted_pos, cell_start, grid, r0, j0, p0);
    if (v1 && !hit1) hit1 = look_above(sorted_pos, cell_start, grid, r1, j1, p1);
    if (v1) *reinterpret_cast<uchar2*>(flag_sorted + j0) = make_uchar2(hit0 ? 1 : 0, hit1 ? 1 : 0);
    else if (v0) flag_sorted[j0] = hit0 ? 1 : 0;

    uint32_t hits = __popc(__ballot_sync(0xffffffffu, hit0)) + __popc(__ballot_sync(0xffffffffu, hit1));
    uint32_t pairs = __reduce_add_sync(0xffffffffu, c0 + c1);
    __syncthreads();  // s_red is reused
    if (lane == 0) {
        s_red[0][warp] = hits;
        s_red[1][warp] = pairs;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t h = 0, pr = 0;
#pragma unroll
        for (int w = 0; w < PAIRED_THREADS / 32; w++) {
            h += s_red[0][w];
            pr += s_red[1][w];
        }
        unsigned long long* stripe = stripes + static_cast<size_t>(blockIdx.x % COUNTER_STRIPES) * COUNTER_STRIDE;
        if (h) atomicAdd(stripe, static_cast<unsigned long long>(h));
        if (pr) atomicAdd(stripe + 1, static_cast<unsigned long long>(pr));
    }
}


}  // namespace

void launch_query_paired(cudaStream_t s, uint32_t n, const float2* sorted_pos, const uint32_t* cell_start, uint8_t* flag_sorted, const GridParams& grid,
                         unsigned long long* stripes) {
    query_paired_kernel<<<(n + 2 * PAIRED_THREADS - 1) / (2 * PAIRED_THREADS), PAIRED_THREADS, 0, s>>>(n, sorted_pos, cell_start, flag_sorted, grid, stripes);
}

}  // namespace msim
