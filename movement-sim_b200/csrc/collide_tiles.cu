// collide_tiles.cu — the odd-tick dispatch on one GPU (default path): who has a neighbour within the collision radius,
// and how many unordered pairs are in range.
//
// Observable contract (reference shader /root/reference/src/sim/shader/random_move.comp):
//   :875-877  every entity turns green, then  :545-547  both members of every in-range pair turn blue;
//   :551-562  in_range(a,b,r): sqrt(dx*dx+dy*dy) < r (strict), evaluated here as d2 < T with the exact binary32 bound T
//             (api.cu exact_hit_threshold), individually rounded operations, no FMA.
//
// Input: positions in cell order (sorted_pos) and the prefix table tab[c] = first slot of cell c (csort.cu).  One thread
// per sorted slot j; every unordered pair is examined from its HIGHER slot: thread j tests the slots below it — the three
// cells (cx-1..cx+1) of the grid row above, one contiguous run [ab_lo, ab_hi), and its own row from cell cx-1 up to j,
// the run [own_lo, j).
//
// What this kernel is built around (ncu of its predecessors, profiles/r2_query_history.md): with ~15 tests per entity the
// query is not bound by the distance tests but by everything around them — the round-1 kernel spent 613 warp instructions
// per 32 entities on 488 useful tests, its main loops ran at 13-15 of 32 lanes, and a third of its stall samples sat at
// CTA-wide barriers.  So:
//   * candidates are read in ALIGNED GROUPS OF FOUR slots (two 128-bit loads), without per-candidate range checks.  Slots
//     just outside a run belong to cells at least two columns away (their distance test fails by itself) as long as the
//     above run and the own run are at least three slots apart; warps where some lane's runs are closer (tiny or nearly
//     empty maps) take an exact scalar path instead.  Only the group that contains slot j itself is masked (k < j);
//   * a hit costs 1.5 instructions: one packed subtract (d2 - T) per two candidates and one LEA.HI that adds the sign bit;
//   * no shared memory and no barrier: lanes of one cell read the same addresses, the 32 slots of a warp and the run above
//     them are a few hundred contiguous bytes that stay in L1 (75 % hit rate); staging them per CTA or per warp with TMA
//     bulk copies was measured slower (mbarrier set-up, hull exchange and waits cost more issue slots than the loads save:
//     191 / 212 us against 156 us), so every warp is independent from its first instruction;
//   * the pair total leaves as one reduction per warp (RED, nothing waits) into striped counters; the flagged total is NOT taken here:
//     fold_counts_kernel (collide.cu) counts the flag bytes this kernel stored and folds the pair stripes (profiles/r2_flag_count_race.md);
//   * lanes that find nothing below look above warp-cooperatively (collide_common.cuh look_above_cooperative): what is left of
//     divergence are the count loops, whose lanes differ in trip count only (one code path, one value per uniform register).
#include <cstdlib>

#include "msim_internal.h"

namespace msim {
namespace {

#include "collide_common.cuh"

// ---- the inner loop: four candidates per step ------------------------------------------------------------------------------
// A group is 32 bytes {x0 y0 x1 y1}{x2 y2 x3 y3}: two LDG.128 through the read-only path.  Per candidate: FADD2 (dx, dy), FMUL2 (dx^2, dy^2), FADD (d2) —
// individually rounded, as the oracle computes them — then per PAIR of candidates one FADD2 (d2 - T) and per candidate one
// LEA.HI that adds the sign bit of the difference to the count: d2 >= +0 and T > 0 are never NaN (positions are finite), so
// the difference is negative exactly when d2 < T.
#ifdef MSIM_HOST_EMU
__device__ __forceinline__ uint32_t below(float d2, float thr) { return d2 < thr ? 1u : 0u; }
__device__ __forceinline__ uint32_t hits_in_pair(float4 a, float2 p, float thr) {
    return below(dist2(make_float2(a.x, a.y), p), thr) + below(dist2(make_float2(a.z, a.w), p), thr);
}
#else
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint32_t below(float d2, float thr) { return __float_as_uint(__fsub_rn(d2, thr)) >> 31; }
__device__ __forceinline__ uint32_t hits_in_pair(float4 a, float2 p, float thr) {
    const unsigned long long pp = pack2(p.x, p.y), tt = pack2(thr, thr);
    unsigned long long d0, d1, t;
    float x0, y0, x1, y1, t0, t1;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d0) : "l"(pack2(a.x, a.y)), "l"(pp));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d1) : "l"(pack2(a.z, a.w)), "l"(pp));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(d0) : "l"(d0));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(d1) : "l"(d1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(y0) : "l"(d0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x1), "=f"(y1) : "l"(d1));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pack2(__fadd_rn(x0, y0), __fadd_rn(x1, y1))), "l"(tt));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
    return (__float_as_uint(t0) >> 31) + (__float_as_uint(t1) >> 31);
}
#endif

struct Group { float4 a, b; };
__device__ __forceinline__ Group load_group(const float4* g) { return Group{__ldg(g), __ldg(g + 1)}; }

// groups [g0, g1) of sorted_pos, no masks
__device__ __forceinline__ uint32_t count_groups(const float4* __restrict__ base, uint32_t g0, uint32_t g1, float2 p, float thr) {
    uint32_t c = 0;
#pragma unroll 1
    for (uint32_t g = g0; g < g1; g++) {
        const Group q = load_group(base + 2u * g);
        c += hits_in_pair(q.a, p, thr) + hits_in_pair(q.b, p, thr);
    }
    return c;
}

// the group that holds slot j itself: only its first `below_j` (0..3) candidates count
__device__ __forceinline__ uint32_t count_last_group(const float4* __restrict__ g, uint32_t below_j, float2 p, float thr) {
    const Group q = load_group(g);
    const uint32_t c0 = below(dist2(make_float2(q.a.x, q.a.y), p), thr), c1 = below(dist2(make_float2(q.a.z, q.a.w), p), thr),
                   c2 = below(dist2(make_float2(q.b.x, q.b.y), p), thr);
    return (below_j > 0u ? c0 : 0u) + (below_j > 1u ? c1 : 0u) + (below_j > 2u ? c2 : 0u);
}

// true iff some candidate of groups [g0, g1) is in range (colours-only mode: the first hit ends the scan)
__device__ __forceinline__ bool any_in_groups(const float4* __restrict__ base, uint32_t g0, uint32_t g1, float2 p, float thr) {
#pragma unroll 1
    for (uint32_t g = g0; g < g1; g++) {
        const Group q = load_group(base + 2u * g);
        if (hits_in_pair(q.a, p, thr) + hits_in_pair(q.b, p, thr)) return true;
    }
    return false;
}

constexpr int TILES_THREADS = 128;  // measured: 64 / 128 / 256 threads per CTA = 170.7 / 169.4 / 173.6 us in the tick

// COUNT_PAIRS = false (MSIM_FLAG_NO_PAIR_COUNT): colours only, a thread stops at its first neighbour.
// GHOSTS (sharded handles): the sorted order also holds the neighbours' boundary rows and this tick's leavers, whose cell rows lie
// outside the band [row_lo, row_hi): candidates for everybody else, but they get no flag and count no pairs here - their owner does
// that.  Every pair is counted by the GPU that owns its higher slot, so the all-reduced sum equals the single-GPU count.  The
// number of sorted slots is the scan's grand total, read from the table itself (n_dev); `n` then only sizes the grid.
template <bool COUNT_PAIRS, bool GHOSTS>
__global__ void __launch_bounds__(TILES_THREADS)
query_tiles_kernel(uint32_t n_host, const uint32_t* __restrict__ n_dev, const float2* __restrict__ sorted_pos, const uint32_t* __restrict__ tab,
                   uint8_t* __restrict__ flag_sorted, GridParams grid, unsigned long long* __restrict__ stripes, int row_lo, int row_hi) {
    const uint32_t n = GHOSTS && n_dev ? *n_dev : n_host;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp_base = (blockIdx.x * TILES_THREADS + threadIdx.x) & ~31u;
    if (warp_base >= n) return;
    const uint32_t j_raw = warp_base + lane;
    const bool in_range = j_raw < n;
    const uint32_t j = in_range ? j_raw : n - 1u;  // the lanes behind the last slot repeat it (and keep their results to themselves)

    // ---- this lane's runs: three loads from the prefix table ---------------------------------------------------------------
    const float2 p = __ldg(sorted_pos + j);
    int cx = __float2int_rd(__fmul_rn(p.x, grid.inv_cell));
    int cy = __float2int_rd(__fmul_rn(p.y, grid.inv_cell));
    cx = min(max(cx, 0), grid.ncx - 1);
    cy = min(max(cy, 0), grid.ncy - 1);
    const bool mine = in_range && (!GHOSTS || (cy >= row_lo && cy < row_hi));
    if (GHOSTS && !__any_sync(0xffffffffu, mine)) {  // a warp of ghosts (the halo rows): nothing to find out here
        if (in_range) flag_sorted[j] = 0;
        return;
    }
    const uint32_t ncx = static_cast<uint32_t>(grid.ncx);
    const uint32_t r = static_cast<uint32_t>(cy) * ncx;                               // < 2^27 cells: 32-bit indices throughout
    const uint32_t x0 = static_cast<uint32_t>(max(cx - 1, 0)), x1e = min(static_cast<uint32_t>(cx) + 2u, ncx);
    uint32_t own_lo = __ldg(tab + (r + x0));
    uint32_t ab_lo = 0u, ab_hi = 0u;
    if (cy > 0) {
        ab_lo = __ldg(tab + (r - ncx + x0));
        ab_hi = __ldg(tab + (r - ncx + x1e));
    }

    uint32_t pairs = 0;
    const float thr = grid.hit_threshold;
    uint32_t scan_hi = j;  // own run [own_lo, scan_hi)
    if (GHOSTS && !mine) {  // a ghost among owned slots: empty runs, nothing counted
        ab_lo = ab_hi = 0u;
        own_lo = scan_hi = j & ~3u;
    }
    // aligned groups are only safe when the slots around a run are far-away cells: above run and own run >= 3 slots apart
    const bool close_runs = ab_hi != ab_lo && own_lo - ab_hi < 3u;
    const float4* G = reinterpret_cast<const float4*>(sorted_pos);
    const uint32_t ga0 = ab_lo >> 2, ga1 = ab_hi != ab_lo ? (ab_hi + 3u) >> 2 : ga0;
    bool hit;
    if (__any_sync(0xffffffffu, close_runs)) {  // exact scalar scans (tiny or nearly empty maps)
        if (COUNT_PAIRS) {
            pairs = count_in_range(sorted_pos, ab_lo, ab_hi, p, thr) + count_in_range(sorted_pos, own_lo, scan_hi, p, thr);
            hit = pairs != 0u;
        } else {
            hit = any_in_range(sorted_pos, own_lo, scan_hi, p, thr) || any_in_range(sorted_pos, ab_lo, ab_hi, p, thr);
        }
    } else if (COUNT_PAIRS) {
        pairs = count_groups(G, ga0, ga1, p, thr) + count_groups(G, own_lo >> 2, scan_hi >> 2, p, thr) +
                count_last_group(G + 2u * (scan_hi >> 2), scan_hi & 3u, p, thr);
        hit = pairs != 0u;
    } else {
        hit = count_last_group(G + 2u * (scan_hi >> 2), scan_hi & 3u, p, thr) != 0u || any_in_groups(G, own_lo >> 2, scan_hi >> 2, p, thr) ||
              any_in_groups(G, ga0, ga1, p, thr);
    }
    // nothing below: look at the slots above (rest of the own row up to the end of cell cx+1, then the three cells of the row below), where the
    // first hit is enough.  Few lanes get here (an entity with neighbours has, as a rule, some of them below it), so the warp serves them one at
    // a time, TOGETHER: the lane's position and its two runs are broadcast, every lane tests another candidate of the concatenated runs.
    // Control flow stays warp-uniform on purpose.  The first version let each such lane walk its runs alone inside a divergent region, and
    // ptxas kept different kernel parameters (row count, threshold, array base) in ONE uniform register across that region's paths: lanes of a
    // warp on different paths overwrote it for each other and, about once per 10^10 entity-ticks, a lane compared its distance with the wrong
    // number and missed its only neighbour (profiles/r2_flag_count_race.md).
    {
        const bool need = !hit && mine;
        uint32_t up0 = 0u, up1 = 0u, dn0 = 0u, dn1 = 0u;
        if (need) {
            up0 = j + 1u;
            up1 = __ldg(tab + (r + x1e));
            if (cy + 1 < grid.ncy) {
                dn0 = __ldg(tab + (r + ncx + x0));
                dn1 = __ldg(tab + (r + ncx + x1e));
            }
        }
        look_above_cooperative(need, p, up0, up1, dn0, dn1, sorted_pos, thr, lane, hit);
    }
    if (in_range) flag_sorted[j] = hit ? 1 : 0;

    // ---- totals: the pairs leave as one reduction (no return value, nothing waits) per warp into a striped counter.  A lane that took the
    // divergent look-above path found nothing below: it carries no pairs, so this sum does not depend on where the warp reconverges.
    // The flagged entities are counted from the flags just stored, by fold_counts_kernel (collide.cu) behind this kernel.
    if (COUNT_PAIRS) {
        pairs = __reduce_add_sync(0xffffffffu, mine ? pairs : 0u);
        if (lane == 0 && pairs) atomicAdd(stripes + ((warp_base >> 5) % COUNTER_STRIPES) * COUNTER_STRIDE + 1, static_cast<unsigned long long>(pairs));
    }
}

}  // namespace

int launch_query_tiles(cudaStream_t s, uint32_t n, const float2* sorted_pos, const uint32_t* tab, uint8_t* flag_sorted, const GridParams& grid, bool count_pairs,
                       Counters* counters, unsigned long long* stripes, Profiler* prof, const uint32_t* n_dev, bool ghosts, uint32_t row_lo, uint32_t row_hi) {
    if (n == 0) return 0;
    const uint32_t blocks = (n + TILES_THREADS - 1) / TILES_THREADS;
    const int lo = static_cast<int>(row_lo), hi = static_cast<int>(row_hi);
    prof->begin(s, K_QUERY);
    if (ghosts) {
        if (count_pairs) query_tiles_kernel<true, true><<<blocks, TILES_THREADS, 0, s>>>(n, n_dev, sorted_pos, tab, flag_sorted, grid, stripes, lo, hi);
        else query_tiles_kernel<false, true><<<blocks, TILES_THREADS, 0, s>>>(n, n_dev, sorted_pos, tab, flag_sorted, grid, stripes, lo, hi);
    } else {
        if (count_pairs) query_tiles_kernel<true, false><<<blocks, TILES_THREADS, 0, s>>>(n, nullptr, sorted_pos, tab, flag_sorted, grid, stripes, 0, 0);
        else query_tiles_kernel<false, false><<<blocks, TILES_THREADS, 0, s>>>(n, nullptr, sorted_pos, tab, flag_sorted, grid, stripes, 0, 0);
    }
    prof->end(s);
    return 1 + launch_fold_counts(s, n, ghosts ? n_dev : nullptr, flag_sorted, stripes, counters, prof);
}

}  // namespace msim
