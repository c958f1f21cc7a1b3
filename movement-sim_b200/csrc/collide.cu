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
build_cells_kernel(uint32_t n_host, const uint32_t* __restrict__ n_dev, const unsigned long long* __restrict__ sorted, const float2* __restrict__ pos,
                   float2* __restrict__ sorted_pos, uint32_t* __restrict__ sorted_idx, uint2* __restrict__ cell_range, Counters* __restrict__ counters) {
    const uint32_t n = n_dev ? *n_dev : n_host;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    const bool in = j < n;
    unsigned long long kv = in ? __ldcs(sorted + j) : ~0ull;
    const uint32_t key = static_cast<uint32_t>(kv >> 32);
    uint32_t prev_key = __shfl_up_sync(0xffffffffu, key, 1);
    if (!in) return;
    if (lane == 0) prev_key = (j == 0) ? 0xffffffffu : static_cast<uint32_t>(sorted[j - 1] >> 32);
    const uint32_t idx = static_cast<uint32_t>(kv);
    sorted_pos[j] = pos[idx];  // 8-byte gather, coalesced store
    sorted_idx[j] = idx;
    if (j == 0 || key != prev_key) {
        cell_range[key].x = j;                       // first slot of this cell
        if (j != 0) cell_range[prev_key].y = ~j;     // one past the last slot of the previous cell
    }
    if (j == n - 1) cell_range[key].y = ~n;
}

#include "collide_common.cuh"

// One thread per sorted slot j.  Every unordered pair is examined from its HIGHER slot only: thread j
// counts its in-range partners among the slots below it — the whole grid row above (cells cx-1..cx+1)
// and its own row up to j.  That is the exact unique-pair count and already decides the flag for
// almost everybody; only when nothing was found below does the thread look at the slots above it
// (rest of its row, row below), where the first hit is enough.
//
// The 256 slots of a CTA are consecutive in cell order, so the candidates they need form two short
// contiguous windows of sorted_pos (one in the row above, one ending at the CTA's own last slot).
// Both windows are staged in shared memory with coalesced loads and every distance test reads LDS:
// the first version issued one L2 round trip per four candidates per thread and was 78 % stalled on
// the long scoreboard (profiles/r1b).  CTAs whose windows do not fit (a row boundary inside the CTA,
// or a pile-up of thousands of entities in one cell) fall back to the global-memory scan.
//
// GHOSTS: slots whose entity index is >= n_owned belong to a neighbouring GPU (halo): they are
// candidates for everybody else but get no flag and count no pairs here — their owner does that.
template <bool COUNT_PAIRS, bool GHOSTS>
__global__ void __launch_bounds__(QUERY_THREADS)
query_kernel(uint32_t n_host, uint32_t n_owned_host, const uint32_t* __restrict__ n_dev, const uint32_t* __restrict__ n_owned_dev,
             const uint32_t* __restrict__ sorted_idx, const float2* __restrict__ sorted_pos,
             const uint2* __restrict__ cell_range, uint8_t* __restrict__ flag_sorted, GridParams grid,
             unsigned long long* __restrict__ stripes) {
    __shared__ __align__(16) float2 s_above[QUERY_WINDOW];
    __shared__ __align__(16) float2 s_own[QUERY_WINDOW];
    __shared__ uint32_t s_red[3][QUERY_THREADS / 32];
    __shared__ __align__(8) unsigned long long s_bar;

    const uint32_t n = n_dev ? *n_dev : n_host;
    const uint32_t n_owned = n_owned_dev ? *n_owned_dev : n_owned_host;
    const uint32_t block_base = blockIdx.x * QUERY_THREADS;
    if (block_base >= n) return;  // grid sized for an upper bound of n (whole CTA, uniform)
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);  // made visible to the CTA by the barrier behind the hull computation
    const uint32_t j = block_base + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t pairs = 0;
    bool hit = false;
    bool mine = j < n;
    if (GHOSTS && mine) {
        mine = __ldcs(sorted_idx + j) < n_owned;
        if (!mine) flag_sorted[j] = 0;
    }

    float2 p = make_float2(0.f, 0.f);
    int cx = 0, cy = 0, x0 = 0, x1 = 0;
    uint32_t own_lo = 0xffffffffu, own_hi = 0, ab_lo = 0xffffffffu, ab_hi = 0;
    if (mine) {
        p = sorted_pos[j];
        cx = __float2int_rd(__fmul_rn(p.x, grid.inv_cell));
        cy = __float2int_rd(__fmul_rn(p.y, grid.inv_cell));
        cx = min(max(cx, 0), grid.ncx - 1);
        cy = min(max(cy, 0), grid.ncy - 1);
        x0 = max(cx - 1, 0);
        x1 = min(cx + 1, grid.ncx - 1);
        row_run(cell_range, grid.ncx, cy, x0, x1, own_lo, own_hi);
        if (cy > 0) row_run(cell_range, grid.ncx, cy - 1, x0, x1, ab_lo, ab_hi);
        else ab_lo = ab_hi = 0;
    }

    // CTA-wide hull of the windows: min of the run starts, max of the run ends
    const bool has_above = mine && ab_lo < ab_hi;
    uint32_t w_own_lo = 0xffffffffu, w_ab_lo = 0xffffffffu, w_ab_hi = 0;
    {
        uint32_t r0 = __reduce_min_sync(0xffffffffu, mine ? own_lo : 0xffffffffu);
        uint32_t r1 = __reduce_min_sync(0xffffffffu, has_above ? ab_lo : 0xffffffffu);
        uint32_t r2 = __reduce_max_sync(0xffffffffu, has_above ? ab_hi : 0u);
        if (lane == 0) {
            s_red[0][warp] = r0;
            s_red[1][warp] = r1;
            s_red[2][warp] = r2;
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < QUERY_THREADS / 32; w++) {
            w_own_lo = min(w_own_lo, s_red[0][w]);
            w_ab_lo = min(w_ab_lo, s_red[1][w]);
            w_ab_hi = max(w_ab_hi, s_red[2][w]);
        }
    }
    const uint32_t w_own_hi = min(block_base + QUERY_THREADS, n);  // nobody needs a slot at or above its own
    if (w_ab_lo > w_ab_hi) w_ab_lo = w_ab_hi = 0;
    const bool any_mine = w_own_lo != 0xffffffffu;
    // windows widened to even slot indices: 16-byte aligned source and size for the bulk copy (sorted_pos holds a multiple of
    // 64 slots, so the widened end stays inside the array; the extra slots are never used as candidates)
    w_own_lo &= ~1u;
    w_ab_lo &= ~1u;
    const uint32_t own_slots = any_mine ? ((w_own_hi - w_own_lo + 1u) & ~1u) : 0u;
    const uint32_t ab_slots = (w_ab_hi - w_ab_lo + 1u) & ~1u;
    const bool tiled = any_mine && own_slots <= QUERY_WINDOW && ab_slots <= QUERY_WINDOW;  // CTA-uniform

    if (tiled) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&s_bar, (own_slots + ab_slots) * static_cast<uint32_t>(sizeof(float2)));
            bulk_copy_g2s(s_own, sorted_pos + w_own_lo, own_slots * static_cast<uint32_t>(sizeof(float2)), &s_bar);
            if (ab_slots) bulk_copy_g2s(s_above, sorted_pos + w_ab_lo, ab_slots * static_cast<uint32_t>(sizeof(float2)), &s_bar);
        }
        mbar_wait(&s_bar, 0);
    }

    if (mine) {
        const float thr = grid.hit_threshold;
        if (tiled) {
            if (COUNT_PAIRS) {
                if (has_above) pairs += count_in_tile(s_above, ab_lo - w_ab_lo, ab_hi - w_ab_lo, p, thr);
                pairs += count_in_tile(s_own, own_lo - w_own_lo, j - w_own_lo, p, thr);
                hit = pairs != 0;
            } else {
                hit = any_in_tile(s_own, own_lo - w_own_lo, j - w_own_lo, p, thr);
                if (!hit && has_above) hit = any_in_tile(s_above, ab_lo - w_ab_lo, ab_hi - w_ab_lo, p, thr);
            }
        } else {
            if (COUNT_PAIRS) {
                if (has_above) pairs += count_in_range(sorted_pos, ab_lo, ab_hi, p, thr);
                pairs += count_in_range(sorted_pos, own_lo, min(j, own_hi), p, thr);
                hit = pairs != 0;
            } else {
                hit = any_in_range(sorted_pos, own_lo, min(j, own_hi), p, thr);
                if (!hit && has_above) hit = any_in_range(sorted_pos, ab_lo, ab_hi, p, thr);
            }
        }
    }
    // nothing below: look above, the whole warp together (collide_common.cuh; every lane of every warp gets here)
    {
        const bool need = mine && !hit;
        uint32_t up0 = 0u, up1 = 0u, dn0 = 0u, dn1 = 0u;
        if (need) {
            up0 = max(j + 1u, own_lo);
            up1 = max(own_hi, up0);
            if (cy + 1 < grid.ncy) {
                row_run(cell_range, grid.ncx, cy + 1, x0, x1, dn0, dn1);
                if (dn0 > dn1) dn0 = dn1;
            }
        }
        look_above_cooperative(need, p, up0, up1, dn0, dn1, sorted_pos, grid.hit_threshold, lane, hit);
    }
    if (mine) flag_sorted[j] = hit ? 1 : 0;
    // CTA-level reduction, then ONE atomic per CTA per counter into a striped counter (64 stripes on
    // separate 128-byte lines).  The first version issued three same-address atomics per warp: ~940 k
    // atomics on one L2 line serialised to ~630 us and hid everything else (profiles/r1d).
    if (COUNT_PAIRS) {  // (flagged entities are counted from the stored flags by fold_counts_kernel)
        pairs = __reduce_add_sync(0xffffffffu, pairs);
        __syncthreads();  // s_red is reused
        if (lane == 0) s_red[1][warp] = pairs;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t pr = 0;
#pragma unroll
            for (int w = 0; w < QUERY_THREADS / 32; w++) pr += s_red[1][w];
            unsigned long long* stripe = stripes + static_cast<size_t>(blockIdx.x % COUNTER_STRIPES) * COUNTER_STRIDE;
            if (pr) atomicAdd(stripe + 1, static_cast<unsigned long long>(pr));
        }
    }
}

// ---- totals of one query -----------------------------------------------------------------------------------------------------
// Pairs: the query kernels reduce them per warp / CTA into the striped counters (a lane that takes the divergent "look above" path
// has found nothing below, so it carries no pairs).  Flagged entities are NOT counted by the query kernels: they are counted here,
// from the flags the query wrote (one byte per sorted slot, 0 / 1; ghost slots hold 0).  Round 2 found the in-kernel count - a
// ballot behind the divergent look-above scans - one short of the flags it had just stored on ~10 % of the ticks at 10 M entities
// (profiles/r2_flag_count_race.md); the flags themselves were almost always right.  This kernel uses no warp-level primitive at all:
// per-thread sums meet in shared-memory atomics behind a CTA barrier and CTAs meet in ONE global atomic each, which carries the CTA's
// count and its ticket - the CTA that completes the set finds the total in the value it gets back (no fence, no second round trip:
// the first version's ticket chain and last-CTA fold were 8 of its 12 us).  CTA 0 folds the pair stripes while its flags are in flight.
constexpr int FOLD_THREADS = 512;
constexpr int FOLD_UNROLL = 8;  // 148 CTAs x 512 threads x 8 x 16 B = 9.7 MB per sweep: one sweep at 10 M entities
constexpr int FOLD_TICKET_SHIFT = 40;  // word behind the stripes: {CTAs done : 24 | flagged so far : 40}
__global__ void __launch_bounds__(FOLD_THREADS)
fold_counts_kernel(uint32_t n_host, const uint32_t* __restrict__ n_dev, const uint8_t* __restrict__ flag_sorted, unsigned long long* __restrict__ stripes,
                   Counters* __restrict__ counters) {
    __shared__ unsigned int s_flagged;
    __shared__ unsigned long long s_pairs;
    if (threadIdx.x == 0) {
        s_flagged = 0u;
        s_pairs = 0ull;
    }
    __syncthreads();
    const uint32_t n = n_dev ? *n_dev : n_host;
    // CTA 0 folds the pair stripes (final since the query kernel ended) while it waits for its flags
    unsigned long long pr = 0ull;
    const bool folds = blockIdx.x == 0 && threadIdx.x < COUNTER_STRIPES;
    unsigned long long* const stripe = stripes + static_cast<size_t>(threadIdx.x % COUNTER_STRIPES) * COUNTER_STRIDE;
    if (folds) pr = stripe[1];
    const uint32_t chunks = n >> 4;  // 16 flags per load; every byte is 0 or 1, so a word's flags are its set bits
    const uint4* f4 = reinterpret_cast<const uint4*>(flag_sorted);
    uint32_t sum = 0;
    const uint32_t stride = gridDim.x * FOLD_THREADS;
    for (uint32_t i = blockIdx.x * FOLD_THREADS + threadIdx.x; i < chunks; i += FOLD_UNROLL * stride) {
        uint4 v[FOLD_UNROLL];
#pragma unroll
        for (int q = 0; q < FOLD_UNROLL; q++) v[q] = i + q * stride < chunks ? __ldcs(f4 + i + q * stride) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int q = 0; q < FOLD_UNROLL; q++) sum += __popc(v[q].x) + __popc(v[q].y) + __popc(v[q].z) + __popc(v[q].w);
    }
    if (blockIdx.x == 0)
        for (uint32_t k = (chunks << 4) + threadIdx.x; k < n; k += FOLD_THREADS) sum += flag_sorted[k];
    if (sum) atomicAdd(&s_flagged, sum);
    if (folds) {
        stripe[0] = 0;
        stripe[1] = 0;
        if (pr) atomicAdd(&s_pairs, pr);
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        counters->pairs_last = s_pairs;
        counters->pairs_total += s_pairs;
    }
    // one atomic per CTA carries its count and its ticket; the CTA that completes the set has the total in the value it got back
    unsigned long long* const acc = stripes + static_cast<size_t>(COUNTER_STRIPES) * COUNTER_STRIDE;
    const unsigned long long old = atomicAdd(acc, (1ull << FOLD_TICKET_SHIFT) | static_cast<unsigned long long>(s_flagged));
    if ((old >> FOLD_TICKET_SHIFT) != gridDim.x - 1u) return;
    const unsigned long long h = (old & ((1ull << FOLD_TICKET_SHIFT) - 1ull)) + s_flagged;
    *acc = 0ull;  // (the next launch is behind this one on the stream)
    counters->flagged_last = h;
    counters->flagged_total += h;
}

__global__ void __launch_bounds__(256)
scatter_flags_kernel(uint32_t n, uint32_t n_owned, const uint32_t* __restrict__ sorted_idx, const uint8_t* __restrict__ flag_sorted,
                     uint8_t* __restrict__ flag_entity) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t idx = __ldcs(sorted_idx + j);
    if (idx >= n_owned) return;  // ghost
    flag_entity[idx] = flag_sorted[j] + 1;  // 1 = green, 2 = blue (0 = "no collision pass yet")
}

}  // namespace

int launch_build_cells(cudaStream_t s, uint32_t n, const uint64_t* sorted, const float2* pos, float2* sorted_pos, uint32_t* sorted_idx,
                       uint2* cell_range, const GridParams& grid, Counters* counters, Profiler* prof, const uint32_t* n_dev) {
    prof->begin(s, K_MEMSET);
    cudaMemsetAsync(cell_range, 0xff, static_cast<size_t>(grid.ncells) * sizeof(uint2), s);
    prof->end(s);
    if (n == 0) return 0;
    prof->begin(s, K_BUILD_CELLS);
    build_cells_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(n, n_dev, reinterpret_cast<const unsigned long long*>(sorted), pos, sorted_pos, sorted_idx, cell_range, counters);
    prof->end(s);
    return 1;
}

int launch_query(cudaStream_t s, uint32_t n, uint32_t n_owned, const uint32_t* sorted_idx, const float2* sorted_pos, const uint2* cell_range,
                 uint8_t* flag_sorted, const GridParams& grid, bool count_pairs, Counters* counters, unsigned long long* stripes, Profiler* prof,
                 const uint32_t* n_dev, const uint32_t* n_owned_dev) {
    if (n == 0) return 0;
    const uint32_t blocks = (n + QUERY_THREADS - 1) / QUERY_THREADS;
    const bool ghosts = n_owned < n || n_owned_dev != nullptr;
    prof->begin(s, K_QUERY);
#define MSIM_QUERY(CP, GH) \
    query_kernel<CP, GH><<<blocks, QUERY_THREADS, 0, s>>>(n, n_owned, n_dev, n_owned_dev, sorted_idx, sorted_pos, cell_range, flag_sorted, grid, stripes)
    if (count_pairs) {
        if (ghosts) MSIM_QUERY(true, true); else MSIM_QUERY(true, false);
    } else {
        if (ghosts) MSIM_QUERY(false, true); else MSIM_QUERY(false, false);
    }
#undef MSIM_QUERY
    prof->end(s);
    return 1 + launch_fold_counts(s, n, n_dev, flag_sorted, stripes, counters, prof);
}

int launch_fold_counts(cudaStream_t s, uint32_t n, const uint32_t* n_dev, const uint8_t* flag_sorted, unsigned long long* stripes, Counters* counters, Profiler* prof) {
    uint32_t blocks = (n / 16u + FOLD_THREADS * FOLD_UNROLL - 1) / (FOLD_THREADS * FOLD_UNROLL);  // n is an upper bound when n_dev is given
    if (blocks > 148u) blocks = 148u;
    if (blocks == 0) blocks = 1;
    prof->begin(s, K_FOLD);
    fold_counts_kernel<<<blocks, FOLD_THREADS, 0, s>>>(n, n_dev, flag_sorted, stripes, counters);
    prof->end(s);
    return 1;
}

size_t query_stripe_bytes() { return static_cast<size_t>(COUNTER_STRIPES) * COUNTER_STRIDE * sizeof(unsigned long long); }

int launch_scatter_flags(cudaStream_t s, uint32_t n, uint32_t n_owned, const uint32_t* sorted_idx, const uint8_t* flag_sorted, uint8_t* flag_entity, Profiler* prof) {
    if (n == 0) return 0;
    prof->begin(s, K_SCATTER_FLAGS);
    scatter_flags_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(n, n_owned, sorted_idx, flag_sorted, flag_entity);
    prof->end(s);
    return 1;
}

}  // namespace msim
