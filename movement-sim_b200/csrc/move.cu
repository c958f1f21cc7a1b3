// move.cu — the even-tick dispatch: advance every entity along the road graph.
//
// Semantics follow the reference shader /root/reference/src/sim/shader/random_move.comp:
//   :725-746  xorshift128 / next_float / next(state,min,max)
//   :778-828  new_target       :830-839  update_direction       :841-852  move
//   :869-873  main(), even tick
// with IEEE binary32 round-to-nearest arithmetic and NO fused multiply-add (every product and sum is
// an explicit __fmul_rn/__fadd_rn; sqrt and divide are the correctly rounded __fsqrt_rn/__fdiv_rn),
// which is what the CPU oracle (oracle/msim_oracle.c) computes — positions come out bit-identical.
//
// B200 mapping: two kernels per move pass.
//   pass A  move_kernel    HBM-bound streaming: per entity read pos (8 B) + target (8 B), write pos
//                          (8 B) [+ cell key 4 B]; two entities per 128-bit load/store, .cs hints so
//                          the one-touch state does not displace the road tables in L2.  Entities
//                          that reach their waypoint are only flagged (1 bit).
//   pass B  arrive_kernel  the ~1/25 flagged entities pick their next waypoint: dependent gathers
//                          into the road / connection tables (read-only path) and the RNG state,
//                          32 arrivals per warp per round, balanced with a warp scan.
// `direction` is never stored: it is a pure function of (previous pos, target, arrival bit) and is
// rebuilt at readback (pack.cu), so the per-tick traffic stays at the 24 B minimum.
#include "msim_internal.h"

namespace msim {
namespace {

constexpr int MOVE_THREADS = 256;
constexpr int MOVE_ITEMS = 2;  // float4 pairs per thread -> 4 entities per thread per iteration
constexpr float SPEED = 1.4f;  // random_move.comp:750

__device__ __forceinline__ float as_f(uint32_t u) { return __uint_as_float(u); }

// random_move.comp:725-736
__device__ __forceinline__ uint32_t xorshift128(uint4& s) {
    uint32_t t = s.w;
    const uint32_t x = s.x;
    s.w = s.z;
    s.z = s.y;
    s.y = x;
    t ^= t << 11;
    t ^= t >> 8;
    s.x = t ^ x ^ (x >> 19);
    return s.x;
}

// random_move.comp:738-746: uint(ceil(float(min) + (next_float(state) * float(max - min + 1)))) - 1
__device__ __forceinline__ uint32_t next_range(uint4& s, uint32_t lo, uint32_t hi) {
    const float f = __fmul_rn(__uint2float_rn(xorshift128(s)), 2.3283064365386962890625e-10f);  // * 2^-32, exact
    const float prod = __fmul_rn(f, __uint2float_rn(hi - lo + 1u));
    const float sum = __fadd_rn(__uint2float_rn(lo), prod);
    return __float2uint_ru(sum) - 1u;  // ceil, then the conversion is exact
}

// connections[] with the canonical out-of-bounds rule: past-the-end reads yield road 0
__device__ __forceinline__ uint32_t read_connection(const uint32_t* __restrict__ conn, uint64_t count, uint64_t idx) {
    return idx < count ? __ldg(conn + idx) : 0u;
}

// random_move.comp:778-828.  `tgt` is the waypoint just reached; returns the new waypoint.
__device__ __forceinline__ float2 new_target(uint32_t e, float2 tgt, uint32_t* __restrict__ road, uint4* __restrict__ rng,
                                             const uint4* __restrict__ roads, const uint32_t* __restrict__ conn,
                                             uint64_t conn_count) {
    const uint32_t cur = road[e];
    const uint4 a = __ldg(roads + 2ull * cur);      // start: pos.x pos.y connectedIndex connectedCount
    const uint4 b = __ldg(roads + 2ull * cur + 1);  // end
    const bool at_start = (tgt.x == as_f(a.x)) && (tgt.y == as_f(a.y));
    const uint4 here = at_start ? a : b;
    const uint4 far = at_start ? b : a;
    if (here.w <= 1u) {  // dead end: turn around, road unchanged (:786-789, :794-797)
        return make_float2(as_f(far.x), as_f(far.y));
    }
    uint32_t next_road;
    if (here.w == 2u) {  // :802-804
        next_road = read_connection(conn, conn_count, static_cast<uint64_t>(here.z) + 1ull);
    } else {  // :805-810
        uint4 s = rng[e];
        const uint32_t off = next_range(s, 1u, here.w);
        rng[e] = s;
        next_road = read_connection(conn, conn_count, static_cast<uint64_t>(here.z) + off);
    }
    const uint4 ns = __ldg(roads + 2ull * next_road);
    const uint4 ne = __ldg(roads + 2ull * next_road + 1);  // same 32-byte sector as ns
    road[e] = next_road;                                  // :820
    const bool from_start = (as_f(ns.x) == tgt.x) && (as_f(ns.y) == tgt.y);  // :814-819
    return from_start ? make_float2(as_f(ne.x), as_f(ne.y)) : make_float2(as_f(ns.x), as_f(ns.y));
}

// ---- pass A: pure streaming ------------------------------------------------------------------
// update_direction(index, pos) + the walking branch of move(index) (:830-847).  Entities that reach
// their waypoint land exactly on it (:848) and are only flagged here.
__device__ __forceinline__ float2 walk(float2 p, float2 t, bool& arrived) {
    const float dx = __fsub_rn(t.x, p.x);
    const float dy = __fsub_rn(t.y, p.y);
    // length(target - pos) == distance(pos, target) bit for bit: the squares are sign-blind
    const float len = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    arrived = !(len > SPEED);
    if (arrived) return t;
    const float dirx = __fmul_rn(__fdiv_rn(dx, len), SPEED);
    const float diry = __fmul_rn(__fdiv_rn(dy, len), SPEED);
    return make_float2(__fadd_rn(p.x, dirx), __fadd_rn(p.y, diry));
}

// Per-cell population for the counting sort (csort.cu).  Storage is kept in cell order, so neighbouring lanes mostly hold the
// same key: each run of equal keys in adjacent lanes issues ONE reduction (by its first lane) - a few times fewer L2 operations
// than one per entity - and the reduction returns nothing (RED): no L2 round trip to wait for, the streaming loop never stalls
// on it.  The slot inside the cell is taken later by the scatter kernel, whose atomics return it directly.
// `full`: every lane of the warp holds an entity that counts (warp-uniform; true everywhere but in the last warp of the array and,
// on sharded handles, in warps that hold a leaver).  Otherwise the lanes that do not count break the runs.
__device__ __forceinline__ void run_count(uint32_t* __restrict__ cell_count, uint32_t key, bool valid, bool full, uint32_t lane) {
    if (!full) key = valid ? key : 0xffffffffu - lane;  // a key nobody shares: its own run, never issued
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane == 0 || key != prev;
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (head && (full || valid)) {
        const uint32_t above = (heads >> lane) >> 1;  // heads behind this lane, shifted down to bit 0
        const uint32_t len = above ? static_cast<uint32_t>(__ffs(above)) : 32u - lane;  // lanes up to the next head
        atomicAdd(&cell_count[key], len);
    }
}

// ---- fused shard pack (multi-GPU bands, shard.cu) ------------------------------------------------
// Whole warp calls this once per entity slot.  Storage is in cell order, so only the warps at either end of
// the slot range ever see a boundary row: everybody else leaves after one vote.
//   leaver (new cell row outside the band)  -> its slot goes on the hole list, its position on the local ghost list, its 72-byte
//                                              record into that side's exchange buffer
//   stays in the band's first / last row    -> its position is appended to the halo list of that side's exchange
//                                              buffer (local send buffer, or the neighbour's receive buffer over
//                                              NVLink peer memory: plain stores, the slot comes from a LOCAL counter)
// a leaver's 72-byte migrant record, written by the thread that moved it (rare: a few hundred entities per boundary per tick; kept out
// of line so that the streaming loop does not carry its registers).  The record holds the state BEFORE pass B (target = the waypoint
// just reached, arrival bit set): pass B runs on whichever GPU owns the entity after the exchange and yields the same result there,
// because new_target() reads nothing but the entity and the replicated road graph.
// Everything but the new position is re-read from memory: nothing of it has to stay in the streaming loop's registers.
__device__ __noinline__ void shard_write_record(void* buf, uint32_t slot, uint32_t e, float2 p_new, bool arrived, const float2* pos_old, const float2* target,
                                                const uint4* rng, const float4* color0, const uint32_t* road, const uint32_t* gid) {
    uint2* rec = records_of(buf) + static_cast<size_t>(slot) * (MIGRANT_BYTES / 8);
    const float2 p_old = pos_old[e], t = target[e];
    const uint4 r = rng[e];
    const float4 c = color0[e];
    rec[0] = make_uint2(__float_as_uint(p_new.x), __float_as_uint(p_new.y));
    rec[1] = make_uint2(__float_as_uint(p_old.x), __float_as_uint(p_old.y));
    rec[2] = make_uint2(__float_as_uint(t.x), __float_as_uint(t.y));
    rec[3] = make_uint2(r.x, r.y);
    rec[4] = make_uint2(r.z, r.w);
    rec[5] = make_uint2(__float_as_uint(c.x), __float_as_uint(c.y));
    rec[6] = make_uint2(__float_as_uint(c.z), __float_as_uint(c.w));
    rec[7] = make_uint2(road[e], gid[e]);
    rec[8] = make_uint2(arrived ? 1u : 0u, 0u);
}

__device__ __forceinline__ void shard_classify(const ShardMoveArgs& sh, uint32_t e, bool valid, uint32_t key, float2 p_new, bool arrived, const float2* pos_old,
                                               const float2* target) {
    const bool low = valid && sh.buf_down && key < sh.lo_key + sh.ncx;   // leaves downwards or sits in the first row
    const bool high = valid && sh.buf_up && key >= sh.hi_key - sh.ncx;   // leaves upwards or sits in the last row
    if (!__any_sync(0xffffffffu, low || high)) return;
    const bool leaves_down = low && key < sh.lo_key, leaves_up = high && key >= sh.hi_key;
    const bool leaves = leaves_down || leaves_up;
    const uint32_t hslot = warp_append(leaves, &sh.ctr[SHARD_CTR_HOLES]);
    if (leaves && hslot < sh.holes_cap) {
        sh.holes[hslot] = e;
        sh.local_ghosts[hslot] = p_new;  // it lands in the neighbour's boundary row: still within reach of ours
    }
    if (sh.buf_down) {
        const uint32_t mslot = warp_append(leaves_down, &sh.ctr[SHARD_CTR_MIG_DOWN]);
        if (leaves_down && mslot < sh.mig_cap) shard_write_record(sh.buf_down, mslot, e, p_new, arrived, pos_old, target, sh.rng, sh.color0, sh.road, sh.gid);  // beyond the capacity the count alone reports the overflow
        const bool halo = low && !leaves;
        const uint32_t slot = warp_append(halo, &sh.ctr[SHARD_CTR_HALO_DOWN]);
        if (halo && slot < sh.halo_cap) halo_of(sh.buf_down, sh.mig_cap)[slot] = p_new;
    }
    if (sh.buf_up) {
        const uint32_t mslot = warp_append(leaves_up, &sh.ctr[SHARD_CTR_MIG_UP]);
        if (leaves_up && mslot < sh.mig_cap) shard_write_record(sh.buf_up, mslot, e, p_new, arrived, pos_old, target, sh.rng, sh.color0, sh.road, sh.gid);
        const bool halo = high && !leaves;
        const uint32_t slot = warp_append(halo, &sh.ctr[SHARD_CTR_HALO_UP]);
        if (halo && slot < sh.halo_cap) halo_of(sh.buf_up, sh.mig_cap)[slot] = p_new;
    }
}

// MODE selects what the pass hands to the neighbour rebuild that follows it:
//   MOVE_PLAIN  nothing (collisions off): the 24 B per entity minimum
//   MOVE_COUNT  the per-cell population (csort.cu, single-GPU default): cell key of the new position -> one RED per run of
//               equal keys in adjacent lanes.  No key or rank is written: the scatter kernel recomputes the key from the
//               position it has to read anyway and takes the slot from an atomic on the scanned table
//   MOVE_KEYS   cell key (4 B) plus either the per-cell population as above (sharded handles: the exchange kernels work on keys) or
//               the radix sort's digit histograms for all passes, accumulated in shared memory and flushed once per CTA (onesweep)
// SHARD (multi-GPU bands, MOVE_KEYS only) additionally does the shard pack for the entities it has just moved (see above).
enum { MOVE_PLAIN = 0, MOVE_COUNT = 1, MOVE_KEYS = 2 };
template <int MODE, bool SHARD>
__global__ void __launch_bounds__(MOVE_THREADS)
move_kernel(uint32_t n_host, const uint32_t* __restrict__ n_dev, const float4* __restrict__ pos_in, float4* __restrict__ pos_out, const float4* __restrict__ target,
            uint32_t* __restrict__ arrived_mask, uint2* __restrict__ keys, GridParams grid, uint32_t* __restrict__ ghist, int hist_passes,
            uint32_t* __restrict__ cell_count, ShardMoveArgs sh) {
    static_assert(!SHARD || MODE == MOVE_KEYS, "the shard pack classifies by key");
    constexpr bool EMIT_KEYS = MODE == MOVE_KEYS;
    __shared__ uint32_t s_hist[EMIT_KEYS ? MAX_SORT_PASSES * RADIX : 1];
    if (EMIT_KEYS) {
        for (int i = threadIdx.x; i < MAX_SORT_PASSES * RADIX; i += MOVE_THREADS) s_hist[i] = 0;
        __syncthreads();
    }
    const uint32_t n = n_dev ? *n_dev : n_host;
    const uint32_t pairs = (n + 1u) >> 1;
    const uint32_t pairs_pad = (pairs + 31u) & ~31u;  // arrays are padded, whole warps stay converged
    const uint32_t lane = threadIdx.x & 31u;
    constexpr uint32_t PER_BLOCK = MOVE_THREADS * MOVE_ITEMS;

    for (uint32_t base = blockIdx.x * PER_BLOCK; base < pairs_pad; base += gridDim.x * PER_BLOCK) {
        float4 P[MOVE_ITEMS], T[MOVE_ITEMS];
        bool live[MOVE_ITEMS];
#pragma unroll
        for (int k = 0; k < MOVE_ITEMS; k++) {
            const uint32_t pi = base + k * MOVE_THREADS + threadIdx.x;
            live[k] = pi < pairs_pad;
            if (live[k]) {
                P[k] = __ldcs(pos_in + pi);
                T[k] = __ldcs(target + pi);
            }
        }
#pragma unroll
        for (int k = 0; k < MOVE_ITEMS; k++) {
            if (!live[k]) continue;  // warp-uniform
            const uint32_t pi = base + k * MOVE_THREADS + threadIdx.x;
            const uint32_t e0 = pi * 2u, e1 = e0 + 1u;
            bool arr0 = false, arr1 = false;
            float2 q0 = make_float2(P[k].x, P[k].y), q1 = make_float2(P[k].z, P[k].w);
            if (e0 < n) q0 = walk(q0, make_float2(T[k].x, T[k].y), arr0);
            if (e1 < n) q1 = walk(q1, make_float2(T[k].z, T[k].w), arr1);
            __stcs(pos_out + pi, make_float4(q0.x, q0.y, q1.x, q1.y));
            const uint32_t m0 = __ballot_sync(0xffffffffu, arr0);
            const uint32_t m1 = __ballot_sync(0xffffffffu, arr1);
            if (lane == 0) {
                const uint32_t w = (pi >> 5) * 2u;
                *reinterpret_cast<uint2*>(arrived_mask + w) = make_uint2(m0, m1);
            }
            if (MODE == MOVE_COUNT) {
                const bool full = ((pi | 31u) * 2u + 1u) < n;  // warp-uniform
                run_count(cell_count, cell_key_of(q0, grid), e0 < n, full, lane);
                run_count(cell_count, cell_key_of(q1, grid), e1 < n, full, lane);
            }
            if (EMIT_KEYS) {
                const uint32_t k0 = cell_key_of(q0, grid), k1 = cell_key_of(q1, grid);
                keys[pi] = make_uint2(k0, k1);
                if (cell_count) {  // counting sort: per-cell population of the entities that stay
                    bool v0 = e0 < n, v1 = e1 < n;
                    if (SHARD) {  // leavers are not part of this band's order any more; arrivals are counted by the integrate kernel
                        v0 = v0 && !((sh.buf_down && k0 < sh.lo_key) || (sh.buf_up && k0 >= sh.hi_key));
                        v1 = v1 && !((sh.buf_down && k1 < sh.lo_key) || (sh.buf_up && k1 >= sh.hi_key));
                    }
                    const bool full = __all_sync(0xffffffffu, v0 && v1);
                    run_count(cell_count, k0, v0, full, lane);
                    run_count(cell_count, k1, v1, full, lane);
                }
                for (int p = 0; p < hist_passes; p++) {
                    if (e0 < n) atomicAdd(&s_hist[p * RADIX + ((k0 >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
                    if (e1 < n) atomicAdd(&s_hist[p * RADIX + ((k1 >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
                }
                if (SHARD) {
                    shard_classify(sh, e0, e0 < n, k0, q0, arr0, reinterpret_cast<const float2*>(pos_in), reinterpret_cast<const float2*>(target));
                    shard_classify(sh, e1, e1 < n, k1, q1, arr1, reinterpret_cast<const float2*>(pos_in), reinterpret_cast<const float2*>(target));
                }
            }
        }
    }
    if (EMIT_KEYS) {
        __syncthreads();
        for (int i = threadIdx.x; i < hist_passes * RADIX; i += MOVE_THREADS)
            if (ghist && s_hist[i]) atomicAdd(&ghist[i], s_hist[i]);
    }
}

// ---- pass B: next waypoint for the entities that arrived (new_target, :778-828) ----------------
// Doing this inside pass A serialised up to four dependent gather chains per thread and left the
// streaming loads waiting behind them (80 % long-scoreboard stalls, profiles/r1a).  Here a warp owns
// 32 words of the arrival bitmask (1024 entities, ~40 arrivals), ranks the set bits with a warp scan
// and hands exactly one arrival to each lane per round: the gather chains run 32-wide and balanced.
constexpr int ARRIVE_THREADS = 256;

// STRIDE = false (default): the grid covers every mask word, one pass.  STRIDE = true (pass B beside a query, MSIM_ARRIVE_BESIDE_CTAS): the grid is capped
// at the resident CTAs and strides over the words (10 M entities need 1221 CTAs where 1184 are resident: no left-over wave).
template <bool STRIDE>
__global__ void __launch_bounds__(ARRIVE_THREADS)
arrive_kernel(uint32_t n_host, const uint32_t* __restrict__ n_dev, const uint32_t* __restrict__ arrived_mask, float2* __restrict__ target,
              uint32_t* __restrict__ road, uint4* __restrict__ rng, const uint4* __restrict__ roads,
              const uint32_t* __restrict__ conn, uint64_t conn_count) {
    const uint32_t n = n_dev ? *n_dev : n_host;
    const uint32_t words = ((n + 63u) >> 6) << 1;  // two mask words per 64-entity chunk
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t w = blockIdx.x * ARRIVE_THREADS + threadIdx.x;  // one mask word per lane
    do {
        const uint32_t mask = (w < words) ? __ldcs(arrived_mask + w) : 0u;
        const uint32_t cnt = __popc(mask);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= static_cast<uint32_t>(d)) incl += up;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        for (uint32_t base = 0; base < total; base += 32u) {
            const uint32_t r = base + lane;  // rank of the arrival this lane handles
            // owner = first lane whose inclusive count exceeds r (binary search over the warp's counts)
            uint32_t lo = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t probe = __shfl_sync(0xffffffffu, incl, (lo + step - 1) & 31u);
                if (probe <= r) lo += step;
            }
            const uint32_t owner = min(lo, 31u);
            const uint32_t owner_mask = __shfl_sync(0xffffffffu, mask, owner);
            const uint32_t owner_incl = __shfl_sync(0xffffffffu, incl, owner);
            const uint32_t owner_cnt = __shfl_sync(0xffffffffu, cnt, owner);
            if (r < total) {
                const uint32_t nth = r - (owner_incl - owner_cnt);          // 0-based among the owner's set bits
                const uint32_t bit = __fns(owner_mask, 0, static_cast<int>(nth) + 1);
                const uint32_t ow = (w - lane) + owner;                     // the owner's mask word index
                const uint32_t e = (ow >> 1) * 64u + bit * 2u + (ow & 1u);  // inverse of arrived_word/arrived_bit
                if (e < n) target[e] = new_target(e, target[e], road, rng, roads, conn, conn_count);
            }
        }
        if (!STRIDE) break;
        w += gridDim.x * ARRIVE_THREADS;
    } while (w - threadIdx.x < words);  // CTA-uniform: whole warps stay together for the shuffles
}

// One thread right behind the fused move + pack kernel (collective exchange): writes the list lengths into the send buffers' headers.
// (Doing this at the end of the move kernel itself - last CTA to take a ticket, one system-scope fence per CTA - was measured: the pass
// grew from 41 to 64 us at 5 M entities, far more than this launch costs.)
__global__ void shard_publish_kernel(ShardMoveArgs sh) { shard_publish(sh); }

__global__ void __launch_bounds__(256)
keygen_kernel(uint32_t n, const float4* __restrict__ pos, uint2* __restrict__ keys, GridParams grid) {
    const uint32_t pairs_pad = (((n + 1u) >> 1) + 31u) & ~31u;
    for (uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x; pi < pairs_pad; pi += gridDim.x * blockDim.x) {
        const float4 p = __ldcs(pos + pi);
        keys[pi] = make_uint2(cell_key_of(make_float2(p.x, p.y), grid), cell_key_of(make_float2(p.z, p.w), grid));
    }
}

}  // namespace

// `keys` non-NULL: MOVE_KEYS (key + per-cell population or digit histograms, optionally the shard pack); `keys` NULL and `cell_count` non-NULL:
// MOVE_COUNT; neither: MOVE_PLAIN.  The grid is 8 CTAs per SM (2048 threads) with a grid-stride loop inside.
int launch_move(cudaStream_t s, int sm_count, uint32_t n, const float2* pos_in, float2* pos_out, const float2* target, uint32_t* arrived,
                uint32_t* keys, const GridParams& grid, uint32_t* hist, int hist_passes, uint32_t* cell_count, Profiler* prof,
                const uint32_t* n_dev, const ShardMoveArgs* shard, int ctas_per_sm) {
    if (n == 0) {  // (a band without entities still has to publish its empty lists)
        if (!shard || !shard->publish) return 0;
        shard_publish_kernel<<<1, 1, 0, s>>>(*shard);
        return 1;
    }
    const uint32_t pairs = (n + 1u) >> 1;
    const uint32_t per_block = MOVE_THREADS * MOVE_ITEMS;
    uint32_t blocks = (pairs + per_block - 1) / per_block;
    const uint32_t resident = static_cast<uint32_t>(sm_count) * static_cast<uint32_t>(ctas_per_sm > 0 && ctas_per_sm < 8 ? ctas_per_sm : 8);
    if (blocks > resident) blocks = resident;
    const float4* pin = reinterpret_cast<const float4*>(pos_in);
    float4* pout = reinterpret_cast<float4*>(pos_out);
    const float4* tgt = reinterpret_cast<const float4*>(target);
    uint2* keys2 = reinterpret_cast<uint2*>(keys);
    const int passes = hist ? hist_passes : 0;
    const ShardMoveArgs none{};
    prof->begin(s, K_MOVE);
    if (keys && shard) {
        move_kernel<MOVE_KEYS, true><<<blocks, MOVE_THREADS, 0, s>>>(n, n_dev, pin, pout, tgt, arrived, keys2, grid, hist, passes, cell_count, *shard);
        if (shard->publish) shard_publish_kernel<<<1, 1, 0, s>>>(*shard);
    } else if (keys)
        move_kernel<MOVE_KEYS, false><<<blocks, MOVE_THREADS, 0, s>>>(n, n_dev, pin, pout, tgt, arrived, keys2, grid, hist, passes, cell_count, none);
    else if (cell_count)
        move_kernel<MOVE_COUNT, false><<<blocks, MOVE_THREADS, 0, s>>>(n, n_dev, pin, pout, tgt, arrived, nullptr, grid, nullptr, 0, cell_count, none);
    else
        move_kernel<MOVE_PLAIN, false><<<blocks, MOVE_THREADS, 0, s>>>(n, n_dev, pin, pout, tgt, arrived, nullptr, grid, nullptr, 0, nullptr, none);
    prof->end(s);
    return 1;
}

int launch_arrive(cudaStream_t s, uint32_t n, float2* target, uint32_t* road, uint4* rng, const uint32_t* arrived, const msim_road* roads,
                  const uint32_t* connections, uint64_t connection_count, Profiler* prof, const uint32_t* n_dev, bool beside, int beside_ctas_per_sm) {
    if (n == 0) return 0;
    const uint32_t words = ((n + 63u) >> 6) << 1;  // grid size (n is an upper bound when n_dev is given)
    uint32_t blocks = (words + ARRIVE_THREADS - 1) / ARRIVE_THREADS;
    // 10 M entities need 1221 CTAs where 1184 are resident (31 registers, 8 per SM): the 37 left over start when the first finish.
    // beside a query: a strided grid of `per_sm` CTAs per SM instead (Tuning::arrive_beside_ctas_per_sm)
    uint32_t cap = 0u;
    const int per_sm = beside_ctas_per_sm >= 0 ? beside_ctas_per_sm : tuning().arrive_beside_ctas_per_sm;
    if (beside && per_sm) cap = 148u * static_cast<uint32_t>(per_sm);
    const bool stride = cap != 0u && blocks > cap;
    prof->begin(s, K_ARRIVE);
    if (stride)
        arrive_kernel<true><<<cap, ARRIVE_THREADS, 0, s>>>(n, n_dev, arrived, target, road, rng, reinterpret_cast<const uint4*>(roads), connections, connection_count);
    else
        arrive_kernel<false><<<blocks, ARRIVE_THREADS, 0, s>>>(n, n_dev, arrived, target, road, rng, reinterpret_cast<const uint4*>(roads), connections, connection_count);
    prof->end(s);
    return 1;
}

int launch_keygen(cudaStream_t s, uint32_t n, const float2* pos, uint32_t* keys, const GridParams& grid, Profiler* prof) {
    if (n == 0) return 0;
    const uint32_t pairs = (n + 1u) >> 1;
    uint32_t blocks = (pairs + 255u) / 256u;
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    prof->begin(s, K_KEYGEN);
    keygen_kernel<<<blocks, 256, 0, s>>>(n, reinterpret_cast<const float4*>(pos), reinterpret_cast<uint2*>(keys), grid);
    prof->end(s);
    return 1;
}

}  // namespace msim
