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
// PREFETCH_RNG (fused pass B only): the RNG state is requested together with the road index instead of after the road record has
// told us that more than two roads meet here — one DRAM latency less on the dependent chain road -> record -> connection -> record,
// for 16 bytes read in vain at dead ends and two-way points.  Same results: the state is only written back when it was drawn from.
template <bool PREFETCH_RNG = false>
__device__ __forceinline__ float2 new_target(uint32_t e, float2 tgt, uint32_t* __restrict__ road, uint4* __restrict__ rng,
                                             const uint4* __restrict__ roads, const uint32_t* __restrict__ conn,
                                             uint64_t conn_count) {
    uint4 early = make_uint4(0u, 0u, 0u, 0u);
    if (PREFETCH_RNG) early = rng[e];
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
        uint4 s = PREFETCH_RNG ? early : rng[e];
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

// Rank of one entity per lane inside its cell = old value of the cell's counter.  Storage is kept in
// cell order, so neighbouring lanes mostly hold the same key: each run of equal keys in adjacent lanes
// issues ONE atomic (by its first lane) and shares the result — a few times fewer L2 atomics than one per
// entity.  Ranks inside a cell are a permutation either way; nothing observable depends on their order.
struct RunRank {
    uint32_t base;     // the run head's atomic result (valid in the head lane only, until finished)
    uint32_t my_head;  // lane of the head of this lane's run
};
// phase 1: find the runs and issue the heads' atomics; the result is not consumed here, so several of these can be
// in flight per thread before anybody waits (the atomics' L2 round trips dominated the kernel: profiles/r1m)
__device__ __forceinline__ RunRank run_rank_issue(uint32_t* __restrict__ cell_count, uint32_t key, bool valid, uint32_t lane) {
    const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool prev_valid = __shfl_up_sync(0xffffffffu, valid ? 1u : 0u, 1) != 0u;
    const bool head = lane == 0 || key != prev || !valid || !prev_valid;
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    RunRank r;
    r.my_head = 31u - __clz(heads & ((2u << lane) - 1u));        // nearest head at or below this lane (lane 0 is one)
    const uint32_t above = r.my_head == 31u ? 0u : (heads >> (r.my_head + 1u)) << (r.my_head + 1u);
    const uint32_t next_head = above ? static_cast<uint32_t>(__ffs(above) - 1) : 32u;
    r.base = 0;
    if (lane == r.my_head && valid) r.base = atomicAdd(&cell_count[key], next_head - r.my_head);
    return r;
}
// phase 2: every lane of a run takes the head's result plus its distance from the head
__device__ __forceinline__ uint32_t run_rank_finish(const RunRank& r, uint32_t lane) {
    return __shfl_sync(0xffffffffu, r.base, r.my_head) + (lane - r.my_head);
}

// ---- fused shard pack (multi-GPU bands, shard.cu) ------------------------------------------------
// Whole warp calls this once per entity slot.  Storage is in cell order, so only the warps at either end of
// the slot range ever see a boundary row: everybody else leaves after one vote.
//   leaver (new cell row outside the band)  -> its slot goes on the hole list; shard_emit_kernel (shard.cu), a
//                                              one-CTA kernel right behind this one, writes the migrant records
//   stays in the band's first / last row    -> its position is appended to the halo list of that side's exchange
//                                              buffer (local send buffer, or the neighbour's receive buffer over
//                                              NVLink peer memory: plain stores, the slot comes from a LOCAL counter)
__device__ __forceinline__ void shard_classify(const ShardMoveArgs& sh, uint32_t e, bool valid, uint32_t key, float2 p_new) {
    const bool low = valid && sh.buf_down && key < sh.lo_key + sh.ncx;   // leaves downwards or sits in the first row
    const bool high = valid && sh.buf_up && key >= sh.hi_key - sh.ncx;   // leaves upwards or sits in the last row
    if (!__any_sync(0xffffffffu, low || high)) return;
    const bool leaves = (low && key < sh.lo_key) || (high && key >= sh.hi_key);
    const uint32_t hslot = warp_append(leaves, &sh.ctr[SHARD_CTR_HOLES]);
    if (leaves && hslot < sh.holes_cap) sh.holes[hslot] = e;
    if (sh.buf_down) {
        const bool halo = low && !leaves;
        const uint32_t slot = warp_append(halo, &sh.ctr[SHARD_CTR_HALO_DOWN]);
        if (halo && slot < sh.halo_cap) halo_of(sh.buf_down, sh.mig_cap)[slot] = p_new;
    }
    if (sh.buf_up) {
        const bool halo = high && !leaves;
        const uint32_t slot = warp_append(halo, &sh.ctr[SHARD_CTR_HALO_UP]);
        if (halo && slot < sh.halo_cap) halo_of(sh.buf_up, sh.mig_cap)[slot] = p_new;
    }
}

// ---- fused pass B (opt-in, MSIM_FLAG_FUSED_ARRIVE) -----------------------------------------------------------
// The arrival bits of the PREVIOUS move pass are consumed by the warp that streams the same entities in THIS pass: bit
// `lane` of the two mask words of a 64-entity chunk belongs to the lane that loads that entity's position and target, so
// nothing has to change lanes.  A lane takes one of its (up to four) pending arrivals per round; one round serves up to
// 32 arrivals of the warp with their gather chains side by side, a second round is rare (two arrivals in one lane).
// The streaming loads of the iteration are already in flight while the chains run, and the mask words are read before
// this iteration's arrival bits overwrite them (same warp, program order).
struct FusedArrive {
    float2* target;  // read-write alias of the move kernel's `target` (which is therefore not read through __restrict__)
    uint32_t* road;
    uint4* rng;
    const uint4* roads;
    const uint32_t* conn;
    unsigned long long conn_count;
    uint32_t consume;  // 0: no pass B is pending (first move after an upload / after a stand-alone pass B)
};

// EMIT_KEYS additionally writes the cell key of the new position (4 B) and accumulates the radix
// sort's digit histograms for all passes in shared memory (flushed once per CTA), so the neighbour
// rebuild needs no separate histogram read of the keys.
// SHARD (multi-GPU bands) additionally does the shard pack for the entities it has just moved (see above).
// MINB = minimum resident CTAs per SM asked of the compiler (register cap 65536 / (256 MINB)).  0 = no cap: 40 / 52 / 58
// registers without / with keys / sharded, i.e. 6 / 4 / 4 resident CTAs; the variants with keys wait on L2 atomics, so
// more resident warps may pay for a few spilled registers (MSIM_MOVE_MIN_BLOCKS = 5 or 6, see tuning() in api.cu).
template <bool EMIT_KEYS, bool SHARD, bool FUSE, int MINB>
__global__ void __launch_bounds__(MOVE_THREADS, MINB)
move_kernel(uint32_t n_host, const uint32_t* __restrict__ n_dev, const float4* __restrict__ pos_in, float4* __restrict__ pos_out, const float4* __restrict__ target,
            uint32_t* __restrict__ arrived_mask, uint2* __restrict__ keys, GridParams grid, uint32_t* __restrict__ ghist, int hist_passes,
            uint32_t* __restrict__ cell_count, uint2* __restrict__ rank, ShardMoveArgs sh, FusedArrive fa) {
    static_assert(!FUSE || MOVE_ITEMS == 2, "the fused pass B selects among 2 x 2 entity slots per lane");
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
        uint32_t pend = 0;  // FUSE: bit 2k + s = entity slot s of item k arrived in the previous pass and has no new waypoint yet
#pragma unroll
        for (int k = 0; k < MOVE_ITEMS; k++) {
            const uint32_t pi = base + k * MOVE_THREADS + threadIdx.x;
            live[k] = pi < pairs_pad;
            if (live[k]) {
                P[k] = __ldcs(pos_in + pi);
                if (FUSE) {
                    T[k] = __ldcs(reinterpret_cast<const float4*>(fa.target) + pi);
                    if (fa.consume) {
                        const uint2 pm = *reinterpret_cast<const uint2*>(arrived_mask + (pi >> 5) * 2u);  // one address per warp
                        const uint32_t e0 = pi * 2u;
                        if (((pm.x >> lane) & 1u) && e0 < n) pend |= 1u << (2 * k);
                        if (((pm.y >> lane) & 1u) && e0 + 1u < n) pend |= 2u << (2 * k);
                    }
                } else {
                    T[k] = __ldcs(target + pi);
                }
            }
        }
        if (FUSE) {
            while (__any_sync(0xffffffffu, pend != 0u)) {
                if (pend) {
                    const uint32_t slot = static_cast<uint32_t>(__ffs(static_cast<int>(pend))) - 1u;
                    pend &= pend - 1u;
                    const uint32_t pi = base + (slot >> 1) * MOVE_THREADS + threadIdx.x;
                    const uint32_t e = pi * 2u + (slot & 1u);
                    const float4 t4 = (slot >> 1) ? T[1] : T[0];
                    const float2 reached = (slot & 1u) ? make_float2(t4.z, t4.w) : make_float2(t4.x, t4.y);
                    const float2 nt = new_target<true>(e, reached, fa.road, fa.rng, fa.roads, fa.conn, fa.conn_count);
                    fa.target[e] = nt;
                    if (slot == 0u) { T[0].x = nt.x; T[0].y = nt.y; }
                    if (slot == 1u) { T[0].z = nt.x; T[0].w = nt.y; }
                    if (slot == 2u) { T[1].x = nt.x; T[1].y = nt.y; }
                    if (slot == 3u) { T[1].z = nt.x; T[1].w = nt.y; }
                }
            }
        }
        RunRank R0[MOVE_ITEMS], R1[MOVE_ITEMS];
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
            if (EMIT_KEYS) {
                const uint32_t k0 = cell_key_of(q0, grid), k1 = cell_key_of(q1, grid);
                keys[pi] = make_uint2(k0, k1);
                if (cell_count) {  // counting sort: the atomic's return value is the entity's rank inside its cell
                    bool v0 = e0 < n, v1 = e1 < n;
                    if (SHARD) {  // leavers are not part of this band's order any more; arrivals take their rank in the integrate kernel
                        v0 = v0 && !((sh.buf_down && k0 < sh.lo_key) || (sh.buf_up && k0 >= sh.hi_key));
                        v1 = v1 && !((sh.buf_down && k1 < sh.lo_key) || (sh.buf_up && k1 >= sh.hi_key));
                    }
                    R0[k] = run_rank_issue(cell_count, k0, v0, lane);
                    R1[k] = run_rank_issue(cell_count, k1, v1, lane);
                }
                for (int p = 0; p < hist_passes; p++) {
                    if (e0 < n) atomicAdd(&s_hist[p * RADIX + ((k0 >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
                    if (e1 < n) atomicAdd(&s_hist[p * RADIX + ((k1 >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
                }
                if (SHARD) {
                    shard_classify(sh, e0, e0 < n, k0, q0);
                    shard_classify(sh, e1, e1 < n, k1, q1);
                }
            }
        }
        if (EMIT_KEYS && cell_count) {  // all of this thread's rank atomics are in flight by now
#pragma unroll
            for (int k = 0; k < MOVE_ITEMS; k++) {
                if (!live[k]) continue;
                const uint32_t pi = base + k * MOVE_THREADS + threadIdx.x;
                rank[pi] = make_uint2(run_rank_finish(R0[k], lane), run_rank_finish(R1[k], lane));
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

// STRIDE = false (default): the grid covers every mask word, one pass.  STRIDE = true (MSIM_ARRIVE_GRID=persistent): the grid is capped
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

__global__ void __launch_bounds__(256)
keygen_kernel(uint32_t n, const float4* __restrict__ pos, uint2* __restrict__ keys, GridParams grid) {
    const uint32_t pairs_pad = (((n + 1u) >> 1) + 31u) & ~31u;
    for (uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x; pi < pairs_pad; pi += gridDim.x * blockDim.x) {
        const float4 p = __ldcs(pos + pi);
        keys[pi] = make_uint2(cell_key_of(make_float2(p.x, p.y), grid), cell_key_of(make_float2(p.z, p.w), grid));
    }
}

// one launch of the chosen variant; the grid is a whole number of resident CTAs per SM (grid-stride loop inside)
template <bool EMIT_KEYS, bool SHARD, bool FUSE, int MINB, typename... Args>
void launch_move_variant(cudaStream_t s, int sm_count, uint32_t blocks_needed, Args... args) {
    uint32_t per_sm = 8u;  // 8 x 256 threads = 2048 threads per SM (what the register budget of MINB = 0 allows is less: see above)
    if (tuning().move_grid_by_occupancy) {
        static int occupancy = 0;  // per instantiation
        if (occupancy == 0 &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occupancy, move_kernel<EMIT_KEYS, SHARD, FUSE, MINB>, MOVE_THREADS, 0) != cudaSuccess)
            occupancy = 8;
        per_sm = static_cast<uint32_t>(occupancy > 0 ? occupancy : 8);
    }
    const uint32_t resident = static_cast<uint32_t>(sm_count) * per_sm;
    const uint32_t blocks = blocks_needed > resident ? resident : blocks_needed;
    move_kernel<EMIT_KEYS, SHARD, FUSE, MINB><<<blocks, MOVE_THREADS, 0, s>>>(args...);
}
template <bool EMIT_KEYS, bool SHARD, bool FUSE, typename... Args>
void launch_move_minb(cudaStream_t s, int sm_count, uint32_t blocks_needed, Args... args) {
    switch (tuning().move_min_blocks) {
        case 5: launch_move_variant<EMIT_KEYS, SHARD, FUSE, 5>(s, sm_count, blocks_needed, args...); break;
        case 6: launch_move_variant<EMIT_KEYS, SHARD, FUSE, 6>(s, sm_count, blocks_needed, args...); break;
        default: launch_move_variant<EMIT_KEYS, SHARD, FUSE, 0>(s, sm_count, blocks_needed, args...); break;
    }
}

}  // namespace

int launch_move(cudaStream_t s, int sm_count, uint32_t n, const float2* pos_in, float2* pos_out, const float2* target, uint32_t* arrived,
                uint32_t* keys, const GridParams& grid, uint32_t* hist, int hist_passes, uint32_t* cell_count, uint32_t* rank, Profiler* prof,
                const uint32_t* n_dev, const ShardMoveArgs* shard, const FusedArriveArgs* fuse) {
    if (n == 0) return 0;
    const uint32_t pairs = (n + 1u) >> 1;
    const uint32_t per_block = MOVE_THREADS * MOVE_ITEMS;
    const uint32_t blocks = (pairs + per_block - 1) / per_block;
    const float4* pin = reinterpret_cast<const float4*>(pos_in);
    float4* pout = reinterpret_cast<float4*>(pos_out);
    const float4* tgt = reinterpret_cast<const float4*>(target);
    uint2* keys2 = reinterpret_cast<uint2*>(keys);
    uint2* rank2 = reinterpret_cast<uint2*>(rank);
    const int passes = hist ? hist_passes : 0;
    const float4* no_target = nullptr;  // the fused variants read the waypoints through FusedArrive::target (read-write alias)
    uint2* no_keys = nullptr;
    uint32_t* no_table = nullptr;
    prof->begin(s, K_MOVE);
    const ShardMoveArgs none{};
    FusedArrive fa{};
    if (fuse) {
        fa.target = fuse->target;
        fa.road = fuse->road;
        fa.rng = fuse->rng;
        fa.roads = reinterpret_cast<const uint4*>(fuse->roads);
        fa.conn = fuse->connections;
        fa.conn_count = fuse->connection_count;
        fa.consume = fuse->consume ? 1u : 0u;
    }
    if (keys && shard)  // sharded handles keep the stand-alone pass B: migrants travel with their pre-arrival state
        launch_move_minb<true, true, false>(s, sm_count, blocks, n, n_dev, pin, pout, tgt, arrived, keys2, grid, hist, passes, cell_count, rank2, *shard, fa);
    else if (keys && fuse)
        launch_move_minb<true, false, true>(s, sm_count, blocks, n, n_dev, pin, pout, no_target, arrived, keys2, grid, hist, passes, cell_count, rank2, none, fa);
    else if (keys)
        launch_move_minb<true, false, false>(s, sm_count, blocks, n, n_dev, pin, pout, tgt, arrived, keys2, grid, hist, passes, cell_count, rank2, none, fa);
    else if (fuse)
        launch_move_minb<false, false, true>(s, sm_count, blocks, n, n_dev, pin, pout, no_target, arrived, no_keys, grid, no_table, 0, no_table, no_keys, none, fa);
    else
        launch_move_minb<false, false, false>(s, sm_count, blocks, n, n_dev, pin, pout, tgt, arrived, no_keys, grid, no_table, 0, no_table, no_keys, none, fa);
    prof->end(s);
    return 1;
}

int launch_arrive(cudaStream_t s, uint32_t n, float2* target, uint32_t* road, uint4* rng, const uint32_t* arrived, const msim_road* roads,
                  const uint32_t* connections, uint64_t connection_count, Profiler* prof, const uint32_t* n_dev, bool beside) {
    if (n == 0) return 0;
    const uint32_t words = ((n + 63u) >> 6) << 1;  // grid size (n is an upper bound when n_dev is given)
    uint32_t blocks = (words + ARRIVE_THREADS - 1) / ARRIVE_THREADS;
    // 10 M entities need 1221 CTAs where 1184 are resident (31 registers, 8 per SM): the 37 left over start when the first finish.
    // MSIM_ARRIVE_GRID=persistent: one resident wave that strides over the words instead
    uint32_t cap = tuning().arrive_persistent ? 148u * 8u : 0u;
    if (beside && tuning().arrive_beside_ctas_per_sm) cap = 148u * static_cast<uint32_t>(tuning().arrive_beside_ctas_per_sm);
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
