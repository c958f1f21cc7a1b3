// shard.cu — kernels of the multi-GPU path: one msim handle per GPU owns a band of
// whole cell rows [row_lo, row_hi) of the neighbour grid; the road graph is replicated.  The reference
// has no multi-GPU path (one kp::Manager, /root/reference/src/sim/Simulator.cpp:52); this is new work
// whose results must equal the single-GPU run: same per-entity state, same colour flags, and a global
// unique-pair count that is the sum of the per-rank counts.
//
// Per sim tick, after the move pass:
//   pack      every owned entity whose new cell row left the band is written as a 72-byte migrant
//             record into the fixed-size buffer for the neighbour below / above and its slot becomes a
//             hole; entities staying in the band's first / last row are appended to the halo list
//             (8-byte positions) of the same buffer.  Leavers are also kept as local ghosts: they
//             land in the neighbour's boundary row, i.e. within reach of our own boundary row.
//   (NCCL)    the two buffers go to the neighbours, two come back — torch.distributed plumbing.
//   place     arrivals are written into holes / appended; relocate closes the remaining holes with
//             entities from the tail; append_ghosts puts halo + leaver positions after the owned
//             entities so the radix sort / cell build / query see them as read-only neighbours.
// Buffer layout (bytes): [32-byte header {n_migrants, n_halo, overflow}] [migrant_capacity x 72]
// [halo_capacity x 8].
#include "msim_internal.h"

namespace msim {
namespace {

__device__ __forceinline__ void set_arrived_bit(uint32_t* mask, uint32_t e, bool value) {
    uint32_t* w = mask + arrived_word(e);
    const uint32_t bit = 1u << arrived_bit(e);
    if (value) atomicOr(w, bit);
    else atomicAnd(w, ~bit);
}

// an entity that joins the cell order after the move kernel has counted the ones that stayed
__device__ __forceinline__ void count_cell(const ShardArrays& a, uint32_t key) {
    if (a.cell_count && key - a.c0 < a.c1 - a.c0) atomicAdd(&a.cell_count[key], 1u);
}

// one tiny launch instead of three memsets: clears the two headers and the hole / ghost counters
__global__ void shard_reset_kernel(void* buf_down, void* buf_up, uint32_t* ctr) {
    const uint32_t t = threadIdx.x;
    if (t < 8u) {
        if (buf_down) reinterpret_cast<uint32_t*>(buf_down)[t] = 0u;
        if (buf_up) reinterpret_cast<uint32_t*>(buf_up)[t] = 0u;
        if (t < SHARD_CTR_COUNT) ctr[t] = 0u;
    }
}

__global__ void __launch_bounds__(256)
shard_pack_kernel(ShardArrays a, uint32_t n_host, const uint32_t* __restrict__ n_dev, int ncx, uint32_t row_lo, uint32_t row_hi, void* buf_down, void* buf_up,
                  uint32_t mig_cap, uint32_t halo_cap, uint32_t* holes, uint32_t holes_cap, float2* local_ghosts, uint32_t* ctr) {
    const uint32_t n = n_dev ? *n_dev : n_host;
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = e < n;
    uint32_t row = 0;
    float2 p = make_float2(0.f, 0.f);
    if (live) {
        row = a.keys[e] / static_cast<uint32_t>(ncx);
        p = a.pos_cur[e];
    }
    const bool go_down = live && buf_down && row < row_lo;
    const bool go_up = live && buf_up && row >= row_hi;
    if (go_down || go_up) {  // rare: a few hundred entities per boundary per tick
        void* buf = go_down ? buf_down : buf_up;
        const uint32_t slot = atomicAdd(&header_of(buf)->n_migrants, 1u);
        if (slot < mig_cap) {
            uint2* rec = records_of(buf) + static_cast<size_t>(slot) * (MIGRANT_BYTES / 8);
            const float2 pp = a.pos_prev[e], t = a.target[e];
            const uint4 s = a.rng[e];
            const float4 c = a.color0[e];
            const uint32_t arrived = (a.arrived[arrived_word(e)] >> arrived_bit(e)) & 1u;
            rec[0] = make_uint2(__float_as_uint(p.x), __float_as_uint(p.y));
            rec[1] = make_uint2(__float_as_uint(pp.x), __float_as_uint(pp.y));
            rec[2] = make_uint2(__float_as_uint(t.x), __float_as_uint(t.y));
            rec[3] = make_uint2(s.x, s.y);
            rec[4] = make_uint2(s.z, s.w);
            rec[5] = make_uint2(__float_as_uint(c.x), __float_as_uint(c.y));
            rec[6] = make_uint2(__float_as_uint(c.z), __float_as_uint(c.w));
            rec[7] = make_uint2(a.road[e], a.gid[e]);
            rec[8] = make_uint2(arrived, 0u);
        } else {
            header_of(buf)->overflow = 1u;
        }
        const uint32_t hslot = atomicAdd(&ctr[SHARD_CTR_HOLES], 1u);
        if (hslot < holes_cap) holes[hslot] = e;
        const uint32_t gslot = atomicAdd(&ctr[SHARD_CTR_LOCAL_GHOSTS], 1u);
        if (gslot < holes_cap) local_ghosts[gslot] = p;
    }
    // halo: owned entities that stay, in the band's first / last row
    const bool halo_down = live && buf_down && !go_down && !go_up && row == row_lo;
    const bool halo_up = live && buf_up && !go_down && !go_up && row + 1u == row_hi;
    if (buf_down) {
        const uint32_t slot = warp_append(halo_down, &header_of(buf_down)->n_halo);
        if (halo_down) {
            if (slot < halo_cap) halo_of(buf_down, mig_cap)[slot] = p;
            else header_of(buf_down)->overflow = 1u;
        }
    }
    if (buf_up) {
        const uint32_t slot = warp_append(halo_up, &header_of(buf_up)->n_halo);
        if (halo_up) {
            if (slot < halo_cap) halo_of(buf_up, mig_cap)[slot] = p;
            else header_of(buf_up)->overflow = 1u;
        }
    }
}

// arrivals: record i of the concatenation [recv_down migrants][recv_up migrants] goes to slot dst[i]
__global__ void __launch_bounds__(128)
shard_place_kernel(ShardArrays a, const void* recv_down, uint32_t n_down, const void* recv_up, uint32_t n_up, const uint32_t* __restrict__ dst,
                   GridParams grid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_down + n_up) return;
    const void* buf = i < n_down ? recv_down : recv_up;
    const uint32_t r = i < n_down ? i : i - n_down;
    const uint2* rec = reinterpret_cast<const uint2*>(static_cast<const char*>(buf) + sizeof(ShardHeader)) + static_cast<size_t>(r) * (MIGRANT_BYTES / 8);
    const uint32_t e = dst[i];
    const uint2 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3], r4 = rec[4], r5 = rec[5], r6 = rec[6], r7 = rec[7], r8 = rec[8];
    const float2 p = make_float2(__uint_as_float(r0.x), __uint_as_float(r0.y));
    a.pos_cur[e] = p;
    a.pos_prev[e] = make_float2(__uint_as_float(r1.x), __uint_as_float(r1.y));
    a.target[e] = make_float2(__uint_as_float(r2.x), __uint_as_float(r2.y));
    a.rng[e] = make_uint4(r3.x, r3.y, r4.x, r4.y);
    a.color0[e] = make_float4(__uint_as_float(r5.x), __uint_as_float(r5.y), __uint_as_float(r6.x), __uint_as_float(r6.y));
    a.road[e] = r7.x;
    a.gid[e] = r7.y;
    const uint32_t key = cell_key_of(p, grid);
    a.keys[e] = key;
    count_cell(a, key);
    set_arrived_bit(a.arrived, e, (r8.x & 1u) != 0u);
}

// close the remaining holes: entity src[i] moves to slot dst[i]
__global__ void __launch_bounds__(128) shard_relocate_kernel(ShardArrays a, const uint2* __restrict__ moves, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t src = moves[i].x, dst = moves[i].y;
    a.pos_cur[dst] = a.pos_cur[src];
    a.pos_prev[dst] = a.pos_prev[src];
    a.target[dst] = a.target[src];
    a.rng[dst] = a.rng[src];
    a.color0[dst] = a.color0[src];
    a.road[dst] = a.road[src];
    a.gid[dst] = a.gid[src];
    a.keys[dst] = a.keys[src];
    set_arrived_bit(a.arrived, dst, ((a.arrived[arrived_word(src)] >> arrived_bit(src)) & 1u) != 0u);
}

// ghosts = halo from below + halo from above + our own leavers, appended behind the owned entities
__global__ void __launch_bounds__(256)
shard_append_ghosts_kernel(ShardArrays a, uint32_t first, const void* recv_down, uint32_t h_down, const void* recv_up, uint32_t h_up,
                           const float2* __restrict__ local_ghosts, uint32_t h_local, uint32_t mig_cap, GridParams grid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h_down + h_up + h_local) return;
    float2 p;
    if (i < h_down) p = halo_of(const_cast<void*>(recv_down), mig_cap)[i];
    else if (i < h_down + h_up) p = halo_of(const_cast<void*>(recv_up), mig_cap)[i - h_down];
    else p = local_ghosts[i - h_down - h_up];
    a.pos_cur[first + i] = p;
    const uint32_t key = cell_key_of(p, grid);
    a.keys[first + i] = key;
    count_cell(a, key);
}

// ---- second half of the fused move + pack ---------------------------------------------------------
// One CTA right behind the move kernel: turns the hole list the move kernel left (slots of the leavers, a few hundred
// per boundary per tick) into migrant records in the exchange buffers, keeps the leavers' positions as local ghosts,
// publishes the list lengths into the buffer headers and - peer-memory exchange - raises the neighbours' flags.
// Records carry the state BEFORE pass B (target = the waypoint just reached, arrival bit set): pass B runs on
// whichever GPU owns the entity after the exchange and yields the same result there, because new_target() reads
// nothing but the entity and the replicated road graph.
// The buffers may live in a neighbour's memory: they only ever see plain stores; every counter is local.  The move
// kernel's halo stores are complete when this kernel starts (stream order); this kernel's own stores are fenced at
// system scope before the flags go up.
constexpr int EMIT_THREADS = 1024;

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

__device__ __forceinline__ void emit_body(const ShardArrays& a, const ShardMoveArgs& sh) {
    __shared__ uint32_t s_mig[2];
    const uint32_t tid = threadIdx.x;
    if (tid < 2) s_mig[tid] = 0;
    __syncthreads();
    const uint32_t k_out = min(sh.ctr[SHARD_CTR_HOLES], sh.holes_cap);
    for (uint32_t i = tid; i < k_out; i += EMIT_THREADS) {
        const uint32_t e = sh.holes[i];
        const bool down = sh.buf_down && a.keys[e] < sh.lo_key;
        const float2 p = a.pos_cur[e];
        const uint32_t slot = atomicAdd(&s_mig[down ? 0 : 1], 1u);
        if (slot < sh.mig_cap) {  // beyond the capacity the count alone reports the overflow
            uint2* rec = records_of(down ? sh.buf_down : sh.buf_up) + static_cast<size_t>(slot) * (MIGRANT_BYTES / 8);
            const float2 pp = a.pos_prev[e], t = a.target[e];
            const uint4 r = a.rng[e];
            const float4 c = a.color0[e];
            const uint32_t arrived = (a.arrived[arrived_word(e)] >> arrived_bit(e)) & 1u;
            rec[0] = make_uint2(__float_as_uint(p.x), __float_as_uint(p.y));
            rec[1] = make_uint2(__float_as_uint(pp.x), __float_as_uint(pp.y));
            rec[2] = make_uint2(__float_as_uint(t.x), __float_as_uint(t.y));
            rec[3] = make_uint2(r.x, r.y);
            rec[4] = make_uint2(r.z, r.w);
            rec[5] = make_uint2(__float_as_uint(c.x), __float_as_uint(c.y));
            rec[6] = make_uint2(__float_as_uint(c.z), __float_as_uint(c.w));
            rec[7] = make_uint2(a.road[e], a.gid[e]);
            rec[8] = make_uint2(arrived, 0u);
        }
        sh.local_ghosts[i] = p;  // it lands in the neighbour's boundary row: still within reach of ours
    }
    // barrier, then ONE thread writes the headers and fences at system scope (fences are cumulative over the barrier)
    // before it raises the flags
    __syncthreads();
    if (tid == 0) {
        sh.ctr[SHARD_CTR_LOCAL_GHOSTS] = k_out;
        if (sh.ctr[SHARD_CTR_HOLES] > sh.holes_cap) atomicOr(sh.error_word, 2u);
        if (sh.buf_down) {
            const uint32_t m = s_mig[0], hl = ld_volatile_u32(sh.ctr + SHARD_CTR_HALO_DOWN);
            ShardHeader* hd = header_of(sh.buf_down);
            hd->n_migrants = m;
            hd->n_halo = hl;
            hd->overflow = (m > sh.mig_cap || hl > sh.halo_cap) ? 1u : 0u;
        }
        if (sh.buf_up) {
            const uint32_t m = s_mig[1], hl = ld_volatile_u32(sh.ctr + SHARD_CTR_HALO_UP);
            ShardHeader* hd = header_of(sh.buf_up);
            hd->n_migrants = m;
            hd->n_halo = hl;
            hd->overflow = (m > sh.mig_cap || hl > sh.halo_cap) ? 1u : 0u;
        }
        __threadfence_system();
        if (sh.peer_flag_down) st_release_sys(sh.peer_flag_down, sh.signal_value);
        if (sh.peer_flag_up) st_release_sys(sh.peer_flag_up, sh.signal_value);
    }
}

__global__ void __launch_bounds__(EMIT_THREADS) shard_emit_kernel(ShardArrays a, ShardMoveArgs sh) { emit_body(a, sh); }

// ---- device-side integrate (asynchronous sharded tick) ------------------------------------------
// Same bookkeeping as msim_shard_integrate's host code, done by ONE CTA so that the tick needs no host
// round trip: arrivals fill the leavers' holes (then append), remaining holes are closed with the live
// entities of the tail, and the new owned / ghost / total counts are left in device memory for the
// kernels that follow on the stream.  Lists are short (hundreds to a few thousand entries per tick).
__device__ __forceinline__ void copy_entity(const ShardArrays& a, uint32_t src, uint32_t dst) {
    a.pos_cur[dst] = a.pos_cur[src];
    a.pos_prev[dst] = a.pos_prev[src];
    a.target[dst] = a.target[src];
    a.rng[dst] = a.rng[src];
    a.color0[dst] = a.color0[src];
    a.road[dst] = a.road[src];
    a.gid[dst] = a.gid[src];
    a.keys[dst] = a.keys[src];
    set_arrived_bit(a.arrived, dst, ((a.arrived[arrived_word(src)] >> arrived_bit(src)) & 1u) != 0u);
}

__device__ __forceinline__ void place_record(const ShardArrays& a, const void* buf, uint32_t r, uint32_t e, const GridParams& grid) {
    const uint2* rec = reinterpret_cast<const uint2*>(static_cast<const char*>(buf) + sizeof(ShardHeader)) + static_cast<size_t>(r) * (MIGRANT_BYTES / 8);
    // .cg: the buffer may have been written by another GPU (peer-memory exchange); L2 is the point of coherence
    const uint2 r0 = __ldcg(rec), r1 = __ldcg(rec + 1), r2 = __ldcg(rec + 2), r3 = __ldcg(rec + 3), r4 = __ldcg(rec + 4), r5 = __ldcg(rec + 5),
                r6 = __ldcg(rec + 6), r7 = __ldcg(rec + 7), r8 = __ldcg(rec + 8);
    const float2 p = make_float2(__uint_as_float(r0.x), __uint_as_float(r0.y));
    a.pos_cur[e] = p;
    a.pos_prev[e] = make_float2(__uint_as_float(r1.x), __uint_as_float(r1.y));
    a.target[e] = make_float2(__uint_as_float(r2.x), __uint_as_float(r2.y));
    a.rng[e] = make_uint4(r3.x, r3.y, r4.x, r4.y);
    a.color0[e] = make_float4(__uint_as_float(r5.x), __uint_as_float(r5.y), __uint_as_float(r6.x), __uint_as_float(r6.y));
    a.road[e] = r7.x;
    a.gid[e] = r7.y;
    const uint32_t key = cell_key_of(p, grid);
    a.keys[e] = key;
    count_cell(a, key);
    set_arrived_bit(a.arrived, e, (r8.x & 1u) != 0u);
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// spin until *flag >= expected; false on timeout (a neighbour that never enqueued its tick must not hang this GPU)
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t expected, unsigned long long timeout_ns) {
    const unsigned long long t0 = global_timer_ns();
    // signed distance: the counter may wrap after 2^32 ticks
    while (static_cast<int32_t>(ld_acquire_sys(flag) - expected) < 0) {
        if (global_timer_ns() - t0 > timeout_ns) return false;
        __nanosleep(200);
    }
    return true;
}

constexpr int INTEGRATE_THREADS = 1024;

struct IntegrateArgs {
    uint32_t* dev_counts;
    const void* sent_down;
    const void* sent_up;
    const void* recv_down;
    const void* recv_up;
    const uint32_t* holes;
    const uint32_t* ctr;
    const float2* local_ghosts;
    uint32_t mig_cap, halo_cap, holes_cap, entity_cap;
    uint32_t* tail_bits;
    uint2* moves;
};

__device__ __forceinline__ void integrate_body(const ShardArrays& a, const IntegrateArgs& ia, const GridParams& grid, const ShardWait& wait) {
    uint32_t* __restrict__ dev_counts = ia.dev_counts;
    const void* sent_down = ia.sent_down;
    const void* sent_up = ia.sent_up;
    const void* recv_down = ia.recv_down;
    const void* recv_up = ia.recv_up;
    const uint32_t* __restrict__ holes = ia.holes;
    const uint32_t* __restrict__ ctr = ia.ctr;
    const uint32_t mig_cap = ia.mig_cap, halo_cap = ia.halo_cap, holes_cap = ia.holes_cap, entity_cap = ia.entity_cap;
    uint32_t* __restrict__ tail_bits = ia.tail_bits;
    uint2* __restrict__ moves = ia.moves;
    __shared__ uint32_t s_n_old, s_n_new, s_k_out, s_in_down, s_in_up, s_ghosts, s_low, s_live, s_err;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        uint32_t err = 0;
        // peer-memory exchange: the neighbours' move kernels write into our receive buffers and then raise our flags
        if ((wait.flag_down || wait.flag_up) && !(dev_counts[DEV_SHARD_ERROR] & 32u)) {  // after one timeout nobody waits again
            if (wait.flag_down && !wait_flag(wait.flag_down, wait.expected, wait.timeout_ns)) err |= 32u;
            if (wait.flag_up && !(err & 32u) && !wait_flag(wait.flag_up, wait.expected, wait.timeout_ns)) err |= 32u;
        }
        if ((err | dev_counts[DEV_SHARD_ERROR]) & 32u) recv_down = recv_up = nullptr;  // timed out: this tick integrates nothing from outside
        const uint32_t n_old = dev_counts[DEV_N_OWNED];
        uint32_t k_out = ctr[SHARD_CTR_HOLES], g_local = ctr[SHARD_CTR_LOCAL_GHOSTS];
        uint32_t in_down = 0, in_up = 0, halo_down = 0, halo_up = 0;
        if (sent_down && static_cast<const ShardHeader*>(sent_down)->overflow) err |= 1u;
        if (sent_up && static_cast<const ShardHeader*>(sent_up)->overflow) err |= 1u;
        if (recv_down) {
            const uint32_t* hd = static_cast<const uint32_t*>(recv_down);  // ShardHeader {n_migrants, n_halo, overflow}
            in_down = __ldcg(hd);
            halo_down = __ldcg(hd + 1);
            if (__ldcg(hd + 2)) err |= 1u;
        }
        if (recv_up) {
            const uint32_t* hd = static_cast<const uint32_t*>(recv_up);
            in_up = __ldcg(hd);
            halo_up = __ldcg(hd + 1);
            if (__ldcg(hd + 2)) err |= 1u;
        }
        if (k_out > holes_cap || g_local > holes_cap) { err |= 2u; k_out = min(k_out, holes_cap); g_local = min(g_local, holes_cap); }
        if (in_down > mig_cap || in_up > mig_cap || halo_down > halo_cap || halo_up > halo_cap) {
            err |= 2u;
            in_down = min(in_down, mig_cap); in_up = min(in_up, mig_cap); halo_down = min(halo_down, halo_cap); halo_up = min(halo_up, halo_cap);
        }
        uint32_t n_new = n_old + in_down + in_up - min(k_out, n_old + in_down + in_up);
        uint32_t ghosts = halo_down + halo_up + g_local;
        if (static_cast<unsigned long long>(n_new) + ghosts > entity_cap) {  // keep every later kernel inside its arrays
            err |= 4u;
            if (n_new > entity_cap) n_new = entity_cap;
            ghosts = entity_cap - n_new;
        }
        // the ghost kernel that follows takes the (clamped) halo counts from here, not from the headers
        if (halo_down + halo_up + g_local > ghosts) {  // capacity error above: drop ghosts from the end
            uint32_t room = ghosts;
            halo_down = min(halo_down, room); room -= halo_down;
            halo_up = min(halo_up, room);
        }
        dev_counts[DEV_HALO_DOWN] = halo_down;
        dev_counts[DEV_HALO_UP] = halo_up;
        s_n_old = n_old; s_n_new = n_new; s_k_out = k_out; s_in_down = in_down; s_in_up = in_up; s_ghosts = ghosts;
        s_low = 0; s_live = 0; s_err = err;
    }
    __syncthreads();
    const uint32_t n_old = s_n_old, n_new = s_n_new, k_out = s_k_out, in_down = s_in_down, in_up = s_in_up, k_in = in_down + in_up;
    // 1. arrivals: holes first, then append
    for (uint32_t i = tid; i < k_in; i += INTEGRATE_THREADS) {
        const uint32_t dst = i < k_out ? holes[i] : n_old + (i - k_out);
        if (dst < entity_cap) place_record(a, i < in_down ? recv_down : recv_up, i < in_down ? i : i - in_down, dst, grid);
    }
    // 2. more leavers than arrivals: the tail [n_new, n_old) goes away; its live entities move into the open holes below n_new
    if (k_out > k_in) {
        const uint32_t tail = n_old - n_new;
        for (uint32_t w = tid; w < (tail + 31u) / 32u; w += INTEGRATE_THREADS) tail_bits[w] = 0;
        __syncthreads();
        for (uint32_t i = k_in + tid; i < k_out; i += INTEGRATE_THREADS) {
            const uint32_t hole = holes[i];
            if (hole >= n_new) atomicOr(&tail_bits[(hole - n_new) >> 5], 1u << ((hole - n_new) & 31u));
            else moves[atomicAdd(&s_low, 1u)].y = hole;
        }
        __syncthreads();
        for (uint32_t t = tid; t < tail; t += INTEGRATE_THREADS)
            if (!((tail_bits[t >> 5] >> (t & 31u)) & 1u)) moves[atomicAdd(&s_live, 1u)].x = n_new + t;
        __syncthreads();
        if (tid == 0 && s_low != s_live) s_err |= 8u;
        const uint32_t m = min(s_low, s_live);
        for (uint32_t i = tid; i < m; i += INTEGRATE_THREADS) copy_entity(a, moves[i].x, moves[i].y);
    }
    __syncthreads();
    if (tid == 0) {
        dev_counts[DEV_N_OWNED] = n_new;
        dev_counts[DEV_N_GHOST] = s_ghosts;
        dev_counts[DEV_N_TOTAL] = n_new + s_ghosts;
        if (s_err) atomicOr(&dev_counts[DEV_SHARD_ERROR], s_err);
    }
}

// ghosts behind the owned entities, counts taken from device memory
__device__ __forceinline__ void ghosts_body(const ShardArrays& a, const uint32_t* __restrict__ dev_counts, const void* recv_down, const void* recv_up,
                                            const float2* __restrict__ local_ghosts, uint32_t mig_cap, const GridParams& grid, uint32_t first_thread,
                                            uint32_t stride) {
    const uint32_t first = dev_counts[DEV_N_OWNED], ghosts = dev_counts[DEV_N_GHOST];
    const uint32_t h_down = dev_counts[DEV_HALO_DOWN], h_up = dev_counts[DEV_HALO_UP];  // validated by the integrate step
    for (uint32_t i = first_thread; i < ghosts; i += stride) {
        float2 p;
        if (i < h_down) p = __ldcg(halo_of(const_cast<void*>(recv_down), mig_cap) + i);
        else if (i < h_down + h_up) p = __ldcg(halo_of(const_cast<void*>(recv_up), mig_cap) + (i - h_down));
        else p = local_ghosts[i - h_down - h_up];
        a.pos_cur[first + i] = p;
        const uint32_t key = cell_key_of(p, grid);
        a.keys[first + i] = key;
        count_cell(a, key);
    }
}

__global__ void __launch_bounds__(INTEGRATE_THREADS) shard_integrate_kernel(ShardArrays a, IntegrateArgs ia, GridParams grid, ShardWait wait) {
    integrate_body(a, ia, grid, wait);
}

__global__ void __launch_bounds__(256)
shard_append_ghosts_device_kernel(ShardArrays a, const uint32_t* __restrict__ dev_counts, const void* recv_down, const void* recv_up,
                                  const float2* __restrict__ local_ghosts, uint32_t mig_cap, GridParams grid) {
    ghosts_body(a, dev_counts, recv_down, recv_up, local_ghosts, mig_cap, grid, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// Peer-memory exchange, the whole middle of a sharded tick in ONE single-CTA launch: publish what the move kernel left
// (emit: records, headers, flags), wait for both neighbours' flags, integrate their buffers, append the ghosts.  Three
// latency-bound launches became one; the ghost lists are a few thousand entries, well within one CTA's reach.
__global__ void __launch_bounds__(INTEGRATE_THREADS)
shard_exchange_kernel(ShardArrays a, ShardMoveArgs sh, int do_emit, IntegrateArgs ia, GridParams grid, ShardWait wait) {
    if (do_emit) {
        emit_body(a, sh);
        __syncthreads();
    }
    integrate_body(a, ia, grid, wait);
    __threadfence_block();
    __syncthreads();  // dev_counts written by thread 0 above
    ghosts_body(a, ia.dev_counts, ia.recv_down, ia.recv_up, ia.local_ghosts, ia.mig_cap, grid, threadIdx.x, INTEGRATE_THREADS);
}

__global__ void __launch_bounds__(256) shard_row_histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n, int ncx, uint32_t* __restrict__ rows) {
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
        atomicAdd(&rows[keys[e] / static_cast<uint32_t>(ncx)], 1u);
}

// raises the neighbours' flags after a pack that was not fused into a move kernel (the init-only first dispatch)
__global__ void shard_signal_kernel(uint32_t* flag_down, uint32_t* flag_up, uint32_t value) {
    __threadfence_system();
    if (flag_down) st_release_sys(flag_down, value);
    if (flag_up) st_release_sys(flag_up, value);
}

}  // namespace

int launch_shard_emit(cudaStream_t s, const ShardArrays& a, const ShardMoveArgs& sh, Profiler* prof) {
    prof->begin(s, K_SHARD);
    shard_emit_kernel<<<1, EMIT_THREADS, 0, s>>>(a, sh);
    prof->end(s);
    return 1;
}

int launch_shard_signal(cudaStream_t s, uint32_t* flag_down, uint32_t* flag_up, uint32_t value) {
    shard_signal_kernel<<<1, 1, 0, s>>>(flag_down, flag_up, value);
    return 1;
}

int launch_shard_reset(cudaStream_t s, void* buf_down, void* buf_up, uint32_t* ctr) {
    shard_reset_kernel<<<1, 32, 0, s>>>(buf_down, buf_up, ctr);
    return 1;
}

int launch_shard_pack(cudaStream_t s, const ShardArrays& a, uint32_t n, int ncx, uint32_t row_lo, uint32_t row_hi, void* buf_down, void* buf_up,
                      uint32_t mig_cap, uint32_t halo_cap, uint32_t* holes, uint32_t holes_cap, float2* local_ghosts, uint32_t* ctr, Profiler* prof,
                      const uint32_t* n_dev, bool reset) {
    if (reset) shard_reset_kernel<<<1, 32, 0, s>>>(buf_down, buf_up, ctr);
    if (n == 0) return 0;
    prof->begin(s, K_SHARD);
    shard_pack_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(a, n, n_dev, ncx, row_lo, row_hi, buf_down, buf_up, mig_cap, halo_cap, holes, holes_cap, local_ghosts, ctr);
    prof->end(s);
    return 1;
}

int launch_shard_place(cudaStream_t s, const ShardArrays& a, const void* recv_down, uint32_t n_down, const void* recv_up, uint32_t n_up,
                       const uint32_t* dst, const GridParams& grid, Profiler* prof) {
    if (n_down + n_up == 0) return 0;
    prof->begin(s, K_SHARD);
    shard_place_kernel<<<(n_down + n_up + 127u) / 128u, 128, 0, s>>>(a, recv_down, n_down, recv_up, n_up, dst, grid);
    prof->end(s);
    return 1;
}

int launch_shard_relocate(cudaStream_t s, const ShardArrays& a, const uint2* moves, uint32_t count, Profiler* prof) {
    if (count == 0) return 0;
    prof->begin(s, K_SHARD);
    shard_relocate_kernel<<<(count + 127u) / 128u, 128, 0, s>>>(a, moves, count);
    prof->end(s);
    return 1;
}

int launch_shard_append_ghosts(cudaStream_t s, const ShardArrays& a, uint32_t first, const void* recv_down, uint32_t h_down, const void* recv_up,
                               uint32_t h_up, const float2* local_ghosts, uint32_t h_local, uint32_t mig_cap, const GridParams& grid, Profiler* prof) {
    const uint32_t total = h_down + h_up + h_local;
    if (total == 0) return 0;
    prof->begin(s, K_SHARD);
    shard_append_ghosts_kernel<<<(total + 255u) / 256u, 256, 0, s>>>(a, first, recv_down, h_down, recv_up, h_up, local_ghosts, h_local, mig_cap, grid);
    prof->end(s);
    return 1;
}

namespace {
IntegrateArgs integrate_args(uint32_t* dev_counts, const void* sent_down, const void* sent_up, const void* recv_down, const void* recv_up,
                             const uint32_t* holes, const uint32_t* ctr, const float2* local_ghosts, uint32_t mig_cap, uint32_t halo_cap, uint32_t holes_cap,
                             uint32_t entity_cap, uint32_t* scratch_bits, uint2* scratch_moves) {
    IntegrateArgs ia{};
    ia.dev_counts = dev_counts;
    ia.sent_down = sent_down;
    ia.sent_up = sent_up;
    ia.recv_down = recv_down;
    ia.recv_up = recv_up;
    ia.holes = holes;
    ia.ctr = ctr;
    ia.local_ghosts = local_ghosts;
    ia.mig_cap = mig_cap;
    ia.halo_cap = halo_cap;
    ia.holes_cap = holes_cap;
    ia.entity_cap = entity_cap;
    ia.tail_bits = scratch_bits;
    ia.moves = scratch_moves;
    return ia;
}
}  // namespace

int launch_shard_integrate_device(cudaStream_t s, const ShardArrays& a, uint32_t* dev_counts, const void* sent_down, const void* sent_up,
                                  const void* recv_down, const void* recv_up, const uint32_t* holes, const uint32_t* ctr, const float2* local_ghosts,
                                  uint32_t mig_cap, uint32_t halo_cap, uint32_t holes_cap, uint32_t entity_cap, uint32_t* scratch_bits, uint2* scratch_moves,
                                  const GridParams& grid, Profiler* prof, const ShardWait* wait) {
    prof->begin(s, K_SHARD);
    const ShardWait none{};
    const IntegrateArgs ia = integrate_args(dev_counts, sent_down, sent_up, recv_down, recv_up, holes, ctr, local_ghosts, mig_cap, halo_cap, holes_cap,
                                            entity_cap, scratch_bits, scratch_moves);
    shard_integrate_kernel<<<1, INTEGRATE_THREADS, 0, s>>>(a, ia, grid, wait ? *wait : none);
    const uint32_t max_ghosts = 2u * halo_cap + holes_cap;
    shard_append_ghosts_device_kernel<<<(max_ghosts + 255u) / 256u, 256, 0, s>>>(a, dev_counts, recv_down, recv_up, local_ghosts, mig_cap, grid);
    prof->end(s);
    return 2;
}

int launch_shard_exchange(cudaStream_t s, const ShardArrays& a, const ShardMoveArgs* emit, uint32_t* dev_counts, const void* recv_down, const void* recv_up,
                          const uint32_t* holes, const uint32_t* ctr, const float2* local_ghosts, uint32_t mig_cap, uint32_t halo_cap, uint32_t holes_cap,
                          uint32_t entity_cap, uint32_t* scratch_bits, uint2* scratch_moves, const GridParams& grid, Profiler* prof, const ShardWait& wait) {
    prof->begin(s, K_SHARD);
    const ShardMoveArgs none{};
    const IntegrateArgs ia = integrate_args(dev_counts, nullptr, nullptr, recv_down, recv_up, holes, ctr, local_ghosts, mig_cap, halo_cap, holes_cap,
                                            entity_cap, scratch_bits, scratch_moves);
    shard_exchange_kernel<<<1, INTEGRATE_THREADS, 0, s>>>(a, emit ? *emit : none, emit ? 1 : 0, ia, grid, wait);
    prof->end(s);
    return 1;
}

int launch_shard_row_histogram(cudaStream_t s, const uint32_t* keys, uint32_t n, int ncx, uint32_t* rows, uint32_t nrows, Profiler* prof) {
    cudaMemsetAsync(rows, 0, static_cast<size_t>(nrows) * sizeof(uint32_t), s);
    if (n == 0) return 0;
    uint32_t blocks = (n + 255u) / 256u;
    if (blocks > 1184u) blocks = 1184u;
    prof->begin(s, K_SHARD);
    shard_row_histogram_kernel<<<blocks, 256, 0, s>>>(keys, n, ncx, rows);
    prof->end(s);
    return 1;
}

}  // namespace msim
