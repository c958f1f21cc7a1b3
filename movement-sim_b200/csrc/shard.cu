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

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

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

// spin until *flag >= expected; false on timeout (a neighbour that never enqueued its tick must not hang this GPU).
// The flag lives in THIS GPU's memory and is written by the neighbour over NVLink behind a system-scope fence: by the time the flag
// write reaches our L2, everything the neighbour stored before it has.  It is polled with plain volatile loads (L2 is the point of
// coherence for peer writes into local memory) and the data behind it is read with .cg loads after a device-scope fence.  A
// system-scope acquire here was measured to make the kernel COMPLETE ~16 us late (any kernel that executes a system-scope fence
// does: 13 us between such a kernel and the next one on its stream), and this kernel is on the tick's critical path.
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t expected, unsigned long long timeout_ns) {
    const unsigned long long t0 = global_timer_ns();
    // signed distance: the counter may wrap after 2^32 ticks
    while (static_cast<int32_t>(ld_volatile_u32(flag) - expected) < 0) {
        if (global_timer_ns() - t0 > timeout_ns) return false;
        __nanosleep(100);
    }
    __threadfence();
    return true;
}

constexpr int INTEGRATE_THREADS = 1024;
constexpr int EXCHANGE_CTAS = 8;  // merged exchange kernel: CTA 0 places the arrivals and closes the holes, all of them share the ghost lists

// The element counts live in device memory in TWO sets that alternate from one integrate to the next: every kernel of a tick
// reads the set the tick's integrate wrote, the next integrate reads that set and writes the other one.  A CTA of the multi-CTA
// exchange kernel that starts late therefore still finds the counts it has to start from.  Error bits are sticky in one word.
struct IntegrateArgs {
    const uint32_t* counts_in;
    uint32_t* counts_out;
    uint32_t* error_word;
    const void* sent_down;
    const void* sent_up;
    const void* recv_down;
    const void* recv_up;
    const uint32_t* holes;
    const uint32_t* ctr;
    uint32_t* ctr_rw;  // the same counters, for the one word the exchange kernel writes (SHARD_CTR_COMPACTED)
    const float2* local_ghosts;
    uint32_t mig_cap, halo_cap, holes_cap, entity_cap;
    uint32_t* tail_bits;
    uint2* moves;
};

// what one integrate works on: a pure function of the counts before it, the counters the move kernel left and the four buffer
// headers - every CTA of the exchange kernel derives the same values for itself
struct ExchangePlan {
    uint32_t n_old, n_new, k_out, in_down, in_up, halo_down, halo_up, ghosts, err;
    const void* recv_down;
    const void* recv_up;
};

// thread 0 of a CTA: wait for the neighbours (peer-memory exchange), read the headers, validate and clamp
__device__ __forceinline__ void plan_exchange(const IntegrateArgs& ia, const ShardWait& wait, unsigned long long* trace, ExchangePlan& pl) {
    const uint32_t mig_cap = ia.mig_cap, halo_cap = ia.halo_cap, holes_cap = ia.holes_cap, entity_cap = ia.entity_cap;
    const void* recv_down = ia.recv_down;
    const void* recv_up = ia.recv_up;
    uint32_t err = 0;
    const uint32_t sticky = ld_volatile_u32(ia.error_word);
    // peer-memory exchange: the neighbours' move kernels write into our receive buffers and then raise our flags
    if ((wait.flag_down || wait.flag_up) && !(sticky & 32u)) {  // after one timeout nobody waits again
        if (wait.flag_down && !wait_flag(wait.flag_down, wait.expected, wait.timeout_ns)) err |= 32u;
        if (wait.flag_up && !(err & 32u) && !wait_flag(wait.flag_up, wait.expected, wait.timeout_ns)) err |= 32u;
    }
    if (trace) {
        const unsigned long long t = global_timer_ns();
        trace[1] += t - trace[6];
        trace[7] = t;
    }
    if ((err | sticky) & 32u) recv_down = recv_up = nullptr;  // timed out: this tick integrates nothing from outside
    const uint32_t n_old = ia.counts_in[DEV_N_OWNED];
    uint32_t k_out = __ldcg(ia.ctr + SHARD_CTR_HOLES), g_local = k_out;  // every leaver left a hole and a local ghost
    uint32_t in_down = 0, in_up = 0, halo_down = 0, halo_up = 0;
    if (ia.sent_down && static_cast<const ShardHeader*>(ia.sent_down)->overflow) err |= 1u;
    if (ia.sent_up && static_cast<const ShardHeader*>(ia.sent_up)->overflow) err |= 1u;
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
    if (halo_down + halo_up + g_local > ghosts) {  // capacity error above: drop ghosts from the end
        uint32_t room = ghosts;
        halo_down = min(halo_down, room); room -= halo_down;
        halo_up = min(halo_up, room);
    }
    pl.n_old = n_old; pl.n_new = n_new; pl.k_out = k_out; pl.in_down = in_down; pl.in_up = in_up;
    pl.halo_down = halo_down; pl.halo_up = halo_up; pl.ghosts = ghosts; pl.err = err;
    pl.recv_down = recv_down; pl.recv_up = recv_up;
}

// Same bookkeeping as msim_shard_integrate's host code, done by ONE CTA so that the tick needs no host round trip: arrivals fill
// the leavers' holes (then append), remaining holes are closed with the live entities of the tail, and the new owned / ghost /
// total counts are left in device memory for the kernels that follow on the stream.  Lists are short (hundreds to a few thousand
// entries per tick).  `pl` lives in shared memory and was filled by this CTA's thread 0.
__device__ __forceinline__ void integrate_body(const ShardArrays& a, const IntegrateArgs& ia, const GridParams& grid, const ExchangePlan& pl) {
    const uint32_t* __restrict__ holes = ia.holes;
    const uint32_t entity_cap = ia.entity_cap;
    uint32_t* __restrict__ tail_bits = ia.tail_bits;
    uint2* __restrict__ moves = ia.moves;
    __shared__ uint32_t s_low, s_live, s_err;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) { s_low = 0; s_live = 0; s_err = pl.err; }
    __syncthreads();
    const uint32_t n_old = pl.n_old, n_new = pl.n_new, k_out = pl.k_out, in_down = pl.in_down, k_in = pl.in_down + pl.in_up;
    // 1. arrivals: holes first, then append
    for (uint32_t i = tid; i < k_in; i += INTEGRATE_THREADS) {
        const uint32_t dst = i < k_out ? __ldcg(holes + i) : n_old + (i - k_out);
        if (dst < entity_cap) place_record(a, i < in_down ? pl.recv_down : pl.recv_up, i < in_down ? i : i - in_down, dst, grid);
    }
    // 2. more leavers than arrivals: the tail [n_new, n_old) goes away; its live entities move into the open holes below n_new
    if (k_out > k_in) {
        const uint32_t tail = n_old - n_new;
        for (uint32_t w = tid; w < (tail + 31u) / 32u; w += INTEGRATE_THREADS) tail_bits[w] = 0;
        __syncthreads();
        for (uint32_t i = k_in + tid; i < k_out; i += INTEGRATE_THREADS) {
            const uint32_t hole = __ldcg(holes + i);
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
        ia.counts_out[DEV_N_OWNED] = n_new;
        ia.counts_out[DEV_N_GHOST] = pl.ghosts;
        ia.counts_out[DEV_N_TOTAL] = n_new + pl.ghosts;
        ia.counts_out[DEV_HALO_DOWN] = pl.halo_down;  // (clamped: the stand-alone ghost kernel takes them from here, not from the headers)
        ia.counts_out[DEV_HALO_UP] = pl.halo_up;
        if (s_err) atomicOr(ia.error_word, s_err);
    }
}

// ghosts = halo from below + halo from above + our own leavers, behind the owned entities
__device__ __forceinline__ void ghosts_body(const ShardArrays& a, uint32_t first, uint32_t ghosts, uint32_t h_down, uint32_t h_up, const void* recv_down,
                                            const void* recv_up, const float2* __restrict__ local_ghosts, uint32_t mig_cap, const GridParams& grid,
                                            uint32_t first_thread, uint32_t stride) {
    for (uint32_t i = first_thread; i < ghosts; i += stride) {
        float2 p;
        if (i < h_down) p = __ldcg(halo_of(const_cast<void*>(recv_down), mig_cap) + i);
        else if (i < h_down + h_up) p = __ldcg(halo_of(const_cast<void*>(recv_up), mig_cap) + (i - h_down));
        else p = __ldcg(local_ghosts + (i - h_down - h_up));
        a.pos_cur[first + i] = p;
        const uint32_t key = cell_key_of(p, grid);
        a.keys[first + i] = key;
        count_cell(a, key);
    }
}

__global__ void __launch_bounds__(INTEGRATE_THREADS) shard_integrate_kernel(ShardArrays a, IntegrateArgs ia, GridParams grid, ShardWait wait) {
    __shared__ ExchangePlan s_plan;
    if (threadIdx.x == 0) plan_exchange(ia, wait, nullptr, s_plan);
    __syncthreads();
    integrate_body(a, ia, grid, s_plan);
}

// (behind shard_integrate_kernel on the stream: counts and clamped halo lengths come from the set it wrote)
__global__ void __launch_bounds__(256)
shard_append_ghosts_device_kernel(ShardArrays a, const uint32_t* __restrict__ counts, const void* recv_down, const void* recv_up,
                                  const float2* __restrict__ local_ghosts, uint32_t mig_cap, GridParams grid) {
    ghosts_body(a, counts[DEV_N_OWNED], counts[DEV_N_GHOST], counts[DEV_HALO_DOWN], counts[DEV_HALO_UP], recv_down, recv_up, local_ghosts, mig_cap, grid,
                blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// Peer-memory exchange, the whole middle of a sharded tick in ONE launch of a few CTAs: wait for both neighbours' flags (their
// move kernels raised them with their last CTA), integrate their buffers, append the ghosts.  Every CTA makes the same plan from the
// same inputs, so nothing is handed from CTA to CTA: CTA 0 places the arrivals and closes the holes while all of them, CTA 0 last,
// work through their share of the ghost lists (a few thousand entries: the serial part of the old single-CTA kernel).
// `trace` (MSIM_SHARD_TRACE=1, else NULL): nanoseconds CTA 0 spent waiting for the flags / integrating / on ghosts, summed over launches, + launch count
__global__ void __launch_bounds__(INTEGRATE_THREADS)
shard_exchange_kernel(ShardArrays a, IntegrateArgs ia, GridParams grid, ShardWait wait, unsigned long long* trace) {
    __shared__ ExchangePlan s_plan;
    const bool lead = blockIdx.x == 0;
    unsigned long long* const trace_all = trace;
    if (!lead) trace = nullptr;
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) {
        if (trace) {
            trace[6] = global_timer_ns();
            if (trace[5]) trace[0] += trace[6] - trace[5];  // gap between the stamp kernel behind the move kernel and the start of this one
        }
        plan_exchange(ia, wait, trace, s_plan);
    }
    __syncthreads();
    // the ghosts go where the tail was when more entities left than arrived: those slots are free once CTA 0 has moved the tail's
    // live entities into the holes (rare and short: a net loss of a few entities)
    const bool tail_in_the_way = s_plan.n_new < s_plan.n_old;
    if (lead) {
        integrate_body(a, ia, grid, s_plan);
        if (tail_in_the_way) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicExch(ia.ctr_rw + SHARD_CTR_COMPACTED, 1u);
        }
        if (trace && threadIdx.x == 0) {
            t0 = global_timer_ns();
            trace[2] += t0 - trace[7];
        }
    } else if (tail_in_the_way) {
        if (threadIdx.x == 0) {
            const unsigned long long w0 = global_timer_ns();
            while (ld_volatile_u32(ia.ctr_rw + SHARD_CTR_COMPACTED) == 0u && global_timer_ns() - w0 < wait.timeout_ns) __nanosleep(100);
            __threadfence();
        }
        __syncthreads();
    }
    ghosts_body(a, s_plan.n_new, s_plan.ghosts, s_plan.halo_down, s_plan.halo_up, s_plan.recv_down, s_plan.recv_up, ia.local_ghosts, ia.mig_cap, grid,
                blockIdx.x * INTEGRATE_THREADS + threadIdx.x, gridDim.x * INTEGRATE_THREADS);
    if (trace_all) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long t1 = global_timer_ns();
            atomicMax(trace_all + 8, t1);  // when the last CTA was done (the stamp kernel behind this one turns it into two sums)
            if (lead) {
                trace[3] += t1 - t0;
                trace[4] += 1;
            }
        }
    }
}

// ---- peer-memory exchange, sender side ------------------------------------------------------------------------------------------
// The fused move + pack kernel fills LOCAL send buffers; this kernel, on a stream of its own behind it, copies what was filled into
// the neighbours' receive buffers over NVLink (plain 8-byte stores), writes the headers, fences at system scope and raises the
// neighbours' flags.  Remote stores are kept out of the kernels of the main stream on purpose: a kernel that has written peer memory
// completes ~13 us late (measured: the gap between such a kernel and the next one on its stream), and the tick has that stream on
// its critical path - this one it has not.  Counter words are read from the local counters the move kernel left.
constexpr int PUSH_THREADS = 1024;
constexpr int PUSH_CTAS = 8;

__device__ __forceinline__ void push_bytes(void* dst, const void* src, size_t bytes, uint32_t first, uint32_t stride) {
    // 8-byte words: records are 72 bytes and the halo list starts behind mig_cap of them, so 8 is the alignment both lists always have
    const uint2* s = static_cast<const uint2*>(src);
    uint2* d = static_cast<uint2*>(dst);
    const size_t n = bytes / 8u;
    for (size_t i = first; i < n; i += stride) d[i] = __ldcg(s + i);
}

struct PushArgs {
    const void* send_down;  // local, filled by the move kernel (NULL: no neighbour on that side)
    const void* send_up;
    void* peer_down;        // the neighbours' receive buffers
    void* peer_up;
    uint32_t* flag_down;    // the neighbours' flag words
    uint32_t* flag_up;
    uint32_t signal_value;
    uint32_t mig_cap, halo_cap, holes_cap;
    uint32_t* ctr;          // local counters of this pack
    uint32_t* error_word;
    uint32_t* ticket;       // CTAs of this launch that are done (cleared by the last one)
    uint32_t counts_in_headers;  // 1: the lists were filled by the stand-alone pack kernel, which counts in the send buffers' headers
};

__global__ void __launch_bounds__(PUSH_THREADS) shard_push_kernel(PushArgs a) {
    __shared__ uint32_t s_last;
    const uint32_t first = blockIdx.x * PUSH_THREADS + threadIdx.x, stride = gridDim.x * PUSH_THREADS;
    const uint32_t holes_total = __ldcg(a.ctr + SHARD_CTR_HOLES);
    const size_t rec0 = sizeof(ShardHeader), halo0 = sizeof(ShardHeader) + static_cast<size_t>(a.mig_cap) * MIGRANT_BYTES;
    uint32_t m_down = 0, h_down = 0, m_up = 0, h_up = 0;
    if (a.send_down) {
        const uint32_t* hd = static_cast<const uint32_t*>(a.send_down);  // ShardHeader {n_migrants, n_halo, overflow}
        m_down = __ldcg(a.counts_in_headers ? hd : a.ctr + SHARD_CTR_MIG_DOWN);
        h_down = __ldcg(a.counts_in_headers ? hd + 1 : a.ctr + SHARD_CTR_HALO_DOWN);
        push_bytes(static_cast<char*>(a.peer_down) + rec0, static_cast<const char*>(a.send_down) + rec0, static_cast<size_t>(min(m_down, a.mig_cap)) * MIGRANT_BYTES, first, stride);
        push_bytes(static_cast<char*>(a.peer_down) + halo0, static_cast<const char*>(a.send_down) + halo0, static_cast<size_t>(min(h_down, a.halo_cap)) * sizeof(float2), first, stride);
    }
    if (a.send_up) {
        const uint32_t* hd = static_cast<const uint32_t*>(a.send_up);
        m_up = __ldcg(a.counts_in_headers ? hd : a.ctr + SHARD_CTR_MIG_UP);
        h_up = __ldcg(a.counts_in_headers ? hd + 1 : a.ctr + SHARD_CTR_HALO_UP);
        push_bytes(static_cast<char*>(a.peer_up) + rec0, static_cast<const char*>(a.send_up) + rec0, static_cast<size_t>(min(m_up, a.mig_cap)) * MIGRANT_BYTES, first, stride);
        push_bytes(static_cast<char*>(a.peer_up) + halo0, static_cast<const char*>(a.send_up) + halo0, static_cast<size_t>(min(h_up, a.halo_cap)) * sizeof(float2), first, stride);
    }
    __syncthreads();  // this CTA's remote stores ...
    if (threadIdx.x == 0) {
        __threadfence_system();  // ... are ordered before its ticket by one cumulative fence
        s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    *a.ticket = 0u;
    if (holes_total > a.holes_cap) atomicOr(a.error_word, 2u);
    if (a.send_down) {
        ShardHeader* hd = header_of(a.peer_down);
        hd->n_migrants = m_down;
        hd->n_halo = h_down;
        hd->overflow = (m_down > a.mig_cap || h_down > a.halo_cap) ? 1u : 0u;
    }
    if (a.send_up) {
        ShardHeader* hd = header_of(a.peer_up);
        hd->n_migrants = m_up;
        hd->n_halo = h_up;
        hd->overflow = (m_up > a.mig_cap || h_up > a.halo_cap) ? 1u : 0u;
    }
    __threadfence_system();  // headers (and, cumulatively, what the tickets ordered before this thread) before the flags
    if (a.flag_down) st_relaxed_sys(a.flag_down, a.signal_value);
    if (a.flag_up) st_relaxed_sys(a.flag_up, a.signal_value);
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

int launch_shard_push(cudaStream_t s, const void* send_down, const void* send_up, void* peer_down, void* peer_up, uint32_t* flag_down, uint32_t* flag_up,
                      uint32_t signal_value, uint32_t mig_cap, uint32_t halo_cap, uint32_t holes_cap, uint32_t* ctr, uint32_t* error_word, uint32_t* ticket,
                      bool counts_in_headers) {
    PushArgs a{};
    a.send_down = send_down; a.send_up = send_up; a.peer_down = peer_down; a.peer_up = peer_up; a.flag_down = flag_down; a.flag_up = flag_up;
    a.signal_value = signal_value; a.mig_cap = mig_cap; a.halo_cap = halo_cap; a.holes_cap = holes_cap; a.ctr = ctr; a.error_word = error_word; a.ticket = ticket;
    a.counts_in_headers = counts_in_headers ? 1u : 0u;
    shard_push_kernel<<<PUSH_CTAS, PUSH_THREADS, 0, s>>>(a);
    return 1;
}

__global__ void shard_stamp_kernel(unsigned long long* trace, int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[slot] = t;
    if (slot == 9 && trace[8]) {  // behind the exchange kernel: start -> last CTA done, last CTA done -> this kernel
        trace[10] += trace[8] - trace[6];
        trace[11] += t - trace[8];
        trace[8] = 0;
    }
}
void launch_shard_stamp(cudaStream_t s, unsigned long long* trace, int slot) { shard_stamp_kernel<<<1, 1, 0, s>>>(trace, slot); }

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
IntegrateArgs integrate_args(const ShardCounts& counts, const void* sent_down, const void* sent_up, const void* recv_down, const void* recv_up,
                             const uint32_t* holes, const uint32_t* ctr, const float2* local_ghosts, uint32_t mig_cap, uint32_t halo_cap, uint32_t holes_cap,
                             uint32_t entity_cap, uint32_t* scratch_bits, uint2* scratch_moves) {
    IntegrateArgs ia{};
    ia.counts_in = counts.in;
    ia.counts_out = counts.out;
    ia.error_word = counts.error_word;
    ia.sent_down = sent_down;
    ia.sent_up = sent_up;
    ia.recv_down = recv_down;
    ia.recv_up = recv_up;
    ia.holes = holes;
    ia.ctr = ctr;
    ia.ctr_rw = const_cast<uint32_t*>(ctr);
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

int launch_shard_integrate_device(cudaStream_t s, const ShardArrays& a, const ShardCounts& counts, const void* sent_down, const void* sent_up,
                                  const void* recv_down, const void* recv_up, const uint32_t* holes, const uint32_t* ctr, const float2* local_ghosts,
                                  uint32_t mig_cap, uint32_t halo_cap, uint32_t holes_cap, uint32_t entity_cap, uint32_t* scratch_bits, uint2* scratch_moves,
                                  const GridParams& grid, Profiler* prof, const ShardWait* wait) {
    prof->begin(s, K_SHARD);
    const ShardWait none{};
    const IntegrateArgs ia = integrate_args(counts, sent_down, sent_up, recv_down, recv_up, holes, ctr, local_ghosts, mig_cap, halo_cap, holes_cap,
                                            entity_cap, scratch_bits, scratch_moves);
    shard_integrate_kernel<<<1, INTEGRATE_THREADS, 0, s>>>(a, ia, grid, wait ? *wait : none);
    const uint32_t max_ghosts = 2u * halo_cap + holes_cap;
    shard_append_ghosts_device_kernel<<<(max_ghosts + 255u) / 256u, 256, 0, s>>>(a, counts.out, recv_down, recv_up, local_ghosts, mig_cap, grid);
    prof->end(s);
    return 2;
}

int launch_shard_exchange(cudaStream_t s, const ShardArrays& a, const ShardCounts& counts, const void* recv_down, const void* recv_up,
                          const uint32_t* holes, const uint32_t* ctr, const float2* local_ghosts, uint32_t mig_cap, uint32_t halo_cap, uint32_t holes_cap,
                          uint32_t entity_cap, uint32_t* scratch_bits, uint2* scratch_moves, const GridParams& grid, Profiler* prof, const ShardWait& wait,
                          unsigned long long* trace) {
    prof->begin(s, K_SHARD);
    const IntegrateArgs ia = integrate_args(counts, nullptr, nullptr, recv_down, recv_up, holes, ctr, local_ghosts, mig_cap, halo_cap, holes_cap,
                                            entity_cap, scratch_bits, scratch_moves);
    shard_exchange_kernel<<<EXCHANGE_CTAS, INTEGRATE_THREADS, 0, s>>>(a, ia, grid, wait, trace);
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
