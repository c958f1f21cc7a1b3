// collide_common.cuh - device helpers of the collision query shared by collide.cu (query_kernel) and collide_tiles.cu (query_tiles_kernel): distance test, candidate scans over global / shared memory, cell-run look-up, 1-D TMA bulk copy.
// Included inside namespace msim { namespace { ... } } of the including translation unit.
#pragma once

// squared distance with individually rounded operations, as the oracle computes it.  The x and y
// lanes go through Blackwell's packed binary32 pipes (FADD2 / FMUL2, sm_100+): same IEEE round-to-nearest
// result per lane as two scalar instructions, half the issue slots — the query kernel is issue-bound.
#ifdef MSIM_HOST_EMU  // tests/cuda_emu: the same individually rounded operations without the packed PTX forms
__device__ __forceinline__ float dist2(float2 a, float2 b) {
    const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}
#else
__device__ __forceinline__ float dist2(float2 a, float2 b) {
    unsigned long long ua, ub, d, sq;
    ua = (static_cast<unsigned long long>(__float_as_uint(a.y)) << 32) | __float_as_uint(a.x);
    ub = (static_cast<unsigned long long>(__float_as_uint(b.y)) << 32) | __float_as_uint(b.x);
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(ua), "l"(ub));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(sq) : "l"(d), "l"(d));
    return __fadd_rn(__uint_as_float(static_cast<uint32_t>(sq)), __uint_as_float(static_cast<uint32_t>(sq >> 32)));
}
#endif

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

// ---- the look-above scan, served by the whole warp ------------------------------------------------------------------------------
// `need`: this lane found nothing among the slots below it and wants to know whether some slot of [up0, up1) (rest of its own row) or
// [dn0, dn1) (row below) is in range.  Few lanes do, so they are served one at a time, TOGETHER: a ballot, the leader's position and runs
// broadcast by shuffles, every lane tests another candidate of the concatenated runs, a vote ends the loop.  Two straight-line tests of the
// next two slots come first (most such lanes have their neighbour right behind them in their own cell).  Control flow is warp-uniform on
// purpose: per-lane divergent scans here were miscompiled into sharing one uniform register between paths (profiles/r2_flag_count_race.md).
// Must be called by all 32 lanes of a converged warp.
__device__ __forceinline__ void look_above_cooperative(bool need, float2 p, uint32_t up0, uint32_t up1, uint32_t dn0, uint32_t dn1,
                                                       const float2* __restrict__ sorted_pos, float thr, uint32_t lane, bool& hit) {
    uint32_t todo = __ballot_sync(0xffffffffu, need);
    if (todo) {  // (warp-uniform)
        bool near = false;
        if (need && up0 < up1) near = dist2(__ldg(sorted_pos + up0), p) < thr;
        if (need && up0 + 1u < up1) near = near || dist2(__ldg(sorted_pos + up0 + 1u), p) < thr;
        if (near) hit = true;
        todo = __ballot_sync(0xffffffffu, need && !near);
        up0 = min(up0 + 2u, up1);
    }
    while (todo) {  // warp-uniform
        const int leader = __ffs(todo) - 1;
        todo &= todo - 1u;
        const float2 lp = make_float2(__shfl_sync(0xffffffffu, p.x, leader), __shfl_sync(0xffffffffu, p.y, leader));
        const uint32_t a0 = __shfl_sync(0xffffffffu, up0, leader), a1 = __shfl_sync(0xffffffffu, up1, leader);
        const uint32_t b0 = __shfl_sync(0xffffffffu, dn0, leader), b1 = __shfl_sync(0xffffffffu, dn1, leader);
        const uint32_t na = a1 - a0, total = na + (b1 - b0);
        bool found = false;
        for (uint32_t base = 0; base < total && !found; base += 32u) {  // warp-uniform: `found` is a vote
            const uint32_t k = base + lane;
            bool h = false;
            if (k < total) h = dist2(__ldg(sorted_pos + (k < na ? a0 + k : b0 + (k - na))), lp) < thr;
            found = __any_sync(0xffffffffu, h);
        }
        if (static_cast<int>(lane) == leader) hit = found;
    }
}

// ---- shared-memory variants of the scans (same arithmetic, candidates already staged) -----------
__device__ __forceinline__ uint32_t count_in_tile(const float2* __restrict__ tile, uint32_t a, uint32_t b, float2 p, float threshold) {
    uint32_t c = 0;
    uint32_t k = a;
    for (; k + 4 <= b; k += 4) {
        const float2 q0 = tile[k], q1 = tile[k + 1], q2 = tile[k + 2], q3 = tile[k + 3];
        c += (dist2(q0, p) < threshold) ? 1u : 0u;
        c += (dist2(q1, p) < threshold) ? 1u : 0u;
        c += (dist2(q2, p) < threshold) ? 1u : 0u;
        c += (dist2(q3, p) < threshold) ? 1u : 0u;
    }
    for (; k < b; k++) c += (dist2(tile[k], p) < threshold) ? 1u : 0u;
    return c;
}

__device__ __forceinline__ bool any_in_tile(const float2* __restrict__ tile, uint32_t a, uint32_t b, float2 p, float threshold) {
    uint32_t k = a;
    for (; k + 4 <= b; k += 4) {
        const float2 q0 = tile[k], q1 = tile[k + 1], q2 = tile[k + 2], q3 = tile[k + 3];
        if ((dist2(q0, p) < threshold) | (dist2(q1, p) < threshold) | (dist2(q2, p) < threshold) | (dist2(q3, p) < threshold)) return true;
    }
    for (; k < b; k++)
        if (dist2(tile[k], p) < threshold) return true;
    return false;
}

// ---- bulk asynchronous copy (TMA, 1-D) of a contiguous window of sorted_pos into shared memory ----------------
// One elected thread arms an mbarrier with the byte count and issues cp.async.bulk; the copy engine fills the window while
// no thread spends issue slots on LDG + STS (the query kernel is issue-bound: the two staging loops were ~10 % of its
// instructions).  Source, destination and size must be multiples of 16 bytes: windows are widened to even slot indices.
#ifdef MSIM_HOST_EMU  // tests/cuda_emu: the elected thread copies at once; waiting on the barrier is a barrier over the block
__device__ __forceinline__ void mbar_init(unsigned long long*, uint32_t) {}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long*, uint32_t) {}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_global, uint32_t bytes, unsigned long long*) { memcpy(dst_smem, src_global, bytes); }
__device__ __forceinline__ void mbar_wait(unsigned long long*, uint32_t) { __syncthreads(); }
#else
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_global, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)), "l"(src_global),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    do {
        // the hint lets the hardware park the warp until the phase flips instead of re-issuing the poll (4 % of the issue slots before)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(smem_addr(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}
#endif

constexpr int QUERY_THREADS = 256;
constexpr int COUNTER_STRIPES = 64;
constexpr int COUNTER_STRIDE = 16;  // in u64 words: 128 bytes between stripes
constexpr uint32_t QUERY_WINDOW = 1728;  // candidates staged per window: 2 x 13.5 KB, so that 8 CTAs (64 warps) fit one SM with the 1 KB per-CTA reserve

