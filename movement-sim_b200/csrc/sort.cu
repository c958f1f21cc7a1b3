// sort.cu — LSD radix sort of (cell key, entity index) pairs: the per-tick rebuild of the neighbour
// structure that replaces the reference's lock-based incremental quadtree
// (/root/reference/src/sim/shader/random_move.comp:100-539: insert / split / merge / update under
// per-node spin locks).  Nothing is ported from there; only the observable result — who is whose
// neighbour — is preserved.
//
// Algorithm: one histogram kernel for all digits, then one "onesweep" kernel per 8-bit digit:
// every CTA ranks a 4096-pair tile with warp-level match_any + per-warp digit counters, resolves its
// global offsets with a decoupled look-back over per-(tile, digit) status words, reorders the tile
// through shared memory and writes digit runs back coalesced.  HBM traffic per pass is the minimum
// for an out-of-place pass: read 8 B + write 8 B per pair (the first pass reads bare 4-byte keys —
// the index is the position).  Tile ids come from an atomic ticket so that a tile can only wait on
// tiles that are already resident: the look-back cannot deadlock, and a watchdog bounds every spin.
#include "msim_internal.h"

namespace msim {
namespace {

constexpr uint32_t FLAG_AGGREGATE = 1u << 30;
constexpr uint32_t FLAG_INCLUSIVE = 2u << 30;
constexpr uint32_t VALUE_MASK = (1u << 30) - 1u;
constexpr uint32_t WATCHDOG_SPINS = 1u << 22;

__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- histogram of every digit in one read of the keys -----------------------------------------
__global__ void __launch_bounds__(256) histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n_host, const uint32_t* __restrict__ n_dev, uint32_t* __restrict__ hist, int passes) {
    __shared__ uint32_t sh[MAX_SORT_PASSES * RADIX];
    const uint32_t n = n_dev ? *n_dev : n_host;
    for (int i = threadIdx.x; i < MAX_SORT_PASSES * RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t quads = n >> 2;
    const uint4* keys4 = reinterpret_cast<const uint4*>(keys);
    for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
        const uint4 k = __ldg(keys4 + q);
        const uint32_t v[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
        for (int j = 0; j < 4; j++)
            for (int p = 0; p < passes; p++) atomicAdd(&sh[p * RADIX + ((v[j] >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3u)) {  // tail
        const uint32_t v = keys[(quads << 2) + threadIdx.x];
        for (int p = 0; p < passes; p++) atomicAdd(&sh[p * RADIX + ((v >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}

// exclusive scan of one value per thread across a 256-thread CTA
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_totals /* [SORT_WARPS] */) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= static_cast<uint32_t>(d)) incl += up;
    }
    if (lane == 31) warp_totals[warp] = incl;
    __syncthreads();
    uint32_t warp_base = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++)
        if (static_cast<uint32_t>(w) < warp) warp_base += warp_totals[w];
    __syncthreads();  // warp_totals may be reused by the caller
    return warp_base + incl - v;
}

// ---- one onesweep pass ------------------------------------------------------------------------
// FIRST: input is the bare u32 key array, the payload (entity index) is the element's position.
template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS, 3) onesweep_kernel(const void* __restrict__ in_raw, uint64_t* __restrict__ out, uint32_t n_host, const uint32_t* __restrict__ n_dev, int shift,
                                                                const uint32_t* __restrict__ hist /* [RADIX], this pass */,
                                                                uint32_t* __restrict__ tile_state /* [tiles][RADIX] */,
                                                                uint32_t* __restrict__ tile_counter, uint32_t* __restrict__ error_flag) {
    __shared__ uint32_t s_warp_hist[SORT_WARPS][RADIX];  // per-warp digit counts -> per-warp exclusive prefixes
    __shared__ uint32_t s_bin_local[RADIX];              // first slot of the digit inside the staged tile
    __shared__ uint32_t s_bin_delta[RADIX];              // global position = s_bin_delta[digit] + staged slot
    __shared__ uint32_t s_scan[SORT_WARPS];
    __shared__ uint32_t s_tile;
    __shared__ uint64_t s_stage[SORT_TILE];

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = tid; i < SORT_WARPS * RADIX; i += SORT_THREADS) (&s_warp_hist[0][0])[i] = 0;
    __syncthreads();
    const uint32_t n = n_dev ? *n_dev : n_host;
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * SORT_TILE;
    if (tile_base >= n) return;  // grid sized for an upper bound of n: surplus CTAs (whole CTA, uniform) have nothing to sort

    // -- load: warp-striped, item i of lane l = warp chunk + i*32 + l (coalesced, order-preserving)
    uint64_t kv[SORT_ITEMS];
    const uint32_t warp_base = tile_base + warp * (32 * SORT_ITEMS);
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t idx = warp_base + i * 32 + lane;
        if (idx < n) {
            if (FIRST) kv[i] = (static_cast<uint64_t>(__ldcs(static_cast<const uint32_t*>(in_raw) + idx)) << 32) | idx;
            else kv[i] = __ldcs(static_cast<const unsigned long long*>(in_raw) + idx);
        } else {
            kv[i] = ~0ull;
        }
    }

    // -- rank inside the warp, in element order (stable).  Two phases so that nothing serialises on a
    //    register dependency: (1) all match_any votes are issued back to back; (2) per item the lowest
    //    peer lane bumps the warp's digit counter with a shared-memory atomic (same-address atomics of
    //    one warp retire in program order, which is element order) and broadcasts the old value.
    uint32_t rank[SORT_ITEMS];  // holds the peer mask between the two phases
    const uint32_t lanes_below = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        // peers = lanes holding the same digit.  Built from one ballot per digit bit instead of
        // match.any: MATCH.ANY throughput capped the kernel at ~20 % issue utilisation (profiles/r1a).
        const uint32_t idx = warp_base + i * 32 + lane;
        const uint32_t digit = static_cast<uint32_t>(kv[i] >> (32 + shift)) & (RADIX - 1);
        uint32_t peers = __ballot_sync(0xffffffffu, idx < n);
#pragma unroll
        for (int b = 0; b < RADIX_BITS; b++) {
            const bool bit = (digit >> b) & 1u;
            const uint32_t vote = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? vote : ~vote;
        }
        rank[i] = peers;
    }
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t idx = warp_base + i * 32 + lane;
        const uint32_t digit = static_cast<uint32_t>(kv[i] >> (32 + shift)) & (RADIX - 1);
        const uint32_t peers = rank[i];
        const int leader = __ffs(peers) - 1;
        uint32_t before = 0;
        if (static_cast<int>(lane) == leader && idx < n && peers != 0u) before = atomicAdd(&s_warp_hist[warp][digit], static_cast<uint32_t>(__popc(peers)));
        before = __shfl_sync(0xffffffffu, before, leader);
        rank[i] = before + __popc(peers & lanes_below);
    }
    __syncthreads();

    // -- per digit (thread b owns digit b): exclusive prefix over the warps, tile total
    uint32_t tile_count = 0;
#pragma unroll
    for (int w = 0; w < SORT_WARPS; w++) {
        const uint32_t c = s_warp_hist[w][tid];
        s_warp_hist[w][tid] = tile_count;
        tile_count += c;
    }

    // -- decoupled look-back: how many pairs with this digit live in earlier tiles
    uint32_t* my_state = tile_state + static_cast<size_t>(tile) * RADIX + tid;
    uint32_t earlier = 0;
    if (tile == 0) {
        st_relaxed(my_state, FLAG_INCLUSIVE | tile_count);
    } else {
        st_relaxed(my_state, FLAG_AGGREGATE | tile_count);
        uint32_t t = tile;
        uint32_t spins = 0;
        while (true) {
            t--;
            const uint32_t* prev = tile_state + static_cast<size_t>(t) * RADIX + tid;
            uint32_t w = ld_relaxed(prev);
            while ((w >> 30) == 0u) {
                if (++spins > WATCHDOG_SPINS) {  // never expected: trip the watchdog instead of hanging the GPU
                    atomicExch(error_flag, 1u);
                    w = FLAG_INCLUSIVE;
                    break;
                }
                w = ld_relaxed(prev);
            }
            earlier += w & VALUE_MASK;
            if ((w >> 30) == 2u || t == 0) break;
        }
        st_relaxed(my_state, FLAG_INCLUSIVE | ((earlier + tile_count) & VALUE_MASK));
    }

    // -- bases: digit start in the whole array (scan of the global histogram) and in this tile
    const uint32_t global_base = block_exclusive_scan(hist[tid], s_scan);
    const uint32_t local_base = block_exclusive_scan(tile_count, s_scan);
    s_bin_local[tid] = local_base;
    s_bin_delta[tid] = global_base + earlier - local_base;
    __syncthreads();

    // -- reorder the tile in shared memory
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; i++) {
        const uint32_t idx = warp_base + i * 32 + lane;
        if (idx < n) {
            const uint32_t digit = static_cast<uint32_t>(kv[i] >> (32 + shift)) & (RADIX - 1);
            s_stage[s_bin_local[digit] + s_warp_hist[warp][digit] + rank[i]] = kv[i];
        }
    }
    __syncthreads();

    // -- write back: consecutive staged slots of one digit are consecutive in global memory
    const uint32_t valid = min(static_cast<uint32_t>(SORT_TILE), n - tile_base);
#pragma unroll 4
    for (uint32_t j = tid; j < valid; j += SORT_THREADS) {
        const uint64_t v = s_stage[j];
        const uint32_t digit = static_cast<uint32_t>(v >> (32 + shift)) & (RADIX - 1);
        out[s_bin_delta[digit] + j] = v;
    }
}

}  // namespace

size_t sort_workspace_bytes(uint32_t capacity) {
    const size_t tiles = (static_cast<size_t>(capacity) + SORT_TILE - 1) / SORT_TILE + 1;
    const size_t words = static_cast<size_t>(MAX_SORT_PASSES) * RADIX   // hist
                         + 64                                            // tile counters + error flag (padded)
                         + static_cast<size_t>(MAX_SORT_PASSES) * tiles * RADIX;
    return words * sizeof(uint32_t);
}

void sort_workspace_bind(SortWorkspace& ws, void* base, uint32_t capacity) {
    const size_t tiles = (static_cast<size_t>(capacity) + SORT_TILE - 1) / SORT_TILE + 1;
    uint32_t* w = static_cast<uint32_t*>(base);
    ws.hist = w;
    ws.tile_counter = w + MAX_SORT_PASSES * RADIX;
    ws.error_flag = ws.tile_counter + 32;
    ws.tile_state = w + MAX_SORT_PASSES * RADIX + 64;
    ws.zero_base = base;
    ws.zero_bytes = sort_workspace_bytes(capacity);
    ws.tiles_cap = static_cast<uint32_t>(tiles);
}

static int sort_passes_for(int key_bits) {
    int passes = (key_bits + RADIX_BITS - 1) / RADIX_BITS;
    if (passes < 1) passes = 1;
    if (passes > MAX_SORT_PASSES) passes = MAX_SORT_PASSES;
    return passes;
}

// zero the digit histograms, the tile tickets and the look-back words one sort of n keys will use
void sort_prepare(cudaStream_t s, uint32_t n, int key_bits, const SortWorkspace& ws, Profiler* prof) {
    if (n == 0) return;
    const int passes = sort_passes_for(key_bits);
    const uint32_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
    const size_t used_words = static_cast<size_t>(MAX_SORT_PASSES) * RADIX + 64 + static_cast<size_t>(passes) * tiles * RADIX;
    prof->begin(s, K_MEMSET);
    cudaMemsetAsync(ws.zero_base, 0, used_words * sizeof(uint32_t), s);
    prof->end(s);
}

int launch_sort(cudaStream_t s, uint32_t n, const uint32_t* keys, uint64_t* buf_a, uint64_t* buf_b, int key_bits, const SortWorkspace& ws,
                uint64_t** result, bool hist_ready, Profiler* prof, const uint32_t* n_dev) {
    *result = buf_a;
    if (n == 0) return 0;
    const int passes = sort_passes_for(key_bits);
    const uint32_t tiles = (n + SORT_TILE - 1) / SORT_TILE;
    int launches = 0;
    if (!hist_ready) {  // keys did not come from the move pass: histogram them here
        sort_prepare(s, n, key_bits, ws, prof);
        uint32_t hblocks = (n / 4 + 255) / 256;
        if (hblocks > 148u * 8u) hblocks = 148u * 8u;
        if (hblocks < 1) hblocks = 1;
        prof->begin(s, K_HISTOGRAM);
        histogram_kernel<<<hblocks, 256, 0, s>>>(keys, n, n_dev, ws.hist, passes);
        prof->end(s);
        launches++;
    }
    const void* in = keys;
    uint64_t* out = buf_a;
    for (int p = 0; p < passes; p++) {
        uint32_t* state = ws.tile_state + static_cast<size_t>(p) * tiles * RADIX;
        prof->begin(s, K_SORT_PASS0 + p);
        if (p == 0)
            onesweep_kernel<true><<<tiles, SORT_THREADS, 0, s>>>(in, out, n, n_dev, p * RADIX_BITS, ws.hist + p * RADIX, state, ws.tile_counter + p, ws.error_flag);
        else
            onesweep_kernel<false><<<tiles, SORT_THREADS, 0, s>>>(in, out, n, n_dev, p * RADIX_BITS, ws.hist + p * RADIX, state, ws.tile_counter + p, ws.error_flag);
        prof->end(s);
        launches++;
        *result = out;
        in = out;
        out = (out == buf_a) ? buf_b : buf_a;
    }
    return launches;
}

}  // namespace msim
