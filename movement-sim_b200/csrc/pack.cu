// pack.cu — the boundary transposes between the reference's 64-byte AoS entity
// (/root/reference/src/sim/Entity.hpp:33-46, shader EntityDescriptor random_move.comp:5-13) and the
// resident structure-of-arrays state.
//
//   unpack : OpTensorSyncDevice side (Simulator.cpp:191-192) — AoS chunk in device staging -> SoA
//   pack   : OpTensorSyncLocal side  (Simulator.cpp:250,262) — SoA -> AoS chunk, re-creating the two
//            fields the hot path never stores:
//              direction  = update_direction() of the last move pass (random_move.comp:830-839):
//                           from the previous position for walkers, from the reached waypoint
//                           towards the new one for entities that arrived (:848-850);
//              colour     = as uploaded until the first collision pass, then green / blue
//                           (:876, :546-547) from the 1-byte collision flag.
#include "msim_internal.h"

namespace msim {
namespace {

constexpr float SPEED = 1.4f;

__device__ __forceinline__ float2 direction_from(float2 from, float2 target) {
    const float dx = __fsub_rn(target.x, from.x);
    const float dy = __fsub_rn(target.y, from.y);
    const float len = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    if (len == 0.0f) return make_float2(0.0f, 0.0f);
    return make_float2(__fmul_rn(__fdiv_rn(dx, len), SPEED), __fmul_rn(__fdiv_rn(dy, len), SPEED));
}

__global__ void __launch_bounds__(256) pack_kernel(uint32_t first, uint32_t count, PackArgs a, float4* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    // external entity id first + i lives in storage slot slot_of[first + i] once the state has been re-sorted
    const uint32_t e = a.slot_of ? a.slot_of[first + i] : first + i;
    const float2 p = a.pos_cur[e];
    const float2 t = a.target[e];
    float2 dir;
    if (a.has_moved) {
        const bool arrived = (a.arrived[arrived_word(e)] >> arrived_bit(e)) & 1u;
        dir = direction_from(arrived ? p : a.pos_prev[e], t);
    } else {
        dir = a.dir0[e];
    }
    float4 color;
    const uint8_t flag = a.flag_entity ? a.flag_entity[e] : 0;
    if (flag == 0) color = a.color0[e];
    else color = (flag == 2) ? make_float4(0.f, 0.f, 1.f, 1.f) : make_float4(0.f, 1.f, 0.f, 1.f);
    const uint4 s = a.rng[e];
    float4* out = dst + static_cast<size_t>(i) * 4;
    out[0] = color;
    out[1] = make_float4(__uint_as_float(s.x), __uint_as_float(s.y), __uint_as_float(s.z), __uint_as_float(s.w));
    out[2] = make_float4(p.x, p.y, t.x, t.y);
    out[3] = make_float4(dir.x, dir.y, __uint_as_float(a.road[e]), __uint_as_float(a.initialized_all));
}

__global__ void __launch_bounds__(256)
unpack_kernel(uint32_t first, uint32_t count, const float4* __restrict__ src, float2* __restrict__ pos, float2* __restrict__ target,
              uint32_t* __restrict__ road, uint4* __restrict__ rng, float4* __restrict__ color0, float2* __restrict__ dir0,
              unsigned int* __restrict__ uninit_count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool uninit = false;
    if (i < count) {
        const uint32_t e = first + i;
        const float4* in = src + static_cast<size_t>(i) * 4;
        const float4 c = in[0], s = in[1], pt = in[2], dr = in[3];
        color0[e] = c;
        rng[e] = make_uint4(__float_as_uint(s.x), __float_as_uint(s.y), __float_as_uint(s.z), __float_as_uint(s.w));
        pos[e] = make_float2(pt.x, pt.y);
        target[e] = make_float2(pt.z, pt.w);
        dir0[e] = make_float2(dr.x, dr.y);
        road[e] = __float_as_uint(dr.z);
        uninit = __float_as_uint(dr.w) == 0u;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, uninit);
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(uninit_count, static_cast<unsigned int>(__popc(m)));
}

__global__ void __launch_bounds__(256) max_road_kernel(uint32_t n, const uint32_t* __restrict__ road, unsigned int* __restrict__ out_max) {
    uint32_t m = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, road[i]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31u) == 0) atomicMax(out_max, m);
}

// ---- cell-ordered storage ---------------------------------------------------------------------
// Every REORDER_EVERY collision passes the resident state is physically permuted into the cell order
// the pass has just computed (new slot j <- old slot sorted_idx[j]).  Consecutive slots then hold
// entities of the same or neighbouring cells for the next few dozen ticks, which turns the scattered
// stores / atomics / gathers of the neighbour rebuild into L2-local traffic (profiles/r1_sort_paths.md).
// ext_id maps slot -> external entity id, slot_of is its inverse; both stay NULL until the first re-sort.
__global__ void __launch_bounds__(256)
reorder_kernel(uint32_t n, const uint32_t* __restrict__ sorted_idx, const uint8_t* __restrict__ flag_sorted, ReorderArrays a) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t src = __ldcs(sorted_idx + j);
    a.pos_prev_new[j] = a.pos_prev[src];
    a.target_new[j] = a.target[src];
    a.road_new[j] = a.road[src];
    a.rng_new[j] = a.rng[src];
    const uint32_t ext = a.ext_id ? a.ext_id[src] : src;
    a.ext_id_new[j] = ext;
    a.slot_of[ext] = j;
    a.flag_entity[j] = flag_sorted[j] + 1;  // 1 = green, 2 = blue: the flags are in slot order already
    if ((a.arrived[arrived_word(src)] >> arrived_bit(src)) & 1u) atomicOr(a.arrived_new + arrived_word(j), 1u << arrived_bit(j));
}

// Sharded handles: new slot j <- old slot sorted_idx[first + j], j < n_owned.  Ghost rows sort before and
// after the band's own rows (row-major keys), so the owned entities are one contiguous run of the order.
__global__ void __launch_bounds__(256)
reorder_sharded_kernel(const uint32_t* __restrict__ sorted_idx, const uint8_t* __restrict__ flag_sorted, ReorderArrays a) {
    const uint32_t n = *a.n_owned_dev, first = *a.first_owned;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t src = __ldcs(sorted_idx + first + j);
    if (src >= n) {  // cannot happen while every owned entity lies inside the band (msim_shard_pack guarantees it)
        atomicOr(a.error_word, 16u);
        return;
    }
    a.pos_cur_new[j] = a.sorted_pos[first + j];
    a.pos_prev_new[j] = a.pos_prev[src];
    a.target_new[j] = a.target[src];
    a.road_new[j] = a.road[src];
    a.rng_new[j] = a.rng[src];
    a.ext_id_new[j] = a.ext_id[src];
    a.flag_entity[j] = flag_sorted[first + j] + 1;
    if ((a.arrived[arrived_word(src)] >> arrived_bit(src)) & 1u) atomicOr(a.arrived_new + arrived_word(j), 1u << arrived_bit(j));
}

__global__ void __launch_bounds__(256) gather_pos_kernel(uint32_t n, const uint32_t* __restrict__ slot_of, const float2* __restrict__ pos, float2* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = pos[slot_of[i]];
}

// readback form of the collision flag: 1 = blue (a neighbour within the radius), 0 = green or no collision pass yet (internally 2 / 1 / 0)
__global__ void __launch_bounds__(256) gather_flag_kernel(uint32_t n, const uint32_t* __restrict__ slot_of, const uint8_t* __restrict__ flag, uint8_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = flag[slot_of ? slot_of[i] : i] == 2 ? 1 : 0;
}

}  // namespace

int launch_reorder(cudaStream_t s, uint32_t n, const uint32_t* sorted_idx, const uint8_t* flag_sorted, const ReorderArrays& a, Profiler* prof) {
    if (n == 0) return 0;
    prof->begin(s, K_REORDER);
    cudaMemsetAsync(a.arrived_new, 0, (static_cast<size_t>(n + 63u) / 64u) * 2u * sizeof(uint32_t), s);
    reorder_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(n, sorted_idx, flag_sorted, a);
    prof->end(s);
    return 1;
}

int launch_reorder_sharded(cudaStream_t s, uint32_t n_upper, uint32_t arrived_words, const uint32_t* sorted_idx, const uint8_t* flag_sorted,
                           const ReorderArrays& a, Profiler* prof) {
    if (n_upper == 0) return 0;
    prof->begin(s, K_REORDER);
    cudaMemsetAsync(a.arrived_new, 0, static_cast<size_t>(arrived_words) * sizeof(uint32_t), s);
    reorder_sharded_kernel<<<(n_upper + 255u) / 256u, 256, 0, s>>>(sorted_idx, flag_sorted, a);
    prof->end(s);
    return 1;
}

int launch_gather_pos(cudaStream_t s, uint32_t n, const uint32_t* slot_of, const float2* pos, float2* out) {
    if (n == 0) return 0;
    gather_pos_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(n, slot_of, pos, out);
    return 1;
}

int launch_gather_flag(cudaStream_t s, uint32_t n, const uint32_t* slot_of, const uint8_t* flag, uint8_t* out) {
    if (n == 0) return 0;
    gather_flag_kernel<<<(n + 255u) / 256u, 256, 0, s>>>(n, slot_of, flag, out);
    return 1;
}

int launch_pack(cudaStream_t s, uint32_t first, uint32_t count, const PackArgs& a, msim_entity* dst, Profiler* prof) {
    if (count == 0) return 0;
    prof->begin(s, K_PACK);
    pack_kernel<<<(count + 255u) / 256u, 256, 0, s>>>(first, count, a, reinterpret_cast<float4*>(dst));
    prof->end(s);
    return 1;
}

int launch_unpack(cudaStream_t s, uint32_t first, uint32_t count, const msim_entity* src, float2* pos, float2* target, uint32_t* road,
                  uint4* rng, float4* color0, float2* dir0, uint8_t* /*init_mask*/, unsigned int* uninit_count, Profiler* prof) {
    if (count == 0) return 0;
    prof->begin(s, K_UNPACK);
    unpack_kernel<<<(count + 255u) / 256u, 256, 0, s>>>(first, count, reinterpret_cast<const float4*>(src), pos, target, road, rng, color0, dir0,
                                                        uninit_count);
    prof->end(s);
    return 1;
}

int launch_max_road(cudaStream_t s, uint32_t n, const uint32_t* road, unsigned int* out_max) {
    if (n == 0) return 0;
    uint32_t blocks = (n + 255u) / 256u;
    if (blocks > 1184u) blocks = 1184u;
    max_road_kernel<<<blocks, 256, 0, s>>>(n, road, out_max);
    return 1;
}

}  // namespace msim
