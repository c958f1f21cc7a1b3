// quadtree.cu — the display quadtree behind msim_read_quadtree_nodes.
//
// The reference keeps a lock-based incremental quadtree on the GPU
// (/root/reference/src/sim/shader/random_move.comp:100-539) whose 64-byte nodes the UI walks to draw a
// grid overlay (src/ui/widgets/opengl/QuadTreeGridGlObject.cpp:10-51: offset, size, contentType and
// the four child links are all it reads).  The hot path here needs no tree, so the nodes are built on
// demand: the tree a fresh insertion of the current positions produces — a node is split iff it holds
// more than entityNodeCap entities and lies above maxDepth (:354) — which is independent of insertion
// order.  (The reference's tree can be deeper where a region used to be crowded: it merges lazily,
// :464-475; and it keeps entities with identical positions together, :301-307.  Both are history /
// tie effects the overlay does not depend on.)
//
// Device part: a histogram of the entities over the 2^(maxDepth-1) x 2^(maxDepth-1) finest cells.  The
// cell of a position is found by the shader's own descent (:319-341): compare with
// offset + width/2 in binary32, halve, repeat — so entities on a boundary go where the shader puts them.
// Host part (api.cu): sum the histogram bottom-up and emit nodes top-down.
#include "msim_internal.h"

namespace msim {
namespace {

__device__ __forceinline__ uint32_t descend(float x, float extent, int levels) {
    float off = 0.0f, width = extent;
    uint32_t index = 0;
    for (int l = 0; l < levels; l++) {
        width = __fmul_rn(width, 0.5f);            // newWidth = width / 2 (exact)
        const float mid = __fadd_rn(off, width);   // offsetXNext = offsetX + width / 2
        index <<= 1;
        if (!(x < mid)) {                          // :325 "if (ePos.x < offsetXNext) left else right"
            off = mid;
            index |= 1u;
        }
    }
    return index;
}

__global__ void __launch_bounds__(256)
leaf_histogram_kernel(uint32_t n, const float2* __restrict__ pos, float world_w, float world_h, int levels, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t s_hist[];
    const uint32_t side = 1u << levels, bins = side * side;
    for (uint32_t i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const float2 p = __ldcs(pos + e);
        atomicAdd(&s_hist[descend(p.y, world_h, levels) * side + descend(p.x, world_w, levels)], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < bins; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

}  // namespace

int launch_leaf_histogram(cudaStream_t s, int sm_count, uint32_t n, const float2* pos, float world_w, float world_h, int levels, uint32_t* hist) {
    const uint32_t bins = 1u << (2 * levels);
    cudaMemsetAsync(hist, 0, bins * sizeof(uint32_t), s);
    if (n == 0) return 0;
    const size_t smem = bins * sizeof(uint32_t);  // 64 KiB at the reference's depth 8
    cudaFuncSetAttribute(leaf_histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    uint32_t blocks = (n + 255u) / 256u;
    const uint32_t cap = static_cast<uint32_t>(sm_count) * 2u;
    if (blocks > cap) blocks = cap;
    leaf_histogram_kernel<<<blocks, 256, smem, s>>>(n, pos, world_w, world_h, levels, hist);
    return 1;
}

}  // namespace msim
