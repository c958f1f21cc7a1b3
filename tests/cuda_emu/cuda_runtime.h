// tests/cuda_emu/cuda_runtime.h - TEST INFRASTRUCTURE: a host SIMT emulator for the handful of CUDA features the kernels of
// movement-sim_b200/csrc use, so that kernels written after the round's GPU budget was spent can at least EXECUTE before they meet hardware.
// It shadows <cuda_runtime.h> (put this directory first on the include path and define MSIM_HOST_EMU).
//
// Model: one OS thread per CUDA thread of a block; the blocks of a grid run one after another.  __syncthreads() is a barrier over the block's
// threads; warp collectives (__ballot_sync, __shfl*_sync, __any_sync, __reduce_*_sync) exchange through a per-warp slot array between two
// barriers over the warp's 32 threads (the kernels only ever use the full mask with all lanes present); atomics are GCC __atomic builtins;
// __shared__ variables are statics (one block at a time); round-to-nearest intrinsics are plain binary32 operations (build with
// -ffp-contract=off).  What it cannot show: anything about timing, memory-model weaknesses of real hardware, or PTX-level code (the few
// inline-PTX helpers have emulator twins next to them in the sources).
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <tuple>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct uchar2 { unsigned char x, y; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline uchar2 make_uchar2(unsigned char x, unsigned char y) { return {x, y}; }

// the host runtime names msim_internal.h mentions; nothing of it is ever called under the emulator
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
typedef int cudaError_t;
inline cudaError_t cudaEventCreate(cudaEvent_t*) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }

namespace cuda_emu {
struct Warp {
    pthread_barrier_t bar;
    unsigned long long slot[32];
};
struct Block {
    pthread_barrier_t bar;
    std::vector<Warp> warps;
};
extern thread_local Block* block;
extern thread_local unsigned lane, warp;
inline void warp_sync() { pthread_barrier_wait(&block->warps[warp].bar); }
template <typename F>
inline auto collective(unsigned long long mine, F combine) {
    Warp& w = block->warps[warp];
    w.slot[lane] = mine;
    warp_sync();
    auto r = combine(w.slot);
    warp_sync();
    return r;
}
}  // namespace cuda_emu

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { pthread_barrier_wait(&cuda_emu::block->bar); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

inline unsigned __ballot_sync(unsigned, int pred) {
    return cuda_emu::collective(pred ? 1ull : 0ull, [](const unsigned long long* s) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= (s[i] ? 1u : 0u) << i; return r; });
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
inline unsigned __shfl_sync(unsigned, unsigned v, unsigned src) {
    return cuda_emu::collective(v, [src](const unsigned long long* s) { return static_cast<unsigned>(s[src & 31u]); });
}
inline unsigned __shfl_up_sync(unsigned, unsigned v, unsigned delta) {
    const unsigned me = cuda_emu::lane;
    return cuda_emu::collective(v, [me, delta](const unsigned long long* s) { return static_cast<unsigned>(me >= delta ? s[me - delta] : s[me]); });
}
inline unsigned long long __shfl_down_sync(unsigned, unsigned long long v, unsigned delta) {
    const unsigned me = cuda_emu::lane;
    return cuda_emu::collective(v, [me, delta](const unsigned long long* s) { return me + delta < 32u ? s[me + delta] : s[me]; });
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    return cuda_emu::collective(v, [](const unsigned long long* s) { unsigned r = 0; for (int i = 0; i < 32; i++) r += static_cast<unsigned>(s[i]); return r; });
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    return cuda_emu::collective(v, [](const unsigned long long* s) { unsigned r = ~0u; for (int i = 0; i < 32; i++) r = std::min(r, static_cast<unsigned>(s[i])); return r; });
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    return cuda_emu::collective(v, [](const unsigned long long* s) { unsigned r = 0; for (int i = 0; i < 32; i++) r = std::max(r, static_cast<unsigned>(s[i])); return r; });
}

inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }

template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }

inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline float __uint2float_rn(unsigned v) { return static_cast<float>(v); }                       // u32 -> binary32, round to nearest even
inline unsigned __float2uint_ru(float v) { return static_cast<unsigned>(std::ceil(v)); }
inline int __float2int_rd(float v) { return v != v ? 0 : static_cast<int>(std::floor(v)); }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz(static_cast<unsigned>(v)) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __fns(unsigned mask, unsigned base, int offset) {  // position of the offset-th set bit at or above `base` (offset > 0)
    for (unsigned b = base; b < 32u; b++)
        if ((mask >> b) & 1u)
            if (--offset == 0) return b;
    return 0xffffffffu;
}
using std::max;
using std::min;
inline unsigned min(unsigned a, int b) { return std::min(a, static_cast<unsigned>(b)); }
inline int min(int a, unsigned b) { return std::min(a, static_cast<int>(b)); }

namespace cuda_emu {
// kernel<<<grid, block>>>(args...) on the host: one block at a time, one OS thread per CUDA thread
template <typename K, typename... A>
void launch(K kernel, unsigned grid, unsigned threads, A... args) {
    for (unsigned b = 0; b < grid; b++) {
        Block blk;
        pthread_barrier_init(&blk.bar, nullptr, threads);
        blk.warps.resize((threads + 31u) / 32u);
        for (unsigned w = 0; w < blk.warps.size(); w++) pthread_barrier_init(&blk.warps[w].bar, nullptr, std::min(32u, threads - 32u * w));
        std::vector<std::thread> pool;
        pool.reserve(threads);
        for (unsigned t = 0; t < threads; t++) {
            pool.emplace_back([&, t, b] {
                block = &blk;
                lane = t & 31u;
                warp = t >> 5;
                threadIdx = {t, 0, 0};
                blockIdx = {b, 0, 0};
                blockDim = dim3(threads);
                gridDim = dim3(grid);
                kernel(args...);
            });
        }
        for (std::thread& th : pool) th.join();
        for (Warp& w : blk.warps) pthread_barrier_destroy(&w.bar);
        pthread_barrier_destroy(&blk.bar);
    }
}
}  // namespace cuda_emu
