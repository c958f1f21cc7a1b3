// tests/cuda_emu/cuda_runtime.h - TEST INFRASTRUCTURE: a host SIMT emulator for the handful of CUDA features the kernels of
// movement-sim_b200/csrc use, so that kernels written after the round's GPU budget was spent can at least EXECUTE before they meet hardware.
// It shadows <cuda_runtime.h> (put this directory first on the include path and define MSIM_HOST_EMU).
//
// Model: one OS thread per CUDA thread of a block; the blocks of a grid run one after another (the same OS threads walk them).  __syncthreads() is a barrier over the block's
// threads; warp collectives (__ballot_sync, __shfl*_sync, __any_sync, __reduce_*_sync) exchange through a per-warp slot array between two
// barriers over the warp's 32 threads (the kernels only ever use the full mask with all lanes present); atomics are GCC __atomic builtins;
// __shared__ variables are statics (one block at a time); round-to-nearest intrinsics are plain binary32 operations (build with
// -ffp-contract=off).  What it cannot show: anything about timing, memory-model weaknesses of real hardware, or PTX-level code (the few
// inline-PTX helpers have emulator twins next to them in the sources).
#pragma once
#include <pthread.h>

#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <tuple>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
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

// ---- the host runtime, synchronous: "device" memory is host memory, streams and events do nothing, copies are memcpy -----------------
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaLimit { cudaLimitPersistingL2CacheSize = 6 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaStreamAttrID { cudaStreamAttributeAccessPolicyWindow = 1 };
enum cudaAccessProperty { cudaAccessPropertyNormal = 0, cudaAccessPropertyStreaming = 1, cudaAccessPropertyPersisting = 2 };
struct cudaAccessPolicyWindow { void* base_ptr; size_t num_bytes; float hitRatio; cudaAccessProperty hitProp, missProp; };
union cudaStreamAttrValue { cudaAccessPolicyWindow accessPolicyWindow; };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount, persistingL2CacheMaxSize, accessPolicyMaxWindowSize; };
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    std::memset(p, 0, sizeof(*p));
    std::strcpy(p->name, "host SIMT emulator (tests/cuda_emu)");
    p->major = 10;
    p->multiProcessorCount = 148;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t bytes) { *p = std::calloc(1, bytes ? bytes : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc(reinterpret_cast<void**>(p), bytes); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t bytes) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(std::malloc(1)); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaStreamSetAttribute(cudaStream_t, cudaStreamAttrID, const cudaStreamAttrValue*) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(std::malloc(1)); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.001f; return cudaSuccess; }
// one process, one address space: an IPC handle is the pointer itself
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof(*h)); std::memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
template <typename K> inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <typename K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 8; return cudaSuccess; }

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
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline void __syncwarp(unsigned = 0xffffffffu) { cuda_emu::warp_sync(); }
inline unsigned __shfl_sync(unsigned, unsigned v, unsigned src) {
    return cuda_emu::collective(v, [src](const unsigned long long* s) { return static_cast<unsigned>(s[src & 31u]); });
}
inline int __shfl_sync(unsigned m, int v, int src) { return static_cast<int>(__shfl_sync(m, static_cast<unsigned>(v), static_cast<unsigned>(src))); }
inline unsigned __shfl_sync(unsigned m, unsigned v, int src) { return __shfl_sync(m, v, static_cast<unsigned>(src)); }
inline float __shfl_sync(unsigned m, float v, int src) {  // binary32 travels as its bit pattern
    unsigned b;
    std::memcpy(&b, &v, sizeof(b));
    b = __shfl_sync(m, b, static_cast<unsigned>(src));
    std::memcpy(&v, &b, sizeof(v));
    return v;
}
inline unsigned __shfl_up_sync(unsigned, unsigned v, unsigned delta) {
    const unsigned me = cuda_emu::lane;
    return cuda_emu::collective(v, [me, delta](const unsigned long long* s) { return static_cast<unsigned>(me >= delta ? s[me - delta] : s[me]); });
}
inline unsigned __shfl_down_sync(unsigned, unsigned v, unsigned delta) {
    const unsigned me = cuda_emu::lane;
    return cuda_emu::collective(v, [me, delta](const unsigned long long* s) { return static_cast<unsigned>(me + delta < 32u ? s[me + delta] : s[me]); });
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
inline unsigned atomicAnd(unsigned* p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
// twins of the system-scope PTX accesses of msim_internal.h
inline void st_release_sys(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline void st_relaxed_sys(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline uint32_t ld_acquire_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }

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
extern thread_local void* dynamic_smem_ptr;
inline void* dynamic_smem() { return dynamic_smem_ptr; }
inline unsigned long long now_ns() { return static_cast<unsigned long long>(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count()); }
// kernel<<<grid, block>>>(args...) on the host: one block at a time, one OS thread per CUDA thread
template <typename K, typename... A>
void launch(K kernel, unsigned grid, unsigned threads, A... args) {
    if (grid == 0 || threads == 0) return;
    // the OS threads are created once per launch and walk the blocks together: the barrier at the end of a block keeps every thread of
    // block b out of block b + 1 until all have left b (static __shared__ variables and the dynamic buffer are reused from block to block)
    Block blk;
    pthread_barrier_init(&blk.bar, nullptr, threads);
    blk.warps.resize((threads + 31u) / 32u);
    for (unsigned w = 0; w < blk.warps.size(); w++) pthread_barrier_init(&blk.warps[w].bar, nullptr, std::min(32u, threads - 32u * w));
    std::vector<std::thread> pool;
    pool.reserve(threads);
    void* const smem = dynamic_smem_ptr;  // the launching thread's setting (cfg below) goes to every thread
    for (unsigned t = 0; t < threads; t++) {
        pool.emplace_back([&, t] {
            block = &blk;
            dynamic_smem_ptr = smem;
            lane = t & 31u;
            warp = t >> 5;
            threadIdx = {t, 0, 0};
            blockDim = dim3(threads);
            gridDim = dim3(grid);
            for (unsigned b = 0; b < grid; b++) {
                blockIdx = {b, 0, 0};
                kernel(args...);
                pthread_barrier_wait(&blk.bar);
            }
        });
    }
    for (std::thread& th : pool) th.join();
    for (Warp& w : blk.warps) pthread_barrier_destroy(&w.bar);
    pthread_barrier_destroy(&blk.bar);
}
// kernel<<<grid, block, smem, stream>>>(args...) is rewritten to cuda_emu::cfg(kernel, grid, block, smem, stream)(args...) by build_emu_lib.py
template <typename K>
struct Configured {
    K kernel;
    unsigned grid, threads;
    size_t smem;
    template <typename... A>
    void operator()(A... args) const {
        if (grid == 0) return;
        std::vector<unsigned long long> dyn((smem + 7) / 8 + 1);
        dynamic_smem_ptr = dyn.data();
        launch(kernel, grid, threads, args...);
        dynamic_smem_ptr = nullptr;
    }
};
template <typename K>
Configured<K> cfg(K kernel, unsigned grid, unsigned threads, size_t smem = 0, cudaStream_t = nullptr) {
    return Configured<K>{kernel, grid, threads, smem};
}
}  // namespace cuda_emu
