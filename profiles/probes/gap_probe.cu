// gap_probe.cu - what makes a tiny kernel expensive on a stream?  Times back-to-back launches of one-CTA kernels that differ in ONE
// ingredient of the shard exchange kernel (system-scope fence, system-scope acquire load, nanosleep, %globaltimer, 1024 threads).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gap_probe gap_probe.cu ; run: ./gap_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_empty(unsigned* p) { if (p && threadIdx.x == 9999) *p = 1; }
__global__ void k_store(unsigned* p) { if (threadIdx.x == 0) *p = 1; }
__global__ void k_fence_gpu(unsigned* p) { if (threadIdx.x == 0) { *p = 1; __threadfence(); p[1] = 2; } }
__global__ void k_fence_sys(unsigned* p) { if (threadIdx.x == 0) { *p = 1; __threadfence_system(); p[1] = 2; } }
__global__ void k_acq_sys(unsigned* p) { if (threadIdx.x == 0) { unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); p[1] = v; } }
__global__ void k_sleep(unsigned* p) { if (threadIdx.x == 0) { __nanosleep(200); p[1] = 2; } }
__global__ void k_timer(unsigned* p) { if (threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); p[1] = (unsigned)t; } }
__global__ void k_atomic(unsigned* p) { if (threadIdx.x == 0) atomicOr(p, 1u); }

template <typename F>
float time_it(F launch, cudaStream_t s, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 20; i++) launch();
    cudaStreamSynchronize(s);
    cudaEventRecord(a, s);
    for (int i = 0; i < reps; i++) launch();
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms * 1000.f / reps;
}

int main() {
    unsigned* p; cudaMalloc(&p, 256); cudaMemset(p, 0, 256);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    const int R = 2000;
#define T(name, call) printf("%-34s %6.2f us per launch\n", name, time_it([&] { call; }, s, R));
    T("empty <<<1,32>>>", (k_empty<<<1, 32, 0, s>>>(p)));
    T("empty <<<1,1024>>>", (k_empty<<<1, 1024, 0, s>>>(p)));
    T("empty <<<8,1024>>>", (k_empty<<<8, 1024, 0, s>>>(p)));
    T("store <<<1,32>>>", (k_store<<<1, 32, 0, s>>>(p)));
    T("store + fence.gpu <<<1,32>>>", (k_fence_gpu<<<1, 32, 0, s>>>(p)));
    T("store + fence.sys <<<1,32>>>", (k_fence_sys<<<1, 32, 0, s>>>(p)));
    T("ld.acquire.sys <<<1,32>>>", (k_acq_sys<<<1, 32, 0, s>>>(p)));
    T("nanosleep(200) <<<1,32>>>", (k_sleep<<<1, 32, 0, s>>>(p)));
    T("globaltimer <<<1,32>>>", (k_timer<<<1, 32, 0, s>>>(p)));
    T("atomicOr <<<1,32>>>", (k_atomic<<<1, 32, 0, s>>>(p)));
    // the same with a second stream busy-free: does an event wait between kernels cost?
    cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaStream_t s2; cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    T("empty + record/wait other stream", (k_empty<<<1, 32, 0, s>>>(p), cudaEventRecord(e, s), cudaStreamWaitEvent(s2, e, 0), k_empty<<<1, 32, 0, s2>>>(p)));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
