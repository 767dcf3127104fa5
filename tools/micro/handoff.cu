// micro-benchmark (developer probe, for the next round): what does it cost to hand a result from one kernel of a decode
// step to the next?  The step is a chain of 81 dependent launches; profiles/r1_timeline_and_micro.txt puts the hand-over
// (previous kernel's end -> the next kernel's griddepcontrol.wait returning) at 2.0-2.5 us each, ~25 % of the step.
//   A  plain stream order (no programmatic dependent launch)
//   B  programmatic dependent launch: launch_dependents at entry, griddepcontrol.wait before the dependent part
//   C  B's early launch, but the data hand-over is a release/acquire counter in global memory instead of kernel
//      completion: the consumer's CTAs are already resident and spinning when the producer's last CTA signals
// Each kernel does ~1 us of dependent work per CTA so that hand-over and work can be told apart (D = work alone).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/handoff tools/micro/handoff.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float spin_work(float x, int iters) {
    for (int i = 0; i < iters; i++) x = __fadd_rn(x, 1.0f);   // dependent chain: ~4.6 cycles per add
    return x;
}
__global__ void k_plain(float* out, int iters) {
    out[blockIdx.x * blockDim.x + threadIdx.x] = spin_work(out[threadIdx.x], iters);
}
__global__ void k_pdl(float* out, int iters) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    out[blockIdx.x * blockDim.x + threadIdx.x] = spin_work(out[threadIdx.x], iters);
}
// flags[i] counts the CTAs of link i that have published their results
__global__ void k_flag(float* out, int iters, unsigned* flags, int link, int wait_every) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (link % wait_every == 0) asm volatile("griddepcontrol.wait;" ::: "memory");   // bounds how deep the chain pre-launches
    if (link > 0 && threadIdx.x == 0) {
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + link - 1) : "memory");
            if (v < gridDim.x) __nanosleep(32);
        } while (v < gridDim.x);
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = spin_work(__ldcg(out + threadIdx.x), iters);
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(flags + link) : "memory");
}

template <typename F> static float time_chain(cudaStream_t st, int reps, F launch_chain) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch_chain(); cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; r++) launch_chain();
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / reps;
}

int main() {
    const int CTAS = 148, THREADS = 256, LINKS = 81, ITERS = 420;   // 420 adds ~ 1 us
    float* out; unsigned* flags;
    cudaMalloc(&out, CTAS * THREADS * 4); cudaMemset(out, 0, CTAS * THREADS * 4);
    cudaMalloc(&flags, LINKS * 4);
    cudaStream_t st; cudaStreamCreate(&st);
    auto launch = [&](auto kernel, bool pdl, auto... args) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(CTAS); cfg.blockDim = dim3(THREADS); cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
        cudaLaunchKernelEx(&cfg, kernel, args...);
    };
    const float work = time_chain(st, 20, [&] { launch(k_plain, false, out, ITERS * LINKS); });
    const float a = time_chain(st, 20, [&] { for (int i = 0; i < LINKS; i++) launch(k_plain, false, out, ITERS); });
    const float b = time_chain(st, 20, [&] { for (int i = 0; i < LINKS; i++) launch(k_pdl, i > 0, out, ITERS); });
    printf("D work alone (one kernel, %d x %d adds)          %8.1f us\n", LINKS, ITERS, work);
    printf("A plain stream order, %d launches               %8.1f us  -> %.2f us per hand-over\n", LINKS, a, (a - work) / LINKS);
    printf("B programmatic dependent launch                  %8.1f us  -> %.2f us per hand-over\n", b, (b - work) / LINKS);
    for (int wait_every : {81, 8, 4}) {
        const float c = time_chain(st, 20, [&] {
            cudaMemsetAsync(flags, 0, LINKS * 4, st);
            for (int i = 0; i < LINKS; i++) launch(k_flag, i > 0, out, ITERS, flags, i, wait_every);
        });
        printf("C flag hand-over, full wait every %2d links       %8.1f us  -> %.2f us per hand-over\n", wait_every, c, (c - work) / LINKS);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
