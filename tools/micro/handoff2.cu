// micro-benchmark (developer probe, round 2): cost of handing a small activation vector from one phase of a decode step
// to the next, for the mechanisms the decode chain could use.  Every link does ~1 us of dependent work per CTA, writes
// its slice of a 2048-float vector and the next link reads the WHOLE vector (what a GEMV prologue does).
//   pdl        programmatic dependent launch + griddepcontrol.wait (what round 1 shipped)
//   flagN      early launch as above, but the hand-over is a release/acquire counter: variants of the poll / signal code
//   persist    one persistent kernel, 148 CTAs, a grid barrier between links (central counter or per-CTA flag words)
// `smem` limits co-residency (100 KB dynamic -> 2 CTAs per SM) like the real kernels.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/handoff2 tools/micro/handoff2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int CTAS = 148, THREADS = 256, LINKS = 80, ITERS = 420, VEC = 2048;

__device__ __forceinline__ float spin_work(float x, int iters) {
    for (int i = 0; i < iters; i++) x = __fadd_rn(x, 1.0f);
    return x;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void red_release(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void red_relaxed(unsigned* p) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// the body every link runs between its wait and its signal: read the whole vector, dependent work, write my slice
__device__ __forceinline__ void link_body(const float* vin, float* vout, int iters) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < VEC; i += THREADS) acc += __ldcg(vin + i);
    acc = spin_work(acc * 1e-30f, iters);
    const int per = (VEC + CTAS - 1) / CTAS;
    if (threadIdx.x < per && blockIdx.x * per + threadIdx.x < VEC) vout[blockIdx.x * per + threadIdx.x] = acc;
}

__global__ void k_pdl(const float* vin, float* vout, int iters) {
    extern __shared__ float sm[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    link_body(vin, vout, iters);
}
// mode bits: 1 = nanosleep(32) back-off, 2 = relaxed poll + acquire fence afterwards, 4 = threadfence + relaxed red,
// 8 = per-CTA flag words polled by a warp
__global__ void k_flag(const float* vin, float* vout, int iters, unsigned* ctr_prev, unsigned* ctr_mine, unsigned target, int mode, int first) {
    extern __shared__ float sm[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (first) asm volatile("griddepcontrol.wait;" ::: "memory");
    else if (mode & 8) {
        if (threadIdx.x < 32) {
            bool ok;
            long long t0 = clock64();
            do {
                ok = true;
                for (int i = threadIdx.x; i < CTAS; i += 32) ok = ok && (ld_acquire(ctr_prev + i) >= target);
                ok = __all_sync(0xffffffffu, ok);
                if (clock64() - t0 > 4000000000LL) __trap();
            } while (!ok);
        }
    } else if (threadIdx.x == 0) {
        long long t0 = clock64();
        if (mode & 2) { while (ld_relaxed(ctr_prev) < target) { if (mode & 1) __nanosleep(32); if (clock64() - t0 > 4000000000LL) __trap(); } __threadfence(); }
        else { while (ld_acquire(ctr_prev) < target) { if (mode & 1) __nanosleep(32); if (clock64() - t0 > 4000000000LL) __trap(); } }
    }
    __syncthreads();
    link_body(vin, vout, iters);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (mode & 8) st_release(ctr_mine + blockIdx.x, target);
        else if (mode & 4) { __threadfence(); red_relaxed(ctr_mine); }
        else red_release(ctr_mine);
    }
}
// persistent: all links inside one kernel; barrier = central counter (mode 0) or per-CTA flag words (mode 8)
__global__ void k_persist(float* va, float* vb, int iters, unsigned* ctr, unsigned base, int mode, int links) {
    extern __shared__ float sm[];
    for (int l = 0; l < links; l++) {
        link_body((l & 1) ? vb : va, (l & 1) ? va : vb, iters);
        __syncthreads();
        const unsigned target = base + l + 1;
        if (mode & 8) {
            if (threadIdx.x == 0) st_release(ctr + blockIdx.x, target);
            if (threadIdx.x < 32) {
                bool ok;
                long long t0 = clock64();
                do {
                    ok = true;
                    for (int i = threadIdx.x; i < CTAS; i += 32) ok = ok && (ld_acquire(ctr + i) >= target);
                    ok = __all_sync(0xffffffffu, ok);
                    if (clock64() - t0 > 4000000000LL) __trap();
                } while (!ok);
            }
        } else if (threadIdx.x == 0) {
            red_release(ctr);
            long long t0 = clock64();
            while (ld_acquire(ctr) < target * CTAS) { if (clock64() - t0 > 4000000000LL) __trap(); }
        }
        __syncthreads();
    }
}

template <typename F> static float time_chain(cudaStream_t st, int reps, F launch_chain) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch_chain(); launch_chain(); cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; r++) launch_chain();
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / reps;
}

int main() {
    float *va, *vb; unsigned* ctrs;
    cudaMalloc(&va, VEC * 4); cudaMalloc(&vb, VEC * 4); cudaMemset(va, 0, VEC * 4); cudaMemset(vb, 0, VEC * 4);
    cudaMalloc(&ctrs, (LINKS + 1) * CTAS * 4); cudaMemset(ctrs, 0, (LINKS + 1) * CTAS * 4);
    cudaStream_t st; cudaStreamCreate(&st);
    for (size_t smem : {(size_t)0, (size_t)100 * 1024}) {
        cudaFuncSetAttribute(k_pdl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_flag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        auto launch = [&](auto kernel, bool pdl, auto... args) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(CTAS); cfg.blockDim = dim3(THREADS); cfg.stream = st; cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
            cudaLaunchKernelEx(&cfg, kernel, args...);
        };
        printf("---- dynamic smem per CTA: %zu KB ----\n", smem / 1024);
        const float work = time_chain(st, 10, [&] { launch(k_persist, false, va, vb, ITERS * LINKS, ctrs, 0u, 0, 0); launch(k_pdl, false, (const float*)va, vb, ITERS * LINKS); }) ;
        printf("work alone (one kernel, %d x %d adds)                 %8.1f us\n", LINKS, ITERS, work);
        const float plain = time_chain(st, 10, [&] { for (int i = 0; i < LINKS; i++) launch(k_pdl, false, (const float*)((i & 1) ? vb : va), (i & 1) ? va : vb, ITERS); });
        printf("plain stream order                                      %8.1f us -> %.2f us per link beyond the work\n", plain, (plain - work) / LINKS);
        const float pdl = time_chain(st, 10, [&] { for (int i = 0; i < LINKS; i++) launch(k_pdl, i > 0, (const float*)((i & 1) ? vb : va), (i & 1) ? va : vb, ITERS); });
        printf("programmatic dependent launch + griddepcontrol.wait     %8.1f us -> %.2f us per link\n", pdl, (pdl - work) / LINKS);
        unsigned epoch = 0;
        for (int mode : {1, 0, 2, 4, 8}) {
            const float c = time_chain(st, 10, [&] {
                epoch++;
                for (int i = 0; i < LINKS; i++) {
                    const unsigned target = (mode & 8) ? epoch : epoch * CTAS;
                    launch(k_flag, i > 0, (const float*)((i & 1) ? vb : va), (i & 1) ? va : vb, ITERS, ctrs + (size_t)i * CTAS, ctrs + (size_t)(i + 1) * CTAS, target, mode, i == 0 ? 1 : 0);
                }
            });
            printf("flag hand-over mode %d %-30s  %8.1f us -> %.2f us per link\n", mode,
                   mode == 1 ? "(acquire poll + nanosleep 32)" : mode == 0 ? "(acquire poll, pure spin)" : mode == 2 ? "(relaxed poll + fence)" : mode == 4 ? "(threadfence + relaxed red)" : "(per-CTA words, warp poll)",
                   c, (c - work) / LINKS);
        }
        cudaMemset(ctrs, 0, (LINKS + 1) * CTAS * 4);
        unsigned base = 0;
        for (int mode : {0, 8}) {
            base = 0; cudaMemset(ctrs, 0, (LINKS + 1) * CTAS * 4);
            const float c = time_chain(st, 10, [&] { launch(k_persist, false, va, vb, ITERS, ctrs, base, mode, LINKS); base += LINKS; });
            printf("persistent kernel, grid barrier mode %d %-14s %8.1f us -> %.2f us per link\n", mode, mode ? "(per-CTA words)" : "(one counter)", c, (c - work) / LINKS);
        }
        printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    }
    return 0;
}
