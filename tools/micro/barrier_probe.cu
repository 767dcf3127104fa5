// micro-benchmark (developer probe, round 2): where does the time of a grid-wide hand-over go?  One persistent kernel,
// 148 CTAs x 256 threads, LINKS rounds of { read the whole 2048-float vector | ~1 us dependent work | write my slice |
// grid barrier }, thread 0 of every CTA stamps clock64 after each part.  Variants of the read and of the barrier.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/barrier_probe tools/micro/barrier_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int CTAS = 148, THREADS = 256, LINKS = 64, ITERS = 420, VEC = 2048, NST = 8;

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
__device__ __forceinline__ unsigned ld_volatile(const unsigned* p) {
    unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void red_release(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void red_relaxed(unsigned* p) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ void fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// rmode: 0 = 256 threads x ld.cg, 1 = one 8 KB bulk copy (TMA) into shared memory + everyone reads shared, 2 = only warp 0 reads (ld.cg x 16 float4)
// bmode: 0 = red.release + ld.acquire spin, 1 = fence + red.relaxed + ld.relaxed spin + fence, 2 = as 1 with ld.volatile,
//        3 = st.release own word + warp polls all words relaxed + fence
__global__ void __launch_bounds__(THREADS, 1) k_probe(float* va, float* vb, unsigned* ctr, unsigned base, int rmode, int bmode, long long* stamps) {
    __shared__ __align__(128) float vs[VEC];
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    for (int l = 0; l < LINKS; l++) {
        const float* vin = (l & 1) ? vb : va;
        float* vout = (l & 1) ? va : vb;
        long long* st = stamps + ((size_t)blockIdx.x * LINKS + l) * NST;
        if (tid == 0) st[0] = clock64();
        float acc = 0.f;
        if (rmode == 0) {
#pragma unroll
            for (int i = 0; i < VEC / THREADS; i++) acc += __ldcg(vin + tid + i * THREADS);
        } else if (rmode == 1) {
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(VEC * 4) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(vs)), "l"(vin), "r"(VEC * 4), "r"(smem_u32(&mbar)) : "memory");
            }
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&mbar)), "r"((unsigned)(l & 1)) : "memory");
#pragma unroll
            for (int i = 0; i < VEC / THREADS; i++) acc += vs[tid + i * THREADS];
        } else {
            if (tid < 32) {
#pragma unroll
                for (int i = 0; i < VEC / 128; i++) { float4 t = __ldcg(reinterpret_cast<const float4*>(vin) + tid + i * 32); reinterpret_cast<float4*>(vs)[tid + i * 32] = t; }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < VEC / THREADS; i++) acc += vs[tid + i * THREADS];
        }
        acc = acc * 1e-30f;
        if (tid == 0) st[1] = clock64() + (long long)(acc > 1e30f);
        acc = spin_work(acc, ITERS);
        if (tid == 0) st[2] = clock64();
        const int per = (VEC + CTAS - 1) / CTAS;
        if (tid < per && blockIdx.x * per + tid < VEC) vout[blockIdx.x * per + tid] = acc;
        __syncthreads();
        if (tid == 0) st[3] = clock64();
        const unsigned round = base + l + 1;
        if (bmode == 3) {
            if (tid == 0) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ctr + blockIdx.x), "r"(round) : "memory"); st[4] = clock64(); }
            if (tid < 32) {
                bool ok;
                long long t0 = clock64();
                do {
                    ok = true;
                    for (int i = tid; i < CTAS; i += 32) ok = ok && (ld_relaxed(ctr + i) >= round);
                    ok = __all_sync(0xffffffffu, ok);
                    if (clock64() - t0 > 4000000000LL) __trap();
                } while (!ok);
                fence_acq_rel();
            }
        } else if (tid == 0) {
            const unsigned target = round * CTAS;
            long long t0 = clock64();
            if (bmode == 0) {
                red_release(ctr);
                st[4] = clock64();
                while (ld_acquire(ctr) < target) { if (clock64() - t0 > 4000000000LL) __trap(); }
            } else {
                fence_acq_rel();
                red_relaxed(ctr);
                st[4] = clock64();
                if (bmode == 1) { while (ld_relaxed(ctr) < target) { if (clock64() - t0 > 4000000000LL) __trap(); } }
                else { while (ld_volatile(ctr) < target) { if (clock64() - t0 > 4000000000LL) __trap(); } }
                fence_acq_rel();
            }
        }
        if (tid == 0) st[5] = clock64();
        __syncthreads();
        if (tid == 0) st[6] = clock64();
    }
}

int main() {
    float *va, *vb; unsigned* ctr; long long* stamps;
    cudaMalloc(&va, VEC * 4); cudaMalloc(&vb, VEC * 4); cudaMemset(va, 0, VEC * 4); cudaMemset(vb, 0, VEC * 4);
    cudaMalloc(&ctr, CTAS * 4);
    cudaMalloc(&stamps, (size_t)CTAS * LINKS * NST * 8);
    std::vector<long long> h((size_t)CTAS * LINKS * NST);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("columns: total us per link (events) | cycles: read, work, store+sync, signal issue, poll wait (mean / max over CTAs), final sync\n");
    for (int rmode = 0; rmode < 3; rmode++)
        for (int bmode = 0; bmode < 4; bmode++) {
            cudaMemset(ctr, 0, CTAS * 4);
            unsigned base = 0;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int w = 0; w < 2; w++) { k_probe<<<CTAS, THREADS>>>(va, vb, ctr, base, rmode, bmode, stamps); base += LINKS; }
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            for (int r = 0; r < 5; r++) { k_probe<<<CTAS, THREADS>>>(va, vb, ctr, base, rmode, bmode, stamps); base += LINKS; }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("rmode %d bmode %d: %s\n", rmode, bmode, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h.data(), stamps, h.size() * 8, cudaMemcpyDeviceToHost);
            double s[6] = {0, 0, 0, 0, 0, 0}, pollmax = 0;
            for (int l = 8; l < LINKS; l++) {
                double mx = 0;
                for (int c = 0; c < CTAS; c++) {
                    const long long* st = &h[((size_t)c * LINKS + l) * NST];
                    for (int k = 0; k < 6; k++) s[k] += (double)(st[k + 1] - st[k]);
                    mx = mx > (double)(st[5] - st[4]) ? mx : (double)(st[5] - st[4]);
                }
                pollmax += mx;
            }
            const double n = (double)CTAS * (LINKS - 8);
            printf("read %d barrier %d: %6.2f us/link | read %5.0f  work %5.0f  store+sync %4.0f  signal %4.0f  poll %5.0f / %5.0f  sync %4.0f  (clock %d MHz)\n", rmode, bmode,
                   ms * 1e3 / 5 / LINKS, s[0] / n, s[1] / n, s[2] / n, s[3] / n, s[4] / n, pollmax / (LINKS - 8), s[5] / n, clk_khz / 1000);
        }
    return 0;
}
