// micro-benchmark (developer probe, round 2): fence-free "flag in the data" hand-over.  Every element of the exchanged
// vector is a 64-bit word (f32 value | sequence number << 32) written with one relaxed 8-byte store; consumers poll the
// words themselves until they carry the expected sequence number -- no release/acquire fence (each costs ~0.5 us on B200,
// tools/micro/barrier_probe.cu), no counter, no kernel boundary.
//   persist : one persistent kernel, LINKS rounds of { poll-read the whole vector | ~1 us work | LL-store my slice }
//   chain   : LINKS kernels launched early (programmatic dependent launch, no griddepcontrol.wait), one LL buffer per link
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/ll_probe tools/micro/ll_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int CTAS = 148, THREADS = 256, LINKS = 64, ITERS = 420, VEC = 2048, NST = 4;

__device__ __forceinline__ float spin_work(float x, int iters) {
    for (int i = 0; i < iters; i++) x = __fadd_rn(x, 1.0f);
    return x;
}
__device__ __forceinline__ void ll_store(uint64_t* p, float v, uint32_t seq) {
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(((uint64_t)seq << 32) | (uint64_t)__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ void ll_ld2(const uint64_t* p, uint64_t& a, uint64_t& b) {
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// read the whole vector (every thread 8 elements = 4 x 16-byte loads), spinning until every word carries seq
__device__ __forceinline__ float ll_read_all(const uint64_t* v, uint32_t seq) {
    uint64_t w[8];
    bool ok;
    long long t0 = clock64();
    do {
        ok = true;
#pragma unroll
        for (int i = 0; i < 4; i++) ll_ld2(v + 2 * (threadIdx.x + i * THREADS), w[2 * i], w[2 * i + 1]);
#pragma unroll
        for (int i = 0; i < 8; i++) ok = ok && ((uint32_t)(w[i] >> 32) == seq);
        if (clock64() - t0 > 4000000000LL) __trap();
    } while (!ok);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) acc += __uint_as_float((uint32_t)w[i]);
    return acc;
}
__device__ __forceinline__ void ll_write_slice(uint64_t* v, float val, uint32_t seq) {
    const int per = (VEC + CTAS - 1) / CTAS;
    if (threadIdx.x < per && blockIdx.x * per + threadIdx.x < VEC) ll_store(v + blockIdx.x * per + threadIdx.x, val, seq);
}

__global__ void __launch_bounds__(THREADS, 1) k_persist(uint64_t* bufs, uint32_t base, long long* stamps) {
    for (int l = 0; l < LINKS; l++) {
        const uint32_t seq = base + l + 1;
        long long* st = stamps + ((size_t)blockIdx.x * LINKS + l) * NST;
        if (threadIdx.x == 0) st[0] = clock64();
        float acc = 0.f;
        if (l > 0) acc = ll_read_all(bufs + (size_t)(l & 1) * VEC, seq - 1);
        acc *= 1e-30f;
        if (threadIdx.x == 0) st[1] = clock64() + (long long)(acc > 1e30f);
        acc = spin_work(acc, ITERS);
        if (threadIdx.x == 0) st[2] = clock64();
        ll_write_slice(bufs + (size_t)((l + 1) & 1) * VEC, acc, seq);
        if (threadIdx.x == 0) st[3] = clock64();
    }
}
__global__ void __launch_bounds__(THREADS) k_link(const uint64_t* vin, uint64_t* vout, uint32_t seq, int first) {
    extern __shared__ float sm[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    float acc = 0.f;
    if (!first) acc = ll_read_all(vin, seq);
    acc = spin_work(acc * 1e-30f, ITERS);
    ll_write_slice(vout, acc, seq);
}

int main() {
    uint64_t* bufs; long long* stamps;
    cudaMalloc(&bufs, (size_t)(LINKS + 2) * VEC * 8); cudaMemset(bufs, 0, (size_t)(LINKS + 2) * VEC * 8);
    cudaMalloc(&stamps, (size_t)CTAS * LINKS * NST * 8);
    std::vector<long long> h((size_t)CTAS * LINKS * NST);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    {
        uint32_t base = 0;
        for (int w = 0; w < 2; w++) { k_persist<<<CTAS, THREADS, 0, st>>>(bufs, base, stamps); base += LINKS; }
        cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        for (int r = 0; r < 5; r++) { k_persist<<<CTAS, THREADS, 0, st>>>(bufs, base, stamps); base += LINKS; }
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("persist: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        cudaMemcpy(h.data(), stamps, h.size() * 8, cudaMemcpyDeviceToHost);
        double s[3] = {0, 0, 0};
        for (int l = 8; l < LINKS; l++) for (int c = 0; c < CTAS; c++) { const long long* t = &h[((size_t)c * LINKS + l) * NST]; for (int k = 0; k < 3; k++) s[k] += (double)(t[k + 1] - t[k]); }
        const double n = (double)CTAS * (LINKS - 8);
        printf("persistent LL exchange: %6.2f us/link | cycles: poll-read %5.0f  work %5.0f  store %4.0f  -> hand-over %.2f us beyond the work\n",
               ms * 1e3 / 5 / LINKS, s[0] / n, s[1] / n, s[2] / n, ms * 1e3 / 5 / LINKS - s[1] / n / 1965.0);
    }
    for (size_t smem : {(size_t)0, (size_t)100 * 1024}) {
        cudaFuncSetAttribute(k_link, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        uint32_t seq = 1000;
        auto chain = [&] {
            seq++;
            for (int i = 0; i < LINKS; i++) {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(CTAS); cfg.blockDim = dim3(THREADS); cfg.stream = st; cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = attr; cfg.numAttrs = i > 0 ? 1 : 0;
                cudaLaunchKernelEx(&cfg, k_link, (const uint64_t*)(bufs + (size_t)i * VEC), bufs + (size_t)(i + 1) * VEC, seq, i == 0 ? 1 : 0);
            }
        };
        chain(); chain(); cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        for (int r = 0; r < 10; r++) chain();
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("kernel chain, early launch + LL exchange, %3zu KB smem/CTA: %6.2f us/link (work ~0.86 us) : %s\n", smem / 1024, ms * 1e3 / 10 / LINKS,
               cudaGetErrorString(cudaDeviceSynchronize()));
    }
    return 0;
}
