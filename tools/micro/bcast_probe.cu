// micro-benchmark (developer probe, round 2): what does it cost when all 148 CTAs read the SAME small vector from L2 at the
// same time (every GEMV prologue of the decode chain does)?  KB = vector size, R = number of replicas at different
// addresses (CTA b reads replica b % R).  Cycles = mean/max over CTAs of the time from a common start (grid barrier) to
// the data being in registers, 256 threads x 16-byte loads.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/bcast_probe tools/micro/bcast_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
constexpr int CTAS = 148, THREADS = 256, ROUNDS = 24;
__device__ __forceinline__ void ld16(const void* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__global__ void __launch_bounds__(THREADS, 1) k(const uint8_t* buf, size_t vec_bytes, int reps, unsigned* ctr, long long* out, unsigned long long* sink) {
    const uint8_t* mine = buf + (size_t)(blockIdx.x % reps) * vec_bytes;
    unsigned long long acc = 0;
    for (int r = 0; r < ROUNDS; r++) {
        __syncthreads();
        if (threadIdx.x == 0) {   // crude grid barrier so that everybody starts reading together
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
            unsigned v; do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < (unsigned)(r + 1) * CTAS);
        }
        __syncthreads();
        const long long t0 = clock64();
        for (size_t off = (size_t)threadIdx.x * 16; off < vec_bytes; off += (size_t)THREADS * 16 * 4) {
            unsigned long long a0 = 0, b0 = 0, a1 = 0, b1 = 0, a2 = 0, b2 = 0, a3 = 0, b3 = 0;
            ld16(mine + off, a0, b0);
            if (off + THREADS * 16 < vec_bytes) ld16(mine + off + THREADS * 16, a1, b1);
            if (off + THREADS * 32 < vec_bytes) ld16(mine + off + THREADS * 32, a2, b2);
            if (off + THREADS * 48 < vec_bytes) ld16(mine + off + THREADS * 48, a3, b3);
            acc += a0 + b0 + a1 + b1 + a2 + b2 + a3 + b3;
        }
        acc = __shfl_xor_sync(0xffffffffu, acc, 1) + acc;
        __syncthreads();
        const long long t1 = clock64();
        if (threadIdx.x == 0) out[(size_t)r * CTAS + blockIdx.x] = t1 - t0;
    }
    if (acc == 0x1234567ull) *sink = acc;
}
int main() {
    uint8_t* buf; unsigned* ctr; long long* out; unsigned long long* sink;
    cudaMalloc(&buf, 64 << 20); cudaMemset(buf, 1, 64 << 20);
    cudaMalloc(&ctr, 4); cudaMalloc(&out, ROUNDS * CTAS * 8); cudaMalloc(&sink, 8);
    std::vector<long long> h(ROUNDS * CTAS);
    printf("  KB  replicas   mean cycles   max cycles  (148 CTAs x 256 threads, 1.965 GHz)\n");
    for (int kb : {2, 4, 8, 16, 32, 64})
        for (int reps : {1, 2, 4, 8, 16, 37, 148}) {
            cudaMemset(ctr, 0, 4);
            k<<<CTAS, THREADS>>>(buf, (size_t)kb * 1024, reps, ctr, out, sink);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
            cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
            double s = 0, mx = 0; int n = 0;
            for (int r = 4; r < ROUNDS; r++) { double m = 0; for (int c = 0; c < CTAS; c++) { s += h[r * CTAS + c]; n++; if (h[r * CTAS + c] > m) m = h[r * CTAS + c]; } mx += m; }
            printf("%4d  %8d   %11.0f  %11.0f\n", kb, reps, s / n, mx / (ROUNDS - 4));
        }
    return 0;
}
