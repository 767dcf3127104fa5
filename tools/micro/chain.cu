// micro-benchmark: cost per step of dependent f32 add chains on sm_100a (developer probe)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_regs(float* out, long long* cyc, float x0) {
    float s = x0; float a[8];
    for (int u = 0; u < 8; u++) a[u] = x0 * (u + 1);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) s = __fadd_rn(s, a[u]);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = s; cyc[0] = t1 - t0; }
}
__global__ void k_lds(float* out, long long* cyc, float x0, int variant) {
    __shared__ float xf[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) xf[i] = x0 * (i % 7);
    __syncthreads();
    if (threadIdx.x >= 32) { return; }
    const int lane = threadIdx.x;
    float s = 0.f;
    long long t0 = clock64();
    if (lane < 8) {
        if (variant == 0) {          // naive
            for (int j = 0; j < 512; j++) { float x = xf[8 * j + lane]; s = __fadd_rn(s, __fmul_rn(x, x)); }
        } else if (variant == 1) {   // products one batch ahead
            float pa[8], pb[8];
            for (int u = 0; u < 8; u++) { float x = xf[8 * u + lane]; pa[u] = __fmul_rn(x, x); }
            for (int b = 1; b < 64; b++) {
#pragma unroll
                for (int u = 0; u < 8; u++) { float x = xf[8 * (b * 8 + u) + lane]; pb[u] = __fmul_rn(x, x); }
#pragma unroll
                for (int u = 0; u < 8; u++) s = __fadd_rn(s, pa[u]);
#pragma unroll
                for (int u = 0; u < 8; u++) pa[u] = pb[u];
            }
#pragma unroll
            for (int u = 0; u < 8; u++) s = __fadd_rn(s, pa[u]);
        } else {                     // squares precomputed in smem: pure add chain, loads one batch ahead
            float pa[8], pb[8];
            for (int u = 0; u < 8; u++) pa[u] = xf[8 * u + lane];
            for (int b = 1; b < 64; b++) {
#pragma unroll
                for (int u = 0; u < 8; u++) pb[u] = xf[8 * (b * 8 + u) + lane];
#pragma unroll
                for (int u = 0; u < 8; u++) s = __fadd_rn(s, pa[u]);
#pragma unroll
                for (int u = 0; u < 8; u++) pa[u] = pb[u];
            }
#pragma unroll
            for (int u = 0; u < 8; u++) s = __fadd_rn(s, pa[u]);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = s; cyc[0] = t1 - t0; }
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    long long h;
    for (int rep = 0; rep < 2; rep++) {
        k_regs<<<1, 32>>>(out, cyc, 1.0f); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("reg chain 512 adds, 1 warp: %lld cycles (%.2f/step)\n", h, h / 512.0);
        for (int nt : {32, 512}) for (int v = 0; v < 3; v++) {
            k_lds<<<1, nt>>>(out, cyc, 1.0f, v); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("lds chain variant %d, block %d: %lld cycles (%.2f/step)\n", v, nt, h, h / 512.0);
        }
    }
    return 0;
}
