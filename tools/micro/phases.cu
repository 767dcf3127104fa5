// micro-benchmark (developer probe): cycles of the latency-bound building blocks of the decode kernels on one SM,
// using the product headers.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false
//        -I lm.rs_b200/csrc -o tools/micro/phases tools/micro/phases.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "attention.cuh"
using namespace lmrs;

constexpr int T = 577, NH = 4, SCS = 644, NTHR = 256, DS = 8;

// ---- rejected variants, kept here so the measurement can be repeated (both measured SLOWER than what ships) --------
// pure-add strided chain: 11.8 cycles per element against 6.9 for serial_av_f32's multiply-in-chain form
template <int VS>
__device__ float serial_sum_strided_f32(const float* __restrict__ base, const int T) {
    float sum = 0.0f;
    int t = 0;
    if (T >= 16) {
        float r[3][8];
#pragma unroll
        for (int u = 0; u < 8; u++) { r[0][u] = base[u * VS]; r[1][u] = base[(8 + u) * VS]; }
        for (; t + 40 <= T; t += 24) {
#pragma unroll
            for (int s = 0; s < 3; s++) {
                const int nx = (s + 2) % 3;
                const float* pn = base + (size_t)(t + 8 * s + 16) * VS;
#pragma unroll
                for (int u = 0; u < 8; u++) r[nx][u] = pn[u * VS];
#pragma unroll
                for (int u = 0; u < 8; u++) sum = __fadd_rn(sum, r[s][u]);
            }
        }
#pragma unroll
        for (int s = 0; s < 2; s++)
#pragma unroll
            for (int u = 0; u < 8; u++) sum = __fadd_rn(sum, r[s][u]);
        t += 16;
    }
    for (; t < T; t++) sum = __fadd_rn(sum, base[(size_t)t * VS]);
    return sum;
}
// rmsnorm over squares formed beforehand: 3056 cycles against 2858 for exact_rnorm (n = 2048)
__device__ float exact_rnorm_sq(const float* xsq, int n, float eps, float* red) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float s = lane < 8 ? serial_sum_strided_f32<8>(xsq + lane, n / 8) : 0.0f;
        const float t = __fadd_rn(s, __shfl_sync(0xffffffffu, s, (lane + 4) & 31));
        const float u = __fadd_rn(t, __shfl_sync(0xffffffffu, t, (lane + 2) & 31));
        float ss = __fadd_rn(u, __shfl_sync(0xffffffffu, u, (lane + 1) & 31));
        if (lane == 0) { ss = __fdiv_rn(ss, (float)n); ss = __fadd_rn(ss, eps); red[0] = __fdiv_rn(1.0f, __fsqrt_rn(ss)); }
    }
    __syncthreads();
    const float r = red[0];
    __syncthreads();
    return r;
}

// exp variants over sc[NH][SCS]
template <int V> __device__ void exp_phase(float* sc_s, const float* mxs, const uint64_t* tab) {
    const int tid = threadIdx.x;
    if (V == 0) {
        for (int h = 0; h < NH; h++) {
            float* sc = sc_s + h * SCS; const float mx = mxs[h];
            for (int t = tid; t < T; t += NTHR) sc[t] = expf_glibc_t(__fsub_rn(sc[t], mx), tab);
        }
    } else {   // flattened (h, t) space, V independent evaluations in flight per thread
        const int total = NH * T;
        for (int i0 = tid; i0 < total; i0 += NTHR * V) {
            constexpr int W = V > 0 ? V : 1; float x[W]; int at[W];
#pragma unroll
            for (int u = 0; u < V; u++) {
                const int i = i0 + u * NTHR;
                const int h = i / T, t = i - h * T;
                at[u] = i < total ? h * SCS + t : -1;
                x[u] = i < total ? __fsub_rn(sc_s[at[u]], mxs[h]) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < V; u++) x[u] = expf_glibc_t(x[u], tab);
#pragma unroll
            for (int u = 0; u < V; u++) if (at[u] >= 0) sc_s[at[u]] = x[u];
        }
    }
}
template <int V> __device__ void div_phase(float* sc_s, const float* sums) {
    const int tid = threadIdx.x;
    if (V == 0) {
        for (int h = 0; h < NH; h++) {
            float* sc = sc_s + h * SCS; const float sum = sums[h];
            for (int t = tid; t < T; t += NTHR) sc[t] = __fdiv_rn(sc[t], sum);
        }
    } else {
        const int total = NH * T;
        for (int i0 = tid; i0 < total; i0 += NTHR * V) {
            constexpr int W = V > 0 ? V : 1; float x[W]; int at[W];
#pragma unroll
            for (int u = 0; u < V; u++) {
                const int i = i0 + u * NTHR;
                const int h = i / T, t = i - h * T;
                at[u] = i < total ? h * SCS + t : -1;
                x[u] = i < total ? __fdiv_rn(sc_s[at[u]], sums[h]) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < V; u++) if (at[u] >= 0) sc_s[at[u]] = x[u];
        }
    }
}
// rmsnorm chain variant with statically rotated register sets (candidate replacement of exact_rnorm's inner loop)
__device__ float rnorm_chain_v1(const float* xf, int n) {
    const int lane = threadIdx.x;
    float s = 0.0f;
    const int steps = n / 8, nb = steps / 8;
    int b = 0;
    if (nb >= 2) {
        float r[3][8];
#pragma unroll
        for (int u = 0; u < 8; u++) { r[0][u] = xf[8 * u + lane]; r[1][u] = xf[8 * (8 + u) + lane]; }
        for (; b + 5 <= nb; b += 3) {
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int nx = (q + 2) % 3;
#pragma unroll
                for (int u = 0; u < 8; u++) r[nx][u] = xf[8 * ((b + q + 2) * 8 + u) + lane];
#pragma unroll
                for (int u = 0; u < 8; u++) s = __fadd_rn(s, __fmul_rn(r[q][u], r[q][u]));
            }
        }
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int u = 0; u < 8; u++) s = __fadd_rn(s, __fmul_rn(r[q][u], r[q][u]));
        b += 2;
    }
    for (int j = b * 8; j < steps; j++) { const float x = xf[8 * j + lane]; s = __fadd_rn(s, __fmul_rn(x, x)); }
    return s;
}

__global__ void __launch_bounds__(NTHR) k(float* out, long long* cyc, float x0) {
    extern __shared__ __align__(16) float big[];   // [640 * 32]
    __shared__ __align__(16) float sc_s[NH * SCS];
    __shared__ __align__(16) float vt[640 * DS];
    __shared__ __align__(16) float xf[4096];
    __shared__ float red[64];
    __shared__ uint64_t tab[32];
    const int tid = threadIdx.x;
    auto fill = [&]() {
        for (int i = tid; i < NH * SCS; i += NTHR) sc_s[i] = x0 * (float)((i * 37) % 101) - 3.0f;
        for (int i = tid; i < 640 * DS; i += NTHR) vt[i] = x0 * (float)((i * 13) % 17) - 0.5f;
        for (int i = tid; i < 640 * 32; i += NTHR) big[i] = x0 * (float)((i * 11) % 19) - 0.4f;
        for (int i = tid; i < 4096; i += NTHR) xf[i] = x0 * (float)((i * 7) % 23) - 0.7f;
        if (tid < 32) tab[tid] = kExp2fTab[tid];
        if (tid < 8) red[tid] = 5.0f + tid;
        __syncthreads();
    };
    int slot = 0;
    long long t0;
    float sink = 0.0f;
#define BEGIN() fill(); __syncthreads(); t0 = clock64();
#define END() __syncthreads(); if (tid == 0) cyc[slot] = clock64() - t0; slot++;
    BEGIN(); exp_phase<0>(sc_s, red, tab); END();          // 0
    BEGIN(); exp_phase<2>(sc_s, red, tab); END();          // 1
    BEGIN(); exp_phase<4>(sc_s, red, tab); END();          // 2
    BEGIN(); exp_phase<8>(sc_s, red, tab); END();          // 3
    BEGIN(); div_phase<0>(sc_s, red); END();               // 4
    BEGIN(); div_phase<4>(sc_s, red); END();               // 5
    BEGIN(); div_phase<8>(sc_s, red); END();               // 6
    BEGIN(); if (tid % 32 == 0 && tid / 32 < NH) sink += serial_sum_f32(sc_s + (tid / 32) * SCS, T); END();   // 7
    BEGIN(); if (tid < NH * DS) sink += serial_av_f32<DS>(sc_s + (tid / DS) * SCS, vt + tid % DS, T); END(); // 8
    BEGIN(); sink += exact_rnorm(xf, 2048, 1e-5f, red); END();   // 9
    BEGIN(); if (tid < 8) sink += rnorm_chain_v1(xf, 2048); END();   // 10
    BEGIN(); sink += exact_rnorm(xf, 4096, 1e-5f, red); END();   // 11
    BEGIN(); if (tid < 8) sink += rnorm_chain_v1(xf, 4096); END();   // 12
    BEGIN(); {   // 13: max phase as in the kernel
        for (int h = 0; h < NH; h++) {
            const float* sc = sc_s + h * SCS; float mx = sc[0];
            for (int t = tid; t < T; t += NTHR) mx = fmaxf(mx, sc[t]);
            mx = warp_max(mx);
            if ((tid & 31) == 0) red[8 + h * 8 + (tid >> 5)] = mx;
        }
    } END();
    BEGIN(); {   // 14: quantize 2048 values (one float4 chunk per thread and pass, as the norm prologue does)
        uint8_t* xq = reinterpret_cast<uint8_t*>(vt); float* xs = red; int* xsum = reinterpret_cast<int*>(red + 32);
        for (int k2 = 0; k2 < 2; k2++) {
            const int c = tid + k2 * NTHR;
            const float4 y = reinterpret_cast<const float4*>(xf)[c];
            quantize_group_to_smem<1>(y, c >> 5, xq, xs, xsum, 2048);
        }
    } END();
    BEGIN(); sink += exact_rnorm_sq(xf, 2048, 1e-5f, red); END();   // 15
    BEGIN(); {   // 16: in-place a*v for one head (DS = 32 dims) + strided pure-add chains
        constexpr int DS2 = 32, DC2 = 8;
        for (int e = tid; e < T * DC2; e += NTHR) {
            const int t = e / DC2; float4* v4 = reinterpret_cast<float4*>(big) + e; const float a = sc_s[t]; float4 v = *v4;
            v.x = __fmul_rn(a, v.x); v.y = __fmul_rn(a, v.y); v.z = __fmul_rn(a, v.z); v.w = __fmul_rn(a, v.w); *v4 = v;
        }
        __syncthreads();
        if (tid == 0) cyc[30] = clock64() - t0;
        if (tid < DS2) sink += serial_sum_strided_f32<DS2>(big + tid, T);
    } END();
    BEGIN(); if (tid < 32) sink += serial_av_f32<32>(sc_s, big + tid, T); END();   // 17: one head, 32 dims, multiply-in-chain
    BEGIN(); {   // 18: the score loop of attn_cluster_kernel<64>: 76 rows x 2 head pairs, K tile rotated, no pushes
        constexpr int HS = 64, C4 = 16; const int myrows = 76, nh = 4, npair = 2;
        const float* q_s = xf; const float* kt = big;
        for (int idx = tid; idx < myrows * npair; idx += NTHR) {
            const int hp = idx / myrows, r = idx - hp * myrows;
            const int ha = hp * 2, hb = min(hp * 2 + 1, nh - 1);
            const float4* qa = reinterpret_cast<const float4*>(q_s + ha * HS);
            const float4* qb = reinterpret_cast<const float4*>(q_s + hb * HS);
            const float4* k4 = reinterpret_cast<const float4*>(kt + r * HS);
            float sa = 0.0f, sb = 0.0f;
            float4 kr[C4];
            { int c = r % C4;
#pragma unroll
              for (int d4 = 0; d4 < C4; d4++) { kr[d4] = k4[c]; c = (c + 1 == C4) ? 0 : c + 1; } }
            float4 q0 = qa[0], q1 = qb[0];
#pragma unroll
            for (int d4 = 0; d4 < C4; d4++) {
                const float4 kv = kr[d4];
                const float a0 = __fmul_rn(q0.x, kv.x), a1 = __fmul_rn(q0.y, kv.y), a2 = __fmul_rn(q0.z, kv.z), a3 = __fmul_rn(q0.w, kv.w);
                const float b0 = __fmul_rn(q1.x, kv.x), b1 = __fmul_rn(q1.y, kv.y), b2 = __fmul_rn(q1.z, kv.z), b3 = __fmul_rn(q1.w, kv.w);
                if (d4 + 1 < C4) { q0 = qa[d4 + 1]; q1 = qb[d4 + 1]; }
                sa = __fadd_rn(sa, a0); sb = __fadd_rn(sb, b0); sa = __fadd_rn(sa, a1); sb = __fadd_rn(sb, b1);
                sa = __fadd_rn(sa, a2); sb = __fadd_rn(sb, b2); sa = __fadd_rn(sa, a3); sb = __fadd_rn(sb, b3);
            }
            sc_s[ha * SCS + r] = __fdiv_rn(sa, 8.0f);
            sc_s[hb * SCS + r] = __fdiv_rn(sb, 8.0f);
        }
    } END();
    BEGIN(); {   // 19: quantize 2048 values, quotient by reciprocal with an exact fallback near rounding boundaries
        uint8_t* xq = reinterpret_cast<uint8_t*>(vt); float* xs = red;
        for (int k2 = 0; k2 < 2; k2++) {
            const int c = tid + k2 * NTHR;
            const float4 y = reinterpret_cast<const float4*>(xf)[c];
            float m = fmaxf(fmaxf(fabsf(y.x), fabsf(y.y)), fmaxf(fabsf(y.z), fabsf(y.w)));
            m = warp_max(m);
            const float scale = __fdiv_rn(m, 127.0f);
            const float inv = __frcp_rn(scale);
            float qv[4] = {y.x * inv, y.y * inv, y.z * inv, y.w * inv};
            const float yy[4] = {y.x, y.y, y.z, y.w};
            int code[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float fr = fabsf(qv[u]) - floorf(fabsf(qv[u]));
                if (fabsf(fr - 0.5f) < 1e-3f || !(fabsf(qv[u]) < 200.0f)) qv[u] = __fdiv_rn(yy[u], scale);   // rare: exact quotient
                code[u] = round_sat_i8(qv[u]);
            }
            reinterpret_cast<uint32_t*>(xq + (size_t)(c >> 5) * 128)[tid & 31] =
                (uint32_t)(code[0] & 0xff) | ((uint32_t)(code[1] & 0xff) << 8) | ((uint32_t)(code[2] & 0xff) << 16) | ((uint32_t)(code[3] & 0xff) << 24);
            if ((tid & 31) == 0) xs[c >> 5] = scale;
        }
    } END();
    BEGIN(); {   // 20: score loop, K row pulled into registers with volatile ld.shared.v4 before the chains start
        constexpr int HS = 64, C4 = 16; const int myrows = 76, nh = 4, npair = 2;
        const float* q_s = xf; const float* kt = big;
        for (int idx = tid; idx < myrows * npair; idx += NTHR) {
            const int hp = idx / myrows, r = idx - hp * myrows;
            const int ha = hp * 2, hb = min(hp * 2 + 1, nh - 1);
            const float4* qa = reinterpret_cast<const float4*>(q_s + ha * HS);
            const float4* qb = reinterpret_cast<const float4*>(q_s + hb * HS);
            float4 kr[C4];
            { int c = r % C4;
#pragma unroll
              for (int d4 = 0; d4 < C4; d4++) {
                  const uint32_t a = smem_u32(kt + r * HS + c * 4);
                  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(kr[d4].x), "=f"(kr[d4].y), "=f"(kr[d4].z), "=f"(kr[d4].w) : "r"(a));
                  c = (c + 1 == C4) ? 0 : c + 1;
              } }
            float sa = 0.0f, sb = 0.0f;
#pragma unroll
            for (int d4 = 0; d4 < C4; d4++) {
                const float4 kv = kr[d4], q0 = qa[d4], q1 = qb[d4];
                sa = __fadd_rn(sa, __fmul_rn(q0.x, kv.x)); sb = __fadd_rn(sb, __fmul_rn(q1.x, kv.x));
                sa = __fadd_rn(sa, __fmul_rn(q0.y, kv.y)); sb = __fadd_rn(sb, __fmul_rn(q1.y, kv.y));
                sa = __fadd_rn(sa, __fmul_rn(q0.z, kv.z)); sb = __fadd_rn(sb, __fmul_rn(q1.z, kv.z));
                sa = __fadd_rn(sa, __fmul_rn(q0.w, kv.w)); sb = __fadd_rn(sb, __fmul_rn(q1.w, kv.w));
            }
            sc_s[ha * SCS + r] = __fdiv_rn(sa, 8.0f);
            sc_s[hb * SCS + r] = __fdiv_rn(sb, 8.0f);
        }
    } END();
    BEGIN(); {   // 21: score loop, thread = (row, head): one chain per thread, all 8 warps busy (304 items)
        constexpr int HS = 64, C4 = 16; const int myrows = 76, nh = 4;
        const float* q_s = xf; const float* kt = big;
        for (int idx = tid; idx < myrows * nh; idx += NTHR) {
            const int h = idx / myrows, r = idx - h * myrows;
            const float4* qa = reinterpret_cast<const float4*>(q_s + h * HS);
            const float4* k4 = reinterpret_cast<const float4*>(kt + r * HS);
            float sa = 0.0f;
            int c = r % C4;
#pragma unroll
            for (int d4 = 0; d4 < C4; d4++) {
                const float4 kv = k4[c], q0 = qa[d4];
                c = (c + 1 == C4) ? 0 : c + 1;
                sa = __fadd_rn(sa, __fmul_rn(q0.x, kv.x)); sa = __fadd_rn(sa, __fmul_rn(q0.y, kv.y));
                sa = __fadd_rn(sa, __fmul_rn(q0.z, kv.z)); sa = __fadd_rn(sa, __fmul_rn(q0.w, kv.w));
            }
            sc_s[h * SCS + r] = __fdiv_rn(sa, 8.0f);
        }
    } END();
    BEGIN(); {   // 22: score loop, q in registers (thread = row, head pair fixed per warp half): K via LDS, no q LDS in the loop
        constexpr int HS = 64, C4 = 16; const int myrows = 76, nh = 4, npair = 2;
        const float* q_s = xf; const float* kt = big;
        for (int idx = tid; idx < myrows * npair; idx += NTHR) {
            const int hp = idx / myrows, r = idx - hp * myrows;
            const int ha = hp * 2, hb = min(hp * 2 + 1, nh - 1);
            const float4* k4 = reinterpret_cast<const float4*>(kt + r * HS);
            float sa = 0.0f, sb = 0.0f;
            int c = r % C4;
#pragma unroll 4
            for (int d4 = 0; d4 < C4; d4++) {
                const float4 kv = k4[c];
                const float4 q0 = *reinterpret_cast<const float4*>(q_s + ha * HS + d4 * 4), q1 = *reinterpret_cast<const float4*>(q_s + hb * HS + d4 * 4);
                c = (c + 1 == C4) ? 0 : c + 1;
                sa = __fadd_rn(sa, __fmul_rn(q0.x, kv.x)); sb = __fadd_rn(sb, __fmul_rn(q1.x, kv.x));
                sa = __fadd_rn(sa, __fmul_rn(q0.y, kv.y)); sb = __fadd_rn(sb, __fmul_rn(q1.y, kv.y));
                sa = __fadd_rn(sa, __fmul_rn(q0.z, kv.z)); sb = __fadd_rn(sb, __fmul_rn(q1.z, kv.z));
                sa = __fadd_rn(sa, __fmul_rn(q0.w, kv.w)); sb = __fadd_rn(sb, __fmul_rn(q1.w, kv.w));
            }
            sc_s[ha * SCS + r] = __fdiv_rn(sa, 8.0f);
            sc_s[hb * SCS + r] = __fdiv_rn(sb, 8.0f);
        }
    } END();
    out[tid] = sink + sc_s[tid] + vt[tid] + big[tid];
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 4 * NTHR); cudaMalloc(&cyc, 8 * 32); cudaMemset(cyc, 0, 8 * 32);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 640 * 32 * 4);
    long long h0[32], h[32];
    for (int rep = 0; rep < 3; rep++) {
        k<<<1, NTHR, 640 * 32 * 4>>>(out, cyc, 0.01f); cudaDeviceSynchronize();
        if (rep == 0) cudaMemcpy(h0, cyc, sizeof h0, cudaMemcpyDeviceToHost);
    }
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const char* names[] = {"exp loop as shipped", "exp ILP2", "exp ILP4", "exp ILP8", "div as shipped", "div ILP4", "div ILP8",
                           "serial_sum_f32 T=577", "serial_av_f32 T=577", "exact_rnorm n=2048", "rnorm chain v1 n=2048",
                           "exact_rnorm n=4096", "rnorm chain v1 n=4096", "max phase", "quantize 2048", "exact_rnorm_sq n=2048",
                           "a*v in place + strided sum", "serial_av_f32<32> T=577", "score loop 76 rows x 4 heads", "quantize 2048 rcp+fallback",
                           "score loop, volatile K row loads", "score loop, one chain per thread", "score loop, rolled (unroll 4)"};
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 23; i++) printf("%-30s %8lld cycles   (first launch, cold: %lld)\n", names[i], h[i], h0[i]);
    printf("  (of test 16: in-place scaling alone %lld cycles)\n", h[30]);
    return 0;
}
