// misc.cuh -- small kernels around the hot path: pending-residual finalize for fill_kv_cache and the
// stand-alone operator kernels behind the operator-level C ABI (quantize / rmsnorm / softmax).
#pragma once
#include "common.cuh"
#include "exact_math.cuh"
#include "gemv.cuh"

namespace lmrs {

// x_row = x_in + delta  (Gemma: + rmsnorm(delta, w_post) with unit offset) -- the residual add that closes a
// transformer block (src/transformer.rs:642-656); in the decode chain it is folded into the next GEMV prologue,
// fill_kv_cache needs it materialised because the caller gets the residual stream back (:678).
struct ResidualParams {
    const float* x_in; const float* delta; const float* w_post;
    int n; float eps;
    float* rows;               // [n_rows][n]; row index = step->token
    const StepParams* step;
    int row_from_block;        // batched prefill: row = blockIdx.x and x_in/delta are row-strided too
};
LMRS_DEVINL void residual_finalize_body(const ResidualParams& p, float* red) {
    const size_t row = p.row_from_block ? (size_t)blockIdx.x : (size_t)p.step->token;
    float* out = p.rows + row * p.n;
    const float* x_in = p.x_in + (p.row_from_block ? row * p.n : 0);
    const float* delta = p.delta + (p.row_from_block ? row * p.n : 0);
    float r = 1.0f;
    if (p.w_post) r = exact_rnorm(delta, p.n, p.eps, red);   // chains read global memory directly
    for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
        float d = __ldcg(delta + i);
        if (p.w_post) d = __fmul_rn(__fadd_rn(1.0f, p.w_post[i]), __fmul_rn(r, d));
        out[i] = __fadd_rn(__ldcg(x_in + i), d);
    }
}
__global__ void __launch_bounds__(256) residual_finalize_kernel(const ResidualParams p) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    residual_finalize_body(p, red);
}

// ---- batched prefill row kernels (fill_kv_cache, src/transformer.rs:672-684 -> forward_layer with sl = N) ----------
// one CTA per token row: the GEMV prologue (residual add, exact rmsnorm, activation quantize) with its result written
// to HBM as the int8 A operand + scales of the tcgen05 GEMM
__global__ void __launch_bounds__(256) rows_prologue_kernel(GemvParams p, uint8_t* xq_out, float* xs_out) {
    extern __shared__ __align__(128) uint8_t rsm[];
    const size_t row = blockIdx.x;
    const int n = p.n, G = n / GS;
    GemvSmem sm;
    sm.xq = rsm;
    sm.xs = reinterpret_cast<float*>(rsm + ((n + 127) / 128) * 128);
    sm.xsum = reinterpret_cast<int*>(sm.xs + G);
    sm.red = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sm.xs) + ((G * 8 + 127) / 128) * 128);
    sm.xf = sm.red + 64;
    sm.exp_tab = kExp2fTab;
    if (p.x_in) p.x_in += row * n;
    if (p.delta) p.delta += row * n;
    if (p.x_out) p.x_out += row * n;
    if (p.act_in) p.act_in += row * n;
    p.xout_all = 1;
    gemv_prologue<1, 8>(p, sm);
    for (int i = threadIdx.x; i < n / 16; i += 256) reinterpret_cast<int4*>(xq_out + row * n)[i] = reinterpret_cast<const int4*>(sm.xq)[i];
    for (int g = threadIdx.x; g < G; g += 256) xs_out[row * G + g] = sm.xs[g];
}
// RoPE on the q rows and on the K rows just written into the cache (src/transformer.rs:443-495), exact (host tables)
__global__ void __launch_bounds__(256) rope_rows_kernel(float* q, float* kcache, const float* rope_cos, const float* rope_sin,
                                                        int n_heads, int n_kv_heads, int hs, int pos0) {
    const int row = blockIdx.x, pos = pos0 + row;
    const int half = hs / 2;
    const float* cs = rope_cos + (size_t)pos * half;
    const float* sn = rope_sin + (size_t)pos * half;
    float* qr = q + (size_t)row * n_heads * hs;
    float* kr = kcache + (size_t)pos * n_kv_heads * hs;
    for (int i = threadIdx.x; i < (n_heads + n_kv_heads) * half; i += 256) {
        const int h = i / half, j = i - h * half;
        float* v = h < n_heads ? qr + (size_t)h * hs : kr + (size_t)(h - n_heads) * hs;
        const float fcr = cs[j], fci = sn[j], v0 = v[j], v1 = v[j + half];
        v[j] = __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci));
        v[j + half] = __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr));
    }
}
// h = act(gate) * up (src/transformer.rs:607-624)
__global__ void glu_rows_kernel(float* h, const float* gate, const float* up, size_t count, int epi) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) h[i] = __fmul_rn(glu_act(epi, gate[i]), up[i]);
}

// ---- operator-level kernels (any group size; one thread walks one group serially like the reference) -------
// src/quantization.rs:44-67
__global__ void quantize_q8_kernel(int8_t* q, float* s, const float* x, int n_groups, int gs) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const float* xg = x + (size_t)g * gs;
    float wmax = 0.0f;
    for (int i = 0; i < gs; i++) wmax = fmaxf(wmax, fabsf(xg[i]));
    const float scale = __fdiv_rn(wmax, 127.0f);
    s[g] = scale;
    for (int i = 0; i < gs; i++) q[(size_t)g * gs + i] = (int8_t)round_sat_i8(__fdiv_rn(xg[i], scale));
}
// src/quantization.rs:69-95
__global__ void quantize_q4_kernel(uint8_t* q, float* s, const float* x, int n_groups, int gs) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const float* xg = x + (size_t)g * gs;
    float wmax = 0.0f;
    for (int i = 0; i < gs; i++) wmax = fmaxf(wmax, fabsf(xg[i]));
    const float scale = __fdiv_rn(wmax, -8.0f);
    s[g] = scale;
    for (int i = 0; i < gs / 2; i++) {
        int a = round_sat_u4(__fdiv_rn(xg[2 * i], scale)), b = round_sat_u4(__fdiv_rn(xg[2 * i + 1], scale));
        q[(size_t)g * gs / 2 + i] = (uint8_t)(a | (b << 4));
    }
}
// src/functional.rs:48-78, exact summation order (exact_rnorm)
__global__ void __launch_bounds__(256) rmsnorm_kernel(float* o, const float* x, const float* w, int size, float eps, int unit) {
    __shared__ float red[32];
    const int n8 = size / 8 * 8;
    // exact_rnorm divides by its n argument: the reference divides by `size` but sums n8 elements
    float r;
    {
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            float s = 0.0f;
            if (lane < 8)
                for (int j = 0; j < n8 / 8; j++) { const float v = x[8 * j + lane]; s = __fadd_rn(s, __fmul_rn(v, v)); }
            const float t = __fadd_rn(s, __shfl_sync(0xffffffffu, s, (lane + 4) & 31));
            const float u = __fadd_rn(t, __shfl_sync(0xffffffffu, t, (lane + 2) & 31));
            float ss = __fadd_rn(u, __shfl_sync(0xffffffffu, u, (lane + 1) & 31));
            if (lane == 0) red[0] = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(__fdiv_rn(ss, (float)size), eps)));
        }
        __syncthreads();
        r = red[0];
    }
    for (int i = threadIdx.x; i < n8; i += 256) {
        const float t = __fmul_rn(r, x[i]);
        o[i] = unit ? __fmul_rn(__fadd_rn(1.0f, w[i]), t) : __fmul_rn(w[i], t);
    }
}
// src/functional.rs:122-140: max, exp(x-max) with glibc's expf, SERIAL sum, divide -- bit-exact
__global__ void __launch_bounds__(256) softmax_kernel(float* x, int n) {
    __shared__ float red[32];
    float mx = x[0];
    for (int i = threadIdx.x; i < n; i += 256) mx = fmaxf(mx, x[i]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < 8; w++) mx = fmaxf(mx, red[w]);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256) x[i] = expf_glibc(__fsub_rn(x[i], mx));
    __syncthreads();
    if (threadIdx.x == 0) {
        float sum = 0.0f;
        for (int i = 0; i < n; i++) sum = __fadd_rn(sum, x[i]);
        red[0] = sum;
    }
    __syncthreads();
    const float sum = red[0];
    for (int i = threadIdx.x; i < n; i += 256) x[i] = __fdiv_rn(x[i], sum);
}

}  // namespace lmrs
