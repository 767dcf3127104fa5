// attention.cuh -- decode-step attention for sm_100a: RoPE + QK^T + soft-cap/window + softmax + .V in ONE
// kernel over the HBM-resident f32 KV cache.  Replaces src/transformer.rs:443-544 for sl = 1.
//
// Grid n_kv_heads * head_chunks: a CTA serves up to 4 query heads that share one KV head (GQA: K rows are
// staged once for all of them).  RoPE uses cos/sin tables computed at load time on the HOST with the same libm
// calls as the reference (powf/cosf/sinf, src/transformer.rs:447-482).  The new K row arrives un-rotated in a
// staging row written by the QKV GEMV; the CTA rotates it, uses it from shared memory and stores it into the
// cache.  All f32 arithmetic follows the reference's operation order exactly (see the kernel comment).
#pragma once
#include "common.cuh"
#include "exact_math.cuh"
#include "gemv.cuh"

namespace lmrs {

constexpr int ATT_THREADS = 256;
constexpr int ATT_QH = 4;      // query heads per CTA (all sharing one KV head)

struct AttnParams {
    const float* q;        // [att_dim] un-rotated
    const float* k_new;    // [kv_dim]  un-rotated K of this step
    float* kcache;         // layer base [seq_len][kv_dim]
    const float* vcache;   // layer base (row `pos` already written by the QKV GEMV)
    const float* rope_cos; // [seq_len][hs/2]
    const float* rope_sin;
    float* out;            // [att_dim]
    float* scores;         // scratch [n_heads][seq_len]: scores -> exp -> probabilities
    int kv_dim, kv_mul, chunks, gemma, seq_len;
    float sqrt_hs;         // sqrtf(head_size): scores are DIVIDED by it (src/transformer.rs:516)
    const StepParams* step;
};

// Bit-exact restatement of src/transformer.rs:501-544 for one token: every f32 operation happens in the
// reference's order (serial dot over d, serial softmax sum over t, serial a*v accumulation over t, separate
// mul and add, exp = glibc expf), only independent chains run in parallel.  One CTA per KV head (x chunks of 4
// query heads): the latency is that of the two T-long dependent add chains (~4 cycles per cached position
// each), which is the price of reproducing the CPU path's rounding exactly -- see exact_math.cuh for why.
template <int HS>
__global__ void __launch_bounds__(ATT_THREADS) attn_decode_kernel(const AttnParams p) {
    constexpr int ATT_TILE = HS > 128 ? 32 : 64;   // cached positions staged per K tile (static smem <= 48 KB)
    __shared__ float q_s[ATT_QH][HS];
    __shared__ float k_s[HS];
    __shared__ float ktile[ATT_TILE][HS + 1];
    __shared__ float red[ATT_THREADS / 32];
    __shared__ float stat[ATT_QH];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kvh = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks;
    const int h0 = kvh * p.kv_mul + chunk * ATT_QH;                  // first query head of this CTA
    const int nh = min(ATT_QH, p.kv_mul - chunk * ATT_QH);           // query heads served here
    pdl_launch_dependents();
    pdl_wait();
    const int pos = (int)p.step->pos;
    const uint32_t mask_base = p.step->mask_base;
    const int T = pos + 1;

    // RoPE on q and on the new k row (rotate-half pairs j, j+HS/2), src/transformer.rs:480-492; tables hold
    // cos/sin(pos*freq)*scale computed on the host with the reference's libm calls.
    const float* cs = p.rope_cos + (size_t)pos * (HS / 2);
    const float* sn = p.rope_sin + (size_t)pos * (HS / 2);
    for (int i = tid; i < nh * (HS / 2); i += ATT_THREADS) {
        const int h = i / (HS / 2), j = i - h * (HS / 2);
        const float fcr = cs[j], fci = sn[j];
        const float v0 = p.q[(size_t)(h0 + h) * HS + j], v1 = p.q[(size_t)(h0 + h) * HS + j + HS / 2];
        q_s[h][j] = __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci));
        q_s[h][j + HS / 2] = __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr));
    }
    for (int j = tid; j < HS / 2; j += ATT_THREADS) {
        const float fcr = cs[j], fci = sn[j];
        const float v0 = p.k_new[(size_t)kvh * HS + j], v1 = p.k_new[(size_t)kvh * HS + j + HS / 2];
        const float r0 = __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci));
        const float r1 = __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr));
        k_s[j] = r0; k_s[j + HS / 2] = r1;
        if (chunk == 0) {   // exactly one CTA per KV head publishes the rotated row into the cache
            p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + j] = r0;
            p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + j + HS / 2] = r1;
        }
    }
    __syncthreads();

    // ---- scores: s[h][t] = (sum_d q[h][d]*k[t][d]) / sqrt(hs)   (:507-528) --------------------------------
    for (int tile0 = 0; tile0 < T; tile0 += ATT_TILE) {
        const int rows = min(ATT_TILE, T - tile0);
        for (int idx = tid; idx < rows * HS; idx += ATT_THREADS) {     // coalesced K rows -> padded smem tile
            const int r = idx / HS, d = idx - r * HS, t = tile0 + r;
            ktile[r][d] = (t == pos) ? k_s[d] : p.kcache[(size_t)t * p.kv_dim + (size_t)kvh * HS + d];
        }
        __syncthreads();
        for (int idx = tid; idx < rows * nh; idx += ATT_THREADS) {
            const int h = idx / rows, r = idx - h * rows, t = tile0 + r;
            float score = 0.0f;
#pragma unroll 8
            for (int d = 0; d < HS; d++) score = __fadd_rn(score, __fmul_rn(q_s[h][d], ktile[r][d]));
            score = __fdiv_rn(score, p.sqrt_hs);
            if (p.gemma) {   // soft-cap 50*tanh(s/50) in f64, window mask on every layer (:518-526)
                score = __fdiv_rn(score, 50.0f);
                score = (float)tanh((double)score);
                score = __fmul_rn(score, 50.0f);
                score = __fadd_rn(score, (mask_base - (uint32_t)t <= 4096u) ? 0.0f : -2.3819763e38f);
            }
            p.scores[(size_t)(h0 + h) * p.seq_len + t] = score;
        }
        __syncthreads();
    }

    // ---- softmax (src/functional.rs:122-140): max, exp(x-max), serial sum, divide ---------------------------
    for (int h = 0; h < nh; h++) {
        float* sc = p.scores + (size_t)(h0 + h) * p.seq_len;
        float mx = sc[0];
        for (int t = tid; t < T; t += ATT_THREADS) mx = fmaxf(mx, sc[t]);
        mx = warp_max(mx);
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = red[0];
#pragma unroll
        for (int w = 1; w < ATT_THREADS / 32; w++) mx = fmaxf(mx, red[w]);
        for (int t = tid; t < T; t += ATT_THREADS) sc[t] = expf_glibc(__fsub_rn(sc[t], mx));
        __syncthreads();
    }
    if (tid < nh) {   // the reference's `sum += x[i]` chain, one thread per head
        const float* sc = p.scores + (size_t)(h0 + tid) * p.seq_len;
        float sum = 0.0f;
        int t = 0;
        for (; t + 8 <= T; t += 8) {
            const float4 a = *reinterpret_cast<const float4*>(sc + t), b = *reinterpret_cast<const float4*>(sc + t + 4);
            sum = __fadd_rn(sum, a.x); sum = __fadd_rn(sum, a.y); sum = __fadd_rn(sum, a.z); sum = __fadd_rn(sum, a.w);
            sum = __fadd_rn(sum, b.x); sum = __fadd_rn(sum, b.y); sum = __fadd_rn(sum, b.z); sum = __fadd_rn(sum, b.w);
        }
        for (; t < T; t++) sum = __fadd_rn(sum, sc[t]);
        stat[tid] = sum;
    }
    __syncthreads();
    for (int h = 0; h < nh; h++) {
        float* sc = p.scores + (size_t)(h0 + h) * p.seq_len;
        const float sum = stat[h];
        for (int t = tid; t < T; t += ATT_THREADS) sc[t] = __fdiv_rn(sc[t], sum);
    }
    __syncthreads();

    // ---- out[h][d] = sum_t a[h][t] * v[t][d], serial over t (:533-542) --------------------------------------
    for (int idx = tid; idx < nh * HS; idx += ATT_THREADS) {
        const int h = idx / HS, d = idx - h * HS;
        const float* sc = p.scores + (size_t)(h0 + h) * p.seq_len;
        const float* vcol = p.vcache + (size_t)kvh * HS + d;
        float acc = 0.0f;
        int t = 0;
        for (; t + 4 <= T; t += 4) {
            const float4 a = *reinterpret_cast<const float4*>(sc + t);
            const float v0 = vcol[(size_t)t * p.kv_dim], v1 = vcol[(size_t)(t + 1) * p.kv_dim];
            const float v2 = vcol[(size_t)(t + 2) * p.kv_dim], v3 = vcol[(size_t)(t + 3) * p.kv_dim];
            acc = __fadd_rn(acc, __fmul_rn(a.x, v0)); acc = __fadd_rn(acc, __fmul_rn(a.y, v1));
            acc = __fadd_rn(acc, __fmul_rn(a.z, v2)); acc = __fadd_rn(acc, __fmul_rn(a.w, v3));
        }
        for (; t < T; t++) acc = __fadd_rn(acc, __fmul_rn(sc[t], vcol[(size_t)t * p.kv_dim]));
        p.out[(size_t)(h0 + h) * HS + d] = acc;
    }
}

// ---- embedding row gather: the reference dequantizes the whole table at load (src/transformer.rs:243-245,
// src/quantization.rs:25-42) and copies a row per token (:324, :659-669); here a row is dequantized on the fly
// (value = code as f32 * scale: one multiply, bit-identical).  Gemma scales by sqrt(dim) (:327-332).
struct EmbedParams {
    const uint8_t* q; const float* s; const float* f32_table;
    int dim, q_type; float scale_mul; int apply_scale;
    const uint32_t* tokens;  // device array (get_embeddings) or nullptr -> step->token
    const StepParams* step;
    float* out;              // [n_tokens][dim]
};
__global__ void embed_kernel(const EmbedParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t tok = p.tokens ? p.tokens[blockIdx.x] : p.step->token;
    float* out = p.out + (size_t)blockIdx.x * p.dim;
    for (int i = threadIdx.x; i < p.dim; i += blockDim.x) {
        const size_t e = (size_t)tok * p.dim + i;
        float v;
        if (p.q_type == 0) v = p.f32_table[e];
        else if (p.q_type == 1) v = __fmul_rn((float)reinterpret_cast<const int8_t*>(p.q)[e], p.s[e / GS]);
        else {
            const int b = p.q[e >> 1];
            const int code = ((e & 1) ? (b >> 4) : (b & 15)) - 8;
            v = __fmul_rn((float)code, p.s[e / GS]);
        }
        if (p.apply_scale) v = __fmul_rn(v, p.scale_mul);
        out[i] = v;
    }
}

}  // namespace lmrs
