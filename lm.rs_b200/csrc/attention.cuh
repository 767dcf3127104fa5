// attention.cuh -- decode-step attention for sm_100a: RoPE + QK^T + soft-cap/window + softmax + .V in ONE
// kernel over the HBM-resident f32 KV cache.  Replaces src/transformer.rs:443-544 for sl = 1.
//
// Grid (n_kv_heads * head_chunks, NSPLIT): a CTA serves up to 4 query heads that share one KV head (GQA:
// K/V rows are read once for all of them) over one slice of the cached positions (split-T, flash-decoding
// style); the last CTA to finish a KV head merges the slices (threadfence + atomic ticket), so no second
// kernel is needed.  RoPE uses cos/sin tables computed at load time on the HOST with the same libm calls as
// the reference (powf/cosf/sinf, src/transformer.rs:447-482), so rotated q/k are bit-identical given
// bit-identical inputs.  The new K row arrives un-rotated in a staging row (written by the QKV GEMV); every
// CTA that needs position `pos` rotates it locally, and exactly one CTA per KV head stores the rotated row
// into the cache (no read/write race between CTAs).
// Softmax / A.V use a parallel, online formulation: summation order differs from the reference's serial
// loops (f32 rounding only; covered by the 1e-3 logits tolerance).
#pragma once
#include "common.cuh"
#include "gemv.cuh"

namespace lmrs {

constexpr int ATT_WARPS = 8;
constexpr int ATT_QH = 4;  // query heads per CTA

struct AttnParams {
    const float* q;        // [att_dim] un-rotated
    const float* k_new;    // [kv_dim]  un-rotated K of this step
    float* kcache;         // layer base [seq_len][kv_dim]
    const float* vcache;   // layer base (row `pos` already written by the QKV GEMV)
    const float* rope_cos; // [seq_len][hs/2]
    const float* rope_sin;
    float* out;            // [att_dim]
    float* part;           // [n_heads][nsplit][hs + 2]   (o[hs], m, l)
    unsigned* tickets;     // [n_kv_heads * chunks]
    int kv_dim, kv_mul, nsplit, chunks, gemma;
    float inv_sqrt_hs_den; // sqrtf(head_size): scores are DIVIDED by it (src/transformer.rs:516)
    const StepParams* step;
};

template <int HS>
__global__ void __launch_bounds__(ATT_WARPS * 32) attn_decode_kernel(const AttnParams p) {
    constexpr int VPL = (HS + 31) / 32;       // dims per lane (strided: d = lane + 32*i)
    __shared__ float q_s[ATT_QH][HS];
    __shared__ float k_s[HS];
    __shared__ float o_s[ATT_WARPS][ATT_QH][HS];
    __shared__ float m_s[ATT_WARPS][ATT_QH], l_s[ATT_WARPS][ATT_QH];
    __shared__ int is_last_cta;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kvh = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks, split = blockIdx.y;
    const int h0 = kvh * p.kv_mul + chunk * ATT_QH;                  // first query head of this CTA
    const int nh = min(ATT_QH, p.kv_mul - chunk * ATT_QH);           // query heads served here
    pdl_launch_dependents();
    pdl_wait();
    const int pos = (int)p.step->pos;
    const uint32_t mask_base = p.step->mask_base;
    const int T = pos + 1;
    int per = (T + p.nsplit - 1) / p.nsplit;
    per = (per + ATT_WARPS - 1) / ATT_WARPS * ATT_WARPS;
    const int t0 = split * per, t1 = min(T, t0 + per);
    const bool owns_pos = (pos >= t0 && pos < t1);

    // RoPE on q (always) and on the new k row (rotate-half pairs j, j+HS/2), src/transformer.rs:480-492
    const float* cs = p.rope_cos + (size_t)pos * (HS / 2);
    const float* sn = p.rope_sin + (size_t)pos * (HS / 2);
    for (int i = tid; i < nh * (HS / 2); i += ATT_WARPS * 32) {
        const int h = i / (HS / 2), j = i - h * (HS / 2);
        const float fcr = cs[j], fci = sn[j];
        const float v0 = p.q[(size_t)(h0 + h) * HS + j], v1 = p.q[(size_t)(h0 + h) * HS + j + HS / 2];
        q_s[h][j] = __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci));
        q_s[h][j + HS / 2] = __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr));
    }
    if (owns_pos) {
        for (int j = tid; j < HS / 2; j += ATT_WARPS * 32) {
            const float fcr = cs[j], fci = sn[j];
            const float v0 = p.k_new[(size_t)kvh * HS + j], v1 = p.k_new[(size_t)kvh * HS + j + HS / 2];
            const float r0 = __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci));
            const float r1 = __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr));
            k_s[j] = r0; k_s[j + HS / 2] = r1;
            if (chunk == 0) {
                p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + j] = r0;
                p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + j + HS / 2] = r1;
            }
        }
    }
    __syncthreads();

    float qr[ATT_QH][VPL], o[ATT_QH][VPL], m[ATT_QH], l[ATT_QH];
#pragma unroll
    for (int h = 0; h < ATT_QH; h++) {
        m[h] = -INFINITY; l[h] = 0.0f;
#pragma unroll
        for (int i = 0; i < VPL; i++) {
            const int d = lane + 32 * i;
            qr[h][i] = (h < nh && d < HS) ? q_s[h][d] : 0.0f;
            o[h][i] = 0.0f;
        }
    }

    for (int t = t0 + warp; t < t1; t += ATT_WARPS) {
        float kr[VPL], vr[VPL];
        const float* vrow = p.vcache + (size_t)t * p.kv_dim + (size_t)kvh * HS;
        if (t == pos) {
#pragma unroll
            for (int i = 0; i < VPL; i++) { const int d = lane + 32 * i; kr[i] = d < HS ? k_s[d] : 0.0f; }
        } else {
            const float* krow = p.kcache + (size_t)t * p.kv_dim + (size_t)kvh * HS;
#pragma unroll
            for (int i = 0; i < VPL; i++) { const int d = lane + 32 * i; kr[i] = d < HS ? krow[d] : 0.0f; }
        }
#pragma unroll
        for (int i = 0; i < VPL; i++) { const int d = lane + 32 * i; vr[i] = d < HS ? vrow[d] : 0.0f; }
#pragma unroll
        for (int h = 0; h < ATT_QH; h++) {
            if (h < nh) {
                float s = 0.0f;
#pragma unroll
                for (int i = 0; i < VPL; i++) s = fmaf(qr[h][i], kr[i], s);
                s = warp_sum(s);
                s = __fdiv_rn(s, p.inv_sqrt_hs_den);
                if (p.gemma) {   // soft-cap 50*tanh(s/50) in f64, window 4096 on every layer (src/transformer.rs:518-526)
                    s = __fdiv_rn(s, 50.0f);
                    s = (float)tanh((double)s);
                    s = __fmul_rn(s, 50.0f);
                    s = __fadd_rn(s, (mask_base - (uint32_t)t <= 4096u) ? 0.0f : -2.3819763e38f);  // u32 wrap kept
                }
                const float mn = fmaxf(m[h], s);
                const float corr = expf(m[h] - mn), pr = expf(s - mn);
                l[h] = fmaf(l[h], corr, pr);
#pragma unroll
                for (int i = 0; i < VPL; i++) o[h][i] = fmaf(o[h][i], corr, pr * vr[i]);
                m[h] = mn;
            }
        }
    }
    // merge the warps of this CTA
#pragma unroll
    for (int h = 0; h < ATT_QH; h++) {
        if (lane == 0) { m_s[warp][h] = m[h]; l_s[warp][h] = l[h]; }
#pragma unroll
        for (int i = 0; i < VPL; i++) { const int d = lane + 32 * i; if (d < HS) o_s[warp][h][d] = o[h][i]; }
    }
    __syncthreads();
    const int stride = HS + 2;
    for (int e = tid; e < nh * HS; e += ATT_WARPS * 32) {
        const int h = e / HS, d = e - h * HS;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < ATT_WARPS; w++) M = fmaxf(M, m_s[w][h]);
        float L = 0.0f, O = 0.0f;
        if (M > -INFINITY) {
#pragma unroll
            for (int w = 0; w < ATT_WARPS; w++) {
                const float sc = expf(m_s[w][h] - M);
                L = fmaf(l_s[w][h], sc, L);
                O = fmaf(o_s[w][h][d], sc, O);
            }
        }
        float* dst = p.part + ((size_t)(h0 + h) * p.nsplit + split) * stride;
        dst[d] = O;
        if (d == 0) { dst[HS] = M; dst[HS + 1] = L; }
    }
    // ticket: the last CTA of this (kv head, chunk) merges all splits
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&p.tickets[blockIdx.x], 1u);
        is_last_cta = (prev == (unsigned)p.nsplit - 1u);
        if (is_last_cta) p.tickets[blockIdx.x] = 0u;   // re-arm for the next launch
    }
    __syncthreads();
    if (!is_last_cta) return;
    __threadfence();
    for (int e = tid; e < nh * HS; e += ATT_WARPS * 32) {
        const int h = e / HS, d = e - h * HS;
        const float* src = p.part + (size_t)(h0 + h) * p.nsplit * stride;
        float M = -INFINITY;
        for (int s = 0; s < p.nsplit; s++) M = fmaxf(M, __ldcg(src + (size_t)s * stride + HS));
        float L = 0.0f, O = 0.0f;
        for (int s = 0; s < p.nsplit; s++) {
            const float ms = __ldcg(src + (size_t)s * stride + HS);
            if (ms > -INFINITY) {
                const float sc = expf(ms - M);
                L = fmaf(__ldcg(src + (size_t)s * stride + HS + 1), sc, L);
                O = fmaf(__ldcg(src + (size_t)s * stride + d), sc, O);
            }
        }
        p.out[(size_t)(h0 + h) * HS + d] = __fdiv_rn(O, L);
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
