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
constexpr int ATT_QH = 4;        // query heads per CTA (all sharing one KV head)
constexpr int ATT_SC_CAP = 2048; // positions whose scores are kept in shared memory (longer contexts: HBM scratch)

struct AttnParams {
    const float* q;        // [att_dim] un-rotated
    const float* k_new;    // [kv_dim]  un-rotated K of this step
    float* kcache;         // layer base [seq_len][kv_dim]
    const float* vcache;   // layer base (row `pos` already written by the QKV GEMV)
    const float* rope_cos; // [seq_len][hs/2]
    const float* rope_sin;
    float* out;            // [att_dim]
    float* scores;         // scratch [n_heads][seq_len] used when pos+1 > ATT_SC_CAP
    int kv_dim, kv_mul, chunks, gemma, seq_len;
    float sqrt_hs;         // sqrtf(head_size): scores are DIVIDED by it (src/transformer.rs:516)
    const StepParams* step;
};

template <int HS> __host__ __device__ constexpr int att_tile_rows() { return HS > 128 ? 32 : 64; }
template <int HS> __host__ __device__ constexpr size_t attn_smem_bytes() {
    return (size_t)(ATT_QH * HS + HS + 2 * att_tile_rows<HS>() * HS + ATT_QH * ATT_SC_CAP + 64) * 4;
}

// Bit-exact restatement of src/transformer.rs:501-544 for one token: every f32 operation happens in the
// reference's order (serial dot over d, serial softmax sum over t, serial a*v accumulation over t, separate
// mul and add, exp = glibc expf); only independent chains run in parallel.  One CTA per KV head (x chunks of 4
// query heads).  K and V tiles are register-prefetched (next tile's global loads are in flight while the
// current tile's dependent-add chains run from shared memory); the K tile is stored column-rotated so that
// the per-position dot products read it conflict-free without breaking the ascending-d summation order.
// Latency floor: the two T-long dependent add chains (softmax sum, a*v) -- the price of exact parity, see
// exact_math.cuh.
template <int HS>
LMRS_DEVINL void attn_decode_body(const AttnParams& p, float* att_smem, const int kvh, const int h0, const int nh,
                                  const bool write_k) {
    constexpr int TILE = att_tile_rows<HS>();
    constexpr int PER = TILE * HS / ATT_THREADS;      // tile elements per thread (16 or 32)
    float* q_s = att_smem;                             // [ATT_QH][HS]
    float* k_s = q_s + ATT_QH * HS;                    // [HS] rotated new K row
    float* tile = k_s + HS;                            // [2][TILE][HS]
    float* sc_s = tile + 2 * TILE * HS;                // [ATT_QH][ATT_SC_CAP]
    float* red = sc_s + ATT_QH * ATT_SC_CAP;           // [64]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pos = (int)p.step->pos;
    const uint32_t mask_base = p.step->mask_base;
    const int T = pos + 1;
    const bool in_smem = T <= ATT_SC_CAP;
    float* sc_base = in_smem ? sc_s : p.scores + (size_t)h0 * p.seq_len;
    const int sc_stride = in_smem ? ATT_SC_CAP : p.seq_len;

    // RoPE on q and on the new k row (rotate-half pairs j, j+HS/2), src/transformer.rs:480-492
    const float* cs = p.rope_cos + (size_t)pos * (HS / 2);
    const float* sn = p.rope_sin + (size_t)pos * (HS / 2);
    for (int i = tid; i < nh * (HS / 2); i += ATT_THREADS) {
        const int h = i / (HS / 2), j = i - h * (HS / 2);
        const float fcr = cs[j], fci = sn[j];
        const float v0 = __ldcg(p.q + (size_t)(h0 + h) * HS + j), v1 = __ldcg(p.q + (size_t)(h0 + h) * HS + j + HS / 2);
        q_s[h * HS + j] = __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci));
        q_s[h * HS + j + HS / 2] = __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr));
    }
    for (int j = tid; j < HS / 2; j += ATT_THREADS) {
        const float fcr = cs[j], fci = sn[j];
        const float v0 = __ldcg(p.k_new + (size_t)kvh * HS + j), v1 = __ldcg(p.k_new + (size_t)kvh * HS + j + HS / 2);
        const float r0 = __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci));
        const float r1 = __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr));
        k_s[j] = r0; k_s[j + HS / 2] = r1;
        if (write_k) {   // exactly one CTA per KV head publishes the rotated row into the cache
            p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + j] = r0;
            p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + j + HS / 2] = r1;
        }
    }
    __syncthreads();

    const int ntiles = (T + TILE - 1) / TILE;
    float pre[PER];
    // ---- scores: s[h][t] = (sum_d q[h][d]*k[t][d]) / sqrt(hs)   (:507-528) --------------------------------
    auto load_k = [&](int tl) {      // tile element e = tid + i*256 -> row e/HS, col e%HS (coalesced rows)
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int e = tid + i * ATT_THREADS, r = e / HS, d = e - r * HS, t = tl * TILE + r;
            pre[i] = (t < T && t != pos) ? __ldcg(p.kcache + (size_t)t * p.kv_dim + (size_t)kvh * HS + d) : (t == pos ? k_s[d] : 0.0f);
        }
    };
    load_k(0);
    for (int tl = 0; tl < ntiles; tl++) {
        float* tb = tile + (tl & 1) * TILE * HS;
#pragma unroll
        for (int i = 0; i < PER; i++) {   // column-rotated store: (r, d) -> r*HS + (d + r) % HS
            const int e = tid + i * ATT_THREADS, r = e / HS, d = e - r * HS;
            int c = d + r; if (c >= HS) c -= HS;
            tb[r * HS + c] = pre[i];
        }
        __syncthreads();
        if (tl + 1 < ntiles) load_k(tl + 1);          // next tile's loads fly during this tile's chains
        const int rows = min(TILE, T - tl * TILE);
        for (int idx = tid; idx < rows * nh; idx += ATT_THREADS) {
            const int h = idx / rows, r = idx - h * rows, t = tl * TILE + r;
            const float* qh = q_s + h * HS;
            const float* kr = tb + r * HS;
            float score = 0.0f;
            int c = r;                                  // column of d = 0
#pragma unroll 8
            for (int d = 0; d < HS; d++) {
                score = __fadd_rn(score, __fmul_rn(qh[d], kr[c]));
                c = (c + 1 == HS) ? 0 : c + 1;
            }
            score = __fdiv_rn(score, p.sqrt_hs);
            if (p.gemma) {   // soft-cap 50*tanh(s/50) in f64, window mask on every layer (:518-526)
                score = __fdiv_rn(score, 50.0f);
                score = (float)tanh((double)score);
                score = __fmul_rn(score, 50.0f);
                score = __fadd_rn(score, (mask_base - (uint32_t)t <= 4096u) ? 0.0f : -2.3819763e38f);
            }
            sc_base[(size_t)h * sc_stride + t] = score;
        }
        // no second barrier needed: the next iteration writes the OTHER buffer, and its __syncthreads orders
        // this tile's reads before the buffer is overwritten two iterations later
    }
    __syncthreads();

    // prefetch the first V tile while the softmax runs
    auto load_v = [&](int tl) {
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int e = tid + i * ATT_THREADS, r = e / HS, d = e - r * HS, t = tl * TILE + r;
            pre[i] = t < T ? __ldcg(p.vcache + (size_t)t * p.kv_dim + (size_t)kvh * HS + d) : 0.0f;
        }
    };
    load_v(0);

    // ---- softmax (src/functional.rs:122-140): max, exp(x-max), serial sum, divide ---------------------------
    for (int h = 0; h < nh; h++) {
        float* sc = sc_base + (size_t)h * sc_stride;
        float mx = sc[0];
        for (int t = tid; t < T; t += ATT_THREADS) mx = fmaxf(mx, sc[t]);
        mx = warp_max(mx);
        if (lane == 0) red[h * 8 + warp] = mx;
    }
    __syncthreads();
    for (int h = 0; h < nh; h++) {
        float* sc = sc_base + (size_t)h * sc_stride;
        float mx = red[h * 8];
#pragma unroll
        for (int w = 1; w < ATT_THREADS / 32; w++) mx = fmaxf(mx, red[h * 8 + w]);
        for (int t = tid; t < T; t += ATT_THREADS) sc[t] = expf_glibc(__fsub_rn(sc[t], mx));
    }
    __syncthreads();
    if (lane == 0 && warp < nh) {   // the reference's `sum += x[i]` chain: one thread per head, in different warps
        const float* sc = sc_base + (size_t)warp * sc_stride;
        float sum = 0.0f;
        int t = 0;
        for (; t + 8 <= T; t += 8) {
            const float4 a = *reinterpret_cast<const float4*>(sc + t), b = *reinterpret_cast<const float4*>(sc + t + 4);
            sum = __fadd_rn(sum, a.x); sum = __fadd_rn(sum, a.y); sum = __fadd_rn(sum, a.z); sum = __fadd_rn(sum, a.w);
            sum = __fadd_rn(sum, b.x); sum = __fadd_rn(sum, b.y); sum = __fadd_rn(sum, b.z); sum = __fadd_rn(sum, b.w);
        }
        for (; t < T; t++) sum = __fadd_rn(sum, sc[t]);
        red[32 + warp] = sum;
    }
    __syncthreads();
    for (int h = 0; h < nh; h++) {
        float* sc = sc_base + (size_t)h * sc_stride;
        const float sum = red[32 + h];
        for (int t = tid; t < T; t += ATT_THREADS) sc[t] = __fdiv_rn(sc[t], sum);
    }
    // (the barrier inside the first V-tile iteration orders these writes before the chains read them)

    // ---- out[h][d] = sum_t a[h][t] * v[t][d], serial over t (:533-542) --------------------------------------
    constexpr int MAXCH = (ATT_QH * HS + ATT_THREADS - 1) / ATT_THREADS;   // chains per thread
    float acc[MAXCH];
#pragma unroll
    for (int k = 0; k < MAXCH; k++) acc[k] = 0.0f;
    for (int tl = 0; tl < ntiles; tl++) {
        float* tb = tile + (tl & 1) * TILE * HS;
#pragma unroll
        for (int i = 0; i < PER; i++) tb[tid + i * ATT_THREADS] = pre[i];
        __syncthreads();
        if (tl + 1 < ntiles) load_v(tl + 1);
        const int rows = min(TILE, T - tl * TILE);
#pragma unroll
        for (int k = 0; k < MAXCH; k++) {
            const int idx = tid + k * ATT_THREADS;
            if (idx < nh * HS) {
                const int h = idx / HS, d = idx - h * HS;
                const float* a = sc_base + (size_t)h * sc_stride + tl * TILE;
                float x = acc[k];
#pragma unroll 8
                for (int r = 0; r < rows; r++) x = __fadd_rn(x, __fmul_rn(a[r], tb[r * HS + d]));
                acc[k] = x;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < MAXCH; k++) {
        const int idx = tid + k * ATT_THREADS;
        if (idx < nh * HS) p.out[(size_t)h0 * HS + idx] = acc[k];
    }
}

template <int HS>
__global__ void __launch_bounds__(ATT_THREADS) attn_decode_kernel(const AttnParams p) {
    extern __shared__ __align__(16) float att_smem_dyn[];
    const int kvh = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks;
    const int h0 = kvh * p.kv_mul + chunk * ATT_QH;                  // first query head of this CTA
    const int nh = min(ATT_QH, p.kv_mul - chunk * ATT_QH);           // query heads served here
    pdl_launch_dependents();
    pdl_wait();
    attn_decode_body<HS>(p, att_smem_dyn, kvh, h0, nh, chunk == 0);
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
