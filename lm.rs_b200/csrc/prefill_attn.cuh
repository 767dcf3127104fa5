// prefill_attn.cuh -- attention of a BATCH of token rows (fill_kv_cache -> forward_layer with sl = N,
// src/transformer.rs:501-544) as two throughput kernels.
//
// The decode kernels (attention.cuh) spend a whole CTA or cluster on one token's dependent chains; with N tokens there
// are N x n_heads independent softmax rows and N x n_heads x head_size independent a*v chains, so here every chain is
// ONE THREAD and the K / V tiles of a KV head are staged once in shared memory for all the tokens and GQA heads of a
// CTA.  The arithmetic of a chain is unchanged -- serial dot over d, /sqrt(hs), max, exp(x - max) with the glibc expf
// restatement, serial sum over t, divide, serial a*v accumulation over t, mul and add unfused -- so the results are
// bit-identical to the reference's loops (and to the decode kernels).
//
//   prefill_scores_kernel  thread = (token, query head): scores of all its positions (4 positions in flight per thread),
//                          then softmax over its own row.  Rows live in an HBM/L2 scratch laid out [kv head][t][token][h]
//                          so that the threads of a CTA read and write consecutive words.
//   prefill_av_kernel      thread = (token, query head, 8 output dims): 8 accumulation chains over t.
#pragma once
#include "attention.cuh"

namespace lmrs {

struct PrefillAttnParams {
    const float* q;        // [n][att_dim] RoPE already applied (rope_rows_kernel)
    const float* kcache;   // layer base [seq_len][kv_dim]; rows pos .. pos+n-1 written (and rotated) by this batch
    const float* vcache;
    float* probs;          // scratch [kv heads][t_cap][n][kv_mul]
    float* out;            // [n][att_dim]
    int n, pos, t_cap;     // batch rows, first position, positions covered by the scratch (>= pos + n)
    int att_dim, kv_dim, kv_mul, gemma;
    uint32_t mask_base;    // Gemma window quirk: the reference tests `pos - t` with the BATCH start (src/transformer.rs:525)
    float sqrt_hs;
    float neg_zero;        // -0.0f at run time: keeps packed products separately rounded (gemm.cuh f2_mul_sep)
};

constexpr int PFA_THREADS = 128;   // scores kernel: (token, head) pairs per CTA
constexpr int PFA_TT = 32;         // positions per staged K / V tile

template <int HS> constexpr size_t prefill_scores_smem() { return ((size_t)HS * PFA_THREADS + (size_t)HS * (PFA_TT + 4) + 64) * 4; }

template <int HS>
__global__ void __launch_bounds__(PFA_THREADS) prefill_scores_kernel(const PrefillAttnParams p) {
    extern __shared__ __align__(16) float pfs[];
    float* q_t = pfs;                                   // [HS][128]   q transposed: consecutive threads, consecutive words
    float* k_t = q_t + HS * PFA_THREADS;                // [HS][TT+4]  K tile transposed: 4 consecutive positions = one float4
    uint64_t* exp_tab = reinterpret_cast<uint64_t*>(k_t + HS * (PFA_TT + 4));
    constexpr int KS = PFA_TT + 4;
    const int tid = threadIdx.x, g = blockIdx.y;
    const int TB = PFA_THREADS / p.kv_mul;              // tokens per CTA
    const int tok_l = tid / p.kv_mul, h_l = tid - tok_l * p.kv_mul;
    const int tok = blockIdx.x * TB + tok_l;
    const bool valid = tok_l < TB && tok < p.n;
    const int my_T = valid ? p.pos + tok + 1 : 0;       // positions this row attends to
    const int cta_T = min(p.pos + p.n, p.pos + (int)(blockIdx.x + 1) * TB);
    if (tid < 32) exp_tab[tid] = kExp2fTab[tid];
    {   // q row of this thread -> transposed shared copy
        const float4* qr = reinterpret_cast<const float4*>(p.q + (size_t)(valid ? tok : 0) * p.att_dim + (size_t)(g * p.kv_mul + (valid ? h_l : 0)) * HS);
#pragma unroll 4
        for (int c = 0; c < HS / 4; c++) {
            const float4 v = valid ? qr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            q_t[(4 * c + 0) * PFA_THREADS + tid] = v.x; q_t[(4 * c + 1) * PFA_THREADS + tid] = v.y;
            q_t[(4 * c + 2) * PFA_THREADS + tid] = v.z; q_t[(4 * c + 3) * PFA_THREADS + tid] = v.w;
        }
    }
    // scratch row of this thread: element t at probs[((g * t_cap + t) * n + tok) * kv_mul + h_l]
    float* srow = p.probs + ((size_t)g * p.t_cap * p.n + (size_t)(valid ? tok : 0)) * p.kv_mul + h_l;
    const size_t sstep = (size_t)p.n * p.kv_mul;
    float mx = -INFINITY;
    for (int t0 = 0; t0 < cta_T; t0 += PFA_TT) {
        __syncthreads();                                // everybody is done with the previous tile (and q_t is written)
        for (int e = tid; e < PFA_TT * (HS / 4); e += PFA_THREADS) {
            const int j = e / (HS / 4), c = e - j * (HS / 4), t = t0 + j;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < cta_T) v = *reinterpret_cast<const float4*>(p.kcache + (size_t)t * p.kv_dim + (size_t)g * HS + 4 * c);
            k_t[(4 * c + 0) * KS + j] = v.x; k_t[(4 * c + 1) * KS + j] = v.y; k_t[(4 * c + 2) * KS + j] = v.z; k_t[(4 * c + 3) * KS + j] = v.w;
        }
        __syncthreads();
        if (t0 >= my_T) continue;                       // (whole warps drop out early only at the causal edge)
#pragma unroll 1
        for (int j4 = 0; j4 < PFA_TT / 4; j4++) {
            const int t = t0 + 4 * j4;
            if (t >= my_T) break;
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;   // four positions in flight: independent ascending-d chains
#pragma unroll 8
            for (int d = 0; d < HS; d++) {
                const float qd = q_t[d * PFA_THREADS + tid];
                const float4 kk = *reinterpret_cast<const float4*>(k_t + d * KS + 4 * j4);
                a0 = __fadd_rn(a0, __fmul_rn(qd, kk.x)); a1 = __fadd_rn(a1, __fmul_rn(qd, kk.y));
                a2 = __fadd_rn(a2, __fmul_rn(qd, kk.z)); a3 = __fadd_rn(a3, __fmul_rn(qd, kk.w));
            }
            const float sv[4] = {a0, a1, a2, a3};
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (t + u < my_T) {
                    float score = __fdiv_rn(sv[u], p.sqrt_hs);
                    if (p.gemma) {   // soft-cap 50*tanh(s/50) in f64, window mask on every layer (:518-526)
                        score = __fdiv_rn(score, 50.0f);
                        score = (float)tanh((double)score);
                        score = __fmul_rn(score, 50.0f);
                        score = __fadd_rn(score, (p.mask_base - (uint32_t)(t + u) <= 4096u) ? 0.0f : -2.3819763e38f);
                    }
                    srow[(size_t)(t + u) * sstep] = score;
                    mx = fmaxf(mx, score);
                }
            }
        }
    }
    // softmax over my own row (src/functional.rs:122-140): exp(x - max), SERIAL sum in ascending t, divide
    float sum = 0.0f;
    for (int t = 0; t < my_T; t++) {
        const float e = expf_glibc_t(__fsub_rn(srow[(size_t)t * sstep], mx), exp_tab);
        srow[(size_t)t * sstep] = e;
        sum = __fadd_rn(sum, e);
    }
    for (int t = 0; t < my_T; t++) srow[(size_t)t * sstep] = __fdiv_rn(srow[(size_t)t * sstep], sum);
}

// tokens per CTA of the a*v kernel: pairs x (HS / 8) threads, about 256
inline int prefill_av_tokens(int hs, int kv_mul) { int t = 256 / (kv_mul * (hs / 8)); return t < 1 ? 1 : t; }
template <int HS> inline size_t prefill_av_smem(int pairs) { return ((size_t)2 * PFA_TT * HS + (size_t)2 * PFA_TT * pairs) * 4; }

template <int HS>
__global__ void __launch_bounds__(512) prefill_av_kernel(const PrefillAttnParams p, const int TBV) {
    extern __shared__ __align__(16) float pfv[];
    constexpr int DG = HS / 8;                          // threads per (token, head) pair, 8 output dims each
    const int pairs = TBV * p.kv_mul;
    float* v_s = pfv;                                   // [2][TT][HS]
    float* p_s = v_s + 2 * PFA_TT * HS;                 // [2][TT][pairs]
    const int tid = threadIdx.x, g = blockIdx.y;
    const int pair_l = tid / DG, dg = tid - pair_l * DG;
    const int tok_l = pair_l / p.kv_mul, h_l = pair_l - tok_l * p.kv_mul;
    const int tok0 = blockIdx.x * TBV, tok = tok0 + tok_l;
    const bool valid = pair_l < pairs && tok < p.n;
    const int my_T = valid ? p.pos + tok + 1 : 0;
    const int cta_T = min(p.pos + p.n, p.pos + tok0 + TBV);
    const int nthr = blockDim.x;
    const float* pbase = p.probs + ((size_t)g * p.t_cap * p.n + (size_t)tok0) * p.kv_mul;   // + t * n * kv_mul + pair
    const size_t sstep = (size_t)p.n * p.kv_mul;
    const int npair_ok = min(pairs, (p.n - tok0) * p.kv_mul);
    auto stage = [&](int tl) {
        const int t0 = tl * PFA_TT, b = tl & 1;
        float* vb = v_s + b * PFA_TT * HS;
        for (int e = tid; e < PFA_TT * (HS / 4); e += nthr) {
            const int j = e / (HS / 4), c = e - j * (HS / 4);
            if (t0 + j < cta_T) cp_async16(vb + j * HS + 4 * c, p.vcache + (size_t)(t0 + j) * p.kv_dim + (size_t)g * HS + 4 * c);
        }
        float* pb = p_s + b * PFA_TT * pairs;
        for (int e = tid; e < PFA_TT * pairs; e += nthr) {
            const int j = e / pairs, q = e - j * pairs;
            pb[e] = (t0 + j < cta_T && q < npair_ok) ? __ldcg(pbase + (size_t)(t0 + j) * sstep + q) : 0.0f;
        }
        cp_async_commit();
    };
    float acc[8];
#pragma unroll
    for (int u = 0; u < 8; u++) acc[u] = 0.0f;
    const int ntiles = (cta_T + PFA_TT - 1) / PFA_TT;
    stage(0);
    for (int tl = 0; tl < ntiles; tl++) {
        if (tl + 1 < ntiles) stage(tl + 1); else cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const float* vb = v_s + (tl & 1) * PFA_TT * HS + dg * 8;
        const float* pb = p_s + (tl & 1) * PFA_TT * pairs + pair_l;
        const int rows = min(PFA_TT, my_T - tl * PFA_TT);
        for (int j = 0; j < rows; j++) {                // one multiply and one dependent add per position, ascending t (:533-542)
            const float a = pb[j * pairs];
            const float4 v0 = *reinterpret_cast<const float4*>(vb + j * HS), v1 = *reinterpret_cast<const float4*>(vb + j * HS + 4);
            acc[0] = __fadd_rn(acc[0], __fmul_rn(a, v0.x)); acc[1] = __fadd_rn(acc[1], __fmul_rn(a, v0.y));
            acc[2] = __fadd_rn(acc[2], __fmul_rn(a, v0.z)); acc[3] = __fadd_rn(acc[3], __fmul_rn(a, v0.w));
            acc[4] = __fadd_rn(acc[4], __fmul_rn(a, v1.x)); acc[5] = __fadd_rn(acc[5], __fmul_rn(a, v1.y));
            acc[6] = __fadd_rn(acc[6], __fmul_rn(a, v1.z)); acc[7] = __fadd_rn(acc[7], __fmul_rn(a, v1.w));
        }
        __syncthreads();                                // the buffer staged two tiles from now is this one
    }
    cp_async_wait<0>();
    if (valid) {
        float4* o = reinterpret_cast<float4*>(p.out + (size_t)tok * p.att_dim + (size_t)(g * p.kv_mul + h_l) * HS + dg * 8);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}


// ---- fused form: scores + softmax + a*v of a tile of token rows in ONE kernel, score rows in shared memory ----------------
//
// CTA = (KV head, tile of `tb` consecutive tokens): RW = tb * kv_mul <= 32 (token, query head) rows.  256 threads =
// 32 row slots x 8 lanes.  The scores never leave shared memory, the softmax rows are warp-local (the 8 lanes of a row
// sit in one warp: no CTA barrier between its passes), and every phase keeps several INDEPENDENT exact-order chains per
// thread so that the kernel is bound by instruction issue, not by the 4-cycle dependent-add latency:
//   scores   lane j of a row owns positions 4j..4j+3 of every 32-position K tile: four ascending-d chains, two packed
//            f32x2 multiply/add pairs per step (K tile transposed to [d][position], q stored as (q, q) pairs);
//   softmax  max over the row, exp(x - max) (glibc expf restated), SERIAL sum in ascending t by one lane per row (32 rows
//            run their chains side by side), divide;
//   a*v      lane j of a row owns output dims [4j, 4j+4) of each HS/8... (HS/32 float4 per lane): one product and one
//            dependent add per position in ascending t, packed two dims per instruction.
// Arithmetic per chain is exactly the reference's (src/transformer.rs:507-542, src/functional.rs:122-140): results are
// bit-identical to the decode kernels and to the two-kernel form above.  Heavy (late) token tiles are scheduled first.
constexpr int PFF_THREADS = 256, PFF_ROWS = 32, PFF_TT = 32, PFF_KS = PFF_TT + 4;
inline size_t prefill_fused_smem(int hs, int rw, int scs) {
    return ((size_t)PFF_ROWS * (hs * 2 + 4) + (size_t)2 * hs * PFF_KS + (size_t)rw * scs + 64 + 8) * 4;
}
constexpr size_t PFF_SMEM_MAX = 216 * 1024;
// score row stride for contexts up to t_max positions: +4 floats de-phases the rows' banks
inline int prefill_fused_scs(int t_max) { return ((t_max + 3) & ~3) + 4; }
// tokens per CTA for contexts up to t_max positions (0: the score rows of even one token do not fit)
inline int prefill_fused_tb(int hs, int kv_mul, int t_max) {
    if (kv_mul > PFF_ROWS) return 0;
    int tb = PFF_ROWS / kv_mul;
    while (tb > 1 && prefill_fused_smem(hs, tb * kv_mul, prefill_fused_scs(t_max)) > PFF_SMEM_MAX) tb--;
    return prefill_fused_smem(hs, tb * kv_mul, prefill_fused_scs(t_max)) <= PFF_SMEM_MAX ? tb : 0;
}

// a separately rounded packed product (an add consumes it): fma(a, b, -0.0) with a run-time -0.0, see gemm.cuh f2_mul_sep
LMRS_DEVINL uint64_t pf2_mul(uint64_t a, uint64_t b, uint64_t nz2) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz2)); return r; }
LMRS_DEVINL uint64_t pf2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
LMRS_DEVINL uint64_t pf2_dup(float a) { uint64_t r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(a)); return r; }
LMRS_DEVINL void pf2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

template <int HS>
__global__ void __launch_bounds__(PFF_THREADS) prefill_attn_fused_kernel(const PrefillAttnParams p, const int tb, const int SCS) {
    static_assert(HS % 32 == 0, "eight lanes x whole float4 per V row");
    constexpr int KS = PFF_KS, TT = PFF_TT;
    constexpr int KPT = TT * (HS / 4) / PFF_THREADS;       // float4 of a K tile staged per thread
    constexpr int VC = HS / 32;                            // float4 of a V row owned by one lane
    extern __shared__ __align__(16) float pff[];
    constexpr int Q2S = HS * 2 + 4;                        // row stride of q2: +16 bytes puts the four rows of a warp on distinct banks
    float* q2 = pff;                                       // [32][HS] (q, q) pairs
    float* tile = q2 + PFF_ROWS * Q2S;                     // [2][HS][KS] transposed K tiles / [2][TT][HS] V tiles (HS*KS >= TT*HS)
    float* sc_s = tile + 2 * HS * KS;                      // [RW][SCS]
    uint64_t* exp_tab = reinterpret_cast<uint64_t*>(sc_s + (size_t)tb * p.kv_mul * SCS);
    const int tid = threadIdx.x, g = blockIdx.y;
    const int ntile = (p.n + tb - 1) / tb;
    const int tix = ntile - 1 - (int)blockIdx.x;           // heavy tiles first
    const int tok0 = tix * tb;
    const int RW = tb * p.kv_mul;
    const int r = tid >> 3, j8 = tid & 7;
    const int tok_l = r / p.kv_mul, h_l = r - tok_l * p.kv_mul;
    const int tok = tok0 + tok_l;
    const bool valid = r < RW && tok < p.n;
    const int my_T = valid ? p.pos + tok + 1 : 0;
    const int cta_T = min(p.pos + p.n, p.pos + tok0 + tb);
    const uint64_t nz2 = pf2_dup(p.neg_zero);
    if (tid < 32) exp_tab[tid] = kExp2fTab[tid];
    {   // q rows as (q, q) pairs
        const float* qr = p.q + (size_t)(valid ? tok : 0) * p.att_dim + (size_t)(g * p.kv_mul + (valid ? h_l : 0)) * HS;
        for (int c = j8; c < HS / 4; c += 8) {
            const float4 v = valid ? *reinterpret_cast<const float4*>(qr + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4* d = reinterpret_cast<float4*>(q2 + (size_t)r * Q2S + 8 * c);
            d[0] = make_float4(v.x, v.x, v.y, v.y); d[1] = make_float4(v.z, v.z, v.w, v.w);
        }
    }
    // ---- scores ------------------------------------------------------------------------------------------------------
    const int ntk = (cta_T + TT - 1) / TT;
    float4 kreg[KPT];
    auto load_k = [&](int tl) {   // element e: position j = e % TT (consecutive lanes: conflict-free transposing stores), chunk c = e / TT
#pragma unroll
        for (int u = 0; u < KPT; u++) {
            const int e = tid + u * PFF_THREADS, j = e % TT, c = e / TT, t = tl * TT + j;
            kreg[u] = t < cta_T ? *reinterpret_cast<const float4*>(p.kcache + (size_t)t * p.kv_dim + (size_t)g * HS + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_k = [&](int b) {
        float* kt = tile + b * HS * KS;
#pragma unroll
        for (int u = 0; u < KPT; u++) {
            const int e = tid + u * PFF_THREADS, j = e % TT, c = e / TT;
            kt[(4 * c + 0) * KS + j] = kreg[u].x; kt[(4 * c + 1) * KS + j] = kreg[u].y;
            kt[(4 * c + 2) * KS + j] = kreg[u].z; kt[(4 * c + 3) * KS + j] = kreg[u].w;
        }
    };
    float mx = -INFINITY;
    float* srow = sc_s + (size_t)(r < RW ? r : 0) * SCS;
    load_k(0);
    for (int tl = 0; tl < ntk; tl++) {
        store_k(tl & 1);
        __syncthreads();                                    // tile tl visible; everybody finished tile tl-1 (the other buffer)
        if (tl + 1 < ntk) load_k(tl + 1);                   // in flight under this tile's arithmetic
        const int t0 = tl * TT + 4 * j8;
        if (t0 < my_T) {
            const float* kt = tile + (tl & 1) * HS * KS + 4 * j8;
            const float* qq = q2 + (size_t)r * Q2S;
            uint64_t a01 = 0ull, a23 = 0ull;               // chains of positions (t0, t0+1) and (t0+2, t0+3), ascending d
#pragma unroll 8
            for (int d = 0; d < HS; d += 2) {
                const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(qq + 2 * d);          // (q_d, q_d), (q_d+1, q_d+1)
                const ulonglong2 k0 = *reinterpret_cast<const ulonglong2*>(kt + d * KS);
                const ulonglong2 k1 = *reinterpret_cast<const ulonglong2*>(kt + (d + 1) * KS);
                a01 = pf2_add(a01, pf2_mul(q.x, k0.x, nz2)); a23 = pf2_add(a23, pf2_mul(q.x, k0.y, nz2));
                a01 = pf2_add(a01, pf2_mul(q.y, k1.x, nz2)); a23 = pf2_add(a23, pf2_mul(q.y, k1.y, nz2));
            }
            float sv[4];
            pf2_unpack(a01, sv[0], sv[1]); pf2_unpack(a23, sv[2], sv[3]);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (t0 + u < my_T) {
                    float score = __fdiv_rn(sv[u], p.sqrt_hs);
                    if (p.gemma) {   // soft-cap 50*tanh(s/50) in f64, window mask on every layer (:518-526)
                        score = __fdiv_rn(score, 50.0f);
                        score = (float)tanh((double)score);
                        score = __fmul_rn(score, 50.0f);
                        score = __fadd_rn(score, (p.mask_base - (uint32_t)(t0 + u) <= 4096u) ? 0.0f : -2.3819763e38f);
                    }
                    srow[t0 + u] = score;
                    mx = fmaxf(mx, score);
                }
            }
        }
    }
    // ---- softmax, row-local (the 8 lanes of a row are neighbours in one warp) -------------------------------------------
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncwarp();
    for (int t = j8; t < my_T; t += 8) srow[t] = expf_glibc_t(__fsub_rn(srow[t], mx), exp_tab);
    __syncwarp();
    float sum = 0.0f;
    if (j8 == 0 && my_T > 0) sum = serial_sum_f32(srow, my_T);   // ascending t, one dependent add per element (src/functional.rs:131-134)
    sum = __shfl_sync(0xffffffffu, sum, (tid & 31) & ~7);
    for (int t = j8; t < my_T; t += 8) srow[t] = __fdiv_rn(srow[t], sum);
    __syncthreads();                                        // all K-tile reads done: the tile buffers now carry V
    // ---- a*v --------------------------------------------------------------------------------------------------------------
    auto stage_v = [&](int tl) {
        float* vb = tile + (tl & 1) * TT * HS;
        for (int e = tid; e < TT * (HS / 4); e += PFF_THREADS) {
            const int j = e / (HS / 4), c = e - j * (HS / 4), t = tl * TT + j;
            if (t < cta_T) cp_async16(vb + j * HS + 4 * c, p.vcache + (size_t)t * p.kv_dim + (size_t)g * HS + 4 * c);
        }
        cp_async_commit();
    };
    uint64_t acc[VC][2];
#pragma unroll
    for (int c = 0; c < VC; c++) acc[c][0] = acc[c][1] = 0ull;
    stage_v(0);
    for (int tl = 0; tl < ntk; tl++) {
        if (tl + 1 < ntk) stage_v(tl + 1); else cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const float* vb = tile + (tl & 1) * TT * HS + 4 * j8;   // lane j owns float4 chunks j, j+8, ... of every row
        const int rows = min(TT, my_T - tl * TT);
        const float* pr = srow + tl * TT;
        // one multiply and one dependent add per position, ascending t (:533-542); probabilities fetched four at a time
        auto step = [&](int j, float a) {
            const uint64_t a2 = pf2_dup(a);
#pragma unroll
            for (int c = 0; c < VC; c++) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(vb + j * HS + 32 * c);
                acc[c][0] = pf2_add(acc[c][0], pf2_mul(a2, v.x, nz2));
                acc[c][1] = pf2_add(acc[c][1], pf2_mul(a2, v.y, nz2));
            }
        };
        int j = 0;
        for (; j + 4 <= rows; j += 4) {
            const float4 a4 = *reinterpret_cast<const float4*>(pr + j);
            step(j, a4.x); step(j + 1, a4.y); step(j + 2, a4.z); step(j + 3, a4.w);
        }
        for (; j < rows; j++) step(j, pr[j]);
        __syncthreads();                                    // the buffer staged two tiles from now is this one
    }
    cp_async_wait<0>();
    if (valid) {
        float* o = p.out + (size_t)tok * p.att_dim + (size_t)(g * p.kv_mul + h_l) * HS + 4 * j8;
#pragma unroll
        for (int c = 0; c < VC; c++) {
            float4 w;
            pf2_unpack(acc[c][0], w.x, w.y); pf2_unpack(acc[c][1], w.z, w.w);
            *reinterpret_cast<float4*>(o + 32 * c) = w;
        }
    }
}

}  // namespace lmrs
