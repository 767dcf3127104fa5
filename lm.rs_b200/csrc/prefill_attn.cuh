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
// CTA = (KV head, tile of `tb` consecutive tokens): RW = tb * kv_mul <= 32 (token, query head) rows, handled as 16 ROW PAIRS:
// every f32x2 instruction carries the SAME step of the two rows' chains in its two halves (q, the score rows and the
// accumulators are stored pair-interleaved) and the K / V operand of a step is a scalar the packed instruction broadcasts.
// The first build of this kernel (one row per 8-lane group, q and K re-read from shared memory for every 4 multiply-adds) was
// bound by the shared-memory pipe, not by arithmetic (ncu: short-scoreboard stalls 5x the issue slots, a 128-bit LDS costs a
// warp four pipe cycles whatever it broadcasts; profiles/r2_ncu_prefill_attn_raw.csv), so the work is register-tiled:
//   scores   NQD warps = row quads x 32 lanes; thread (quad, lane j) owns the 4 x 4 tile of rows {two pairs} x positions
//            4j..4j+3 of every 128-position K tile (K transposed to [d][position]): per d ONE 16-byte load of the two q pairs
//            and ONE of four K values feed 8 packed multiply/add pairs = 16 ascending-d chains;
//   softmax  warp-local (a quad's lanes are one warp; lanes 0..15 serve its first pair, 16..31 the second): max, exp(x - max)
//            (glibc expf restated), SERIAL sums of both rows in ascending t by one lane (two independent chains), divide;
//            positions beyond a row's own length are set to +0.0 (adding +-0 never changes an accumulator that started from
//            +0.0, so the longer row of a pair may drive the loops);
//   a*v      thread (pair, lane j of 16) owns float4 chunks j, j+16, .. of the head: one packed product and one dependent packed
//            add per position and dim, ascending t, both rows of the pair per instruction.
// Arithmetic per chain is exactly the reference's (src/transformer.rs:507-542, src/functional.rs:122-140): results are
// bit-identical to the decode kernels and to the two-kernel form above.  Heavy (late) token tiles are scheduled first.
// NQD = row quads (= warps) per CTA: 8 (32 rows) or 4 (16 rows).  The causal triangle makes the late token tiles the long
// poles of the launch (a 32-row CTA at T = 512 runs ~2x the average and as long as 3/4 of the whole kernel, ncu: SMs idle a
// third of the time); 16-row CTAs halve the longest CTA and fit three to an SM.
constexpr int PFF_TT = 128, PFF_KS = PFF_TT + 4, PFF_VT = 64;
inline int prefill_fused_scs(int t_max) { return ((t_max + 3) & ~3) + 2; }   // (pairs per score row) even; + 2 de-phases neighbouring pairs
inline size_t prefill_fused_smem(int hs, int nqd, int rw, int scs) {
    const size_t tile = (size_t)hs * PFF_KS > (size_t)2 * PFF_VT * hs ? (size_t)hs * PFF_KS : (size_t)2 * PFF_VT * hs;
    return ((size_t)nqd * (hs * 4 + 8) + tile + (size_t)2 * ((rw + 1) / 2) * scs + 64 + 8) * 4;
}
constexpr size_t PFF_SMEM_MAX = 216 * 1024;
// tokens per CTA for contexts up to t_max positions (0: the score rows of even one token do not fit)
inline int prefill_fused_tb(int hs, int nqd, int kv_mul, int t_max) {
    if (kv_mul > 4 * nqd) return 0;
    int tb = 4 * nqd / kv_mul;
    while (tb > 1 && prefill_fused_smem(hs, nqd, tb * kv_mul, prefill_fused_scs(t_max)) > PFF_SMEM_MAX) tb--;
    return prefill_fused_smem(hs, nqd, tb * kv_mul, prefill_fused_scs(t_max)) <= PFF_SMEM_MAX ? tb : 0;
}
inline int prefill_fused_nqd(int kv_mul, int want) { return (want == 4 && kv_mul <= 16) ? 4 : 8; }

// a separately rounded packed product (an add consumes it): fma(a, b, -0.0) with a run-time -0.0, see gemm.cuh f2_mul_sep
LMRS_DEVINL uint64_t pf2_mul(uint64_t a, uint64_t b, uint64_t nz2) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz2)); return r; }
LMRS_DEVINL uint64_t pf2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
LMRS_DEVINL uint64_t pf2_dup(float a) { uint64_t r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(a)); return r; }   // (ptxas folds it into the .F32 broadcast operand form)
LMRS_DEVINL void pf2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

template <int HS, int NQD>
__global__ void __launch_bounds__(32 * NQD, HS <= 128 ? (NQD == 4 ? 3 : 2) : 1) prefill_attn_fused_kernel(const PrefillAttnParams p, const int tb, const int SCS) {
    static_assert(HS % 32 == 0, "whole float4 per lane");
    constexpr int PFF_THREADS = 32 * NQD, PFF_ROWS = 4 * NQD;
    constexpr int KS = PFF_KS, TT = PFF_TT, VT = PFF_VT;
    constexpr int NQ = HS / 4, NQT = (NQ + 15) / 16;       // float4 chunks of a K / V row; V chunks per lane
    constexpr int KE = TT * NQ / PFF_THREADS;              // float4 of a K tile staged per thread
    constexpr int KPF = KE <= 8 ? KE : 8;                  // of those, fetched one tile ahead into registers (the rest at store time)
    constexpr int Q4S = HS * 4 + 8;                        // floats per row-quad row of q4
    constexpr int TILE_F = HS * KS > 2 * VT * HS ? HS * KS : 2 * VT * HS;
    extern __shared__ __align__(16) float pff[];
    float* q4 = pff;                                       // [8][HS] (pair A row0, row1, pair B row0, row1)
    float* tile = q4 + (PFF_ROWS / 4) * Q4S;               // [HS][KS] transposed K tile / [2][VT][HS] V tiles
    float* sc2 = tile + TILE_F;                            // [pairs][SCS] (score row0, score row1) pairs
    const int tid = threadIdx.x, g = blockIdx.y;
    const int ntile = (p.n + tb - 1) / tb;
    const int tix = ntile - 1 - (int)blockIdx.x;           // heavy tiles first
    const int tok0 = tix * tb;
    const int RW = tb * p.kv_mul;
    uint64_t* exp_tab = reinterpret_cast<uint64_t*>(sc2 + (size_t)2 * ((RW + 1) / 2) * SCS);
    const int rq = tid >> 5, lane = tid & 31;              // row quad (= warp), lane
    const int jq = lane & 15, rp = 2 * rq + (lane >> 4);   // softmax / a*v roles: lane j of 16 of row pair rp
    const int cta_T = min(p.pos + p.n, p.pos + tok0 + tb);
    const uint64_t nz2 = pf2_dup(p.neg_zero);
    auto row_T = [&](int r) { const int tk = tok0 + r / p.kv_mul; return (r < RW && tk < p.n) ? p.pos + tk + 1 : 0; };
    if (tid < 32) exp_tab[tid] = kExp2fTab[tid];
    {   // q of the quad, interleaved (A0, A1, B0, B1) per d; lanes cover the d chunks
        float4 qr[4];
        for (int c = lane; c < NQ; c += 32) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int r = 4 * rq + e, tl_ = r / p.kv_mul, h = r - tl_ * p.kv_mul, tk = tok0 + tl_;
                qr[e] = (r < RW && tk < p.n) ? *reinterpret_cast<const float4*>(p.q + (size_t)tk * p.att_dim + (size_t)(g * p.kv_mul + h) * HS + 4 * c)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4* d = reinterpret_cast<float4*>(q4 + (size_t)rq * Q4S + 16 * c);
            d[0] = make_float4(qr[0].x, qr[1].x, qr[2].x, qr[3].x); d[1] = make_float4(qr[0].y, qr[1].y, qr[2].y, qr[3].y);
            d[2] = make_float4(qr[0].z, qr[1].z, qr[2].z, qr[3].z); d[3] = make_float4(qr[0].w, qr[1].w, qr[2].w, qr[3].w);
        }
    }
    // ---- scores ------------------------------------------------------------------------------------------------------
    const int Tq = max(max(row_T(4 * rq), row_T(4 * rq + 1)), max(row_T(4 * rq + 2), row_T(4 * rq + 3)));   // longest row of the quad
    const int ntk = (cta_T + TT - 1) / TT;
    float4 kreg[KPF];
    auto k_src = [&](int tl, int u) -> float4 {   // element e: position j = e % TT (consecutive lanes: conflict-free transposing stores), chunk c = e / TT
        const int e = tid + u * PFF_THREADS, j = e % TT, c = e / TT, t = tl * TT + j;
        return t < cta_T ? *reinterpret_cast<const float4*>(p.kcache + (size_t)t * p.kv_dim + (size_t)g * HS + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto k_put = [&](int u, const float4 v) {
        const int e = tid + u * PFF_THREADS, j = e % TT, c = e / TT;
        tile[(4 * c + 0) * KS + j] = v.x; tile[(4 * c + 1) * KS + j] = v.y; tile[(4 * c + 2) * KS + j] = v.z; tile[(4 * c + 3) * KS + j] = v.w;
    };
#pragma unroll
    for (int u = 0; u < KPF; u++) kreg[u] = k_src(0, u);
    float mxr[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int tl = 0; tl < ntk; tl++) {
        __syncthreads();                                    // everybody finished the previous tile (and q4 is written)
#pragma unroll
        for (int u = 0; u < KPF; u++) k_put(u, kreg[u]);
#pragma unroll
        for (int u = KPF; u < KE; u++) k_put(u, k_src(tl, u));
        __syncthreads();
        if (tl + 1 < ntk) {                                 // in flight under this tile's arithmetic
#pragma unroll
            for (int u = 0; u < KPF; u++) kreg[u] = k_src(tl + 1, u);
        }
        const int t0 = tl * TT + 4 * lane;
        if (t0 < Tq) {
            const float* kt = tile + 4 * lane;
            const float* qq = q4 + (size_t)rq * Q4S;
            uint64_t acc[2][4];                             // [pair][position]: (row0, row1) chains, ascending d
#pragma unroll
            for (int u = 0; u < 4; u++) acc[0][u] = acc[1][u] = 0ull;
#pragma unroll 8
            for (int d = 0; d < HS; d++) {
                const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(qq + 4 * d);        // (A0, A1), (B0, B1)
                const float4 kk = *reinterpret_cast<const float4*>(kt + d * KS);
                acc[0][0] = pf2_add(acc[0][0], pf2_mul(q.x, pf2_dup(kk.x), nz2)); acc[1][0] = pf2_add(acc[1][0], pf2_mul(q.y, pf2_dup(kk.x), nz2));
                acc[0][1] = pf2_add(acc[0][1], pf2_mul(q.x, pf2_dup(kk.y), nz2)); acc[1][1] = pf2_add(acc[1][1], pf2_mul(q.y, pf2_dup(kk.y), nz2));
                acc[0][2] = pf2_add(acc[0][2], pf2_mul(q.x, pf2_dup(kk.z), nz2)); acc[1][2] = pf2_add(acc[1][2], pf2_mul(q.y, pf2_dup(kk.z), nz2));
                acc[0][3] = pf2_add(acc[0][3], pf2_mul(q.x, pf2_dup(kk.w), nz2)); acc[1][3] = pf2_add(acc[1][3], pf2_mul(q.y, pf2_dup(kk.w), nz2));
            }
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int pr_ = 2 * rq + s;
                if (2 * pr_ >= RW) continue;
                const int T0 = row_T(2 * pr_), T1 = row_T(2 * pr_ + 1), Tpp = max(T0, T1);
                float* srow_w = sc2 + (size_t)pr_ * SCS * 2;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    float sv[2];
                    pf2_unpack(acc[s][u], sv[0], sv[1]);
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        float score = __fdiv_rn(sv[e], p.sqrt_hs);
                        if (p.gemma) {   // soft-cap 50*tanh(s/50) in f64, window mask on every layer (:518-526)
                            score = __fdiv_rn(score, 50.0f);
                            score = (float)tanh((double)score);
                            score = __fmul_rn(score, 50.0f);
                            score = __fadd_rn(score, (p.mask_base - (uint32_t)(t0 + u) <= 4096u) ? 0.0f : -2.3819763e38f);
                        }
                        sv[e] = score;
                    }
                    if (t0 + u < T0) mxr[2 * s] = fmaxf(mxr[2 * s], sv[0]);
                    if (t0 + u < T1) mxr[2 * s + 1] = fmaxf(mxr[2 * s + 1], sv[1]);
                    if (t0 + u < Tpp) *reinterpret_cast<float2*>(srow_w + 2 * (t0 + u)) = make_float2(sv[0], sv[1]);
                }
            }
        }
    }
    // ---- softmax, warp-local: lanes 0..15 of the warp take the quad's first pair, lanes 16..31 the second --------------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int e = 0; e < 4; e++) mxr[e] = fmaxf(mxr[e], __shfl_xor_sync(0xffffffffu, mxr[e], o));
    const int Tr0 = row_T(2 * rp), Tr1 = row_T(2 * rp + 1), Tp = max(Tr0, Tr1);
    const float mx0 = (lane >> 4) ? mxr[2] : mxr[0], mx1 = (lane >> 4) ? mxr[3] : mxr[1];
    float* srow = sc2 + (size_t)(2 * rp < RW ? rp : 0) * SCS * 2;
    __syncwarp();
    for (int t = jq; t < Tp; t += 16) {
        float2 s = *reinterpret_cast<float2*>(srow + 2 * t);
        s.x = t < Tr0 ? expf_glibc_t(__fsub_rn(s.x, mx0), exp_tab) : 0.0f;
        s.y = t < Tr1 ? expf_glibc_t(__fsub_rn(s.y, mx1), exp_tab) : 0.0f;
        *reinterpret_cast<float2*>(srow + 2 * t) = s;
    }
    __syncwarp();
    float sum0 = 0.0f, sum1 = 0.0f;
    if (jq == 0 && Tp > 0) {   // ascending t, one dependent add per element and row (src/functional.rs:131-134); entries beyond a row's length are +0.0
        const float4* s4 = reinterpret_cast<const float4*>(srow);   // (t: row0, row1, t+1: row0, row1)
        int t = 0;
        float4 a = s4[0], b = Tp > 2 ? s4[1] : make_float4(0.f, 0.f, 0.f, 0.f);
        for (; t + 4 < Tp; t += 4) {                        // two float4 (four positions) per trip, the next two already loaded
            const float4 a1 = s4[t / 2 + 2], b1 = t + 6 < Tp ? s4[t / 2 + 3] : make_float4(0.f, 0.f, 0.f, 0.f);
            sum0 = __fadd_rn(sum0, a.x); sum1 = __fadd_rn(sum1, a.y); sum0 = __fadd_rn(sum0, a.z); sum1 = __fadd_rn(sum1, a.w);
            sum0 = __fadd_rn(sum0, b.x); sum1 = __fadd_rn(sum1, b.y); sum0 = __fadd_rn(sum0, b.z); sum1 = __fadd_rn(sum1, b.w);
            a = a1; b = b1;
        }
        // tail: up to four positions left in a, b; only those below Tp count (the rest of the buffer is not this step's data)
        const float tv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (t + k < Tp) { sum0 = __fadd_rn(sum0, tv[2 * k]); sum1 = __fadd_rn(sum1, tv[2 * k + 1]); }
    }
    sum0 = __shfl_sync(0xffffffffu, sum0, lane & 16); sum1 = __shfl_sync(0xffffffffu, sum1, lane & 16);
    for (int t = jq; t < Tp; t += 16) {
        float2 s = *reinterpret_cast<float2*>(srow + 2 * t);
        s.x = t < Tr0 ? __fdiv_rn(s.x, sum0) : 0.0f;
        s.y = t < Tr1 ? __fdiv_rn(s.y, sum1) : 0.0f;
        *reinterpret_cast<float2*>(srow + 2 * t) = s;
    }
    __syncthreads();                                        // all K-tile reads done: the tile buffers now carry V
    // ---- a*v --------------------------------------------------------------------------------------------------------------
    const int nvt = (cta_T + VT - 1) / VT;
    auto stage_v = [&](int tl) {
        float* vb = tile + (tl & 1) * VT * HS;
        for (int e = tid; e < VT * NQ; e += PFF_THREADS) {
            const int j = e / NQ, c = e - j * NQ, t = tl * VT + j;
            if (t < cta_T) cp_async16(vb + j * HS + 4 * c, p.vcache + (size_t)t * p.kv_dim + (size_t)g * HS + 4 * c);
        }
        cp_async_commit();
    };
    uint64_t av[NQT][4];
#pragma unroll
    for (int i = 0; i < NQT; i++) av[i][0] = av[i][1] = av[i][2] = av[i][3] = 0ull;
    stage_v(0);
    for (int tl = 0; tl < nvt; tl++) {
        if (tl + 1 < nvt) stage_v(tl + 1); else cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const float* vb = tile + (tl & 1) * VT * HS;
        const int rows = min(VT, Tp - tl * VT);
        const float* pr = srow + 2 * tl * VT;
        // one multiply and one dependent add per position and dim, ascending t (:533-542), both rows of the pair per instruction
        auto step = [&](int j, uint64_t pp) {
#pragma unroll
            for (int i = 0; i < NQT; i++) {
                const int c = jq + 16 * i;
                if (NQ % 16 == 0 || c < NQ) {
                    const float4 v = *reinterpret_cast<const float4*>(vb + j * HS + 4 * c);
                    av[i][0] = pf2_add(av[i][0], pf2_mul(pp, pf2_dup(v.x), nz2)); av[i][1] = pf2_add(av[i][1], pf2_mul(pp, pf2_dup(v.y), nz2));
                    av[i][2] = pf2_add(av[i][2], pf2_mul(pp, pf2_dup(v.z), nz2)); av[i][3] = pf2_add(av[i][3], pf2_mul(pp, pf2_dup(v.w), nz2));
                }
            }
        };
        int j = 0;
        for (; j + 2 <= rows; j += 2) {
            const ulonglong2 pp = *reinterpret_cast<const ulonglong2*>(pr + 2 * j);
            step(j, pp.x); step(j + 1, pp.y);
        }
        if (j < rows) step(j, *reinterpret_cast<const uint64_t*>(pr + 2 * j));
        __syncthreads();                                    // the buffer staged two tiles from now is this one
    }
    cp_async_wait<0>();
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const int r = 2 * rp + e, tl_ = r / p.kv_mul, h = r - tl_ * p.kv_mul, tk = tok0 + tl_;
        if (!(r < RW && tk < p.n)) continue;
        float* o = p.out + (size_t)tk * p.att_dim + (size_t)(g * p.kv_mul + h) * HS;
#pragma unroll
        for (int i = 0; i < NQT; i++) {
            const int c = jq + 16 * i;
            if (NQ % 16 == 0 || c < NQ) {
                float lo[4], hi[4];
#pragma unroll
                for (int k = 0; k < 4; k++) pf2_unpack(av[i][k], lo[k], hi[k]);
                *reinterpret_cast<float4*>(o + 4 * c) = e == 0 ? make_float4(lo[0], lo[1], lo[2], lo[3]) : make_float4(hi[0], hi[1], hi[2], hi[3]);
            }
        }
    }
}

}  // namespace lmrs
