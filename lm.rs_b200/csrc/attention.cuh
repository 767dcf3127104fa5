// attention.cuh -- attention for sm_100a: RoPE + QK^T + soft-cap/window + softmax + .V in ONE kernel over the
// HBM-resident f32 KV cache.  Replaces src/transformer.rs:443-544.
//
//   attn_cluster_kernel  decode steps: a thread-block cluster per KV head, scores exchanged through distributed shared
//                        memory (second half of this file)
//   attn_decode_kernel   one CTA per KV head (x chunks of 4 query heads): batched prefill rows, contexts beyond the
//                        cluster kernel's shared-memory buckets, and the body the persistent megakernel runs
//   attn_scores_kernel   q.k products of batched prefill spread over the whole GPU
//
// A CTA (or cluster) serves up to 4 query heads that share one KV head (GQA: K rows are staged once for all of them).
// RoPE uses cos/sin tables computed at load time on the HOST with the same libm calls as the reference (powf/cosf/sinf,
// src/transformer.rs:447-482).  The new K row arrives un-rotated in a staging row written by the QKV GEMV; the kernel
// rotates it, uses it from shared memory and stores it into the cache.  All f32 arithmetic follows the reference's
// operation order exactly (see the kernel comments).
#pragma once
#include "common.cuh"
#include "exact_math.cuh"
#include "gemv.cuh"

namespace lmrs {

constexpr int ATT_THREADS = 256;   // stand-alone kernel; the megakernel runs the same body with its own block size
constexpr int ATT_QH = 4;        // query heads per CTA (all sharing one KV head)
constexpr int ATT_SC_CAP = 2048; // positions whose scores are kept in shared memory (longer contexts: HBM scratch)
// K/V tile ring geometry.  BIG (decode, one CTA per kv head): 128-row tiles x 3 in flight -- per-tile fixed costs
// (barrier, cp.async issue, fix-ups) are a large part of the phase, measured 5.5% of the whole decode step;
// small (batched prefill, thousands of CTAs): 64-row tiles x 4, half the shared memory so two CTAs fit per SM.
template <bool BIG> __host__ __device__ constexpr int att_nt() { return BIG ? 3 : 4; }

struct AttnParams {
    const void* q;         // [att_dim] un-rotated                      (f32, or LL words when ll: common.cuh)
    const void* k_new;     // [kv_dim]  un-rotated K of this step        (same)
    const void* v_new;     // ll only: [kv_dim] V of this step (plain mode: the QKV GEMV wrote row `pos` of the cache itself)
    float* kcache;         // layer base [seq_len][kv_dim]
    float* vcache;         // layer base
    const float* rope_cos; // [seq_len][hs/2]
    const float* rope_sin;
    void* out;             // [att_dim]                                  (f32 or LL words)
    int ll, ll_nowait;     // decode chain: activations in/out are LL words, no griddepcontrol.wait
    float* scores;         // scratch [n_heads][seq_len] used when pos+1 > ATT_SC_CAP
    int kv_dim, kv_mul, chunks, gemma, seq_len;
    int batch, q_stride;   // batched prefill: grid.y = token index; pos = step->pos + blockIdx.y, q/out rows strided
    int scores_ready;      // scores already computed by attn_scores_kernel (global scratch [row][head][seq_len])
    float sqrt_hs;         // sqrtf(head_size): scores are DIVIDED by it (src/transformer.rs:516)
    int trace_slot;        // LMRS_TRACE builds: timeline slot of this launch (-1: none)
    int dev_skip;          // -DLMRS_DEV_PROBES builds only: phases to leave out when timing (results are then wrong)
    const StepParams* step;
    // L2 prefetch of the weights the NEXT kernels of the step will stream (decode chain): the attention phase is a latency
    // chain that leaves HBM idle, so its CTAs ask the L2 to fetch [l2pf_ptr, +l2pf_bytes) meanwhile (cp.async.bulk.prefetch.L2)
    const uint8_t* l2pf_ptr; unsigned long long l2pf_bytes; int l2pf_chunk;
};

template <int HS, bool BIG = false> __host__ __device__ constexpr int att_tile_rows() { return (BIG ? 128 : 64) / (HS <= 64 ? 1 : (HS <= 128 ? 2 : 4)); }
template <int HS, bool BIG = false> __host__ __device__ constexpr size_t attn_smem_bytes() {
    return (size_t)(ATT_QH * HS + 2 * HS + att_nt<BIG>() * att_tile_rows<HS, BIG>() * HS + ATT_QH * ATT_SC_CAP + 128 + 64) * 4;
}

LMRS_DEVINL void cp_async16(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
LMRS_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> LMRS_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// sum of sc[0..T) accumulated in ascending order with one dependent f32 add per element (src/functional.rs:131-134).
// Blocks of 8 values rotate through three statically indexed register sets (the loop body is unrolled three times), so
// two blocks of loads are always in flight ahead of the adds and no register moves sit behind a load.
LMRS_DEVINL float serial_sum_f32(const float* __restrict__ sc, const int T) {
    float sum = 0.0f;
    int t = 0;
    if (T >= 16) {
        float4 r[3][2];
        r[0][0] = *reinterpret_cast<const float4*>(sc); r[0][1] = *reinterpret_cast<const float4*>(sc + 4);
        r[1][0] = *reinterpret_cast<const float4*>(sc + 8); r[1][1] = *reinterpret_cast<const float4*>(sc + 12);
        for (; t + 40 <= T; t += 24) {
#pragma unroll
            for (int s = 0; s < 3; s++) {
                const int nx = (s + 2) % 3;
                r[nx][0] = *reinterpret_cast<const float4*>(sc + t + 8 * s + 16);
                r[nx][1] = *reinterpret_cast<const float4*>(sc + t + 8 * s + 20);
                sum = __fadd_rn(sum, r[s][0].x); sum = __fadd_rn(sum, r[s][0].y); sum = __fadd_rn(sum, r[s][0].z); sum = __fadd_rn(sum, r[s][0].w);
                sum = __fadd_rn(sum, r[s][1].x); sum = __fadd_rn(sum, r[s][1].y); sum = __fadd_rn(sum, r[s][1].z); sum = __fadd_rn(sum, r[s][1].w);
            }
        }
#pragma unroll
        for (int s = 0; s < 2; s++) {
            sum = __fadd_rn(sum, r[s][0].x); sum = __fadd_rn(sum, r[s][0].y); sum = __fadd_rn(sum, r[s][0].z); sum = __fadd_rn(sum, r[s][0].w);
            sum = __fadd_rn(sum, r[s][1].x); sum = __fadd_rn(sum, r[s][1].y); sum = __fadd_rn(sum, r[s][1].z); sum = __fadd_rn(sum, r[s][1].w);
        }
        t += 16;
    }
    for (; t < T; t++) sum = __fadd_rn(sum, sc[t]);
    return sum;
}

// x = sum_t a[t] * v[t * VS], one multiply and one dependent add per position in ascending order (src/transformer.rs:
// 533-542).  Blocks of 8 positions alternate between two statically indexed register sets: while the adds of block b
// run, the products of block b+1 are formed from registers loaded one block earlier and the loads of block b+2 are issued.
template <int VS>
LMRS_DEVINL float serial_av_f32(const float* __restrict__ pa, const float* __restrict__ tv, const int T) {
    float x = 0.0f;
    int t = 0;
    if (T >= 16) {
        float pr[2][8], v[2][8];
        float4 a[2][2];
        a[0][0] = *reinterpret_cast<const float4*>(pa); a[0][1] = *reinterpret_cast<const float4*>(pa + 4);
#pragma unroll
        for (int u = 0; u < 8; u++) v[0][u] = tv[u * VS];
        a[1][0] = *reinterpret_cast<const float4*>(pa + 8); a[1][1] = *reinterpret_cast<const float4*>(pa + 12);
#pragma unroll
        for (int u = 0; u < 8; u++) v[1][u] = tv[(8 + u) * VS];
        pr[0][0] = __fmul_rn(a[0][0].x, v[0][0]); pr[0][1] = __fmul_rn(a[0][0].y, v[0][1]); pr[0][2] = __fmul_rn(a[0][0].z, v[0][2]); pr[0][3] = __fmul_rn(a[0][0].w, v[0][3]);
        pr[0][4] = __fmul_rn(a[0][1].x, v[0][4]); pr[0][5] = __fmul_rn(a[0][1].y, v[0][5]); pr[0][6] = __fmul_rn(a[0][1].z, v[0][6]); pr[0][7] = __fmul_rn(a[0][1].w, v[0][7]);
        // invariant at the top of a half-trip s: pr[s] = products of block b, (a, v)[s ^ 1] = loaded values of block b+1
        for (; t + 32 <= T; t += 16) {
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int o = s ^ 1;
                const float* pn = pa + t + 8 * s + 16;
                const float* tn = tv + (size_t)(t + 8 * s + 16) * VS;
                a[s][0] = *reinterpret_cast<const float4*>(pn); a[s][1] = *reinterpret_cast<const float4*>(pn + 4);
#pragma unroll
                for (int u = 0; u < 8; u++) v[s][u] = tn[u * VS];
                pr[o][0] = __fmul_rn(a[o][0].x, v[o][0]); pr[o][1] = __fmul_rn(a[o][0].y, v[o][1]); pr[o][2] = __fmul_rn(a[o][0].z, v[o][2]); pr[o][3] = __fmul_rn(a[o][0].w, v[o][3]);
                pr[o][4] = __fmul_rn(a[o][1].x, v[o][4]); pr[o][5] = __fmul_rn(a[o][1].y, v[o][5]); pr[o][6] = __fmul_rn(a[o][1].z, v[o][6]); pr[o][7] = __fmul_rn(a[o][1].w, v[o][7]);
#pragma unroll
                for (int u = 0; u < 8; u++) x = __fadd_rn(x, pr[s][u]);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) x = __fadd_rn(x, pr[0][u]);
        x = __fadd_rn(x, __fmul_rn(a[1][0].x, v[1][0])); x = __fadd_rn(x, __fmul_rn(a[1][0].y, v[1][1])); x = __fadd_rn(x, __fmul_rn(a[1][0].z, v[1][2])); x = __fadd_rn(x, __fmul_rn(a[1][0].w, v[1][3]));
        x = __fadd_rn(x, __fmul_rn(a[1][1].x, v[1][4])); x = __fadd_rn(x, __fmul_rn(a[1][1].y, v[1][5])); x = __fadd_rn(x, __fmul_rn(a[1][1].z, v[1][6])); x = __fadd_rn(x, __fmul_rn(a[1][1].w, v[1][7]));
        t += 16;
    }
    for (; t < T; t++) x = __fadd_rn(x, __fmul_rn(pa[t], tv[(size_t)t * VS]));
    return x;
}

// two chains (two query heads sharing the V element) continued over `rows` positions of a staged tile; same rotation
template <int VS>
LMRS_DEVINL void serial_av2_f32(float& xa_io, float& xb_io, const float* __restrict__ pa, const float* __restrict__ pb,
                                const float* __restrict__ tv, const int rows) {
    float xa = xa_io, xb = xb_io;
    int t = 0;
    if (rows >= 16) {
        float pra[2][8], prb[2][8], v[2][8];
        float4 a[2][2], b[2][2];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            a[k][0] = *reinterpret_cast<const float4*>(pa + 8 * k); a[k][1] = *reinterpret_cast<const float4*>(pa + 8 * k + 4);
            b[k][0] = *reinterpret_cast<const float4*>(pb + 8 * k); b[k][1] = *reinterpret_cast<const float4*>(pb + 8 * k + 4);
#pragma unroll
            for (int u = 0; u < 8; u++) v[k][u] = tv[(8 * k + u) * VS];
        }
        auto products = [&](const int k) {
            pra[k][0] = __fmul_rn(a[k][0].x, v[k][0]); pra[k][1] = __fmul_rn(a[k][0].y, v[k][1]); pra[k][2] = __fmul_rn(a[k][0].z, v[k][2]); pra[k][3] = __fmul_rn(a[k][0].w, v[k][3]);
            pra[k][4] = __fmul_rn(a[k][1].x, v[k][4]); pra[k][5] = __fmul_rn(a[k][1].y, v[k][5]); pra[k][6] = __fmul_rn(a[k][1].z, v[k][6]); pra[k][7] = __fmul_rn(a[k][1].w, v[k][7]);
            prb[k][0] = __fmul_rn(b[k][0].x, v[k][0]); prb[k][1] = __fmul_rn(b[k][0].y, v[k][1]); prb[k][2] = __fmul_rn(b[k][0].z, v[k][2]); prb[k][3] = __fmul_rn(b[k][0].w, v[k][3]);
            prb[k][4] = __fmul_rn(b[k][1].x, v[k][4]); prb[k][5] = __fmul_rn(b[k][1].y, v[k][5]); prb[k][6] = __fmul_rn(b[k][1].z, v[k][6]); prb[k][7] = __fmul_rn(b[k][1].w, v[k][7]);
        };
        products(0);
        for (; t + 32 <= rows; t += 16) {
#pragma unroll
            for (int s = 0; s < 2; s++) {
                const int o = s ^ 1;
                const int nt = t + 8 * s + 16;
                a[s][0] = *reinterpret_cast<const float4*>(pa + nt); a[s][1] = *reinterpret_cast<const float4*>(pa + nt + 4);
                b[s][0] = *reinterpret_cast<const float4*>(pb + nt); b[s][1] = *reinterpret_cast<const float4*>(pb + nt + 4);
#pragma unroll
                for (int u = 0; u < 8; u++) v[s][u] = tv[(nt + u) * VS];
                products(o);
#pragma unroll
                for (int u = 0; u < 8; u++) { xa = __fadd_rn(xa, pra[s][u]); xb = __fadd_rn(xb, prb[s][u]); }
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) { xa = __fadd_rn(xa, pra[0][u]); xb = __fadd_rn(xb, prb[0][u]); }
        products(1);
#pragma unroll
        for (int u = 0; u < 8; u++) { xa = __fadd_rn(xa, pra[1][u]); xb = __fadd_rn(xb, prb[1][u]); }
        t += 16;
    }
    for (; t < rows; t++) {
        const float vv = tv[t * VS];
        xa = __fadd_rn(xa, __fmul_rn(pa[t], vv));
        xb = __fadd_rn(xb, __fmul_rn(pb[t], vv));
    }
    xa_io = xa; xb_io = xb;
}

// Bit-exact restatement of src/transformer.rs:501-544 for one token: every f32 operation happens in the
// reference's order (serial dot over d, serial softmax sum over t, serial a*v accumulation over t, separate
// mul and add, exp = glibc expf); only independent chains run in parallel.  One CTA per KV head (x chunks of 4
// query heads).  K and V stream through a 4-deep cp.async (LDGSTS) tile ring so that several tiles are in
// flight while the dependent-add chains of the current tile run; K tiles are stored with their 16-byte columns
// rotated by the row index, which makes the per-position LDS.128 dot products bank-conflict-free without
// touching the ascending-d summation order.  Latency floor: the two T-long dependent add chains (softmax sum,
// a*v) -- the price of exact parity, see exact_math.cuh.
template <int HS, int NTHR, bool BIG = false>
LMRS_DEVINL void attn_decode_body(const AttnParams& p, float* att_smem, const int kvh, const int h0, const int nh,
                                  const bool write_k) {
    constexpr int NWARP = NTHR / 32;
    constexpr int ATT_NT = att_nt<BIG>();
    constexpr int TILE = att_tile_rows<HS, BIG>();
    constexpr int C4 = HS / 4;                         // 16-byte chunks per row
    constexpr int CHUNKS = TILE * C4;                  // per tile
    constexpr int PERT = (CHUNKS + NTHR - 1) / NTHR;
    float* q_s = att_smem;                             // [ATT_QH][HS]
    float* k_s = q_s + ATT_QH * HS;                    // [HS] rotated new K row
    float* v_s = k_s + HS;                             // [HS] new V row (LL mode: arrives in a staging row, filed into the cache here)
    float* tile = v_s + HS;                            // [ATT_NT][TILE][HS]
    float* sc_s = tile + ATT_NT * TILE * HS;           // [ATT_QH][ATT_SC_CAP]
    float* red = sc_s + ATT_QH * ATT_SC_CAP;           // [128]: per-head per-warp maxima, then sums at [96..]
    uint64_t* exp_tab = reinterpret_cast<uint64_t*>(red + 128);   // [32] expf table copy

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int brow = p.batch ? (int)blockIdx.y : 0;
    const int pos = (int)p.step->pos + brow;
    const uint32_t mask_base = p.step->mask_base;
    const bool ll = p.ll != 0;                   // (never with batch)
    const bool nowait = p.ll_nowait != 0;
    const uint32_t seq = ll ? p.step->seq : 0u;
    const float* q_in = reinterpret_cast<const float*>(p.q) + (size_t)brow * p.q_stride;
    const llword_t* q_ll = reinterpret_cast<const llword_t*>(p.q);
    float* out_row = reinterpret_cast<float*>(p.out) + (size_t)brow * p.q_stride;
    llword_t* out_ll = reinterpret_cast<llword_t*>(p.out);
    const bool have_knew = p.k_new != nullptr;   // decode: K row of this step arrives un-rotated in a staging row
    const bool have_vnew = ll;                   // ... and so does the V row in LL mode
    const int T = pos + 1;
    const bool in_smem = T <= ATT_SC_CAP;
    float* const sc_glob = p.scores + ((size_t)brow * p.kv_mul * (gridDim.x / p.chunks) + h0) * p.seq_len;
    const int ntiles = (T + TILE - 1) / TILE;

    // everything below runs once, specialised on where the scores live: a pointer the compiler can prove to be shared
    // memory turns every access of the softmax and of the a*v chains into LDS/STS instead of generic LD/ST
    auto run = [&](float* const sc_base, const int sc_stride) {
    // tile tl of K (rotated columns) or V (plain) -> ring slot tl % ATT_NT; rows >= T and the row `pos` of K
    // (not in the cache yet) are skipped; always commits a group so the wait counts stay uniform
    auto issue_tile = [&](const float* base, int tl, bool rotate) {
        if (tl < ntiles) {
            float* tb = tile + (tl % ATT_NT) * TILE * HS;
#pragma unroll
            for (int i = 0; i < PERT; i++) {
                const int e = tid + i * NTHR;
                if (e < CHUNKS) {
                    const int r = e / C4, c = e - r * C4, t = tl * TILE + r;
                    if (t < T && !((rotate ? have_knew : have_vnew) && t == pos)) {
                        int cc = c;
                        if (rotate) { cc = c + r; cc = cc % C4; }
                        cp_async16(tb + r * HS + cc * 4, base + (size_t)t * p.kv_dim + (size_t)kvh * HS + c * 4);
                    }
                }
            }
        }
        cp_async_commit();
    };
    if (warp == NWARP - 1) exp_tab[lane] = kExp2fTab[lane];   // by the last warp: warp 0 must not stall before the K prefetch
    trace_event(200);
    // K tiles start flowing before anything else (rows < pos are in the cache since earlier steps)
    if (!p.scores_ready) {
#pragma unroll
        for (int k = 0; k < ATT_NT - 1; k++) issue_tile(p.kcache, k, true);
    }

    if (ll) ll_canary_wait(q_ll + (size_t)h0 * HS, seq, nowait);   // park the CTA until the QKV kernel starts delivering
    // RoPE on q and on the new k row (rotate-half pairs j, j+HS/2), src/transformer.rs:480-492
    const float* cs = p.rope_cos + (size_t)pos * (HS / 2);
    const float* sn = p.rope_sin + (size_t)pos * (HS / 2);
    for (int i = tid; i < nh * (HS / 2); i += NTHR) {
        const int h = i / (HS / 2), j = i - h * (HS / 2);
        const float fcr = cs[j], fci = sn[j];
        const size_t qi = (size_t)(h0 + h) * HS + j;
        const float v0 = ll ? ll_wait1(q_ll + qi, seq, nowait) : __ldcg(q_in + qi);
        const float v1 = ll ? ll_wait1(q_ll + qi + HS / 2, seq, nowait) : __ldcg(q_in + qi + HS / 2);
        q_s[h * HS + j] = p.batch ? v0 : __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci));       // batched prefill: q rows
        q_s[h * HS + j + HS / 2] = p.batch ? v1 : __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr));   // were rotated by rope_rows_kernel
    }
    if (have_knew && !p.scores_ready)
    for (int j = tid; j < HS / 2; j += NTHR) {
        const float fcr = cs[j], fci = sn[j];
        const size_t ki = (size_t)kvh * HS + j;
        const float v0 = ll ? ll_wait1(reinterpret_cast<const llword_t*>(p.k_new) + ki, seq, nowait) : __ldcg(reinterpret_cast<const float*>(p.k_new) + ki);
        const float v1 = ll ? ll_wait1(reinterpret_cast<const llword_t*>(p.k_new) + ki + HS / 2, seq, nowait) : __ldcg(reinterpret_cast<const float*>(p.k_new) + ki + HS / 2);
        const float r0 = __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci));
        const float r1 = __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr));
        k_s[j] = r0; k_s[j + HS / 2] = r1;
        if (write_k) {   // exactly one CTA per KV head publishes the rotated row into the cache
            p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + j] = r0;
            p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + j + HS / 2] = r1;
        }
    }
    if (have_vnew)
    for (int d = tid; d < HS; d += NTHR) {
        const float vv = ll_wait1(reinterpret_cast<const llword_t*>(p.v_new) + (size_t)kvh * HS + d, seq, nowait);
        v_s[d] = vv;
        if (write_k) p.vcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + d] = vv;
    }
    __syncthreads();

    trace_event(201);
    // ---- scores: s[h][t] = (sum_d q[h][d]*k[t][d]) / sqrt(hs)   (:507-528) --------------------------------
    if (p.scores_ready) {   // dot products were spread over the whole GPU by attn_scores_kernel
        const float* sg = p.scores + ((size_t)brow * p.kv_mul * (gridDim.x / p.chunks) + h0) * p.seq_len;
        if (in_smem)
            for (int h = 0; h < nh; h++)
                for (int t = tid; t < T; t += NTHR) sc_s[h * ATT_SC_CAP + t] = __ldcg(sg + (size_t)h * p.seq_len + t);
    }
    for (int tl = 0; tl < (p.scores_ready ? 0 : ntiles); tl++) {
        cp_async_wait<ATT_NT - 2>();                   // this thread's copies of tile tl have landed
        float* tb = tile + (tl % ATT_NT) * TILE * HS;
        if (have_knew && pos / TILE == tl) {           // the new K row comes from shared memory, same rotated layout
            const int r = pos - tl * TILE;
            for (int d = tid; d < HS; d += NTHR) {
                const int c = d >> 2;
                tb[r * HS + ((c + r) % C4) * 4 + (d & 3)] = k_s[d];
            }
        }
        __syncthreads();                               // tile tl visible; everyone is done with tile tl-1
        issue_tile(p.kcache, tl + ATT_NT - 1, true);   // refill the slot tile tl-1 used
        const int rows = min(TILE, T - tl * TILE);
        // thread = (row r, pair of heads): one LDS.128 of K feeds two independent dot-product chains (ILP 2),
        // products of chunk d4+1 are formed while chunk d4's dependent adds run
        const int npair = (nh + 1) / 2;
        for (int idx = tid; idx < rows * npair; idx += NTHR) {
            const int hp = idx / rows, r = idx - hp * rows, t = tl * TILE + r;
            const int ha = hp * 2, hb = min(hp * 2 + 1, nh - 1);
            const float4* qa = reinterpret_cast<const float4*>(q_s + ha * HS);
            const float4* qb = reinterpret_cast<const float4*>(q_s + hb * HS);
            const float4* k4 = reinterpret_cast<const float4*>(tb + r * HS);
            float sa = 0.0f, sb = 0.0f;
            // the whole K row first (C4 independent LDS.128: their latency overlaps), un-rotating on the fly
            float4 kr[C4];
            {
                int c = r % C4;
#pragma unroll
                for (int d4 = 0; d4 < C4; d4++) { kr[d4] = k4[c]; c = (c + 1 == C4) ? 0 : c + 1; }
            }
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
#pragma unroll
            for (int w = 0; w < 2; w++) {
                const int h = w ? hb : ha;
                if (w == 1 && hb == ha) break;
                float score = __fdiv_rn(w ? sb : sa, p.sqrt_hs);
                if (p.gemma) {   // soft-cap 50*tanh(s/50) in f64, window mask on every layer (:518-526)
                    score = __fdiv_rn(score, 50.0f);
                    score = (float)tanh((double)score);
                    score = __fmul_rn(score, 50.0f);
                    score = __fadd_rn(score, (mask_base - (uint32_t)t <= 4096u) ? 0.0f : -2.3819763e38f);
                }
                sc_base[(size_t)h * sc_stride + t] = score;
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    trace_event(202);
    // V tiles start flowing now; the softmax below runs while they land
#pragma unroll
    for (int k = 0; k < ATT_NT - 1; k++) issue_tile(p.vcache, k, false);

    // ---- softmax (src/functional.rs:122-140): max, exp(x-max), serial sum, divide ---------------------------
    for (int h = 0; h < nh; h++) {
        float* sc = sc_base + (size_t)h * sc_stride;
        float mx = sc[0];
        for (int t = tid; t < T; t += NTHR) mx = fmaxf(mx, sc[t]);
        mx = warp_max(mx);
        if (lane == 0) red[h * NWARP + warp] = mx;
    }
    __syncthreads();
    for (int h = 0; h < nh; h++) {
        float* sc = sc_base + (size_t)h * sc_stride;
        float mx = red[h * NWARP];
#pragma unroll
        for (int w = 1; w < NWARP; w++) mx = fmaxf(mx, red[h * NWARP + w]);
        for (int t = tid; t < T; t += NTHR) sc[t] = expf_glibc_t(__fsub_rn(sc[t], mx), exp_tab);
    }
    __syncthreads();
    trace_event(203);
    if (lane == 0 && warp < nh)   // the reference's `sum += x[i]` chain: one thread per head, in different warps
        red[96 + warp] = serial_sum_f32(sc_base + (size_t)warp * sc_stride, T);
    __syncthreads();
    trace_event(204);
    for (int h = 0; h < nh; h++) {
        float* sc = sc_base + (size_t)h * sc_stride;
        const float sum = red[96 + h];
        for (int t = tid; t < T; t += NTHR) sc[t] = __fdiv_rn(sc[t], sum);
    }
    // (the barrier inside the first V-tile iteration orders these writes before the chains read them)

    // ---- out[h][d] = sum_t a[h][t] * v[t][d], serial over t (:533-542) --------------------------------------
    trace_event(205);
    // thread = (dim d, pair of heads): one V element feeds two independent accumulation chains (ILP 2); the
    // products of the next four rows are formed while the current four dependent adds run
    constexpr int NPMAX = ATT_QH / 2;
    constexpr int MAXCH = (NPMAX * HS + NTHR - 1) / NTHR;   // (d, head-pair) items per thread
    const int npair = (nh + 1) / 2;
    float acca[MAXCH], accb[MAXCH];
#pragma unroll
    for (int k = 0; k < MAXCH; k++) { acca[k] = 0.0f; accb[k] = 0.0f; }
    for (int tl = 0; tl < ntiles; tl++) {
        cp_async_wait<ATT_NT - 2>();
        if (have_vnew && pos / TILE == tl) {           // the new V row comes from shared memory
            float* tw = tile + (tl % ATT_NT) * TILE * HS + (pos - tl * TILE) * HS;
            for (int d = tid; d < HS; d += NTHR) tw[d] = v_s[d];
        }
        __syncthreads();
        issue_tile(p.vcache, tl + ATT_NT - 1, false);
        const float* tb = tile + (tl % ATT_NT) * TILE * HS;
        const int rows = min(TILE, T - tl * TILE);
#pragma unroll
        for (int k = 0; k < MAXCH; k++) {
            const int idx = tid + k * NTHR;
            if (idx < npair * HS) {
                const int hp = idx / HS, d = idx - hp * HS;
                const int ha = hp * 2, hb = min(hp * 2 + 1, nh - 1);
                const float* pa = sc_base + (size_t)ha * sc_stride + tl * TILE;
                const float* pb = sc_base + (size_t)hb * sc_stride + tl * TILE;
                const float* tv = tb + d;
                float xa = acca[k], xb = accb[k];
                serial_av2_f32<HS>(xa, xb, pa, pb, tv, rows);
                acca[k] = xa; accb[k] = xb;
            }
        }
    }
    cp_async_wait<0>();
    trace_event(206);
#pragma unroll
    for (int k = 0; k < MAXCH; k++) {
        const int idx = tid + k * NTHR;
        if (idx < npair * HS) {
            const int hp = idx / HS, d = idx - hp * HS;
            if (ll) {
                ll_store(out_ll + (size_t)(h0 + hp * 2) * HS + d, acca[k], seq);
                if (hp * 2 + 1 < nh) ll_store(out_ll + (size_t)(h0 + hp * 2 + 1) * HS + d, accb[k], seq);
            } else {
                out_row[(size_t)(h0 + hp * 2) * HS + d] = acca[k];
                if (hp * 2 + 1 < nh) out_row[(size_t)(h0 + hp * 2 + 1) * HS + d] = accb[k];
            }
        }
    }
    };
    if (in_smem) run(sc_s, ATT_SC_CAP);
    else run(sc_glob, p.seq_len);
    __syncthreads();   // the tile ring / score buffers may be reused by the caller
}

template <int HS, bool BIG>
__global__ void __launch_bounds__(ATT_THREADS) attn_decode_kernel(const AttnParams p) {
    extern __shared__ __align__(16) float att_smem_dyn[];
    const int kvh = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks;
    const int h0 = kvh * p.kv_mul + chunk * ATT_QH;                  // first query head of this CTA
    const int nh = min(ATT_QH, p.kv_mul - chunk * ATT_QH);           // query heads served here
    pdl_launch_dependents();
    if (!p.ll) pdl_wait();
    attn_decode_body<HS, ATT_THREADS, BIG>(p, att_smem_dyn, kvh, h0, nh, chunk == 0);
}

// ---- cluster decode attention -------------------------------------------------------------------------------------
// The single-CTA kernel above is bounded by what ONE SM can pull from L2 (~40 B/clk) and by running every phase of a
// whole KV head on one SM.  Here a thread-block CLUSTER of CL CTAs serves one KV head (nh <= 4 query heads):
//   scores : CTA `rank` owns positions [rank*RP, rank*RP+RP) -- 1/CL of the K rows, staged with cp.async (LDGSTS) --
//            and PUSHES every score it computes into the shared memory of the CTAs that need it with st.async
//            (SASS STAS), which also counts the bytes on the receiver's mbarrier: a consumer simply waits until its
//            complete score rows have arrived -- no cluster-wide barrier (ptxas lowers barrier.cluster.arrive.release
//            to MEMBAR.ALL.GPU) and no remote reads
//   softmax: the cluster is split into G head groups of R = CL/G CTAs; a CTA runs max / exp / serial sum / divide
//            only for its group's nh/G heads (G = nh for Llama-style GQA: one head per CTA pair)
//   a*v    : CTA (group, part) owns output dims [part*HS/R, ...) of its heads -- that slice of every V row is
//            prefetched during the phases above
// The arithmetic and its order are those of attn_decode_body (bit-exact with src/transformer.rs:501-544); only the
// placement of independent chains changes.  Everything a CTA needs lives in shared memory sized by `cap` (the graph
// variant's position bucket), so contexts beyond the largest bucket use the single-CTA kernel.
LMRS_DEVINL uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
LMRS_DEVINL void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
LMRS_DEVINL void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// push one f32 into CTA `rank`'s shared memory (same offsets as mine) and count its 4 bytes on that CTA's mbarrier:
// st.async (SASS STAS) -- the receiver just waits for its expected byte count, no cluster-wide barrier or fence
LMRS_DEVINL void st_async_f32(const float* local_ptr, const uint64_t* local_bar, uint32_t rank, float v) {
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(__float_as_uint(v)), "r"(rb)
                 : "memory");
}

constexpr int ATTC_MAX_CL = 8;     // portable cluster size
// floats of dynamic shared memory for positions <= cap (cap % 32 == 0), clusters of cl CTAs in g head groups serving nh heads
__host__ __device__ constexpr size_t attc_smem_floats(int hs, int cap, int cl, int g, int nh) {
    return (size_t)ATT_QH * hs + (size_t)(nh / g) * (cap + 4) + (size_t)(cap / cl) * hs + (size_t)cap * (hs * g / cl) + 128 + 64 + 4;
}

// flattened (head, position) loops with two independent evaluations in flight per thread (the f64 exp and the IEEE
// divide are long dependent instruction sequences: measured 20-25% faster than one at a time, more in flight is slower)
template <typename F>
LMRS_DEVINL void rowwise_ilp2(float* sc_s, const int nhl, const int SCS, const int T, const int tid, const int nthr, F f) {
    const int total = nhl * T;
    for (int i0 = tid; i0 < total; i0 += 2 * nthr) {
        const int i1 = i0 + nthr;
        const int ha = i0 / T, ta = i0 - ha * T;
        const bool two = i1 < total;
        const int hb = two ? i1 / T : ha, tb = two ? i1 - hb * T : ta;
        const float xa = sc_s[ha * SCS + ta], xb = sc_s[hb * SCS + tb];
        const float ya = f(xa, ha), yb = f(xb, hb);
        sc_s[ha * SCS + ta] = ya;
        if (two) sc_s[hb * SCS + tb] = yb;
    }
}

template <int HS, int CL, int G>
__global__ void __launch_bounds__(ATT_THREADS) attn_cluster_kernel(const AttnParams p, const int cap) {
    constexpr int NTHR = ATT_THREADS, NWARP = NTHR / 32, C4 = HS / 4;
    constexpr int R = CL / G;                  // CTAs per head group
    constexpr int DS = HS / R, DC = DS / 4;    // output dims (and 16-byte chunks per V row) owned by this CTA
    static_assert(CL % G == 0 && HS % (4 * R) == 0, "every CTA owns whole 16-byte chunks of a V row");
    extern __shared__ __align__(16) float att_smem_dyn[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int cid = blockIdx.x / CL;
    const int kvh = cid / p.chunks, chunk = cid % p.chunks;
    const int h0 = kvh * p.kv_mul + chunk * ATT_QH;
    const int nh = min(ATT_QH, p.kv_mul - chunk * ATT_QH);   // the host guarantees nh % G == 0
    const int nhl = nh / G;                    // heads of my group
    const int grp = rank / R, part = rank - grp * R;
#ifdef LMRS_DEV_PROBES
    const int skip = p.dev_skip;
#else
    constexpr int skip = 0;
#endif
    const int SCS = cap + 4;                   // score row stride: +4 floats puts the heads' float4 reads on distinct banks
    float* q_s = att_smem_dyn;                         // [ATT_QH][HS]  all heads of the KV head (scores need them all)
    float* sc_s = q_s + ATT_QH * HS;                   // [nhl][SCS]    complete score rows of my group's heads
    float* kt = sc_s + nhl * SCS;                      // [cap/CL][HS]  this CTA's K rows, 16-byte columns rotated by the row
    float* vt = kt + (size_t)(cap / CL) * HS;          // [cap][DS]     this CTA's slice of every V row
    float* red = vt + (size_t)cap * DS;                // [128]
    uint64_t* exp_tab = reinterpret_cast<uint64_t*>(red + 128);
    uint64_t* push_bar = exp_tab + 32;                 // counts the score bytes pushed into this CTA

    // positions < pos were written by earlier steps (complete: a step's first kernel is never launched programmatically),
    // so their K/V rows are requested BEFORE the dependency wait and land while the QKV GEMV is still finishing
    const long long c0 = ktrace_c0();
    const int pos = (int)p.step->pos;
    const uint32_t mask_base = p.step->mask_base;
    const int T = pos + 1;
    if (tid == 0) {                                    // this CTA will receive the complete score rows of its heads
        mbar_init(push_bar, 1);
        fence_barrier_init();
        mbar_expect_tx(push_bar, (uint32_t)(nhl * T * 4));
    }
    cluster_arrive();                                  // "my shared memory and barrier exist": peers push after their wait
    const int RP = (((T + CL - 1) / CL) + 3) & ~3;     // positions per CTA
    const int r0 = min(T, rank * RP), r1 = min(T, r0 + RP), myrows = r1 - r0;
    {
        const float* kb = p.kcache + (size_t)kvh * HS;
        for (int e = tid; e < ((skip & 64) ? 0 : myrows * C4); e += NTHR) {
            const int r = e / C4, c = e - r * C4, t = r0 + r;
            if (t != pos) cp_async16(kt + r * HS + ((c + r) % C4) * 4, kb + (size_t)t * p.kv_dim + c * 4);
        }
        cp_async_commit();
        const float* vb = p.vcache + (size_t)kvh * HS + part * DS;
        for (int e = tid; e < ((skip & 64) ? 0 : pos * DC); e += NTHR) {
            const int t = e / DC, c = e - t * DC;
            cp_async16(vt + t * DS + c * 4, vb + (size_t)t * p.kv_dim + c * 4);
        }
        cp_async_commit();
    }
    // RoPE factors of this position (constant tables: src/transformer.rs:447-482 evaluated on the host at load time)
    const float* cs = p.rope_cos + (size_t)pos * (HS / 2);
    const float* sn = p.rope_sin + (size_t)pos * (HS / 2);
    float fcr[(ATT_QH * (HS / 2) + NTHR - 1) / NTHR], fci[(ATT_QH * (HS / 2) + NTHR - 1) / NTHR];
#pragma unroll
    for (int k = 0; k < (ATT_QH * (HS / 2) + NTHR - 1) / NTHR; k++) {
        const int j = (tid + k * NTHR) % (HS / 2);
        fcr[k] = cs[j]; fci[k] = sn[j];
    }
    if (warp == NWARP - 1) exp_tab[lane] = kExp2fTab[lane];
    if (blockIdx.x == 0 && tid == 0) ktrace(p.trace_slot, 0);
    pdl_launch_dependents();
    if (p.l2pf_bytes && tid == 32 * (NWARP - 2)) l2_prefetch_slice(p.l2pf_ptr, p.l2pf_bytes, p.l2pf_chunk);   // after this CTA's own K/V requests
    const bool ll = p.ll != 0, nowait = p.ll_nowait != 0;
    const uint32_t seq = ll ? p.step->seq : 0u;
    if (!ll) pdl_wait();
    else ll_canary_wait(reinterpret_cast<const llword_t*>(p.q) + (size_t)h0 * HS, seq, nowait);   // park until the QKV kernel delivers
    if (tid == 0) { ktrace(p.trace_slot, 1); ktrace_c(p.trace_slot, 1, c0); }
    // this step's inputs: q (all heads of the KV head), the new K row (only the CTA that scores position `pos`) and the new
    // V row slice.  LL mode: every load of a thread is issued before any is validated (one L2 round trip on the critical
    // path, not one per input), and only incomplete words are polled again.
    constexpr int QK = (ATT_QH * (HS / 2) + NTHR - 1) / NTHR;
    const bool own_pos = pos >= r0 && pos < r1;        // exactly one CTA of the cluster scores (and publishes) the new row
    float kn0 = 0.0f, kn1 = 0.0f;
    float qv0[QK], qv1[QK];
    if (!ll) {
        if (tid < DC) cp_async16(vt + pos * DS + tid * 4, p.vcache + (size_t)pos * p.kv_dim + (size_t)kvh * HS + part * DS + tid * 4);
        if (own_pos && tid < HS / 2) {
            const float* kn = reinterpret_cast<const float*>(p.k_new) + (size_t)kvh * HS + tid;
            kn0 = __ldcg(kn); kn1 = __ldcg(kn + HS / 2);
        }
#pragma unroll
        for (int k = 0; k < QK; k++) {
            const int i = tid + k * NTHR;
            qv0[k] = qv1[k] = 0.0f;
            if (i < nh * (HS / 2)) {
                const int h = i / (HS / 2), jj = i - h * (HS / 2);
                const float* qp = reinterpret_cast<const float*>(p.q) + (size_t)(h0 + h) * HS + jj;
                qv0[k] = __ldcg(qp); qv1[k] = __ldcg(qp + HS / 2);
            }
        }
    } else {
        const llword_t* vn = reinterpret_cast<const llword_t*>(p.v_new) + (size_t)kvh * HS + part * DS + tid * 4;
        const llword_t* kn = reinterpret_cast<const llword_t*>(p.k_new) + (size_t)kvh * HS + tid;
        const bool want_v = tid < DC, want_k = own_pos && tid < HS / 2;
        bool got_v = !want_v, got_k = !want_k, got_q[QK];
        float4 vv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < QK; k++) { got_q[k] = !(tid + k * NTHR < nh * (HS / 2)); qv0[k] = qv1[k] = 0.0f; }
        const LLSpin sp = ll_spin_begin();
        for (;;) {
            bool all = true;
            llword_t wk0 = 0, wk1 = 0, wq0[QK], wq1[QK];
            if (!got_k) { wk0 = ll_ld1(kn); wk1 = ll_ld1(kn + HS / 2); }
#pragma unroll
            for (int k = 0; k < QK; k++)
                if (!got_q[k]) {
                    const int i = tid + k * NTHR, h = i / (HS / 2), jj = i - h * (HS / 2);
                    const llword_t* qp = reinterpret_cast<const llword_t*>(p.q) + (size_t)(h0 + h) * HS + jj;
                    wq0[k] = ll_ld1(qp); wq1[k] = ll_ld1(qp + HS / 2);
                }
            if (!got_v) { if (ll_try4(vn, seq, nowait, vv)) got_v = true; else all = false; }
            if (!got_k) { if (nowait || (ll_ok(wk0, seq) && ll_ok(wk1, seq))) { kn0 = ll_val(wk0); kn1 = ll_val(wk1); got_k = true; } else all = false; }
#pragma unroll
            for (int k = 0; k < QK; k++)
                if (!got_q[k]) { if (nowait || (ll_ok(wq0[k], seq) && ll_ok(wq1[k], seq))) { qv0[k] = ll_val(wq0[k]); qv1[k] = ll_val(wq1[k]); got_q[k] = true; } else all = false; }
            if (all) break;
            __nanosleep(20);
            ll_spin_check(sp);
        }
        if (want_v) {
            *reinterpret_cast<float4*>(vt + pos * DS + tid * 4) = vv;
            if (grp == 0 && chunk == 0) *reinterpret_cast<float4*>(p.vcache + (size_t)pos * p.kv_dim + (size_t)kvh * HS + part * DS + tid * 4) = vv;   // filed for later steps
        }
    }
    cp_async_commit();

    // RoPE on q and on the new k row (src/transformer.rs:480-492)
#pragma unroll
    for (int k = 0; k < QK; k++) {
        const int i = tid + k * NTHR;
        if (i < nh * (HS / 2)) {
            const int h = i / (HS / 2), jj = i - h * (HS / 2);
            q_s[h * HS + jj] = __fsub_rn(__fmul_rn(qv0[k], fcr[k]), __fmul_rn(qv1[k], fci[k]));
            q_s[h * HS + jj + HS / 2] = __fadd_rn(__fmul_rn(qv0[k], fci[k]), __fmul_rn(qv1[k], fcr[k]));
        }
    }
    if (own_pos && tid < HS / 2) {                      // (tid < HS/2 <= NTHR/2: these threads hold factor j = tid in slot 0)
        const int j = tid, j1 = j + HS / 2, r = pos - r0;
        const float k0 = __fsub_rn(__fmul_rn(kn0, fcr[0]), __fmul_rn(kn1, fci[0]));
        const float k1 = __fadd_rn(__fmul_rn(kn0, fci[0]), __fmul_rn(kn1, fcr[0]));
        kt[r * HS + (((j >> 2) + r) % C4) * 4 + (j & 3)] = k0;
        kt[r * HS + (((j1 >> 2) + r) % C4) * 4 + (j1 & 3)] = k1;
        kn0 = k0; kn1 = k1;                            // published into the cache at the end of the kernel
    }
    if (tid == 0) ktrace_c(p.trace_slot, 2, c0);       // rope done (thread 0)
    cp_async_wait<2>();                                // this thread's K copies have landed
    __syncthreads();
    cluster_wait();                                    // every peer is running: its shared memory and barrier may be written
    if (tid == 0) ktrace_c(p.trace_slot, 3, c0);       // K visible

    // ---- scores of this CTA's positions (:507-528): thread = (row, pair of heads), two dot-product chains ------
    {
        const int npair = (nh + 1) / 2;
        for (int idx = tid; idx < ((skip & 1) ? 0 : myrows * npair); idx += NTHR) {
            const int hp = idx / myrows, r = idx - hp * myrows, t = r0 + r;
            const int ha = hp * 2, hb = min(hp * 2 + 1, nh - 1);
            const float4* qa = reinterpret_cast<const float4*>(q_s + ha * HS);
            const float4* qb = reinterpret_cast<const float4*>(q_s + hb * HS);
            const float4* k4 = reinterpret_cast<const float4*>(kt + r * HS);
            float sa = 0.0f, sb = 0.0f;
            float4 kr[C4];
            {
                int c = r % C4;
#pragma unroll
                for (int d4 = 0; d4 < C4; d4++) { kr[d4] = k4[c]; c = (c + 1 == C4) ? 0 : c + 1; }
            }
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
#pragma unroll
            for (int w = 0; w < 2; w++) {
                const int h = w ? hb : ha;
                if (w == 1 && hb == ha) break;
                float score = __fdiv_rn(w ? sb : sa, p.sqrt_hs);
                if (p.gemma) {   // soft-cap 50*tanh(s/50) in f64, window mask on every layer (:518-526)
                    score = __fdiv_rn(score, 50.0f);
                    score = (float)tanh((double)score);
                    score = __fmul_rn(score, 50.0f);
                    score = __fadd_rn(score, (mask_base - (uint32_t)t <= 4096u) ? 0.0f : -2.3819763e38f);
                }
                const int g = h / nhl, hl = h - g * nhl;   // deliver to the R CTAs of the head's group (maybe myself)
                const float* slot = sc_s + hl * SCS + t;
#pragma unroll
                for (int c = 0; c < R; c++) st_async_f32(slot, push_bar, (uint32_t)(g * R + c), score);
            }
        }
    }
    if (tid == 0) ktrace_c(p.trace_slot, 4, c0);       // own scores done (thread 0)
    if (!(skip & 1)) mbar_wait(push_bar, 0);           // every score of my heads has landed: complete rows, no remote access below
    if (tid == 0) { ktrace(p.trace_slot, 2); ktrace_c(p.trace_slot, 5, c0); }   // scores complete

    // ---- softmax (src/functional.rs:122-140) of my group's heads: max, exp(x-max), serial sum, divide ---------
    {
        float mx[ATT_QH];
#pragma unroll
        for (int h = 0; h < ATT_QH; h++) mx[h] = h < nhl ? sc_s[h * SCS] : 0.0f;
        for (int t = tid; t < T; t += NTHR)
#pragma unroll
            for (int h = 0; h < ATT_QH; h++) if (h < nhl) mx[h] = fmaxf(mx[h], sc_s[h * SCS + t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int h = 0; h < ATT_QH; h++) mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], o));   // independent chains
        if (lane == 0)
#pragma unroll
            for (int h = 0; h < ATT_QH; h++) if (h < nhl) red[h * NWARP + warp] = mx[h];
    }
    __syncthreads();
    if (tid < ATT_QH) {
        float m = red[tid * NWARP];
#pragma unroll
        for (int w = 1; w < NWARP; w++) m = fmaxf(m, red[tid * NWARP + w]);
        red[64 + tid] = m;
    }
    __syncthreads();
    if (!(skip & 2))
        rowwise_ilp2(sc_s, nhl, SCS, T, tid, NTHR, [&](float x, int h) { return expf_glibc_t(__fsub_rn(x, red[64 + h]), exp_tab); });
    __syncthreads();
    if (tid == 0) ktrace_c(p.trace_slot, 7, c0);       // max + exp done
    if (lane == 0 && warp < nhl && !(skip & 4))   // the reference's `sum += x[i]` chain: one thread per head, in different warps
        red[96 + warp] = serial_sum_f32(sc_s + warp * SCS, T);
    if (tid == 0) ktrace_c(p.trace_slot, 8, c0);       // serial sum done
    cp_async_wait<0>();                                // V slice (issued long ago) -- visible after the next barrier
    __syncthreads();
    if (tid == 0) { ktrace(p.trace_slot, 4); ktrace_c(p.trace_slot, 9, c0); }   // V visible
    if (!(skip & 32))
        rowwise_ilp2(sc_s, nhl, SCS, T, tid, NTHR, [&](float x, int h) { return __fdiv_rn(x, red[96 + h]); });
    __syncthreads();

    if (tid == 0) ktrace_c(p.trace_slot, 10, c0);      // divide done
    // ---- out[h][d] = sum_t a[h][t] * v[t][d], serial over t (:533-542): thread = (head of my group, owned dim), one
    // chain each (serial_av_f32: loads two blocks ahead, products one block ahead of the dependent adds)
    if (tid < nhl * DS && !(skip & 8)) {
        const int hl = tid / DS, d = tid - hl * DS;
        const float o = serial_av_f32<DS>(sc_s + hl * SCS, vt + d, T);
        const size_t oi = (size_t)(h0 + grp * nhl + hl) * HS + part * DS + d;
        if (ll) ll_store(reinterpret_cast<llword_t*>(p.out) + oi, o, seq);
        else reinterpret_cast<float*>(p.out)[oi] = o;
    }
    if (own_pos && chunk == 0 && tid < HS / 2) {       // the rotated K row of this step enters the cache
        p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + tid] = kn0;
        p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + tid + HS / 2] = kn1;
    }
    if (tid == 0) { ktrace(p.trace_slot, 3); ktrace_c(p.trace_slot, 11, c0); }
}

// ---- embedding row gather: the reference dequantizes the whole table at load (src/transformer.rs:243-245,
// src/quantization.rs:25-42) and copies a row per token (:324, :659-669); here a row is dequantized on the fly
// (value = code as f32 * scale: one multiply, bit-identical).  Gemma scales by sqrt(dim) (:327-332).
struct EmbedParams {
    const uint8_t* q; const float* s; const float* f32_table;   // q: BP16 table (s unused)
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
        else if (p.q_type == 1) v = bp_value<1>(p.q, e);
        else v = bp_value<2>(p.q, e);
        if (p.apply_scale) v = __fmul_rn(v, p.scale_mul);
        out[i] = v;
    }
}

}  // namespace lmrs

// =====================================================================================================================
// Score kernel for batched prefill (and LMRS_B200_ATT_SPLIT=1 decode): the per-position dot products are independent, so
// `attn_scores_kernel` spreads them over the whole GPU (grid = kv heads x position splits [x token rows]): thread =
// cached position, four query heads per thread (ILP 4), the K row read straight from L2 into registers -- no tile
// staging, no block barriers in the loop.  The two serial chains of the reference (softmax sum over t, a*v accumulation
// over t) then run in attn_decode_kernel with `scores_ready`.  Arithmetic and operation order are identical to
// attn_decode_body: bit-identical results.
// =====================================================================================================================
namespace lmrs {

constexpr int ATTS_THREADS = 64;    // scores kernel: positions per CTA pass

template <int HS>
__global__ void __launch_bounds__(ATTS_THREADS) attn_scores_kernel(const AttnParams p, const int nsplit) {
    constexpr int C4 = HS / 4;
    __shared__ __align__(16) float q_s[ATT_QH][HS];
    __shared__ __align__(16) float k_s[HS];
    const int tid = threadIdx.x;
    const int kvh = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks, split = blockIdx.y;
    const int brow = p.batch ? (int)blockIdx.z : 0;
    const int h0 = kvh * p.kv_mul + chunk * ATT_QH;
    const int nh = min(ATT_QH, p.kv_mul - chunk * ATT_QH);
    pdl_launch_dependents();
    pdl_wait();
    const int pos = (int)p.step->pos + brow;
    const uint32_t mask_base = p.step->mask_base;
    const int T = pos + 1;
    const int per = (T + nsplit - 1) / nsplit;
    const int t0 = split * per, t1 = min(T, t0 + per);
    const bool have_knew = p.k_new != nullptr;
    const bool owns_pos = have_knew && pos >= t0 && pos < t1;
    if (t0 >= t1) return;
    const float* q_in = reinterpret_cast<const float*>(p.q) + (size_t)brow * p.q_stride;   // (plain mode only)
    const float* cs = p.rope_cos + (size_t)pos * (HS / 2);
    const float* sn = p.rope_sin + (size_t)pos * (HS / 2);
    for (int i = tid; i < ATT_QH * (HS / 2); i += ATTS_THREADS) {
        const int h = i / (HS / 2), j = i - h * (HS / 2);
        float r0 = 0.0f, r1 = 0.0f;
        if (h < nh) {
            const float v0 = __ldcg(q_in + (size_t)(h0 + h) * HS + j), v1 = __ldcg(q_in + (size_t)(h0 + h) * HS + j + HS / 2);
            if (p.batch) { r0 = v0; r1 = v1; }   // rotated by rope_rows_kernel already
            else { const float fcr = cs[j], fci = sn[j]; r0 = __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci)); r1 = __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr)); }
        }
        q_s[h][j] = r0; q_s[h][j + HS / 2] = r1;
    }
    if (owns_pos) {
        for (int j = tid; j < HS / 2; j += ATTS_THREADS) {
            const float fcr = cs[j], fci = sn[j];
            const float* kn = reinterpret_cast<const float*>(p.k_new);
            const float v0 = __ldcg(kn + (size_t)kvh * HS + j), v1 = __ldcg(kn + (size_t)kvh * HS + j + HS / 2);
            const float r0 = __fsub_rn(__fmul_rn(v0, fcr), __fmul_rn(v1, fci));
            const float r1 = __fadd_rn(__fmul_rn(v0, fci), __fmul_rn(v1, fcr));
            k_s[j] = r0; k_s[j + HS / 2] = r1;
            if (chunk == 0) {   // exactly one CTA per KV head publishes the rotated row into the cache
                p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + j] = r0;
                p.kcache[(size_t)pos * p.kv_dim + (size_t)kvh * HS + j + HS / 2] = r1;
            }
        }
    }
    __syncthreads();
    float* sc_out = p.scores + ((size_t)brow * p.kv_mul * (gridDim.x / p.chunks) + h0) * p.seq_len;   // [row][head][stride]
    for (int t = t0 + tid; t < t1; t += ATTS_THREADS) {
        float4 kr[C4];
        if (have_knew && t == pos) {
#pragma unroll
            for (int c = 0; c < C4; c++) kr[c] = reinterpret_cast<const float4*>(k_s)[c];
        } else {
            const float4* krow = reinterpret_cast<const float4*>(p.kcache + (size_t)t * p.kv_dim + (size_t)kvh * HS);
#pragma unroll
            for (int c = 0; c < C4; c++) kr[c] = __ldcg(krow + c);
        }
        float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
        for (int c = 0; c < C4; c++) {   // four independent dot-product chains, ascending d
            const float4 kv = kr[c];
            const float4 a = reinterpret_cast<const float4*>(q_s[0])[c], b = reinterpret_cast<const float4*>(q_s[1])[c];
            const float4 e = reinterpret_cast<const float4*>(q_s[2])[c], f = reinterpret_cast<const float4*>(q_s[3])[c];
            s0 = __fadd_rn(s0, __fmul_rn(a.x, kv.x)); s1 = __fadd_rn(s1, __fmul_rn(b.x, kv.x)); s2 = __fadd_rn(s2, __fmul_rn(e.x, kv.x)); s3 = __fadd_rn(s3, __fmul_rn(f.x, kv.x));
            s0 = __fadd_rn(s0, __fmul_rn(a.y, kv.y)); s1 = __fadd_rn(s1, __fmul_rn(b.y, kv.y)); s2 = __fadd_rn(s2, __fmul_rn(e.y, kv.y)); s3 = __fadd_rn(s3, __fmul_rn(f.y, kv.y));
            s0 = __fadd_rn(s0, __fmul_rn(a.z, kv.z)); s1 = __fadd_rn(s1, __fmul_rn(b.z, kv.z)); s2 = __fadd_rn(s2, __fmul_rn(e.z, kv.z)); s3 = __fadd_rn(s3, __fmul_rn(f.z, kv.z));
            s0 = __fadd_rn(s0, __fmul_rn(a.w, kv.w)); s1 = __fadd_rn(s1, __fmul_rn(b.w, kv.w)); s2 = __fadd_rn(s2, __fmul_rn(e.w, kv.w)); s3 = __fadd_rn(s3, __fmul_rn(f.w, kv.w));
        }
        const float sv[4] = {s0, s1, s2, s3};
#pragma unroll
        for (int h = 0; h < ATT_QH; h++) {
            if (h < nh) {
                float score = __fdiv_rn(sv[h], p.sqrt_hs);
                if (p.gemma) {   // soft-cap 50*tanh(s/50) in f64, window mask on every layer (:518-526)
                    score = __fdiv_rn(score, 50.0f);
                    score = (float)tanh((double)score);
                    score = __fmul_rn(score, 50.0f);
                    score = __fadd_rn(score, (mask_base - (uint32_t)t <= 4096u) ? 0.0f : -2.3819763e38f);
                }
                sc_out[(size_t)h * p.seq_len + t] = score;
            }
        }
    }
}

}  // namespace lmrs
