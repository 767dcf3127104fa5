// gemv.cuh -- decode-path (seq = 1) quantized matrix-vector kernel for sm_100a.
//
// Replaces functional.rs::matmul_q8 / matmul_q4 (src/functional.rs:173-250) for one activation row, with the
// surrounding glue of transformer.rs::forward_layer fused in:
//   prologue  PRO_NORM   residual add (+ Gemma post-norm) + rmsnorm + activation quantize
//                        (src/transformer.rs:409-411,427 / :562-580,596 / :642-656 / :343,356)
//             PRO_QUANT  activation quantize of an f32 vector (src/transformer.rs:553,633; src/quantization.rs:44-95)
//             PRO_RAW    pre-quantized activation given by the caller (operator-level C ABI)
//   epilogue  EPI_STORE / EPI_QKV (q buffer, new-K staging row, V cache row) / EPI_GLU_* (act(gate)*up,
//             src/transformer.rs:607-624) / EPI_LOGITS (Gemma soft-cap quirk, src/transformer.rs:375-381)
//
// Data movement (HBM-bound kernel, ~2 int-ops/byte): every warp owns two independent "streams" of
// consecutive weight rows (one per half-warp; for GLU the halves stream the same rows of w1 and w3).  A
// stream is a contiguous byte range of the row-major [o][n] int8 matrix plus the matching contiguous range
// of f32 group scales, so lane 0 moves it with 1-D bulk async copies (cp.async.bulk -> SASS UBLKCP, the TMA
// engine) into a per-warp shared-memory ring guarded by mbarriers: no registers are tied up by loads in
// flight, and the first DEPTH stages are issued BEFORE griddepcontrol.wait so weight traffic overlaps the
// previous kernel's tail and this kernel's own prologue (programmatic dependent launch).
//
// Arithmetic: a stage gives each lane one whole 128-element group (8 x LDS.128 of weights, 8 x LDS.128 of
// the quantized activation, 32 x IDP.4A) -- the int32 group sum needs no cross-lane reduction.  The f32 part
// follows the reference bit for bit: term = ((ival as f32) * ws) * xs, and terms are added to the row
// accumulator in ascending group order starting from 0.0 (a 16-step shuffle scan per half-warp), so
// matmul_q8/matmul_q4 results are BIT-IDENTICAL to the CPU path; -fmad=false keeps mul/add unfused.
//
// The kernel is specialised on (prologue, epilogue, exchange mode): a launch only carries the code it runs (round 1
// shipped ONE kernel with three prologues x five epilogues whose top stall was instruction fetch).  Exchange mode LL:
// the activations between the kernels of a decode step travel as (value, sequence) words polled by the consumer
// (common.cuh), the kernels are launched early (programmatic dependent launch) and never call griddepcontrol.wait, so
// consecutive kernels are co-resident (<= 128 registers, two CTAs per SM) and a hand-over costs a poll, not a kernel
// boundary.  Plain mode (operator ABI, batched prefill rows, NCCL-sharded steps): f32 arrays + griddepcontrol.wait.
#pragma once
#include "common.cuh"
#include "exact_math.cuh"

namespace lmrs {

constexpr int GS = 128;          // quantization group size (elements); the exporter always writes 128 (utils/io.py:21)
constexpr int SG = 16;           // groups per half-warp stream per stage
constexpr int NORM_MAX_DIM = 4096;  // PRO_NORM keeps the vector in registers: dim <= 4096

enum { PRO_NORM = 0, PRO_QUANT = 1, PRO_RAW = 2 };
enum { EPI_STORE = 0, EPI_QKV = 1, EPI_GLU_SILU = 2, EPI_GLU_GELU = 3, EPI_LOGITS = 4 };

struct StepParams {
    uint32_t token;      // decode: token id; serial prefill: row index into the staged embeddings
    uint32_t pos;        // position of this step
    uint32_t mask_base;  // Gemma window quirk: the reference tests `pos - t` with the BATCH start pos
    uint32_t seq;        //   (src/transformer.rs:525); decode passes pos.  seq: launch number (grid-barrier base)
};

struct GemvParams {
    const uint8_t* wq_a; const float* ws_a;   // matrix A: block-packed (BP16) weights; ws_* unused (scales ride in the blocks)
    const uint8_t* wq_b; const float* ws_b;   // matrix B (GLU: w3), else unused
    int n, o, row_gran;
    int pro, epi;      // host-side dispatch (the kernels are specialised on both)
    int ll;            // host-side dispatch: activation pointers below are LL word arrays (llword_t) instead of f32 arrays
    int ll_nowait;     // LL measurement passes: do not wait for the sequence number
    int pre_stages;    // plain mode: ring stages requested BEFORE griddepcontrol.wait (the rest right after it)
    // activations in/out: `const float*` in plain mode, `const llword_t*` in LL mode (same element indexing)
    const void* x_in; const void* delta; const float* w_post; const float* w_norm; void* x_out;
    int x_in_plain;    // LL mode: x_in is nevertheless a plain f32 array (serial prefill: the staged embedding rows)
    int x_in_stride;   // serial prefill: x_in += step->token * x_in_stride (row of the staged embeddings)
    int xout_all;      // batched prefill (one CTA per row): every CTA writes its x_out
    float eps; int unit_offset;
    // PRO_NORM may take x_in from the embedding table instead (decode step, src/transformer.rs:324-332):
    const uint8_t* emb_q; const float* emb_s; int emb_qtype; float emb_mul; int emb_apply_mul;   // emb_q: BP16 table
    const void* act_in;
    const uint8_t* raw_q; const float* raw_s;
    void* out; void* out_k; void* out_v;   // EPI_QKV: q / new K row / V (plain: cache base, row `pos`; LL: staging row); EPI_LOGITS: always f32
    int att_dim, kv_dim;
    const StepParams* step;
    int softcap_rows;
    int trace_slot;    // LMRS_TRACE builds: timeline slot of this launch (-1: none)
    // L2 prefetch of weights LATER kernels of the step will stream, requested while this kernel's CTAs sit in the dependency
    // wait (common.cuh l2_prefetch_slice; chunk < 0: evict_first)
    const uint8_t* l2pf_ptr; unsigned long long l2pf_bytes; int l2pf_chunk;
    // row-sharded N-GPU mode, peer exchange (common.cuh): px_world > 1
    int px_world;
    const llword_t* px_in;               // PRO_NORM: this GPU's exchange vector [px_world][n]; delta = slot 0 + slot 1 + ... in rank order
    llword_t* px_out[PX_MAX_WORLD];      // EPI_STORE: this rank's slot in EVERY GPU's exchange vector (peer-mapped), indexed by row
    float* px_logits[PX_MAX_WORLD];      // EPI_LOGITS: every GPU's logits buffer + this rank's vocabulary offset
};

template <int QT> struct QTraits;
template <> struct QTraits<1> { static constexpr int QB = 128; };  // bytes of codes per group, Q8_0
template <> struct QTraits<2> { static constexpr int QB = 64; };   // Q4_0

// Weights live in HBM in a block-packed layout ("BP16"): the matrix's quantization groups in row-major order, 16 per
// block, each block = 16 x QB code bytes followed by the 16 f32 group scales (2112 B for Q8_0, 1088 B for Q4_0).  A
// half-warp stage is exactly one block, i.e. ONE bulk copy that brings codes and scales together (the file keeps
// them in two separate arrays; lmrs_b200.cu repacks at load time).  The last block is zero-padded.
template <int QT> __host__ __device__ constexpr int blk_bytes() { return SG * QTraits<QT>::QB + SG * 4; }
template <int QT> __host__ __device__ constexpr int gemv_stage_bytes() { return 2 * blk_bytes<QT>(); }
// element e of a block-packed Q8/Q4 tensor (embedding gather): value = code as f32 * scale
template <int QT> LMRS_DEVINL float bp_value(const uint8_t* base, size_t e) {
    constexpr int QB = QTraits<QT>::QB;
    const size_t f = e / GS; const int k = (int)(e % GS);
    const uint8_t* blk = base + (f / SG) * blk_bytes<QT>();
    const int gi = (int)(f % SG);
    const float sc = reinterpret_cast<const float*>(blk + SG * QB)[gi];
    int code;
    if (QT == 1) code = reinterpret_cast<const int8_t*>(blk)[gi * QB + k];
    else { const int b = blk[gi * QB + (k >> 1)]; code = ((k & 1) ? (b >> 4) : (b & 15)) - 8; }
    return __fmul_rn((float)code, sc);
}
// file layout ([o][n] codes + [o][n/128] scales) -> BP16
template <int QT> __global__ void repack_bp16_kernel(uint8_t* dst, const uint8_t* src_q, const float* src_s, size_t n_groups) {
    constexpr int QB = QTraits<QT>::QB;
    const size_t b = blockIdx.x;
    uint8_t* out = dst + b * blk_bytes<QT>();
    for (int i = threadIdx.x; i < SG * QB / 16; i += blockDim.x) {   // 16-byte chunks of codes
        const size_t f = b * SG + (size_t)(i * 16) / QB;
        int4 v = make_int4(0, 0, 0, 0);
        if (f < n_groups) v = *reinterpret_cast<const int4*>(src_q + b * SG * QB + (size_t)i * 16);
        reinterpret_cast<int4*>(out)[i] = v;
    }
    if (threadIdx.x < SG) {
        const size_t f = b * SG + threadIdx.x;
        reinterpret_cast<float*>(out + SG * QB)[threadIdx.x] = f < n_groups ? src_s[f] : 0.0f;
    }
}

// shared-memory footprint of one CTA (must match the carve-up in the kernel)
template <int QT, int WARPS, int DEPTH> inline size_t gemv_smem_bytes(int n, bool norm) {
    size_t ring = (size_t)WARPS * DEPTH * gemv_stage_bytes<QT>();
    size_t xq = (size_t)n;                         // Q8: n codes; Q4: n/2 even + n/2 odd signed bytes
    size_t xs = (size_t)(n / GS) * 4 * 2;          // scales + per-group code sums (Q4)
    size_t xf = norm ? (size_t)n * 4 + 128 : 0;    // PRO_NORM: staging of the squares, eight padded chain rows (exact_rnorm_t)
    return ring + ((xq + 127) / 128) * 128 + ((xs + 127) / 128) * 128 + 64 * 4 + (size_t)WARPS * DEPTH * 8 + 128 + xf + 256;
}

struct RowRange { int row0, nrows; };
LMRS_DEVINL RowRange slot_rows(int slot, int nslots, int o, int gran) {
    int units = (o + gran - 1) / gran;               // the last unit may be short
    int a = (int)(((long long)slot * units) / nslots);
    int b = (int)(((long long)(slot + 1) * units) / nslots);
    int r0 = a * gran, r1 = min(b * gran, o);
    return {r0, max(r1 - r0, 0)};
}

template <int WARPS> LMRS_DEVINL float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = lane < WARPS ? red[lane] : 0.0f;
    t = warp_sum(t);
    __syncthreads();
    return t;
}

// 1/sqrt(mean(x^2)+eps) with the reference's exact summation order (src/functional.rs:49-62): eight lane
// accumulators ss[k] += x[8j+k]*x[8j+k] walked serially over j (mul and add unfused), then wide's
// f32x8::reduce_add order ((a0+a4)+(a2+a6))+((a1+a5)+(a3+a7)) (see oracle/lmrs_ref.c header), /size, +eps,
// 1/sqrt.  Eight threads walk the chains (size/8 dependent adds, ~4 cycles each); everybody gets the result.
LMRS_DEVINL float exact_rnorm(const float* xf, int n, float eps, float* red) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float s = 0.0f;
        if (lane < 8) {
            // software pipeline, two batches deep: while batch k's eight dependent adds run (the critical path,
            // ~4 cycles each), batch k+1's products are already in registers and batch k+2's loads are in flight
            const int steps = n / 8;
            const int nb = steps / 8;
            float pa[8], pb[8], xc[8];
            if (nb >= 2) {
#pragma unroll
                for (int u = 0; u < 8; u++) { const float x = xf[8 * u + lane]; pa[u] = __fmul_rn(x, x); }
#pragma unroll
                for (int u = 0; u < 8; u++) xc[u] = xf[8 * (8 + u) + lane];
                for (int b = 0; b + 2 < nb; b++) {
#pragma unroll
                    for (int u = 0; u < 8; u++) pb[u] = __fmul_rn(xc[u], xc[u]);
#pragma unroll
                    for (int u = 0; u < 8; u++) xc[u] = xf[8 * ((b + 2) * 8 + u) + lane];
#pragma unroll
                    for (int u = 0; u < 8; u++) s = __fadd_rn(s, pa[u]);
#pragma unroll
                    for (int u = 0; u < 8; u++) pa[u] = pb[u];
                }
#pragma unroll
                for (int u = 0; u < 8; u++) pb[u] = __fmul_rn(xc[u], xc[u]);
#pragma unroll
                for (int u = 0; u < 8; u++) s = __fadd_rn(s, pa[u]);
#pragma unroll
                for (int u = 0; u < 8; u++) s = __fadd_rn(s, pb[u]);
                for (int j = nb * 8; j < steps; j++) { const float x = xf[8 * j + lane]; s = __fadd_rn(s, __fmul_rn(x, x)); }
            } else {
                for (int j = 0; j < steps; j++) { const float x = xf[8 * j + lane]; s = __fadd_rn(s, __fmul_rn(x, x)); }
            }
        }
        const float t = __fadd_rn(s, __shfl_sync(0xffffffffu, s, (lane + 4) & 31));   // lanes 0..3: a_l + a_{l+4}
        const float u = __fadd_rn(t, __shfl_sync(0xffffffffu, t, (lane + 2) & 31));   // lane 0: s0+s2, lane 1: s1+s3
        float ss = __fadd_rn(u, __shfl_sync(0xffffffffu, u, (lane + 1) & 31));        // lane 0: (s0+s2)+(s1+s3)
        if (lane == 0) {
            ss = __fdiv_rn(ss, (float)n);
            ss = __fadd_rn(ss, eps);
            red[0] = __fdiv_rn(1.0f, __fsqrt_rn(ss));
        }
    }
    __syncthreads();
    const float r = red[0];
    __syncthreads();
    return r;
}

// quantize 4 consecutive activations held by each lane of a full warp == one 128-group
// Q8: src/quantization.rs:44-67.  Q4: src/quantization.rs:69-95, stored as signed (nibble-8) bytes split
// into even/odd element planes + the per-group sum of those bytes (used to fold the weight-side "-8").
template <int QT>
LMRS_DEVINL void quantize_group_to_smem(float4 y, int g, uint8_t* xq, float* xs, int* xsum, int n) {
    const int lane = threadIdx.x & 31;
    float m = fmaxf(fmaxf(fabsf(y.x), fabsf(y.y)), fmaxf(fabsf(y.z), fabsf(y.w)));
    m = warp_max(m);
    if (QT == 1) {
        float scale = __fdiv_rn(m, 127.0f);
        int a = round_sat_i8(__fdiv_rn(y.x, scale)), b = round_sat_i8(__fdiv_rn(y.y, scale));
        int c = round_sat_i8(__fdiv_rn(y.z, scale)), d = round_sat_i8(__fdiv_rn(y.w, scale));
        uint32_t packed = (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) |
                          ((uint32_t)(d & 0xff) << 24);
        reinterpret_cast<uint32_t*>(xq + (size_t)g * GS)[lane] = packed;
        if (lane == 0) xs[g] = scale;
    } else {
        float scale = __fdiv_rn(m, -8.0f);
        int a = round_sat_u4(__fdiv_rn(y.x, scale)) - 8, b = round_sat_u4(__fdiv_rn(y.y, scale)) - 8;
        int c = round_sat_u4(__fdiv_rn(y.z, scale)) - 8, d = round_sat_u4(__fdiv_rn(y.w, scale)) - 8;
        uint8_t* xe = xq + (size_t)g * (GS / 2);
        uint8_t* xo = xq + (size_t)(n / 2) + (size_t)g * (GS / 2);
        reinterpret_cast<uint16_t*>(xe)[lane] = (uint16_t)((a & 0xff) | ((c & 0xff) << 8));
        reinterpret_cast<uint16_t*>(xo)[lane] = (uint16_t)((b & 0xff) | ((d & 0xff) << 8));
        int sum = a + b + c + d;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) { xs[g] = scale; xsum[g] = sum; }
    }
}


// The same chain on a TRANSPOSED staging of the squares (what the GEMV prologues use): the owner of element e = 8j + k
// stores x[e]*x[e] (the reference's unfused product, src/functional.rs:56) at row k, column j of eight rows padded to
// n/8 + 4 floats, so chain k reads ITS products as consecutive float4 (one LDS.128 per four dependent adds, the eight
// lanes on distinct banks) instead of one strided scalar load per add: the chain runs at the dependent-add latency.
LMRS_DEVINL int rnorm_t_stride(int n) { return n / 8 + 4; }
LMRS_DEVINL void rnorm_t_stage4(float* xt, int n, int c, float4 v) {   // chunk c = elements 4c .. 4c+3
    const int st = rnorm_t_stride(n), k0 = 4 * (c & 1), j = c >> 1;
    xt[(k0 + 0) * st + j] = __fmul_rn(v.x, v.x); xt[(k0 + 1) * st + j] = __fmul_rn(v.y, v.y);
    xt[(k0 + 2) * st + j] = __fmul_rn(v.z, v.z); xt[(k0 + 3) * st + j] = __fmul_rn(v.w, v.w);
}
LMRS_DEVINL float exact_rnorm_t(const float* xt, int n, float eps, float* red) {   // n % 128 == 0
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float s = 0.0f;
        if (lane < 8) {
            const float4* row = reinterpret_cast<const float4*>(xt + lane * rnorm_t_stride(n));
            const int nq = n / 32;   // float4 per chain; a multiple of 4
            float4 a0 = row[0], a1 = row[1], a2 = row[2], a3 = row[3];
            for (int q = 4; q < nq; q += 4) {   // loads run four float4 (16 adds = 64 cycles) ahead of the chain
                const float4 b0 = row[q], b1 = row[q + 1], b2 = row[q + 2], b3 = row[q + 3];
                s = __fadd_rn(s, a0.x); s = __fadd_rn(s, a0.y); s = __fadd_rn(s, a0.z); s = __fadd_rn(s, a0.w);
                s = __fadd_rn(s, a1.x); s = __fadd_rn(s, a1.y); s = __fadd_rn(s, a1.z); s = __fadd_rn(s, a1.w);
                s = __fadd_rn(s, a2.x); s = __fadd_rn(s, a2.y); s = __fadd_rn(s, a2.z); s = __fadd_rn(s, a2.w);
                s = __fadd_rn(s, a3.x); s = __fadd_rn(s, a3.y); s = __fadd_rn(s, a3.z); s = __fadd_rn(s, a3.w);
                a0 = b0; a1 = b1; a2 = b2; a3 = b3;
            }
            s = __fadd_rn(s, a0.x); s = __fadd_rn(s, a0.y); s = __fadd_rn(s, a0.z); s = __fadd_rn(s, a0.w);
            s = __fadd_rn(s, a1.x); s = __fadd_rn(s, a1.y); s = __fadd_rn(s, a1.z); s = __fadd_rn(s, a1.w);
            s = __fadd_rn(s, a2.x); s = __fadd_rn(s, a2.y); s = __fadd_rn(s, a2.z); s = __fadd_rn(s, a2.w);
            s = __fadd_rn(s, a3.x); s = __fadd_rn(s, a3.y); s = __fadd_rn(s, a3.z); s = __fadd_rn(s, a3.w);
        }
        const float t = __fadd_rn(s, __shfl_sync(0xffffffffu, s, (lane + 4) & 31));   // lanes 0..3: a_l + a_{l+4}
        const float u = __fadd_rn(t, __shfl_sync(0xffffffffu, t, (lane + 2) & 31));   // lane 0: s0+s2, lane 1: s1+s3
        float ss = __fadd_rn(u, __shfl_sync(0xffffffffu, u, (lane + 1) & 31));        // lane 0: (s0+s2)+(s1+s3)
        if (lane == 0) {
            ss = __fdiv_rn(ss, (float)n);
            ss = __fadd_rn(ss, eps);
            red[0] = __fdiv_rn(1.0f, __fsqrt_rn(ss));
        }
    }
    __syncthreads();
    const float r = red[0];
    __syncthreads();
    return r;
}

// ---- shared-memory views of one CTA -----------------------------------------------------------------------------
struct GemvSmem {
    uint8_t* xq;    // quantized activation (Q8: n codes; Q4: n/2 even + n/2 odd signed bytes)
    float* xs;      // [G] activation scales
    int* xsum;      // [G] per-group sums of the signed activation bytes (Q4)
    float* red;     // [64] reduction scratch
    float* xf;      // [n] f32 staging for the exact rmsnorm chains (PRO_NORM)
    const uint64_t* exp_tab;   // shared-memory copy of the expf table (GLU epilogue), or the global one
};

// ---- one warp's two weight streams for one matrix -------------------------------------------------------------------
template <int QT> struct WarpStreams {
    RowRange r0, r1;
    int ng0, ng1, nst, G;
    const uint8_t *src0, *src1;      // first block of each half's stream
    bool glu;
};
// the part of GemvParams that defines the weight streams (kept in shared memory by the megakernel)
struct StreamDesc {
    const uint8_t* wq_a; const float* ws_a; const uint8_t* wq_b; const float* ws_b;
    int n, o, row_gran, epi;
};
LMRS_DEVINL StreamDesc stream_desc(const GemvParams& p) { return {p.wq_a, p.ws_a, p.wq_b, p.ws_b, p.n, p.o, p.row_gran, p.epi}; }
template <int QT>
LMRS_DEVINL WarpStreams<QT> make_streams(const StreamDesc& p, int wslot, int n_wslots) {
    WarpStreams<QT> w;
    w.glu = (p.epi == EPI_GLU_SILU || p.epi == EPI_GLU_GELU);
    w.G = p.n / GS;
    const int nslots = n_wslots * (w.glu ? 1 : 2);
    w.r0 = slot_rows(w.glu ? wslot : wslot * 2, nslots, p.o, p.row_gran);
    w.r1 = w.glu ? w.r0 : slot_rows(wslot * 2 + 1, nslots, p.o, p.row_gran);
    w.ng0 = w.r0.nrows * w.G; w.ng1 = w.r1.nrows * w.G;
    w.nst = max((w.ng0 + SG - 1) / SG, (w.ng1 + SG - 1) / SG);
    w.src0 = p.wq_a + ((size_t)w.r0.row0 * w.G / SG) * blk_bytes<QT>();      // row0 * G is a multiple of 16 (row_gran)
    w.src1 = (w.glu ? p.wq_b : p.wq_a) + ((size_t)w.r1.row0 * w.G / SG) * blk_bytes<QT>();
    return w;
}
// lane 0: stream stage s of both halves into one ring slot: one block-sized bulk copy per half, one mbarrier.
// Weights are read exactly once per token: L2 evict_first keeps the KV cache and the activations resident instead.
template <int QT>
LMRS_DEVINL void issue_stage(const WarpStreams<QT>& w, int s, uint8_t* buf, uint64_t* bar, uint64_t pol) {
    constexpr int BLK = blk_bytes<QT>();
    const bool h0 = w.ng0 > SG * s, h1 = w.ng1 > SG * s;
    mbar_expect_tx(bar, (uint32_t)((h0 ? BLK : 0) + (h1 ? BLK : 0)));
    if (h0) bulk_g2s_hint(buf, w.src0 + (size_t)s * BLK, BLK, bar, pol);
    if (h1) bulk_g2s_hint(buf + BLK, w.src1 + (size_t)s * BLK, BLK, bar, pol);
}

// ---- activation access, plain f32 arrays or LL word arrays (common.cuh) ----------------------------------------------
// gather this thread's float4 chunks c = tid + k*THREADS (< nchunks) of an LL vector: all loads of a round are issued
// together and only the chunks that were not complete yet are polled again
template <int NC, int THREADS>
LMRS_DEVINL void ll_gather(const llword_t* base, int nchunks, uint32_t seq, bool nowait, float4 (&v)[NC]) {
    const int tid = threadIdx.x;
    bool done[NC];
#pragma unroll
    for (int k = 0; k < NC; k++) { done[k] = !(tid + k * THREADS < nchunks); v[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
    const LLSpin sp = ll_spin_begin();
    for (;;) {
        bool all = true;
#pragma unroll
        for (int k = 0; k < NC; k++)
            if (!done[k]) {
                if (ll_try4(base + 4 * (size_t)(tid + k * THREADS), seq, nowait, v[k])) done[k] = true;
                else all = false;
            }
        if (all) break;
        __nanosleep(20);
        ll_spin_check(sp);
    }
}
// two vectors at once (residual stream + pending contribution): one L2 round trip instead of two on the critical path
template <int NC, int THREADS>
LMRS_DEVINL void ll_gather2(const llword_t* base_a, const llword_t* base_b, int nchunks, uint32_t seq, bool nowait, float4 (&va)[NC], float4 (&vb)[NC]) {
    const int tid = threadIdx.x;
    bool da[NC], db[NC];
#pragma unroll
    for (int k = 0; k < NC; k++) {
        da[k] = db[k] = !(tid + k * THREADS < nchunks);
        va[k] = vb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const LLSpin sp = ll_spin_begin();
    for (;;) {
        bool all = true;
#pragma unroll
        for (int k = 0; k < NC; k++) {
            if (!da[k]) { if (ll_try4(base_a + 4 * (size_t)(tid + k * THREADS), seq, nowait, va[k])) da[k] = true; else all = false; }
            if (!db[k]) { if (ll_try4(base_b + 4 * (size_t)(tid + k * THREADS), seq, nowait, vb[k])) db[k] = true; else all = false; }
        }
        if (all) break;
        __nanosleep(20);
        ll_spin_check(sp);
    }
}
// peer exchange: this thread's chunks of the N partial vectors the GPUs pushed into this GPU's exchange buffer, added in
// ascending rank order starting from rank 0's (the order the oracle's k-shard mode restates).  Up to four slots are
// requested together (one L2 round trip per batch), only incomplete chunks are polled again.
template <int NC, int THREADS>
LMRS_DEVINL void px_gather_sum(const llword_t* base, int world, int n, int nchunks, uint32_t seq, bool nowait, float4 (&dv)[NC]) {
    constexpr int PXB = NC <= 2 ? 4 : 2;   // slots requested together (bounded by the registers of the wide-CTA builds)
    const int tid = threadIdx.x;
    for (int r0 = 0; r0 < world; r0 += PXB) {
        float4 t[PXB][NC];
        bool done[PXB][NC];
#pragma unroll
        for (int b = 0; b < PXB; b++)
#pragma unroll
            for (int k = 0; k < NC; k++) { done[b][k] = !(r0 + b < world && tid + k * THREADS < nchunks); t[b][k] = make_float4(0.f, 0.f, 0.f, 0.f); }
        const LLSpin sp = ll_spin_begin();
        for (;;) {
            bool all = true;
#pragma unroll
            for (int b = 0; b < PXB; b++)
#pragma unroll
                for (int k = 0; k < NC; k++)
                    if (!done[b][k]) {
                        if (ll_try4_sys(base + (size_t)(r0 + b) * n + 4 * (size_t)(tid + k * THREADS), seq, nowait, t[b][k])) done[b][k] = true;
                        else all = false;
                    }
            if (all) break;
            __nanosleep(40);
            px_spin_check(sp);
        }
#pragma unroll
        for (int b = 0; b < PXB; b++)
            if (r0 + b < world) {
#pragma unroll
                for (int k = 0; k < NC; k++) {
                    if (r0 + b == 0) dv[k] = t[b][k];
                    else { dv[k].x = __fadd_rn(dv[k].x, t[b][k].x); dv[k].y = __fadd_rn(dv[k].y, t[b][k].y); dv[k].z = __fadd_rn(dv[k].z, t[b][k].z); dv[k].w = __fadd_rn(dv[k].w, t[b][k].w); }
                }
            }
    }
}
LMRS_DEVINL void ll_store4(llword_t* p, float4 v, uint32_t seq) {
    const unsigned long long hi = (unsigned long long)seq << 32;
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(hi | __float_as_uint(v.x)), "l"(hi | __float_as_uint(v.y)) : "memory");
    asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p + 2), "l"(hi | __float_as_uint(v.z)), "l"(hi | __float_as_uint(v.w)) : "memory");
}
template <bool LL> LMRS_DEVINL void act_store(void* base, size_t i, float v, uint32_t seq) {
    if constexpr (LL) ll_store(reinterpret_cast<llword_t*>(base) + i, v, seq);
    else reinterpret_cast<float*>(base)[i] = v;
}

// ---- prologue: build the quantized activation in shared memory (whole CTA, ends with __syncthreads) ----------------
template <int QT, int WARPS, int PRO, bool LL, int MAXC = NORM_MAX_DIM / 4 / (WARPS * 32)>
LMRS_DEVINL void gemv_prologue(const GemvParams& p, const GemvSmem& sm, const uint32_t seq) {
    constexpr int THREADS = WARPS * 32;
    constexpr int NORM_MAXC = MAXC;   // float4 chunks per thread (default: enough for dim <= NORM_MAX_DIM)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = p.n, G = n / GS;
    const bool nowait = (LL || p.px_world > 1) && p.ll_nowait;
    const long long cP = ktrace_c0();
    trace_event(100 + PRO);
    if constexpr (PRO == PRO_NORM) {
        const int nchunks = n / 4;
        float4 v[NORM_MAXC], wnv[NORM_MAXC];
        {   // the norm weights do not depend on the previous phase: get them in flight first
            const float4* wn = reinterpret_cast<const float4*>(p.w_norm);
#pragma unroll
            for (int k = 0; k < NORM_MAXC; k++) {
                const int c = tid + k * THREADS;
                wnv[k] = c < nchunks ? wn[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (p.emb_q) {   // embedding row dequantized on the fly: code as f32 * scale (src/quantization.rs:25-42)
            const uint32_t tok = p.step->token;
#pragma unroll
            for (int k = 0; k < NORM_MAXC; k++) {
                const int c = tid + k * THREADS;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c < nchunks) {
                    const size_t e = (size_t)tok * n + (size_t)c * 4;
                    t = make_float4(bp_value<QT>(p.emb_q, e), bp_value<QT>(p.emb_q, e + 1), bp_value<QT>(p.emb_q, e + 2), bp_value<QT>(p.emb_q, e + 3));
                    if (p.emb_apply_mul) { t.x = __fmul_rn(t.x, p.emb_mul); t.y = __fmul_rn(t.y, p.emb_mul); t.z = __fmul_rn(t.z, p.emb_mul); t.w = __fmul_rn(t.w, p.emb_mul); }
                }
                v[k] = t;
            }
        } else if (LL && !p.x_in_plain) {
            // gathered together with delta below (one round trip); a vector without a pending contribution is read alone
            if (!p.delta) ll_gather<NORM_MAXC, THREADS>(reinterpret_cast<const llword_t*>(p.x_in), nchunks, seq, nowait, v);
        } else {
            // N-GPU consumer: this kernel did NOT wait for its predecessor (see the kernel body).  The first words of every
            // GPU's partial prove that the predecessor is past ITS dependency wait, i.e. the kernel that wrote x_in has
            // completed and flushed; only then is x_in read (through L2).
            if (!LL && p.px_world > 1 && p.px_in) px_canary_wait(p.px_in, p.px_world, n, seq, nowait);
            const float4* xin = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.x_in) + (size_t)(p.x_in_stride ? p.step->token : 0u) * p.x_in_stride);
#pragma unroll
            for (int k = 0; k < NORM_MAXC; k++) {
                const int c = tid + k * THREADS;
                v[k] = c < nchunks ? __ldcg(&xin[c]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (p.delta) {
            float4 dv[NORM_MAXC], wpv[NORM_MAXC];
#pragma unroll
            for (int k = 0; k < NORM_MAXC; k++) {
                const int c = tid + k * THREADS;
                wpv[k] = (p.w_post && c < nchunks) ? reinterpret_cast<const float4*>(p.w_post)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if constexpr (LL) {
                if (!p.emb_q && !p.x_in_plain)
                    ll_gather2<NORM_MAXC, THREADS>(reinterpret_cast<const llword_t*>(p.x_in), reinterpret_cast<const llword_t*>(p.delta), nchunks, seq, nowait, v, dv);
                else
                    ll_gather<NORM_MAXC, THREADS>(reinterpret_cast<const llword_t*>(p.delta), nchunks, seq, nowait, dv);
            } else if (p.px_world > 1) {   // N-GPU mode: the contribution is the rank-ordered sum of the partials every GPU pushed here
                if (p.emb_q) px_canary_wait(p.px_in, p.px_world, n, seq, nowait);   // (otherwise parked on the canaries before x_in was read)
                px_gather_sum<NORM_MAXC, THREADS>(p.px_in, p.px_world, n, nchunks, seq, nowait, dv);
            } else {
                const float4* din = reinterpret_cast<const float4*>(p.delta);
#pragma unroll
                for (int k = 0; k < NORM_MAXC; k++) {
                    const int c = tid + k * THREADS;
                    dv[k] = c < nchunks ? __ldcg(&din[c]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (p.w_post) {  // Gemma: x += rmsnorm(delta, w_post) with unit offset (src/transformer.rs:564,645)
#pragma unroll
                for (int k = 0; k < NORM_MAXC; k++) {
                    const int c = tid + k * THREADS;
                    if (c < nchunks) rnorm_t_stage4(sm.xf, n, c, dv[k]);
                }
                __syncthreads();
                const float r = exact_rnorm_t(sm.xf, n, p.eps, sm.red);
#pragma unroll
                for (int k = 0; k < NORM_MAXC; k++) {
                    const int c = tid + k * THREADS;
                    if (c < nchunks) {
                        const float4 w = wpv[k];
                        dv[k].x = __fmul_rn(__fadd_rn(1.0f, w.x), __fmul_rn(r, dv[k].x));
                        dv[k].y = __fmul_rn(__fadd_rn(1.0f, w.y), __fmul_rn(r, dv[k].y));
                        dv[k].z = __fmul_rn(__fadd_rn(1.0f, w.z), __fmul_rn(r, dv[k].z));
                        dv[k].w = __fmul_rn(__fadd_rn(1.0f, w.w), __fmul_rn(r, dv[k].w));
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < NORM_MAXC; k++) {
                v[k].x = __fadd_rn(v[k].x, dv[k].x); v[k].y = __fadd_rn(v[k].y, dv[k].y);
                v[k].z = __fadd_rn(v[k].z, dv[k].z); v[k].w = __fadd_rn(v[k].w, dv[k].w);
            }
        }
        if (p.x_out && (blockIdx.x == 0 || p.xout_all)) {  // decode: exactly one CTA publishes the updated residual stream
#pragma unroll
            for (int k = 0; k < NORM_MAXC; k++) {
                const int c = tid + k * THREADS;
                if (c < nchunks) {
                    if constexpr (LL) ll_store4(reinterpret_cast<llword_t*>(p.x_out) + 4 * (size_t)c, v[k], seq);
                    else reinterpret_cast<float4*>(p.x_out)[c] = v[k];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < NORM_MAXC; k++) {
            const int c = tid + k * THREADS;
            if (c < nchunks) rnorm_t_stage4(sm.xf, n, c, v[k]);
        }
        __syncthreads();
        if (lane == 0) ktrace_c(p.trace_slot, 4, cP);   // inputs loaded, residual formed
        trace_event(110);
        const float r = exact_rnorm_t(sm.xf, n, p.eps, sm.red);   // src/functional.rs:48-62, exact order
        if (lane == 0) ktrace_c(p.trace_slot, 5, cP);   // 1/rms known
        trace_event(111);
#pragma unroll
        for (int k = 0; k < NORM_MAXC; k++) {
            const int c = tid + k * THREADS;       // chunk c = 4 elements; 32 consecutive chunks = one warp = one group
            if (c < nchunks) {
                const float4 w = wnv[k];
                float4 y;
                if (p.unit_offset) {
                    y.x = __fmul_rn(__fadd_rn(1.0f, w.x), __fmul_rn(r, v[k].x));
                    y.y = __fmul_rn(__fadd_rn(1.0f, w.y), __fmul_rn(r, v[k].y));
                    y.z = __fmul_rn(__fadd_rn(1.0f, w.z), __fmul_rn(r, v[k].z));
                    y.w = __fmul_rn(__fadd_rn(1.0f, w.w), __fmul_rn(r, v[k].w));
                } else {
                    y.x = __fmul_rn(w.x, __fmul_rn(r, v[k].x)); y.y = __fmul_rn(w.y, __fmul_rn(r, v[k].y));
                    y.z = __fmul_rn(w.z, __fmul_rn(r, v[k].z)); y.w = __fmul_rn(w.w, __fmul_rn(r, v[k].w));
                }
                quantize_group_to_smem<QT>(y, c >> 5, sm.xq, sm.xs, sm.xsum, n);
            }
        }
    } else if constexpr (PRO == PRO_QUANT) {
        for (int g0 = warp; g0 < G; g0 += WARPS * 8) {   // 8 groups per warp in flight: one L2 round trip, not eight
            float4 y[8];
            if constexpr (LL) {
                const llword_t* ain = reinterpret_cast<const llword_t*>(p.act_in);
                bool done[8];
#pragma unroll
                for (int u = 0; u < 8; u++) { done[u] = !(g0 + u * WARPS < G); y[u] = make_float4(0.f, 0.f, 0.f, 0.f); }
                const LLSpin sp = ll_spin_begin();
                for (;;) {
                    bool all = true;
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if (!done[u]) {
                            if (ll_try4(ain + 4 * ((size_t)(g0 + u * WARPS) * 32 + lane), seq, nowait, y[u])) done[u] = true;
                            else all = false;
                        }
                    if (__all_sync(0xffffffffu, all)) break;
                    __nanosleep(20);
                    ll_spin_check(sp);
                }
            } else {
                const float4* ain = reinterpret_cast<const float4*>(p.act_in);
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int g = g0 + u * WARPS;
                    y[u] = g < G ? __ldcg(&ain[g * 32 + lane]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int g = g0 + u * WARPS;
                if (g < G) quantize_group_to_smem<QT>(y[u], g, sm.xq, sm.xs, sm.xsum, n);
            }
        }
    } else {  // PRO_RAW: caller-supplied codes (Q8: i8[n]; Q4: packed nibbles u8[n/2]) and scales
        if (QT == 1) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(p.raw_q);
            for (int i = tid; i < n / 4; i += THREADS) reinterpret_cast<uint32_t*>(sm.xq)[i] = src[i];
            for (int g = tid; g < G; g += THREADS) sm.xs[g] = p.raw_s[g];
        } else {
            for (int g = warp; g < G; g += WARPS) {   // unpack one group per warp: 64 bytes = 2 per lane
                const uint16_t two = reinterpret_cast<const uint16_t*>(p.raw_q + (size_t)g * 64)[lane];
                const int b0 = two & 0xff, b1 = two >> 8;
                const int e0 = (b0 & 15) - 8, o0 = (b0 >> 4) - 8, e1 = (b1 & 15) - 8, o1 = (b1 >> 4) - 8;
                reinterpret_cast<uint16_t*>(sm.xq + (size_t)g * 64)[lane] = (uint16_t)((e0 & 0xff) | ((e1 & 0xff) << 8));
                reinterpret_cast<uint16_t*>(sm.xq + (size_t)(n / 2) + (size_t)g * 64)[lane] =
                    (uint16_t)((o0 & 0xff) | ((o1 & 0xff) << 8));
                int sum = e0 + o0 + e1 + o1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (lane == 0) { sm.xs[g] = p.raw_s[g]; sm.xsum[g] = sum; }
            }
        }
    }
    __syncthreads();
    if (lane == 0) ktrace_c(p.trace_slot, 6, cP);       // quantized activations in shared memory
    trace_event(119);
}

// ---- one stage of one warp: 32 group dot products, ordered f32 accumulation, epilogue -------------------------------
// Per-phase consumer state.  `aligned` (G % 16 == 0, every real model): a half-warp stage lies inside one row, so the
// row/group bookkeeping is warp-uniform and incremental (no integer division, no per-step selects in the scan).
// `xreg` (Q8, G == 16): lane l always meets activation group l, which then lives in 32 registers for the whole phase.
template <int QT> struct Consumer {
    float acc;
    int row_l, g_base;
    bool aligned, xreg;
    int4 xr[8];
};
template <int QT>
LMRS_DEVINL void consumer_begin(Consumer<QT>& c, const WarpStreams<QT>& w, const GemvSmem& sm) {
    const int l16 = threadIdx.x & 15;
    c.acc = 0.0f; c.row_l = 0; c.g_base = 0;
    c.aligned = (w.G % SG) == 0;
    c.xreg = (QT == 1) && w.G == SG;
    if (c.xreg) {
        const int4* xv = reinterpret_cast<const int4*>(sm.xq + (size_t)l16 * GS);
#pragma unroll
        for (int i = 0; i < 8; i++) c.xr[i] = xv[(i + l16) & 7];
    }
}
LMRS_DEVINL float glu_act(int epi, float val, const uint64_t* exp_tab = kExp2fTab) {
    if (epi == EPI_GLU_GELU) {  // tanh-GELU, tanh in f64 (src/transformer.rs:614)
        const float inner = __fadd_rn(val, __fmul_rn(__fmul_rn(__fmul_rn(0.044715f, val), val), val));
        const float th = (float)tanh(0.7978845608028654 * (double)inner);
        return __fmul_rn(val, __fmul_rn(0.5f, __fadd_rn(1.0f, th)));
    }
    return __fmul_rn(val, __fdiv_rn(1.0f, __fadd_rn(1.0f, expf_glibc_t(-val, exp_tab))));   // SiLU (:617), exp = glibc expf
}
template <int EPI, bool LL>
LMRS_DEVINL void store_row(const GemvParams& p, int row, float v, uint32_t pos, uint32_t seq) {
    if constexpr (EPI == EPI_QKV) {
        if (row < p.att_dim) act_store<LL>(p.out, row, v, seq);
        else if (row < p.att_dim + p.kv_dim) act_store<LL>(p.out_k, row - p.att_dim, v, seq);
        else if constexpr (LL) act_store<true>(p.out_v, row - p.att_dim - p.kv_dim, v, seq);   // staging row: the attention kernel files it
        else reinterpret_cast<float*>(p.out_v)[(size_t)pos * p.kv_dim + (row - p.att_dim - p.kv_dim)] = v;
    } else if constexpr (EPI == EPI_LOGITS) {
        if (row < p.softcap_rows) {
            float t = __fdiv_rn(v, 30.0f);   // src/transformer.rs:375-381
            t = (float)tanh((double)t);
            v = __fmul_rn(t, 30.0f);
        }
        if (p.px_world > 1) {   // N-GPU mode: this rank's vocabulary rows go straight into every GPU's logits buffer
            for (int r = 0; r < p.px_world; r++) p.px_logits[r][row] = v;
        } else {
            reinterpret_cast<float*>(p.out)[row] = v;   // logits leave the step: plain f32
        }
    } else {
        if (!LL && p.px_world > 1) {   // N-GPU mode: push the partial into this rank's slot on every GPU (NVLink stores)
            for (int r = 0; r < p.px_world; r++) ll_store_sys(p.px_out[r] + row, v, seq);
        } else {
            act_store<LL>(p.out, row, v, seq);
        }
    }
}
template <int QT, int EPI, bool LL>
LMRS_DEVINL void consume_stage(const GemvParams& p, const WarpStreams<QT>& w, int s, const uint8_t* buf, const GemvSmem& sm,
                               Consumer<QT>& c, uint32_t pos, uint32_t seq) {
    constexpr int QB = QTraits<QT>::QB;
    constexpr bool GLU = (EPI == EPI_GLU_SILU || EPI == EPI_GLU_GELU);
    const int lane = threadIdx.x & 31, half = lane >> 4, l16 = lane & 15;
    const int n = p.n, G = w.G;
    const RowRange rr = half ? w.r1 : w.r0;
    const int ng = half ? w.ng1 : w.ng0;
    const int f = s * SG + l16;           // index of my group inside my stream
    const bool valid = f < ng;
    int row_l, g;
    if (c.aligned) { row_l = c.row_l; g = c.g_base + l16; }
    else { row_l = f / G; g = f - row_l * G; }
    float t = 0.0f;
    if (valid) {
        int iv0 = 0, iv1 = 0;
        const int4* wv = reinterpret_cast<const int4*>(buf + half * blk_bytes<QT>() + l16 * QB);
        if (QT == 1) {
            if (c.xreg) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int4 w4 = wv[(i + l16) & 7], x4 = c.xr[i];
                    iv0 = dp4a_ss(w4.x, x4.x, iv0); iv1 = dp4a_ss(w4.y, x4.y, iv1);
                    iv0 = dp4a_ss(w4.z, x4.z, iv0); iv1 = dp4a_ss(w4.w, x4.w, iv1);
                }
            } else {
                const int4* xv = reinterpret_cast<const int4*>(sm.xq + (size_t)g * GS);
#pragma unroll
                for (int i = 0; i < 8; i++) {   // 16-byte column rotated by lane: conflict-free LDS.128
                    const int cc = (i + l16) & 7;
                    const int4 w4 = wv[cc], x4 = xv[cc];
                    iv0 = dp4a_ss(w4.x, x4.x, iv0); iv1 = dp4a_ss(w4.y, x4.y, iv1);
                    iv0 = dp4a_ss(w4.z, x4.z, iv0); iv1 = dp4a_ss(w4.w, x4.w, iv1);
                }
            }
            iv0 += iv1;
        } else {
            const int4* ev = reinterpret_cast<const int4*>(sm.xq + (size_t)g * (GS / 2));
            const int4* ov = reinterpret_cast<const int4*>(sm.xq + (size_t)(n / 2) + (size_t)g * (GS / 2));
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int cc = (i + (l16 >> 1)) & 3;
                const int4 w4 = wv[cc], e4 = ev[cc], o4 = ov[cc];
                iv0 = dp4a_su(e4.x, (uint32_t)w4.x & 0x0F0F0F0Fu, iv0); iv1 = dp4a_su(o4.x, ((uint32_t)w4.x >> 4) & 0x0F0F0F0Fu, iv1);
                iv0 = dp4a_su(e4.y, (uint32_t)w4.y & 0x0F0F0F0Fu, iv0); iv1 = dp4a_su(o4.y, ((uint32_t)w4.y >> 4) & 0x0F0F0F0Fu, iv1);
                iv0 = dp4a_su(e4.z, (uint32_t)w4.z & 0x0F0F0F0Fu, iv0); iv1 = dp4a_su(o4.z, ((uint32_t)w4.z >> 4) & 0x0F0F0F0Fu, iv1);
                iv0 = dp4a_su(e4.w, (uint32_t)w4.w & 0x0F0F0F0Fu, iv0); iv1 = dp4a_su(o4.w, ((uint32_t)w4.w >> 4) & 0x0F0F0F0Fu, iv1);
            }
            iv0 = iv0 + iv1 - 8 * sm.xsum[g];   // sum x_s*(w_u - 8) = sum x_s*w_u - 8*sum x_s
        }
        const float wsc = reinterpret_cast<const float*>(buf + half * blk_bytes<QT>() + SG * QB)[l16];
        t = __fmul_rn(__fmul_rn((float)iv0, wsc), sm.xs[g]);   // (ival*ws)*xs, src/functional.rs:207,246
    }
    if (c.aligned) {
        // the whole half-warp stage belongs to one row: plain ordered sum of its 16 terms, ascending group index
        float a = (c.g_base == 0) ? 0.0f : c.acc;
#pragma unroll
        for (int j = 0; j < SG; j++) a = __fadd_rn(a, __shfl_sync(0xffffffffu, t, j, 16));
        c.acc = a;
        const bool row_done = (c.g_base + SG == G);
        if constexpr (GLU) {
            const float up = __shfl_sync(0xffffffffu, a, 16);
            if (row_done && lane == 0 && valid) act_store<LL>(p.out, rr.row0 + row_l, __fmul_rn(glu_act(EPI, a, sm.exp_tab), up), seq);
        } else if (row_done && l16 == 0 && valid) {
            store_row<EPI, LL>(p, rr.row0 + row_l, a, pos, seq);
        }
        c.g_base += SG;
        if (c.g_base == G) { c.g_base = 0; c.row_l++; }
        return;
    }
    // generic path (G not a multiple of 16: tiny test models, Gemma-2-9B's dim 3584): rows may start/end anywhere inside the stage
    const bool is_last = valid && (g == G - 1);
    const uint32_t first_mask = __ballot_sync(0xffffffffu, valid && g == 0) >> (half * 16);
    float mine = 0.0f, acc = c.acc;
#pragma unroll
    for (int j = 0; j < SG; j++) {
        const float tj = __shfl_sync(0xffffffffu, t, j, 16);
        acc = __fadd_rn(((first_mask >> j) & 1u) ? 0.0f : acc, tj);
        if (j == l16) mine = acc;
    }
    c.acc = acc;
    if constexpr (GLU) {
        const float up = __shfl_sync(0xffffffffu, mine, l16 + 16);
        if (is_last && half == 0) act_store<LL>(p.out, rr.row0 + row_l, __fmul_rn(glu_act(EPI, mine, sm.exp_tab), up), seq);
    } else if (is_last) {
        store_row<EPI, LL>(p, rr.row0 + row_l, mine, pos, seq);
    }
}

// carve one CTA's activation area (after the ring) -- must match gemv_smem_bytes
LMRS_DEVINL GemvSmem carve_gemv_smem(uint8_t* base, int n, int n_bars) {
    GemvSmem sm;
    const int G = n / GS;
    sm.xq = base;
    sm.xs = reinterpret_cast<float*>(sm.xq + ((n + 127) / 128) * 128);
    sm.xsum = reinterpret_cast<int*>(sm.xs + G);
    sm.red = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sm.xs) + ((G * 8 + 127) / 128) * 128);
    sm.xf = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sm.red + 64) + n_bars * 8 + 64);
    sm.exp_tab = kExp2fTab;
    return sm;
}

// the quantized matrix-vector kernel (name chosen not to collide with library GEMV symbols in launch classifiers)
#ifndef LMRS_GEMV_MINB
#define LMRS_GEMV_MINB 2   // CTAs per SM the register allocation must allow: 2 = consecutive kernels of the chain co-reside
#endif
template <int QT, int WARPS, int DEPTH, int PRO, int EPI, bool LL>
__global__ void __launch_bounds__(WARPS * 32, WARPS <= 8 ? LMRS_GEMV_MINB : 1) lmrs_q_matvec_kernel(const GemvParams p) {
    constexpr int STAGE = gemv_stage_bytes<QT>();
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = smem;
    GemvSmem sm = carve_gemv_smem(ring + (size_t)WARPS * DEPTH * STAGE, p.n, WARPS * DEPTH);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm.red + 64) + warp * DEPTH;
    StreamDesc sd = stream_desc(p);
    sd.epi = EPI;
    const WarpStreams<QT> w = make_streams<QT>(sd, blockIdx.x * WARPS + warp, gridDim.x * WARPS);

    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < DEPTH; d++) mbar_init(&bars[d], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint64_t pol = l2_policy_evict_first();
    const long long c0 = ktrace_c0();
    if (blockIdx.x == 0 && threadIdx.x == 0) ktrace(p.trace_slot, 0);
    pdl_launch_dependents();
    const uint32_t seq = (LL || p.px_world > 1) ? p.step->seq : 0u;
    // LL mode: the kernel is resident long before its inputs exist.  It parks on ONE word of its freshest input (the pending
    // residual contribution / the activation to quantize) and only then starts streaming: a weight prefetch issued at CTA
    // start lands in the middle of the PREVIOUS kernel's latency-bound stream and delays it by more than it saves here
    // (measured: co-resident prefetching kernels cost ~60 us per step), while after the canary it overlaps this kernel's
    // own prologue.  Plain mode: stream first, then griddepcontrol.wait (inside the prologue).
    if constexpr (LL) {
        const void* fresh = PRO == PRO_NORM ? p.delta : (PRO == PRO_QUANT ? p.act_in : nullptr);
        if (fresh) ll_canary_wait(reinterpret_cast<const llword_t*>(fresh), seq, p.ll_nowait != 0);
    }
    // Plain mode: weights never depend on the previous kernel, but what is requested before the dependency wait competes
    // with the previous kernel's last, latency-bound stages; `pre_stages` of the ring are requested now, the others right
    // after the wait (they still land during this kernel's prologue).
    const int pre = LL ? DEPTH : min(p.pre_stages, DEPTH);
    if (lane == 0)
        for (int s = 0; s < pre && s < w.nst; s++) issue_stage<QT>(w, s, ring + (size_t)(warp * DEPTH + s) * STAGE, &bars[s], pol);
    if constexpr (!LL) {
        // this kernel's CTAs were launched early and now idle until the previous kernel completes: one otherwise unused
        // thread asks the L2 to fetch weights that later kernels of the step will stream (after this CTA's own requests)
        if (p.l2pf_bytes && threadIdx.x == (WARPS - 1) * 32 + 1) l2_prefetch_slice(p.l2pf_ptr, p.l2pf_bytes, p.l2pf_chunk);
        // the norm weights are read once per token and have long left the L2 when their block comes round again (1.3 GB of
        // weights stream through it per step): ask for them now, the prologue's loads after the wait then hit the L2
        if constexpr (PRO == PRO_NORM) {
            const int off = threadIdx.x * 32;   // floats: one 128-byte line per thread
            if (off < p.n) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.w_norm + off));
                if (p.w_post) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.w_post + off));
            }
        }
        // N-GPU mode, consumer of an exchanged vector: everything this kernel reads from its predecessor arrives as
        // (value, sequence) words it polls anyway, so it does not wait for the predecessor's COMPLETION (which includes the
        // NVLink round trip of that kernel's stores into the peers) -- only for the words themselves (gemv_prologue)
        if (!(PRO == PRO_NORM && p.px_world > 1 && p.px_in != nullptr))
            pdl_wait();   // upstream activations are complete and visible from here on
        if (lane == 0)
            for (int s = pre; s < DEPTH && s < w.nst; s++) issue_stage<QT>(w, s, ring + (size_t)(warp * DEPTH + s) * STAGE, &bars[s], pol);
    }
    if constexpr (EPI == EPI_GLU_SILU) {   // expf table -> shared memory (last 256 B), by the LAST warp, after the weight prefetch was issued
        uint64_t* tab = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sm.xf) + (PRO == PRO_NORM ? (size_t)p.n * 4 + 128 : 0));
        if (warp == WARPS - 1) tab[lane] = kExp2fTab[lane];
        sm.exp_tab = tab;
    }
    // LL mode: no dependency wait at all -- the prologue polls the activation words; the step parameters were written before
    // the step's first (ordinary) launch.  (A warm-up pass that ran prologue + first stage on dummy data before the inputs arrive, to prime
    // the instruction cache, measured no gain and was removed.)
    if (lane == 0) { ktrace(p.trace_slot, 1); ktrace_c(p.trace_slot, 1, c0); }
    gemv_prologue<QT, WARPS, PRO, LL>(p, sm, seq);
    if (lane == 0) { ktrace(p.trace_slot, 2); ktrace_c(p.trace_slot, 2, c0); }
    const uint32_t pos = (EPI == EPI_QKV && !LL) ? p.step->pos : 0u;
    Consumer<QT> cs;
    consumer_begin<QT>(cs, w, sm);
    for (int s = 0; s < w.nst; s++) {
        const int d = s % DEPTH;
        mbar_wait(&bars[d], (uint32_t)((s / DEPTH) & 1));
        consume_stage<QT, EPI, LL>(p, w, s, ring + (size_t)(warp * DEPTH + d) * STAGE, sm, cs, pos, seq);
        __syncwarp();
        if (lane == 0 && s + DEPTH < w.nst) issue_stage<QT>(w, s + DEPTH, ring + (size_t)(warp * DEPTH + d) * STAGE, &bars[d], pol);
    }
    if (lane == 0) { ktrace(p.trace_slot, 3); ktrace_c(p.trace_slot, 3, c0); }
}

}  // namespace lmrs
