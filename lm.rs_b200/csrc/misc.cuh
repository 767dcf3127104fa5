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
    const void* x_in; const void* delta; const float* w_post;   // f32 arrays, or LL words when ll (common.cuh)
    int n; float eps;
    float* rows;               // [n_rows][n]; row index = step->token
    const StepParams* step;
    int row_from_block;        // batched prefill: row = blockIdx.x and x_in/delta are row-strided too
    int ll, ll_nowait;         // serial prefill inside the LL decode chain (never with row_from_block)
    float* scratch;            // ll / px: [n] f32 scratch for the unpacked contribution
    int px_world;              // N-GPU peer exchange: delta = rank-ordered sum of the px_world slots of px_in ([px_world][n] words)
    const llword_t* px_in;
};
LMRS_DEVINL void residual_finalize_body(const ResidualParams& p, float* red) {
    const size_t row = p.row_from_block ? (size_t)blockIdx.x : (size_t)p.step->token;
    float* out = p.rows + row * p.n;
    const float* x_in = reinterpret_cast<const float*>(p.x_in) + (p.row_from_block ? row * p.n : 0);
    const float* delta = reinterpret_cast<const float*>(p.delta) + (p.row_from_block ? row * p.n : 0);
    const uint32_t seq = p.ll ? p.step->seq : 0u;
    if (p.px_world > 1) {   // N-GPU mode: sum the partials the GPUs pushed here, in rank order (see gemv.cuh px_gather_sum)
        const uint32_t pseq = p.step->seq;
        px_canary_wait(p.px_in, p.px_world, p.n, pseq, p.ll_nowait != 0);
        for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
            float d = px_wait1(p.px_in + i, pseq, p.ll_nowait != 0);
            for (int r = 1; r < p.px_world; r++) d = __fadd_rn(d, px_wait1(p.px_in + (size_t)r * p.n + i, pseq, p.ll_nowait != 0));
            p.scratch[i] = d;
        }
        __syncthreads();
        delta = p.scratch;
    } else if (p.ll) {   // unpack delta once (the exact chain below wants a plain array)
        ll_canary_wait(reinterpret_cast<const llword_t*>(p.delta), seq, p.ll_nowait != 0);
        for (int i = threadIdx.x; i < p.n; i += blockDim.x)
            p.scratch[i] = ll_wait1(reinterpret_cast<const llword_t*>(p.delta) + i, seq, p.ll_nowait != 0);
        __syncthreads();
        delta = p.scratch;
    }
    float r = 1.0f;
    if (p.w_post) r = exact_rnorm(delta, p.n, p.eps, red);   // chains read global memory directly
    for (int i = threadIdx.x; i < p.n; i += blockDim.x) {
        float d = (p.ll || p.px_world > 1) ? delta[i] : __ldcg(delta + i);
        if (p.w_post) d = __fmul_rn(__fadd_rn(1.0f, p.w_post[i]), __fmul_rn(r, d));
        const float xv = p.ll ? ll_wait1(reinterpret_cast<const llword_t*>(p.x_in) + i, seq, p.ll_nowait != 0) : __ldcg(x_in + i);
        out[i] = __fadd_rn(xv, d);
    }
}
__global__ void __launch_bounds__(256) residual_finalize_kernel(const ResidualParams p) {
    __shared__ float red[32];
    pdl_launch_dependents();
    if (!p.ll) pdl_wait();
    residual_finalize_body(p, red);
}

// ---- N-GPU mode: "my logits rows have landed everywhere" ------------------------------------------------------------------
// Runs after the classifier (griddepcontrol.wait: all its stores, including the ones into the peers' logits buffers, are
// performed).  Thread r tells GPU r so (release at system scope) and waits until GPU r has told this GPU the same: when the
// kernel ends, this GPU's logits buffer holds every rank's rows of the step.
struct PeerFlagParams {
    uint32_t* flag_peer[PX_MAX_WORLD];   // GPU r's flag array (peer-mapped); this rank writes element `rank`
    const uint32_t* flag_local;          // this GPU's flag array; element r is written by GPU r
    int world, rank, nowait;
    const StepParams* step;              // decode: the flag value is the step's sequence number (waits for equality)
    uint32_t value; int use_value;       // batched prefill: a counter that only grows (waits for >=, wrap-around aware)
};
__global__ void __launch_bounds__(32) peer_flag_kernel(const PeerFlagParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const int r = threadIdx.x;
    if (r >= p.world) return;
    const uint32_t seq = p.use_value ? p.value : p.step->seq;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flag_peer[r] + p.rank), "r"(seq) : "memory");
    if (p.nowait) return;
    const LLSpin sp = ll_spin_begin();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.flag_local + r) : "memory");
        if (p.use_value ? (int32_t)(v - seq) >= 0 : v == seq) break;
        __nanosleep(100);
        px_spin_check(sp);
    }
}

// ---- N-GPU batched prefill: the partial [T][dim] result of a K-sharded GEMM travels to every GPU as plain f32 rows (bandwidth-
// bound, 4-7 MB: no per-word flags), one flag exchange (peer_flag_kernel) publishes it, and the sum over the GPUs is formed
// in ascending rank order like the decode path's (the oracle's k-shard mode restates exactly this order).
struct PxRowsParams {
    const float* src;                    // push: this GPU's partial [count4 * 4]
    float* dst[PX_MAX_WORLD];            // push: slot `rank` of the exchange buffer on every GPU (peer-mapped)
    const float* slots;                  // sum: this GPU's exchange buffer, slot r at slots + r * slot_stride
    float* out;                          // sum: [count4 * 4]
    int world; size_t count4, slot_stride;
};
__global__ void __launch_bounds__(256) px_rows_push_kernel(const PxRowsParams p) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= p.count4) return;
    const float4 v = __ldcg(reinterpret_cast<const float4*>(p.src) + i);
    for (int r = 0; r < p.world; r++) reinterpret_cast<float4*>(p.dst[r])[i] = v;
}
__global__ void __launch_bounds__(256) px_rows_sum_kernel(const PxRowsParams p) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= p.count4) return;
    float4 v = __ldcg(reinterpret_cast<const float4*>(p.slots) + i);
    for (int r = 1; r < p.world; r++) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(p.slots + (size_t)r * p.slot_stride) + i);
        v.x = __fadd_rn(v.x, t.x); v.y = __fadd_rn(v.y, t.y); v.z = __fadd_rn(v.z, t.z); v.w = __fadd_rn(v.w, t.w);
    }
    reinterpret_cast<float4*>(p.out)[i] = v;
}

// ---- batched prefill row kernels (fill_kv_cache, src/transformer.rs:672-684 -> forward_layer with sl = N) ----------
// one CTA per token row: the GEMV prologue (residual add, exact rmsnorm, activation quantize) with its result written
// to HBM as the int8 A operand + scales of the tcgen05 GEMM
// MAXC: float4 chunks of the row per thread (n <= 1024 * MAXC): the fewer, the fewer registers and the more rows per SM
template <int PRO, int MAXC>
__global__ void __launch_bounds__(256, MAXC <= 2 ? 3 : 2) rows_prologue_kernel(GemvParams p, uint8_t* xq_out, float* xs_out) {
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
    if (p.x_in) p.x_in = reinterpret_cast<const float*>(p.x_in) + row * n;
    if (p.delta) p.delta = reinterpret_cast<const float*>(p.delta) + row * n;
    if (p.x_out) p.x_out = reinterpret_cast<float*>(p.x_out) + row * n;
    if (p.act_in) p.act_in = reinterpret_cast<const float*>(p.act_in) + row * n;
    p.xout_all = 1;
    gemv_prologue<1, 8, PRO, false, MAXC>(p, sm, 0u);
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

// ---- greedy sampler on the device: Sampler::sample_argmax (src/sampler.rs:29-41) -------------------------------------
// The reference scans the logits serially with a strict `>`: the FIRST maximum wins, a NaN never replaces the running
// maximum, and a NaN at index 0 makes every comparison false (result 0).  (value, index) pairs ordered by
// "greater value, then lower index" reproduce that scan exactly under any reduction tree; NaNs at i >= 1 are mapped to
// -inf (they can never win) and a NaN at index 0 is answered directly.  One launch: every CTA reduces a contiguous slice,
// the last CTA to arrive (atomic ticket) reduces the per-CTA partials and writes the 4-byte token id.
struct ArgmaxParams {
    const float* x; int n;
    float* pmax; int* pidx;        // [gridDim.x] per-CTA partials
    unsigned* ticket;              // zero on entry, left zero
    uint32_t* out;                 // device word the host copies back (4 bytes instead of vocab * 4)
    StepParams* advance;           // device-side generate loop: feed the token into the next step (token, pos+1, seq+1)
};
LMRS_DEVINL bool argmax_better(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }
LMRS_DEVINL void argmax_block_reduce(float& v, int& i, float* sv, int* si) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (argmax_better(ov, oi, v, i)) { v = ov; i = oi; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sv[warp] = v; si[warp] = i; }
    __syncthreads();
    if (warp == 0) {
        v = lane < (int)(blockDim.x >> 5) ? sv[lane] : -INFINITY;
        i = lane < (int)(blockDim.x >> 5) ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (argmax_better(ov, oi, v, i)) { v = ov; i = oi; }
        }
    }
    __syncthreads();
}
__global__ void __launch_bounds__(256) argmax_kernel(const ArgmaxParams p) {
    __shared__ float sv[8];
    __shared__ int si[8];
    __shared__ bool last;
    pdl_launch_dependents();
    pdl_wait();
    const int per = (p.n + gridDim.x - 1) / gridDim.x;
    const int i0 = blockIdx.x * per, i1 = min(p.n, i0 + per);
    float v = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        float t = __ldcg(p.x + i);
        if (t != t) t = -INFINITY;
        if (argmax_better(t, i, v, idx)) { v = t; idx = i; }
    }
    argmax_block_reduce(v, idx, sv, si);
    if (threadIdx.x == 0) {
        p.pmax[blockIdx.x] = v; p.pidx[blockIdx.x] = idx;
        __threadfence();
        last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    v = -INFINITY; idx = 0x7fffffff;
    for (int c = threadIdx.x; c < (int)gridDim.x; c += blockDim.x) {
        const float t = __ldcg(p.pmax + c);
        const int ti = __ldcg(p.pidx + c);
        if (argmax_better(t, ti, v, idx)) { v = t; idx = ti; }
    }
    argmax_block_reduce(v, idx, sv, si);
    if (threadIdx.x == 0) {
        const float x0 = __ldcg(p.x);
        const uint32_t tok = (x0 != x0 || idx == 0x7fffffff) ? 0u : (uint32_t)idx;
        *p.out = tok;
        *p.ticket = 0u;
        if (p.advance) { p.advance->token = tok; p.advance->pos += 1u; p.advance->mask_base = p.advance->pos; p.advance->seq += 1u; if (p.advance->seq == 0u) p.advance->seq = 1u; }
    }
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
// src/functional.rs:80-114 (the vision tower's norm): mean and variance each accumulated in eight lane partials walked
// serially over j (x[8j+k]; then (x - mean)^2 with mul and add unfused), wide's horizontal order, /size, +eps, 1/sqrt;
// out = ((x - mean) * inv_std) * w + b, unfused.  One CTA per row; the size % 8 tail is left untouched like the reference.
__global__ void __launch_bounds__(256) layernorm_rows_kernel(float* o, const float* x, const float* w, const float* b, int size, float eps) {
    __shared__ float red[2];
    const float* xr = x + (size_t)blockIdx.x * size;
    float* orow = o + (size_t)blockIdx.x * size;
    const int n8 = size / 8 * 8;
    auto lanes_reduce = [&](float s) {   // ((a0+a4)+(a2+a6))+((a1+a5)+(a3+a7)), see exact_rnorm
        const int lane = threadIdx.x;
        const float t = __fadd_rn(s, __shfl_sync(0xffffffffu, s, (lane + 4) & 31));
        const float u = __fadd_rn(t, __shfl_sync(0xffffffffu, t, (lane + 2) & 31));
        return __fadd_rn(u, __shfl_sync(0xffffffffu, u, (lane + 1) & 31));
    };
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float s = 0.0f;
        if (lane < 8)
            for (int j = 0; j < n8 / 8; j++) s = __fadd_rn(s, xr[8 * j + lane]);
        const float mean = __fdiv_rn(__shfl_sync(0xffffffffu, lanes_reduce(s), 0), (float)size);
        float v = 0.0f;
        if (lane < 8)
            for (int j = 0; j < n8 / 8; j++) { const float d = __fsub_rn(xr[8 * j + lane], mean); v = __fadd_rn(v, __fmul_rn(d, d)); }
        const float var = __fadd_rn(__fdiv_rn(lanes_reduce(v), (float)size), eps);
        if (lane == 0) { red[0] = mean; red[1] = __fdiv_rn(1.0f, __fsqrt_rn(var)); }
    }
    __syncthreads();
    const float mean = red[0], inv_std = red[1];
    for (int i = threadIdx.x; i < n8; i += 256)
        orow[i] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(xr[i], mean), inv_std), w[i]), b[i]);
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
