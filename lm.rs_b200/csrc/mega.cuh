// mega.cuh -- the whole decode step (Transformer::forward, src/transformer.rs:316-384) as ONE persistent kernel.
//
// Why: a decode step is 5*n_layers+1 dependent matrix-vector products of 1-34 MB each.  As separate launches each one
// pays a launch gap, a pipeline fill and a drain (measured ~3.5 us per kernel, 82 kernels per Llama-1B token, against
// ~2.4 us of pure streaming time each), so HBM idles most of the time.  Here one CTA per SM stays resident for the
// whole step and every warp walks ONE continuous stream of weight stages that runs across phase boundaries:
//
//   * the schedule is static (row ranges per warp per matrix are a pure function of the warp id), so after consuming
//     a stage a warp immediately issues the bulk copy (TMA engine, cp.async.bulk + mbarrier) for the stage DEPTH
//     positions ahead in its stream -- which may belong to the next matrix or the next layer.  Weight traffic keeps
//     flowing while the CTA sits in a grid barrier, runs the exact-order rmsnorm chain or the attention phase;
//   * phases are separated by grid-wide barriers (one 64-bit arrival counter in HBM, release/acquire), 5 per layer;
//   * attention (RoPE, QK^T, softmax, AV -- attention.cuh) runs on n_kv_heads CTAs while the others wait with full
//     rings; the activation scratch and the attention scratch alias the same shared-memory region.
//
// Arithmetic is exactly that of gemv.cuh / attention.cuh (shared device functions), so results stay bit-identical.
#pragma once
#include "attention.cuh"
#include "common.cuh"
#include "gemv.cuh"
#include "misc.cuh"

namespace lmrs {

constexpr int MEGA_WARPS = 16;         // 4 warps per scheduler: the per-stage dependent chains need the extra TLP
constexpr int MEGA_MAX_DEPTH = 4;

enum { PH_GEMV = 0, PH_ATTN = 1, PH_FINALIZE = 2 };

struct MegaPhase {
    int kind;
    int pad;
    GemvParams g;
    AttnParams a;
    ResidualParams r;
};

struct MegaParams {
    const MegaPhase* phases;
    const StreamDesc* streams;     // [n_phases] compact weight-stream descriptors (o = 0 for non-GEMV phases)
    int n_phases;
    int depth;                     // ring stages per warp
    int act_n;                     // largest GEMV input length (bytes of quantized activation)
    int norm_n;                    // dim (f32 staging of the normed vector)
    int n_kv_heads, att_chunks, head_size;
    unsigned long long* bar_ctr;   // grid barrier arrival counter (monotonic across launches)
    const StepParams* step;        // step->seq numbers the launches of this kernel variant
    unsigned long long* timing;    // optional [n_phases][4][gridDim] globaltimer stamps (profiling aid), else nullptr
};

LMRS_DEVINL unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
LMRS_DEVINL unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// all CTAs of the (co-resident) grid arrive; writes before the barrier are visible to every CTA after it
LMRS_DEVINL void grid_barrier(unsigned long long* ctr, unsigned long long target) {
    __syncthreads();   // every thread's writes happen-before thread 0's release (cumulativity)
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ULL) : "memory");
        while (ld_acquire_u64(ctr) < target) __nanosleep(32);   // back off: 148 pollers on one L2 line slow everyone
    }
    __syncthreads();
}

template <int QT> __host__ __device__ constexpr size_t mega_ring_bytes(int depth) {
    return (size_t)MEGA_WARPS * depth * gemv_stage_bytes<QT>();
}
__host__ __device__ inline size_t mega_desc_bytes(int n_phases) { return ((size_t)n_phases * sizeof(StreamDesc) + 127) / 128 * 128; }
__host__ __device__ inline size_t mega_phase_bytes() { return (sizeof(MegaPhase) + 127) / 128 * 128; }
inline size_t mega_act_bytes(int act_n, int norm_n) {
    return 64 * 4 + (size_t)norm_n * 4 + (size_t)((act_n + 127) / 128) * 128 + (size_t)((act_n / GS * 8 + 127) / 128) * 128 + 128;
}

template <int QT, int HS>
__global__ void __launch_bounds__(MEGA_WARPS * 32, 1) decode_mega_kernel(const MegaParams mp) {
    constexpr int STAGE = gemv_stage_bytes<QT>();
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int depth = mp.depth;
    uint8_t* ring = smem + (size_t)warp * depth * STAGE;                                  // this warp's ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + mega_ring_bytes<QT>(depth)) + warp * MEGA_MAX_DEPTH;
    StreamDesc* sd_s = reinterpret_cast<StreamDesc*>(smem + mega_ring_bytes<QT>(depth) + MEGA_WARPS * MEGA_MAX_DEPTH * 8);
    MegaPhase* ph_s = reinterpret_cast<MegaPhase*>(reinterpret_cast<uint8_t*>(sd_s) + mega_desc_bytes(mp.n_phases));
    uint8_t* uni = reinterpret_cast<uint8_t*>(ph_s) + 2 * mega_phase_bytes();   // activation / attention union
    float* red = reinterpret_cast<float*>(uni);
    float* xf = red + 64;
    uint8_t* xq = reinterpret_cast<uint8_t*>(xf + mp.norm_n);

    trace_reset();
    if (lane == 0) {
        for (int d = 0; d < depth; d++) mbar_init(&bars[d], 1);
        fence_barrier_init();
    }
    // stream descriptors of every phase and the parameters of phase 0 -> shared memory (one round trip)
    for (int i = threadIdx.x; i < mp.n_phases * (int)(sizeof(StreamDesc) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(sd_s)[i] = reinterpret_cast<const uint32_t*>(mp.streams)[i];
    auto load_phase = [&](int ph) {   // copy phase parameters into the double buffer (consumed after the next barrier)
        const uint32_t* src = reinterpret_cast<const uint32_t*>(mp.phases + ph);
        uint32_t* dst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ph_s) + (ph & 1) * mega_phase_bytes());
        for (int i = threadIdx.x; i < (int)(sizeof(MegaPhase) / 4); i += blockDim.x) dst[i] = src[i];
    };
    load_phase(0);
    __syncthreads();

    const int wslot = blockIdx.x * MEGA_WARPS + warp, n_wslots = gridDim.x * MEGA_WARPS;
    // ---- prefetch iterator: the next (phase, stage) of this warp's weight stream --------------------------------
    int pf_phase = -1, pf_stage = 0;
    WarpStreams<QT> pf_w;
    pf_w.nst = 0;
    auto pf_seek = [&]() {   // move to the next existing stage (skipping phases where this warp owns nothing)
        while (pf_phase < mp.n_phases && (pf_phase < 0 || pf_stage >= pf_w.nst)) {
            pf_phase++;
            pf_stage = 0;
            pf_w.nst = 0;
            if (pf_phase < mp.n_phases && sd_s[pf_phase].o > 0)
                pf_w = make_streams<QT>(sd_s[pf_phase], wslot, n_wslots);
        }
    };
    uint32_t issued = 0, consumed = 0;
    const uint64_t pol = l2_policy_evict_first();
    auto pf_issue = [&]() {
        const uint32_t slot = issued % (uint32_t)depth;
        if (lane == 0) issue_stage<QT>(pf_w, pf_stage, ring + (size_t)slot * STAGE, &bars[slot], pol);
        issued++;
        pf_stage++;
        pf_seek();
    };
    pf_seek();
    while (issued < (uint32_t)depth && pf_phase < mp.n_phases) pf_issue();   // weights do not depend on anything

    const unsigned long long nbar = (unsigned long long)(mp.n_phases - 1);
    const unsigned long long base = (unsigned long long)mp.step->seq * nbar * gridDim.x;
    const uint32_t pos = mp.step->pos;

    auto stamp = [&](int ph, int k) {
        if (mp.timing && threadIdx.x == 0) mp.timing[((size_t)ph * 4 + k) * gridDim.x + blockIdx.x] = globaltimer_ns();
    };
    for (int ph = 0; ph < mp.n_phases; ph++) {
        const MegaPhase& P = *reinterpret_cast<const MegaPhase*>(reinterpret_cast<uint8_t*>(ph_s) + (ph & 1) * mega_phase_bytes());
        stamp(ph, 0);
        trace_event(1000 + ph);
        if (ph + 1 < mp.n_phases) load_phase(ph + 1);
        if (P.kind == PH_GEMV) {
            const GemvParams& g = P.g;
            GemvSmem sm;
            sm.red = red; sm.xf = xf; sm.xq = xq;
            sm.xs = reinterpret_cast<float*>(xq + ((g.n + 127) / 128) * 128);
            sm.xsum = reinterpret_cast<int*>(sm.xs + g.n / GS);
            sm.exp_tab = kExp2fTab;
            gemv_prologue<QT, MEGA_WARPS>(g, sm);
            stamp(ph, 1);
            const WarpStreams<QT> w = make_streams<QT>(sd_s[ph], wslot, n_wslots);
            Consumer<QT> cs;
            consumer_begin<QT>(cs, w, sm);
            for (int s = 0; s < w.nst; s++) {
                const uint32_t slot = consumed % (uint32_t)depth;
                mbar_wait(&bars[slot], (consumed / (uint32_t)depth) & 1u);
                consume_stage<QT>(g, w, s, ring + (size_t)slot * STAGE, sm, cs, pos);
                __syncwarp();
                consumed++;
                if (pf_phase < mp.n_phases) pf_issue();   // refill the slot just freed with the stage DEPTH ahead
            }
        } else if (P.kind == PH_ATTN) {
            const int units = mp.n_kv_heads * mp.att_chunks;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int kvh = u / mp.att_chunks, chunk = u % mp.att_chunks;
                const int h0 = kvh * P.a.kv_mul + chunk * ATT_QH;
                const int nh = min(ATT_QH, P.a.kv_mul - chunk * ATT_QH);
                __syncthreads();
                attn_decode_body<HS, MEGA_WARPS * 32>(P.a, reinterpret_cast<float*>(uni), kvh, h0, nh, chunk == 0);
            }
        } else {   // PH_FINALIZE: residual stream row back to the caller (fill_kv_cache), one CTA
            if (blockIdx.x == 0) residual_finalize_body(P.r, red);
        }
        if (mp.timing) { __syncthreads(); stamp(ph, 2); }
        trace_event(900);
        if (ph + 1 < mp.n_phases) grid_barrier(mp.bar_ctr, base + (unsigned long long)(ph + 1) * gridDim.x);
        stamp(ph, 3);
    }
}

}  // namespace lmrs
