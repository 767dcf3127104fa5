// common.cuh -- sm_100a PTX helpers shared by the lmrs_b200 kernels: mbarrier, 1-D bulk async copies
// (TMA engine, SASS UBLKCP), programmatic dependent launch, exact-rounding float helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LMRS_DEVINL __device__ __forceinline__

namespace lmrs {

LMRS_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier (shared::cta) ----------------------------------------------------------------------------
LMRS_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make barrier inits (generic proxy) visible to the async proxy (bulk copies complete_tx on them)
LMRS_DEVINL void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
LMRS_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
LMRS_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
LMRS_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- 1-D bulk async copy global -> shared, completion signalled on an mbarrier (cp.async.bulk) ---------
// dst, src 16-byte aligned, bytes a multiple of 16.
LMRS_DEVINL void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// same with an L2 eviction-priority hint (weights are streamed once per token: evict_first)
LMRS_DEVINL uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
LMRS_DEVINL void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}

// fire-and-forget request to bring [src, src + bytes) into L2 (no shared-memory destination, no completion to wait for)
LMRS_DEVINL void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
LMRS_DEVINL void bulk_prefetch_l2_hint(const void* src_gmem, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src_gmem), "r"(bytes), "l"(pol) : "memory");
}
// one thread per CTA: this CTA's 1/gridDim.x slice of the range, in `chunk`-byte requests (everything 16-byte aligned)
// evict_first (chunk < 0 selects it): the lines are read exactly once, soon; they must not push the KV cache out of L2
LMRS_DEVINL void l2_prefetch_slice(const uint8_t* base, unsigned long long bytes, int chunk) {
    const bool ef = chunk < 0;
    if (ef) chunk = -chunk;
    const uint64_t pol = ef ? l2_policy_evict_first() : 0ull;
    const unsigned long long per = ((bytes / gridDim.x) + 4095ull) & ~4095ull;
    const unsigned long long lo = (unsigned long long)blockIdx.x * per;
    const unsigned long long hi = lo + per < bytes ? lo + per : bytes;
    for (unsigned long long o = lo; o < hi; o += (unsigned long long)chunk) {
        const uint32_t nb = (uint32_t)((hi - o < (unsigned long long)chunk ? hi - o : (unsigned long long)chunk) & ~15ull);
        if (ef) bulk_prefetch_l2_hint(base + o, nb, pol); else bulk_prefetch_l2(base + o, nb);
    }
}

// ---- programmatic dependent launch ---------------------------------------------------------------------
// pdl_wait(): block until every kernel this launch depends on has completed and flushed (no-op when the
// kernel was launched without the PDL attribute).  pdl_launch_dependents(): allow the next kernel in the
// stream to start its pre-wait portion (weight prefetch) now.
LMRS_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
LMRS_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- "LL" activation exchange between co-resident kernels of one decode step --------------------------------
// A gpu-scope release/acquire pair costs ~1 us on B200 (each fence ~0.5 us, tools/micro/barrier_probe.cu) and a
// kernel boundary 2-3 us (griddepcontrol.wait + first loads).  The decode chain therefore hands activations over
// WITHOUT fences: every element travels as one 64-bit word (f32 bits | step sequence number << 32) written with a
// single relaxed 8-byte store (single-copy atomic), and a consumer polls the words themselves until they carry the
// sequence number of the current step.  Every buffer is written once per step (one buffer per phase), steps are
// separated by an ordinary, fully ordered launch, so there is no reuse hazard.  `nowait` (measurement passes that run
// only some of the kernels): accept whatever is there.
typedef unsigned long long llword_t;
LMRS_DEVINL void ll_store(llword_t* p, float v, uint32_t seq) {
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(v)) : "memory");
}
LMRS_DEVINL void ll_ld2(const llword_t* p, llword_t& a, llword_t& b) {
    asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
LMRS_DEVINL llword_t ll_ld1(const llword_t* p) {
    llword_t a;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
LMRS_DEVINL bool ll_ok(llword_t w, uint32_t seq) { return (uint32_t)(w >> 32) == seq; }
LMRS_DEVINL float ll_val(llword_t w) { return __uint_as_float((uint32_t)w); }
// a spinning consumer gives up loudly (trap -> launch failure) instead of hanging the GPU if its producer never runs
struct LLSpin { long long t0; };
LMRS_DEVINL LLSpin ll_spin_begin() { LLSpin s; s.t0 = clock64(); return s; }
LMRS_DEVINL void ll_spin_check(const LLSpin& s) { if (clock64() - s.t0 > 6000000000LL) __trap(); }
// Whole-CTA wait for a "canary" word of a vector another kernel is still producing: ONE thread polls (with a short
// back-off), the others sit in the barrier.  Kernels of later phases are resident long before their inputs exist; if all
// their threads polled, the L2 would be saturated by polling traffic (measured: the step got SLOWER than with kernel-
// boundary hand-overs).  After the canary the caller gathers its elements (each still validated individually).
LMRS_DEVINL void ll_canary_wait(const llword_t* p, uint32_t seq, bool nowait) {
    if (!nowait && threadIdx.x == 0) {
        const LLSpin sp = ll_spin_begin();
        while (!ll_ok(ll_ld1(p), seq)) { __nanosleep(40); ll_spin_check(sp); }
    }
    __syncthreads();
}
// one element / four consecutive elements (32-byte aligned), spinning until they belong to step `seq`
LMRS_DEVINL float ll_wait1(const llword_t* p, uint32_t seq, bool nowait) {
    llword_t w = ll_ld1(p);
    if (!nowait && !ll_ok(w, seq)) { const LLSpin sp = ll_spin_begin(); do { __nanosleep(20); ll_spin_check(sp); w = ll_ld1(p); } while (!ll_ok(w, seq)); }
    return ll_val(w);
}
LMRS_DEVINL bool ll_try4(const llword_t* p, uint32_t seq, bool nowait, float4& out) {
    llword_t a, b, c, d;
    ll_ld2(p, a, b); ll_ld2(p + 2, c, d);
    if (!nowait && !(ll_ok(a, seq) && ll_ok(b, seq) && ll_ok(c, seq) && ll_ok(d, seq))) return false;
    out = make_float4(ll_val(a), ll_val(b), ll_val(c), ll_val(d));
    return true;
}
LMRS_DEVINL float4 ll_wait4(const llword_t* p, uint32_t seq, bool nowait) {
    float4 v;
    if (!ll_try4(p, seq, nowait, v)) { const LLSpin sp = ll_spin_begin(); do { __nanosleep(20); ll_spin_check(sp); } while (!ll_try4(p, seq, nowait, v)); }
    return v;
}

// ---- the same words BETWEEN GPUs (row-sharded N-GPU mode): every GPU pushes its partial result vector straight into the
// exchange buffers of all its peers over NVLink (peer-mapped pointers), one 8-byte (value, sequence) word per element --
// the protocol NCCL's LL path uses, but issued from the epilogue of the producing kernel, so no collective sits in the chain.
// System scope: the consumer polls its OWN memory, which the peers write through this GPU's L2.
LMRS_DEVINL void ll_store_sys(llword_t* p, float v, uint32_t seq) {
    asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(p), "l"(((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(v)) : "memory");
}
LMRS_DEVINL llword_t ll_ld1_sys(const llword_t* p) {
    llword_t a;
    asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
LMRS_DEVINL bool ll_try4_sys(const llword_t* p, uint32_t seq, bool nowait, float4& out) {
    llword_t a, b, c, d;
    asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 2) : "memory");
    if (!nowait && !(ll_ok(a, seq) && ll_ok(b, seq) && ll_ok(c, seq) && ll_ok(d, seq))) return false;
    out = make_float4(ll_val(a), ll_val(b), ll_val(c), ll_val(d));
    return true;
}
// a peer may be a whole host-side launch behind (graph instantiation, a slow rank): wait longer before giving up
LMRS_DEVINL void px_spin_check(const LLSpin& s) { if (clock64() - s.t0 > 40000000000LL) __trap(); }
LMRS_DEVINL float px_wait1(const llword_t* p, uint32_t seq, bool nowait) {
    llword_t w = ll_ld1_sys(p);
    if (!nowait && !ll_ok(w, seq)) { const LLSpin sp = ll_spin_begin(); do { __nanosleep(40); px_spin_check(sp); w = ll_ld1_sys(p); } while (!ll_ok(w, seq)); }
    return ll_val(w);
}
// whole CTA: one thread per source GPU parks on the first word of that GPU's slot (stride n words), everybody else sits
// in the barrier -- 148 CTAs polling with every thread would flood the L2 the peers' stores have to get through
LMRS_DEVINL void px_canary_wait(const llword_t* base, int world, int n, uint32_t seq, bool nowait) {
    if (!nowait && (int)threadIdx.x < world) {
        const llword_t* p = base + (size_t)threadIdx.x * n;
        const LLSpin sp = ll_spin_begin();
        while (!ll_ok(ll_ld1_sys(p), seq)) { __nanosleep(40); px_spin_check(sp); }
    }
    __syncthreads();
}
constexpr int PX_MAX_WORLD = 8;

// ---- integer dot products ------------------------------------------------------------------------------
LMRS_DEVINL int dp4a_ss(int a, int b, int c) { return __dp4a(a, b, c); }
// a: 4 x s8, b: 4 x u8
LMRS_DEVINL int dp4a_su(int a, uint32_t b, int c) {
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---- Rust-semantics scalar helpers (no FMA contraction anywhere on the parity path) ---------------------
// f32::round (half away from zero) followed by `as i8` (saturating, NaN -> 0): src/quantization.rs:63
LMRS_DEVINL int round_sat_i8(float v) {
    float r = roundf(v);
    if (!(r == r)) return 0;
    r = fminf(fmaxf(r, -128.0f), 127.0f);
    return (int)r;
}
// (v + 8.0).round() as u8, clamp(0, 15): src/quantization.rs:89-90
LMRS_DEVINL int round_sat_u4(float v) {
    float r = roundf(__fadd_rn(v, 8.0f));
    if (!(r == r)) return 0;
    r = fminf(fmaxf(r, 0.0f), 15.0f);
    return (int)r;
}

// ---- developer trace: (tag | clock, globaltimer) events of CTA 0 / thread 0 when enabled (LMRS_B200_TIMING=1) ----
// The event counter lives in a register-like __shared__ word so that an event costs ~40 cycles, not an L2 round trip.
__device__ unsigned long long* g_trace_buf = nullptr;   // [2 * cap]
LMRS_DEVINL unsigned int& trace_counter() {
    __shared__ unsigned int n;
    return n;
}
LMRS_DEVINL void trace_reset() {
    if (threadIdx.x == 0) trace_counter() = 0;
}
LMRS_DEVINL void trace_value(int tag, unsigned long long value) {   // value in the timestamp slot
#ifndef LMRS_TRACE
    (void)tag; (void)value;
    return;   // compiled out by default: even the disabled check is a global load on the critical path
#endif
    if (g_trace_buf != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned int i = trace_counter()++;
        if (i < 8192u) { g_trace_buf[2 * i] = ((unsigned long long)clock64() << 16) | (unsigned long long)(tag & 0xffff); g_trace_buf[2 * i + 1] = value; }
    }
}
LMRS_DEVINL void trace_event(int tag) {
#ifndef LMRS_TRACE
    (void)tag;
    return;
#endif
    if (g_trace_buf != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        trace_value(tag, t);
    }
}

// per-launch timeline (LMRS_TRACE builds, LMRS_B200_TIMING=1): slot = launch index within the step, 8 words per slot;
// word k keeps the LATEST globaltimer value any calling thread reported for event k (k = 0: first CTA's start)
LMRS_DEVINL void ktrace(const int slot, const int k) {
#ifdef LMRS_TRACE
    if (g_trace_buf != nullptr && slot >= 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        atomicMax(&g_trace_buf[(size_t)slot * 8 + k], t);
    }
#else
    (void)slot; (void)k;
#endif
}

// cycle-resolution companion: word 1024*8 + slot*16 + k keeps the largest (clock64() - c0) any calling thread reported,
// c0 being that CTA's own clock at kernel entry (ktrace_c0)
LMRS_DEVINL long long ktrace_c0() {
#ifdef LMRS_TRACE
    return clock64();
#else
    return 0;
#endif
}
LMRS_DEVINL void ktrace_c(const int slot, const int k, const long long c0) {
#ifdef LMRS_TRACE
    if (g_trace_buf != nullptr && slot >= 0) atomicMax(&g_trace_buf[8192 + (size_t)slot * 16 + k], (unsigned long long)(clock64() - c0));
#else
    (void)slot; (void)k; (void)c0;
#endif
}

LMRS_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
LMRS_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace lmrs
