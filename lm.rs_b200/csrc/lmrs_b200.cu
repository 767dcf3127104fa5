// lmrs_b200.cu -- host side of liblmrs_b200.so: LMRS v4 loader, HBM layout, CUDA-graph decode step and
// the C ABI declared in include/lmrs_b200.h.  Mirrors src/transformer.rs of samuel-vitorino/lm.rs:
//   Transformer::new :134-314, forward :316-384, forward_layer :388-657, get_embeddings :659-669,
//   fill_kv_cache :672-684, Drop :688-711.
// There is no CPU fallback: every entry point fails with a message when CUDA/sm_100 is unavailable.
#include "../../include/lmrs_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <algorithm>
#include <vector>

#include "attention.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "gemv.cuh"
#include "misc.cuh"
#include "prefill_attn.cuh"
#include "f32_ops.cuh"
#include "shard.h"

#include <map>

using namespace lmrs;

// one phase of a step (src/transformer.rs:316-384 + :388-657): a quantized matrix-vector product with its fused glue,
// the attention of a block, or the pending-residual write-back of a serial fill_kv_cache step
enum { PH_GEMV = 0, PH_ATTN = 1, PH_FINALIZE = 2, PH_PEERFLAG = 3 };
struct Phase {
    int kind;
    int comm;     // row-sharded (NCCL) mode: 1 = all-reduce the output, 2 = all-gather the logits
    GemvParams g;
    AttnParams a;
    ResidualParams r;
    PeerFlagParams f;
};

static thread_local std::string g_err;
extern "C" const char* lmrs_b200_last_error(void) { return g_err.c_str(); }
extern "C" const char* lmrs_b200_version(void) { return "lmrs_b200 0.1 sm_100a"; }
static int fail(const std::string& m) { g_err = m; return 1; }

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" +   \
                        std::to_string(__LINE__) + ")");                                           \
    } while (0)

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// ---- device matrix views ---------------------------------------------------------------------------------
struct Mat {
    const uint8_t* q = nullptr;     // BP16 block-packed weights (decode GEMV)
    const float* s = nullptr;
    int o = 0, n = 0, gran = 1;
    const uint8_t* dq = nullptr;    // dense file-layout copy [o][n] int8 + [o][n/128] f32 (prefill GEMM operand; Q8 only)
    const float* ds = nullptr;
    CUtensorMap tmap, tmap64;       // TMA descriptors of dq as a [o][n] byte matrix, 128 B x 128 / 64 rows box, 128B swizzle
    bool has_tmap = false;
};
struct Layer {
    Mat qkv, wo, w1, w3, w2;
    const float *rms_att = nullptr, *rms_post_att = nullptr, *rms_pre_ffn = nullptr, *rms_post_ffn = nullptr;
};

struct lmrs_b200 {
    lmrs_args_t args{};
    int device = 0, sms = 148;
    int rank = 0, world = 1;
    // per-rank shard geometry (== full model when world == 1)
    int l_heads = 0, l_kv_heads = 0, l_att_dim = 0, l_kv_dim = 0, l_hidden = 0, l_vocab = 0, vocab_off = 0;
    uint8_t* d_arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<Layer> layers;
    Mat emb, cls;
    const float* rms_final = nullptr;
    float *d_kcache = nullptr, *d_vcache = nullptr, *d_rope_cos = nullptr, *d_rope_sin = nullptr;
    float *d_x[2] = {nullptr, nullptr}, *d_q = nullptr, *d_knew = nullptr, *d_att = nullptr, *d_wo_out = nullptr;
    float *d_h = nullptr, *d_down_out = nullptr, *d_logits = nullptr, *d_scores = nullptr, *d_rows = nullptr;
    uint8_t* d_dense = nullptr;           // file-layout weights kept for the tcgen05 prefill GEMM (Q8 models)
    // batched prefill activations, capacity pf_cap rows
    size_t pf_cap = 0, pf_sc_bytes = 0;
    uint8_t* pf_xq = nullptr;
    float *pf_xs = nullptr, *pf_q = nullptr, *pf_att = nullptr, *pf_wo = nullptr, *pf_g = nullptr, *pf_u = nullptr, *pf_h = nullptr, *pf_down = nullptr, *pf_scores = nullptr;
    bool use_gemm = true;
    int pfa_quads = 4;         // fused batched-rows attention: row quads (warps) per CTA, 4 or 8 (LMRS_B200_PFA_QUADS)
    int use_pf_attn = 1;       // batched-rows attention: 1 = fused throughput kernel (prefill_attn.cuh), 2 = its two-kernel form, 0 = the per-row decode kernel
    size_t rows_cap = 0;
    StepParams* d_step = nullptr;
    StepParams* h_step_ring = nullptr;  // pinned
    int step_slot = 0;
    float* h_logits = nullptr;  // pinned
    // on-device greedy sampler (src/sampler.rs:29-41): per-CTA partials, ticket, the chosen token ids of a generate call
    float* d_amax = nullptr; int* d_aidx = nullptr; unsigned* d_ticket = nullptr; uint32_t* d_next = nullptr;
    uint32_t* d_gen = nullptr; uint32_t* h_gen = nullptr; size_t gen_cap = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    std::vector<Phase> ph_decode, ph_prefill;
    unsigned long long* d_trace = nullptr;
    // LL exchange (common.cuh): one (value, sequence) word buffer per activation per layer, written once per step;
    // step_seq numbers the steps of this handle (decode and serial prefill alike), 0 is never used
    llword_t* d_ll = nullptr;
    size_t ll_words = 0;
    bool use_ll = true;
    uint32_t step_seq = 0;
    cudaEvent_t ev_pf0 = nullptr, ev_pf1 = nullptr;   // device time of the last fill_kv_cache (kernels only, copies excluded)
    float last_prefill_ms = -1.0f;
    float* d_fin_scratch = nullptr;
    std::map<const void*, size_t> smem_optin;   // per handle (= per device): kernels whose >48 KB opt-in has been set
    // one CUDA graph per attention variant: 0..5 = cluster attention variants (setup_attn_cluster), 7 = single-CTA kernel
    cudaGraphExec_t g_decode[8] = {}, g_prefill[8] = {};
    cudaStream_t g_decode_stream[8] = {}, g_prefill_stream[8] = {};
    int n_decode_kernels[8] = {}, n_prefill_kernels[8] = {};
    int att_cl = 0;                 // CTAs per cluster of attn_cluster_kernel (0: disabled)
    struct AttVar { int cap, g; } att_var[6] = {};   // cluster-kernel variants: positions covered, head groups
    int att_nvar = 0;
    int att_variant = 7;            // variant of the step being enqueued (7: single-CTA kernel)
    uint64_t launches = 0;
    int att_chunks = 1;
    bool use_graph = true, use_pdl = true;
    int gemv_cfg = 0, gemv_ctas_per_sm = 1;
    int l2pf = 0, l2pf_cls_mb = 24, l2pf_chunk = 32768, l2pf_ef = 1;   // L2 prefetch of upcoming weights during the attention phase (make_phases)
    Shard shard;  // multi-GPU bootstrap / NCCL fallback (shard.h); inert when world == 1
    // peer exchange (common.cuh, N-GPU mode): ONE cudaMalloc'ed block per GPU, mapped into every peer through CUDA IPC:
    //   [n_layers][2][world][dim] LL words (slot r of a vector = rank r's partial) | flags | logits[vocab] |
    //   batched prefill: [2][world][pf_xchg_rows][dim] f32 partial rows
    bool use_peer = false;
    // in-process N-GPU mode (lmrs_b200_create_multi): the rank-0 handle leads a group of one handle per GPU, all driven
    // by the caller's thread; `group` is empty for ordinary handles and lists every member (itself first) on the leader
    std::vector<lmrs_b200*> group;
    bool in_process_group = false;   // member of such a group: peers are mapped with cudaDeviceEnablePeerAccess, not CUDA IPC
    uint8_t* d_xchg = nullptr;
    uint8_t* xchg_peer[PX_MAX_WORLD] = {};   // the same block on every GPU (entry `rank` = d_xchg)
    size_t xchg_flags_off = 0, xchg_logits_off = 0, xchg_pf_off = 0;
    int pf_xchg_rows = 512;          // rows per batched-prefill pass in N-GPU mode (capacity of the exchange buffers)
    uint32_t pf_xchg_count = 0;      // exchanges so far (same on every GPU): flag value and buffer parity
};

static llword_t* px_vec(const lmrs_b200* m, int gpu, size_t layer, int which) {   // exchange vector (layer, 0: after Wo / 1: after W2) on GPU `gpu`
    return reinterpret_cast<llword_t*>(m->xchg_peer[gpu]) + ((layer * 2 + (size_t)which) * m->world) * (size_t)m->args.dim;
}

static cudaError_t smem_optin(lmrs_b200* m, const void* fn, size_t smem);

// ---- kernel launch helper (optionally with the programmatic-dependent-launch attribute) -------------------
template <typename... KArgs, typename... Args>
static cudaError_t launch(lmrs_b200* m, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = m->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = m->use_pdl ? 1 : 0;
    cudaError_t ce = smem_optin(m, (const void*)kernel, smem);
    if (ce != cudaSuccess) return ce;
    m->launches++;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// function attributes are per device: the opt-in to > 48 KB of dynamic shared memory is tracked per handle
// Every kernel also asks for the maximum shared-memory carveout: an SM re-partitions L1/shared memory only when it is
// idle, so a kernel that prefers a different carveout than the one resident cannot start next to it (measured: the
// attention kernel of the LL chain started only after the QKV kernel had drained).  With one carveout for all kernels
// consecutive kernels of the chain are co-resident as far as registers and shared memory allow.
static cudaError_t smem_optin(lmrs_b200* m, const void* fn, size_t smem) {
    auto it = m->smem_optin.find(fn);
    if (it == m->smem_optin.end()) {
        static const int carve = env_int("LMRS_B200_CARVEOUT", 100);   // developer knob: -1 = leave the driver's default
        cudaError_t e = carve >= 0 ? cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, carve) : cudaSuccess;
        if (e != cudaSuccess) return e;
        it = m->smem_optin.emplace(fn, (size_t)0).first;
    }
    if (smem <= it->second || smem <= 48 * 1024) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) it->second = smem;
    return e;
}

// ---- GEMV dispatch ----------------------------------------------------------------------------------------
// ring geometries (warps, stages per warp); 8 warps x <= 128 registers: two CTAs (consecutive kernels of the chain) per SM
struct GemvCfg { int warps, depth; };
static const GemvCfg kGemvCfgs[] = {{8, 2}, {8, 3}, {8, 4}, {16, 2}, {16, 3}};
constexpr int N_GEMV_CFG = 5;
typedef void (*gemv_fn)(const GemvParams);
template <int QT, int W, int D, bool LL> static gemv_fn gemv_kernel_pe(int pro, int epi) {
    if (pro == PRO_RAW) return LL ? nullptr : (gemv_fn)lmrs_q_matvec_kernel<QT, W, D, PRO_RAW, EPI_STORE, false>;
    if (pro == PRO_QUANT) return epi == EPI_STORE ? (gemv_fn)lmrs_q_matvec_kernel<QT, W, D, PRO_QUANT, EPI_STORE, LL> : nullptr;
    switch (epi) {
        case EPI_QKV: return lmrs_q_matvec_kernel<QT, W, D, PRO_NORM, EPI_QKV, LL>;
        case EPI_GLU_SILU: return lmrs_q_matvec_kernel<QT, W, D, PRO_NORM, EPI_GLU_SILU, LL>;
        case EPI_GLU_GELU: return lmrs_q_matvec_kernel<QT, W, D, PRO_NORM, EPI_GLU_GELU, LL>;
        case EPI_LOGITS: return lmrs_q_matvec_kernel<QT, W, D, PRO_NORM, EPI_LOGITS, LL>;
        default: return nullptr;
    }
}
template <int QT> static gemv_fn gemv_kernel_for(int cfg, int pro, int epi, bool ll) {
    if (cfg == 1) return ll ? gemv_kernel_pe<QT, 8, 3, true>(pro, epi) : gemv_kernel_pe<QT, 8, 3, false>(pro, epi);
    if (cfg == 2) return ll ? gemv_kernel_pe<QT, 8, 4, true>(pro, epi) : gemv_kernel_pe<QT, 8, 4, false>(pro, epi);
    if (cfg == 3) return ll ? gemv_kernel_pe<QT, 16, 2, true>(pro, epi) : gemv_kernel_pe<QT, 16, 2, false>(pro, epi);
    if (cfg == 4) return ll ? gemv_kernel_pe<QT, 16, 3, true>(pro, epi) : gemv_kernel_pe<QT, 16, 3, false>(pro, epi);
    return ll ? gemv_kernel_pe<QT, 8, 2, true>(pro, epi) : gemv_kernel_pe<QT, 8, 2, false>(pro, epi);
}
template <int QT> static size_t gemv_smem_for(int cfg, int n, bool norm) {
    switch (cfg) {
        case 1: return gemv_smem_bytes<QT, 8, 3>(n, norm);
        case 2: return gemv_smem_bytes<QT, 8, 4>(n, norm);
        case 3: return gemv_smem_bytes<QT, 16, 2>(n, norm);
        case 4: return gemv_smem_bytes<QT, 16, 3>(n, norm);
        default: return gemv_smem_bytes<QT, 8, 2>(n, norm);
    }
}
// stream boundaries must fall on BP16 block boundaries: row0 * G must be a multiple of 16 groups
static int gran_for(int n) {
    int G = n / GS;
    for (int r = 1; r <= 16; r *= 2)
        if ((r * G) % SG == 0) return r;
    return 16;
}
static size_t bp16_bytes(int q_type, size_t o, size_t n) {
    size_t groups = o * (n / GS), blocks = (groups + SG - 1) / SG;
    return blocks * (q_type == 1 ? blk_bytes<1>() : blk_bytes<2>());
}
static cudaError_t repack_bp16(int q_type, uint8_t* dst, const uint8_t* src_q, const float* src_s, size_t o, size_t n, cudaStream_t st) {
    size_t groups = o * (n / GS), blocks = (groups + SG - 1) / SG;
    if (q_type == 1) repack_bp16_kernel<1><<<(unsigned)blocks, 128, 0, st>>>(dst, src_q, src_s, groups);
    else repack_bp16_kernel<2><<<(unsigned)blocks, 128, 0, st>>>(dst, src_q, src_s, groups);
    return cudaGetLastError();
}

static cudaError_t launch_gemv(lmrs_b200* m, int q_type, GemvParams p) {
    static const int cfg_long = env_int("LMRS_B200_GEMV_CFG_LONG", -1);   // developer knob: geometry of the down projection
    const int cfg = (cfg_long >= 0 && cfg_long < N_GEMV_CFG && p.pro == PRO_QUANT && p.n >= 4096) ? cfg_long : m->gemv_cfg;
    const GemvCfg c = kGemvCfgs[cfg];
    gemv_fn fn = q_type == 1 ? gemv_kernel_for<1>(cfg, p.pro, p.epi, p.ll != 0) : gemv_kernel_for<2>(cfg, p.pro, p.epi, p.ll != 0);
    if (!fn) return cudaErrorInvalidValue;
    size_t smem = q_type == 1 ? gemv_smem_for<1>(cfg, p.n, p.pro == PRO_NORM) : gemv_smem_for<2>(cfg, p.n, p.pro == PRO_NORM);
    static const int pre_stages = env_int("LMRS_B200_PRE", 99);   // developer knob: ring stages requested before the dependency wait
    p.pre_stages = pre_stages;
    static const int pad_kb = env_int("LMRS_B200_GEMV_PAD_KB", 0);   // developer knob: unused shared memory (limits co-residency)
    smem += (size_t)pad_kb * 1024;
    cudaError_t se = smem_optin(m, (const void*)fn, smem);
    if (se != cudaSuccess) return se;
    int grid = m->sms * m->gemv_ctas_per_sm;
    // never launch more CTAs than there are row units to hand out (tiny matrices)
    const bool glu = p.epi == EPI_GLU_SILU || p.epi == EPI_GLU_GELU;
    (void)glu;
    int units = (p.o + p.row_gran - 1) / p.row_gran;
    if (grid > units) grid = units < 1 ? 1 : units;   // at least one row unit per CTA; otherwise every SM takes part
    return launch(m, fn, dim3(grid), dim3(c.warps * 32), smem, p);
}

static GemvParams gemv_base(const Mat& a, const Mat* b) {
    GemvParams p{};
    p.wq_a = a.q; p.ws_a = a.s;
    if (b) { p.wq_b = b->q; p.ws_b = b->s; }
    p.n = a.n; p.o = a.o; p.row_gran = a.gran;
    return p;
}

template <int HS, bool BIG> static cudaError_t launch_attn_grid_hs(lmrs_b200* m, const AttnParams& p0, int n_kv_heads, int rows) {
    {
        cudaError_t e = smem_optin(m, (const void*)attn_decode_kernel<HS, BIG>, attn_smem_bytes<HS, BIG>());
        if (e != cudaSuccess) return e;
    }
    AttnParams p = p0;
    const bool split_scores = !p.ll && env_int("LMRS_B200_ATT_SPLIT", rows > 1 ? 1 : 0) != 0;   // decode: measured no gain (extra launch)
    if (split_scores) {
        // the independent q.k dot products cover the whole GPU (decode: position splits; prefill: token rows) ...
        int nsplit = 1;
        if (rows == 1) { nsplit = m->sms / (n_kv_heads * p.chunks); if (nsplit < 1) nsplit = 1; if (nsplit > 32) nsplit = 32; }
        cudaError_t e = launch(m, attn_scores_kernel<HS>, dim3(n_kv_heads * p.chunks, nsplit, rows), dim3(ATTS_THREADS), 0, p, nsplit);
        if (e != cudaSuccess) return e;
        p.scores_ready = 1;   // ... and the serial softmax / a*v chains run per kv head
    }
    return launch(m, attn_decode_kernel<HS, BIG>, dim3(n_kv_heads * p.chunks, rows), dim3(ATT_THREADS), attn_smem_bytes<HS, BIG>(), p);
}
template <int HS> static cudaError_t launch_attn_grid_t(lmrs_b200* m, const AttnParams& p, int n_kv_heads, int rows) {
    return rows == 1 ? launch_attn_grid_hs<HS, true>(m, p, n_kv_heads, rows) : launch_attn_grid_hs<HS, false>(m, p, n_kv_heads, rows);
}
static cudaError_t launch_attn_grid(lmrs_b200* m, const AttnParams& p, int n_kv_heads, int rows) {
    switch (m->args.head_size) {
        case 64: return launch_attn_grid_t<64>(m, p, n_kv_heads, rows);
        case 96: return launch_attn_grid_t<96>(m, p, n_kv_heads, rows);
        case 128: return launch_attn_grid_t<128>(m, p, n_kv_heads, rows);
        case 256: return launch_attn_grid_t<256>(m, p, n_kv_heads, rows);
        default: return cudaErrorInvalidValue;
    }
}
static cudaError_t launch_attn(lmrs_b200* m, const AttnParams& p, int n_kv_heads) { return launch_attn_grid(m, p, n_kv_heads, 1); }

// decode attention on thread-block clusters (attention.cuh: attn_cluster_kernel), one cluster per KV head
constexpr int ATT_LEGACY = 7;   // variant index of the single-CTA kernel (any context length)
template <int HS, int CL, int G> static cudaError_t launch_attn_cluster_t(lmrs_b200* m, const AttnParams& p, int n_kv_heads, int cap) {
    if constexpr (CL % G != 0 || HS % (4 * (CL / G)) != 0) {
        return cudaErrorInvalidValue;
    } else {
        const int nh = std::min<int>(ATT_QH, p.kv_mul);
        const size_t smem = attc_smem_floats(HS, cap, CL, G, nh) * 4;
        {
            cudaError_t e = smem_optin(m, (const void*)attn_cluster_kernel<HS, CL, G>, smem);
            if (e != cudaSuccess) return e;
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(n_kv_heads * p.chunks * CL));
        cfg.blockDim = dim3(ATT_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = m->stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = m->use_pdl ? 2 : 1;
        m->launches++;
        return cudaLaunchKernelEx(&cfg, attn_cluster_kernel<HS, CL, G>, p, cap);
    }
}
template <int HS, int CL> static cudaError_t launch_attn_cluster_g(lmrs_b200* m, const AttnParams& p, int n_kv_heads, int cap, int g) {
    switch (g) {
        case 4: return launch_attn_cluster_t<HS, CL, 4>(m, p, n_kv_heads, cap);
        case 2: return launch_attn_cluster_t<HS, CL, 2>(m, p, n_kv_heads, cap);
        case 1: return launch_attn_cluster_t<HS, CL, 1>(m, p, n_kv_heads, cap);
        default: return cudaErrorInvalidValue;
    }
}
template <int HS> static cudaError_t launch_attn_cluster_hs(lmrs_b200* m, const AttnParams& p, int n_kv_heads, int cap, int g) {
    switch (m->att_cl) {
        case 8: return launch_attn_cluster_g<HS, 8>(m, p, n_kv_heads, cap, g);
        case 4: return launch_attn_cluster_g<HS, 4>(m, p, n_kv_heads, cap, g);
        default: return cudaErrorInvalidValue;
    }
}
static cudaError_t launch_attn_cluster(lmrs_b200* m, const AttnParams& p, int n_kv_heads, int cap, int g) {
    switch (m->args.head_size) {
        case 64: return launch_attn_cluster_hs<64>(m, p, n_kv_heads, cap, g);
        case 96: return launch_attn_cluster_hs<96>(m, p, n_kv_heads, cap, g);
        case 128: return launch_attn_cluster_hs<128>(m, p, n_kv_heads, cap, g);
        case 256: return launch_attn_cluster_hs<256>(m, p, n_kv_heads, cap, g);
        default: return cudaErrorInvalidValue;
    }
}
// Variants of the cluster kernel, one CUDA graph each: position buckets (short contexts need little shared memory, so
// the neighbouring GEMVs' CTAs co-reside and prefetch) in the head-group layout, then -- when heads are grouped -- one
// more bucket in the ungrouped layout, whose smaller V slices reach longer contexts within ~200 KB of shared memory.
static void setup_attn_cluster(lmrs_b200* m) {
    const int hs = (int)m->args.head_size;
    const int kv_mul = (int)(m->args.n_heads / m->args.n_kv_heads);
    const int nh = std::min<int>(ATT_QH, kv_mul);
    int cl = env_int("LMRS_B200_ATT_CLUSTER", ATTC_MAX_CL);
    cl = cl >= 8 ? 8 : (cl >= 4 ? 4 : 0);
    m->att_cl = cl;
    m->att_nvar = 0;
    if (cl == 0) return;
    // every chunk of query heads (4, and kv_mul % 4 for the last one) must split evenly into the groups
    int g = kv_mul % 4 == 0 ? 4 : (kv_mul % 4 == 2 ? 2 : 1);
    g = std::min(g, env_int("LMRS_B200_ATT_GROUPS", 4));
    while (g > 1 && (cl % g != 0 || hs % (4 * (cl / g)) != 0)) g /= 2;
    auto max_cap = [&](int gg) {
        int cap = 0;
        while (cap + 32 <= ATT_SC_CAP && attc_smem_floats(hs, cap + 32, cl, gg, nh) * 4 <= (size_t)200 * 1024) cap += 32;
        return cap;
    };
    const int cap_g = max_cap(g);
    if (cap_g < 64) { m->att_cl = 0; return; }
    const int num[4] = {5, 8, 12, 16};
    int prev = 0;
    for (int b = 0; b < 4; b++) {
        const int cap = std::min(cap_g, ((cap_g * num[b] / 16) + 31) & ~31);
        if (cap > prev) { m->att_var[m->att_nvar].cap = cap; m->att_var[m->att_nvar].g = g; m->att_nvar++; prev = cap; }
    }
    if (g > 1) {
        const int cap_1 = max_cap(1);
        if (cap_1 > prev) { m->att_var[m->att_nvar].cap = cap_1; m->att_var[m->att_nvar].g = 1; m->att_nvar++; }
    }
}
static int attn_variant_for(const lmrs_b200* m, uint32_t pos) {
    if (m->att_cl == 0) return ATT_LEGACY;
    for (int b = 0; b < m->att_nvar; b++)
        if ((int)pos + 1 <= m->att_var[b].cap) return b;
    return ATT_LEGACY;
}

// ---- TMA descriptors (driver entry point resolved at run time: no link-time dependency on libcuda) ---------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// [rows][row_bytes] byte matrix, box 128 bytes x 128 rows, 128B swizzle (the K-major UMMA operand layout)
static int make_tmap_2d(CUtensorMap* map, const void* base, size_t row_bytes, size_t rows, int box_rows = 128) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail("cuTensorMapEncodeTiled is not available from this driver");
    cuuint64_t gdim[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {128, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return 0;
}
static bool gemm_shape_ok(int n, int o) { return n % 128 == 0 && o % 128 == 0; }
// out[t][i] for t < T: tcgen05 int8 GEMM of xq [T][n] (with xs [T][n/128]) against the dense copy of w.  Tile shape: 128
// output rows of w per CTA, or 64 when that is what it takes to give every SM a tile (Wo / W2 of the 1B..3B models).
static int launch_gemm(lmrs_b200* m, const Mat& w, const uint8_t* xq, const float* xs, int T, GemmParams gp) {
    CUtensorMap ta;
    if (make_tmap_2d(&ta, xq, (size_t)w.n, (size_t)T)) return 1;
    gp.T = T; gp.n = w.n; gp.o = w.o; gp.ws = w.ds; gp.xs = xs; gp.neg_zero = -0.0f;
    const int mt = (T + GEMM_M - 1) / GEMM_M;
    static const int force_bn = env_int("LMRS_B200_GEMM_BN", 0);   // developer knob
    // 64-column tiles also where the q | k | v output segments end on a 64- but not a 128-column boundary (KV heads of 64 on 8 GPUs)
    const bool narrow = (gp.c1 % 128) || (gp.c2 % 128) || (force_bn ? force_bn == 64 : (w.o / 128) * mt < m->sms);
    if (narrow) {
        const size_t smem = gemm_smem_bytes<64>(w.n);
        CK(smem_optin(m, (const void*)gemm_q8_kernel<64, false>, smem));
        gemm_q8_kernel<64, false><<<dim3(w.o / 64, mt), gemm_threads<64>(), smem, m->stream>>>(ta, w.tmap64, w.tmap64, gp);
    } else {
        const size_t smem = gemm_smem_bytes<128>(w.n);
        CK(smem_optin(m, (const void*)gemm_q8_kernel<128, false>, smem));
        gemm_q8_kernel<128, false><<<dim3(w.o / 128, mt), gemm_threads<128>(), smem, m->stream>>>(ta, w.tmap, w.tmap, gp);
    }
    m->launches++;
    CK(cudaGetLastError());
    return 0;
}
// h[t][i] = act(w1[i] . x[t]) * (w3[i] . x[t]): gate and up in ONE launch, each CTA tile = 64 rows of w1 + the same 64 rows of w3
static int launch_gemm_glu(lmrs_b200* m, const Mat& w1, const Mat& w3, const uint8_t* xq, const float* xs, int T, float* h, int ld, int epi) {
    CUtensorMap ta;
    if (make_tmap_2d(&ta, xq, (size_t)w1.n, (size_t)T)) return 1;
    GemmParams gp{};
    gp.T = T; gp.n = w1.n; gp.o = w1.o; gp.ws = w1.ds; gp.ws2 = w3.ds; gp.xs = xs; gp.out0 = h; gp.ld0 = ld; gp.glu_epi = epi; gp.neg_zero = -0.0f;
    const size_t smem = gemm_smem_bytes<128>(w1.n);
    CK(smem_optin(m, (const void*)gemm_q8_kernel<128, true>, smem));
    gemm_q8_kernel<128, true><<<dim3(w1.o / 64, (T + GEMM_M - 1) / GEMM_M), gemm_threads<128>(), smem, m->stream>>>(ta, w1.tmap64, w3.tmap64, gp);
    m->launches++;
    CK(cudaGetLastError());
    return 0;
}
static GemmParams gemm_out1(float* out, int ld, int o) {
    GemmParams g{};
    g.out0 = out; g.ld0 = ld; g.c1 = o; g.out1 = out; g.ld1 = ld; g.c2 = o; g.out2 = out; g.ld2 = ld;
    return g;
}

// ---- RoPE frequency (src/transformer.rs:445-478), evaluated on the host with libm like the reference -------
static void rope_freq(int model_type, float rope_theta, int head_size, int j, float* freq_out, float* mscale) {
    static const double short_factor[48] = {
        1.08, 1.1, 1.1300000000000001, 1.2800000000000002, 1.3100000000000003, 1.4500000000000004,
        1.4500000000000004, 1.9500000000000008, 2.030000000000001, 2.4299999999999926, 2.5699999999999896,
        2.9499999999999815, 3.729999999999965, 3.869999999999962, 4.189999999999955, 4.43999999999995,
        4.6399999999999455, 4.979999999999938, 5.159999999999934, 5.279999999999932, 5.759999999999922,
        5.889999999999919, 5.889999999999919, 5.969999999999917, 6.089999999999915, 6.2799999999999105,
        6.7699999999999, 6.8899999999998975, 7.109999999999893, 7.129999999999892, 7.179999999999891,
        7.289999999999889, 7.339999999999888, 7.559999999999883, 7.619999999999882, 7.69999999999988,
        7.879999999999876, 7.879999999999876, 7.879999999999876, 7.939999999999875, 7.949999999999875,
        7.979999999999874, 8.19999999999987, 8.439999999999864, 8.469999999999864, 8.589999999999861,
        8.809999999999857, 8.999999999999853};
    volatile float fj = (float)(2 * j) / (float)head_size;
    volatile float freq = 1.0f / powf(rope_theta, fj);
    float scaling = 1.0f;
    if (model_type == 1) {  // LLAMA: llama-3 scaling constants hard-coded in the reference (:451-470)
        volatile float wavelen = (2.0f * 3.14159265358979323846f) / freq;
        const float factor = 32.0f, low = 1.0f, high = 4.0f, old_ctx = 8192.0f;
        volatile float low_wl = old_ctx / low, high_wl = old_ctx / high;
        if (wavelen > low_wl) {
            freq = freq / factor;
        } else if (wavelen <= low_wl && wavelen >= high_wl) {
            volatile float a = old_ctx / wavelen;
            volatile float smooth = (a - low) / (high - low);
            volatile float t1 = (1.0f - smooth) * freq;
            volatile float t2 = t1 / factor;
            volatile float t3 = smooth * freq;
            freq = t2 + t3;
        }
    }
    if (model_type == 2) {  // PHI (:472-478)
        volatile float inv = (float)(1.0 / short_factor[j % 48]);
        freq = freq * inv;
        volatile float scale = 131072.0f / 4096.0f;
        volatile float r = logf(scale) / logf(4096.0f);
        scaling = sqrtf(1.0f + r);
    }
    *freq_out = freq;
    *mscale = scaling;
}

// ---- LL exchange buffers: per layer one word array per activation that crosses a kernel boundary ------------------
struct LLLayout { size_t xo1, q, k_new, v_new, att, wo_out, xo0, h, down_out, per_layer; };
static LLLayout ll_layout(const lmrs_b200* m) {
    LLLayout L{};
    size_t cur = 0;
    auto take = [&](size_t n) { size_t o = cur; cur += (n + 15) / 16 * 16; return o; };   // 128-byte aligned slices
    L.xo1 = take(m->args.dim); L.q = take(m->l_att_dim); L.k_new = take(m->l_kv_dim); L.v_new = take(m->l_kv_dim);
    L.att = take(m->l_att_dim); L.wo_out = take(m->args.dim); L.xo0 = take(m->args.dim); L.h = take(m->l_hidden);
    L.down_out = take(m->args.dim);
    L.per_layer = cur;
    return L;
}

// ---- loader ----------------------------------------------------------------------------------------------
struct FileTensor { size_t q_off, s_off, q_bytes, s_bytes; };   // one quantized tensor of one layer

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int build_model(lmrs_b200* m, const uint8_t* file, size_t len, size_t* end_offset) {
    if (len < 256 || file[0] != 0x6c || file[1] != 0x6d || file[2] != 0x72 || file[3] != 0x73)
        return fail("Model not in lm.rs format.");  // src/transformer.rs:135
    memcpy(&m->args, file + 8, sizeof(lmrs_args_t));
    lmrs_args_t& a = m->args;
    if (a.seq_len > 8192) a.seq_len = 8192;  // :158-160
    if (a.q_type != 1 && a.q_type != 2)
        return fail("lmrs_b200: only Q8_0 / Q4_0 models are supported on the B200 path (q_type=" + std::to_string(a.q_type) + ")");
    if (a.model_type > 2) return fail("unknown model_type");
    if (!a.dim || !a.hidden_dim || !a.n_layers || !a.n_heads || !a.n_kv_heads || !a.head_size || !a.vocab_size || !a.seq_len)
        return fail("malformed header: zero-sized model dimension");
    if (a.group_size != GS) return fail("lmrs_b200: group_size must be 128 (the exporter always quantizes with 128, utils/io.py:21)");
    const size_t dim = a.dim, hd = a.hidden_dim, L = a.n_layers, hs = a.head_size;
    const size_t att_dim = (size_t)a.n_heads * hs, kv_dim = (size_t)a.n_kv_heads * hs;
    if (dim % GS || hd % GS || att_dim % GS) return fail("dim/hidden_dim/att_dim must be multiples of the group size");
    if (hs != 64 && hs != 96 && hs != 128 && hs != 256) return fail("unsupported head_size (64/96/128/256)");
    if (a.n_heads % a.n_kv_heads) return fail("n_heads must be a multiple of n_kv_heads");
    if (a.model_type == 2 && hs != 96) return fail("PHI requires head_size 96 (48 rope short factors, src/transformer.rs:473)");
    if (a.vocab_size % 4 || kv_dim % 4) return fail("row counts must be multiples of 4 (src/functional.rs:179)");
    const int W = m->world, R = m->rank;
    if (a.n_kv_heads % W || (hd / GS) % W || (a.vocab_size / 4) % W)
        return fail("world size must divide n_kv_heads, hidden_dim/128 and vocab_size/4");
    m->l_kv_heads = a.n_kv_heads / W;
    m->l_heads = a.n_heads / W;
    m->l_att_dim = m->l_heads * hs;
    m->l_kv_dim = m->l_kv_heads * hs;
    m->l_hidden = hd / W;
    m->l_vocab = a.vocab_size / W;
    m->vocab_off = R * m->l_vocab;
    if (W > 1 && (m->l_att_dim % GS)) return fail("sharded att_dim must stay a multiple of 128");
    const int qdiv = a.q_type == 2 ? 2 : 1;

    // walk the file exactly like src/transformer.rs:241-270
    size_t off = 256;
    auto take_q = [&](size_t n_layers, size_t elems, std::vector<FileTensor>& out) -> bool {
        for (size_t l = 0; l < n_layers; l++) {
            FileTensor t;
            t.q_off = off; t.q_bytes = elems / qdiv; off += t.q_bytes;
            t.s_off = off; t.s_bytes = elems / GS * 4; off += t.s_bytes;
            if (off > len) return false;
            out.push_back(t);
        }
        return true;
    };
    auto take_f = [&](size_t n_layers, size_t elems, std::vector<size_t>& out) -> bool {
        for (size_t l = 0; l < n_layers; l++) { out.push_back(off); off += elems * 4; if (off > len) return false; }
        return true;
    };
    std::vector<FileTensor> f_emb, f_wq, f_wk, f_wv, f_wo, f_w1, f_w2, f_w3, f_head;
    std::vector<size_t> f_rms_att, f_rms_post, f_rms_pre, f_rms_postffn, f_rms_final;
    bool ok = take_q(1, (size_t)a.vocab_size * dim, f_emb) && take_f(L, dim, f_rms_att) &&
              take_q(L, dim * att_dim, f_wq) && take_q(L, dim * kv_dim, f_wk) && take_q(L, dim * kv_dim, f_wv) &&
              take_q(L, dim * att_dim, f_wo) && take_f(L, dim, f_rms_post);
    if (ok && a.model_type == 0) ok = take_f(L, dim, f_rms_pre);
    ok = ok && take_q(L, dim * hd, f_w1) && take_q(L, dim * hd, f_w2) && take_q(L, dim * hd, f_w3);
    if (ok && a.model_type == 0) ok = take_f(L, dim, f_rms_postffn);
    ok = ok && take_f(1, dim, f_rms_final);
    if (ok && a.model_type == 2) ok = take_q(1, dim * a.vocab_size, f_head);
    if (!ok) return fail("file truncated");
    if (end_offset) *end_offset = off;

    // ---- plan the HBM arena: everything 256-byte aligned, QKV rows concatenated per layer --------------------
    struct Piece { size_t dst; const uint8_t* src; size_t bytes; size_t src_stride, dst_stride, rows; };
    std::vector<Piece> pieces;
    size_t cur = 0;
    auto place = [&](size_t bytes) { size_t o = cur; cur = align_up(cur + bytes, 256); return o; };
    // rows [r0, r0+nr) of a row-major [o][n] quantized tensor (contiguous)
    auto add_rows = [&](size_t dst_q, size_t dst_s, const FileTensor& t, size_t n, size_t r0, size_t nr) {
        pieces.push_back({dst_q, file + t.q_off + r0 * (n / qdiv), nr * (n / qdiv), 0, 0, 1});
        pieces.push_back({dst_s, file + t.s_off + r0 * (n / GS) * 4, nr * (n / GS) * 4, 0, 0, 1});
    };
    // columns [c0, c0+nc) of every row (K-shard for wo / w2): strided gather
    auto add_cols = [&](size_t dst_q, size_t dst_s, const FileTensor& t, size_t n, size_t o, size_t c0, size_t nc) {
        pieces.push_back({dst_q, file + t.q_off + c0 / qdiv, nc / qdiv, n / qdiv, nc / qdiv, o});
        pieces.push_back({dst_s, file + t.s_off + (c0 / GS) * 4, (nc / GS) * 4, (n / GS) * 4, (nc / GS) * 4, o});
    };
    struct MatPlan { size_t q, s; int o, n; };
    auto plan_mat = [&](int o, int n) { MatPlan p; p.o = o; p.n = n; p.q = place((size_t)o * n / qdiv); p.s = place((size_t)o * n / GS * 4); return p; };
    auto plan_vec = [&](size_t file_off) { size_t d = place(dim * 4); pieces.push_back({d, file + file_off, dim * 4, 0, 0, 1}); return d; };

    struct LayerPlan { MatPlan qkv, wo, w1, w3, w2; size_t rms_att, rms_post, rms_pre, rms_postffn; };
    std::vector<LayerPlan> lp(L);
    const size_t la = m->l_att_dim, lk = m->l_kv_dim, lh = m->l_hidden;
    for (size_t l = 0; l < L; l++) {
        LayerPlan& P = lp[l];
        P.qkv = plan_mat((int)(la + 2 * lk), (int)dim);
        add_rows(P.qkv.q, P.qkv.s, f_wq[l], dim, R * la, la);
        add_rows(P.qkv.q + la * dim / qdiv, P.qkv.s + la * (dim / GS) * 4, f_wk[l], dim, R * lk, lk);
        add_rows(P.qkv.q + (la + lk) * dim / qdiv, P.qkv.s + (la + lk) * (dim / GS) * 4, f_wv[l], dim, R * lk, lk);
        P.wo = plan_mat((int)dim, (int)la);
        if (W == 1) add_rows(P.wo.q, P.wo.s, f_wo[l], att_dim, 0, dim); else add_cols(P.wo.q, P.wo.s, f_wo[l], att_dim, dim, R * la, la);
        P.w1 = plan_mat((int)lh, (int)dim);
        add_rows(P.w1.q, P.w1.s, f_w1[l], dim, R * lh, lh);
        P.w3 = plan_mat((int)lh, (int)dim);
        add_rows(P.w3.q, P.w3.s, f_w3[l], dim, R * lh, lh);
        P.w2 = plan_mat((int)dim, (int)lh);
        if (W == 1) add_rows(P.w2.q, P.w2.s, f_w2[l], hd, 0, dim); else add_cols(P.w2.q, P.w2.s, f_w2[l], hd, dim, R * lh, lh);
        P.rms_att = plan_vec(f_rms_att[l]);
        P.rms_post = plan_vec(f_rms_post[l]);
        P.rms_pre = a.model_type == 0 ? plan_vec(f_rms_pre[l]) : 0;
        P.rms_postffn = a.model_type == 0 ? plan_vec(f_rms_postffn[l]) : 0;
    }
    MatPlan p_emb = plan_mat((int)a.vocab_size, (int)dim);   // full table on every rank (row gather)
    add_rows(p_emb.q, p_emb.s, f_emb[0], dim, 0, a.vocab_size);
    MatPlan p_cls = p_emb;
    if (a.model_type == 2) {
        p_cls = plan_mat(m->l_vocab, (int)dim);
        add_rows(p_cls.q, p_cls.s, f_head[0], dim, m->vocab_off, m->l_vocab);
    } else if (W > 1) {  // tied classifier: this rank's vocab rows are a sub-range of the table
        p_cls.o = m->l_vocab;
        p_cls.q = p_emb.q + (size_t)m->vocab_off * dim / qdiv;
        p_cls.s = p_emb.s + (size_t)m->vocab_off * (dim / GS) * 4;
    }
    size_t rms_final = plan_vec(f_rms_final[0]);
    // ---- upload in file layout to a staging buffer, then repack every matrix into the BP16 arena ------------------
    uint8_t* d_stage = nullptr;
    CK(cudaMalloc(&d_stage, cur));
    m->d_dense = d_stage;   // owned by the handle from here on: an error return below is cleaned up by lmrs_b200_destroy
    for (const Piece& pc : pieces) {
        if (pc.rows <= 1) CK(cudaMemcpy(d_stage + pc.dst, pc.src, pc.bytes, cudaMemcpyHostToDevice));
        else CK(cudaMemcpy2D(d_stage + pc.dst, pc.dst_stride, pc.src, pc.src_stride, pc.bytes, pc.rows, cudaMemcpyHostToDevice));
    }
    size_t pcur = 0;
    auto pplace = [&](size_t bytes) { size_t o = pcur; pcur = align_up(pcur + bytes, 256); return o; };
    struct PackJob { size_t dst; MatPlan src; };
    std::vector<PackJob> jobs;
    struct VecJob { size_t dst, src; };
    std::vector<VecJob> vjobs;
    auto pk = [&](const MatPlan& p) { size_t d = pplace(bp16_bytes(a.q_type, p.o, p.n)); jobs.push_back({d, p}); return d; };
    auto pv = [&](size_t src) { size_t d = pplace(dim * 4); vjobs.push_back({d, src}); return d; };
    struct LayerPk { size_t qkv, wo, w1, w3, w2, rms_att, rms_post, rms_pre, rms_postffn; };
    std::vector<LayerPk> lpk(L);
    for (size_t l = 0; l < L; l++) {
        lpk[l].qkv = pk(lp[l].qkv); lpk[l].wo = pk(lp[l].wo); lpk[l].w1 = pk(lp[l].w1); lpk[l].w3 = pk(lp[l].w3); lpk[l].w2 = pk(lp[l].w2);
        lpk[l].rms_att = pv(lp[l].rms_att); lpk[l].rms_post = pv(lp[l].rms_post);
        lpk[l].rms_pre = a.model_type == 0 ? pv(lp[l].rms_pre) : 0;
        lpk[l].rms_postffn = a.model_type == 0 ? pv(lp[l].rms_postffn) : 0;
    }
    const size_t emb_pk = pk(p_emb);
    size_t cls_pk = emb_pk;
    if (a.model_type == 2) cls_pk = pk(p_cls);
    else if (W > 1) {   // tied classifier: this rank's vocab rows are a block-aligned sub-range of the packed table
        if (((size_t)m->vocab_off * (dim / GS)) % SG) return fail("vocab shard is not BP16 block aligned");
        cls_pk = emb_pk + ((size_t)m->vocab_off * (dim / GS) / SG) * (a.q_type == 1 ? blk_bytes<1>() : blk_bytes<2>());
    }
    const size_t rms_final_pk = pv(rms_final);
    m->arena_bytes = pcur;
    CK(cudaMalloc(&m->d_arena, m->arena_bytes));
    for (const PackJob& j : jobs)
        CK(repack_bp16(a.q_type, m->d_arena + j.dst, d_stage + j.src.q, (const float*)(d_stage + j.src.s), j.src.o, j.src.n, 0));
    for (const VecJob& j : vjobs) CK(cudaMemcpy(m->d_arena + j.dst, d_stage + j.src, dim * 4, cudaMemcpyDeviceToDevice));
    CK(cudaDeviceSynchronize());
    m->use_gemm = a.q_type == 1 && env_int("LMRS_B200_GEMM", 1) != 0 && (W == 1 || m->use_peer);   // N-GPU: needs the peer exchange
    if (!m->use_gemm) { cudaFree(d_stage); m->d_dense = nullptr; }   // otherwise the file-layout copy doubles as the GEMM's B operand
    int tmap_err = 0;
    auto mkd = [&](Mat& x, const MatPlan& pl) {   // dense views + TMA descriptor
        if (!m->use_gemm) return;
        x.dq = m->d_dense + pl.q; x.ds = (const float*)(m->d_dense + pl.s);
        if (gemm_shape_ok(pl.n, pl.o)) {
            if (make_tmap_2d(&x.tmap, x.dq, (size_t)pl.n, (size_t)pl.o) || make_tmap_2d(&x.tmap64, x.dq, (size_t)pl.n, (size_t)pl.o, 64)) tmap_err = 1;
            else x.has_tmap = true;
        }
    };
    auto mk = [&](size_t off, int o, int n) { Mat x; x.q = m->d_arena + off; x.s = nullptr; x.o = o; x.n = n; x.gran = gran_for(n); return x; };
    auto fp = [&](size_t o) { return (const float*)(m->d_arena + o); };
    m->layers.resize(L);
    for (size_t l = 0; l < L; l++) {
        Layer& Y = m->layers[l];
        Y.qkv = mk(lpk[l].qkv, lp[l].qkv.o, lp[l].qkv.n); Y.wo = mk(lpk[l].wo, lp[l].wo.o, lp[l].wo.n);
        Y.w1 = mk(lpk[l].w1, lp[l].w1.o, lp[l].w1.n); Y.w3 = mk(lpk[l].w3, lp[l].w3.o, lp[l].w3.n); Y.w2 = mk(lpk[l].w2, lp[l].w2.o, lp[l].w2.n);
        Y.rms_att = fp(lpk[l].rms_att); Y.rms_post_att = fp(lpk[l].rms_post);
        if (a.model_type == 0) { Y.rms_pre_ffn = fp(lpk[l].rms_pre); Y.rms_post_ffn = fp(lpk[l].rms_postffn); }
        mkd(Y.qkv, lp[l].qkv); mkd(Y.wo, lp[l].wo); mkd(Y.w1, lp[l].w1); mkd(Y.w3, lp[l].w3); mkd(Y.w2, lp[l].w2);
    }
    if (tmap_err) m->use_gemm = false;
    m->emb = mk(emb_pk, p_emb.o, p_emb.n);
    m->cls = mk(cls_pk, p_cls.o, p_cls.n);
    m->rms_final = fp(rms_final_pk);

    // ---- state: f32 KV cache (src/transformer.rs:302-303), RoPE tables, activations ---------------------------
    const size_t kv_elems = L * (size_t)a.seq_len * lk;
    CK(cudaMalloc(&m->d_kcache, kv_elems * 4));
    CK(cudaMalloc(&m->d_vcache, kv_elems * 4));
    CK(cudaMemset(m->d_kcache, 0, kv_elems * 4));
    CK(cudaMemset(m->d_vcache, 0, kv_elems * 4));
    {
        std::vector<float> cs((size_t)a.seq_len * (hs / 2)), sn((size_t)a.seq_len * (hs / 2));
        for (size_t j = 0; j < hs / 2; j++) {
            float freq, ms;
            rope_freq(a.model_type, a.rope_theta, (int)hs, (int)j, &freq, &ms);
            for (size_t p = 0; p < a.seq_len; p++) {
                volatile float val = (float)(uint32_t)p * freq;       // :480
                volatile float c = cosf(val), s = sinf(val);
                cs[p * (hs / 2) + j] = c * ms;                        // :481-482
                sn[p * (hs / 2) + j] = s * ms;
            }
        }
        CK(cudaMalloc(&m->d_rope_cos, cs.size() * 4));
        CK(cudaMalloc(&m->d_rope_sin, sn.size() * 4));
        CK(cudaMemcpy(m->d_rope_cos, cs.data(), cs.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(m->d_rope_sin, sn.data(), sn.size() * 4, cudaMemcpyHostToDevice));
    }
    m->att_chunks = (a.n_heads / a.n_kv_heads + ATT_QH - 1) / ATT_QH;
    CK(cudaMalloc(&m->d_x[0], dim * 4));
    CK(cudaMalloc(&m->d_x[1], dim * 4));
    CK(cudaMalloc(&m->d_q, la * 4));
    CK(cudaMalloc(&m->d_knew, lk * 4));
    CK(cudaMalloc(&m->d_att, la * 4));
    CK(cudaMalloc(&m->d_wo_out, dim * 4));
    CK(cudaMalloc(&m->d_h, lh * 4));
    CK(cudaMalloc(&m->d_down_out, dim * 4));
    if (!m->use_peer) CK(cudaMalloc(&m->d_logits, (size_t)a.vocab_size * 4));   // peer mode: lives in the exchange block (setup_peer_exchange)
    CK(cudaMalloc(&m->d_scores, (size_t)m->l_heads * align_up(a.seq_len, 4) * 4));
    {
        const LLLayout Y0 = ll_layout(m);
        m->ll_words = Y0.per_layer * L;
        CK(cudaMalloc(&m->d_ll, m->ll_words * sizeof(llword_t)));
        CK(cudaMemset(m->d_ll, 0, m->ll_words * sizeof(llword_t)));   // sequence number 0 = "never written"
        CK(cudaMalloc(&m->d_fin_scratch, dim * 4));
    }
    CK(cudaMalloc(&m->d_amax, 256 * 4)); CK(cudaMalloc(&m->d_aidx, 256 * 4));
    CK(cudaMalloc(&m->d_ticket, 4)); CK(cudaMemset(m->d_ticket, 0, 4));
    CK(cudaMalloc(&m->d_next, 4));
    CK(cudaMalloc(&m->d_step, sizeof(StepParams)));
    CK(cudaMallocHost(&m->h_step_ring, sizeof(StepParams) * 64));
    CK(cudaMallocHost(&m->h_logits, (size_t)a.vocab_size * 4));
    CK(cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking));
    m->stream = m->own_stream;
    return 0;
}

// ---- the step as a table of phases (src/transformer.rs:316-384 + :388-657 with sl = 1) ------------------------------
// `serial_prefill`: x comes from row `token` of the staged embeddings, no classifier, and a final phase writes the
// residual stream back (fill_kv_cache semantics).  `ll`: the activations between the phases are LL word buffers (one per
// phase output per layer) and the kernels never wait for a kernel boundary; otherwise plain f32 buffers reused by every
// layer and griddepcontrol.wait (row-sharded NCCL mode, LMRS_B200_LL=0).
static std::vector<Phase> make_phases(lmrs_b200* m, bool serial_prefill) {
    const lmrs_args_t& a = m->args;
    const bool gemma = a.model_type == 0;
    const bool ll = m->use_ll;
    const size_t L = a.n_layers;
    const LLLayout Y0 = ll_layout(m);
    std::vector<Phase> ph;
    const void* delta = nullptr;        // pending residual contribution
    const float* w_post = nullptr;      // Gemma: norm applied to it before the add
    const void* x_cur = ll ? nullptr : (const void*)m->d_x[0];   // residual stream before the pending add
    auto gemv_phase = [&](const Mat& A, const Mat* B) {
        Phase P;
        memset(&P, 0, sizeof P);
        P.kind = PH_GEMV;
        P.g = gemv_base(A, B);
        P.g.step = m->d_step;
        P.g.ll = ll;
        return P;
    };
    // L2 prefetch (LMRS_B200_L2PF): while a block's latency-bound middle runs, HBM is idle; the rest of the block's weights
    // (through the next block's [Wq;Wk;Wv]; after the last block: the first part of a tied classifier) form ONE contiguous
    // arena range (pack order of build_model) that idle CTAs ask the L2 to fetch.  1: the attention kernel's CTAs request
    // from Wo on; 2: the Wo kernel's CTAs (parked in their dependency wait while the attention runs) request from W1 on.
    auto l2pf_range = [&](size_t l, const uint8_t* lo, const uint8_t*& ptr, unsigned long long& bytes, int& chunk) {
        const Layer& Y = m->layers[l];
        const uint8_t* hi;
        if (l + 1 < L) hi = m->layers[l + 1].qkv.q + bp16_bytes(a.q_type, m->layers[l + 1].qkv.o, m->layers[l + 1].qkv.n);
        else {
            const size_t extra = (m->cls.q == m->emb.q && m->world == 1) ? ((size_t)m->l2pf_cls_mb << 20) : (size_t)65536;
            hi = std::min<const uint8_t*>(m->d_arena + m->arena_bytes, Y.w2.q + bp16_bytes(a.q_type, Y.w2.o, Y.w2.n) + extra);
        }
        if (hi > lo) { ptr = lo; bytes = ((unsigned long long)(hi - lo)) & ~15ull; chunk = m->l2pf_ef ? -m->l2pf_chunk : m->l2pf_chunk; }
    };
    const bool peer = m->use_peer;
    const llword_t* px_pending = nullptr;   // peer mode: the pending contribution lives in this GPU's exchange vector
    auto px_push = [&](GemvParams& g, size_t layer, int which) {   // the producer writes its slot on every GPU
        g.px_world = m->world;
        for (int r = 0; r < m->world; r++) g.px_out[r] = px_vec(m, r, layer, which) + (size_t)m->rank * a.dim;
    };
    auto px_take = [&](GemvParams& g) {                             // the consumer sums the slots in rank order
        if (!px_pending) return;
        g.px_world = m->world; g.px_in = px_pending; g.delta = px_pending;
    };
    for (size_t l = 0; l < L; l++) {
        const Layer& Y = m->layers[l];
        float* kc = m->d_kcache + l * (size_t)a.seq_len * m->l_kv_dim;
        float* vc = m->d_vcache + l * (size_t)a.seq_len * m->l_kv_dim;
        llword_t* lb = m->d_ll + l * Y0.per_layer;
        void* b_xo1 = ll ? (void*)(lb + Y0.xo1) : (void*)m->d_x[1];
        void* b_q = ll ? (void*)(lb + Y0.q) : (void*)m->d_q;
        void* b_knew = ll ? (void*)(lb + Y0.k_new) : (void*)m->d_knew;
        void* b_vnew = ll ? (void*)(lb + Y0.v_new) : (void*)vc;
        void* b_att = ll ? (void*)(lb + Y0.att) : (void*)m->d_att;
        void* b_wo = ll ? (void*)(lb + Y0.wo_out) : (void*)m->d_wo_out;
        void* b_xo0 = ll ? (void*)(lb + Y0.xo0) : (void*)m->d_x[0];
        void* b_h = ll ? (void*)(lb + Y0.h) : (void*)m->d_h;
        void* b_down = ll ? (void*)(lb + Y0.down_out) : (void*)m->d_down_out;
        {   // x(+delta) -> rmsnorm(w_rms_att) -> quantize -> [Wq;Wk;Wv]   (:409-431)
            Phase P = gemv_phase(Y.qkv, nullptr);
            GemvParams& p = P.g;
            p.pro = PRO_NORM; p.x_in = x_cur; p.delta = delta; p.w_post = w_post; p.w_norm = Y.rms_att;
            p.x_out = b_xo1; p.eps = a.rms_norm_eps; p.unit_offset = gemma;
            if (l == 0) {
                if (serial_prefill) { p.x_in = m->d_rows; p.x_in_stride = (int)a.dim; p.x_in_plain = 1; }
                else {   // embedding row (:324) with Gemma's sqrt(dim) scaling (:327-332) folded into the prologue
                    p.emb_q = m->emb.q; p.emb_s = m->emb.s; p.emb_qtype = a.q_type;
                    p.emb_apply_mul = gemma; p.emb_mul = sqrtf((float)a.dim);
                }
            }
            p.epi = EPI_QKV; p.out = b_q; p.out_k = b_knew; p.out_v = b_vnew;
            p.att_dim = m->l_att_dim; p.kv_dim = m->l_kv_dim;
            if (peer) px_take(p);
            ph.push_back(P);
        }
        {   // RoPE + attention (:443-544)
            Phase P;
            memset(&P, 0, sizeof P);
            P.kind = PH_ATTN;
            AttnParams& p = P.a;
            p.q = b_q; p.k_new = b_knew; p.v_new = ll ? b_vnew : nullptr; p.kcache = kc; p.vcache = vc;
            p.rope_cos = m->d_rope_cos; p.rope_sin = m->d_rope_sin; p.out = b_att; p.scores = m->d_scores;
            p.kv_dim = m->l_kv_dim; p.kv_mul = a.n_heads / a.n_kv_heads;
            p.chunks = m->att_chunks; p.gemma = gemma; p.seq_len = (int)align_up(a.seq_len, 4);
            p.sqrt_hs = sqrtf((float)a.head_size); p.step = m->d_step; p.ll = ll;
            if (m->l2pf == 1 && !serial_prefill) l2pf_range(l, Y.wo.q, p.l2pf_ptr, p.l2pf_bytes, p.l2pf_chunk);   // requested by the attention CTAs
            ph.push_back(P);
        }
        {   // quantize(att) -> Wo (:546-560)
            Phase P = gemv_phase(Y.wo, nullptr);
            P.g.pro = PRO_QUANT; P.g.act_in = b_att; P.g.epi = EPI_STORE; P.g.out = b_wo;
            if (m->l2pf == 2 && !serial_prefill && !ll) l2pf_range(l, Y.w1.q, P.g.l2pf_ptr, P.g.l2pf_bytes, P.g.l2pf_chunk);
            P.comm = peer ? 0 : 1;   // row-sharded mode: push the partial to every GPU / NCCL all-reduce of the output
            if (peer) { px_push(P.g, l, 0); px_pending = px_vec(m, m->rank, l, 0); }
            ph.push_back(P);
        }
        {   // x += wo_out (Gemma: normed) -> rmsnorm -> quantize -> gate/up -> act*up (:562-624)
            Phase P = gemv_phase(Y.w1, &Y.w3);
            GemvParams& p = P.g;
            p.pro = PRO_NORM; p.x_in = b_xo1; p.delta = b_wo; p.w_post = gemma ? Y.rms_post_att : nullptr;
            p.w_norm = gemma ? Y.rms_pre_ffn : Y.rms_post_att; p.x_out = b_xo0; p.eps = a.rms_norm_eps;
            p.unit_offset = gemma; p.epi = gemma ? EPI_GLU_GELU : EPI_GLU_SILU; p.out = b_h;
            if (peer) px_take(p);
            ph.push_back(P);
        }
        {   // quantize(hidden) -> W2 (:626-640)
            Phase P = gemv_phase(Y.w2, nullptr);
            P.g.pro = PRO_QUANT; P.g.act_in = b_h; P.g.epi = EPI_STORE; P.g.out = b_down;
            P.comm = peer ? 0 : 1;
            if (peer) { px_push(P.g, l, 1); px_pending = px_vec(m, m->rank, l, 1); }
            ph.push_back(P);
        }
        x_cur = b_xo0;
        delta = b_down;
        w_post = gemma ? Y.rms_post_ffn : nullptr;   // (:642-656) applied by the next prologue
    }
    if (!serial_prefill) {   // final rmsnorm + classifier (:343-371) + Gemma soft-cap quirk (:375-381)
        Phase P = gemv_phase(m->cls, nullptr);
        GemvParams& p = P.g;
        p.pro = PRO_NORM; p.x_in = x_cur; p.delta = delta; p.w_post = w_post; p.w_norm = m->rms_final;
        p.x_out = nullptr; p.eps = a.rms_norm_eps; p.unit_offset = gemma;
        p.epi = EPI_LOGITS; p.out = m->d_logits + m->vocab_off;
        int cap = gemma ? (int)a.dim - m->vocab_off : 0;
        p.softcap_rows = cap < 0 ? 0 : (cap > m->l_vocab ? m->l_vocab : cap);
        P.comm = peer ? 0 : 2;   // row-sharded mode: rows written into every GPU's logits buffer / NCCL all-gather
        if (peer) {
            px_take(p);
            p.px_world = m->world;
            for (int r = 0; r < m->world; r++) p.px_logits[r] = reinterpret_cast<float*>(m->xchg_peer[r] + m->xchg_logits_off) + m->vocab_off;
        }
        ph.push_back(P);
        if (peer) {   // "my rows have landed everywhere" / wait for everybody else's
            Phase F;
            memset(&F, 0, sizeof F);
            F.kind = PH_PEERFLAG;
            for (int r = 0; r < m->world; r++) F.f.flag_peer[r] = reinterpret_cast<uint32_t*>(m->xchg_peer[r] + m->xchg_flags_off);
            F.f.flag_local = reinterpret_cast<const uint32_t*>(m->d_xchg + m->xchg_flags_off);
            F.f.world = m->world; F.f.rank = m->rank; F.f.step = m->d_step;
            ph.push_back(F);
        }
    } else {                 // fill_kv_cache returns the residual stream: apply the pending add (:642-656)
        Phase P;
        memset(&P, 0, sizeof P);
        P.kind = PH_FINALIZE;
        P.r.x_in = x_cur; P.r.delta = delta; P.r.w_post = w_post; P.r.n = a.dim; P.r.eps = a.rms_norm_eps;
        P.r.rows = m->d_rows; P.r.step = m->d_step; P.r.ll = ll; P.r.scratch = m->d_fin_scratch;
        if (peer && px_pending) { P.r.px_world = m->world; P.r.px_in = px_pending; }
        ph.push_back(P);
    }
    return ph;
}

// one kernel per phase, chained with programmatic dependent launch.  LL mode: the kernels are launched early and hand
// their results over through the LL buffers (no griddepcontrol.wait); plain mode: every kernel waits for its predecessor
// (the only mode with world > 1: NCCL collectives sit between the kernels).
static int enqueue_step(lmrs_b200* m, bool decode, bool nowait = false, int only_kind = -1) {
    const std::vector<Phase>& ph = decode ? m->ph_decode : m->ph_prefill;
    const bool pdl = m->use_pdl;
    bool first = true;
    int slot = 0;
    for (const Phase& P0 : ph) {
        Phase P = P0;
        P.g.trace_slot = P.a.trace_slot = -1;
        if (m->d_trace && P.kind == PH_GEMV) P.g.trace_slot = slot;
        if (m->d_trace && P.kind == PH_ATTN) P.a.trace_slot = slot;
        slot++;
        if (only_kind >= 0 && P.kind != only_kind) continue;
        P.g.ll_nowait = P.a.ll_nowait = P.r.ll_nowait = P.f.nowait = nowait;
        // the first kernel of a step is an ordinary launch: it starts after EVERYTHING earlier in the stream has
        // completed, which is what lets later kernels of the step touch older KV rows (and reuse the LL buffers of the
        // previous step) without any further synchronisation
        m->use_pdl = pdl && !first;
        first = false;
        int rc = 0;
        cudaError_t e = cudaSuccess;
        if (P.kind == PH_GEMV) {
            e = launch_gemv(m, m->args.q_type, P.g);
            if (e == cudaSuccess && m->world > 1 && P.comm == 1) { rc = shard_allreduce(m->shard, (float*)P.g.out, m->args.dim, m->stream); m->launches++; }
            if (e == cudaSuccess && m->world > 1 && P.comm == 2) { rc = shard_allgather_logits(m->shard, m->d_logits, m->l_vocab, m->stream); m->launches++; }
        } else if (P.kind == PH_ATTN) {
            if (m->att_variant != ATT_LEGACY) e = launch_attn_cluster(m, P.a, m->l_kv_heads, m->att_var[m->att_variant].cap, m->att_var[m->att_variant].g);
            else e = launch_attn(m, P.a, m->l_kv_heads);
        } else if (P.kind == PH_PEERFLAG) {
            e = launch(m, peer_flag_kernel, dim3(1), dim3(32), 0, P.f);
        } else {
            e = launch(m, residual_finalize_kernel, dim3(1), dim3(256), 0, P.r);
        }
        if (e != cudaSuccess || rc) {
            m->use_pdl = pdl;
            return fail(rc ? std::string(shard_error()) : std::string("kernel launch: ") + cudaGetErrorString(e));
        }
    }
    m->use_pdl = pdl;
    return 0;
}

// build (once) and replay the step as a CUDA graph of the PDL-chained kernel sequence
static int run_graph(lmrs_b200* m, cudaGraphExec_t* exec, cudaStream_t* built_on, int* n_kernels, bool decode) {
    if (!m->use_graph) return enqueue_step(m, decode);
    if (!*exec || *built_on != m->stream) {
        if (*exec) { cudaGraphExecDestroy(*exec); *exec = nullptr; }
        cudaGraph_t graph = nullptr;
        uint64_t before = m->launches;
        CK(cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_step(m, decode);
        cudaError_t e = cudaStreamEndCapture(m->stream, &graph);
        if (rc || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            m->launches = before;
            if (rc) return 1;
            return fail(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
        }
        *n_kernels = (int)(m->launches - before);
        m->launches = before;
        cudaError_t ie = cudaGraphInstantiate(exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { *exec = nullptr; return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); }
        *built_on = m->stream;
    }
    CK(cudaGraphLaunch(*exec, m->stream));
    m->launches += *n_kernels;
    return 0;
}

static int setup_trace(lmrs_b200* m) {
    if (env_int("LMRS_B200_TIMING", 0)) {
        unsigned long long* tb;
        CK(cudaMalloc(&tb, 2 * 8192 * 8));
        CK(cudaMemset(tb, 0, 2 * 8192 * 8));
        CK(cudaMemcpyToSymbol(g_trace_buf, &tb, sizeof(tb)));
        m->d_trace = tb;
    }
    return 0;
}

static uint32_t next_seq(lmrs_b200* m) {   // LL sequence numbers: one per step of this handle, never 0
    if (++m->step_seq == 0) ++m->step_seq;
    return m->step_seq;
}
static int push_step(lmrs_b200* m, uint32_t token, uint32_t pos, uint32_t mask_base) {
    const uint32_t seq = next_seq(m);
    if (m->step_slot == 64) { CK(cudaStreamSynchronize(m->stream)); m->step_slot = 0; }
    StepParams* s = &m->h_step_ring[m->step_slot++];
    s->token = token; s->pos = pos; s->mask_base = mask_base; s->seq = seq;
    CK(cudaMemcpyAsync(m->d_step, s, sizeof(StepParams), cudaMemcpyHostToDevice, m->stream));
    return 0;
}

// fill_kv_cache, serial form: the decode-shaped block chain once per token (Q4 models, shapes the GEMM tiles do not
// cover, contexts beyond the attention kernel's shared-memory score capacity).  For LLAMA/PHI this is mathematically the
// reference's batched forward_layer; the Gemma window quirk is reproduced through mask_base.
static int prefill_serial(lmrs_b200* m, size_t n, uint32_t pos) {
    for (size_t i = 0; i < n; i++) {
        if (push_step(m, (uint32_t)i, pos + (uint32_t)i, pos)) return 1;
        const int v = m->att_variant = attn_variant_for(m, pos + (uint32_t)i);
        if (run_graph(m, &m->g_prefill[v], &m->g_prefill_stream[v], &m->n_prefill_kernels[v], false)) return 1;
    }
    return 0;
}

// batched-rows attention (prefill_attn.cuh): scores + softmax per (token, head) thread, then a*v per (token, head, 8 dims) thread
template <int HS> static int launch_prefill_attn_t(lmrs_b200* m, const PrefillAttnParams& p, int n_kv_heads) {
    if (m->use_pf_attn == 1) {   // fused form: score rows in shared memory
        const int scs = prefill_fused_scs(p.pos + p.n);
        const int nqd = prefill_fused_nqd(p.kv_mul, m->pfa_quads);
        const int tb = prefill_fused_tb(HS, nqd, p.kv_mul, p.pos + p.n);
        const size_t smem = prefill_fused_smem(HS, nqd, tb * p.kv_mul, scs);
        if (tb > 0) {
            const dim3 grid((unsigned)((p.n + tb - 1) / tb), (unsigned)n_kv_heads);
            if (nqd == 4) {
                CK(smem_optin(m, (const void*)prefill_attn_fused_kernel<HS, 4>, smem));
                prefill_attn_fused_kernel<HS, 4><<<grid, 128, smem, m->stream>>>(p, tb, scs);
            } else {
                CK(smem_optin(m, (const void*)prefill_attn_fused_kernel<HS, 8>, smem));
                prefill_attn_fused_kernel<HS, 8><<<grid, 256, smem, m->stream>>>(p, tb, scs);
            }
            m->launches++;
            CK(cudaGetLastError());
            return 0;
        }
    }
    const int TB = PFA_THREADS / p.kv_mul;
    CK(smem_optin(m, (const void*)prefill_scores_kernel<HS>, prefill_scores_smem<HS>()));
    prefill_scores_kernel<HS><<<dim3((unsigned)((p.n + TB - 1) / TB), (unsigned)n_kv_heads), PFA_THREADS, prefill_scores_smem<HS>(), m->stream>>>(p);
    m->launches++;
    CK(cudaGetLastError());
    const int TBV = prefill_av_tokens(HS, p.kv_mul), pairs = TBV * p.kv_mul;
    const int threads = (pairs * (HS / 8) + 31) / 32 * 32;
    CK(smem_optin(m, (const void*)prefill_av_kernel<HS>, prefill_av_smem<HS>(pairs)));
    prefill_av_kernel<HS><<<dim3((unsigned)((p.n + TBV - 1) / TBV), (unsigned)n_kv_heads), threads, prefill_av_smem<HS>(pairs), m->stream>>>(p, TBV);
    m->launches++;
    CK(cudaGetLastError());
    return 0;
}
static int launch_prefill_attn(lmrs_b200* m, const PrefillAttnParams& p, int n_kv_heads) {
    switch (m->args.head_size) {
        case 64: return launch_prefill_attn_t<64>(m, p, n_kv_heads);
        case 96: return launch_prefill_attn_t<96>(m, p, n_kv_heads);
        case 128: return launch_prefill_attn_t<128>(m, p, n_kv_heads);
        case 256: return launch_prefill_attn_t<256>(m, p, n_kv_heads);
        default: return fail("unsupported head_size");
    }
}

static bool gemm_prefill_ok(const lmrs_b200* m, size_t n, uint32_t pos) {
    if (!m->use_gemm || n < 8) return false;
    // attention of the batch: the fused kernel keeps whole score rows in shared memory (any context the KV cache holds for
    // the shipped shapes); the older forms stop at ATT_SC_CAP positions
    const bool fused = m->use_pf_attn == 1 && prefill_fused_tb((int)m->args.head_size, prefill_fused_nqd((int)(m->args.n_heads / m->args.n_kv_heads), m->pfa_quads), (int)(m->args.n_heads / m->args.n_kv_heads), (int)(pos + n)) > 0;
    if (!fused && pos + n > (size_t)ATT_SC_CAP) return false;
    for (const Layer& Y : m->layers)
        for (const Mat* x : {&Y.qkv, &Y.wo, &Y.w1, &Y.w3, &Y.w2})
            if (!x->has_tmap) return false;
    return m->l_att_dim % 128 == 0 && m->l_kv_dim % 64 == 0;
}

static int launch_rows_prologue(lmrs_b200* m, GemvParams p, int T) {
    const int n = p.n, G = n / GS;
    size_t smem = (size_t)((n + 127) / 128) * 128 + (size_t)((G * 8 + 127) / 128) * 128 + 64 * 4 + (p.pro == PRO_NORM ? (size_t)n * 4 + 128 : 0) + 128;
    if (p.pro == PRO_NORM && n > 2048) {
        CK(smem_optin(m, (const void*)rows_prologue_kernel<PRO_NORM, 4>, smem));
        rows_prologue_kernel<PRO_NORM, 4><<<T, 256, smem, m->stream>>>(p, m->pf_xq, m->pf_xs);
    } else if (p.pro == PRO_NORM) {
        CK(smem_optin(m, (const void*)rows_prologue_kernel<PRO_NORM, 2>, smem));
        rows_prologue_kernel<PRO_NORM, 2><<<T, 256, smem, m->stream>>>(p, m->pf_xq, m->pf_xs);
    } else {   // (the quantize-only prologue keeps nothing per thread across its loop: any n)
        CK(smem_optin(m, (const void*)rows_prologue_kernel<PRO_QUANT, 2>, smem));
        rows_prologue_kernel<PRO_QUANT, 2><<<T, 256, smem, m->stream>>>(p, m->pf_xq, m->pf_xs);
    }
    m->launches++;
    CK(cudaGetLastError());
    return 0;
}

// fill_kv_cache, batched form (src/transformer.rs:672-684 -> forward_layer(sl = N), :388-657): per block
//   rows prologue (residual, exact rmsnorm, quantize) -> tcgen05 GEMM [Wq;Wk;Wv] (K/V straight into the cache) -> RoPE rows
//   -> exact attention per (token, kv head) -> quantize -> GEMM Wo -> rows prologue -> GEMM W1, W3 -> act*up -> quantize
//   -> GEMM W2; the block-closing residual is folded into the next prologue, the last one is materialised at the end.
// fill_kv_cache, batched form, in three parts so that an in-process group can walk its GPUs block by block:
//   pf_begin  buffers + step parameters;  pf_layer  one transformer block of T rows;  pf_end  the pending residual
struct PfCtx { int T; uint32_t pos; float* rows; const float* delta; const float* w_post; bool fused_attn; size_t sc_stride; };

// N-GPU mode: `part` holds this GPU's partial [T][dim] product (K-sharded Wo / W2); on return it holds the sum over the
// GPUs in rank order.  Two alternating exchange buffers per GPU: a peer can be at most one exchange ahead of this GPU's
// reads (it needs this GPU's flag of exchange e to get past e).
static int pf_exchange(lmrs_b200* m, float* part, int T) {
    if (m->world == 1) return 0;
    const uint32_t e = m->pf_xchg_count++;
    const size_t dim = m->args.dim, slot = (size_t)m->pf_xchg_rows * dim;   // floats per slot
    PxRowsParams p{};
    p.src = part; p.world = m->world; p.count4 = (size_t)T * dim / 4; p.slot_stride = slot;
    for (int r = 0; r < m->world; r++)
        p.dst[r] = reinterpret_cast<float*>(m->xchg_peer[r] + m->xchg_pf_off) + ((size_t)(e & 1) * m->world + m->rank) * slot;
    p.slots = reinterpret_cast<const float*>(m->d_xchg + m->xchg_pf_off) + (size_t)(e & 1) * m->world * slot;
    p.out = part;
    const unsigned grid = (unsigned)((p.count4 + 255) / 256);
    px_rows_push_kernel<<<grid, 256, 0, m->stream>>>(p);
    m->launches++;
    CK(cudaGetLastError());
    PeerFlagParams f{};
    for (int r = 0; r < m->world; r++) f.flag_peer[r] = reinterpret_cast<uint32_t*>(m->xchg_peer[r] + m->xchg_flags_off + 128);
    f.flag_local = reinterpret_cast<const uint32_t*>(m->d_xchg + m->xchg_flags_off + 128);
    f.world = m->world; f.rank = m->rank; f.step = m->d_step; f.value = e + 1; f.use_value = 1;
    peer_flag_kernel<<<1, 32, 0, m->stream>>>(f);
    m->launches++;
    CK(cudaGetLastError());
    px_rows_sum_kernel<<<grid, 256, 0, m->stream>>>(p);
    m->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int pf_begin(lmrs_b200* m, size_t n, uint32_t pos, float* rows, PfCtx& c) {
    const lmrs_args_t& a = m->args;
    const int dim = a.dim, att = m->l_att_dim, hid = m->l_hidden;
    c.T = (int)n; c.pos = pos; c.rows = rows; c.delta = nullptr; c.w_post = nullptr;
    c.fused_attn = m->use_pf_attn == 1 && prefill_fused_tb((int)a.head_size, prefill_fused_nqd((int)(a.n_heads / a.n_kv_heads), m->pfa_quads), (int)(a.n_heads / a.n_kv_heads), (int)(pos + n)) > 0;
    c.sc_stride = c.fused_attn ? 4 : align_up(std::min<size_t>(a.seq_len, ATT_SC_CAP), 4);   // score scratch of the unfused forms (pos + n <= ATT_SC_CAP)
    if (m->pf_cap < n) {
        for (void* p : {(void*)m->pf_xq, (void*)m->pf_xs, (void*)m->pf_q, (void*)m->pf_att, (void*)m->pf_wo, (void*)m->pf_h, (void*)m->pf_down}) cudaFree(p);
        m->pf_xq = nullptr; m->pf_xs = m->pf_q = m->pf_att = m->pf_wo = m->pf_h = m->pf_down = nullptr;
        m->pf_cap = 0;   // a failed allocation below must not leave a capacity behind
        const size_t nmax = std::max<size_t>(std::max<size_t>(dim, att), hid);
        CK(cudaMalloc(&m->pf_xq, n * nmax)); CK(cudaMalloc(&m->pf_xs, n * (nmax / GS) * 4));
        CK(cudaMalloc(&m->pf_q, n * att * 4)); CK(cudaMalloc(&m->pf_att, n * att * 4)); CK(cudaMalloc(&m->pf_wo, n * dim * 4));
        CK(cudaMalloc(&m->pf_h, n * hid * 4));   // (gate and up never leave the fused GEMM)
        CK(cudaMalloc(&m->pf_down, n * dim * 4));
        m->pf_cap = n;
    }
    if (!c.fused_attn && m->pf_sc_bytes < n * (size_t)m->l_heads * c.sc_stride * 4) {   // score scratch of the unfused attention forms
        cudaFree(m->pf_scores); m->pf_scores = nullptr; m->pf_sc_bytes = 0;
        CK(cudaMalloc(&m->pf_scores, n * (size_t)m->l_heads * c.sc_stride * 4));
        m->pf_sc_bytes = n * (size_t)m->l_heads * c.sc_stride * 4;
    }
    return push_step(m, 0, pos, pos);   // attention: pos = step->pos + token index, mask_base = batch start
}

static int pf_layer(lmrs_b200* m, size_t l, PfCtx& c) {
    const lmrs_args_t& a = m->args;
    const int T = c.T, dim = a.dim, att = m->l_att_dim, kvd = m->l_kv_dim, hid = m->l_hidden;
    const uint32_t pos = c.pos;
    const bool gemma = a.model_type == 0;
    const Layer& Y = m->layers[l];
    float* kc = m->d_kcache + l * (size_t)a.seq_len * kvd;
    float* vc = m->d_vcache + l * (size_t)a.seq_len * kvd;
    {
        GemvParams p{};
        p.n = dim; p.pro = PRO_NORM; p.x_in = c.rows; p.delta = c.delta; p.w_post = c.w_post; p.w_norm = Y.rms_att;
        p.x_out = c.rows; p.eps = a.rms_norm_eps; p.unit_offset = gemma; p.step = m->d_step;
        if (launch_rows_prologue(m, p, T)) return 1;
        GemmParams g{};
        g.out0 = m->pf_q; g.ld0 = att; g.c1 = att;
        g.out1 = kc + (size_t)pos * kvd; g.ld1 = kvd; g.c2 = att + kvd;
        g.out2 = vc + (size_t)pos * kvd; g.ld2 = kvd;
        if (launch_gemm(m, Y.qkv, m->pf_xq, m->pf_xs, T, g)) return 1;
    }
    rope_rows_kernel<<<T, 256, 0, m->stream>>>(m->pf_q, kc, m->d_rope_cos, m->d_rope_sin, m->l_heads, m->l_kv_heads, a.head_size, (int)pos);
    m->launches++;
    CK(cudaGetLastError());
    if (m->use_pf_attn) {
        PrefillAttnParams p{};
        p.q = m->pf_q; p.kcache = kc; p.vcache = vc; p.probs = m->pf_scores; p.out = m->pf_att;
        p.n = T; p.pos = (int)pos; p.t_cap = (int)c.sc_stride; p.att_dim = att; p.kv_dim = kvd; p.kv_mul = a.n_heads / a.n_kv_heads;
        p.gemma = gemma; p.mask_base = pos; p.sqrt_hs = sqrtf((float)a.head_size); p.neg_zero = -0.0f;
        if (launch_prefill_attn(m, p, m->l_kv_heads)) return 1;
    } else {
        AttnParams p{};
        p.q = m->pf_q; p.k_new = nullptr; p.kcache = kc; p.vcache = vc; p.rope_cos = m->d_rope_cos; p.rope_sin = m->d_rope_sin;
        p.out = m->pf_att; p.scores = m->pf_scores; p.kv_dim = kvd; p.kv_mul = a.n_heads / a.n_kv_heads; p.chunks = m->att_chunks;
        p.gemma = gemma; p.seq_len = (int)c.sc_stride; p.sqrt_hs = sqrtf((float)a.head_size); p.step = m->d_step;
        p.batch = 1; p.q_stride = att;
        bool pdl = m->use_pdl; m->use_pdl = false;
        cudaError_t e = launch_attn_grid(m, p, m->l_kv_heads, T);
        m->use_pdl = pdl;
        CK(e);
    }
    {
        GemvParams p{};
        p.n = att; p.pro = PRO_QUANT; p.act_in = m->pf_att; p.step = m->d_step;
        if (launch_rows_prologue(m, p, T)) return 1;
        if (launch_gemm(m, Y.wo, m->pf_xq, m->pf_xs, T, gemm_out1(m->pf_wo, dim, dim))) return 1;
        if (pf_exchange(m, m->pf_wo, T)) return 1;
    }
    {
        GemvParams p{};
        p.n = dim; p.pro = PRO_NORM; p.x_in = c.rows; p.delta = m->pf_wo; p.w_post = gemma ? Y.rms_post_att : nullptr;
        p.w_norm = gemma ? Y.rms_pre_ffn : Y.rms_post_att; p.x_out = c.rows; p.eps = a.rms_norm_eps; p.unit_offset = gemma; p.step = m->d_step;
        if (launch_rows_prologue(m, p, T)) return 1;
        // gate and up in one launch with act(gate) * up in its epilogue (:607-624): g and u never travel through HBM
        if (launch_gemm_glu(m, Y.w1, Y.w3, m->pf_xq, m->pf_xs, T, m->pf_h, hid, gemma ? EPI_GLU_GELU : EPI_GLU_SILU)) return 1;
    }
    {
        GemvParams p{};
        p.n = hid; p.pro = PRO_QUANT; p.act_in = m->pf_h; p.step = m->d_step;
        if (launch_rows_prologue(m, p, T)) return 1;
        if (launch_gemm(m, Y.w2, m->pf_xq, m->pf_xs, T, gemm_out1(m->pf_down, dim, dim))) return 1;
        if (pf_exchange(m, m->pf_down, T)) return 1;
    }
    c.delta = m->pf_down;
    c.w_post = gemma ? Y.rms_post_ffn : nullptr;
    return 0;
}

static int pf_end(lmrs_b200* m, PfCtx& c) {
    ResidualParams r{};
    r.x_in = c.rows; r.delta = c.delta; r.w_post = c.w_post; r.n = m->args.dim; r.eps = m->args.rms_norm_eps; r.rows = c.rows; r.step = m->d_step;
    r.row_from_block = 1;
    residual_finalize_kernel<<<c.T, 256, 0, m->stream>>>(r);
    m->launches++;
    CK(cudaGetLastError());
    return 0;
}

// rows per pass: everything at once on one GPU; N-GPU mode walks the batch in chunks the exchange buffers hold (a later
// chunk attends to the earlier ones through the KV cache: same results as one pass)
static size_t pf_chunk_rows(const lmrs_b200* m, size_t n) { return m->world > 1 ? std::min<size_t>(n, (size_t)m->pf_xchg_rows) : n; }

static int prefill_gemm(lmrs_b200* m, size_t n, uint32_t pos) {
    for (size_t r0 = 0; r0 < n;) {
        const size_t cn = pf_chunk_rows(m, n - r0);
        PfCtx c;
        if (pf_begin(m, cn, pos + (uint32_t)r0, m->d_rows + r0 * m->args.dim, c)) return 1;
        for (size_t l = 0; l < m->args.n_layers; l++)
            if (pf_layer(m, l, c)) return 1;
        if (pf_end(m, c)) return 1;
        r0 += cn;
    }
    return 0;
}

static int prefill_batched(lmrs_b200* m, size_t n, uint32_t pos) {
    if (gemm_prefill_ok(m, n, pos)) return prefill_gemm(m, n, pos);
    return prefill_serial(m, n, pos);
}

// N-GPU mode: allocate this GPU's exchange block and map every peer's (CUDA IPC; handles travel through one NCCL all-gather)
static int alloc_peer_exchange(lmrs_b200* m) {
    const size_t W = (size_t)m->world, dim = m->args.dim, L = m->args.n_layers;
    if (W > (size_t)PX_MAX_WORLD) return fail("peer exchange supports up to 8 GPUs");
    const size_t px_bytes = L * 2 * W * dim * sizeof(llword_t);
    m->xchg_flags_off = align_up(px_bytes, 256);
    m->xchg_logits_off = m->xchg_flags_off + 256;
    m->pf_xchg_rows = std::max(8, env_int("LMRS_B200_PF_ROWS", 512));
    m->xchg_pf_off = align_up(m->xchg_logits_off + (size_t)m->args.vocab_size * 4, 256);
    const size_t total = m->xchg_pf_off + (m->use_gemm ? (size_t)2 * W * m->pf_xchg_rows * dim * 4 : 0);
    CK(cudaMalloc(&m->d_xchg, total));
    CK(cudaMemset(m->d_xchg, 0, total));   // sequence number 0 = "never written"
    CK(cudaDeviceSynchronize());
    m->xchg_peer[m->rank] = m->d_xchg;
    m->d_logits = reinterpret_cast<float*>(m->d_xchg + m->xchg_logits_off);
    return 0;
}
static int setup_peer_exchange(lmrs_b200* m) {
    const size_t W = (size_t)m->world;
    if (alloc_peer_exchange(m)) return 1;
    cudaIpcMemHandle_t mine;
    CK(cudaIpcGetMemHandle(&mine, m->d_xchg));
    uint8_t* d_h = nullptr;
    CK(cudaMalloc(&d_h, (W + 1) * sizeof(mine)));
    CK(cudaMemcpy(d_h, &mine, sizeof(mine), cudaMemcpyHostToDevice));
    if (shard_allgather_bytes(m->shard, d_h, d_h + sizeof(mine), sizeof(mine), m->stream)) { cudaFree(d_h); return fail(shard_error()); }
    CK(cudaStreamSynchronize(m->stream));
    std::vector<cudaIpcMemHandle_t> all(W);
    CK(cudaMemcpy(all.data(), d_h + sizeof(mine), W * sizeof(mine), cudaMemcpyDeviceToHost));
    cudaFree(d_h);
    for (size_t r = 0; r < W; r++) {
        if ((int)r == m->rank) { m->xchg_peer[r] = m->d_xchg; continue; }
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return fail(std::string("cudaIpcOpenMemHandle (peer exchange between GPUs; LMRS_B200_PEER=0 selects the NCCL data path): ") + cudaGetErrorString(e));
        m->xchg_peer[r] = (uint8_t*)ptr;
    }
    return 0;
}

// ---- C ABI -----------------------------------------------------------------------------------------------
static int alloc_peer_exchange(lmrs_b200* m);
// phase 1 of a handle: device checks, knobs, weights and state.  `inproc`: a member of an in-process group (no NCCL, the
// peer mapping and the phase tables follow once every member exists: finish_create)
static int create_common(const uint8_t* file, size_t len, int device, int rank, int world, const void* nccl_id,
                         lmrs_b200_t** out, size_t* end_offset, bool inproc = false) {
    if (!file || !out) return fail("null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("lmrs_b200: no CUDA device available (this library has no CPU fallback)");
    if (device < 0) CK(cudaGetDevice(&device));
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(std::string("lmrs_b200: device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                    "; this library contains sm_100a code only");
    lmrs_b200* m = new lmrs_b200();
    m->device = device;
    m->sms = prop.multiProcessorCount;
    m->rank = rank; m->world = world;
    m->use_graph = env_int("LMRS_B200_GRAPH", 1) != 0;
    m->use_pdl = env_int("LMRS_B200_PDL", 1) != 0;
    // ring geometry: 16 warps x 2 stages measured best for the decode chain (profiles/r2_decode_experiments.md): the exact-
    // order prologues and the per-stage scans are latency chains that want warps, deeper rings measured slower
    m->gemv_cfg = env_int("LMRS_B200_GEMV_CFG", 3);
    if (m->gemv_cfg < 0 || m->gemv_cfg >= N_GEMV_CFG) m->gemv_cfg = 0;
    m->gemv_ctas_per_sm = env_int("LMRS_B200_GEMV_CTAS", 1);
    // Hand-over between the kernels of a step.  Default: kernel boundaries (programmatic dependent launch +
    // griddepcontrol.wait), one matrix-vector CTA per SM.  LMRS_B200_LL=1: fence-free LL exchange between co-resident
    // kernels (common.cuh) -- bit-identical and covered by the tests, but measured SLOWER on B200 (an L2 round trip costs
    // 0.3-0.6 us, a hand-over needs two or three of them, and co-resident waiting kernels slow the running one; see
    // profiles/r2_decode_experiments.md), so it is opt-in.
    m->use_ll = env_int("LMRS_B200_LL", 0) != 0 && world == 1;
    m->use_pf_attn = env_int("LMRS_B200_PF_ATTN", 1);
    m->pfa_quads = env_int("LMRS_B200_PFA_QUADS", 4) == 8 ? 8 : 4;
    m->l2pf = env_int("LMRS_B200_L2PF", 0);
    m->l2pf_ef = env_int("LMRS_B200_L2PF_EF", 1);
    m->l2pf_cls_mb = env_int("LMRS_B200_L2PF_CLS_MB", 24);
    m->l2pf_chunk = std::max(1024, env_int("LMRS_B200_L2PF_CHUNK", 32768)) & ~15;
    // N-GPU data path: partial vectors pushed between the GPUs by the kernels themselves (default), or NCCL collectives
    // between the kernels (LMRS_B200_PEER=0)
    m->use_peer = world > 1 && (inproc || env_int("LMRS_B200_PEER", 1) != 0);
    m->in_process_group = inproc;
    if (build_model(m, file, len, end_offset)) { lmrs_b200_destroy(m); return 1; }
    if ((int)m->args.dim > NORM_MAX_DIM) {
        lmrs_b200_destroy(m);
        return fail("dim too large for the fused norm prologue of this GEMV configuration");
    }
    if (inproc) {   // the group wires the peers and finishes the members together
        if (alloc_peer_exchange(m)) { lmrs_b200_destroy(m); return 1; }
        *out = m;
        return 0;
    }
    if (world > 1 && shard_init(m->shard, rank, world, nccl_id, m->args.dim)) { lmrs_b200_destroy(m); return fail(shard_error()); }
    if (m->use_peer && setup_peer_exchange(m)) { lmrs_b200_destroy(m); return 1; }
    if (m->use_peer) shard_destroy(m->shard);   // NCCL only carried the IPC handles: no communicator is kept (or torn down at exit) in peer mode
    setup_attn_cluster(m);
    if (setup_trace(m)) { lmrs_b200_destroy(m); return 1; }
    m->ph_decode = make_phases(m, false);
    *out = m;
    return 0;
}

// Transformer::new on n_gpus GPUs of this process (devices 0 .. n_gpus-1): one handle per GPU, row-sharded like
// create_sharded, peers mapped with cudaDeviceEnablePeerAccess (no NCCL, no second process).  The returned handle leads
// the group: every entry point fans out to the members from the caller's thread (the kernels of the members wait for
// one another on the device, the host never does).
extern "C" int lmrs_b200_create_multi(const uint8_t* file, size_t len, int n_gpus, lmrs_b200_t** out, size_t* end_offset) {
    if (!file || !out) return fail("null argument");
    if (n_gpus <= 1) return create_common(file, len, n_gpus == 1 ? 0 : -1, 0, 1, nullptr, out, end_offset);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < n_gpus) return fail("lmrs_b200_create_multi: fewer CUDA devices than n_gpus");
    if (n_gpus > PX_MAX_WORLD) return fail("lmrs_b200_create_multi: at most 8 GPUs");
    std::vector<lmrs_b200*> g((size_t)n_gpus, nullptr);
    auto bail = [&]() { std::string e = g_err; for (lmrs_b200* x : g) if (x) { x->group.clear(); lmrs_b200_destroy(x); } g_err = e; return 1; };
    for (int r = 0; r < n_gpus; r++)
        if (create_common(file, len, r, r, n_gpus, nullptr, &g[r], r == 0 ? end_offset : nullptr, /*inproc=*/true)) return bail();
    for (int a = 0; a < n_gpus; a++) {
        if (cudaSetDevice(g[a]->device) != cudaSuccess) { fail("cudaSetDevice failed"); return bail(); }
        for (int b = 0; b < n_gpus; b++) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, g[a]->device, g[b]->device);
            if (!can) { fail("lmrs_b200_create_multi: GPUs " + std::to_string(a) + " and " + std::to_string(b) + " have no peer access"); return bail(); }
            cudaError_t e = cudaDeviceEnablePeerAccess(g[b]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { fail(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return bail(); }
            cudaGetLastError();
        }
        for (int b = 0; b < n_gpus; b++) g[a]->xchg_peer[b] = g[b]->d_xchg;
    }
    for (int r = 0; r < n_gpus; r++) {
        if (cudaSetDevice(g[r]->device) != cudaSuccess) { fail("cudaSetDevice failed"); return bail(); }
        setup_attn_cluster(g[r]);
        if (setup_trace(g[r])) return bail();
        g[r]->ph_decode = make_phases(g[r], false);
    }
    g[0]->group = g;
    *out = g[0];
    return 0;
}
// the handles an entry point has to drive: the members of a group, or the handle itself
static std::vector<lmrs_b200*> members(lmrs_b200* m) { return m->group.empty() ? std::vector<lmrs_b200*>{m} : m->group; }

extern "C" int lmrs_b200_create(const uint8_t* file, size_t len, int device, lmrs_b200_t** out, size_t* end_offset) {
    return create_common(file, len, device, 0, 1, nullptr, out, end_offset);
}
extern "C" int lmrs_b200_create_sharded(const uint8_t* file, size_t len, int device, int rank, int world,
                                        const void* nccl_unique_id, lmrs_b200_t** out, size_t* end_offset) {
    if (world < 1 || rank < 0 || rank >= world) return fail("bad rank/world");
    if (world > 1 && !nccl_unique_id) return fail("nccl_unique_id required when world > 1");
    return create_common(file, len, device, rank, world, nccl_unique_id, out, end_offset);
}
extern "C" int lmrs_b200_nccl_unique_id(void* out128) {
    if (shard_unique_id(out128)) return fail(shard_error());
    return 0;
}

extern "C" void lmrs_b200_destroy(lmrs_b200_t* m) {
    if (!m) return;
    if (!m->group.empty()) {   // leader of an in-process group: the members first (all idle), then itself
        std::vector<lmrs_b200*> g = m->group;
        for (lmrs_b200* x : g) { cudaSetDevice(x->device); if (x->own_stream) cudaStreamSynchronize(x->own_stream); }
        m->group.clear();
        for (size_t i = 1; i < g.size(); i++) lmrs_b200_destroy(g[i]);
    }
    cudaSetDevice(m->device);
    if (m->own_stream) cudaStreamSynchronize(m->own_stream);
    for (int v = 0; v < 8; v++) {
        if (m->g_decode[v]) cudaGraphExecDestroy(m->g_decode[v]);
        if (m->g_prefill[v]) cudaGraphExecDestroy(m->g_prefill[v]);
    }
    for (int r = 0; r < m->world && r < PX_MAX_WORLD; r++)
        if (!m->in_process_group && m->xchg_peer[r] && m->xchg_peer[r] != m->d_xchg) cudaIpcCloseMemHandle(m->xchg_peer[r]);
    // An exchange block exported through CUDA IPC is NOT freed here: cudaFree of exported memory while another process still
    // has it mapped is undefined behaviour, and nothing orders this rank's destroy against its peers' cudaIpcCloseMemHandle
    // (a collective in a destructor would hang as soon as one rank has died).  The block (a few MB to ~140 MB) goes with the
    // process; the driver reference-counts it across processes.  In-process groups free theirs normally.
    if (m->in_process_group || m->world == 1) cudaFree(m->d_xchg);
    shard_destroy(m->shard);
    cudaFree(m->d_arena); cudaFree(m->d_dense); cudaFree(m->pf_xq); cudaFree(m->pf_xs); cudaFree(m->pf_q); cudaFree(m->pf_att); cudaFree(m->pf_wo);
    cudaFree(m->pf_g); cudaFree(m->pf_u); cudaFree(m->pf_h); cudaFree(m->pf_down); cudaFree(m->pf_scores); cudaFree(m->d_kcache); cudaFree(m->d_vcache); cudaFree(m->d_rope_cos); cudaFree(m->d_rope_sin);
    cudaFree(m->d_x[0]); cudaFree(m->d_x[1]); cudaFree(m->d_q); cudaFree(m->d_knew); cudaFree(m->d_att);
    cudaFree(m->d_wo_out); cudaFree(m->d_h); cudaFree(m->d_down_out); if (!m->d_xchg) cudaFree(m->d_logits); cudaFree(m->d_scores);
    cudaFree(m->d_step); cudaFree(m->d_rows); cudaFree(m->d_ll); cudaFree(m->d_fin_scratch);
    cudaFree(m->d_amax); cudaFree(m->d_aidx); cudaFree(m->d_ticket); cudaFree(m->d_next); cudaFree(m->d_gen);
    if (m->h_gen) cudaFreeHost(m->h_gen);
    if (m->ev_pf0) cudaEventDestroy(m->ev_pf0);
    if (m->ev_pf1) cudaEventDestroy(m->ev_pf1);
    if (m->h_step_ring) cudaFreeHost(m->h_step_ring);
    if (m->h_logits) cudaFreeHost(m->h_logits);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    delete m;
}

extern "C" int lmrs_b200_args(const lmrs_b200_t* m, lmrs_args_t* out) {
    if (!m || !out) return fail("null argument");
    *out = m->args;
    return 0;
}

static int step_one(lmrs_b200* m, bool with_params, uint32_t token, uint32_t pos) {   // one decode step of ONE handle
    CK(cudaSetDevice(m->device));
    if (with_params) { if (push_step(m, token, pos, pos)) return 1; }
    else next_seq(m);   // the device advanced (token, pos, seq) itself: keep the host's sequence counter in step
    const int v = m->att_variant = attn_variant_for(m, pos);
    return run_graph(m, &m->g_decode[v], &m->g_decode_stream[v], &m->n_decode_kernels[v], true);
}
extern "C" int lmrs_b200_forward_device(lmrs_b200_t* m, uint32_t token, uint32_t pos) {
    if (!m) return fail("null handle");
    if (token >= m->args.vocab_size) return fail("token out of range");
    if (pos >= m->args.seq_len) return fail("position out of range (seq_len is clamped to 8192, src/transformer.rs:158)");
    for (lmrs_b200* g : members(m))   // in-process group: every GPU's step is enqueued from this thread; they meet on the device
        if (step_one(g, true, token, pos)) return 1;
    return 0;
}

extern "C" int lmrs_b200_forward(lmrs_b200_t* m, uint32_t token, uint32_t pos, float** logits_host) {
    if (!logits_host) return fail("null argument");
    if (lmrs_b200_forward_device(m, token, pos)) return 1;
    CK(cudaSetDevice(m->device));
    CK(cudaMemcpyAsync(m->h_logits, m->d_logits, (size_t)m->args.vocab_size * 4, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    *logits_host = m->h_logits;
    return 0;
}

// ---- greedy sampling on the device (Sampler::sample with temperature 0 -> sample_argmax, src/sampler.rs:29-41,112-113)
static int ensure_gen(lmrs_b200* m, size_t n) {
    if (m->gen_cap >= n) return 0;
    cudaFree(m->d_gen); m->d_gen = nullptr;
    if (m->h_gen) { cudaFreeHost(m->h_gen); m->h_gen = nullptr; }
    m->gen_cap = 0;
    CK(cudaMalloc(&m->d_gen, n * 4));
    CK(cudaMallocHost(&m->h_gen, n * 4));
    m->gen_cap = n;
    return 0;
}
static int launch_argmax(lmrs_b200* m, uint32_t* d_out, bool advance) {
    ArgmaxParams ap{};
    ap.x = m->d_logits; ap.n = (int)m->args.vocab_size; ap.pmax = m->d_amax; ap.pidx = m->d_aidx; ap.ticket = m->d_ticket;
    ap.out = d_out; ap.advance = advance ? m->d_step : nullptr;
    int grid = (int)std::min<size_t>((size_t)m->sms, (m->args.vocab_size + 1023) / 1024);
    if (grid < 1) grid = 1;
    if (grid > 256) grid = 256;
    CK(launch(m, argmax_kernel, dim3(grid), dim3(256), 0, ap));   // PDL: waits for the classifier (griddepcontrol.wait)
    return 0;
}
extern "C" int lmrs_b200_forward_argmax(lmrs_b200_t* m, uint32_t token, uint32_t pos, uint32_t* next_token) {
    if (!next_token) return fail("null argument");
    if (m && m->world > 1 && !m->use_peer) return fail("forward_argmax: not with the NCCL data path (LMRS_B200_PEER=0)");   // peer mode: every GPU holds all logits rows
    if (lmrs_b200_forward_device(m, token, pos)) return 1;
    CK(cudaSetDevice(m->device));   // (a group leader holds every rank's logits rows: its own pick is the answer)
    if (ensure_gen(m, 64)) return 1;
    if (launch_argmax(m, m->d_next, false)) return 1;
    CK(cudaMemcpyAsync(m->h_gen, m->d_next, 4, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    *next_token = m->h_gen[0];
    return 0;
}
// the generate loop of src/bin/chat.rs:188-226 at temperature 0 with the sampled token fed back ON THE DEVICE: the
// argmax kernel writes the next step's parameters (token, pos + 1), so consecutive steps need no host round trip; the
// host only reads the produced ids back in chunks to look for `eos` (steps enqueued past it only touch cache rows that
// a later call overwrites).
extern "C" int lmrs_b200_generate_greedy(lmrs_b200_t* m, uint32_t first_token, uint32_t pos, uint32_t max_new, int32_t eos,
                                         uint32_t* out_tokens, uint32_t* n_out) {
    if (!m || !out_tokens || !n_out) return fail("null argument");
    if (m->world > 1 && !m->use_peer) return fail("generate_greedy: not with the NCCL data path (LMRS_B200_PEER=0)");
    if (first_token >= m->args.vocab_size) return fail("token out of range");
    if ((size_t)pos + max_new > m->args.seq_len) return fail("position out of range (seq_len is clamped to 8192, src/transformer.rs:158)");
    *n_out = 0;
    if (max_new == 0) return 0;
    const std::vector<lmrs_b200*> grp = members(m);
    for (lmrs_b200* g : grp) { CK(cudaSetDevice(g->device)); if (ensure_gen(g, max_new)) return 1; }
    const uint32_t chunk = 32;
    for (uint32_t i0 = 0; i0 < max_new; i0 += chunk) {
        const uint32_t i1 = std::min(max_new, i0 + chunk);
        for (uint32_t i = i0; i < i1; i++)
            for (lmrs_b200* g : grp) {   // every GPU of a group picks the same token from the same logits and advances itself
                if (step_one(g, i == 0, first_token, pos + i)) return 1;
                if (launch_argmax(g, g->d_gen + i, true)) return 1;
            }
        CK(cudaSetDevice(m->device));
        CK(cudaMemcpyAsync(m->h_gen + i0, m->d_gen + i0, (size_t)(i1 - i0) * 4, cudaMemcpyDeviceToHost, m->stream));
        CK(cudaStreamSynchronize(m->stream));
        for (uint32_t i = i0; i < i1; i++) {
            out_tokens[i] = m->h_gen[i];
            *n_out = i + 1;
            if (eos >= 0 && m->h_gen[i] == (uint32_t)eos) return 0;
        }
    }
    return 0;
}

extern "C" int lmrs_b200_bench_gemv_pass(lmrs_b200_t* m, uint32_t pos, int* n_launches) {
    if (!m) return fail("null handle");
    if (pos >= m->args.seq_len) return fail("position out of range");
    for (lmrs_b200* g : members(m)) {
        CK(cudaSetDevice(g->device));
        if (push_step(g, 0, pos, pos)) return 1;
        const uint64_t before = g->launches;
        if (enqueue_step(g, true, /*nowait=*/true, PH_GEMV)) return 1;   // LL / peer-exchange kernels take whatever their input buffers hold
        if (n_launches && g == m) *n_launches = (int)(g->launches - before);
    }
    return 0;
}

extern "C" int lmrs_b200_bench_attn_pass(lmrs_b200_t* m, uint32_t pos, int* n_launches) {
    if (!m) return fail("null handle");
    if (pos >= m->args.seq_len) return fail("position out of range");
    CK(cudaSetDevice(m->device));
    if (push_step(m, 0, pos, pos)) return 1;
    m->att_variant = attn_variant_for(m, pos);
    const uint64_t before = m->launches;
    if (enqueue_step(m, true, /*nowait=*/true, PH_ATTN)) return 1;
    if (n_launches) *n_launches = (int)(m->launches - before);
    return 0;
}

extern "C" int lmrs_b200_logits_device(lmrs_b200_t* m, float** logits_dev) {
    if (!m || !logits_dev) return fail("null argument");
    *logits_dev = m->d_logits;
    return 0;
}
extern "C" int lmrs_b200_set_stream(lmrs_b200_t* m, void* s) {
    if (!m) return fail("null handle");
    if (!m->group.empty() && s) return fail("set_stream: an in-process multi-GPU handle drives one library-owned stream per GPU");
    CK(cudaSetDevice(m->device));
    CK(cudaStreamSynchronize(m->stream));
    m->stream = s ? (cudaStream_t)s : m->own_stream;
    return 0;
}
extern "C" int lmrs_b200_synchronize(lmrs_b200_t* m) {
    if (!m) return fail("null handle");
    for (lmrs_b200* g : members(m)) { CK(cudaSetDevice(g->device)); CK(cudaStreamSynchronize(g->stream)); }
    CK(cudaSetDevice(m->device));
    return 0;
}
extern "C" int lmrs_b200_kernel_launches(const lmrs_b200_t* m, uint64_t* count) {
    if (!m || !count) return fail("null argument");
    *count = 0;
    for (lmrs_b200* g : members(const_cast<lmrs_b200*>(m))) *count += g->launches;   // a group counts the kernels of all its GPUs
    return 0;
}

extern "C" int lmrs_b200_get_embeddings(const lmrs_b200_t* cm, const uint32_t* tokens, size_t n, float* out) {
    lmrs_b200* m = const_cast<lmrs_b200*>(cm);
    if (!m || !tokens || !out) return fail("null argument");
    if (n == 0) return 0;
    for (size_t i = 0; i < n; i++)
        if (tokens[i] >= m->args.vocab_size) return fail("token out of range");
    CK(cudaSetDevice(m->device));
    uint32_t* d_tok; float* d_out;
    CK(cudaMalloc(&d_tok, n * 4));
    CK(cudaMalloc(&d_out, n * (size_t)m->args.dim * 4));
    CK(cudaMemcpyAsync(d_tok, tokens, n * 4, cudaMemcpyHostToDevice, m->stream));
    EmbedParams p{};
    p.dim = m->args.dim; p.q_type = m->args.q_type; p.q = m->emb.q; p.s = m->emb.s; p.apply_scale = 0;   // :659-669: no Gemma scaling
    p.tokens = d_tok; p.step = m->d_step; p.out = d_out;
    bool pdl = m->use_pdl; m->use_pdl = false;
    cudaError_t e = launch(m, embed_kernel, dim3((unsigned)n), dim3(256), 0, p);
    m->use_pdl = pdl;
    CK(e);
    CK(cudaMemcpyAsync(out, d_out, n * (size_t)m->args.dim * 4, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    cudaFree(d_tok); cudaFree(d_out);
    return 0;
}

extern "C" int lmrs_b200_fill_kv_cache(lmrs_b200_t* m, float* emb, size_t n_floats, uint32_t pos, uint32_t* new_pos) {
    if (!m || !emb || !new_pos) return fail("null argument");
    const size_t dim = m->args.dim;
    const size_t n = n_floats / dim;
    if (pos + n > m->args.seq_len) return fail("position out of range (seq_len is clamped to 8192, src/transformer.rs:158)");
    {   // the reference walks the batch in att_dim-sized chunks of an n*dim buffer (:501-503): one chunk too many -> panic there
        const size_t att_dim = (size_t)m->args.n_heads * m->args.head_size;
        if (att_dim < dim && n * (dim - att_dim) >= att_dim)
            return fail("sl*(dim-att_dim) >= att_dim is out of bounds in the reference (src/transformer.rs:501-503)");
    }
    CK(cudaSetDevice(m->device));
    if (n == 0) { *new_pos = pos; return 0; }
    const std::vector<lmrs_b200*> grp = members(m);
    for (lmrs_b200* g : grp) {   // every GPU of a group stages the same embedding rows
        CK(cudaSetDevice(g->device));
        if (g->rows_cap < n * dim) {
            cudaFree(g->d_rows);
            g->d_rows = nullptr; g->rows_cap = 0;
            CK(cudaMalloc(&g->d_rows, n * dim * 4));
            g->rows_cap = n * dim;
            for (int v = 0; v < 8; v++)
                if (g->g_prefill[v]) { cudaGraphExecDestroy(g->g_prefill[v]); g->g_prefill[v] = nullptr; }
            g->ph_prefill = make_phases(g, true);   // the phase table embeds the staging buffer's address
        }
        if (!g->ev_pf0) { CK(cudaEventCreate(&g->ev_pf0)); CK(cudaEventCreate(&g->ev_pf1)); }
        CK(cudaMemcpyAsync(g->d_rows, emb, n * dim * 4, cudaMemcpyHostToDevice, g->stream));
        CK(cudaEventRecord(g->ev_pf0, g->stream));
    }
    if (grp.size() == 1) {
        if (prefill_batched(m, n, pos)) return 1;
    } else if (gemm_prefill_ok(m, n, pos)) {   // in-process group, batched form: block by block across the GPUs
        for (size_t r0 = 0; r0 < n;) {
            const size_t cn = pf_chunk_rows(m, n - r0);
            std::vector<PfCtx> cs(grp.size());
            for (size_t gi = 0; gi < grp.size(); gi++) {
                CK(cudaSetDevice(grp[gi]->device));
                if (pf_begin(grp[gi], cn, pos + (uint32_t)r0, grp[gi]->d_rows + r0 * dim, cs[gi])) return 1;
            }
            for (size_t l = 0; l < m->args.n_layers; l++)
                for (size_t gi = 0; gi < grp.size(); gi++) {
                    CK(cudaSetDevice(grp[gi]->device));
                    if (pf_layer(grp[gi], l, cs[gi])) return 1;
                }
            for (size_t gi = 0; gi < grp.size(); gi++) {
                CK(cudaSetDevice(grp[gi]->device));
                if (pf_end(grp[gi], cs[gi])) return 1;
            }
            r0 += cn;
        }
    } else {   // in-process group: the per-token chain, token by token across the GPUs (they meet on the device)
        for (size_t i = 0; i < n; i++)
            for (lmrs_b200* g : grp) {
                CK(cudaSetDevice(g->device));
                if (push_step(g, (uint32_t)i, pos + (uint32_t)i, pos)) return 1;
                const int v = g->att_variant = attn_variant_for(g, pos + (uint32_t)i);
                if (run_graph(g, &g->g_prefill[v], &g->g_prefill_stream[v], &g->n_prefill_kernels[v], false)) return 1;
            }
    }
    for (lmrs_b200* g : grp) { CK(cudaSetDevice(g->device)); CK(cudaEventRecord(g->ev_pf1, g->stream)); }
    CK(cudaSetDevice(m->device));
    CK(cudaMemcpyAsync(emb, m->d_rows, n * dim * 4, cudaMemcpyDeviceToHost, m->stream));
    for (lmrs_b200* g : grp) { CK(cudaSetDevice(g->device)); CK(cudaStreamSynchronize(g->stream)); }
    CK(cudaSetDevice(m->device));
    if (cudaEventElapsedTime(&m->last_prefill_ms, m->ev_pf0, m->ev_pf1) != cudaSuccess) m->last_prefill_ms = -1.0f;
    *new_pos = pos + (uint32_t)n;
    return 0;
}

extern "C" int lmrs_b200_last_prefill_device_ms(const lmrs_b200_t* m, float* ms) {
    if (!m || !ms) return fail("null argument");
    *ms = m->last_prefill_ms;
    return 0;
}

extern "C" int lmrs_b200_read_kv(lmrs_b200_t* m, uint32_t layer, uint32_t pos0, uint32_t n, float* k_out, float* v_out) {
    if (!m || !k_out || !v_out) return fail("null argument");
    if (layer >= m->args.n_layers || pos0 + n > m->args.seq_len) return fail("out of range");
    const std::vector<lmrs_b200*> grp = members(m);
    // a group returns whole rows [n][kv_dim] (GPU r holds columns [r * l_kv_dim, (r + 1) * l_kv_dim)); a single sharded
    // handle returns its own slice [n][l_kv_dim]
    const size_t out_ld = grp.size() > 1 ? (size_t)m->l_kv_dim * grp.size() : (size_t)m->l_kv_dim;
    for (size_t r = 0; r < grp.size(); r++) {
        lmrs_b200* g = grp[r];
        CK(cudaSetDevice(g->device));
        CK(cudaStreamSynchronize(g->stream));
        const size_t base = ((size_t)layer * g->args.seq_len + pos0) * g->l_kv_dim;
        CK(cudaMemcpy2D(k_out + r * g->l_kv_dim, out_ld * 4, g->d_kcache + base, (size_t)g->l_kv_dim * 4, (size_t)g->l_kv_dim * 4, n, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy2D(v_out + r * g->l_kv_dim, out_ld * 4, g->d_vcache + base, (size_t)g->l_kv_dim * 4, (size_t)g->l_kv_dim * 4, n, cudaMemcpyDeviceToHost));
    }
    CK(cudaSetDevice(m->device));
    return 0;
}

extern "C" int lmrs_b200_debug_buffer(lmrs_b200_t* m, const char* name, float* out, size_t* n) {
    if (!m || !name || !out || !n) return fail("null argument");
    const std::string nm(name);
    CK(cudaSetDevice(m->device));
    if (nm == "trace_reset") {   // the device-side counter restarts with every launch
        CK(cudaStreamSynchronize(m->stream));
        if (m->d_trace) CK(cudaMemsetAsync(m->d_trace, 0, 2 * 8192 * 8, m->stream));
        *n = 0;
        return 0;
    }
    if (nm == "trace") {
        if (!m->d_trace) return fail("trace not enabled (LMRS_B200_TIMING=1)");
        const size_t cnt = 2 * 8192 * 2;
        if (*n < cnt) return fail("buffer too small");
        CK(cudaStreamSynchronize(m->stream));
        CK(cudaMemcpy(out, m->d_trace, cnt * 4, cudaMemcpyDeviceToHost));
        *n = cnt;
        return 0;
    }
    // activation buffers of the LAST block: plain f32 arrays, or LL word arrays whose low halves are the values
    const LLLayout Y0 = ll_layout(m);
    const llword_t* lb = m->d_ll + (size_t)(m->args.n_layers - 1) * Y0.per_layer;
    const float* src = nullptr; const llword_t* lsrc = nullptr;
    size_t cnt = 0;
    if (nm == "x0") { src = m->d_x[0]; lsrc = lb + Y0.xo0; cnt = m->args.dim; }
    else if (nm == "x1") { src = m->d_x[1]; lsrc = lb + Y0.xo1; cnt = m->args.dim; }
    else if (nm == "q") { src = m->d_q; lsrc = lb + Y0.q; cnt = m->l_att_dim; }
    else if (nm == "k_new") { src = m->d_knew; lsrc = lb + Y0.k_new; cnt = m->l_kv_dim; }
    else if (nm == "att") { src = m->d_att; lsrc = lb + Y0.att; cnt = m->l_att_dim; }
    else if (nm == "wo_out") { src = m->d_wo_out; lsrc = lb + Y0.wo_out; cnt = m->args.dim; }
    else if (nm == "h") { src = m->d_h; lsrc = lb + Y0.h; cnt = m->l_hidden; }
    else if (nm == "down_out") { src = m->d_down_out; lsrc = lb + Y0.down_out; cnt = m->args.dim; }
    else return fail("unknown buffer name");
    if (*n < cnt) return fail("buffer too small");
    CK(cudaStreamSynchronize(m->stream));
    if (m->use_ll) {
        std::vector<llword_t> w(cnt);
        CK(cudaMemcpy(w.data(), lsrc, cnt * sizeof(llword_t), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < cnt; i++) { const uint32_t b = (uint32_t)w[i]; memcpy(&out[i], &b, 4); }
    } else {
        CK(cudaMemcpy(out, src, cnt * 4, cudaMemcpyDeviceToHost));
    }
    *n = cnt;
    return 0;
}

#include "ops_abi.inc"
