// gemm.cuh -- batched (sl > 1) Q8_0 matmul on the 5th-generation tensor cores: tcgen05.mma kind::i8 with TMEM
// accumulators and TMA-staged operand tiles.  Replaces functional.rs::matmul_q8 (src/functional.rs:173-214) for the
// fill_kv_cache path (src/transformer.rs:672-684 -> forward_layer with sl = N).
//
//   out[t][i] = sum_g ((sum_{k in g} xq[t][k] * wq[i][k]) as f32 * ws[i][g]) * xs[t][g]      (g ascending, from 0.0)
//
// The per-group scale is rank-1 PER 128-wide K group, so the int32 accumulator has to leave TMEM after every group
// (4 MMAs of K = 32): CTA tile 128 tokens x 128 output rows, K streamed one quantization group per pipeline stage;
//   warp 0      TMA producer   A tile [128 tok][128 B], B tile [128 rows][128 B], 128B-swizzled, 4-stage mbarrier ring
//   warp 1      MMA issuer     4 x tcgen05.mma.cta_group::1.kind::i8 (M128 N128 K32) per group into one of two TMEM
//                              accumulator buffers; tcgen05.commit frees the smem stage and publishes the buffer
//   warps 2..17 epilogue       tcgen05.ld 32x32b (thread = token row x 16/32 columns), s32 -> f32, * ws[col][g] * xs[row][g], added
//                              into f32 register accumulators in the reference's order -> results are bit-identical
//                              to the CPU path; double-buffered TMEM lets group g+1's MMAs run under group g's epilogue.
// The CUDA-core mini-epilogue runs after EVERY group (the scale is rank-1 per K group), so it bounds the kernel: it is written
// with packed f32x2 arithmetic and an exact integer->float conversion without I2F (2.5 issue slots per element per group
// instead of ~5 with a quarter-rate I2F; 3 per element), the tile's weight scales are staged once for all groups, and tiles come in three
// shapes so that every matrix fills the GPU: 128 x 128, 128 x 64 (Wo / W2: 2048 output rows), and the fused gate/up tile.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemv.cuh"   // glu_act

namespace lmrs {

constexpr int GEMM_M = 128, GEMM_K = 128, GEMM_STAGES = 4;
// epilogue warps: four per TMEM lane quadrant for the 128-column tiles (the per-group epilogue is a latency-bound instruction
// stream: 8 -> 16 warps measured 86 -> 68 us on the fused gate/up GEMM), two for the 64-column tiles (16 warps measured slower)
template <int BNM> __host__ __device__ constexpr int gemm_epi_warps() { return BNM == 128 ? 16 : 8; }
template <int BNM> __host__ __device__ constexpr int gemm_threads() { return 64 + 32 * gemm_epi_warps<BNM>(); }   // TMA warp, MMA warp, epilogue warps
constexpr int GEMM_TILE_BYTES = GEMM_M * GEMM_K;   // 16 KB A tile (128 token rows x one quantization group)
// BNM = accumulator columns of one MMA (output rows of w per CTA tile): 128 or 64.  Shared memory: operand ring, the tile's
// weight scales for ALL groups transposed to [group][column], barriers.
template <int BNM> inline size_t gemm_smem_bytes(int n) {
    return 1024 + (size_t)GEMM_STAGES * (GEMM_TILE_BYTES + BNM * GEMM_K) + (size_t)(n / GEMM_K) * BNM * 4 + 256;
}

struct GemmParams {
    int T, n, o;          // rows of x, input features, output rows of w
    const float* ws;      // dense [o][n/128] weight scales (file layout)
    const float* ws2;     // GLU: scales of the second matrix (w3)
    const float* xs;      // [T][n/128] activation scales
    // output: columns [0,c1) -> out0, [c1,c2) -> out1, [c2,o) -> out2 (segment boundaries are multiples of 128)
    float* out0; int ld0; int c1;
    float* out1; int ld1; int c2;
    float* out2; int ld2;
    int glu_epi;          // GLU: EPI_GLU_SILU / EPI_GLU_GELU (gemv.cuh)
    float neg_zero;       // -0.0f, passed at run time (see f2_mul_sep)
};

LMRS_DEVINL void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
LMRS_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major operand tile, rows of exactly 128 bytes, SWIZZLE_128B: 8-row atoms of 1024 B (SBO), version 1 (sm_100)
LMRS_DEVINL uint64_t umma_desc_sw128(const void* smem_tile) {
    const uint32_t addr = smem_u32(smem_tile);
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                          // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: next 8-row atom
    d |= (uint64_t)1 << 46;                          // descriptor version
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = s32, A = B = signed int8, both K-major, M = 128, N = BNM, dense, no saturate
template <int BNM> __host__ __device__ constexpr uint32_t umma_idesc_i8_m128() {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BNM >> 3) << 17) | ((uint32_t)(GEMM_M >> 4) << 24);
}
LMRS_DEVINL void umma_i8(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
LMRS_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
LMRS_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

LMRS_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
template <int N> LMRS_DEVINL void tmem_ld(uint32_t taddr, uint32_t (&v)[N]) {
    if constexpr (N == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
}

// ---- packed f32x2 arithmetic (Blackwell FADD2 / FMUL2: two IEEE round-to-nearest results per issue slot) ----------------
LMRS_DEVINL uint64_t f2_pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
LMRS_DEVINL void f2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
LMRS_DEVINL uint64_t f2_mul(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
LMRS_DEVINL uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// A product that must stay a SEPARATELY ROUNDED product when an add consumes it.  ptxas (12.9) contracts mul.rn.f32x2
// followed by add.rn.f32x2 into one FFMA2 -- a single rounding -- although both carry an explicit rounding modifier (it
// does not do that to the scalar forms).  Written as fma(a, b, -0.0) with the -0.0 taken from a kernel PARAMETER the
// product is exact (a*b + -0.0 rounds like a*b and keeps the sign of a zero product), costs the same FFMA2 issue slot,
// and cannot be contracted any further: the following add stays an FADD2.
LMRS_DEVINL uint64_t f2_mul_sep(uint64_t a, uint64_t b, uint64_t neg_zero2) {
    uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(neg_zero2)); return r;
}
// (ival as f32) for |ival| < 2^22 without I2F (a quarter-rate conversion that bounded the first build's epilogue): the
// integer is added into the mantissa of 1.5 * 2^23 and the constant subtracted again -- both steps exact.
// |ival| <= 128 * 127 * 127 = 2,064,512 for one quantization group.
constexpr int GEMM_MAGIC_I = 0x4B400000;
constexpr float GEMM_MAGIC_F = 12582912.0f;

// acc[0 .. N/2) (pairs) += ((v as f32) * ws) * xs for N consecutive accumulator columns: src/functional.rs:207's term and the
// ascending-group f32 accumulation, two columns per instruction
template <int N>
LMRS_DEVINL void gemm_epi_chunk(uint64_t* acc, const uint32_t (&v)[N], const float* ws32, uint64_t xs2, uint64_t neg_magic2, uint64_t nz2) {
    const float4* w4 = reinterpret_cast<const float4*>(ws32);
#pragma unroll
    for (int j = 0; j < N / 4; j++) {
        const float4 w = w4[j];   // the same address for every lane: one broadcast LDS.128
        const uint64_t f0 = f2_add(f2_pack(__int_as_float((int)v[4 * j] + GEMM_MAGIC_I), __int_as_float((int)v[4 * j + 1] + GEMM_MAGIC_I)), neg_magic2);
        const uint64_t f1 = f2_add(f2_pack(__int_as_float((int)v[4 * j + 2] + GEMM_MAGIC_I), __int_as_float((int)v[4 * j + 3] + GEMM_MAGIC_I)), neg_magic2);
        acc[2 * j] = f2_add(acc[2 * j], f2_mul_sep(f2_mul(f0, f2_pack(w.x, w.y)), xs2, nz2));
        acc[2 * j + 1] = f2_add(acc[2 * j + 1], f2_mul_sep(f2_mul(f1, f2_pack(w.z, w.w)), xs2, nz2));
    }
}

// BNM: accumulator columns per CTA tile (128 or 64).  GLU (BNM = 128): accumulator columns 0..63 are rows n0h..n0h+63 of
// w1 (gate), columns 64..127 the same rows of w3 (up); the epilogue writes act(gate) * up (src/transformer.rs:607-624)
// for 64 hidden columns, so gate/up never travel through HBM.
template <int BNM, bool GLU>
__global__ void __launch_bounds__(gemm_threads<BNM>(), 1)
gemm_q8_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_b2,
               const GemmParams p) {
    static_assert(BNM == 128 || BNM == 64, "tile width");
    constexpr int GEMM_EPI_WARPS = gemm_epi_warps<BNM>();
    static_assert(!GLU || BNM == 128, "GLU tiles pair 64 gate with 64 up columns");
    constexpr int B_TILE_BYTES = BNM * GEMM_K;
    constexpr int STAGE_BYTES = GEMM_TILE_BYTES + B_TILE_BYTES;
    constexpr int OUT_COLS = GLU ? 64 : BNM;                              // output columns (rows of w) per CTA
    extern __shared__ uint8_t gsm_raw[];
    // 1024-byte alignment for the 128B-swizzled operand tiles, computed on the SHARED-window offset and applied as pointer
    // arithmetic on gsm_raw: a round trip through uintptr_t makes the compiler forget the address space, and every access
    // behind it (the per-group scale loads of the epilogue) became a generic LD.E instead of LDS in the first builds
    uint8_t* gsm = gsm_raw + (((smem_u32(gsm_raw) + 1023u) & ~1023u) - smem_u32(gsm_raw));
    const int G = p.n / GEMM_K;
    uint8_t* tiles = gsm;                                                   // [STAGES][A 16 KB | B]
    float* ws_t = reinterpret_cast<float*>(tiles + (size_t)GEMM_STAGES * STAGE_BYTES);   // [G][BNM]
    uint64_t* full = reinterpret_cast<uint64_t*>(ws_t + (size_t)G * BNM);   // [STAGES]
    uint64_t* empty = full + GEMM_STAGES;                                   // [STAGES]
    uint64_t* tfull = empty + GEMM_STAGES;                                  // [2]
    uint64_t* tempty = tfull + 2;                                           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * GEMM_M, n0 = blockIdx.x * OUT_COLS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GEMM_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], GEMM_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) {   // TMEM: two BNM-column s32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BNM) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // ---- TMA producer ----
            for (int g = 0; g < G; g++) {
                const int s = g % GEMM_STAGES;
                mbar_wait(&empty[s], ((g / GEMM_STAGES) & 1) ^ 1);
                uint8_t* a_t = tiles + (size_t)s * STAGE_BYTES;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                tma_load_2d(a_t, &tm_a, g * GEMM_K, m0, &full[s]);
                if (GLU) {
                    tma_load_2d(a_t + GEMM_TILE_BYTES, &tm_b, g * GEMM_K, n0, &full[s]);
                    tma_load_2d(a_t + GEMM_TILE_BYTES + 64 * GEMM_K, &tm_b2, g * GEMM_K, n0, &full[s]);
                } else {
                    tma_load_2d(a_t + GEMM_TILE_BYTES, &tm_b, g * GEMM_K, n0, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---- MMA issuer ----
            const uint32_t idesc = umma_idesc_i8_m128<BNM>();
            for (int g = 0; g < G; g++) {
                const int s = g % GEMM_STAGES, b = g & 1;
                mbar_wait(&tempty[b], ((g >> 1) & 1) ^ 1);          // epilogue has drained this accumulator
                mbar_wait(&full[s], (g / GEMM_STAGES) & 1);         // operands have landed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint8_t* a_t = tiles + (size_t)s * STAGE_BYTES;
                const uint64_t da = umma_desc_sw128(a_t), db = umma_desc_sw128(a_t + GEMM_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < GEMM_K / 32; k++)   // advance 32 bytes along K inside the swizzle atom
                    umma_i8(tmem_base + b * BNM, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k > 0);
                umma_commit(&empty[s]);                              // smem stage reusable once these MMAs finished
                umma_commit(&tfull[b]);                              // accumulator complete
            }
        }
    } else {   // ---- epilogue: TMEM lane quadrant = warp % 4, column slice = (warp - 2) / 4; thread = one token row x CW columns per chunk ----
        constexpr int NSUB = GEMM_EPI_WARPS / 4;                     // column slices
        constexpr int NCH = GLU ? 2 : 1;                             // chunks per thread and group (GLU: gate slice + matching up slice)
        constexpr int CW = (GLU ? 64 : BNM) / NSUB;                  // columns per chunk: 32 (128-wide tiles) or 16
        constexpr int EPI_THREADS = 32 * GEMM_EPI_WARPS;
        static_assert(CW == 16 || CW == 32, "TMEM load shapes x16 / x32");
        const int quad = warp & 3;
        const int csub = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const int et = threadIdx.x - 64;                             // index among the epilogue threads
        const bool row_ok = m0 + row < p.T;
        auto col_of = [&](int c) { return (GLU ? c * 64 : 0) + csub * CW; };   // first accumulator column of chunk c
        // the tile's weight scales for every group, transposed to [g][column] (file layout is [row of w][g])
        for (int e = et; e < G * BNM; e += EPI_THREADS) {
            const int c = e / G, g = e - c * G;
            float v;
            if (GLU) v = c < 64 ? p.ws[(size_t)(n0 + c) * G + g] : p.ws2[(size_t)(n0 + c - 64) * G + g];
            else v = (n0 + c < p.o) ? p.ws[(size_t)(n0 + c) * G + g] : 0.0f;
            ws_t[g * BNM + c] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        uint64_t acc[NCH][CW / 2];
#pragma unroll
        for (int c = 0; c < NCH; c++)
#pragma unroll
            for (int j = 0; j < CW / 2; j++) acc[c][j] = 0ull;       // (+0.0f, +0.0f)
        const uint64_t neg_magic2 = f2_pack(-GEMM_MAGIC_F, -GEMM_MAGIC_F), nz2 = f2_pack(p.neg_zero, p.neg_zero);
        const float* xs_row = p.xs + (size_t)(row_ok ? m0 + row : 0) * G;
        float xs_next = row_ok ? xs_row[0] : 0.0f;
        for (int g = 0; g < G; g++) {
            const int b = g & 1;
            const float xsc = xs_next;
            if (g + 1 < G) xs_next = row_ok ? xs_row[g + 1] : 0.0f;   // in flight under this group's arithmetic
            const uint64_t xs2 = f2_pack(xsc, xsc);
            mbar_wait(&tfull[b], (g >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // all of this thread's accumulator words of the group are requested back to back; once they sit in registers the
            // TMEM buffer goes back to the MMA warp BEFORE the arithmetic, so group g+2's MMAs run under group g's epilogue
            uint32_t v[NCH][CW];
#pragma unroll
            for (int c = 0; c < NCH; c++)
                tmem_ld<CW>(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * BNM + col_of(c)), v[c]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[b]);
#pragma unroll
            for (int c = 0; c < NCH; c++) gemm_epi_chunk<CW>(acc[c], v[c], ws_t + g * BNM + col_of(c), xs2, neg_magic2, nz2);
        }
        if (row_ok) {
            if constexpr (GLU) {
                float4* o4 = reinterpret_cast<float4*>(p.out0 + (size_t)(m0 + row) * p.ld0 + n0 + csub * CW);
#pragma unroll
                for (int j = 0; j < CW / 4; j++) {
                    float g0, g1, g2, g3, u0, u1, u2, u3;
                    f2_unpack(acc[0][2 * j], g0, g1); f2_unpack(acc[0][2 * j + 1], g2, g3);
                    f2_unpack(acc[1][2 * j], u0, u1); f2_unpack(acc[1][2 * j + 1], u2, u3);
                    o4[j] = make_float4(__fmul_rn(glu_act(p.glu_epi, g0), u0), __fmul_rn(glu_act(p.glu_epi, g1), u1),
                                        __fmul_rn(glu_act(p.glu_epi, g2), u2), __fmul_rn(glu_act(p.glu_epi, g3), u3));
                }
            } else {
                float* dst; int ld, cbase;
                if (n0 < p.c1) { dst = p.out0; ld = p.ld0; cbase = 0; }
                else if (n0 < p.c2) { dst = p.out1; ld = p.ld1; cbase = p.c1; }
                else { dst = p.out2; ld = p.ld2; cbase = p.c2; }
                float4* o4 = reinterpret_cast<float4*>(dst + (size_t)(m0 + row) * ld + (n0 - cbase) + csub * CW);
#pragma unroll
                for (int j = 0; j < CW / 4; j++) {
                    float a0, a1, a2, a3;
                    f2_unpack(acc[0][2 * j], a0, a1); f2_unpack(acc[0][2 * j + 1], a2, a3);
                    o4[j] = make_float4(a0, a1, a2, a3);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BNM) : "memory");
    }
}

}  // namespace lmrs
