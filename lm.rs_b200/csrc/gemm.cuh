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
//   warps 2..9  epilogue       tcgen05.ld 32x32b (thread = token row x 64 columns), s32 -> f32, * ws[col][g] * xs[row][g], added
//                              into 64 f32 register accumulators in the reference's order -> results are bit-identical
//                              to the CPU path; double-buffered TMEM lets group g+1's MMAs run under group g's epilogue.
// The CUDA-core mini-epilogue (4 instructions per element per group) is about twice the MMA time of a tile: the
// honest int8 tensor-core roofline fraction of this formulation is bounded by it (hard part 1).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace lmrs {

constexpr int GEMM_M = 128, GEMM_N = 128, GEMM_K = 128, GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quadrant, 64 columns each)
constexpr int GEMM_TILE_BYTES = GEMM_M * GEMM_K;   // 16 KB per operand tile
constexpr size_t GEMM_SMEM = 1024 + (size_t)GEMM_STAGES * 2 * GEMM_TILE_BYTES + 2 * GEMM_N * 4 + 256;

struct GemmParams {
    int T, n, o;          // rows of x, input features, output rows of w
    const float* ws;      // dense [o][n/128] weight scales (file layout)
    const float* xs;      // [T][n/128] activation scales
    // output: columns [0,c1) -> out0, [c1,c2) -> out1, [c2,o) -> out2 (segment boundaries are multiples of 128)
    float* out0; int ld0; int c1;
    float* out1; int ld1; int c2;
    float* out2; int ld2;
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
// instruction descriptor: D = s32, A = B = signed int8, both K-major, M = 128, N = 128, dense, no saturate
__host__ __device__ constexpr uint32_t umma_idesc_i8_m128_n128() {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(GEMM_N >> 3) << 17) | ((uint32_t)(GEMM_M >> 4) << 24);
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

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_q8_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmParams p) {
    extern __shared__ uint8_t gsm_raw[];
    uint8_t* gsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gsm_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* tiles = gsm;                                                   // [STAGES][A 16 KB | B 16 KB]
    float* ws_s = reinterpret_cast<float*>(tiles + (size_t)GEMM_STAGES * 2 * GEMM_TILE_BYTES);   // [2][128]
    uint64_t* full = reinterpret_cast<uint64_t*>(ws_s + 2 * GEMM_N);        // [STAGES]
    uint64_t* empty = full + GEMM_STAGES;                                   // [STAGES]
    uint64_t* tfull = empty + GEMM_STAGES;                                  // [2]
    uint64_t* tempty = tfull + 2;                                           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * GEMM_M, n0 = blockIdx.x * GEMM_N;
    const int G = p.n / GEMM_K;

    if (threadIdx.x == 0) {
        for (int s = 0; s < GEMM_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; b++) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 8); }
        fence_barrier_init();
    }
    if (warp == 1) {   // TMEM: 256 columns = two 128-column s32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
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
                uint8_t* a_t = tiles + (size_t)s * 2 * GEMM_TILE_BYTES;
                mbar_expect_tx(&full[s], 2 * GEMM_TILE_BYTES);
                tma_load_2d(a_t, &tm_a, g * GEMM_K, m0, &full[s]);
                tma_load_2d(a_t + GEMM_TILE_BYTES, &tm_b, g * GEMM_K, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {   // ---- MMA issuer ----
            const uint32_t idesc = umma_idesc_i8_m128_n128();
            for (int g = 0; g < G; g++) {
                const int s = g % GEMM_STAGES, b = g & 1;
                mbar_wait(&tempty[b], ((g >> 1) & 1) ^ 1);          // epilogue has drained this accumulator
                mbar_wait(&full[s], (g / GEMM_STAGES) & 1);         // operands have landed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint8_t* a_t = tiles + (size_t)s * 2 * GEMM_TILE_BYTES;
                const uint64_t da = umma_desc_sw128(a_t), db = umma_desc_sw128(a_t + GEMM_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < GEMM_K / 32; k++)   // advance 32 bytes along K inside the swizzle atom
                    umma_i8(tmem_base + b * GEMM_N, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k > 0);
                umma_commit(&empty[s]);                              // smem stage reusable once these MMAs finished
                umma_commit(&tfull[b]);                              // accumulator complete
            }
        }
    } else {   // ---- epilogue: warps 2..9, TMEM lane quadrant = warp % 4, thread = one token row x 64 columns ----
        constexpr int EC = GEMM_N / 2;                               // columns per epilogue thread
        const int quad = warp & 3;
        const int chalf = (warp - 2) >> 2;                           // 0: columns 0..63, 1: columns 64..127
        const int row = quad * 32 + lane;
        const int et = threadIdx.x - 64;                             // 0..255 among the epilogue threads
        const bool row_ok = m0 + row < p.T;
        float acc[EC];
#pragma unroll
        for (int j = 0; j < EC; j++) acc[j] = 0.0f;
        for (int g = 0; g < G; g++) {
            const int b = g & 1;
            // stage this group's 128 column scales (file layout [o][G]) and fetch my row's activation scale
            if (et < GEMM_N) ws_s[b * GEMM_N + et] = (n0 + et < p.o) ? p.ws[(size_t)(n0 + et) * G + g] : 0.0f;
            const float xsc = row_ok ? p.xs[(size_t)(m0 + row) * G + g] : 0.0f;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(&tfull[b], (g >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < EC / 32; c++) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * GEMM_N + chalf * EC + c * 32), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const float t = __fmul_rn(__fmul_rn((float)(int)v[j], ws_s[b * GEMM_N + chalf * EC + c * 32 + j]), xsc);
                    acc[c * 32 + j] = __fadd_rn(acc[c * 32 + j], t);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[b]);
        }
        if (row_ok) {
            float* dst; int ld, cbase;
            if (n0 < p.c1) { dst = p.out0; ld = p.ld0; cbase = 0; }
            else if (n0 < p.c2) { dst = p.out1; ld = p.ld1; cbase = p.c1; }
            else { dst = p.out2; ld = p.ld2; cbase = p.c2; }
            float4* o4 = reinterpret_cast<float4*>(dst + (size_t)(m0 + row) * ld + (n0 - cbase) + chalf * EC);
#pragma unroll
            for (int j = 0; j < EC / 4; j++) o4[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    }
}

}  // namespace lmrs
