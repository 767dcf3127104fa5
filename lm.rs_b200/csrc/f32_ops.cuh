// f32_ops.cuh -- the unquantized matmul operators of the reference (`matmul` src/functional.rs:142-171 and
// `matmul_rest` :252-280), used by unquantized models and by the vision tower's patch embedding (src/vision.rs:264).
//
// One output element = one thread, evaluated by a __host__ __device__ function in the reference's exact f32 order, so
// that the SAME function body can be (and is: tests/test_f32_ops_host.py) compiled for the host and compared bit for
// bit with the CPU oracle without a GPU.  Device build: -fmad=false keeps every a*b+c as a separate multiply and add.
// Memory: thread i walks row i of w in 32-byte steps (one full sector per 8-lane chunk), so DRAM sectors are fully used
// even though a warp's accesses are not contiguous; these are operator-ABI kernels, not the decode hot path.
#pragma once
#include "exact_math.cuh"   // LMRS_HD

namespace lmrs {

// wide::f32x8::reduce_add as the oracle fixes it: ((p0+p4)+(p2+p6))+((p1+p5)+(p3+p7))
LMRS_HD float f32x8_reduce_add(const float* p) {
    const float s0 = p[0] + p[4], s1 = p[1] + p[5], s2 = p[2] + p[6], s3 = p[3] + p[7];
    const float d0 = s0 + s2, d1 = s1 + s3;
    return d0 + d1;
}

// `matmul` (:142-171): xout = 0; for every 8-element chunk: xout += (x_vec * w_vec).reduce_add(); the n % 8 tail is dropped
LMRS_HD float matmul_f32_element(const float* xr, const float* wr, int n) {
    float acc = 0.0f;
    for (int j = 0; j < n / 8; j++) {
        float p[8];
        for (int k = 0; k < 8; k++) p[k] = xr[j * 8 + k] * wr[j * 8 + k];
        acc += f32x8_reduce_add(p);
    }
    return acc;
}

// `matmul_rest` (:252-280): eight lane accumulators over the chunks, final_sum = 0 + reduce_add, then the tail -- which
// reads x[r] of the FIRST row whatever row is being computed (:273-275; kept)
LMRS_HD float matmul_rest_element(const float* xr, const float* x_row0, const float* wr, int n) {
    float lanes[8] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    const int n_simd = n / 8;
    for (int j = 0; j < n_simd; j++)
        for (int k = 0; k < 8; k++) lanes[k] += wr[j * 8 + k] * xr[j * 8 + k];
    float fs = 0.0f;
    fs += f32x8_reduce_add(lanes);
    for (int t = n_simd * 8; t < n; t++) fs += wr[t] * x_row0[t];
    return fs;
}

#ifdef __CUDACC__
// xout[r][i] for r < rows, i < o; rest = 0: `matmul` (o is a multiple of 4 here), rest = 1: `matmul_rest`
__global__ void matmul_f32_kernel(float* __restrict__ xout, const float* __restrict__ x, const float* __restrict__ w,
                                  const int rows, const int n, const int o, const int rest) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * o) return;
    const int r = (int)(idx / o), i = (int)(idx - (size_t)r * o);
    const float* xr = x + (size_t)r * n;
    const float* wr = w + (size_t)i * n;
    xout[idx] = rest ? matmul_rest_element(xr, x, wr, n) : matmul_f32_element(xr, wr, n);
}
#endif

}  // namespace lmrs
