// exact_math.cuh -- transcendental functions that must return the SAME BITS as the reference's libm calls.
//
// Why: lm.rs re-quantizes activations to int8/int4 four times per block (src/transformer.rs:427,553,596,633).
// Rounding is discontinuous, so a 1-ulp difference in an f32 activation occasionally flips a code, and the
// flip is amplified by the following layers far beyond the 1e-3 logits tolerance.  The only robust way to
// "match the CPU path" is to reproduce every f32 operation that feeds a quantizer bit for bit.  IEEE add /
// mul / div / sqrt are exact by construction (-fmad=false, __f*_rn); f32::exp is the one libm call on the
// path (softmax src/functional.rs:133, SiLU src/transformer.rs:617).  Rust's f32::exp lowers to the
// platform expf, i.e. glibc's (sysdeps/ieee754/flt-32/e_expf.c, glibc >= 2.27): a double-precision
// degree-3 polynomial around a 32-entry 2^(i/32) table, rounded once to float.  expf_glibc() below restates
// that algorithm with the constants of `__exp2f_data` (table = correctly rounded 2^(i/32) minus i<<47;
// polynomial and 32/ln2 constants read out of libm.so.6 2.39 and equal to the published source).
// tests/test_exact_math.py checks it against the host's glibc expf on tens of millions of inputs.
#pragma once
#include <stdint.h>
#include <string.h>
#ifdef __CUDACC__
#define LMRS_HD __host__ __device__ __forceinline__
#else
#define LMRS_HD static inline
#endif

namespace lmrs {

#ifdef __CUDA_ARCH__
__device__ static const uint64_t kExp2fTab[32] = {
#else
static const uint64_t kExp2fTab[32] = {
#endif
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

// T: the 32-entry table (kExp2fTab or a shared-memory copy of it: a kernel's first touch of the global table is an
// L2/HBM round trip on the critical path of every launch, so device code stages it in shared memory up front)
LMRS_HD float expf_glibc_t(float x, const uint64_t* T) {
    uint32_t ix;
#ifdef __CUDA_ARCH__
    ix = __float_as_uint(x);
#else
    memcpy(&ix, &x, 4);
#endif
    const uint32_t abstop = (ix >> 20) & 0x7ff;
    if (abstop >= 0x42b) {                        // |x| >= 88 or NaN  (top12(88.0f) = 0x42b)
        if (ix == 0xff800000u) return 0.0f;        // -inf
        if (abstop >= 0x7f8) return x + x;         // inf / NaN
        if (x > 88.72283172607421875f) return x * 3.4028234663852886e38f;   // overflow -> +inf  (x > 0x1.62e42ep6)
        if (x < -103.972076416015625f) return 0.0f;                         // underflow          (x < -0x1.9fe368p6)
    }
    const double xd = (double)x;
    const double InvLn2N = 0x1.71547652b82fep+5;   // 32/ln2
    const double Shift = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-20, C1 = 0x1.ebfce50fac4f3p-13, C2 = 0x1.62e42ff0c52d6p-6;
#ifdef __CUDA_ARCH__
    const double z = __dmul_rn(InvLn2N, xd);
    double kd = __dadd_rn(z, Shift);
    const uint64_t ki = (uint64_t)__double_as_longlong(kd);
    kd = __dsub_rn(kd, Shift);
    const double r = __dsub_rn(z, kd);
    const uint64_t t = T[ki & 31] + (ki << 47);
    const double s = __longlong_as_double((long long)t);
    const double zz = __dadd_rn(__dmul_rn(C0, r), C1);
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
    y = __dadd_rn(__dmul_rn(zz, r2), y);
    y = __dmul_rn(y, s);
    return __double2float_rn(y);
#else
    const double z = InvLn2N * xd;
    volatile double kdv = z + Shift;
    double kd = kdv;
    uint64_t ki;
    memcpy(&ki, &kd, 8);
    kd -= Shift;
    const double r = z - kd;
    const uint64_t t = T[ki & 31] + (ki << 47);
    double s;
    memcpy(&s, &t, 8);
    const double zz = C0 * r + C1;
    const double r2 = r * r;
    double y = C2 * r + 1.0;
    y = zz * r2 + y;
    y = y * s;
    return (float)y;
#endif
}

LMRS_HD float expf_glibc(float x) { return expf_glibc_t(x, kExp2fTab); }

}  // namespace lmrs
