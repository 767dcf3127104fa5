/*
 * oracle/lmrs_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT.  See oracle/lmrs_ref.h.
 *
 * CPU restatement of the lm.rs forward path, one C function per reference function, same
 * operation order.  Every function cites the reference file:line it follows (paths relative
 * to the reference repo root).  Build with -ffp-contract=off: Rust never contracts a*b+c
 * into an FMA, so neither may this file.
 *
 * PARITY UNPINNED: no golden vectors exist in the reference and its Rust crate cannot be
 * compiled here, so this restatement is validated only by hand-derived KATs
 * (tests/test_oracle_kat.py) and the reference's Python exporter (tests/golden/).
 *
 * Third-party arithmetic that is NOT in /root/reference: crate `wide` ^0.7.28
 * (Cargo.toml:14; no Cargo.lock) supplies f32x8/i32x8.  Integer lanes are order-independent.
 * For f32x8::reduce_add this file uses the crate's AVX sequence as published
 * (add hi/lo 128-bit halves, then movehl, then lane 1):
 *     ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7))
 * which could not be re-checked offline; the 1e-3 logits tolerance absorbs any other order.
 * f32::exp/powf/sin/cos and f64::tanh in Rust lower to the platform libm -- the same glibc
 * this file calls.
 */
#define _GNU_SOURCE
#include "lmrs_ref.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char g_err[256];
const char* lmrs_ref_last_error(void) { return g_err; }
static int fail(const char* msg) { snprintf(g_err, sizeof g_err, "%s", msg); return 1; }

void lmrs_ref_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int lmrs_ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- scalar semantics of Rust casts ------------------------------------------------------ */
/* `f32 as i8`: saturating, NaN -> 0 (Rust reference, "as" casts; used at src/quantization.rs:63) */
static inline int8_t f32_as_i8(float v) {
    if (v != v) return 0;
    if (v <= -128.0f) return -128;
    if (v >= 127.0f) return 127;
    return (int8_t)v; /* truncation toward zero; callers pass already-rounded values */
}
/* `f32 as u8`: saturating, NaN -> 0 (src/quantization.rs:89-90) */
static inline uint8_t f32_as_u8(float v) {
    if (v != v) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
/* wide::f32x8::reduce_add, AVX order (see file header) */
static inline float reduce_add8(const float a[8]) {
    float s0 = a[0] + a[4], s1 = a[1] + a[5], s2 = a[2] + a[6], s3 = a[3] + a[7];
    float d0 = s0 + s2, d1 = s1 + s3;
    return d0 + d1;
}

/* ---- src/functional.rs ------------------------------------------------------------------- */

/* src/functional.rs:48-78 */
void lmrs_ref_rmsnorm(float* o, const float* x, const float* w, int size, float eps, int add_unit_offset) {
    int n_simd = size / 8;
    float lanes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < n_simd; j++)
        for (int k = 0; k < 8; k++) lanes[k] += x[j * 8 + k] * x[j * 8 + k];
    float ss = reduce_add8(lanes);
    ss /= (float)size;
    ss += eps;
    ss = 1.0f / sqrtf(ss);
    for (int j = 0; j < n_simd * 8; j++) {
        float t = ss * x[j];
        o[j] = add_unit_offset ? (1.0f + w[j]) * t : w[j] * t;
    }
}

/* src/functional.rs:122-140 */
/* src/functional.rs:80-114: mean and variance in eight lane partials each (x[8j+k] / (x[8j+k]-mean)^2 walked over j),
 * wide's reduce_add order (see the header comment), / size, + eps, 1/sqrt; o = ((x - mean) * inv_std) * w + b with every
 * operation rounded separately (operator overloads of wide::f32x8: no fused multiply-add).  Elements beyond size/8*8
 * are not written. */
void lmrs_ref_layernorm(float* o, const float* x, const float* w, const float* b, int size, float eps) {
    int n8 = size / 8;
    float m[8] = {0}, v[8] = {0};
    for (int j = 0; j < n8; j++)
        for (int k = 0; k < 8; k++) m[k] += x[8 * j + k];
    float mean = (((m[0] + m[4]) + (m[2] + m[6])) + ((m[1] + m[5]) + (m[3] + m[7]))) / (float)size;
    for (int j = 0; j < n8; j++)
        for (int k = 0; k < 8; k++) { float d = x[8 * j + k] - mean; float d2 = d * d; v[k] += d2; }
    float var = (((v[0] + v[4]) + (v[2] + v[6])) + ((v[1] + v[5]) + (v[3] + v[7]))) / (float)size + eps;
    float inv_std = 1.0f / sqrtf(var);
    for (int i = 0; i < n8 * 8; i++) {
        float nrm = (x[i] - mean) * inv_std;
        float t = nrm * w[i];
        o[i] = t + b[i];
    }
}

void lmrs_ref_softmax(float* x, int n) {
    float sum = 0.0f, max_val = x[0];
    for (int i = 0; i < n; i++)
        if (x[i] > max_val) max_val = x[i];
    for (int i = 0; i < n; i++) {
        x[i] = expf(x[i] - max_val);
        sum += x[i];
    }
    for (int i = 0; i < n; i++) x[i] /= sum;
}

/* src/functional.rs:142-171 (o % 4 trailing outputs and n % 8 trailing inputs are dropped) */
void lmrs_ref_matmul_f32(float* xout, const float* x, const float* w, int rows, int n, int o) {
    int n_simd = n / 8;
    int o4 = o / 4 * 4;
#pragma omp parallel for collapse(2) schedule(static)
    for (int r = 0; r < rows; r++) {
        for (int i = 0; i < o4; i++) {
            const float* xr = x + (size_t)r * n;
            const float* wr = w + (size_t)i * n;
            float acc = 0.0f;
            for (int j = 0; j < n_simd; j++) {
                float p[8];
                for (int k = 0; k < 8; k++) p[k] = xr[j * 8 + k] * wr[j * 8 + k];
                acc += reduce_add8(p);
            }
            xout[(size_t)r * o + i] = acc;
        }
    }
}

/* src/functional.rs:252-280, INCLUDING the tail bug at :273-275 (x[r], not x[xi+r]) */
void lmrs_ref_matmul_rest(float* xout, const float* x, const float* w, int rows, int n, int o) {
    int n_simd = n / 8, rest = n_simd * 8;
#pragma omp parallel for collapse(2) schedule(static)
    for (int r = 0; r < rows; r++) {
        for (int i = 0; i < o; i++) {
            const float* xr = x + (size_t)r * n;
            const float* wr = w + (size_t)i * n;
            float lanes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < n_simd; j++)
                for (int k = 0; k < 8; k++) lanes[k] += wr[j * 8 + k] * xr[j * 8 + k];
            float fs = 0.0f;
            fs += reduce_add8(lanes);
            for (int t = rest; t < n; t++) fs += wr[t] * x[t]; /* row 0 of x: reference bug kept */
            xout[(size_t)r * o + i] = fs;
        }
    }
}

/* src/functional.rs:173-214.  Integer part exact (order-free); f32 part: per group, ascending,
 * xout += ((ival as f32) * ws) * xs, starting from 0.  Outputs beyond o/4*4 are not written
 * (par_chunks_exact_mut(4), :179). */
static void matmul_q8_kshards(float* xout, const int8_t* xq, const float* xs, const int8_t* wq, const float* ws,
                              int rows, int n, int o, int gs, int shards);
void lmrs_ref_matmul_q8(float* xout, const int8_t* xq, const float* xs, const int8_t* wq, const float* ws,
                        int rows, int n, int o, int gs) {
    matmul_q8_kshards(xout, xq, xs, wq, ws, rows, n, o, gs, 1);
}
/* `shards` > 1 is NOT a reference feature: it restates the summation order of lmrs_b200's row-sharded N-GPU mode
 *  so that N-GPU runs can be checked bit for bit: the K groups are split into `shards` equal
 * contiguous ranges, each range is accumulated from 0.0 in ascending order (one GPU's partial), and the partials are
 * added in ascending rank order starting from rank 0's. */
static void matmul_q8_kshards(float* xout, const int8_t* xq, const float* xs, const int8_t* wq, const float* ws,
                              int rows, int n, int o, int gs, int shards) {
    int o4 = o / 4 * 4;
    int ng = n / gs; /* (0..=(n-gs)).step_by(gs) */
#pragma omp parallel for collapse(2) schedule(static)
    for (int r = 0; r < rows; r++) {
        for (int i4 = 0; i4 < o4; i4 += 4) {
            const int8_t* xr = xq + (size_t)r * n;
            for (int u = 0; u < 4; u++) {
                int i = i4 + u;
                const int8_t* wr = wq + (size_t)i * n;
                float total = 0.0f;
                for (int sh = 0; sh < shards; sh++) {
                    float acc = 0.0f;
                    for (int g = sh * (ng / shards); g < (sh + 1 == shards ? ng : (sh + 1) * (ng / shards)); g++) {
                        int32_t ival = 0;
                        const int8_t* xa = xr + g * gs;
                        const int8_t* wa = wr + g * gs;
                        for (int k = 0; k < gs / 8 * 8; k++) ival += (int32_t)xa[k] * (int32_t)wa[k];
                        float t = (float)ival * ws[((size_t)i * n + (size_t)g * gs) / gs];
                        t = t * xs[((size_t)r * n + (size_t)g * gs) / gs];
                        acc += t;
                    }
                    total = sh == 0 ? acc : total + acc;
                }
                xout[(size_t)r * o + i] = total;
            }
        }
    }
}

/* src/functional.rs:216-250.  4-bit weights AND activations, both stored nibble+8, low nibble =
 * even element.  Reference bug (:224 `xi = j*n` although packed rows are n/2 bytes) makes rows>0
 * undefined in the reference; here rows>0 index x row-wise (xq + r*n/2, xs + r*n/gs), i.e. the
 * result every row would get if it were passed alone as row 0. */
static void matmul_q4_kshards(float* xout, const uint8_t* xq, const float* xs, const uint8_t* wq, const float* ws,
                              int rows, int n, int o, int gs, int shards);
void lmrs_ref_matmul_q4(float* xout, const uint8_t* xq, const float* xs, const uint8_t* wq, const float* ws,
                        int rows, int n, int o, int gs) {
    matmul_q4_kshards(xout, xq, xs, wq, ws, rows, n, o, gs, 1);
}
static void matmul_q4_kshards(float* xout, const uint8_t* xq, const float* xs, const uint8_t* wq, const float* ws,
                              int rows, int n, int o, int gs, int shards) {
    int gb = gs / 2;     /* bytes per group */
    int ng = (n / 2) / gb;
#pragma omp parallel for collapse(2) schedule(static)
    for (int r = 0; r < rows; r++) {
        for (int i = 0; i < o; i++) {
            const uint8_t* xr = xq + (size_t)r * (n / 2);
            const float* xsr = xs + (size_t)r * (n / gs);
            const uint8_t* wr = wq + (size_t)i * (n / 2);
            float total = 0.0f;
            for (int sh = 0; sh < shards; sh++) {
                float acc = 0.0f; /* Iterator::sum::<f32>() */
                for (int g = sh * (ng / shards); g < (sh + 1 == shards ? ng : (sh + 1) * (ng / shards)); g++) {
                    int32_t ival = 0;
                    for (int k = 0; k < gb / 8 * 8; k++) {
                        int xb = xr[g * gb + k], wb = wr[g * gb + k];
                        ival += ((xb & 0x0F) - 8) * ((wb & 0x0F) - 8);
                        ival += (((xb & 0xF0) >> 4) - 8) * (((wb & 0xF0) >> 4) - 8);
                    }
                    float t = (float)ival * ws[((size_t)i * (n / 2) + (size_t)g * gb) / gb];
                    t = t * xsr[g];
                    acc += t;
                }
                total = sh == 0 ? acc : total + acc;
            }
            xout[(size_t)r * o + i] = total;
        }
    }
}

/* ---- src/quantization.rs ----------------------------------------------------------------- */

/* src/quantization.rs:44-67: scale = max|x| / 127; q = round_half_away(x / scale) as i8
 * (all-zero group: x/0 = NaN -> 0). Serial in the reference. */
void lmrs_ref_quantize_q8(int8_t* q, float* s, const float* x, int n, int gs) {
    int ng = n / gs;
    for (int g = 0; g < ng; g++) {
        float wmax = 0.0f;
        for (int i = 0; i < gs; i++) {
            float v = fabsf(x[g * gs + i]);
            if (v > wmax) wmax = v;
        }
        float scale = wmax / 127.0f;
        s[g] = scale;
        for (int i = 0; i < gs; i++) q[g * gs + i] = f32_as_i8(roundf(x[g * gs + i] / scale));
    }
}

/* src/quantization.rs:69-95: scale = max|x| / -8; nibble = clamp(round(x/scale + 8) as u8, 0, 15);
 * byte = even | odd << 4. */
void lmrs_ref_quantize_q4(uint8_t* q, float* s, const float* x, int n, int gs) {
    int ng = n / gs;
    for (int g = 0; g < ng; g++) {
        float wmax = 0.0f;
        for (int i = 0; i < gs; i++) {
            float v = fabsf(x[g * gs + i]);
            if (v > wmax) wmax = v;
        }
        float scale = wmax / -8.0f;
        s[g] = scale;
        for (int i = 0; i < gs / 2; i++) {
            float a = x[g * gs + i * 2] / scale, b = x[g * gs + i * 2 + 1] / scale;
            uint8_t qa = f32_as_u8(roundf(a + 8.0f)), qb = f32_as_u8(roundf(b + 8.0f));
            if (qa > 15) qa = 15;
            if (qb > 15) qb = 15;
            q[g * gs / 2 + i] = (uint8_t)(qa | (qb << 4));
        }
    }
}

/* src/quantization.rs:17-42 */
void lmrs_ref_dequantize(float* x, const void* qv, const float* s, int n, int gs, int q_type) {
    if (q_type == 1) {
        const int8_t* q = (const int8_t*)qv;
        for (int i = 0; i < n; i++) x[i] = (float)q[i] * s[i / gs];
    } else if (q_type == 2) {
        const uint8_t* q = (const uint8_t*)qv;
        for (int i = 0; i < n / 2; i++) {
            int a = (q[i] & 0x0F) - 8, b = ((q[i] & 0xF0) >> 4) - 8;
            float scale = s[(i * 2) / gs];
            x[i * 2] = (float)a * scale;
            x[i * 2 + 1] = (float)b * scale;
        }
    }
}

/* ---- src/transformer.rs ------------------------------------------------------------------ */

typedef struct { const void* q; const float* s; } qt_t;

struct lmrs_ref {
    lmrs_ref_args_t args;
    const uint8_t* data;
    size_t len;
    int quantized;
    /* f32 model */
    const float *emb_f, *wq_f, *wk_f, *wv_f, *wo_f, *w1_f, *w2_f, *w3_f, *lm_head_f;
    /* quantized model: per-layer views (init_param_quant, src/transformer.rs:24-48) */
    qt_t emb_q, lm_head_q;
    qt_t *wq_q, *wk_q, *wv_q, *wo_q, *w1_q, *w2_q, *w3_q;
    const float *rms_att, *rms_post_att, *rms_pre_ffn, *rms_post_ffn, *rms_final;
    /* state (src/transformer.rs:116-125,299-305) */
    float* logits;
    float *key_cache, *value_cache;
};

static const float* take_f32(const uint8_t* data, size_t len, size_t* off, size_t count, int* err) {
    size_t bytes = count * 4;
    if (*off + bytes > len) { *err = 1; return NULL; }
    if (((uintptr_t)(data + *off)) % 4 != 0) { *err = 2; return NULL; } /* functional.rs:20-25 */
    const float* p = (const float*)(data + *off);
    *off += bytes;
    return p;
}

/* src/transformer.rs:24-48 */
static qt_t* take_quant(const uint8_t* data, size_t len, size_t* off, uint32_t n, size_t size_each, uint32_t gs,
                        int q_type, int* err) {
    qt_t* res = (qt_t*)calloc(n, sizeof(qt_t));
    size_t groups = size_each / gs;
    size_t size = q_type == 2 ? size_each / 2 : size_each;
    for (uint32_t i = 0; i < n; i++) {
        if (*off + size > len) { *err = 1; return res; }
        res[i].q = data + *off;
        *off += size;
        res[i].s = take_f32(data, len, off, groups, err);
        if (*err) return res;
    }
    return res;
}

/* src/transformer.rs:134-314 */
int lmrs_ref_create(const uint8_t* data, size_t len, lmrs_ref_t** out, size_t* end_offset) {
    if (len < 256 || data[0] != 0x6c || data[1] != 0x6d || data[2] != 0x72 || data[3] != 0x73)
        return fail("Model not in lm.rs format."); /* :135 */
    lmrs_ref_t* m = (lmrs_ref_t*)calloc(1, sizeof *m);
    memcpy(&m->args, data + 8, sizeof(lmrs_ref_args_t)); /* :141-145 */
    lmrs_ref_args_t* c = &m->args;
    if (c->seq_len > 8192) c->seq_len = 8192; /* :158-160 */
    m->data = data;
    m->len = len;
    m->quantized = c->q_type != 0;
    size_t off = 256; /* :151 */
    int err = 0;
    size_t L = c->n_layers, dim = c->dim, hd = c->hidden_dim;
    size_t att_dim = (size_t)c->n_heads * c->head_size, kv_dim = (size_t)c->head_size * c->n_kv_heads;
    int gemma = c->model_type == 0, phi = c->model_type == 2;
    if (c->q_type > 2 || c->model_type > 2) { free(m); return fail("bad q_type/model_type"); }
    if (m->quantized && (c->group_size == 0 || dim % c->group_size || hd % c->group_size || att_dim % c->group_size)) {
        free(m);
        return fail("group_size does not divide dim/hidden_dim/att_dim");
    }

    if (!m->quantized) { /* :169-237 */
        m->emb_f = take_f32(data, len, &off, (size_t)c->vocab_size * dim, &err);
        m->rms_att = take_f32(data, len, &off, L * dim, &err);
        m->wq_f = take_f32(data, len, &off, L * dim * att_dim, &err);
        m->wk_f = take_f32(data, len, &off, L * dim * kv_dim, &err);
        m->wv_f = take_f32(data, len, &off, L * dim * kv_dim, &err);
        m->wo_f = take_f32(data, len, &off, L * dim * att_dim, &err);
        m->rms_post_att = take_f32(data, len, &off, L * dim, &err);
        if (gemma) m->rms_pre_ffn = take_f32(data, len, &off, L * dim, &err);
        m->w1_f = take_f32(data, len, &off, L * dim * hd, &err);
        m->w2_f = take_f32(data, len, &off, L * dim * hd, &err);
        m->w3_f = take_f32(data, len, &off, L * dim * hd, &err);
        if (gemma) m->rms_post_ffn = take_f32(data, len, &off, L * dim, &err);
        m->rms_final = take_f32(data, len, &off, dim, &err);
        if (phi) m->lm_head_f = take_f32(data, len, &off, dim * c->vocab_size, &err);
    } else { /* :241-270 */
        uint32_t gs = c->group_size;
        int qt = c->q_type;
        qt_t* e = take_quant(data, len, &off, 1, (size_t)c->vocab_size * dim, gs, qt, &err);
        m->emb_q = e[0];
        free(e);
        m->rms_att = take_f32(data, len, &off, L * dim, &err);
        m->wq_q = take_quant(data, len, &off, L, dim * att_dim, gs, qt, &err);
        m->wk_q = take_quant(data, len, &off, L, dim * kv_dim, gs, qt, &err);
        m->wv_q = take_quant(data, len, &off, L, dim * kv_dim, gs, qt, &err);
        m->wo_q = take_quant(data, len, &off, L, dim * att_dim, gs, qt, &err);
        m->rms_post_att = take_f32(data, len, &off, L * dim, &err);
        if (gemma) m->rms_pre_ffn = take_f32(data, len, &off, L * dim, &err);
        m->w1_q = take_quant(data, len, &off, L, dim * hd, gs, qt, &err);
        m->w2_q = take_quant(data, len, &off, L, dim * hd, gs, qt, &err);
        m->w3_q = take_quant(data, len, &off, L, dim * hd, gs, qt, &err);
        if (gemma) m->rms_post_ffn = take_f32(data, len, &off, L * dim, &err);
        m->rms_final = take_f32(data, len, &off, dim, &err);
        if (phi) {
            qt_t* h = take_quant(data, len, &off, 1, dim * c->vocab_size, gs, qt, &err);
            m->lm_head_q = h[0];
            free(h);
        }
    }
    if (err) {
        lmrs_ref_destroy(m);
        return fail(err == 1 ? "file truncated" : "Data was not aligned correctly");
    }
    m->logits = (float*)calloc(c->vocab_size, sizeof(float));
    m->key_cache = (float*)calloc(L * c->seq_len * kv_dim, sizeof(float));
    m->value_cache = (float*)calloc(L * c->seq_len * kv_dim, sizeof(float));
    if (!m->logits || !m->key_cache || !m->value_cache) {
        lmrs_ref_destroy(m);
        return fail("out of memory");
    }
    *out = m;
    if (end_offset) *end_offset = off;
    return 0;
}

void lmrs_ref_destroy(lmrs_ref_t* m) { /* Drop, src/transformer.rs:688-711 */
    if (!m) return;
    free(m->wq_q); free(m->wk_q); free(m->wv_q); free(m->wo_q);
    free(m->w1_q); free(m->w2_q); free(m->w3_q);
    free(m->logits); free(m->key_cache); free(m->value_cache);
    free(m);
}

int lmrs_ref_args(const lmrs_ref_t* m, lmrs_ref_args_t* out) { *out = m->args; return 0; }
const float* lmrs_ref_key_cache(const lmrs_ref_t* m) { return m->key_cache; }
const float* lmrs_ref_value_cache(const lmrs_ref_t* m) { return m->value_cache; }

/* One embedding row of the f32 table the reference builds at load by dequantize()
 * (src/transformer.rs:243-245): value = q as f32 * s -- computed on demand, bit-identical. */
static void embedding_row(const lmrs_ref_t* m, uint32_t token, float* out) {
    const lmrs_ref_args_t* c = &m->args;
    size_t dim = c->dim;
    if (!m->quantized) {
        memcpy(out, m->emb_f + (size_t)token * dim, dim * 4);
    } else if (c->q_type == 1) {
        const int8_t* q = (const int8_t*)m->emb_q.q + (size_t)token * dim;
        const float* s = m->emb_q.s + (size_t)token * dim / c->group_size;
        for (size_t i = 0; i < dim; i++) out[i] = (float)q[i] * s[i / c->group_size];
    } else {
        const uint8_t* q = (const uint8_t*)m->emb_q.q + (size_t)token * dim / 2;
        const float* s = m->emb_q.s + (size_t)token * dim / c->group_size;
        for (size_t i = 0; i < dim / 2; i++) {
            int a = (q[i] & 0x0F) - 8, b = ((q[i] & 0xF0) >> 4) - 8;
            float scale = s[(i * 2) / c->group_size];
            out[i * 2] = (float)a * scale;
            out[i * 2 + 1] = (float)b * scale;
        }
    }
}

/* src/transformer.rs:445-478: frequency of rotary pair j, plus the PHI magnitude scale */
void lmrs_ref_rope_freq(int model_type, float rope_theta, int head_size, int j, float* freq_out, float* mscale) {
    static const double short_factor[48] = { /* :473, su-scaled rope of Phi-3.5 */
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
    uint32_t head_dim = (uint32_t)j * 2;
    float freq = 1.0f / powf(rope_theta, (float)head_dim / (float)head_size); /* :447 */
    float scaling_factor = 1.0f;
    if (model_type == 1) { /* LLAMA :451-470, llama-3 constants hard-coded */
        float wavelen = (2.0f * 3.14159265358979323846f) / freq;
        float factor = 32.0f, low_freq_factor = 1.0f, high_freq_factor = 4.0f, old_context_len = 8192.0f;
        float low_freq_wavelen = old_context_len / low_freq_factor;
        float high_freq_wavelen = old_context_len / high_freq_factor;
        if (wavelen > low_freq_wavelen) {
            freq /= factor;
        } else if (wavelen <= low_freq_wavelen && wavelen >= high_freq_wavelen) {
            float smooth = (old_context_len / wavelen - low_freq_factor) / (high_freq_factor - low_freq_factor);
            freq = (1.0f - smooth) * freq / factor + smooth * freq;
        }
    }
    if (model_type == 2) { /* PHI :472-478 */
        freq *= (float)(1.0 / short_factor[j % 48]);
        float scale = 131072.0f / 4096.0f;
        scaling_factor = sqrtf(1.0f + logf(scale) / logf(4096.0f));
    }
    *freq_out = freq;
    *mscale = scaling_factor;
}

/* quantize + matmul dispatch used 4x per layer (src/transformer.rs:424-438,550-558,593-603,630-638) */
/* test access to the k-shard accumulation order (tests/test_oracle_kat.py pins it against explicit column-slice partial sums) */
void lmrs_ref_matmul_q8_kshards(float* xout, const int8_t* xq, const float* xs, const int8_t* wq, const float* ws,
                                int rows, int n, int o, int gs, int shards) {
    matmul_q8_kshards(xout, xq, xs, wq, ws, rows, n, o, gs, shards < 1 ? 1 : shards);
}
static int g_kshards = 1;   /* see matmul_q8_kshards: models lmrs_b200's N-GPU partial-sum order for Wo / W2 (not a reference feature) */
void lmrs_ref_set_kshards(int n) { g_kshards = n < 1 ? 1 : n; }
static void qmatmul_sh(const lmrs_ref_t* m, float* out, const float* in, const qt_t* w, int rows, int n, int o, int shards);
static void qmatmul(const lmrs_ref_t* m, float* out, const float* in, const qt_t* w, int rows, int n, int o) {
    qmatmul_sh(m, out, in, w, rows, n, o, 1);
}
static void qmatmul_sh(const lmrs_ref_t* m, float* out, const float* in, const qt_t* w, int rows, int n, int o, int shards) {
    int gs = m->args.group_size;
    if (m->args.q_type == 1) {
        int8_t* q = (int8_t*)malloc((size_t)rows * n);
        float* s = (float*)malloc((size_t)rows * n / gs * 4);
        lmrs_ref_quantize_q8(q, s, in, rows * n, gs);
        matmul_q8_kshards(out, q, s, (const int8_t*)w->q, w->s, rows, n, o, gs, shards);
        free(q); free(s);
    } else {
        uint8_t* q = (uint8_t*)malloc((size_t)rows * n / 2);
        float* s = (float*)malloc((size_t)rows * n / gs * 4);
        lmrs_ref_quantize_q4(q, s, in, rows * n, gs);
        matmul_q4_kshards(out, q, s, (const uint8_t*)w->q, w->s, rows, n, o, gs, shards);
        free(q); free(s);
    }
}

/* src/transformer.rs:388-657 */
static int forward_layer(lmrs_ref_t* m, float* x, uint32_t sl, uint32_t l, uint32_t pos) {
    const lmrs_ref_args_t p = m->args;
    const uint32_t dim = p.dim, head_size = p.head_size;
    const uint32_t att_dim = p.n_heads * head_size, kv_dim = head_size * p.n_kv_heads;
    const uint32_t kv_mul = p.n_heads / p.n_kv_heads, hidden_dim = p.hidden_dim;
    const int gemma = p.model_type == 0;
    if (pos + sl > p.seq_len) return fail("position out of range (seq_len is clamped to 8192)");
    /* :501-503 walks `embeddings` (sl*dim floats) in chunks of att_dim: floor(sl*dim/att_dim) chunks.  One chunk per token
     * is what the code means; as soon as there is one more (sl*(dim-att_dim) >= att_dim) its `sq` slice is out of bounds
     * and the reference panics.  Below that the batch is well defined (Gemma-2-2B: up to 7 rows). */
    if (att_dim < dim && (size_t)sl * (dim - att_dim) >= att_dim)
        return fail("sl*(dim-att_dim) >= att_dim is out of bounds in the reference (src/transformer.rs:501-503)");

    size_t total = (size_t)sl * dim, total_h = (size_t)sl * hidden_dim;
    size_t emb_len = total > (size_t)sl * att_dim ? total : (size_t)sl * att_dim; /* :497-499 */
    float* embeddings = (float*)calloc(emb_len, 4);
    float* temp_embeddings = (float*)calloc(total, 4);
    float* hidden = (float*)calloc(total_h, 4);
    float* temp_hidden = (float*)calloc(total_h, 4);
    float* sq = (float*)calloc((size_t)att_dim * sl, 4);

    /* :409-411 attention rmsnorm per token */
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < sl; i++)
        lmrs_ref_rmsnorm(embeddings + (size_t)i * dim, x + (size_t)i * dim, m->rms_att + (size_t)l * dim, dim,
                         p.rms_norm_eps, gemma);

    size_t loff = (size_t)l * p.seq_len * kv_dim; /* :413 */
    float* k = m->key_cache + loff + (size_t)pos * kv_dim;
    float* v = m->value_cache + loff + (size_t)pos * kv_dim;

    if (!m->quantized) { /* :419-422 */
        lmrs_ref_matmul_f32(sq, embeddings, m->wq_f + (size_t)l * dim * att_dim, sl, dim, att_dim);
        lmrs_ref_matmul_f32(k, embeddings, m->wk_f + (size_t)l * dim * kv_dim, sl, dim, kv_dim);
        lmrs_ref_matmul_f32(v, embeddings, m->wv_f + (size_t)l * dim * kv_dim, sl, dim, kv_dim);
    } else if (p.q_type == 1) { /* :426-431: one quantize shared by the three matmuls */
        int gs = p.group_size;
        int8_t* q = (int8_t*)malloc(total);
        float* s = (float*)malloc(total / gs * 4);
        lmrs_ref_quantize_q8(q, s, embeddings, (int)total, gs);
        lmrs_ref_matmul_q8(sq, q, s, (const int8_t*)m->wq_q[l].q, m->wq_q[l].s, sl, dim, att_dim, gs);
        lmrs_ref_matmul_q8(k, q, s, (const int8_t*)m->wk_q[l].q, m->wk_q[l].s, sl, dim, kv_dim, gs);
        lmrs_ref_matmul_q8(v, q, s, (const int8_t*)m->wv_q[l].q, m->wv_q[l].s, sl, dim, kv_dim, gs);
        free(q); free(s);
    } else { /* :432-437 */
        int gs = p.group_size;
        uint8_t* q = (uint8_t*)malloc(total / 2);
        float* s = (float*)malloc(total / gs * 4);
        lmrs_ref_quantize_q4(q, s, embeddings, (int)total, gs);
        lmrs_ref_matmul_q4(sq, q, s, (const uint8_t*)m->wq_q[l].q, m->wq_q[l].s, sl, dim, att_dim, gs);
        lmrs_ref_matmul_q4(k, q, s, (const uint8_t*)m->wk_q[l].q, m->wk_q[l].s, sl, dim, kv_dim, gs);
        lmrs_ref_matmul_q4(v, q, s, (const uint8_t*)m->wv_q[l].q, m->wv_q[l].s, sl, dim, kv_dim, gs);
        free(q); free(s);
    }

    /* RoPE :443-495 -- rotate-half pairs (j, j+hs/2); q always, k iff the index is inside kv_dim */
#pragma omp parallel for schedule(static)
    for (uint32_t idx = 0; idx < sl; idx++) {
        float* tk = k + (size_t)idx * kv_dim;
        float* tq = sq + (size_t)idx * att_dim;
        for (uint32_t i = 0; i < p.n_heads; i++) {
            for (uint32_t j = 0; j < head_size / 2; j++) {
                float freq, scaling_factor;
                lmrs_ref_rope_freq(p.model_type, p.rope_theta, head_size, j, &freq, &scaling_factor);
                float val = (float)(pos + idx) * freq;
                float fcr = cosf(val) * scaling_factor;
                float fci = sinf(val) * scaling_factor;
                uint32_t rotn = (i * head_size + j + head_size / 2 < kv_dim) ? 2 : 1;
                for (uint32_t vv = 0; vv < rotn; vv++) {
                    float* vec = vv == 0 ? tq : tk;
                    float v0 = vec[i * head_size + j], v1 = vec[i * head_size + j + head_size / 2];
                    vec[i * head_size + j] = v0 * fcr - v1 * fci;
                    vec[i * head_size + j + head_size / 2] = v0 * fci + v1 * fcr;
                }
            }
        }
    }

    /* attention :501-544 -- serial f32 dot, /sqrt(hs), Gemma soft-cap + window, softmax, serial AV */
    int fail_flag = 0;
#pragma omp parallel
    {
        float* att = (float*)malloc((size_t)p.seq_len * 4);
#pragma omp for collapse(2) schedule(dynamic, 1)
        for (uint32_t i = 0; i < sl; i++) {
            for (uint32_t h = 0; h < p.n_heads; h++) {
                float* xb = embeddings + (size_t)i * att_dim + (size_t)h * head_size;
                const float* q = sq + (size_t)i * att_dim + (size_t)h * head_size;
                uint32_t nt = pos + i + 1;
                for (uint32_t t = 0; t < nt; t++) {
                    const float* kk = m->key_cache + loff + (size_t)t * kv_dim + (size_t)(h / kv_mul) * head_size;
                    float score = 0.0f;
                    for (uint32_t d = 0; d < head_size; d++) score += q[d] * kk[d];
                    score /= sqrtf((float)head_size);
                    if (gemma) {
                        score /= 50.0f;
                        score = (float)tanh((double)score);
                        score *= 50.0f;
                        uint32_t dist = pos - t; /* u32 wrap when t > pos (release build), :525 */
                        score += dist <= 4096u ? 0.0f : -2.3819763e38f;
                    }
                    att[t] = score;
                }
                lmrs_ref_softmax(att, (int)nt);
                for (uint32_t d = 0; d < head_size; d++) xb[d] = 0.0f;
                for (uint32_t t = 0; t < nt; t++) {
                    const float* vv = m->value_cache + loff + (size_t)t * kv_dim + (size_t)(h / kv_mul) * head_size;
                    float a = att[t];
                    for (uint32_t d = 0; d < head_size; d++) xb[d] += a * vv[d];
                }
            }
        }
        free(att);
    }
    (void)fail_flag;

    /* output projection :546-560 */
    if (!m->quantized)
        lmrs_ref_matmul_f32(temp_embeddings, embeddings, m->wo_f + (size_t)l * dim * att_dim, sl, att_dim, dim);
    else
        qmatmul_sh(m, temp_embeddings, embeddings, &m->wo_q[l], sl, att_dim, dim, g_kshards);

    /* residual + norm :562-580 (rows of `embeddings` are re-used as dim-strided scratch) */
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < sl; i++) {
        float* xe = x + (size_t)i * dim;
        float* emb = embeddings + (size_t)i * dim;
        const float* temb = temp_embeddings + (size_t)i * dim;
        if (gemma) {
            lmrs_ref_rmsnorm(emb, temb, m->rms_post_att + (size_t)l * dim, dim, p.rms_norm_eps, 1);
            for (uint32_t d = 0; d < dim; d++) xe[d] += emb[d];
            lmrs_ref_rmsnorm(emb, xe, m->rms_pre_ffn + (size_t)l * dim, dim, p.rms_norm_eps, 1);
        } else {
            for (uint32_t d = 0; d < dim; d++) xe[d] += temb[d];
            lmrs_ref_rmsnorm(emb, xe, m->rms_post_att + (size_t)l * dim, dim, p.rms_norm_eps, 0);
        }
    }

    /* gate / up :588-605 */
    if (!m->quantized) {
        lmrs_ref_matmul_f32(hidden, embeddings, m->w1_f + (size_t)l * dim * hidden_dim, sl, dim, hidden_dim);
        lmrs_ref_matmul_f32(temp_hidden, embeddings, m->w3_f + (size_t)l * dim * hidden_dim, sl, dim, hidden_dim);
    } else if (p.q_type == 1) {
        int gs = p.group_size;
        int8_t* q = (int8_t*)malloc(total);
        float* s = (float*)malloc(total / gs * 4);
        lmrs_ref_quantize_q8(q, s, embeddings, (int)total, gs);
        lmrs_ref_matmul_q8(hidden, q, s, (const int8_t*)m->w1_q[l].q, m->w1_q[l].s, sl, dim, hidden_dim, gs);
        lmrs_ref_matmul_q8(temp_hidden, q, s, (const int8_t*)m->w3_q[l].q, m->w3_q[l].s, sl, dim, hidden_dim, gs);
        free(q); free(s);
    } else {
        int gs = p.group_size;
        uint8_t* q = (uint8_t*)malloc(total / 2);
        float* s = (float*)malloc(total / gs * 4);
        lmrs_ref_quantize_q4(q, s, embeddings, (int)total, gs);
        lmrs_ref_matmul_q4(hidden, q, s, (const uint8_t*)m->w1_q[l].q, m->w1_q[l].s, sl, dim, hidden_dim, gs);
        lmrs_ref_matmul_q4(temp_hidden, q, s, (const uint8_t*)m->w3_q[l].q, m->w3_q[l].s, sl, dim, hidden_dim, gs);
        free(q); free(s);
    }

    /* activation :607-624 */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < total_h; i++) {
        float val = hidden[i];
        if (gemma) { /* tanh-GELU with the tanh evaluated in f64, :614 */
            float inner = val + 0.044715f * val * val * val;
            val *= 0.5f * (1.0f + (float)tanh(0.7978845608028654 * (double)inner));
        } else { /* SiLU :617 */
            val *= 1.0f / (1.0f + expf(-val));
        }
        val *= temp_hidden[i];
        hidden[i] = val;
    }

    /* down :626-640 -> rows of `embeddings` at stride dim */
    if (!m->quantized)
        lmrs_ref_matmul_f32(embeddings, hidden, m->w2_f + (size_t)l * dim * hidden_dim, sl, hidden_dim, dim);
    else
        qmatmul_sh(m, embeddings, hidden, &m->w2_q[l], sl, hidden_dim, dim, g_kshards);

    /* final residual :642-656 */
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < sl; i++) {
        float* xe = x + (size_t)i * dim;
        const float* emb = embeddings + (size_t)i * dim;
        float* temb = temp_embeddings + (size_t)i * dim;
        if (gemma) {
            lmrs_ref_rmsnorm(temb, emb, m->rms_post_ffn + (size_t)l * dim, dim, p.rms_norm_eps, 1);
            for (uint32_t d = 0; d < dim; d++) xe[d] += temb[d];
        } else {
            for (uint32_t d = 0; d < dim; d++) xe[d] += emb[d];
        }
    }

    free(embeddings); free(temp_embeddings); free(hidden); free(temp_hidden); free(sq);
    return 0;
}

/* src/transformer.rs:316-384 */
int lmrs_ref_forward(lmrs_ref_t* m, uint32_t token, uint32_t pos, float** logits) {
    const lmrs_ref_args_t p = m->args;
    uint32_t dim = p.dim;
    if (token >= p.vocab_size) return fail("token out of range");
    float* x = (float*)malloc((size_t)dim * 4);
    float* xb = (float*)malloc((size_t)dim * 4);
    embedding_row(m, token, x); /* :324 */
    if (p.model_type == 0) {    /* :327-332 */
        float normalizer = sqrtf((float)dim);
        for (uint32_t i = 0; i < dim; i++) x[i] *= normalizer;
    }
    for (uint32_t l = 0; l < p.n_layers; l++)
        if (forward_layer(m, x, 1, l, pos)) { free(x); free(xb); return 1; }
    memcpy(xb, x, (size_t)dim * 4);
    lmrs_ref_rmsnorm(x, xb, m->rms_final, dim, p.rms_norm_eps, p.model_type == 0); /* :343 */
    if (!m->quantized) { /* :346-351 */
        lmrs_ref_matmul_f32(m->logits, x, p.model_type != 2 ? m->emb_f : m->lm_head_f, 1, dim, p.vocab_size);
    } else { /* :353-371 */
        const qt_t* w = p.model_type != 2 ? &m->emb_q : &m->lm_head_q;
        qmatmul(m, m->logits, x, w, 1, dim, p.vocab_size);
    }
    if (p.model_type == 0) { /* :375-381 -- loops 0..dim (not vocab): reference quirk kept */
        for (uint32_t d = 0; d < dim && d < p.vocab_size; d++) {
            m->logits[d] /= 30.0f;
            m->logits[d] = (float)tanh((double)m->logits[d]);
            m->logits[d] *= 30.0f;
        }
    }
    free(x); free(xb);
    *logits = m->logits;
    return 0;
}

/* src/transformer.rs:659-669 */
int lmrs_ref_get_embeddings(const lmrs_ref_t* m, const uint32_t* tokens, size_t n, float* out) {
    for (size_t t = 0; t < n; t++) {
        if (tokens[t] >= m->args.vocab_size) return fail("token out of range");
        embedding_row(m, tokens[t], out + t * m->args.dim);
    }
    return 0;
}

/* src/transformer.rs:672-684 */
int lmrs_ref_fill_kv_cache(lmrs_ref_t* m, float* emb, size_t n_floats, uint32_t pos, uint32_t* new_pos) {
    uint32_t num = (uint32_t)(n_floats / m->args.dim);
    for (uint32_t l = 0; l < m->args.n_layers; l++)
        if (forward_layer(m, emb, num, l, pos)) return 1;
    *new_pos = pos + num;
    return 0;
}
