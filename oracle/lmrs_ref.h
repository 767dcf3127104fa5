/*
 * oracle/lmrs_ref.h -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * C restatement of the CPU arithmetic of samuel-vitorino/lm.rs for the hot path
 * (src/functional.rs, src/quantization.rs, src/transformer.rs).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (liblmrs_b200.so) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned".  The reference ships no tests, golden vectors
 * or fixtures for this path and its Rust crate cannot be built
 * in this image (no rustc/cargo; deps `wide`, `rayon` un-vendored).  The oracle is
 * pinned only by known-answer vectors hand-derived from the cited source lines
 * (tests/test_oracle_kat.py) and by the reference's *Python* exporter for the
 * file format / weight quantiser (tests/golden/).
 */
#ifndef LMRS_REF_H
#define LMRS_REF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 47-byte packed header, file bytes 8..55 (src/transformer.rs:57-74, export.py:54-80). */
#pragma pack(push, 1)
typedef struct lmrs_ref_args {
    uint32_t dim, hidden_dim, n_layers, n_heads, head_size, n_kv_heads, vocab_size, seq_len;
    float    rms_norm_eps, rope_theta;
    uint8_t  q_type;      /* 0 none, 1 Q8_0, 2 Q4_0   (src/quantization.rs:1-6) */
    uint8_t  model_type;  /* 0 GEMMA, 1 LLAMA, 2 PHI  (src/transformer.rs:50-55) */
    uint32_t group_size;
    uint8_t  multimodal;
} lmrs_ref_args_t;
#pragma pack(pop)

typedef struct lmrs_ref lmrs_ref_t;

/* Transformer::new (src/transformer.rs:134-314).  `file` must outlive the handle (weights are
 * borrowed, like the reference's Mmap).  Returns 0 on success. */
int  lmrs_ref_create(const uint8_t* file, size_t len, lmrs_ref_t** out, size_t* end_offset);
void lmrs_ref_destroy(lmrs_ref_t* m);
int  lmrs_ref_args(const lmrs_ref_t* m, lmrs_ref_args_t* out);
const char* lmrs_ref_last_error(void);

/* Transformer::forward (src/transformer.rs:316-384): returns library-owned logits[vocab]. */
int lmrs_ref_forward(lmrs_ref_t* m, uint32_t token, uint32_t pos, float** logits);
/* Transformer::get_embeddings (src/transformer.rs:659-669). */
int lmrs_ref_get_embeddings(const lmrs_ref_t* m, const uint32_t* tokens, size_t n, float* out);
/* Transformer::fill_kv_cache (src/transformer.rs:672-684): in-place on emb, returns new pos. */
int lmrs_ref_fill_kv_cache(lmrs_ref_t* m, float* emb_inout, size_t n_floats, uint32_t pos, uint32_t* new_pos);
/* test access to the f32 KV cache [layer][seq_len][kv_dim] (src/transformer.rs:302-303) */
const float* lmrs_ref_key_cache(const lmrs_ref_t* m);
const float* lmrs_ref_value_cache(const lmrs_ref_t* m);

/* src/functional.rs */
void lmrs_ref_rmsnorm(float* o, const float* x, const float* w, int size, float eps, int add_unit_offset);
void lmrs_ref_layernorm(float* o, const float* x, const float* w, const float* b, int size, float eps);
void lmrs_ref_softmax(float* x, int n);
void lmrs_ref_matmul_f32(float* xout, const float* x, const float* w, int rows, int n, int o);
void lmrs_ref_matmul_rest(float* xout, const float* x, const float* w, int rows, int n, int o);
void lmrs_ref_matmul_q8(float* xout, const int8_t* xq, const float* xs, const int8_t* wq, const float* ws,
                        int rows, int n, int o, int gs);
void lmrs_ref_matmul_q4(float* xout, const uint8_t* xq, const float* xs, const uint8_t* wq, const float* ws,
                        int rows, int n, int o, int gs);
/* src/quantization.rs */
void lmrs_ref_quantize_q8(int8_t* q, float* s, const float* x, int n, int gs);
void lmrs_ref_quantize_q4(uint8_t* q, float* s, const float* x, int n, int gs);
void lmrs_ref_dequantize(float* x, const void* q, const float* s, int n, int gs, int q_type);
/* RoPE frequency for pair index j (src/transformer.rs:445-478); writes freq and magnitude scale. */
void lmrs_ref_rope_freq(int model_type, float rope_theta, int head_size, int j, float* freq, float* mscale);

/* NOT a reference feature: Wo / W2 accumulate `n` contiguous K ranges separately and add the partials in ascending
 * order -- the summation order of lmrs_b200's N-GPU row-sharded mode, so that N-GPU runs can be checked bit for bit. */
void lmrs_ref_set_kshards(int n);
void lmrs_ref_matmul_q8_kshards(float* xout, const int8_t* xq, const float* xs, const int8_t* wq, const float* ws,
                                int rows, int n, int o, int gs, int shards);

int lmrs_ref_num_threads(void);
void lmrs_ref_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
