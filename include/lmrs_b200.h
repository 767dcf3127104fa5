/*
 * lmrs_b200.h -- C ABI of liblmrs_b200.so: the Blackwell (sm_100a) implementation of the lm.rs
 * quantized transformer forward path.
 *
 * The reference (samuel-vitorino/lm.rs) has no FFI/plugin interface; its drop-in boundary is the pub
 * API of `lmrs::transformer` that the chat/backend bins call.  Every entry point
 * below names the reference item it replaces (paths relative to the reference repo).  A Rust shim with
 * the identical pub surface forwards to these symbols (lm.rs_b200/rust/).
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function returns 0 on success and a
 * non-zero status on failure; lmrs_b200_last_error() then returns a thread-local message.  (The
 * reference panics on the same conditions; the Rust shim turns a non-zero status into panic!.)
 * There is NO CPU fallback: every call fails loudly if no sm_100 device is usable.
 * A handle is not re-entrant (the reference takes &mut self); distinct handles are independent and may
 * be driven from different threads (src/bin/backend.rs:87-110 creates one Transformer per connection).
 */
#ifndef LMRS_B200_H
#define LMRS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* TransformerArgs, the 47-byte #[repr(C, packed)] header at file bytes 8..55
 * (src/transformer.rs:57-74; written by export.py:54-80). */
#pragma pack(push, 1)
typedef struct lmrs_args {
    uint32_t dim, hidden_dim, n_layers, n_heads, head_size, n_kv_heads, vocab_size, seq_len;
    float    rms_norm_eps, rope_theta;
    uint8_t  q_type;      /* QuantType: 0 None, 1 Q8_0, 2 Q4_0  (src/quantization.rs:1-6) */
    uint8_t  model_type;  /* ModelType: 0 GEMMA, 1 LLAMA, 2 PHI (src/transformer.rs:50-55) */
    uint32_t group_size;
    uint8_t  multimodal;
} lmrs_args_t;
#pragma pack(pop)

typedef struct lmrs_b200 lmrs_b200_t;

/* ---- Transformer::new(&Mmap) -> (Transformer, usize)            src/transformer.rs:134-314 ------------
 * Parses the LMRS v4 image [file, file+len), copies every tensor to HBM (no pointer into `file` is kept),
 * allocates the f32 KV cache [n_layers][min(seq_len,8192)][kv_dim] x2, the RoPE tables and the pinned
 * logits buffer.  *end_offset = byte offset where the vision section starts (the tuple's usize).
 * device = CUDA ordinal (-1: current device). */
int lmrs_b200_create(const uint8_t* file, size_t len, int device, lmrs_b200_t** out, size_t* end_offset);

/* Row-sharded variant for N GPUs, one process per GPU (BASELINE north_star): rank r
 * of world owns head-aligned output-row shards of wq/wk/wv/w1/w3 and input-column shards of wo/w2, and
 * all-reduces the residual contribution twice per layer.  nccl_unique_id: the 128 bytes produced by
 * lmrs_b200_nccl_unique_id() on rank 0 and broadcast by the caller (torch.distributed/MPI/...). */
int lmrs_b200_create_sharded(const uint8_t* file, size_t len, int device, int rank, int world,
                             const void* nccl_unique_id, lmrs_b200_t** out, size_t* end_offset);
int lmrs_b200_nccl_unique_id(void* out128);

/* impl Drop for Transformer                                      src/transformer.rs:688-711 */
void lmrs_b200_destroy(lmrs_b200_t* m);

/* Transformer::new on `n_gpus` GPUs of THIS process (devices 0 .. n_gpus-1), the form SURVEY.md section 8b proposes: one
 * handle per GPU behind the returned one, weights row-sharded like create_sharded, the GPUs mapped into one another with
 * cudaDeviceEnablePeerAccess and exchanging their partial results from inside the kernels (no NCCL, no second process, no
 * host synchronisation between the GPUs).  Every entry point below accepts the returned handle and drives all GPUs from
 * the caller's thread; logits / token ids / the residual stream come back from GPU 0.  n_gpus <= 1: same as lmrs_b200_create. */
int lmrs_b200_create_multi(const uint8_t* file, size_t len, int n_gpus, lmrs_b200_t** out, size_t* end_offset);

/* pub args: TransformerArgs (vocab_size, model_type, multimodal are the pub fields the bins read:
 * src/bin/chat.rs:85,135,158); seq_len is returned clamped to 8192 as at src/transformer.rs:158-160. */
int lmrs_b200_args(const lmrs_b200_t* m, lmrs_args_t* out);

/* ---- Transformer::forward(&mut self, token, pos) -> &mut [f32]   src/transformer.rs:316-384 -----------
 * One decode step.  *logits_host points at library-owned PINNED host memory of vocab_size floats, fully
 * rewritten by every call and valid until the next call; the caller may write into it (the reference's
 * sampler does, src/sampler.rs:115-117). */
int lmrs_b200_forward(lmrs_b200_t* m, uint32_t token, uint32_t pos, float** logits_host);

/* ---- Sampler::sample at temperature 0 == Sampler::sample_argmax      src/sampler.rs:29-41,112-113 ------------
 * forward + the greedy pick fused on the device: 4 bytes come back instead of vocab_size floats.  The scan order of
 * the reference (strict `>`: the FIRST maximum wins, NaN never wins, NaN at index 0 answers 0) is reproduced exactly.
 * The logits of the step stay readable through lmrs_b200_logits_device(). */
int lmrs_b200_forward_argmax(lmrs_b200_t* m, uint32_t token, uint32_t pos, uint32_t* next_token);
/* The generate loop of src/bin/chat.rs:188-226 at temperature 0: feeds first_token at pos, then every picked token at
 * the next position, up to max_new tokens or until `eos` (< 0: none) is produced; the picked token is handed to the next
 * step on the device (no host round trip per token).  out_tokens[0..*n_out) = the picked ids (eos included). */
int lmrs_b200_generate_greedy(lmrs_b200_t* m, uint32_t first_token, uint32_t pos, uint32_t max_new, int32_t eos,
                              uint32_t* out_tokens, uint32_t* n_out);

/* ---- Transformer::get_embeddings(&self, &[u32]) -> Vec<f32>      src/transformer.rs:659-669 ----------- */
int lmrs_b200_get_embeddings(const lmrs_b200_t* m, const uint32_t* tokens, size_t n_tokens, float* out);

/* ---- Transformer::fill_kv_cache(&mut self, &mut [f32], pos) -> u32   src/transformer.rs:672-684 -------
 * Batched prefill of n_floats/dim embeddings starting at pos: no final norm, no logits; the residual
 * stream is written back into emb_inout (in-place semantics of the reference); *new_pos = pos + n. */
int lmrs_b200_fill_kv_cache(lmrs_b200_t* m, float* emb_inout, size_t n_floats, uint32_t pos, uint32_t* new_pos);

/* ---- device-resident stepping (no reference counterpart; used by bench.py's HBM-resident `value` leg
 * and by callers that sample on the device).  forward_device enqueues one decode step on the handle's
 * stream and returns without synchronising or copying logits; logits stay in HBM. */
int lmrs_b200_forward_device(lmrs_b200_t* m, uint32_t token, uint32_t pos);
int lmrs_b200_logits_device(lmrs_b200_t* m, float** logits_dev);
int lmrs_b200_set_stream(lmrs_b200_t* m, void* cuda_stream); /* cudaStream_t; NULL = library-owned stream */
int lmrs_b200_synchronize(lmrs_b200_t* m);
/* number of CUDA kernels this handle has launched (graph replays count their kernel nodes) */
int lmrs_b200_kernel_launches(const lmrs_b200_t* m, uint64_t* count);
/* measurement aid: enqueue only the matrix-vector launches (gemv_kernel) of one decode step -- the 4 per block + the
 * classifier, with their real prologues/epilogues, attention skipped -- so that bench.py can time the HBM-bound kernel
 * by itself with CUDA events on the handle's stream.  *n_launches = launches enqueued.  Results are meaningless. */
int lmrs_b200_bench_gemv_pass(lmrs_b200_t* m, uint32_t pos, int* n_launches);
/* same for the attention launches (one per block) of a decode step at position `pos`; rewrites K row `pos` of every
 * layer from whatever the staging row holds, so only call it on a handle used for measurement. */
int lmrs_b200_bench_attn_pass(lmrs_b200_t* m, uint32_t pos, int* n_launches);
/* measurement aid: device time (CUDA events on the handle's stream) of the kernels of the last fill_kv_cache call, host
 * copies excluded; -1 before the first call */
int lmrs_b200_last_prefill_device_ms(const lmrs_b200_t* m, float* ms);
/* test access: copy K and V rows [pos0, pos0+n) of one layer to host (f32 [n][kv_dim] each) */
int lmrs_b200_read_kv(lmrs_b200_t* m, uint32_t layer, uint32_t pos0, uint32_t n, float* k_out, float* v_out);
/* test access: copy one activation buffer of the LAST executed block to host.  name: "x0","x1" (residual
 * ping-pong, dim), "q" (un-rotated, att_dim), "k_new" (kv_dim), "att" (att_dim), "wo_out" (dim), "h"
 * (hidden_dim), "down_out" (dim).  Returns the element count through *n (capacity in, count out). */
int lmrs_b200_debug_buffer(lmrs_b200_t* m, const char* name, float* out, size_t* n);

/* ---- operator level: the free functions vision.rs / processor.rs import (src/vision.rs:1-3,
 * src/processor.rs:1-3).  Host pointers in and out; each call stages through HBM and runs the same
 * kernels the model path uses.
 *   matmul_q8   src/functional.rs:173-214    xout[rows*o];  x {q i8[rows*n], s f32[rows*n/gs]};  w {q i8[o*n], s}
 *   matmul_q4   src/functional.rs:216-250    4-bit x and w, two values per byte, low nibble = even index
 *   quantize    src/quantization.rs:44-67    quantize_q4  src/quantization.rs:69-95
 *   rmsnorm     src/functional.rs:48-78      softmax      src/functional.rs:122-140
 *   matmul (f32) src/functional.rs:142-171   xout[rows*o] = x[rows*n] . w[o*n]^T, 8-lane chunks reduced then added
 *                                            serially; the n % 8 tail is dropped like the reference does
 *   matmul_rest src/functional.rs:252-280    any n; its tail reads x[r] of row 0 (reference quirk, kept) */
int lmrs_b200_matmul_q8(float* xout, const int8_t* xq, const float* xs, const int8_t* wq, const float* ws,
                        int rows, int n, int o, int gs);
int lmrs_b200_matmul_q4(float* xout, const uint8_t* xq, const float* xs, const uint8_t* wq, const float* ws,
                        int rows, int n, int o, int gs);
int lmrs_b200_matmul_f32(float* xout, const float* x, const float* w, int rows, int n, int o);
int lmrs_b200_matmul_rest(float* xout, const float* x, const float* w, int rows, int n, int o);
int lmrs_b200_quantize_q8(int8_t* q, float* s, const float* x, int n, int gs);
int lmrs_b200_quantize_q4(uint8_t* q, float* s, const float* x, int n, int gs);
int lmrs_b200_rmsnorm(float* o, const float* x, const float* w, int size, float eps, int add_unit_offset);
int lmrs_b200_softmax(float* x, int n);
/* layernorm   src/functional.rs:80-114     `rows` independent rows of `size` elements (src/vision.rs normalises every token
 *                                          row); the size % 8 tail of o is left as the caller passed it, like the reference */
int lmrs_b200_layernorm(float* o, const float* x, const float* w, const float* b, int rows, int size, float eps);

/* ---- quantized weights resident in HBM: QuantizedTensor views that src/vision.rs / src/processor.rs build once with
 * init_param_quant (src/transformer.rs:24-48) and then apply to every token row.  upload copies the matrix (q: i8[o*n]
 * for Q8_0, u8[o*n/2] for Q4_0; s: f32[o*n/gs]) to the current device once; matmul_w is matmul_q8 / matmul_q4
 * (src/functional.rs:173-250) of `rows` quantized activation rows against it -- only the rows and the result cross PCIe.
 * rows >= 8 of a Q8_0 matrix with 128-aligned shapes run on the tcgen05 GEMM, everything else on the matrix-vector kernel. */
typedef struct lmrs_b200_weights lmrs_b200_weights_t;
int  lmrs_b200_weights_upload(int q_type, const void* wq, const float* ws, int n, int o, int gs, lmrs_b200_weights_t** out);
int  lmrs_b200_matmul_w(float* xout, const void* xq, const float* xs, const lmrs_b200_weights_t* w, int rows);
void lmrs_b200_weights_free(lmrs_b200_weights_t* w);

const char* lmrs_b200_last_error(void);
/* "lmrs_b200 <version> sm_100a" */
const char* lmrs_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LMRS_B200_H */
