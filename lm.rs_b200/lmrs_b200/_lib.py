"""Loader for liblmrs_b200.so (the C ABI of include/lmrs_b200.h).  No fallback: if the shared library is
missing the import error says how to build it; if no sm_100 GPU is present every compute call raises."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LMRS_B200_SO") or os.path.join(_HERE, "liblmrs_b200.so")   # LMRS_B200_SO: A/B builds
PKG_ROOT = os.path.dirname(_HERE)

# every symbol include/lmrs_b200.h declares (tests/test_abi.py checks the .so exports all of them)
ABI_SYMBOLS = [
    "lmrs_b200_create", "lmrs_b200_create_multi", "lmrs_b200_create_sharded", "lmrs_b200_nccl_unique_id", "lmrs_b200_destroy",
    "lmrs_b200_args", "lmrs_b200_forward", "lmrs_b200_get_embeddings", "lmrs_b200_fill_kv_cache",
    "lmrs_b200_forward_argmax", "lmrs_b200_generate_greedy", "lmrs_b200_forward_device", "lmrs_b200_logits_device", "lmrs_b200_set_stream", "lmrs_b200_synchronize",
    "lmrs_b200_kernel_launches", "lmrs_b200_bench_gemv_pass", "lmrs_b200_bench_attn_pass", "lmrs_b200_last_prefill_device_ms", "lmrs_b200_read_kv", "lmrs_b200_debug_buffer", "lmrs_b200_matmul_q8", "lmrs_b200_matmul_q4", "lmrs_b200_matmul_f32", "lmrs_b200_matmul_rest",
    "lmrs_b200_quantize_q8", "lmrs_b200_quantize_q4", "lmrs_b200_rmsnorm", "lmrs_b200_softmax", "lmrs_b200_layernorm",
    "lmrs_b200_weights_upload", "lmrs_b200_matmul_w", "lmrs_b200_weights_free",
    "lmrs_b200_last_error", "lmrs_b200_version",
]


class Args(C.Structure):
    """lmrs_args_t == TransformerArgs (src/transformer.rs:57-74)"""
    _pack_ = 1
    _fields_ = [(n, C.c_uint32) for n in ("dim", "hidden_dim", "n_layers", "n_heads", "head_size", "n_kv_heads",
                                         "vocab_size", "seq_len")] + \
               [("rms_norm_eps", C.c_float), ("rope_theta", C.c_float), ("q_type", C.c_uint8),
                ("model_type", C.c_uint8), ("group_size", C.c_uint32), ("multimodal", C.c_uint8)]


def build(force: bool = False) -> str:
    """Compile liblmrs_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    cmd = ["make", "-C", PKG_ROOT, "-s"] + (["-B"] if force else [])
    subprocess.check_call(cmd)
    return SO_PATH


_lib = None
_f32p = C.POINTER(C.c_float)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(f"{SO_PATH} not built: run `make -C {PKG_ROOT}` (or __graft_entry__.build()); "
                          "lmrs_b200 has no CPU fallback")
    L = C.CDLL(SO_PATH)
    vp, sz, u32, i = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int
    L.lmrs_b200_last_error.restype = C.c_char_p
    L.lmrs_b200_version.restype = C.c_char_p
    L.lmrs_b200_create.argtypes = [vp, sz, i, C.POINTER(vp), C.POINTER(sz)]
    L.lmrs_b200_create_multi.argtypes = [vp, sz, i, C.POINTER(vp), C.POINTER(sz)]
    L.lmrs_b200_create_sharded.argtypes = [vp, sz, i, i, i, vp, C.POINTER(vp), C.POINTER(sz)]
    L.lmrs_b200_nccl_unique_id.argtypes = [vp]
    L.lmrs_b200_destroy.argtypes = [vp]
    L.lmrs_b200_destroy.restype = None
    L.lmrs_b200_args.argtypes = [vp, C.POINTER(Args)]
    L.lmrs_b200_forward.argtypes = [vp, u32, u32, C.POINTER(_f32p)]
    L.lmrs_b200_forward_argmax.argtypes = [vp, u32, u32, C.POINTER(u32)]
    L.lmrs_b200_generate_greedy.argtypes = [vp, u32, u32, u32, C.c_int32, vp, C.POINTER(u32)]
    L.lmrs_b200_get_embeddings.argtypes = [vp, vp, sz, vp]
    L.lmrs_b200_fill_kv_cache.argtypes = [vp, vp, sz, u32, C.POINTER(u32)]
    L.lmrs_b200_forward_device.argtypes = [vp, u32, u32]
    L.lmrs_b200_logits_device.argtypes = [vp, C.POINTER(vp)]
    L.lmrs_b200_set_stream.argtypes = [vp, vp]
    L.lmrs_b200_synchronize.argtypes = [vp]
    L.lmrs_b200_kernel_launches.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.lmrs_b200_bench_gemv_pass.argtypes = [vp, u32, C.POINTER(i)]
    L.lmrs_b200_bench_attn_pass.argtypes = [vp, u32, C.POINTER(i)]
    L.lmrs_b200_last_prefill_device_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.lmrs_b200_read_kv.argtypes = [vp, u32, u32, u32, vp, vp]
    L.lmrs_b200_debug_buffer.argtypes = [vp, C.c_char_p, vp, C.POINTER(sz)]
    L.lmrs_b200_matmul_q8.argtypes = [vp] * 5 + [i] * 4
    L.lmrs_b200_matmul_q4.argtypes = [vp] * 5 + [i] * 4
    L.lmrs_b200_matmul_f32.argtypes = [vp] * 3 + [i] * 3
    L.lmrs_b200_matmul_rest.argtypes = [vp] * 3 + [i] * 3
    L.lmrs_b200_quantize_q8.argtypes = [vp] * 3 + [i] * 2
    L.lmrs_b200_quantize_q4.argtypes = [vp] * 3 + [i] * 2
    L.lmrs_b200_rmsnorm.argtypes = [vp] * 3 + [i, C.c_float, i]
    L.lmrs_b200_softmax.argtypes = [vp, i]
    L.lmrs_b200_layernorm.argtypes = [vp] * 4 + [i, i, C.c_float]
    L.lmrs_b200_weights_upload.argtypes = [i, vp, vp, i, i, i, C.POINTER(vp)]
    L.lmrs_b200_matmul_w.argtypes = [vp, vp, vp, vp, i]
    L.lmrs_b200_weights_free.argtypes = [vp]
    L.lmrs_b200_weights_free.restype = None
    _lib = L
    return L


class LmrsError(RuntimeError):
    """Raised where the reference would panic (assert!/expect/slice OOB)."""


def check(rc):
    if rc != 0:
        raise LmrsError(lib().lmrs_b200_last_error().decode(errors="replace"))
