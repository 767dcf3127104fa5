"""ctypes binding of the CPU oracle (oracle/_build/liblmrs_ref.so) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (lm.rs_b200/lmrs_b200) never does.  PARITY UNPINNED, see oracle/lmrs_ref.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblmrs_ref.so")


def usable_cores():
    """CPU cores this process may actually use: affinity mask, capped by the cgroup v2/v1 CPU quota."""
    n = len(os.sched_getaffinity(0))
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(per))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return n


def set_threads(n):
    lib().lmrs_ref_set_num_threads(int(n))


def set_kshards(n):
    """NOT a reference feature: Wo / W2 accumulate n contiguous K ranges separately and add the partials in ascending
    order -- the summation order of lmrs_b200's N-GPU row-sharded mode (checked bit for bit by the multi-GPU tests)."""
    lib().lmrs_ref_set_kshards(int(n))


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "lmrs_ref.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class Args(C.Structure):
    _pack_ = 1
    _fields_ = [(n, C.c_uint32) for n in ("dim", "hidden_dim", "n_layers", "n_heads", "head_size", "n_kv_heads",
                                         "vocab_size", "seq_len")] + \
               [("rms_norm_eps", C.c_float), ("rope_theta", C.c_float), ("q_type", C.c_uint8),
                ("model_type", C.c_uint8), ("group_size", C.c_uint32), ("multimodal", C.c_uint8)]


_lib = None
_f32p = C.POINTER(C.c_float)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.lmrs_ref_set_num_threads.argtypes = [C.c_int]
        L.lmrs_ref_set_num_threads.restype = None
        if "OMP_NUM_THREADS" not in os.environ:
            # shared hosts: nproc may far exceed the cgroup's CPU quota, and oversubscribed OpenMP spin-waits are
            # catastrophic (measured 11 s/token at 128 threads); default to the quota, capped at 32
            L.lmrs_ref_set_num_threads(min(32, usable_cores()))
        L.lmrs_ref_matmul_q8_kshards.argtypes = [C.c_void_p] * 5 + [C.c_int] * 5
        L.lmrs_ref_matmul_q8_kshards.restype = None
        L.lmrs_ref_set_kshards.argtypes = [C.c_int]
        L.lmrs_ref_set_kshards.restype = None
        L.lmrs_ref_last_error.restype = C.c_char_p
        L.lmrs_ref_create.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.lmrs_ref_destroy.argtypes = [C.c_void_p]
        L.lmrs_ref_destroy.restype = None
        L.lmrs_ref_args.argtypes = [C.c_void_p, C.POINTER(Args)]
        L.lmrs_ref_forward.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(_f32p)]
        L.lmrs_ref_get_embeddings.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.lmrs_ref_fill_kv_cache.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.POINTER(C.c_uint32)]
        L.lmrs_ref_key_cache.argtypes = [C.c_void_p]
        L.lmrs_ref_key_cache.restype = _f32p
        L.lmrs_ref_value_cache.argtypes = [C.c_void_p]
        L.lmrs_ref_value_cache.restype = _f32p
        L.lmrs_ref_rmsnorm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int]
        L.lmrs_ref_rmsnorm.restype = None
        L.lmrs_ref_layernorm.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_float]
        L.lmrs_ref_layernorm.restype = None
        L.lmrs_ref_softmax.argtypes = [C.c_void_p, C.c_int]
        L.lmrs_ref_softmax.restype = None
        for name in ("lmrs_ref_matmul_f32", "lmrs_ref_matmul_rest"):
            getattr(L, name).argtypes = [C.c_void_p] * 3 + [C.c_int] * 3
            getattr(L, name).restype = None
        for name in ("lmrs_ref_matmul_q8", "lmrs_ref_matmul_q4"):
            getattr(L, name).argtypes = [C.c_void_p] * 5 + [C.c_int] * 4
            getattr(L, name).restype = None
        for name in ("lmrs_ref_quantize_q8", "lmrs_ref_quantize_q4"):
            getattr(L, name).argtypes = [C.c_void_p] * 3 + [C.c_int] * 2
            getattr(L, name).restype = None
        L.lmrs_ref_dequantize.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3
        L.lmrs_ref_dequantize.restype = None
        L.lmrs_ref_rope_freq.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, _f32p, _f32p]
        L.lmrs_ref_rope_freq.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a


# ---- operator level (src/functional.rs, src/quantization.rs) ---------------------------------------------

def rmsnorm(x, w, eps, add_unit_offset=False):
    x, w = _c(x, np.float32), _c(w, np.float32)
    o = np.zeros_like(x)
    lib().lmrs_ref_rmsnorm(_p(o), _p(x), _p(w), x.size, eps, int(add_unit_offset))
    return o


def layernorm(x, w, b, eps):
    x, w, b = _c(x, np.float32), _c(w, np.float32), _c(b, np.float32)
    o = np.zeros_like(x)
    lib().lmrs_ref_layernorm(_p(o), _p(x), _p(w), _p(b), x.size, eps)
    return o


def softmax(x):
    x = _c(x, np.float32).copy()
    lib().lmrs_ref_softmax(_p(x), x.size)
    return x


def quantize_q8(x, gs):
    x = _c(x, np.float32)
    q, s = np.zeros(x.size, np.int8), np.zeros(x.size // gs, np.float32)
    lib().lmrs_ref_quantize_q8(_p(q), _p(s), _p(x), x.size, gs)
    return q, s


def quantize_q4(x, gs):
    x = _c(x, np.float32)
    q, s = np.zeros(x.size // 2, np.uint8), np.zeros(x.size // gs, np.float32)
    lib().lmrs_ref_quantize_q4(_p(q), _p(s), _p(x), x.size, gs)
    return q, s


def dequantize(q, s, n, gs, q_type):
    out = np.zeros(n, np.float32)
    q = np.ascontiguousarray(q)
    lib().lmrs_ref_dequantize(_p(out), _p(q), _p(_c(s, np.float32)), n, gs, q_type)
    return out


def matmul_q8(xq, xs, wq, ws, rows, n, o, gs):
    xq, xs, wq, ws = _c(xq, np.int8), _c(xs, np.float32), _c(wq, np.int8), _c(ws, np.float32)
    out = np.zeros(rows * o, np.float32)
    lib().lmrs_ref_matmul_q8(_p(out), _p(xq), _p(xs), _p(wq), _p(ws), rows, n, o, gs)
    return out


def matmul_q8_kshards(xq, xs, wq, ws, rows, n, o, gs, shards):
    """matmul_q8 with the K groups accumulated in `shards` contiguous ranges whose partials are added in ascending order
    (the N-GPU order of lmrs_b200; not a reference feature)."""
    xq, xs, wq, ws = _c(xq, np.int8), _c(xs, np.float32), _c(wq, np.int8), _c(ws, np.float32)
    out = np.zeros(rows * o, np.float32)
    lib().lmrs_ref_matmul_q8_kshards(_p(out), _p(xq), _p(xs), _p(wq), _p(ws), rows, n, o, gs, shards)
    return out


def matmul_q4(xq, xs, wq, ws, rows, n, o, gs):
    xq, xs, wq, ws = _c(xq, np.uint8), _c(xs, np.float32), _c(wq, np.uint8), _c(ws, np.float32)
    out = np.zeros(rows * o, np.float32)
    lib().lmrs_ref_matmul_q4(_p(out), _p(xq), _p(xs), _p(wq), _p(ws), rows, n, o, gs)
    return out


def matmul_f32(x, w, rows, n, o, rest=False):
    x, w = _c(x, np.float32), _c(w, np.float32)
    out = np.zeros(rows * o, np.float32)
    (lib().lmrs_ref_matmul_rest if rest else lib().lmrs_ref_matmul_f32)(_p(out), _p(x), _p(w), rows, n, o)
    return out


def rope_freq(model_type, theta, head_size, j):
    f, m = C.c_float(), C.c_float()
    lib().lmrs_ref_rope_freq(model_type, theta, head_size, j, C.byref(f), C.byref(m))
    return np.float32(f.value), np.float32(m.value)


# ---- model level (src/transformer.rs) ---------------------------------------------------------------------

class RefTransformer:
    """Mirror of lmrs::transformer::Transformer over the oracle (new/forward/get_embeddings/fill_kv_cache)."""

    def __init__(self, data: np.ndarray):
        self._data = np.ascontiguousarray(data, dtype=np.uint8)  # borrowed for the handle's lifetime
        h, end = C.c_void_p(), C.c_size_t()
        if lib().lmrs_ref_create(_p(self._data), self._data.size, C.byref(h), C.byref(end)):
            raise RuntimeError(lib().lmrs_ref_last_error().decode())
        self._h, self.end_offset = h, end.value
        self.args = Args()
        lib().lmrs_ref_args(self._h, C.byref(self.args))

    def forward(self, token, pos):
        out = _f32p()
        if lib().lmrs_ref_forward(self._h, token, pos, C.byref(out)):
            raise RuntimeError(lib().lmrs_ref_last_error().decode())
        return np.ctypeslib.as_array(out, shape=(self.args.vocab_size,))

    def get_embeddings(self, tokens):
        t = _c(tokens, np.uint32)
        out = np.zeros(t.size * self.args.dim, np.float32)
        if lib().lmrs_ref_get_embeddings(self._h, _p(t), t.size, _p(out)):
            raise RuntimeError(lib().lmrs_ref_last_error().decode())
        return out

    def fill_kv_cache(self, emb, pos):
        assert emb.dtype == np.float32 and emb.flags.c_contiguous
        newpos = C.c_uint32()
        if lib().lmrs_ref_fill_kv_cache(self._h, _p(emb), emb.size, pos, C.byref(newpos)):
            raise RuntimeError(lib().lmrs_ref_last_error().decode())
        return newpos.value

    def kv_cache(self):
        a = self.args
        n = a.n_layers * a.seq_len * a.head_size * a.n_kv_heads
        shape = (a.n_layers, a.seq_len, a.head_size * a.n_kv_heads)
        k = np.ctypeslib.as_array(lib().lmrs_ref_key_cache(self._h), shape=(n,)).reshape(shape)
        v = np.ctypeslib.as_array(lib().lmrs_ref_value_cache(self._h), shape=(n,)).reshape(shape)
        return k, v

    def close(self):
        if self._h:
            lib().lmrs_ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
