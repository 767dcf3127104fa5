"""Host-side mirror of `lmrs::transformer` (src/transformer.rs) over the C ABI.

Same names, argument meaning and error behaviour as the Rust pub surface the bins use:
    Transformer.new(data) -> (Transformer, end_offset)      src/transformer.rs:134
    Transformer.forward(token, pos) -> logits[vocab]        :316
    Transformer.get_embeddings(tokens) -> f32[n*dim]        :659
    Transformer.fill_kv_cache(embeddings, pos) -> new_pos   :672
    Transformer.args.{vocab_size, model_type, multimodal}   :66,71,73
Where the reference panics this raises LmrsError.  All arithmetic happens in liblmrs_b200.so on the GPU.
"""
import ctypes as C
from enum import IntEnum

import numpy as np

from ._lib import Args, LmrsError, check, lib


class ModelType(IntEnum):  # src/transformer.rs:50-55
    GEMMA = 0
    LLAMA = 1
    PHI = 2


class QuantType(IntEnum):  # src/quantization.rs:1-6
    NONE = 0
    Q8_0 = 1
    Q4_0 = 2


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


class Transformer:
    def __init__(self, handle, args, keepalive=None, world=1):
        self._h, self.args, self._keep, self._world = handle, args, keepalive, world

    @classmethod
    def new(cls, data, device: int = -1):
        """Transformer::new(&Mmap) -> (Transformer, usize).  `data`: bytes-like LMRS v4 image (np.uint8 array,
        mmap, bytes).  Weights are copied to HBM; `data` need not outlive the call."""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        h, end = C.c_void_p(), C.c_size_t()
        check(lib().lmrs_b200_create(_vp(buf), buf.size, device, C.byref(h), C.byref(end)))
        a = Args()
        check(lib().lmrs_b200_args(h, C.byref(a)))
        return cls(h, a), end.value

    @classmethod
    def new_multi(cls, data, n_gpus: int):
        """Transformer::new on n_gpus GPUs of this process (lmrs_b200_create_multi): one handle drives all of them."""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        h, end = C.c_void_p(), C.c_size_t()
        check(lib().lmrs_b200_create_multi(_vp(buf), buf.size, n_gpus, C.byref(h), C.byref(end)))
        a = Args()
        check(lib().lmrs_b200_args(h, C.byref(a)))
        return cls(h, a), end.value   # (read_kv of a group returns whole rows: world stays 1 for the mirror)

    @classmethod
    def new_sharded(cls, data, device: int, rank: int, world: int, nccl_unique_id: bytes):
        """One process per GPU, output rows sharded across `world` ranks."""
        buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, np.uint8)
        h, end = C.c_void_p(), C.c_size_t()
        idbuf = C.create_string_buffer(bytes(nccl_unique_id), 128) if world > 1 else None
        check(lib().lmrs_b200_create_sharded(_vp(buf), buf.size, device, rank, world, idbuf, C.byref(h), C.byref(end)))
        a = Args()
        check(lib().lmrs_b200_args(h, C.byref(a)))
        return cls(h, a, world=world), end.value

    # -- the four methods of the drop-in boundary -------------------------------------------------------------
    def forward(self, token: int, pos: int) -> np.ndarray:
        """&mut [f32] of vocab_size logits in library-owned pinned memory, valid until the next call."""
        out = C.POINTER(C.c_float)()
        check(lib().lmrs_b200_forward(self._h, token, pos, C.byref(out)))
        return np.ctypeslib.as_array(out, shape=(self.args.vocab_size,))

    def forward_argmax(self, token: int, pos: int) -> int:
        """forward + Sampler::sample at temperature 0 (sample_argmax, src/sampler.rs:29-41) fused on the device."""
        nxt = C.c_uint32()
        check(lib().lmrs_b200_forward_argmax(self._h, token, pos, C.byref(nxt)))
        return nxt.value

    def generate_greedy(self, first_token: int, pos: int, max_new: int, eos: int = -1) -> np.ndarray:
        """chat.rs:188-226 at temperature 0, token feedback on the device; returns the picked ids (eos included)."""
        out, n = np.zeros(max(max_new, 1), np.uint32), C.c_uint32()
        check(lib().lmrs_b200_generate_greedy(self._h, first_token, pos, max_new, eos, _vp(out), C.byref(n)))
        return out[: n.value].copy()

    def get_embeddings(self, tokens) -> np.ndarray:
        t = np.ascontiguousarray(tokens, dtype=np.uint32)
        out = np.zeros(t.size * self.args.dim, np.float32)
        check(lib().lmrs_b200_get_embeddings(self._h, _vp(t), t.size, _vp(out)))
        return out

    def fill_kv_cache(self, embeddings: np.ndarray, curr_pos: int) -> int:
        if embeddings.dtype != np.float32 or not embeddings.flags.c_contiguous:
            raise LmrsError("embeddings must be a contiguous float32 array (it is updated in place)")
        new_pos = C.c_uint32()
        check(lib().lmrs_b200_fill_kv_cache(self._h, _vp(embeddings), embeddings.size, curr_pos, C.byref(new_pos)))
        return new_pos.value

    # -- device-resident extras -------------------------------------------------------------------------------
    def forward_device(self, token: int, pos: int) -> None:
        check(lib().lmrs_b200_forward_device(self._h, token, pos))

    def logits_device_ptr(self) -> int:
        p = C.c_void_p()
        check(lib().lmrs_b200_logits_device(self._h, C.byref(p)))
        return p.value

    def set_stream(self, cuda_stream: int) -> None:
        check(lib().lmrs_b200_set_stream(self._h, C.c_void_p(cuda_stream)))

    def synchronize(self) -> None:
        check(lib().lmrs_b200_synchronize(self._h))

    def kernel_launches(self) -> int:
        n = C.c_uint64()
        check(lib().lmrs_b200_kernel_launches(self._h, C.byref(n)))
        return n.value

    def bench_gemv_pass(self, pos: int) -> int:
        """enqueue only the gemv_kernel launches of one decode step (measurement aid); returns the launch count"""
        n = C.c_int()
        check(lib().lmrs_b200_bench_gemv_pass(self._h, pos, C.byref(n)))
        return n.value

    def bench_attn_pass(self, pos: int) -> int:
        """enqueue only the attention launches of one decode step at `pos` (measurement aid); returns the launch count"""
        n = C.c_int()
        check(lib().lmrs_b200_bench_attn_pass(self._h, pos, C.byref(n)))
        return n.value

    def last_prefill_device_ms(self) -> float:
        ms = C.c_float()
        check(lib().lmrs_b200_last_prefill_device_ms(self._h, C.byref(ms)))
        return ms.value

    def read_kv(self, layer: int, pos0: int, n: int):
        kvd = self.args.head_size * self.args.n_kv_heads // self._world   # a sharded handle holds its own KV heads only
        k, v = np.zeros((n, kvd), np.float32), np.zeros((n, kvd), np.float32)
        check(lib().lmrs_b200_read_kv(self._h, layer, pos0, n, _vp(k), _vp(v)))
        return k, v

    def debug_buffer(self, name: str) -> np.ndarray:
        cap = max(self.args.hidden_dim, self.args.dim, self.args.n_heads * self.args.head_size,
                  (5 * self.args.n_layers + 2) * 4 * 160 * 2 if name == "timing" else 0, 2 * 8192 * 2 if name == "trace" else 0)
        out, n = np.zeros(cap, np.float32), C.c_size_t(cap)
        check(lib().lmrs_b200_debug_buffer(self._h, name.encode(), _vp(out), C.byref(n)))
        return out[: n.value].copy()

    def close(self):  # impl Drop
        if self._h:
            lib().lmrs_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib().lmrs_b200_nccl_unique_id(buf))
    return buf.raw
