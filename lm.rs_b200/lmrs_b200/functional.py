"""Mirror of `lmrs::functional` (src/functional.rs) over the operator-level C ABI: same names and argument
order as the Rust free functions (xout first), numpy arrays instead of slices.  Runs on the GPU only."""
import ctypes as C

import numpy as np

from ._lib import check, lib


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def matmul_q8(xout: np.ndarray, x, w, n: int, o: int, gs: int) -> None:
    """src/functional.rs:173-214.  x, w: objects with .q (int8) and .s (float32) like (Mutable)QuantizedTensor."""
    rows = xout.size // o
    xq, xs, wq, ws = _c(x.q, np.int8), _c(x.s, np.float32), _c(w.q, np.int8), _c(w.s, np.float32)
    assert xout.dtype == np.float32 and xout.flags.c_contiguous
    check(lib().lmrs_b200_matmul_q8(_vp(xout), _vp(xq), _vp(xs), _vp(wq), _vp(ws), rows, n, o, gs))


def matmul_q4(xout: np.ndarray, x, w, n: int, o: int, gs: int) -> None:
    """src/functional.rs:216-250 (rows > 0 use row-wise indexing; the reference's own is broken there, :224)."""
    rows = xout.size // o
    xq, xs, wq, ws = _c(x.q, np.uint8), _c(x.s, np.float32), _c(w.q, np.uint8), _c(w.s, np.float32)
    assert xout.dtype == np.float32 and xout.flags.c_contiguous
    check(lib().lmrs_b200_matmul_q4(_vp(xout), _vp(xq), _vp(xs), _vp(wq), _vp(ws), rows, n, o, gs))


def matmul(xout: np.ndarray, x, w, n: int, o: int) -> None:
    """src/functional.rs:142-171: unquantized operands, xout[rows*o] = x[rows*n] . w[o*n]^T (the n % 8 tail is dropped)."""
    rows = xout.size // o
    x, w = _c(x, np.float32), _c(w, np.float32)
    assert xout.dtype == np.float32 and xout.flags.c_contiguous
    check(lib().lmrs_b200_matmul_f32(_vp(xout), _vp(x), _vp(w), rows, n, o))


def matmul_rest(xout: np.ndarray, x, w, n: int, o: int) -> None:
    """src/functional.rs:252-280: any n; the tail elements read row 0 of x (reference quirk, kept)."""
    rows = xout.size // o
    x, w = _c(x, np.float32), _c(w, np.float32)
    assert xout.dtype == np.float32 and xout.flags.c_contiguous
    check(lib().lmrs_b200_matmul_rest(_vp(xout), _vp(x), _vp(w), rows, n, o))


def rmsnorm(o: np.ndarray, x, weight, size: int, eps: float, add_unit_offset: bool) -> None:
    """src/functional.rs:48-78"""
    x, weight = _c(x, np.float32), _c(weight, np.float32)
    assert o.dtype == np.float32 and o.flags.c_contiguous
    check(lib().lmrs_b200_rmsnorm(_vp(o), _vp(x), _vp(weight), size, eps, int(add_unit_offset)))


def softmax(x: np.ndarray) -> None:
    """src/functional.rs:122-140 (in place)"""
    assert x.dtype == np.float32 and x.flags.c_contiguous
    check(lib().lmrs_b200_softmax(_vp(x), x.size))


def layernorm(o: np.ndarray, x, weight, bias, size: int, eps: float) -> None:
    """src/functional.rs:80-114; o / x may hold several rows of `size` elements (one call per token row in src/vision.rs)."""
    x, weight, bias = _c(x, np.float32), _c(weight, np.float32), _c(bias, np.float32)
    assert o.dtype == np.float32 and o.flags.c_contiguous and o.size == x.size and x.size % size == 0
    check(lib().lmrs_b200_layernorm(_vp(o), _vp(x), _vp(weight), _vp(bias), x.size // size, size, eps))


class ResidentWeights:
    """A QuantizedTensor kept in HBM (include/lmrs_b200.h lmrs_b200_weights_*): what src/vision.rs / src/processor.rs get
    from init_param_quant once and then pass to matmul_q8 / matmul_q4 for every token row."""

    def __init__(self, w, n: int, o: int, gs: int, q_type: int = 1):
        self.n, self.o, self.gs, self.q_type = n, o, gs, q_type
        wq = _c(w.q, np.int8 if q_type == 1 else np.uint8)
        ws = _c(w.s, np.float32)
        h = C.c_void_p()
        check(lib().lmrs_b200_weights_upload(q_type, _vp(wq), _vp(ws), n, o, gs, C.byref(h)))
        self._h = h

    def matmul(self, xout: np.ndarray, x) -> None:
        """matmul_q8 / matmul_q4(xout, x, self, n, o, gs)"""
        rows = xout.size // self.o
        xq, xs = _c(x.q, np.int8 if self.q_type == 1 else np.uint8), _c(x.s, np.float32)
        assert xout.dtype == np.float32 and xout.flags.c_contiguous
        check(lib().lmrs_b200_matmul_w(_vp(xout), _vp(xq), _vp(xs), self._h, rows))

    def close(self):
        if self._h:
            lib().lmrs_b200_weights_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
