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
