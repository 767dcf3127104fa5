"""Mirror of `lmrs::quantization` (src/quantization.rs): QuantType, tensor views, runtime activation quantizers."""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import check, lib
from .transformer import QuantType  # noqa: F401  (re-export, src/quantization.rs:1-6)


@dataclass
class QuantizedTensor:  # src/quantization.rs:8-15 (also used for MutableQuantizedTensor)
    q: np.ndarray
    s: np.ndarray


MutableQuantizedTensor = QuantizedTensor


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def quantize(qx: QuantizedTensor, x, n: int, gs: int) -> None:
    """src/quantization.rs:44-67 -- fills qx.q (int8[n]) and qx.s (float32[n/gs])."""
    x = np.ascontiguousarray(x, np.float32)
    check(lib().lmrs_b200_quantize_q8(_vp(qx.q), _vp(qx.s), _vp(x), n, gs))


def quantize_q4(qx: QuantizedTensor, x, n: int, gs: int) -> None:
    """src/quantization.rs:69-95 -- fills qx.q (uint8[n/2], low nibble = even element) and qx.s."""
    x = np.ascontiguousarray(x, np.float32)
    check(lib().lmrs_b200_quantize_q4(_vp(qx.q), _vp(qx.s), _vp(x), n, gs))
