"""lmrs_b200 -- Blackwell (sm_100a) implementation of the lm.rs quantized transformer forward path.

The product is liblmrs_b200.so (hand-written CUDA behind the C ABI of include/lmrs_b200.h); this package is
the thin host-side mirror of the reference crate's modules used by tests and bench.py:
    lmrs_b200.transformer   <->  src/transformer.rs
    lmrs_b200.functional    <->  src/functional.rs
    lmrs_b200.quantization  <->  src/quantization.rs
    lmrs_b200.lmrs_file     <->  export.py / utils/ (LMRS v4 layout, synthetic-model writer)
"""
from ._lib import ABI_SYMBOLS, SO_PATH, Args, LmrsError, build, lib  # noqa: F401
from .transformer import ModelType, QuantType, Transformer, nccl_unique_id  # noqa: F401
from . import functional, lmrs_file, quantization  # noqa: F401
