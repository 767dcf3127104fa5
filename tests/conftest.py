import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lm.rs_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ref():
    """The CPU oracle (test infrastructure): oracle/lmrs_ref.py over oracle/_build/liblmrs_ref.so."""
    import lmrs_ref
    lmrs_ref.build()
    return lmrs_ref


@pytest.fixture(scope="session")
def lf():
    from lmrs_b200 import lmrs_file
    return lmrs_file


@pytest.fixture(scope="session")
def gpu_lib():
    """liblmrs_b200.so through the host mirror; fails (not skips) if the extension is missing."""
    import lmrs_b200
    lmrs_b200.lib()
    return lmrs_b200


_model_cache = {}


@pytest.fixture(scope="session")
def synth(lf):
    def make(name, q_type, seed=0, mode="exact"):
        key = (name, q_type, seed, mode)
        if key not in _model_cache:
            _model_cache[key] = lf.write_synthetic(lf.model_args(name, q_type), seed=seed, mode=mode)
        return _model_cache[key]
    return make


def prompt_tokens(vocab, n, seed=1):
    return np.random.default_rng(seed).integers(0, vocab, n).astype(np.uint32)
