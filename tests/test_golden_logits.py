"""The oracle frozen against itself: tests/golden/<name>.oracle.npz (tests/golden/make_golden_logits.py) holds the
residual stream, logits and KV rows the C restatement produced on the exporter-written fixtures when they were committed.
CPU: the oracle must still reproduce them bit for bit.  GPU: liblmrs_b200 must reproduce them without the oracle in the
loop (LLAMA / PHI exactly, GEMMA within 1e-3).  Parity with the Rust binary itself stays unpinned (no rustc here)."""
import os
import sys

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
from make_golden_logits import NAMES, run  # noqa: E402


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_its_frozen_outputs(ref, name):
    want = np.load(os.path.join(GOLD, name + ".oracle.npz"))
    got = run(name)
    for key in ("tokens", "stream", "logits", "k_last", "v_last"):
        assert np.array_equal(got[key], want[key]), f"{name}: oracle output '{key}' changed"


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_the_frozen_oracle_outputs(gpu_lib, name):
    want = np.load(os.path.join(GOLD, name + ".oracle.npz"))
    buf = np.fromfile(os.path.join(GOLD, name + ".lmrs"), dtype=np.uint8)
    m, _ = gpu_lib.Transformer.new(buf)
    exact = m.args.model_type != 0
    toks = want["tokens"]
    emb = m.get_embeddings(toks[:6])
    assert m.fill_kv_cache(emb, 0) == 6
    logits = np.stack([m.forward(int(t), 6 + i).copy() for i, t in enumerate(toks[6:])])
    k, v = m.read_kv(m.args.n_layers - 1, 0, 10)
    for got, key in ((emb, "stream"), (logits, "logits"), (k, "k_last"), (v, "v_last")):
        if exact:
            assert np.array_equal(got, want[key]), f"{name}: {key}"
        else:
            assert float(np.abs(got - want[key]).max()) <= 1e-3, f"{name}: {key}"
    m.close()
