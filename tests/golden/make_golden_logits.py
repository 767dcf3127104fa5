"""Freezes the ORACLE's outputs on the exporter-written fixtures (tests/golden/*.lmrs) so that a later change to
oracle/lmrs_ref.c cannot silently move the goalposts of every parity test:

    tests/golden/<name>.oracle.npz   tokens, residual stream after a 6-embedding fill_kv_cache, logits of 4 decode steps,
                                     K / V rows of the last block

These are outputs of the C RESTATEMENT, not of the Rust binary (no rustc in this image): they pin the oracle against
itself over time, the exporter fixtures pin the file format -- parity with the real reference stays "unpinned".
usage: python tests/golden/make_golden_logits.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle")]
import lmrs_ref  # noqa: E402

NAMES = ("ref_llama_q8", "ref_gemma_q4", "ref_phi_q8")


def run(name):
    buf = np.fromfile(os.path.join(HERE, name + ".lmrs"), dtype=np.uint8)
    m = lmrs_ref.RefTransformer(buf)
    toks = np.random.default_rng(7).integers(0, m.args.vocab_size, 10).astype(np.uint32)
    emb = m.get_embeddings(toks[:6])
    assert m.fill_kv_cache(emb, 0) == 6
    logits = np.stack([m.forward(int(t), 6 + i).copy() for i, t in enumerate(toks[6:])])
    k, v = m.kv_cache()
    L = m.args.n_layers
    return dict(tokens=toks, stream=emb, logits=logits, k_last=k[L - 1, :10].copy(), v_last=v[L - 1, :10].copy())


if __name__ == "__main__":
    for n in NAMES:
        np.savez_compressed(os.path.join(HERE, n + ".oracle.npz"), **run(n))
        print("wrote", n + ".oracle.npz")
