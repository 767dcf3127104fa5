"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): the hot path on tiny models -- batched prefill
(tcgen05 GEMM + fused attention), decode steps on the cluster attention kernel (DSMEM st.async + mbarrier), on-device
argmax -- checked against the oracle so that a sanitizer run also proves the instrumented kernels still compute the
right bits.   usage: compute-sanitizer --tool racecheck python tools/sanitize_smoke.py [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "lm.rs_b200"), os.path.join(ROOT, "oracle")]
import numpy as np
import lmrs_b200, lmrs_ref
from lmrs_b200 import lmrs_file as lf

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for name, q in (("tiny-llama", 1), ("tiny-phi", 2)):
    buf = lf.write_synthetic(lf.model_args(name, q))
    g, _ = lmrs_b200.Transformer.new(buf)
    c = lmrs_ref.RefTransformer(buf)
    toks = np.random.default_rng(2).integers(0, g.args.vocab_size, 12 + steps).astype(np.uint32)
    eg, ec = g.get_embeddings(toks[:12]), c.get_embeddings(toks[:12])
    assert g.fill_kv_cache(eg, 0) == c.fill_kv_cache(ec, 0) == 12 and np.array_equal(eg, ec)
    for i, t in enumerate(toks[12:]):
        assert np.array_equal(g.forward(int(t), 12 + i), c.forward(int(t), 12 + i)), (name, i)
    nxt = g.forward_argmax(int(toks[3]), 12 + steps)
    assert nxt == int(np.argmax(c.forward(int(toks[3]), 12 + steps)))
    g.close(); c.close()
print("sanitize smoke ok")
