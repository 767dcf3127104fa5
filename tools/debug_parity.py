"""Developer probe: per-layer K/V and per-position logits diff between the GPU path and the CPU oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "lm.rs_b200"), os.path.join(ROOT, "oracle")]
import numpy as np
import lmrs_b200, lmrs_ref
from lmrs_b200 import lmrs_file as lf

name, q = sys.argv[1], int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
buf = lf.write_synthetic(lf.model_args(name, q))
cpu = lmrs_ref.RefTransformer(buf)
gpu, _ = lmrs_b200.Transformer.new(buf)
toks = np.random.default_rng(1).integers(0, gpu.args.vocab_size, n)
for pos, t in enumerate(toks):
    lg, le = gpu.forward(int(t), pos), cpu.forward(int(t), pos)
    kc, vc = cpu.kv_cache()
    line = [f"pos {pos} logits {np.abs(lg-le).max():.2e}"]
    for l in range(gpu.args.n_layers):
        k, v = gpu.read_kv(l, pos, 1)
        line.append(f"L{l} k {np.abs(k[0]-kc[l,pos]).max():.2e} v {np.abs(v[0]-vc[l,pos]).max():.2e}")
    print(" | ".join(line))
