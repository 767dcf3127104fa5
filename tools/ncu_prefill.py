"""Workload for ncu captures of the batched prefill (tcgen05 GEMM).  usage: python tools/ncu_prefill.py [model] [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lm.rs_b200"))
import numpy as np
import lmrs_b200
from lmrs_b200 import lmrs_file as lf
model = sys.argv[1] if len(sys.argv) > 1 else "llama-3.2-1b"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
a = lf.model_args(model, 1)
path = f"/tmp/{model}-1.lmrs"
buf = np.fromfile(path, dtype=np.uint8) if os.path.exists(path) else lf.write_synthetic(a, mode="fast")
if not os.path.exists(path): buf.tofile(path)
m, _ = lmrs_b200.Transformer.new(buf)
emb = m.get_embeddings(np.arange(n, dtype=np.uint32) + 5)
import time
for i in range(2):
    e = emb.copy(); t0 = time.perf_counter(); m.fill_kv_cache(e, 0); print("fill ms", (time.perf_counter() - t0) * 1e3)
