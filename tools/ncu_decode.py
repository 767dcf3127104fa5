"""Workload for ncu captures: a few decode steps of a synthetic model at a given position.
usage: python tools/ncu_decode.py [model] [q_type] [pos0] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lm.rs_b200"))
import numpy as np
import lmrs_b200
from lmrs_b200 import lmrs_file as lf

model = sys.argv[1] if len(sys.argv) > 1 else "llama-3.2-1b"
q = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pos0 = int(sys.argv[3]) if len(sys.argv) > 3 else 512
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
a = lf.model_args(model, q)
path = f"/tmp/{model}-{q}.lmrs"
if os.path.exists(path):
    buf = np.fromfile(path, dtype=np.uint8)
else:
    buf = lf.write_synthetic(a, mode="fast"); buf.tofile(path)
m, _ = lmrs_b200.Transformer.new(buf)
for i in range(steps):
    m.forward_device(7 + i, pos0 + i)
m.synchronize()
print("done", m.kernel_launches())
