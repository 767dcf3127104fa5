"""Developer probe (not the benchmark): time the decode attention launches of one step alone, per position, optionally
with phases knocked out (needs a -DLMRS_DEV_PROBES build: LMRS_B200_SO=..., LMRS_B200_DEV_SKIP bit mask:
1 scores, 2 exp, 4 serial sum, 8 a*v, 16 DSMEM gather, 32 divide, 64 K/V loads).
usage: python tools/attn_probe.py [model] [q_type] [pos,pos,...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lm.rs_b200"))
import numpy as np
import torch
import lmrs_b200
from lmrs_b200 import lmrs_file as lf

model = sys.argv[1] if len(sys.argv) > 1 else "llama-3.2-1b"
q = int(sys.argv[2]) if len(sys.argv) > 2 else 1
poss = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "64,576,1000,2000").split(",")]
a = lf.model_args(model, q)
path = f"/tmp/{model}-{q}.lmrs"
if os.path.exists(path):
    buf = np.fromfile(path, dtype=np.uint8)
else:
    buf = lf.write_synthetic(a, mode="fast"); buf.tofile(path)
m, _ = lmrs_b200.Transformer.new(buf)
st = torch.cuda.Stream()
m.set_stream(st.cuda_stream)
m.forward_device(1, 0); m.synchronize()
out = []
for pos in poss:
    if pos >= a.seq_len:
        continue
    for _ in range(3):
        n = m.bench_attn_pass(pos)
    m.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(reps):
            m.bench_attn_pass(pos)
        e1.record(st)
    e1.synchronize()
    out.append(f"pos {pos}: {e0.elapsed_time(e1) / reps / n * 1e3:.2f} us/launch")
print(os.environ.get("LMRS_B200_ATT_CLUSTER", "8"), os.environ.get("LMRS_B200_DEV_SKIP", "0"), " | ".join(out), flush=True)
