"""Developer probe (not the benchmark): decode-step time (HBM-resident leg) under environment variants.
usage: python tools/step_time.py "LMRS_B200_LL=0" "LMRS_B200_GEMV_CFG=1" ...   (each argument = one variant, comma-separated assignments)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, os.path.join(%r, "lm.rs_b200"))
import numpy as np, lmrs_b200
from lmrs_b200 import lmrs_file as lf
model, q, pos0, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
a = lf.model_args(model, q)
path = f"/tmp/{model}-{q}.lmrs"
if os.path.exists(path): buf = np.fromfile(path, dtype=np.uint8)
else:
    buf = lf.write_synthetic(a, mode="fast"); buf.tofile(path)
m, _ = lmrs_b200.Transformer.new(buf)
toks = np.random.default_rng(1).integers(0, a.vocab_size, steps + 8)
for i in range(8): m.forward_device(int(toks[i]), pos0 + i)
m.synchronize()
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()
    for i in range(steps): m.forward_device(int(toks[8 + i]), pos0 + i)
    m.synchronize()
    best = min(best, (time.perf_counter() - t0) / steps)
print(f"{best * 1e6:8.1f} us/step  {1 / best:8.1f} tok/s")
''' % ROOT
model = os.environ.get("MODEL", "llama-3.2-1b"); q = os.environ.get("QUANT", "1"); pos = os.environ.get("POS", "512")
for var in sys.argv[1:] or [""]:
    env = dict(os.environ)
    for kv in filter(None, var.split(",")):
        k, v = kv.split("="); env[k] = v
    r = subprocess.run([sys.executable, "-c", CHILD, model, q, pos, "64"], env=env, capture_output=True, text=True)
    print(f"{var or 'default':45s} {r.stdout.strip() or r.stderr[-300:]}", flush=True)
