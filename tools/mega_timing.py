"""Developer probe: per-phase time breakdown inside the decode megakernel (LMRS_B200_TIMING=1 stamps).
usage: LMRS_B200_TIMING=1 python tools/mega_timing.py [model] [q] [pos]"""
import os, sys
os.environ["LMRS_B200_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lm.rs_b200"))
import numpy as np
import lmrs_b200
from lmrs_b200 import lmrs_file as lf

model = sys.argv[1] if len(sys.argv) > 1 else "llama-3.2-1b"
q = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pos0 = int(sys.argv[3]) if len(sys.argv) > 3 else 512
a = lf.model_args(model, q)
path = f"/tmp/{model}-{q}.lmrs"
buf = np.fromfile(path, dtype=np.uint8) if os.path.exists(path) else lf.write_synthetic(a, mode="fast")
m, _ = lmrs_b200.Transformer.new(buf)
for p in range(pos0 + 4):
    m.forward_device(7 + p % 100, p)
m.synchronize()
raw = m.debug_buffer("timing").view(np.uint64)
nph = 5 * a.n_layers + 1
sms = raw.size // (nph * 4)
t = raw.reshape(nph, 4, sms).astype(np.int64)
t0 = t[0, 0].min()
names = ["qkv", "attn", "wo", "gateup", "down"]
agg = {}
for ph in range(nph):
    kind = names[ph % 5] if ph < nph - 1 else "cls"
    start = t[ph, 0].max() - t0 if ph == 0 else t[ph - 1, 3].max() - t0
    pro = (t[ph, 1] - t[ph, 0]).mean() if kind != "attn" else 0
    work_end_max = t[ph, 2].max()
    work_end_mean = t[ph, 2].mean()
    bar_end = t[ph, 3].max()
    phase_total = bar_end - (t[ph, 0].min())
    d = agg.setdefault(kind, [])
    d.append((pro, work_end_mean - t[ph, 0].mean(), work_end_max - t[ph, 0].min(), bar_end - work_end_max, phase_total))
print(f"total step {(t[-1,3].max()-t0)/1e3:.1f} us over {nph} phases, {sms} CTAs")
print("phase    n   prologue  work(mean)  work(max)  barrier-after-last  total   [us]")
for k, v in agg.items():
    v = np.array(v, dtype=np.float64) / 1e3
    print(f"{k:7s} {len(v):3d}   {v[:,0].mean():7.2f}   {v[:,1].mean():8.2f}   {v[:,2].mean():8.2f}   {v[:,3].mean():10.2f}      {v[:,4].mean():7.2f}   sum {v[:,4].sum():8.1f}")

# ---- fine trace of CTA 0 for one step ---------------------------------------------------------------------------
m.debug_buffer("trace_reset")
m.forward_device(11, pos0 + 4)
m.synchronize()
tr = m.debug_buffer("trace").view(np.uint64).reshape(-1, 2)
tr = tr[tr[:, 1] > 0]
clk = (tr[:, 0] >> np.uint64(16)).astype(np.int64)
tr = np.stack([(tr[:, 0] & np.uint64(0xffff)).astype(np.int64), tr[:, 1].astype(np.int64)], axis=1)
print("trace events:", len(tr), " SM clock over the step: %.0f MHz" % ((clk[-1] - clk[0]) / ((tr[-1, 1] - tr[0, 1]) / 1e3)))
t00 = int(tr[0, 1])
# print layer 2 (phases 10..14) in detail; tags 300-311 carry cycle counters, not timestamps
sel = False
prev = None
for tag, ts in tr:
    tag = int(tag); ts = int(ts)
    if tag == 1010: sel = True
    if tag == 1016: break
    if sel:
        if 300 <= tag < 400:
            print(f"  tag {tag:5d}  value {ts}")
        else:
            print(f"  tag {tag:5d}  t={ (ts - t00)/1e3:9.2f} us  (+{0 if prev is None else (ts-prev)/1e3:6.2f})")
            prev = ts
