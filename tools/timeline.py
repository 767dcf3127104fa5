"""Developer probe: per-launch timeline of one decode step on the default (one kernel per phase, PDL-chained) path.
Needs a -DLMRS_TRACE build (LMRS_B200_SO=...) and runs with LMRS_B200_TIMING=1.
usage: LMRS_B200_SO=lm.rs_b200/lmrs_b200/liblmrs_b200_trace.so python tools/timeline.py [model] [q] [pos]"""
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
for p in range(pos0, pos0 + 8):
    m.forward_device(7 + p % 100, p)
m.synchronize()
nl = 5 * a.n_layers + 1
acc = []
for rep in range(5):
    m.debug_buffer("trace_reset")
    m.forward_device(11, pos0 + 8 + rep)
    m.synchronize()
    raw = m.debug_buffer("trace").view(np.uint64)
    tr = raw[: nl * 8].reshape(nl, 8).astype(np.int64)
    cyc = raw[8192: 8192 + nl * 16].reshape(nl, 16).astype(np.int64)
    acc.append(tr)
names = ["qkv", "attn", "wo", "gateup", "down"]
tr = acc[-1]
t0 = tr[5, 0]
print("launch kind    start   wait_ret  pro_end   end    | prewait  handoff  prologue  main   total-since-prev-end [us]")
rows = {}
def plausible(v):   # a stamp that was never written (0) or belongs to another clock domain is left out of the table
    return abs(int(v) - int(t0)) < 100_000_000
for i in range(nl):
    kind = names[i % 5] if i < nl - 1 else "cls"
    if i < 5 or not all(plausible(tr[j, k]) for j in (i - 1, i) for k in range(4)):
        continue
    s, w, pe, e = (tr[i, 0] - t0) / 1e3, (tr[i, 1] - t0) / 1e3, (tr[i, 2] - t0) / 1e3, (tr[i, 3] - t0) / 1e3
    prev_end = (tr[i - 1, 3] - t0) / 1e3
    rows.setdefault(kind, []).append((w - s, w - prev_end, pe - w, e - pe, e - prev_end))
    if 10 <= i < 21 or i == nl - 1:
        print(f"{i:4d} {kind:7s} {s:8.2f} {w:8.2f} {pe:8.2f} {e:8.2f} | {w - s:7.2f} {w - prev_end:7.2f} {pe - w:8.2f} {e - pe:7.2f} {e - prev_end:7.2f}")
print(f"step: {(tr[-1, 3] - t0) / 1e3:.1f} us (first start -> last end)")
print("kind      n   handoff(prev end -> all waits returned)  prologue   main    per-launch   sum")
for k, v in rows.items():
    v = np.array(v)
    print(f"{k:7s} {len(v):3d}   {v[:, 1].mean():8.2f}   {v[:, 2].mean():8.2f}  {v[:, 3].mean():8.2f}  {v[:, 4].mean():8.2f}  {v[:, 4].sum():8.1f}")

ghz = 1.965
print("\ncycle-resolution phases (max over CTAs of clock64 since the CTA's entry), us at %.3f GHz; layers 1..n-1 averaged" % ghz)
def avg(kind_idx, k):
    sel = [i for i in range(5, nl - 1) if i % 5 == kind_idx]
    return np.mean([cyc[i, k] for i in sel]) / ghz / 1e3
for ki, kind in enumerate(names):
    if kind == "attn":
        labels = {1: "wait returned", 2: "rope done", 3: "K visible", 4: "own scores", 5: "cluster barrier",
                  7: "max+exp done", 8: "serial sum done", 9: "V visible", 10: "divide done", 11: "a*v done"}
        print("attn   :", "  ".join(f"{labels[k]} {avg(ki, k):.2f}" for k in sorted(labels)))
    else:
        print(f"{kind:7s}: since entry: wait returned {avg(ki, 1):.2f}  prologue done {avg(ki, 2):.2f}  end {avg(ki, 3):.2f} |"
              f" inside prologue: inputs {avg(ki, 4):.2f}  1/rms {avg(ki, 5):.2f}  quantized {avg(ki, 6):.2f}")
