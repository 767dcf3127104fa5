"""Developer probe (not the benchmark): decode-step time of a synthetic model under several kernel configs.
usage: python tools/quick_decode.py [model] [q_type] [pos0] [steps]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lm.rs_b200"))


def child(model, q_type, pos0, steps):
    import numpy as np
    import lmrs_b200
    from lmrs_b200 import lmrs_file as lf
    a = lf.model_args(model, q_type)
    path = f"/tmp/{model}-{q_type}.lmrs"
    if os.path.exists(path):
        buf = np.fromfile(path, dtype=np.uint8)
    else:
        buf = lf.write_synthetic(a, mode="fast")
        buf.tofile(path)
    m, _ = lmrs_b200.Transformer.new(buf)
    rng = np.random.default_rng(1)
    toks = rng.integers(0, a.vocab_size, pos0 + steps + 8)
    t0 = time.time()
    for p in range(pos0):
        m.forward_device(int(toks[p]), p)
    m.synchronize()
    t_fill = time.time() - t0
    for p in range(pos0, pos0 + 8):
        m.forward_device(int(toks[p]), p)
    m.synchronize()
    t0 = time.time()
    for p in range(pos0 + 8, pos0 + 8 + steps):
        m.forward_device(int(toks[p]), p)
    m.synchronize()
    dt = (time.time() - t0) / steps
    by = lf.decode_bytes_per_token(a, pos0 + 8 + steps // 2)
    print(json.dumps({"us_per_tok": round(dt * 1e6, 1), "tok_s": round(1 / dt, 1), "GBs": round(by / dt / 1e9, 1),
                      "fill_us_per_tok": round(t_fill / max(pos0, 1) * 1e6, 1)}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
        sys.exit(0)
    model = sys.argv[1] if len(sys.argv) > 1 else "llama-3.2-1b"
    q_type = sys.argv[2] if len(sys.argv) > 2 else "1"
    pos0 = sys.argv[3] if len(sys.argv) > 3 else "512"
    steps = sys.argv[4] if len(sys.argv) > 4 else "128"
    configs = [dict(), dict(LMRS_B200_PDL="0"), dict(LMRS_B200_GRAPH="0"), dict(LMRS_B200_GRAPH="0", LMRS_B200_PDL="0")]
    for cfg in range(5):
        for ctas in (1, 2):
            configs.append(dict(LMRS_B200_GEMV_CFG=str(cfg), LMRS_B200_GEMV_CTAS=str(ctas)))
    for ns in (4, 8, 32):
        configs.append(dict(LMRS_B200_NSPLIT=str(ns)))
    if os.environ.get("QUICK_CONFIGS"):   # e.g. QUICK_CONFIGS='[{}, {"LMRS_B200_ATT_CLUSTER": "0"}]'
        configs = json.loads(os.environ["QUICK_CONFIGS"])
    for c in configs:
        env = dict(os.environ, **c)
        r = subprocess.run([sys.executable, __file__, "--child", model, q_type, pos0, steps], env=env, capture_output=True, text=True)
        print(c, r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
