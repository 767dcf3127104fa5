"""Developer probe: re-derive the LAST block's intermediates from the GPU's own buffers with oracle ops
and report where the GPU chain first deviates.  usage: debug_layer.py model q_type npos"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "lm.rs_b200"), os.path.join(ROOT, "oracle")]
import numpy as np
import lmrs_b200, lmrs_ref as R
from lmrs_b200 import lmrs_file as lf

name, q, npos = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
a = lf.model_args(name, q)
buf = lf.write_synthetic(a)
offs, _ = lf.tensor_offsets(a)
gpu, _ = lmrs_b200.Transformer.new(buf)
cpu = R.RefTransformer(buf)
L = a.n_layers - 1
gs = 128

def qt(nm, l, elems):
    qo, so = offs[nm][l]
    qb = elems // 2 if q == 2 else elems
    return buf[qo:qo + qb], buf[so:so + elems // gs * 4].view(np.float32)
def fv(nm, l):
    o = offs[nm][l]; return buf[o:o + a.dim * 4].view(np.float32)
def qmm(x, nm, l, n, o):
    wq, ws = qt(nm, l, n * o)
    if q == 1:
        xq, xs = R.quantize_q8(x, gs); return R.matmul_q8(xq, xs, wq.view(np.int8), ws, 1, n, o, gs)
    xq, xs = R.quantize_q4(x, gs); return R.matmul_q4(xq, xs, wq, ws, 1, n, o, gs)

toks = np.random.default_rng(1).integers(0, a.vocab_size, npos)
for pos, t in enumerate(toks):
    lg = gpu.forward(int(t), pos); le = cpu.forward(int(t), pos)
    B = {k: gpu.debug_buffer(k) for k in ("x0", "x1", "q", "k_new", "att", "wo_out", "h", "down_out")}
    hs, kvm = a.head_size, a.n_heads // a.n_kv_heads
    K, V = gpu.read_kv(L, 0, pos + 1)
    # attention reference (float64) from the GPU's q and cache
    att = np.zeros(a.att_dim)
    for h in range(a.n_heads):
        qv = B["q"][h * hs:(h + 1) * hs].astype(np.float64)
        # rope q
        fr = np.array([R.rope_freq(a.model_type, a.rope_theta, hs, j) for j in range(hs // 2)])
        ang = np.float32(pos) * fr[:, 0].astype(np.float32)
        c, s_ = np.cos(ang.astype(np.float32)) * fr[:, 1], np.sin(ang.astype(np.float32)) * fr[:, 1]
        q0, q1 = qv[:hs // 2].copy(), qv[hs // 2:].copy()
        qr = np.concatenate([q0 * c - q1 * s_, q0 * s_ + q1 * c])
        kk = K[:, (h // kvm) * hs:(h // kvm + 1) * hs].astype(np.float64)
        sc = kk @ qr / np.sqrt(np.float32(hs))
        if a.model_type == 0:
            sc = 50 * np.tanh(sc / 50)
        p = np.exp(sc - sc.max()); p /= p.sum()
        att[h * hs:(h + 1) * hs] = p @ V[:, (h // kvm) * hs:(h // kvm + 1) * hs].astype(np.float64)
    d_att = np.abs(att - B["att"]).max()
    wo = qmm(B["att"], "wo", L, a.att_dim, a.dim)
    d_wo = np.abs(wo - B["wo_out"]).max()
    # ffn from x1 + wo_out
    if a.model_type == 0:
        xn = B["x1"] + R.rmsnorm(B["wo_out"], fv("rms_post_att", L), a.rms_norm_eps, True)
        hin = R.rmsnorm(xn, fv("rms_pre_ffn", L), a.rms_norm_eps, True)
    else:
        xn = B["x1"] + B["wo_out"]
        hin = R.rmsnorm(xn, fv("rms_post_att", L), a.rms_norm_eps, False)
    d_x0 = np.abs(xn - B["x0"]).max()
    g = qmm(hin, "w1", L, a.dim, a.hidden_dim); u = qmm(hin, "w3", L, a.dim, a.hidden_dim)
    if a.model_type == 0:
        inner = (g + np.float32(0.044715) * g * g * g).astype(np.float32)
        act = g * (np.float32(0.5) * (np.float32(1) + np.tanh(0.7978845608028654 * inner.astype(np.float64)).astype(np.float32)))
    else:
        act = g * (np.float32(1) / (np.float32(1) + np.exp(-g)))
    hh = (act * u).astype(np.float32)
    d_h = np.abs(hh - B["h"]).max()
    dn = qmm(B["h"], "w2", L, a.hidden_dim, a.dim)
    d_dn = np.abs(dn - B["down_out"]).max()
    print(f"pos {pos} logits {np.abs(lg-le).max():.2e} | att {d_att:.2e} wo {d_wo:.2e} x0 {d_x0:.2e} h {d_h:.2e} (|h| {np.abs(hh).max():.2f}) down {d_dn:.2e}")
