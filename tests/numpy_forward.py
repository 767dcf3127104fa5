"""Independent numpy restatement of Transformer::forward (src/transformer.rs:316-657) used to cross-check the
C oracle on tiny models: f64 accumulation for the float reductions, the oracle's own quantize/matmul ops for
the integer parts.  Test infrastructure only."""
import numpy as np


class NumpyForward:
    def __init__(self, buf, lf, R):
        self.buf, self.lf, self.R = buf, lf, R
        self.a = a = lf.parse_header(buf)
        self.offs, _ = lf.tensor_offsets(a)
        self.K = np.zeros((a.n_layers, min(a.seq_len, 8192), a.kv_dim), np.float32)
        self.V = np.zeros_like(self.K)

    def qt(self, nm, l, elems):
        a = self.a
        qo, so = self.offs[nm][l]
        qb = elems // 2 if a.q_type == 2 else elems
        return self.buf[qo:qo + qb], self.buf[so:so + elems // 128 * 4].view(np.float32)

    def fv(self, nm, l):
        o = self.offs[nm][l]
        return self.buf[o:o + self.a.dim * 4].view(np.float32)

    def qmm(self, x, nm, l, n, o):
        R, a = self.R, self.a
        wq, ws = self.qt(nm, l, n * o)
        x = np.ascontiguousarray(x, np.float32)
        if a.q_type == 1:
            xq, xs = R.quantize_q8(x, 128)
            return R.matmul_q8(xq, xs, wq.view(np.int8), ws, 1, n, o, 128)
        xq, xs = R.quantize_q4(x, 128)
        return R.matmul_q4(xq, xs, wq, ws, 1, n, o, 128)

    def norm(self, x, w, unit):
        a = self.a
        ss = np.float32(np.sum(x.astype(np.float64) ** 2) / a.dim) + np.float32(a.rms_norm_eps)
        r = np.float32(1.0) / np.sqrt(np.float32(ss))
        return ((1 + w) if unit else w) * (r * x)

    def forward(self, token, pos):
        a, R = self.a, self.R
        gem = a.model_type == 0
        dim, hs = a.dim, a.head_size
        eq, es = self.qt("emb", 0, a.vocab_size * dim)
        row = R.dequantize(eq[token * dim // (2 if a.q_type == 2 else 1):(token + 1) * dim // (2 if a.q_type == 2 else 1)],
                           es[token * dim // 128:(token + 1) * dim // 128], dim, 128, a.q_type)
        x = row * np.float32(np.sqrt(np.float32(dim))) if gem else row
        fr = np.array([R.rope_freq(a.model_type, a.rope_theta, hs, j) for j in range(hs // 2)], np.float32)
        ang = (np.float32(pos) * fr[:, 0]).astype(np.float32)
        c, s_ = np.cos(ang) * fr[:, 1], np.sin(ang) * fr[:, 1]
        kvm = a.n_heads // a.n_kv_heads
        for l in range(a.n_layers):
            hn = self.norm(x, self.fv("rms_att", l), gem).astype(np.float32)
            q = self.qmm(hn, "wq", l, dim, a.att_dim)
            k = self.qmm(hn, "wk", l, dim, a.kv_dim)
            v = self.qmm(hn, "wv", l, dim, a.kv_dim)
            def rope(vec, nh):
                vec = vec.reshape(nh, hs).copy()
                v0, v1 = vec[:, :hs // 2].copy(), vec[:, hs // 2:].copy()
                vec[:, :hs // 2] = v0 * c - v1 * s_
                vec[:, hs // 2:] = v0 * s_ + v1 * c
                return vec.reshape(-1)
            q, k = rope(q, a.n_heads), rope(k, a.n_kv_heads)
            self.K[l, pos], self.V[l, pos] = k, v
            att = np.zeros(a.att_dim, np.float32)
            for h in range(a.n_heads):
                kk = self.K[l, :pos + 1, (h // kvm) * hs:(h // kvm + 1) * hs].astype(np.float64)
                sc = (kk @ q[h * hs:(h + 1) * hs].astype(np.float64)) / np.sqrt(np.float32(hs))
                if gem:
                    sc = 50 * np.tanh(sc / 50)
                p = np.exp(sc - sc.max())
                p /= p.sum()
                att[h * hs:(h + 1) * hs] = p @ self.V[l, :pos + 1, (h // kvm) * hs:(h // kvm + 1) * hs].astype(np.float64)
            wo = self.qmm(att, "wo", l, a.att_dim, dim)
            if gem:
                x = x + self.norm(wo, self.fv("rms_post_att", l), True)
                hin = self.norm(x, self.fv("rms_pre_ffn", l), True)
            else:
                x = x + wo
                hin = self.norm(x, self.fv("rms_post_att", l), False)
            g = self.qmm(hin, "w1", l, dim, a.hidden_dim).astype(np.float64)
            u = self.qmm(hin, "w3", l, dim, a.hidden_dim).astype(np.float64)
            if gem:
                act = g * 0.5 * (1 + np.tanh(0.7978845608028654 * (g + 0.044715 * g ** 3)))
            else:
                act = g / (1 + np.exp(-g))
            dn = self.qmm((act * u).astype(np.float32), "w2", l, a.hidden_dim, dim)
            x = x + (self.norm(dn, self.fv("rms_post_ffn", l), True) if gem else dn)
            x = x.astype(np.float32)
        y = self.norm(x, self.fv("rms_final", 0), gem).astype(np.float32)
        logits = self.qmm(y, "lm_head" if a.model_type == 2 else "emb", 0, dim, a.vocab_size)
        if gem:
            logits[:dim] = (30 * np.tanh(logits[:dim].astype(np.float64) / 30)).astype(np.float32)
        return logits
