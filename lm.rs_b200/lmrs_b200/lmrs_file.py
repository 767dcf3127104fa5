"""LMRS v4 model-file layout: header pack/parse, tensor table, synthetic-model writer.

The reference never ships weights and this box has no network, so every test and the
benchmark run on *synthetic* LMRS files that carry the real models' headers and shapes.
This module restates the writer side of the format:

  * header            export.py:54-84      (reader: src/transformer.rs:134-160)
  * tensor order      export.py:87-125     (reader: src/transformer.rs:241-270)
  * Q(n) = q bytes then n/gs f32 scales, per layer   utils/io.py:21-52 (reader :24-48)
  * Q8_0 weights      utils/quantization.py:42-66  (scale = max|w|/127, torch.round = half-even)
  * Q4_0 weights      utils/quantization.py:4-39   (scale = max|w|/-7.5, nibble = round(w/s+8), lo = even)

It is host-side utility code (numpy only); nothing here touches the GPU.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, asdict

import numpy as np

MAGIC = 0x73726D6C  # "lmrs"
VERSION = 4
HEADER_BYTES = 256
GEMMA, LLAMA, PHI = 0, 1, 2
Q_NONE, Q8_0, Q4_0 = 0, 1, 2
MODEL_TYPE_NAMES = {GEMMA: "GEMMA", LLAMA: "LLAMA", PHI: "PHI"}


@dataclass
class LmrsArgs:
    """The 47-byte packed TransformerArgs (src/transformer.rs:57-74)."""
    dim: int
    hidden_dim: int
    n_layers: int
    n_heads: int
    head_size: int
    n_kv_heads: int
    vocab_size: int
    seq_len: int
    rms_norm_eps: float
    rope_theta: float
    q_type: int = Q8_0
    model_type: int = LLAMA
    group_size: int = 128
    multimodal: int = 0

    @property
    def att_dim(self): return self.n_heads * self.head_size
    @property
    def kv_dim(self): return self.n_kv_heads * self.head_size

    def pack(self) -> bytes:
        b = struct.pack("<II", MAGIC, VERSION)
        b += struct.pack("<IIIIIIIIff", self.dim, self.hidden_dim, self.n_layers, self.n_heads, self.head_size,
                         self.n_kv_heads, self.vocab_size, self.seq_len, self.rms_norm_eps, self.rope_theta)
        b += struct.pack("<BB", self.q_type, self.model_type)
        b += struct.pack("<I", self.group_size)  # unaligned at offset 50, as in export.py:78
        b += struct.pack("<B", self.multimodal)
        assert len(b) == 55
        return b + b"\0" * (HEADER_BYTES - len(b))


def parse_header(buf) -> LmrsArgs:
    b = bytes(memoryview(buf)[:HEADER_BYTES])
    magic, version = struct.unpack_from("<II", b, 0)
    if magic != MAGIC:
        raise ValueError("Model not in lm.rs format.")
    f = struct.unpack_from("<IIIIIIIIffBBIB", b, 8)
    return LmrsArgs(*f)


# name -> header values of the checkpoints the reference README lists (SURVEY.md section 8 preamble)
MODEL_SHAPES = {
    "llama-3.2-1b": dict(dim=2048, hidden_dim=8192, n_layers=16, n_heads=32, head_size=64, n_kv_heads=8,
                         vocab_size=128256, seq_len=131072, rms_norm_eps=1e-5, rope_theta=500000.0, model_type=LLAMA),
    "llama-3.2-3b": dict(dim=3072, hidden_dim=8192, n_layers=28, n_heads=24, head_size=128, n_kv_heads=8,
                         vocab_size=128256, seq_len=131072, rms_norm_eps=1e-5, rope_theta=500000.0, model_type=LLAMA),
    "gemma-2-2b": dict(dim=2304, hidden_dim=9216, n_layers=26, n_heads=8, head_size=256, n_kv_heads=4,
                       vocab_size=256000, seq_len=8192, rms_norm_eps=1e-6, rope_theta=10000.0, model_type=GEMMA),
    "gemma-2-9b": dict(dim=3584, hidden_dim=14336, n_layers=42, n_heads=16, head_size=256, n_kv_heads=8,
                       vocab_size=256000, seq_len=8192, rms_norm_eps=1e-6, rope_theta=10000.0, model_type=GEMMA),
    "phi-3.5-mini": dict(dim=3072, hidden_dim=8192, n_layers=32, n_heads=32, head_size=96, n_kv_heads=32,
                         vocab_size=32064, seq_len=131072, rms_norm_eps=1e-5, rope_theta=10000.0, model_type=PHI),
    # CI-sized models exercising every structural case
    "tiny-llama": dict(dim=256, hidden_dim=512, n_layers=2, n_heads=4, head_size=64, n_kv_heads=2,
                       vocab_size=512, seq_len=256, rms_norm_eps=1e-5, rope_theta=500000.0, model_type=LLAMA),
    "tiny-gemma": dict(dim=256, hidden_dim=512, n_layers=2, n_heads=4, head_size=128, n_kv_heads=2,  # att_dim > dim
                       vocab_size=512, seq_len=256, rms_norm_eps=1e-6, rope_theta=10000.0, model_type=GEMMA),
    "tiny-gemma-narrow": dict(dim=384, hidden_dim=512, n_layers=2, n_heads=2, head_size=128, n_kv_heads=1,  # att_dim < dim
                              vocab_size=512, seq_len=256, rms_norm_eps=1e-6, rope_theta=10000.0, model_type=GEMMA),
    "tiny-gemma-narrow2": dict(dim=640, hidden_dim=512, n_layers=2, n_heads=4, head_size=128, n_kv_heads=2,  # att_dim 512 < dim: batches of <= 3 rows are defined
                               vocab_size=512, seq_len=256, rms_norm_eps=1e-6, rope_theta=10000.0, model_type=GEMMA),
    "tiny-phi": dict(dim=384, hidden_dim=512, n_layers=2, n_heads=4, head_size=96, n_kv_heads=4,
                     vocab_size=512, seq_len=256, rms_norm_eps=1e-5, rope_theta=10000.0, model_type=PHI),
    "small-llama": dict(dim=1024, hidden_dim=2048, n_layers=4, n_heads=16, head_size=64, n_kv_heads=4,
                        vocab_size=4096, seq_len=2048, rms_norm_eps=1e-5, rope_theta=500000.0, model_type=LLAMA),
}


def model_args(name: str, q_type: int = Q8_0, **overrides) -> LmrsArgs:
    kw = dict(MODEL_SHAPES[name])
    kw.update(overrides)
    return LmrsArgs(q_type=q_type, **kw)


def tensor_table(a: LmrsArgs):
    """[(name, kind, n_layers, elems_each)] in file order; kind 'f' = f32, 'q' = quantised Q(n).
    export.py:87-125 / src/transformer.rs:241-270 (and :169-195 for q_type 0)."""
    L, dim, hd = a.n_layers, a.dim, a.hidden_dim
    wk = "q" if a.q_type != Q_NONE else "f"
    t = [("emb", wk, 1, a.vocab_size * dim), ("rms_att", "f", L, dim),
         ("wq", wk, L, dim * a.att_dim), ("wk", wk, L, dim * a.kv_dim), ("wv", wk, L, dim * a.kv_dim),
         ("wo", wk, L, dim * a.att_dim), ("rms_post_att", "f", L, dim)]
    if a.model_type == GEMMA:
        t.append(("rms_pre_ffn", "f", L, dim))
    t += [("w1", wk, L, dim * hd), ("w2", wk, L, dim * hd), ("w3", wk, L, dim * hd)]
    if a.model_type == GEMMA:
        t.append(("rms_post_ffn", "f", L, dim))
    t.append(("rms_final", "f", 1, dim))
    if a.model_type == PHI:
        t.append(("lm_head", wk, 1, dim * a.vocab_size))
    return t


def q_bytes(a: LmrsArgs, elems: int) -> int:
    return elems // 2 if a.q_type == Q4_0 else elems


def tensor_offsets(a: LmrsArgs):
    """{name: [(q_off, s_off) | f_off per layer]} and the end offset (= where the vision section starts)."""
    off, out = HEADER_BYTES, {}
    for name, kind, n, elems in tensor_table(a):
        lst = []
        for _ in range(n):
            if kind == "f":
                lst.append(off)
                off += elems * 4
            else:
                qo = off
                off += q_bytes(a, elems)
                lst.append((qo, off))
                off += elems // a.group_size * 4
        out[name] = lst
    return out, off


def file_size(a: LmrsArgs) -> int:
    return tensor_offsets(a)[1]


def decode_bytes_per_token(a: LmrsArgs, pos: int | None = None) -> int:
    """ALGORITHMIC HBM bytes one decode step must move: every layer matrix and the
    classifier (q bytes + scales) once, the norm vectors, plus -- if pos is given -- the K/V rows read
    ((pos+1) * kv_dim * 4 * 2 per layer) and written (kv_dim*4*2 per layer)."""
    total = 0
    for name, kind, n, elems in tensor_table(a):
        if name == "emb" and a.model_type == PHI:
            continue  # PHI has a separate lm_head; the embedding table is only gathered (1 row)
        if kind == "f":
            total += n * elems * 4
        else:
            total += n * (q_bytes(a, elems) + elems // a.group_size * 4)
    if pos is not None:
        total += a.n_layers * a.kv_dim * 4 * 2 * (pos + 1)  # K and V rows read
        total += a.n_layers * a.kv_dim * 4 * 2              # this token's K and V written
    return total


def prefill_int8_ops(a: LmrsArgs, t: int) -> int:
    """2 * MACs of the layer matmuls for t tokens (fill_kv_cache runs no classifier)."""
    per_tok = a.n_layers * (a.dim * a.att_dim * 2 + a.dim * a.kv_dim * 2 + 3 * a.dim * a.hidden_dim)
    return 2 * per_tok * t


# ---- weight quantisers of the exporter (numpy restatement; np.round == torch.round == half-even) ----------

def quantize_q80(w: np.ndarray, gs: int = 128):
    """utils/quantization.py:42-66"""
    w = w.astype(np.float32).reshape(-1, gs)
    wmax = np.abs(w).max(axis=1)
    scale = (wmax / np.float32(127.0)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        quant = w / scale[:, None]
    q = np.nan_to_num(np.round(quant), nan=0.0).astype(np.int8)
    return q.reshape(-1), scale


def quantize_q40(w: np.ndarray, gs: int = 128):
    """utils/quantization.py:4-39"""
    w = w.astype(np.float32).reshape(-1, gs)
    wmax = np.abs(w).max(axis=1)
    scale = (wmax / np.float32(-7.5)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        quant = w / scale[:, None]
    u = np.clip(np.nan_to_num(np.round(quant + 8), nan=0.0), 0, 255).astype(np.uint8).clip(0, 15)
    u = u.reshape(u.shape[0], gs // 2, 2)
    packed = (u[..., 0] | (u[..., 1] << 4)).astype(np.uint8)
    return packed.reshape(-1), scale


def write_synthetic(a: LmrsArgs, seed: int = 0, mode: str = "exact") -> np.ndarray:
    """Return the whole LMRS file as a uint8 array (64-byte aligned by numpy).

    mode "exact": f32 weights ~ N(0, 1/n_in) quantised with the exporter's formulas (tests, small models).
    mode "fast":  int8/nibble codes drawn directly (uniform) with per-group scales U(0.5,1.5)*sigma/std(code);
                  same value distribution in aggregate, seconds instead of minutes for GB-sized files.
    Norm weights are 1 + 0.1 N(0,1) (GEMMA: 0.1 N(0,1), the model applies 1+w)."""
    offs, end = tensor_offsets(a)
    buf = np.zeros(end, dtype=np.uint8)
    buf[:HEADER_BYTES] = np.frombuffer(a.pack(), dtype=np.uint8)
    gs = a.group_size
    for ti, (name, kind, n, elems) in enumerate(tensor_table(a)):
        rng = np.random.default_rng([seed, ti])
        n_in = {"emb": a.dim, "wq": a.dim, "wk": a.dim, "wv": a.dim, "wo": a.att_dim, "w1": a.dim, "w3": a.dim,
                "w2": a.hidden_dim, "lm_head": a.dim}.get(name, a.dim)
        sigma = 0.05 if name == "emb" else 1.0 / np.sqrt(n_in)
        for l in range(n):
            if kind == "f" and name.startswith("rms"):
                w = (0.1 * rng.standard_normal(elems)).astype(np.float32)
                if a.model_type != GEMMA:
                    w += np.float32(1.0)
                buf[offs[name][l]:offs[name][l] + elems * 4] = w.view(np.uint8)
            elif kind == "f":
                w = (sigma * rng.standard_normal(elems)).astype(np.float32)
                buf[offs[name][l]:offs[name][l] + elems * 4] = w.view(np.uint8)
            else:
                qo, so = offs[name][l]
                ng = elems // gs
                if mode == "exact":
                    w = (sigma * rng.standard_normal(elems)).astype(np.float32)
                    q, s = quantize_q80(w, gs) if a.q_type == Q8_0 else quantize_q40(w, gs)
                    q = q.view(np.uint8)
                elif a.q_type == Q8_0:
                    q = rng.integers(-127, 128, size=elems, dtype=np.int8).view(np.uint8)
                    s = (rng.uniform(0.5, 1.5, ng) * (sigma / 73.3)).astype(np.float32)
                else:
                    q = rng.integers(0, 256, size=elems // 2, dtype=np.uint8)
                    s = (-rng.uniform(0.5, 1.5, ng) * (sigma / 4.61)).astype(np.float32)
                buf[qo:qo + q.size] = q
                buf[so:so + ng * 4] = s.view(np.uint8)
    return buf


def describe(a: LmrsArgs) -> dict:
    d = asdict(a)
    d["model_type"] = MODEL_TYPE_NAMES[a.model_type]
    d["q_type"] = {0: "f32", 1: "Q8_0", 2: "Q4_0"}[a.q_type]
    return d
