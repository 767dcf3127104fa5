"""Known-answer tests that pin the CPU oracle (oracle/lmrs_ref.c) to the reference SOURCE semantics.

The reference ships no tests or golden vectors for this path and its Rust crate cannot be built here
("parity unpinned"), so these vectors are hand-derived from the cited lines."""
import numpy as np
import pytest


def test_quantize_q8_round_half_away_from_zero(ref):
    # src/quantization.rs:57-64: scale = max/127, f32::round is half-away-from-zero (numpy would give 0,0,2)
    q, s = ref.quantize_q8(np.array([127, -127, 63.5, -63.5, 0.5, -0.5, 2.5, 0], np.float32), 8)
    assert s.tolist() == [1.0]
    assert q.tolist() == [127, -127, 64, -64, 1, -1, 3, 0]


def test_quantize_q8_zero_group_nan_casts_to_zero(ref):
    # scale 0 -> x/0 = NaN -> `NaN as i8` = 0 (Rust saturating cast)
    q, s = ref.quantize_q8(np.zeros(16, np.float32), 8)
    assert s.tolist() == [0.0, 0.0] and not q.any()


def test_quantize_q4_negative_scale_and_packing(ref):
    # src/quantization.rs:71,82,89-92: scale = max/-8, nibble = clamp(round(x/s + 8), 0, 15), low nibble = even index
    x = np.array([8, -8, 4, -4, 0, 1, -1, 7.5], np.float32)
    q, s = ref.quantize_q4(x, 8)
    assert s.tolist() == [-1.0]
    assert q.tobytes().hex() == "f0c47819"
    assert ref.dequantize(q, s, 8, 8, 2).tolist() == [8, -7, 4, -4, 0, 1, -1, 7]


def test_matmul_q8_f32_accumulation_order(ref):
    # src/functional.rs:207: xout += ((ival as f32) * ws) * xs per group, ascending
    xq = np.array(list(range(1, 9)) + [-i for i in range(1, 9)], np.int8)
    wq = np.tile(np.array([127] * 8 + [-127] * 8, np.int8), 4)
    out = ref.matmul_q8(xq, [0.5, 0.25], wq, [0.01, 0.02] * 4, 1, 16, 4, 8)
    assert out.view(np.uint32).tolist() == [0x4236E147] * 4          # 45.719997
    g0 = np.float32(np.float32(4572) * np.float32(0.01)) * np.float32(0.5)
    g1 = np.float32(np.float32(4572) * np.float32(0.02)) * np.float32(0.25)
    assert out[0] == np.float32(g0 + g1)


def test_matmul_q8_drops_rows_beyond_multiple_of_4(ref):
    # par_chunks_exact_mut(4) (src/functional.rs:179): trailing o % 4 rows are never written
    xq = np.ones(8, np.int8); wq = np.ones(6 * 8, np.int8)
    out = ref.matmul_q8(xq, [1.0], wq, [1.0] * 6, 1, 8, 6, 8)
    assert out.tolist() == [8, 8, 8, 8, 0, 0]


def test_matmul_q4_nibble_semantics(ref):
    # src/functional.rs:218,236-246: (nibble - 8) on both sides, lo nibble = even element; the SIMD loop covers
    # (gs/2)/8 chunks per group, so gs=16 is the smallest group the reference processes at all (gs=8 -> zero work)
    x = np.array([1, -2, 3, -4, 5, -6, 7, -8] * 2, np.float32)
    xq, xs = ref.quantize_q4(x, 16)
    w = np.array([[-8, 7, -6, 5, -4, 3, -2, 1] * 2] * 4, np.float32).reshape(-1)
    wq, ws = ref.quantize_q4(w, 16)
    out = ref.matmul_q4(xq, xs, wq, ws, 1, 16, 4, 16)
    xd, wd = ref.dequantize(xq, xs, 16, 16, 2), ref.dequantize(wq, ws, 64, 16, 2)[:16]
    assert abs(out[0] - float(np.dot(xd.astype(np.float64), wd.astype(np.float64)))) < 1e-4
    assert ref.matmul_q4(xq[:4], xs[:1], wq[:16], ws[:4], 1, 8, 4, 8).tolist() == [0, 0, 0, 0]


def test_llama3_rope_bands(ref):
    # src/transformer.rs:451-469 with theta 5e5, head 64: j<=14 unchanged, 15..17 smoothed, >=18 divided by 32
    base = lambda j: np.float32(1.0) / np.float32(np.float32(500000.0) ** np.float32(np.float32(2 * j) / np.float32(64)))
    for j in range(0, 15):
        assert ref.rope_freq(1, 500000.0, 64, j)[0] == pytest.approx(float(base(j)), rel=2e-7)
    assert ref.rope_freq(1, 500000.0, 64, 15)[0] == pytest.approx(1.290548e-3, rel=1e-6)
    assert ref.rope_freq(1, 500000.0, 64, 18)[0] == pytest.approx(1.946164e-5, rel=1e-6)
    for j in range(18, 32):
        assert ref.rope_freq(1, 500000.0, 64, j)[0] == pytest.approx(float(base(j)) / 32, rel=1e-6)
    # GEMMA: plain
    assert ref.rope_freq(0, 10000.0, 256, 3)[0] == pytest.approx(float(np.float32(1.0) / np.float32(10000.0) ** np.float32(6 / 256)), rel=1e-6)


def test_phi_rope_short_factor_and_scale(ref):
    f, m = ref.rope_freq(2, 10000.0, 96, 0)      # src/transformer.rs:472-478
    assert f == pytest.approx(1 / 1.08, rel=1e-6) and m == pytest.approx(1.19023807, rel=1e-6)
    f47, _ = ref.rope_freq(2, 10000.0, 96, 47)
    assert f47 == pytest.approx((10000.0 ** (-94 / 96)) / 8.999999999999853, rel=1e-5)


def test_rmsnorm_lane_order(ref):
    # src/functional.rs:48-78: eight lane sums then horizontal add; (1+w) variant for Gemma
    rng = np.random.default_rng(0)
    x = rng.standard_normal(64).astype(np.float32); w = rng.standard_normal(64).astype(np.float32)
    lanes = np.zeros(8, np.float32)
    for j in range(8):
        lanes += x[8 * j:8 * j + 8] * x[8 * j:8 * j + 8]
    ss = ((lanes[0] + lanes[4]) + (lanes[2] + lanes[6])) + ((lanes[1] + lanes[5]) + (lanes[3] + lanes[7]))
    ss = np.float32(ss / np.float32(64)) + np.float32(1e-5)
    r = np.float32(1.0) / np.sqrt(np.float32(ss))
    assert np.array_equal(ref.rmsnorm(x, w, 1e-5, False), w * (r * x))
    assert np.array_equal(ref.rmsnorm(x, w, 1e-5, True), (np.float32(1) + w) * (r * x))


def test_softmax_serial(ref):
    x = np.array([1.0, 2.0, 3.0, -1e30], np.float32)
    e = np.exp((x - np.float32(3.0)).astype(np.float32)).astype(np.float32)
    s = np.float32(0)
    for v in e:
        s = np.float32(s + v)
    got = ref.softmax(x)
    np.testing.assert_allclose(got, e / s, rtol=1e-6)
    assert got[3] == 0.0


def test_matmul_rest_tail_bug_is_kept(ref):
    # src/functional.rs:273-275: the tail multiplies by x[r] (row 0), not x[xi+r]
    n, o = 10, 2
    x = np.arange(20, dtype=np.float32); w = np.ones(o * n, np.float32)
    out = ref.matmul_f32(x, w, 2, n, o, rest=True)
    assert out[0] == x[:10].sum()
    assert out[2] == x[10:18].sum() + x[8] + x[9]        # second row re-uses row 0's tail elements


def test_forward_quirks_on_tiny_gemma(ref, lf):
    a = lf.model_args("tiny-gemma", 1)
    m = ref.RefTransformer(lf.write_synthetic(a))
    lg = m.forward(3, 0)
    assert np.isfinite(lg).all()
    # soft-cap 30*tanh(x/30) only on the first `dim` vocabulary entries (src/transformer.rs:375-381)
    assert np.abs(lg[:a.dim]).max() < 30.0
    assert m.args.seq_len == 256
    with pytest.raises(RuntimeError):
        m.forward(a.vocab_size, 0)


def test_fill_kv_cache_equals_token_by_token_for_llama(ref, lf):
    """sl>1 forward_layer (src/transformer.rs:672-684) is the same math as sl=1 steps for LLAMA/PHI."""
    buf = lf.write_synthetic(lf.model_args("tiny-llama", 1))
    a, b = ref.RefTransformer(buf), ref.RefTransformer(buf)
    toks = np.array([5, 9, 200, 31, 7, 77], np.uint32)
    emb = a.get_embeddings(toks)
    assert a.fill_kv_cache(emb, 0) == 6
    for p, t in enumerate(toks):
        b.forward(int(t), p)
    ka, va = a.kv_cache(); kb, vb = b.kv_cache()
    assert np.array_equal(ka[:, :6], kb[:, :6]) and np.array_equal(va[:, :6], vb[:, :6])
    assert np.array_equal(a.forward(11, 6), b.forward(11, 6))


@pytest.mark.parametrize("name,q_type", [("tiny-llama", 1), ("tiny-phi", 1), ("tiny-gemma", 1), ("tiny-llama", 2)])
def test_oracle_forward_agrees_with_an_independent_numpy_forward(ref, lf, name, q_type):
    """tests/numpy_forward.py restates Transformer::forward (src/transformer.rs:316-657) a second time, vectorised and
    with f64 float reductions (only the integer group dot products are shared with the oracle).  The two must agree far
    inside the 1e-3 logits tolerance on 2-layer models: a wrong head mapping, RoPE pairing, residual order or Gemma
    quirk in either restatement shows up as an O(1) difference."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from numpy_forward import NumpyForward
    buf = lf.write_synthetic(lf.model_args(name, q_type))
    cpu, nf = ref.RefTransformer(buf), NumpyForward(buf, lf, ref)
    for pos, tok in enumerate([3, 17, 5, 200, 9, 1]):
        a, b = cpu.forward(tok, pos).copy(), nf.forward(tok, pos)
        assert float(np.abs(a - b).max()) <= 1e-4, (name, pos)


@pytest.mark.parametrize("shards", [2, 4, 8])
def test_kshard_mode_is_the_rank_ordered_sum_of_column_slice_products(ref, shards):
    """The oracle's k-shard accumulation (the order of lmrs_b200's N-GPU peer exchange; NOT a reference feature) equals what it
    claims to be: matmul_q8 of every contiguous K range on its own (one GPU's partial), partials added in ascending order
    starting from range 0 -- and it differs from the unsharded order in the low bits, which is why the N-GPU tests need it."""
    rng = np.random.default_rng(shards)
    n, o, rows = 4096, 64, 3   # 32 groups: every shard count leaves several groups per shard
    xq = rng.integers(-127, 128, rows * n, dtype=np.int8); xs = (rng.uniform(0.5, 1.5, rows * n // 128) * 0.01).astype(np.float32)
    wq = rng.integers(-127, 128, o * n, dtype=np.int8); ws = (rng.uniform(0.5, 1.5, o * n // 128) * 0.002).astype(np.float32)
    got = ref.matmul_q8_kshards(xq, xs, wq, ws, rows, n, o, 128, shards)
    ns = n // shards
    total = None
    for r in range(shards):
        xq_r = np.ascontiguousarray(xq.reshape(rows, n)[:, r * ns:(r + 1) * ns]).reshape(-1)
        xs_r = np.ascontiguousarray(xs.reshape(rows, n // 128)[:, r * ns // 128:(r + 1) * ns // 128]).reshape(-1)
        wq_r = np.ascontiguousarray(wq.reshape(o, n)[:, r * ns:(r + 1) * ns]).reshape(-1)
        ws_r = np.ascontiguousarray(ws.reshape(o, n // 128)[:, r * ns // 128:(r + 1) * ns // 128]).reshape(-1)
        part = ref.matmul_q8(xq_r, xs_r, wq_r, ws_r, rows, ns, o, 128)
        total = part if total is None else (total + part).astype(np.float32)
    assert np.array_equal(got, total)
    assert np.array_equal(ref.matmul_q8_kshards(xq, xs, wq, ws, rows, n, o, 128, 1), ref.matmul_q8(xq, xs, wq, ws, rows, n, o, 128))
    assert not np.array_equal(got, ref.matmul_q8(xq, xs, wq, ws, rows, n, o, 128))
