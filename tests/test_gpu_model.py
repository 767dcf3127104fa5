"""Model-level parity on the GPU: Transformer.{new, forward, get_embeddings, fill_kv_cache} through the C ABI
against the CPU oracle on identical synthetic LMRS files and prompts.

Bar: BASELINE.json north_star asks for logits within 1e-3 max-abs of the CPU path.  Because the path re-quantizes
activations (discontinuous rounding), that is only reliably achievable by reproducing every f32 operation bit for
bit (see lm.rs_b200/csrc/exact_math.cuh), so LLAMA/PHI logits, KV rows and residual streams are compared for
EXACT equality; GEMMA goes through an f64 tanh (CUDA libdevice vs glibc) and keeps the 1e-3 tolerance."""
import os
import numpy as np
import pytest

from conftest import prompt_tokens

pytestmark = pytest.mark.gpu
TOL = 1e-3

TINY = [("tiny-llama", 1), ("tiny-llama", 2), ("tiny-gemma", 1), ("tiny-gemma", 2), ("tiny-gemma-narrow", 1),
        ("tiny-phi", 1), ("tiny-phi", 2), ("small-llama", 1)]


@pytest.mark.parametrize("name,q_type", TINY)
def test_forward_logits_match_oracle(gpu_lib, ref, synth, name, q_type):
    buf = synth(name, q_type)
    cpu = ref.RefTransformer(buf)
    gpu, end = gpu_lib.Transformer.new(buf)
    assert end == cpu.end_offset == buf.size
    assert bytes(gpu.args) == bytes(cpu.args)
    toks = prompt_tokens(gpu.args.vocab_size, 24)
    exact = gpu.args.model_type != 0
    worst = 0.0
    for pos, t in enumerate(toks):
        lg = gpu.forward(int(t), pos)
        le = cpu.forward(int(t), pos)
        assert np.isfinite(lg).all()
        worst = max(worst, float(np.abs(lg - le).max()))
        if exact:
            assert np.array_equal(lg, le), f"pos {pos}: logits differ (max abs {np.abs(lg - le).max()})"
    assert worst <= TOL, f"max-abs logits diff {worst}"
    kc, vc = cpu.kv_cache()
    for l in range(gpu.args.n_layers):
        k, v = gpu.read_kv(l, 0, len(toks))
        if exact:
            assert np.array_equal(k, kc[l, :len(toks)]) and np.array_equal(v, vc[l, :len(toks)])
        else:
            np.testing.assert_allclose(k, kc[l, :len(toks)], atol=TOL, rtol=0)
            np.testing.assert_allclose(v, vc[l, :len(toks)], atol=TOL, rtol=0)
    gpu.close(); cpu.close()


@pytest.mark.parametrize("name", ["ref_llama_q8", "ref_gemma_q4", "ref_phi_q8"])
def test_files_written_by_the_reference_exporter(gpu_lib, ref, name):
    """tests/golden/*.lmrs come from /root/reference/export.py itself (tests/golden/make_golden.py)."""
    buf = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", name + ".lmrs"), dtype=np.uint8)
    cpu = ref.RefTransformer(buf)
    gpu, end = gpu_lib.Transformer.new(buf)
    assert end == buf.size
    for pos, t in enumerate(prompt_tokens(gpu.args.vocab_size, 12, seed=5)):
        lg, le = gpu.forward(int(t), pos), cpu.forward(int(t), pos)
        if gpu.args.model_type != 0:
            assert np.array_equal(lg, le)
        assert float(np.abs(lg - le).max()) <= TOL


@pytest.mark.parametrize("name,q_type", [("tiny-llama", 1), ("tiny-gemma", 1), ("tiny-phi", 2)])
def test_get_embeddings_bit_exact(gpu_lib, ref, synth, name, q_type):
    buf = synth(name, q_type)
    cpu = ref.RefTransformer(buf)
    gpu, _ = gpu_lib.Transformer.new(buf)
    toks = prompt_tokens(gpu.args.vocab_size, 9)
    assert np.array_equal(gpu.get_embeddings(toks), cpu.get_embeddings(toks))
    assert gpu.get_embeddings(np.zeros(0, np.uint32)).size == 0
    with pytest.raises(gpu_lib.LmrsError):
        gpu.get_embeddings([gpu.args.vocab_size])


@pytest.mark.parametrize("name,q_type", [("tiny-llama", 1), ("tiny-gemma", 1), ("tiny-phi", 1), ("tiny-llama", 2)])
def test_fill_kv_cache_then_decode(gpu_lib, ref, synth, name, q_type):
    """The multimodal path of the bins: get_embeddings ++ features -> fill_kv_cache -> forward (chat.rs:110-119)."""
    buf = synth(name, q_type)
    cpu = ref.RefTransformer(buf)
    gpu, _ = gpu_lib.Transformer.new(buf)
    # 3 decode steps first so the batch starts at pos 3 (exercises pos offset and the Gemma mask_base quirk)
    for pos, t in enumerate([5, 6, 7]):
        gpu.forward(t, pos); cpu.forward(t, pos)
    toks = prompt_tokens(gpu.args.vocab_size, 17, seed=3)
    eg, ec = gpu.get_embeddings(toks), cpu.get_embeddings(toks)
    pg, pc = gpu.fill_kv_cache(eg, 3), cpu.fill_kv_cache(ec, 3)
    assert pg == pc == 20
    lg, le = gpu.forward(11, pg), cpu.forward(11, pc)
    if gpu.args.model_type != 0:
        assert np.array_equal(eg, ec)                             # residual stream returned in place
        assert np.array_equal(lg, le)
    np.testing.assert_allclose(eg, ec, atol=TOL, rtol=1e-4)
    assert float(np.abs(lg - le).max()) <= TOL


def test_errors_where_the_reference_panics(gpu_lib, synth):
    buf = synth("tiny-llama", 1)
    bad = buf.copy(); bad[0] = 0
    with pytest.raises(gpu_lib.LmrsError, match="lm.rs format"):
        gpu_lib.Transformer.new(bad)
    with pytest.raises(gpu_lib.LmrsError, match="truncated"):
        gpu_lib.Transformer.new(buf[: buf.size // 2])
    gpu, _ = gpu_lib.Transformer.new(buf)
    with pytest.raises(gpu_lib.LmrsError):
        gpu.forward(gpu.args.vocab_size, 0)
    with pytest.raises(gpu_lib.LmrsError):
        gpu.forward(0, gpu.args.seq_len)


def test_seq_len_clamped_to_8192(gpu_lib, lf):
    a = lf.model_args("tiny-llama", 1, seq_len=131072)
    gpu, _ = gpu_lib.Transformer.new(lf.write_synthetic(a))
    assert gpu.args.seq_len == 8192      # src/transformer.rs:158-160


def test_two_handles_are_independent(gpu_lib, ref, synth):
    """backend.rs creates one Transformer per connection (src/bin/backend.rs:87-110)."""
    buf = synth("tiny-llama", 1)
    a, _ = gpu_lib.Transformer.new(buf)
    b, _ = gpu_lib.Transformer.new(buf)
    cpu = ref.RefTransformer(buf)
    la0 = a.forward(3, 0).copy()
    lb0 = b.forward(9, 0).copy()
    la1 = a.forward(4, 1).copy()
    e0 = cpu.forward(3, 0).copy(); e1 = cpu.forward(4, 1).copy()
    assert np.abs(la0 - e0).max() <= TOL and np.abs(la1 - e1).max() <= TOL
    assert not np.array_equal(la0, lb0)


def test_llama_1b_q8_full_shape(gpu_lib, ref, lf):
    """BASELINE config 2 shape (Llama-3.2-1B Q8_0, synthetic weights): logits within 1e-3 at a few positions, incl.
    after a 96-embedding fill_kv_cache."""
    buf = lf.write_synthetic(lf.model_args("llama-3.2-1b", 1), seed=0, mode="fast")
    cpu = ref.RefTransformer(buf)
    gpu, _ = gpu_lib.Transformer.new(buf)
    toks = prompt_tokens(gpu.args.vocab_size, 100, seed=1)
    eg, ec = gpu.get_embeddings(toks[:96]), cpu.get_embeddings(toks[:96])
    assert np.array_equal(eg, ec)
    pg, pc = gpu.fill_kv_cache(eg, 0), cpu.fill_kv_cache(ec, 0)
    assert pg == pc == 96
    assert np.array_equal(eg, ec)
    for i, t in enumerate(toks[96:]):
        lg, le = gpu.forward(int(t), 96 + i), cpu.forward(int(t), 96 + i)
        assert np.array_equal(lg, le), f"pos {96 + i}: max abs {np.abs(lg - le).max()}"


@pytest.mark.parametrize("name,q_type,layers", [("llama-3.2-3b", 2, 3), ("phi-3.5-mini", 1, 2), ("gemma-2-9b", 1, 2), ("llama-3.2-3b", 1, 2)])
def test_full_shapes_of_the_other_baseline_configs(gpu_lib, ref, lf, name, q_type, layers):
    """BASELINE configs 3-5 at their real dim / hidden / heads / head_size / vocabulary (Llama-3.2-3B Q4_0 and Q8_0,
    Phi-3.5-mini Q8_0 with its separate lm_head, Gemma-2-9B Q8_0 with dim 3584 = 28 groups and att_dim > dim), truncated to
    a few blocks so the CPU oracle finishes in seconds: a 40-embedding fill_kv_cache (tcgen05 GEMM path for Q8_0, per-token
    chain for Q4_0), then decode steps.  LLAMA / PHI bit for bit, GEMMA (f64 tanh in two libms) within 1e-3."""
    a = lf.model_args(name, q_type, n_layers=layers)
    buf = lf.write_synthetic(a, seed=3, mode="fast")
    cpu = ref.RefTransformer(buf)
    gpu, _ = gpu_lib.Transformer.new(buf)
    exact = a.model_type != 0
    toks = prompt_tokens(a.vocab_size, 44, seed=5)
    eg, ec = gpu.get_embeddings(toks[:40]), cpu.get_embeddings(toks[:40])
    assert np.array_equal(eg, ec)
    assert gpu.fill_kv_cache(eg, 0) == cpu.fill_kv_cache(ec, 0) == 40
    assert np.array_equal(eg, ec) if exact else float(np.abs(eg - ec).max()) <= TOL, f"{name}: residual stream"
    for i, t in enumerate(toks[40:]):
        lg, le = gpu.forward(int(t), 40 + i), cpu.forward(int(t), 40 + i)
        assert np.array_equal(lg, le) if exact else float(np.abs(lg - le).max()) <= TOL, f"{name} pos {40 + i}: max abs {np.abs(lg - le).max()}"
    gpu.close(); cpu.close()


def test_fill_kv_cache_edge_cases(gpu_lib, ref, lf):
    """empty batch, batch crossing the shared-memory score capacity (serial fallback + HBM score scratch), out of range."""
    a = lf.model_args("tiny-llama", 1, seq_len=4096)
    buf = lf.write_synthetic(a)
    cpu = ref.RefTransformer(buf)
    gpu, _ = gpu_lib.Transformer.new(buf)
    assert gpu.fill_kv_cache(np.zeros(0, np.float32), 7) == 7                      # no embeddings: position unchanged
    toks = prompt_tokens(a.vocab_size, 24, seed=9)
    eg, ec = gpu.get_embeddings(toks), cpu.get_embeddings(toks)
    assert gpu.fill_kv_cache(eg, 2040) == cpu.fill_kv_cache(ec, 2040) == 2064      # crosses 2048: per-token chain, scores in HBM scratch
    assert np.array_equal(eg, ec)
    assert np.array_equal(gpu.forward(5, 2064), cpu.forward(5, 2064))
    with pytest.raises(gpu_lib.LmrsError):
        gpu.fill_kv_cache(gpu.get_embeddings(toks), 4090)                           # beyond seq_len


@pytest.mark.parametrize("name", ["tiny-llama", "tiny-gemma", "tiny-phi"])
def test_decode_attention_position_buckets(gpu_lib, ref, lf, name):
    """The cluster attention kernel has one graph variant per position bucket and hands over to the single-CTA kernel
    beyond the largest one: sweep positions around every hand-over (and the first few, where most CTAs own no rows)."""
    a = lf.model_args(name, 1, seq_len=4096)
    buf = lf.write_synthetic(a)
    cpu = ref.RefTransformer(buf)
    gpu, _ = gpu_lib.Transformer.new(buf)
    edges = [444, 572, 636, 700, 892, 1020, 1372, 1788, 2044]
    positions = list(range(0, 40)) + [p + d for p in edges for d in range(0, 10)]
    for i, pos in enumerate(positions):
        t = (7 * i + 3) % a.vocab_size
        lg, le = gpu.forward(t, pos), cpu.forward(t, pos)
        if a.model_type != 0:
            assert np.array_equal(lg, le), f"{name} pos {pos}: max abs {np.abs(lg - le).max()}"
        assert float(np.abs(lg - le).max()) <= TOL, f"{name} pos {pos}"


@pytest.mark.parametrize("name", ["tiny-llama", "tiny-gemma", "tiny-phi", "small-llama"])
def test_batched_prefill_ragged_multi_tile(gpu_lib, ref, lf, name):
    """fill_kv_cache over several token tiles of the fused attention kernel, a ragged second batch at pos > 0 whose tiles
    straddle the 32-position K/V tiles, then decode: residual stream, KV rows and logits against the oracle."""
    a = lf.model_args(name, 1, seq_len=4096)
    buf = lf.write_synthetic(a)
    cpu = ref.RefTransformer(buf)
    gpu, _ = gpu_lib.Transformer.new(buf)
    toks = prompt_tokens(a.vocab_size, 300 + 77 + 2, seed=11)
    exact = a.model_type != 0
    pos = 0
    for n in (300, 77):
        eg, ec = gpu.get_embeddings(toks[pos:pos + n]), cpu.get_embeddings(toks[pos:pos + n])
        assert gpu.fill_kv_cache(eg, pos) == cpu.fill_kv_cache(ec, pos) == pos + n
        assert np.array_equal(eg, ec) if exact else float(np.abs(eg - ec).max()) <= TOL, f"{name}: residual stream after {pos}+{n}"
        pos += n
    kg, vg = gpu.read_kv(a.n_layers - 1, 0, pos)
    kc, vc = cpu.kv_cache()
    if exact:
        assert np.array_equal(kg, kc[a.n_layers - 1, :pos]) and np.array_equal(vg, vc[a.n_layers - 1, :pos])
    for i in range(2):
        lg, le = gpu.forward(int(toks[pos + i]), pos + i), cpu.forward(int(toks[pos + i]), pos + i)
        assert np.array_equal(lg, le) if exact else float(np.abs(lg - le).max()) <= TOL, f"{name}: logits at {pos + i}"


def test_batches_with_att_dim_below_dim(gpu_lib, ref, lf):
    """att_dim < dim (Gemma-2-2B-like): the reference's chunking (src/transformer.rs:501-503) is well defined while
    n * (dim - att_dim) < att_dim and panics beyond; same boundary here (3 rows run, 4 are refused)."""
    buf = lf.write_synthetic(lf.model_args("tiny-gemma-narrow2", 1))
    cpu = ref.RefTransformer(buf)
    gpu, _ = gpu_lib.Transformer.new(buf)
    toks = prompt_tokens(gpu.args.vocab_size, 5, seed=2)
    eg, ec = gpu.get_embeddings(toks[:3]), cpu.get_embeddings(toks[:3])
    assert gpu.fill_kv_cache(eg, 0) == cpu.fill_kv_cache(ec, 0) == 3
    assert float(np.abs(eg - ec).max()) <= TOL
    assert float(np.abs(gpu.forward(int(toks[3]), 3) - cpu.forward(int(toks[3]), 3)).max()) <= TOL
    with pytest.raises(gpu_lib.LmrsError):
        gpu.fill_kv_cache(gpu.get_embeddings(toks[:4]), 4)


def test_q4_prefill_uses_the_per_token_chain(gpu_lib, ref, synth):
    """matmul_q4 with sl > 1 is undefined in the reference (src/functional.rs:224): row-wise semantics, no GEMM path."""
    buf = synth("tiny-llama", 2)
    cpu = ref.RefTransformer(buf)
    gpu, _ = gpu_lib.Transformer.new(buf)
    toks = prompt_tokens(gpu.args.vocab_size, 20, seed=4)
    eg, ec = gpu.get_embeddings(toks), cpu.get_embeddings(toks)
    assert gpu.fill_kv_cache(eg, 0) == cpu.fill_kv_cache(ec, 0) == 20
    assert np.array_equal(eg, ec)
    assert np.array_equal(gpu.forward(3, 20), cpu.forward(3, 20))


@pytest.mark.parametrize("env", [{"LMRS_B200_LL": "1"}, {"LMRS_B200_LL": "1", "LMRS_B200_GEMV_CFG": "0"}, {"LMRS_B200_ATT_SPLIT": "1"}, {"LMRS_B200_GRAPH": "0", "LMRS_B200_PDL": "0"},
                                 {"LMRS_B200_GRAPH": "0"}, {"LMRS_B200_GEMM": "0"}, {"LMRS_B200_GEMV_CFG": "1"}, {"LMRS_B200_GEMV_CFG": "4"},
                                 {"LMRS_B200_ATT_CLUSTER": "0"}, {"LMRS_B200_ATT_CLUSTER": "4"}, {"LMRS_B200_ATT_GROUPS": "1"},
                                 {"LMRS_B200_PF_ATTN": "2"}, {"LMRS_B200_PF_ATTN": "0"}, {"LMRS_B200_PFA_QUADS": "8"}, {"LMRS_B200_L2PF": "1"}, {"LMRS_B200_L2PF": "2"}, {"LMRS_B200_GEMM_BN": "128"},
                                 {"LMRS_B200_GEMM_BN": "64"}])
def test_alternative_execution_modes_stay_bit_exact(env):
    """fence-free LL exchange between co-resident kernels instead of kernel-boundary hand-overs (16- and 8-warp rings) /
    GPU-wide score kernel / no graph, no PDL / no graph / serial prefill / other ring geometries / single-CTA attention /
    clusters of 4 / ungrouped heads / two-kernel and per-row prefill attention / L2 prefetch modes / forced GEMM tile widths:
    same bits as the default path."""
    import subprocess
    import sys
    code = r'''
import sys, os, numpy as np
sys.path[:0] = [os.path.join(os.getcwd(), "lm.rs_b200"), os.path.join(os.getcwd(), "oracle")]
import lmrs_b200, lmrs_ref
from lmrs_b200 import lmrs_file as lf
for name, q in (("tiny-llama", 1), ("tiny-phi", 2), ("small-llama", 1)):
    buf = lf.write_synthetic(lf.model_args(name, q))
    g, _ = lmrs_b200.Transformer.new(buf); c = lmrs_ref.RefTransformer(buf)
    toks = np.random.default_rng(2).integers(0, g.args.vocab_size, 30).astype(np.uint32)
    eg, ec = g.get_embeddings(toks[:18]), c.get_embeddings(toks[:18])
    assert g.fill_kv_cache(eg, 0) == c.fill_kv_cache(ec, 0) == 18 and np.array_equal(eg, ec)
    for i, t in enumerate(toks[18:]):
        assert np.array_equal(g.forward(int(t), 18 + i), c.forward(int(t), 18 + i)), (name, i)
print("ok")
'''
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
