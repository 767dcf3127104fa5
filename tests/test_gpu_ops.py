"""GPU operator parity (through the C ABI) against the CPU oracle.

Bar: BIT-EXACT for every operator -- the kernels reproduce the reference's f32 operation order (ascending
group accumulation, 8-lane rmsnorm partial sums, serial softmax sum) and glibc's expf algorithm."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class QT:
    def __init__(self, q, s):
        self.q, self.s = q, s


def _rand_q8(rng, rows, n, scale):
    q = rng.integers(-127, 128, size=rows * n, dtype=np.int8)
    s = (rng.uniform(0.5, 1.5, rows * n // 128) * scale).astype(np.float32)
    return q, s


@pytest.mark.parametrize("n,o,rows", [(128, 4, 1), (256, 8, 1), (2048, 2048, 1), (2048, 512, 2), (8192, 2048, 1),
                                      (3072, 1024, 1), (2304, 64, 1), (3584, 4096, 1), (2048, 3072, 3),
                                      (14336, 256, 1), (384, 128256 // 32, 1)])
def test_matmul_q8_bit_exact(gpu_lib, ref, n, o, rows):
    rng = np.random.default_rng(n * 7 + o)
    xq, xs = _rand_q8(rng, rows, n, 0.01)
    wq, ws = _rand_q8(rng, o, n, 0.002)
    exp = ref.matmul_q8(xq, xs, wq, ws, rows, n, o, 128)
    got = np.full(rows * o, np.nan, np.float32)
    gpu_lib.functional.matmul_q8(got, QT(xq, xs), QT(wq, ws), n, o, 128)
    assert np.array_equal(got, exp), f"max abs diff {np.abs(got - exp).max()}"


@pytest.mark.parametrize("n,o,rows", [(128, 128, 8), (256, 128, 130), (2048, 2048, 128), (2048, 3072, 512), (2048, 8192, 300), (8192, 2048, 200),
                                      (3584, 256, 64), (384, 1152, 17)])
def test_matmul_q8_batched_tcgen05_gemm_bit_exact(gpu_lib, ref, n, o, rows):
    """rows >= 8 and 128-aligned shapes run the tcgen05/TMEM int8 GEMM (the fill_kv_cache kernel): same bits as the CPU path."""
    rng = np.random.default_rng(n * 13 + o + rows)
    xq, xs = _rand_q8(rng, rows, n, 0.01)
    wq, ws = _rand_q8(rng, o, n, 0.002)
    exp = ref.matmul_q8(xq, xs, wq, ws, rows, n, o, 128)
    got = np.full(rows * o, np.nan, np.float32)
    gpu_lib.functional.matmul_q8(got, QT(xq, xs), QT(wq, ws), n, o, 128)
    assert np.array_equal(got, exp), f"max abs diff {np.abs(got - exp).max()}, mismatches {(got != exp).sum()}"


@pytest.mark.parametrize("n,o,rows", [(128, 4, 1), (256, 8, 1), (3072, 3072, 1), (8192, 3072, 1), (3072, 1024, 2),
                                      (2304, 64, 1), (2048, 2048, 1)])
def test_matmul_q4_bit_exact(gpu_lib, ref, n, o, rows):
    rng = np.random.default_rng(n * 11 + o)
    xq = rng.integers(0, 256, size=rows * n // 2, dtype=np.uint8)
    xs = (-rng.uniform(0.5, 1.5, rows * n // 128) * 0.1).astype(np.float32)   # runtime scales are negative (max/-8)
    wq = rng.integers(0, 256, size=o * n // 2, dtype=np.uint8)
    ws = (-rng.uniform(0.5, 1.5, o * n // 128) * 0.02).astype(np.float32)
    exp = ref.matmul_q4(xq, xs, wq, ws, rows, n, o, 128)
    got = np.full(rows * o, np.nan, np.float32)
    gpu_lib.functional.matmul_q4(got, QT(xq, xs), QT(wq, ws), n, o, 128)
    assert np.array_equal(got, exp), f"max abs diff {np.abs(got - exp).max()}"


def test_matmul_rejects_bad_shapes(gpu_lib):
    out = np.zeros(4, np.float32)
    z8, zf = np.zeros(128, np.int8), np.zeros(1, np.float32)
    with pytest.raises(gpu_lib.LmrsError):
        gpu_lib.functional.matmul_q8(out, QT(z8, zf), QT(np.zeros(512, np.int8), np.zeros(4, np.float32)), 128, 4, 64)
    with pytest.raises(gpu_lib.LmrsError):   # o % 4 != 0: the reference silently drops rows (functional.rs:179); we refuse
        gpu_lib.functional.matmul_q8(np.zeros(3, np.float32), QT(z8, zf), QT(np.zeros(384, np.int8), np.zeros(3, np.float32)), 128, 3, 128)


@pytest.mark.parametrize("gs", [8, 32, 128])
def test_quantize_q8_bit_exact(gpu_lib, ref, gs):
    rng = np.random.default_rng(gs)
    x = (rng.standard_normal(gs * 257) * 3).astype(np.float32)
    x[gs:2 * gs] = 0.0                                  # all-zero group: scale 0, NaN -> 0 (quantization.rs:57-64)
    x[2 * gs:2 * gs + 8] = [127, -127, 63.5, -63.5, 0.5, -0.5, 2.5, 0]   # half-away-from-zero ties
    eq, es = ref.quantize_q8(x, gs)
    q = gpu_lib.quantization.QuantizedTensor(np.zeros(x.size, np.int8), np.zeros(x.size // gs, np.float32))
    gpu_lib.quantization.quantize(q, x, x.size, gs)
    assert np.array_equal(q.q, eq) and np.array_equal(q.s.view(np.uint32), es.view(np.uint32))


@pytest.mark.parametrize("gs", [8, 32, 128])
def test_quantize_q4_bit_exact(gpu_lib, ref, gs):
    rng = np.random.default_rng(gs + 1)
    x = (rng.standard_normal(gs * 129) * 2).astype(np.float32)
    x[gs:2 * gs] = 0.0
    x[2 * gs:2 * gs + 8] = [8, -8, 4, -4, 0, 1, -1, 7.5]
    eq, es = ref.quantize_q4(x, gs)
    q = gpu_lib.quantization.QuantizedTensor(np.zeros(x.size // 2, np.uint8), np.zeros(x.size // gs, np.float32))
    gpu_lib.quantization.quantize_q4(q, x, x.size, gs)
    assert np.array_equal(q.q, eq) and np.array_equal(q.s.view(np.uint32), es.view(np.uint32))


@pytest.mark.parametrize("size,unit", [(256, False), (2048, False), (3584, True), (4096, True)])
def test_rmsnorm(gpu_lib, ref, size, unit):
    rng = np.random.default_rng(size)
    x = rng.standard_normal(size).astype(np.float32) * 3
    w = (1 + 0.1 * rng.standard_normal(size)).astype(np.float32)
    exp = ref.rmsnorm(x, w, 1e-5, unit)
    got = np.zeros(size, np.float32)
    gpu_lib.functional.rmsnorm(got, x, w, size, 1e-5, unit)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


@pytest.mark.parametrize("n", [1, 7, 513, 8192])
def test_softmax(gpu_lib, ref, n):
    x = (np.random.default_rng(n).standard_normal(n) * 4).astype(np.float32)
    exp = ref.softmax(x)
    got = x.copy()
    gpu_lib.functional.softmax(got)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32))


@pytest.mark.parametrize("rows,n,o", [(1, 64, 8), (3, 588, 12), (2, 1024, 1024), (577, 1024, 64), (5, 7, 4), (4, 37, 20)])
@pytest.mark.parametrize("rest", [False, True])
def test_matmul_f32_and_matmul_rest_bit_exact(gpu_lib, ref, rows, n, o, rest):
    """`matmul` / `matmul_rest` (src/functional.rs:142-171, 252-280; the CLIP patch embedding has n = 588): one thread per
    output element running the function bodies that tests/test_f32_ops_host.py pins against the oracle on the host."""
    rng = np.random.default_rng(rows * 1000 + n + o)
    x = (rng.standard_normal(rows * n) * 3).astype(np.float32)
    w = (rng.standard_normal(o * n) * 0.7).astype(np.float32)
    x[::17] = -0.0
    exp = ref.matmul_f32(x, w, rows, n, o, rest=rest)
    got = np.full(rows * o, np.nan, np.float32)
    (gpu_lib.functional.matmul_rest if rest else gpu_lib.functional.matmul)(got, x, w, n, o)
    assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), f"mismatches {(got != exp).sum()}"


def test_matmul_f32_refuses_an_output_count_the_reference_would_truncate(gpu_lib):
    with pytest.raises(gpu_lib.LmrsError, match="multiple of 4"):
        gpu_lib.functional.matmul(np.zeros(6, np.float32), np.zeros(8, np.float32), np.zeros(48, np.float32), 8, 6)


@pytest.mark.parametrize("rows,size", [(1, 1024), (577, 1024), (3, 1032), (2, 1027), (1, 8)])
def test_layernorm_rows_bit_exact(gpu_lib, ref, rows, size):
    """src/functional.rs:80-114 for every token row of a vision-encoder activation (CLIP hidden size 1024; sizes with a
    size % 8 tail, which the reference leaves untouched)."""
    rng = np.random.default_rng(size + rows)
    x = rng.standard_normal((rows, size)).astype(np.float32) * 3 + 0.5
    w = (1 + 0.1 * rng.standard_normal(size)).astype(np.float32)
    b = (0.1 * rng.standard_normal(size)).astype(np.float32)
    exp = np.stack([ref.layernorm(x[r], w, b, 1e-5) for r in range(rows)])     # the oracle zero-fills the tail of its output
    got = np.zeros((rows, size), np.float32)
    gpu_lib.functional.layernorm(got, x, w, b, size, 1e-5)
    assert np.array_equal(got, exp), f"max abs diff {np.abs(got - exp).max()}"


@pytest.mark.parametrize("q_type,n,o,rows", [(1, 1024, 1024, 577), (1, 1024, 4096, 64), (1, 4096, 1024, 5), (1, 384, 100, 3),
                                               (2, 1024, 1024, 3)])
def test_resident_weights_matmul_bit_exact(gpu_lib, ref, q_type, n, o, rows):
    """weights uploaded once (lmrs_b200_weights_upload), applied to several batches of rows (tcgen05 GEMM for >= 8 Q8_0 rows
    of 128-aligned shapes, the matrix-vector kernel otherwise): same bits as matmul_q8 / matmul_q4 of the oracle."""
    rng = np.random.default_rng(n + o + rows)
    if q_type == 1:
        wq, ws = _rand_q8(rng, o, n, 0.002)
    else:
        wq = rng.integers(0, 256, size=o * n // 2, dtype=np.uint8)
        ws = (rng.uniform(0.5, 1.5, o * n // 128) * -0.003).astype(np.float32)
    W = gpu_lib.functional.ResidentWeights(QT(wq, ws), n, o, 128, q_type)
    for rr in (rows, 1, rows):
        if q_type == 1:
            xq, xs = _rand_q8(rng, rr, n, 0.01)
            exp = ref.matmul_q8(xq, xs, wq, ws, rr, n, o, 128)
        else:
            xq = rng.integers(0, 256, size=rr * n // 2, dtype=np.uint8)
            xs = (rng.uniform(0.5, 1.5, rr * n // 128) * -0.01).astype(np.float32)
            exp = ref.matmul_q4(xq, xs, wq, ws, rr, n, o, 128)
        got = np.full(rr * o, np.nan, np.float32)
        W.matmul(got, QT(xq, xs))
        assert np.array_equal(got, exp), f"rows {rr}: max abs diff {np.abs(got - exp).max()}"
    W.close()
