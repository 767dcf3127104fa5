"""The per-element functions of lm.rs_b200/csrc/f32_ops.cuh (`matmul`, `matmul_rest`, src/functional.rs:142-171,252-280)
are __host__ __device__: this test compiles the SAME function bodies for the host and checks them bit for bit against
the CPU oracle, so the arithmetic order of the CUDA operator is pinned without a GPU (the launch geometry is covered by
tests/test_gpu_ops.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "f32_ops.cuh"
extern "C" void host_matmul_f32(float* xout, const float* x, const float* w, int rows, int n, int o, int rest) {
    for (int r = 0; r < rows; r++)
        for (int i = 0; i < o; i++)
            xout[(size_t)r * o + i] = rest ? lmrs::matmul_rest_element(x + (size_t)r * n, x, w + (size_t)i * n, n)
                                           : lmrs::matmul_f32_element(x + (size_t)r * n, w + (size_t)i * n, n);
}
'''


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("f32host")
    src, so = tmp / "h.cpp", tmp / "h.so"
    src.write_text(SRC)
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "lm.rs_b200", "csrc"),
                           "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    L.host_matmul_f32.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4
    L.host_matmul_f32.restype = None
    return L


@pytest.mark.parametrize("rows,n,o", [(1, 64, 8), (3, 588, 12), (2, 1024, 16), (5, 7, 4), (2, 8, 4), (4, 37, 20)])
@pytest.mark.parametrize("rest", [False, True])
def test_f32_operator_bodies_match_the_oracle_bit_for_bit(host_lib, ref, rows, n, o, rest):
    rng = np.random.default_rng(rows * 1000 + n + o)
    x = (rng.standard_normal(rows * n) * 3).astype(np.float32)
    w = (rng.standard_normal(o * n) * 0.7).astype(np.float32)
    x[::17] = -0.0                                   # signed zeros: 0.0f + (-0.0f) must stay +0.0f as in the reference
    out = np.full(rows * o, 7.0, np.float32)
    host_lib.host_matmul_f32(out.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                             rows, n, o, int(rest))
    want = ref.matmul_f32(x, w, rows, n, o, rest=rest)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
