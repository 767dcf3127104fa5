"""The C-ABI shared library: loads on a CPU-only box, exports every symbol include/lmrs_b200.h declares, and
fails loudly (no CPU fallback) when asked to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "lmrs_b200.h")).read()
    return sorted(set(re.findall(r"\b(lmrs_b200_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(gpu_lib):
    lib = ctypes.CDLL(gpu_lib.SO_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lmrs_b200.h but not exported"
    assert sorted(gpu_lib.ABI_SYMBOLS) == names, "host mirror's symbol list is out of date"
    assert b"sm_100a" in gpu_lib.lib().lmrs_b200_version()


def test_sass_contains_the_blackwell_copy_engine_path():
    """The GEMV streams weights with cp.async.bulk (SASS UBLKCP) and contains IDP.4A integer dot products."""
    import shutil
    import subprocess
    so = os.path.join(ROOT, "lm.rs_b200", "lmrs_b200", "liblmrs_b200.so")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "IDP.4A" in sass and "SYNCS" in sass
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", so], capture_output=True, text=True).stdout


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present: the no-GPU failure path cannot be exercised")
def test_no_gpu_means_loud_failure_not_fallback(gpu_lib, lf):
    buf = lf.write_synthetic(lf.model_args("tiny-llama", 1))
    with pytest.raises(gpu_lib.LmrsError, match="no CUDA device|no CPU fallback"):
        gpu_lib.Transformer.new(buf)
    with pytest.raises(gpu_lib.LmrsError):
        gpu_lib.functional.softmax(np.zeros(4, np.float32))


def test_sass_contains_the_cluster_and_tensor_core_paths():
    """Decode attention pushes scores between the CTAs of a cluster with st.async (SASS STAS) and stages K/V with
    cp.async (LDGSTS); the prefill GEMM issues tcgen05 int8 MMAs (UTCIMMA) fed by 2-D TMA loads (UTMALDG)."""
    import shutil
    import subprocess
    so = os.path.join(ROOT, "lm.rs_b200", "lmrs_b200", "liblmrs_b200.so")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True).stdout
    for mnemonic in ("STAS", "LDGSTS", "UCGABAR_ARV", "UTCIMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_compiled_caller_links_against_the_c_abi_and_fails_loudly_without_a_gpu(lf, tmp_path):
    """lm.rs_b200/examples/generate.cpp: the chat.rs greedy loop written against include/lmrs_b200.h only."""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "lm.rs_b200"), "-s", "examples"])
    exe = os.path.join(ROOT, "lm.rs_b200", "examples", "generate")
    path = tmp_path / "tiny.lmrs"
    lf.write_synthetic(lf.model_args("tiny-llama", 1)).tofile(path)
    r = subprocess.run([exe, str(path), "4", "3", "5"], capture_output=True, text=True)
    if os.path.exists("/dev/nvidia0"):
        assert r.returncode == 0 and len(r.stdout.split()) == 4, r.stderr
    else:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_oracle_is_not_linked_into_the_product():
    import subprocess
    so = os.path.join(ROOT, "lm.rs_b200", "lmrs_b200", "liblmrs_b200.so")
    out = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert "lmrs_ref_" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "lm.rs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert "lmrs_ref" not in src.replace("lmrs_ref.c header", ""), f"{f} references the oracle"
