"""expf_glibc (lm.rs_b200/csrc/exact_math.cuh) must return the same bits as the host libm's expf -- the
function Rust's f32::exp calls in softmax and SiLU (src/functional.rs:133, src/transformer.rs:617)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "exact_math.cuh"
#include <math.h>
#include <stdio.h>
int main() {
    unsigned long long bad = 0, n = 0;
    for (uint32_t b = 0; b < 0x7f800000u; b += 101) {
        for (int sg = 0; sg < 2; sg++) {
            uint32_t u = b | (sg ? 0x80000000u : 0); float x; memcpy(&x, &u, 4);
            if (!(fabsf(x) < 200.0f)) continue;
            float a = expf(x), c = lmrs::expf_glibc(x); uint32_t ua, uc; memcpy(&ua, &a, 4); memcpy(&uc, &c, 4);
            n++; if (ua != uc) bad++;
        }
    }
    float specials[] = {-INFINITY, INFINITY, 88.0f, 88.7f, 88.8f, 100.0f, -87.0f, -103.0f, -103.9f, -104.0f, -200.0f, 0.0f, -0.0f};
    for (float x : specials) { float a = expf(x), c = lmrs::expf_glibc(x); uint32_t ua, uc; memcpy(&ua, &a, 4); memcpy(&uc, &c, 4); n++; if (ua != uc) bad++; }
    printf("%llu %llu\n", n, bad);
    return 0;
}
'''


def test_expf_restatement_matches_host_libm_bit_for_bit():
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "t.cpp")
        open(src, "w").write(SRC)
        exe = os.path.join(tmp, "t")
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-ffp-contract=off", "-I", os.path.join(ROOT, "lm.rs_b200", "csrc"), "-o", exe, src, "-lm"])
        n, bad = map(int, subprocess.check_output([exe]).split())
    assert n > 20_000_000 and bad == 0
