"""CPU tests: the product's float-exact restatement headers (pbrt-v4_b200/csrc/*.h are host+device) compiled with g++ and
compared bit for bit with golden vectors written by the reference itself (oracle/ref_dump.cpp -> tests/golden/ref_vectors.bin).
This pins the BxDFs (conductor, rough dielectric, LayeredBxDF), the image infinite light and the participating-media code
(DDA majorant iterator, SampleT_maj, grid lookup, Henyey-Greenstein) without a GPU; the GPU tests then check that the device
build of the same headers renders the same images."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool", ["bxdf_host_check", "envmap_host_check", "media_host_check", "nee_host_check", "camera_host_check", "triintr_host_check", "material_host_check", "texture_host_check", "sampler_host_check", "rgbmedium_host_check", "quadric_host_check", "bilinear_host_check"])
def test_host_restatement_matches_reference_golden(tool, tmp_path):
    exe = str(tmp_path / tool)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-std=c++17", "-I/usr/local/cuda/include",
                           "-I" + os.path.join(ROOT, "pbrt-v4_b200", "csrc"), os.path.join(ROOT, "tests", "tools", tool + ".cpp"), "-o", exe])
    r = subprocess.run([exe], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_exact_libm_restatement_matches_host_libm(tmp_path):
    """csrc/libm_exact.h (glibc's float logf / sinf / cosf algorithms in double arithmetic, the device option BPT_EXACT_LIBM)
    against the host libm: every 13th float of the whole domain (the exhaustive run, stride 1, takes ~20 CPU-seconds and was
    0 mismatches in 2.1e9 + 2 x 2.2e9 arguments)."""
    exe = str(tmp_path / "libmcheck")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-mfma", "-fopenmp", os.path.join(ROOT, "tests", "tools", "libm_exact_check.c"),
                           "-I" + os.path.join(ROOT, "pbrt-v4_b200", "csrc"), "-o", exe, "-lm"])
    r = subprocess.run([exe, "13"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
