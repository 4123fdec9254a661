"""CPU tests: the product's float-exact restatement headers (pbrt-v4_b200/csrc/*.h are host+device) compiled with g++ and
compared bit for bit with golden vectors written by the reference itself (oracle/ref_dump.cpp -> tests/golden/ref_vectors.bin).
This pins the BxDFs (conductor, rough dielectric, LayeredBxDF), the image infinite light and the participating-media code
(DDA majorant iterator, SampleT_maj, grid lookup, Henyey-Greenstein) without a GPU; the GPU tests then check that the device
build of the same headers renders the same images."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool", ["bxdf_host_check", "envmap_host_check", "media_host_check"])
def test_host_restatement_matches_reference_golden(tool, tmp_path):
    exe = str(tmp_path / tool)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-std=c++17", "-I/usr/local/cuda/include",
                           "-I" + os.path.join(ROOT, "pbrt-v4_b200", "csrc"), os.path.join(ROOT, "tests", "tools", tool + ".cpp"), "-o", exe])
    r = subprocess.run([exe], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
