"""The C++ host drop-in (dicomautomaton_b200/host/Alignment_Demons.{h,cc}) under the REFERENCE'S OWN unit tests:
/root/reference/src/Alignment_Demons_Tests.cc is compiled in place, unmodified, against our AlignViaDemons and a minimal
Ygor stand-in (host/build_host_tests.py), and linked to libdcma_b200.so.  The binary is built in the container where
/root/reference exists and travels to the GPU box."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dicomautomaton_b200", "host"))


@pytest.fixture(scope="module")
def exe():
    import build_host_tests
    p = build_host_tests.build(verbose=False)
    if not p:
        pytest.skip("reference test source not available and no prebuilt binary")
    return p


def _run(exe, *args, mode=None):
    env = dict(os.environ, YLOG_VERBOSITY="0")
    if mode:
        env["DCMA_B200_FIELD_MODE"] = mode
    r = subprocess.run([exe, *args], capture_output=True, text=True, env=env, timeout=600)
    m = re.search(r"test cases:\s+(\d+)\s+\|\s+(\d+) passed\s+\|\s+(\d+) failed", r.stdout)
    assert m, r.stdout[-2000:]
    return int(m.group(1)), int(m.group(2)), int(m.group(3)), r.stdout


def _has_gpu():
    try:
        from dicomautomaton_b200 import capi
        return capi.lib().dcma_b200_device_count() > 0
    except Exception:
        return False


def test_dropin_host_logic_against_a_stand_in_library(tmp_path):
    """Host side of AlignViaDemons on CPU: host/Alignment_Demons.cc linked against tests/native/fake_register_capi.cc (a
    library that returns a KNOWN field, in chunks that straddle slices) instead of the GPU library.  Checks what the C call
    is given (volumes, grids, which entry point) and that every voxel of the returned deformation_field is the one the
    library wrote -- the chunk-by-chunk copy from the notification callback (tests/native/dropin_chunk_check.cc)."""
    host = os.path.join(ROOT, "dicomautomaton_b200", "host")
    inc = ["-I", os.path.join(host, "shim", "dcma_min"), "-I", os.path.join(host, "shim", "ygor_min"), "-I", host,
           "-I", os.path.join(ROOT, "include")]
    exe = os.path.join(str(tmp_path), "dropin_chunk_check")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-Wno-unused-function", *inc,
                           os.path.join(ROOT, "tests", "native", "dropin_chunk_check.cc"),
                           os.path.join(ROOT, "tests", "native", "fake_register_capi.cc"),
                           os.path.join(host, "Alignment_Demons.cc"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, YLOG_VERBOSITY="0"))
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr


def test_reference_host_side_cases_pass_without_gpu(exe):
    """histogram_match (exact values 10, 17.5, 25, 32.5) and the empty-input contract are host code and need no device."""
    total, passed, failed, out = _run(exe, "-tc=histogram_match*,AlignViaDemons handles empty inputs")
    assert total == 3 and failed == 0, out[-2000:]


def test_gpu_backed_helpers_fail_loudly_without_a_device(exe):
    """resample_image_to_reference_grid and warp_image_with_field run on the GPU for regular stacks (the reference runs
    them as serial host loops, cc:709-755 and cc:896-941).  There is no CPU fallback: on a box without a CUDA device they
    must fail with the library's message, not quietly compute something on the host."""
    if _has_gpu():
        pytest.skip("a CUDA device is present: the cases run for real in the gpu-marked test below")
    total, passed, failed, out = _run(exe, "-tc=resample_image_to_reference_grid*,warp_image_with_field*")
    assert total == 3 and failed == 3, out[-2000:]
    assert out.count("no CUDA device is available (dcma_b200 has no CPU fallback)") >= 3, out[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp64", "fp64_strict", "fp32"])
def test_reference_unit_tests_pass_against_the_gpu_dropin(exe, mode):
    """All 14 doctest cases of the reference (src/Alignment_Demons_Tests.cc) against the CUDA path."""
    total, passed, failed, out = _run(exe, mode=mode)
    assert total == 14 and failed == 0, out[-3000:]
