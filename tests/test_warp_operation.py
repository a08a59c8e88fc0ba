"""Row f1 at the operation: the reference's OWN src/Operations/WarpImages.cc is compiled in place against stand-ins for its
headers (host/build_host_tests.py::build_warp_operation) -- unmodified, and with patches/WarpImages.patch applied, which
hands the per-voxel body of a deformation-field warp (WarpImages.cc:418-456) to host/Warp_Images_Field.cc ->
dcma_b200_warp_to_grid -- and driven the way the dispatcher drives it (tests/native/warp_operation_check.cc).

CPU: the argument surface; the UNMODIFIED operation runs its own per-voxel path and satisfies the driver's checks (this
pins the driver and the stand-ins, not the product); the PATCHED operation without a device fails loudly (no fallback);
the patch is current and touches only what it says.  GPU (`-m gpu`): the patched operation satisfies the same checks --
same bounded voxels, same values as the statements of WarpImages.cc:424-452 evaluated on the host.
"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, has_gpu

sys.path.insert(0, os.path.join(ROOT, "dicomautomaton_b200", "host"))

# src/Operations/WarpImages.cc:64-175
EXPECTED_ARGS = [("ImageSelection", "last"), ("ReferenceImageSelection", "first"), ("TransformSelection", "last"),
                 ("NormalizedROILabelRegex", ".*"), ("ROILabelRegex", ".*"), ("ROISelection", "all"),
                 ("ContourOverlap", "ignore"), ("Inclusivity", "center"), ("Channel", "0")]


def _exe(patched):
    import build_host_tests
    p = build_host_tests.build_warp_operation(verbose=False, patched=patched)
    if not p:
        pytest.skip("reference operation source not available and no prebuilt binary")
    return p


def _run(exe, mode):
    return subprocess.run([exe, mode], capture_output=True, text=True, timeout=600, env=dict(os.environ, YLOG_VERBOSITY="0"))


@pytest.mark.parametrize("patched", [False, True])
def test_warp_operation_argument_surface(patched):
    r = _run(_exe(patched), "doc")
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "operation=WarpImages"
    assert [tuple(l.split("=", 1)) for l in lines[1:]] == EXPECTED_ARGS


def test_unmodified_operation_satisfies_the_driver_on_cpu():
    r = _run(_exe(False), "run")
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr


def test_patched_operation_has_no_cpu_fallback():
    if has_gpu():
        pytest.skip("a device is present")
    r = _run(_exe(True), "run")
    assert r.returncode == 3 and "no CUDA device" in r.stdout, r.stdout + r.stderr


def test_patched_operation_host_logic_on_cpu(tmp_path):
    """The PATCHED operation end to end on CPU: its objects are linked against tests/native/fake_warp_capi.cc -- the
    definition of dcma_b200_warp_to_grid evaluated on the host, malloc for the pinned allocator -- instead of the GPU
    library.  Checks everything host-side that the patch and host/Warp_Images_Field.cc add: which voxels are marked, the
    order of the mask, the marshalling of the three stacks and the write-back; the kernel itself is the GPU test's."""
    import build_host_tests
    if _exe(True) is None:
        pytest.skip("reference sources not available")
    objs = [os.path.join(build_host_tests.OUT, o) for o in ("ref_warp_operation_patched.o", "warp_operation_check.o", "Warp_Images_Field.o")]
    if not all(os.path.isfile(o) for o in objs):
        pytest.skip("objects of build_warp_operation not present (prebuilt executable only)")
    fake = os.path.join(str(tmp_path), "fake_warp_capi.o")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-c",
                           os.path.join(ROOT, "tests", "native", "fake_warp_capi.cc"), "-o", fake])
    exe = os.path.join(str(tmp_path), "warp_operation_check_cpu")
    subprocess.check_call(["g++", "-pthread", "-o", exe, *objs, fake])
    r = _run(exe, "run")
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr


def test_warp_patch_is_current_and_minimal():
    ref = os.environ.get("DCMA_REFERENCE_SRC", "/root/reference/src")
    src = os.path.join(ref, "Operations", "WarpImages.cc")
    if not os.path.isfile(src):
        pytest.skip("reference sources not available")
    patch = os.path.join(ROOT, "patches", "WarpImages.patch")
    text = open(patch).read()
    removed = [l[1:].strip() for l in text.splitlines() if l.startswith("-") and not l.startswith("---")]
    # the only line the patch takes away is the head of the operation's own mutator call, which it re-emits behind the
    # `!warped_on_gpu &&` guard
    assert removed == ["if(!edit_ia_ptr->imagecoll.Process_Images_Parallel( GroupIndividualImages,"]
    out = subprocess.run(["patch", "-s", "-o", "-", src, patch], capture_output=True, text=True)
    assert out.returncode == 0
    assert out.stdout.count("dcma_b200_host::warp_stack_with_field") == 1 and out.stdout.count("Process_Images_Parallel") == 2
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_warp_patch", os.path.join(ROOT, "tools", "make_warp_patch.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.patched(open(src).read()) == out.stdout


@pytest.mark.gpu
def test_patched_operation_warps_on_the_gpu():
    r = _run(_exe(True), "run")
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
