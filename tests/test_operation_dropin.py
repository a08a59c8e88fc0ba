"""The drop-in one level up, at the operation: the reference's OWN src/Operations/RegisterImagesDemons.cc is compiled in
place, unmodified, against our Alignment_Demons.h / AlignViaDemons and stand-ins for Structs.h / Regex_Selectors.h
(host/build_host_tests.py: build_operation), and driven the way the dispatcher drives it -- arguments as strings,
defaults from its OperationDoc (tests/native/operation_check.cc).

Checked here: the argument surface (16 arguments, names and defaults) and the error contract without a
device (AlignViaDemons warns and returns nullopt, the operation throws "Demons registration failed",
RegisterImagesDemons.cc:337-341).  With a device (`-m gpu`) the same binary runs the registration and checks what the
operation leaves in the Drover (RegisterImagesDemons.cc:346-381: one Transform3 holding a deformation_field, its
metadata, the KeepDeformedImages array) and that the warped moving image is close to the fixed one.
"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT, has_gpu

sys.path.insert(0, os.path.join(ROOT, "dicomautomaton_b200", "host"))

# SURVEY.md section 8(b), "Argument surface (names, defaults)"; RegisterImagesDemons.cc:84-230
EXPECTED_ARGS = [("MovingImageSelection", "last"), ("FixedImageSelection", "first"), ("MaxIterations", "100"),
                 ("ConvergenceThreshold", "0.001"), ("DeformationFieldSmoothingSigma", "1.0"),
                 ("UpdateFieldSmoothingSigma", "0.5"), ("UseDiffeomorphic", "false"), ("UseHistogramMatching", "false"),
                 ("HistogramBins", "256"), ("HistogramOutlierFraction", "0.01"), ("NormalizationFactor", "1.0"),
                 ("MaxUpdateMagnitude", "2.0"), ("Verbosity", "1"), ("TransformName", "demons_registration"),
                 ("Metadata", ""), ("KeepDeformedImages", "false")]


@pytest.fixture(scope="module")
def exe():
    import build_host_tests
    p = build_host_tests.build_operation(verbose=False)
    if not p:
        pytest.skip("reference operation source not available and no prebuilt binary")
    return p


def test_operation_argument_surface_is_the_references(exe):
    r = subprocess.run([exe, "doc"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "operation=RegisterImagesDemons"
    got = [tuple(l.split("=", 1)) for l in lines[1:]]
    assert got == EXPECTED_ARGS
    # the parameter struct behind them has the same defaults (Alignment_Demons.h:20-61 <-> dcma_b200_demons_params_default)
    from dicomautomaton_b200 import capi
    p = capi.default_params()
    d = dict(got)
    assert p.max_iterations == int(d["MaxIterations"]) and p.convergence_threshold == float(d["ConvergenceThreshold"])
    assert p.deformation_field_smoothing_sigma == float(d["DeformationFieldSmoothingSigma"])
    assert p.update_field_smoothing_sigma == float(d["UpdateFieldSmoothingSigma"])
    assert p.histogram_bins == int(d["HistogramBins"]) and p.histogram_outlier_fraction == float(d["HistogramOutlierFraction"])
    assert p.normalization_factor == float(d["NormalizationFactor"]) and p.max_update_magnitude == float(d["MaxUpdateMagnitude"])
    assert not p.use_diffeomorphic and not p.use_histogram_matching


def test_operation_fails_like_the_reference_without_a_device(exe):
    if has_gpu():
        pytest.skip("a device is present")
    r = subprocess.run([exe, "run"], capture_output=True, text=True, timeout=600, env=dict(os.environ, YLOG_VERBOSITY="0"))
    # no CPU fallback: nullopt -> std::runtime_error("Demons registration failed") -> the dispatcher returns false
    assert r.returncode == 3, r.stdout + r.stderr
    assert "THROW: Demons registration failed" in r.stdout
    assert "no CUDA device is available" in r.stderr


@pytest.mark.gpu
def test_operation_registers_through_the_gpu_dropin(exe):
    """On a B200 (profiles/r1_operation_check.log): MSE before 91.1138 after 0.4843, every check passes."""
    r = subprocess.run([exe, "run"], capture_output=True, text=True, timeout=600, env=dict(os.environ, YLOG_VERBOSITY="0"))
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr


# ---- the operation with patches/RegisterImagesDemons.patch applied: the additive arguments of the drop-in ----------
ADDED_ARGS = [("FieldMode", "fp64"), ("PyramidLevels", "1"), ("PyramidIterations", ""), ("UseSymmetricForce", "false"),
              ("Devices", "1")]


@pytest.fixture(scope="module")
def exe_patched():
    import build_host_tests
    p = build_host_tests.build_operation(verbose=False, patched=True)
    if not p:
        pytest.skip("reference operation source not available and no prebuilt binary")
    return p


def test_patch_adds_the_extension_arguments_with_reference_preserving_defaults(exe_patched):
    """patches/RegisterImagesDemons.patch applies to the reference's file, the result compiles against the drop-in, the 16
    reference arguments are untouched and the five additive ones follow `Verbosity` with defaults that reproduce the
    reference (new args must be additive with reference-preserving defaults)."""
    r = subprocess.run([exe_patched, "doc"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    got = [tuple(l.split("=", 1)) for l in r.stdout.strip().splitlines()[1:]]
    i = [n for n, _ in got].index("Verbosity")
    assert got[i + 1:i + 1 + len(ADDED_ARGS)] == ADDED_ARGS
    assert [g for g in got if g not in ADDED_ARGS] == EXPECTED_ARGS


def test_patched_operation_fails_like_the_reference_without_a_device(exe_patched):
    if has_gpu():
        pytest.skip("a device is present")
    r = subprocess.run([exe_patched, "run3"], capture_output=True, text=True, timeout=600, env=dict(os.environ, YLOG_VERBOSITY="0"))
    assert r.returncode == 3 and "THROW: Demons registration failed" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_config3_style_run_through_the_operation_surface(exe_patched):
    """PyramidLevels=3, UseSymmetricForce=true, KeepDeformedImages=true through RegisterImagesDemons() itself: the
    multi-resolution symmetric-force registration (BASELINE config 3's kind of run) is reachable from the dispatcher."""
    r = subprocess.run([exe_patched, "run3"], capture_output=True, text=True, timeout=600, env=dict(os.environ, YLOG_VERBOSITY="0"))
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([exe_patched, "run"], capture_output=True, text=True, timeout=600, env=dict(os.environ, YLOG_VERBOSITY="0"))
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
