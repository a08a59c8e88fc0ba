"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares, its POD
structs have the layout the ctypes binding assumes, defaults equal the reference's, and -- with no GPU -- every compute
entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dcma_b200", "demons_c_api.h")


def _declared_symbols():
    txt = open(HEADER).read()
    return sorted(set(re.findall(r"DCMA_B200_API[^;(]*?\b(dcma_b200_\w+)\s*\(", txt)))


def test_header_declares_what_the_binding_binds():
    from dicomautomaton_b200 import capi
    assert _declared_symbols() == sorted(capi.SYMBOLS.keys())


def test_library_exports_every_declared_symbol():
    from dicomautomaton_b200 import capi
    lib = capi.lib()
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (dcma_b200_\w+)", out))
    assert set(_declared_symbols()) <= exported
    # nothing else leaks out of the library (hidden visibility): only the C ABI is exported
    leaked = [l for l in out.splitlines() if " T " in l and "dcma_b200_" not in l]
    assert not leaked, leaked[:5]
    assert lib.dcma_b200_abi_version() == 1


def test_library_is_built_for_sm_100a():
    from dicomautomaton_b200 import capi
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_struct_layouts_match_the_header(tmp_path):
    """Compile a C program against the real header and compare sizeof/offsetof with the ctypes mirrors."""
    from dicomautomaton_b200 import capi
    fields = {
        "dcma_b200_geom": (capi.Geom, [f[0] for f in capi.Geom._fields_]),
        "dcma_b200_demons_params": (capi.Params, [f[0] for f in capi.Params._fields_]),
        "dcma_b200_run_stats": (capi.RunStats, [f[0] for f in capi.RunStats._fields_]),
        "dcma_b200_grid": (capi.Grid, [f[0] for f in capi.Grid._fields_]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dcma_b200/demons_c_api.h"', 'int main(void){']
    for name, (_, fl) in fields.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for f in fl:
            lines.append(f'printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines.append('return 0;}')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for name, (ct, fl) in fields.items():
        assert int(got[name]) == C.sizeof(ct), name
        for f in fl:
            assert int(got[f"{name}.{f}"]) == getattr(ct, f).offset, f"{name}.{f}"


def test_default_params_equal_the_reference_defaults():
    """src/Alignment_Demons.h:20-61."""
    from dicomautomaton_b200 import capi
    from dicomautomaton_b200.demons import AlignViaDemonsParams
    p = capi.default_params()
    ref = dict(max_iterations=100, convergence_threshold=0.001, deformation_field_smoothing_sigma=1.0,
               update_field_smoothing_sigma=0.5, use_diffeomorphic=0, use_histogram_matching=0, histogram_bins=256,
               histogram_outlier_fraction=0.01, normalization_factor=1.0, max_update_magnitude=2.0, verbosity=1)
    for k, v in ref.items():
        assert getattr(p, k) == v, k
    # extensions default to reference behaviour
    assert p.pyramid_levels == 1 and p.use_symmetric_force == 0 and p.fused == 1 and p.device == -1
    q = AlignViaDemonsParams()
    for k, v in ref.items():
        assert getattr(q, k) == v, k


def test_bad_arguments_are_rejected_without_a_device():
    from dicomautomaton_b200 import capi
    lib = capi.lib()
    err = C.create_string_buffer(256)
    h = C.c_void_p()
    p = capi.default_params()
    assert lib.dcma_b200_engine_create(C.byref(h), None, C.byref(p), None, err, len(err)) == capi.EINVAL
    g = capi.Geom(0, 4, 4, 1, 1.0, 1.0, 1.0)
    assert lib.dcma_b200_engine_create(C.byref(h), C.byref(g), C.byref(p), None, err, len(err)) == capi.EINVAL
    assert b">= 1" in err.value
    g = capi.Geom(4, 4, 4, 1, 1.0, -1.0, 1.0)
    assert lib.dcma_b200_engine_create(C.byref(h), C.byref(g), C.byref(p), None, err, len(err)) == capi.EINVAL
    g = capi.Geom(4, 4, 4, 1, 1.0, 1.0, 1.0)
    p.field_mode = 17
    assert lib.dcma_b200_engine_create(C.byref(h), C.byref(g), C.byref(p), None, err, len(err)) == capi.EINVAL
    assert lib.dcma_b200_engine_run(None, 1, 0, None, None) == capi.EINVAL
    assert lib.dcma_b200_last_error()


def _no_gpu():
    from dicomautomaton_b200 import capi
    return capi.lib().dcma_b200_device_count() == 0


def test_no_cpu_fallback_exists():
    """Without a CUDA device the product path must FAIL, not compute on the CPU."""
    from dicomautomaton_b200 import capi
    from dicomautomaton_b200.demons import AlignViaDemons, AlignViaDemonsParams, DemonsEngine, demons_volume
    if not _no_gpu():
        pytest.skip("a GPU is present")
    v = demons_volume(np.zeros((4, 6, 8), np.float32))
    with pytest.raises(capi.DcmaError) as ei:
        DemonsEngine(v, v, AlignViaDemonsParams(verbosity=0))
    assert ei.value.code == capi.ENODEVICE
    # the reference-shaped entry point reports failure the reference's way: nullopt
    assert AlignViaDemons(AlignViaDemonsParams(verbosity=0), v, v) is None
    # the entry points added in round 2 keep the rule
    from dicomautomaton_b200.demons import make_grid, register_stacks, virtual_field_v1, warp_to_grid
    g = make_grid((4, 6, 8), (1.0, 1.0, 1.0))
    img = np.zeros((4, 6, 8), np.float32)
    for call in (lambda: register_stacks(AlignViaDemonsParams(verbosity=0, max_iterations=2), g, img, g, img),
                 lambda: virtual_field_v1("vortex_v1", (4, 6, 8)),
                 lambda: warp_to_grid(g, np.zeros((4, 6, 8, 3)), g, img, g, img)):
        with pytest.raises(capi.DcmaError) as ei:
            call()
        assert ei.value.code == capi.ENODEVICE
    # the device pool is pure bookkeeping: usable (and empty) without a device
    capi.lib().dcma_b200_device_pool_trim()
    assert capi.lib().dcma_b200_device_pool_bytes() == 0


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under dicomautomaton_b200/ may import, load or link it."""
    pkg = os.path.join(ROOT, "dicomautomaton_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath or os.sep + "lib" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "libdemons_oracle" not in txt and "libdemons_ref" not in txt, f
    code = "import sys; import dicomautomaton_b200.demons, dicomautomaton_b200.capi, dicomautomaton_b200.phantom; " \
           "sys.exit(1 if any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules) else 0)"
    assert subprocess.run([sys.executable, "-c", code], cwd=ROOT).returncode == 0


def test_empty_input_returns_none_like_the_reference():
    """AlignViaDemons with an empty image array returns nullopt (src/Alignment_Demons.cc:953-956)."""
    from dicomautomaton_b200.demons import AlignViaDemons, AlignViaDemonsParams, demons_volume
    empty = demons_volume(np.zeros((0, 4, 4), np.float32))
    full = demons_volume(np.zeros((2, 4, 4), np.float32))
    assert AlignViaDemons(AlignViaDemonsParams(verbosity=0), empty, full) is None
    assert AlignViaDemons(AlignViaDemonsParams(verbosity=0), full, empty) is None


def test_header_is_plain_c_and_a_c_program_links_against_the_library(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 with warnings as errors, and examples/register_from_c.c links
    against libdcma_b200.so; without a device it reports DCMA_B200_ENODEVICE and exits 0 (no CPU fallback)."""
    from dicomautomaton_b200 import capi
    hdr = tmp_path / "only_header.c"
    hdr.write_text('#include "dcma_b200/demons_c_api.h"\nint main(void){return DCMA_B200_ABI_VERSION == 1 ? 0 : 1;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-fsyntax-only", str(hdr)])
    exe = tmp_path / "register_from_c"
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "register_from_c.c"), "-L", libdir, "-ldcma_b200",
                           f"-Wl,-rpath,{libdir}", "-lm", "-o", str(exe)])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "dcma_b200 ABI 1" in r.stdout
    if capi.lib().dcma_b200_device_count() == 0:
        assert "registration failed (-2)" in r.stdout
    else:
        assert "MSE" in r.stdout
