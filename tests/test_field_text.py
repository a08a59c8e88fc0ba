"""Row f3: the text form of a stored deformation field (reference src/Alignment_Field.cc:322-436).

CPU tests: the conversion arithmetic the kernels compile (csrc/text_format.cuh) against this C library's printf/strtod
-- which is what the reference's `os << val` / `is >> val` execute --, the Python oracle against known answers and the
reference's round-trip cases, and the no-fallback rule.  GPU tests: the C-ABI calls against the oracle, byte for byte
and bit for bit, including sizes that cross the library's internal chunking, and the host-side writer/reader.
"""
import io
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_gpu

NATIVE = os.path.join(ROOT, "tests", "native")
BUILD = os.path.join(NATIVE, "_build")
LIBDIR = os.path.join(ROOT, "dicomautomaton_b200", "lib")
HOST = os.path.join(ROOT, "dicomautomaton_b200", "host")
RPATH = "-Wl,-rpath,$ORIGIN/../../../dicomautomaton_b200/lib"


def _mixed_values(n, seed=0):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    any_bits = bits.view(np.float64).copy()
    any_bits[~np.isfinite(any_bits)] = 1.5
    dyadic = np.ldexp((rng.integers(0, 2**44, size=n) | 1).astype(np.float64), -rng.integers(0, 80, size=n))
    grid = 0.05 * rng.integers(-2000, 2001, size=n)
    disp = rng.normal(0, 3, size=n) * 10.0 ** -rng.integers(0, 12, size=n)
    kind = rng.integers(0, 5, size=n)
    out = np.where(kind == 0, any_bits, np.where(kind == 1, dyadic, np.where(kind == 2, grid, disp)))
    out[kind == 4] = np.where(rng.random(np.count_nonzero(kind == 4)) < 0.1, 0.0, np.round(out[kind == 4] * 1000))
    return np.ascontiguousarray(out, dtype=np.float64)


# ---- CPU ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1], ids=["fast-path", "exact-path"])
def test_conversion_arithmetic_equals_the_c_library(tmp_path, mode):
    """format_g17 == printf("%.17g") and parse_double == strtod over ~2 million values per mode: random bit patterns,
    displacement-like values, exact ties, powers of two and ten, subnormals, long literals, malformed tokens."""
    exe = tmp_path / "text_format_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", os.path.join(NATIVE, "text_format_check.cc"), "-o", str(exe)])
    r = subprocess.run([str(exe), "300000" if mode == 0 else "60000", str(mode)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " mismatches 0 " in r.stdout and "undecided 0" in r.stdout


def test_pow10_table_is_what_the_generator_writes(tmp_path):
    src = open(os.path.join(ROOT, "dicomautomaton_b200", "csrc", "pow10_table.inc")).read().splitlines()
    rows = [l for l in src if l.startswith("{")]
    assert len(rows) == 701
    for k in (-350, -1, 0, 1, 27, 55, 56, 350):
        hi, lo, e2 = rows[k + 350].strip("{},").replace("ull", "").split(",")
        g = (int(hi, 16) << 64) | int(lo, 16)
        assert g >> 127 == 1
        # 10^k in [g, g+1) * 2^e2, exact for 0 <= k <= 55
        e2 = int(e2)
        if k >= 0:
            lo_ok = (g << e2 if e2 >= 0 else g) <= (10**k if e2 >= 0 else 10**k << -e2)
            hi_ok = (10**k if e2 >= 0 else 10**k << -e2) < ((g + 1) << e2 if e2 >= 0 else g + 1)
            assert lo_ok and hi_ok
            if k <= 55:
                assert (g << e2 if e2 >= 0 else g) == (10**k if e2 >= 0 else 10**k << -e2)
        else:
            assert g * 10**-k <= 1 << -e2 < (g + 1) * 10**-k


def test_oracle_known_answers_and_reference_round_trip_cases():
    from oracle import field_text as ft
    assert ft.format_double(0.1) == "0.10000000000000001"
    assert ft.format_double(-0.2) == "-0.20000000000000001"
    assert ft.format_double(0.05) == "0.050000000000000003"
    assert ft.format_double(100001 / 2**18) == "0.38147354125976562"   # exact tie, half-even
    assert ft.format_double(1e23) == "9.9999999999999992e+22"
    assert ft.format_double(5e-324) == "4.9406564584124654e-324"
    assert ft.format_double(0.0) == "0" and ft.format_double(-0.0) == "-0" and ft.format_double(123456.0) == "123456"
    assert ft.format_double(1e16) == "10000000000000000" and ft.format_double(1e17) == "1e+17" and ft.format_double(1e-5) == "1.0000000000000001e-05"

    # src/Alignment_Field_Tests.cc:196-248: 2 x 3 x 4 field, (0.1 col, -0.2 row, 0.05 slice), metadata with special characters
    def image(s, rows, cols, fn, meta):
        data = [c for r in range(rows) for col in range(cols) for c in fn(s, r, col)]
        return dict(rows=rows, cols=cols, pxl_dx=1.0, pxl_dy=1.0, pxl_dz=1.0, anchor=(0, 0, 0), offset=(0, 0, float(s)),
                    row_unit=(1, 0, 0), col_unit=(0, 1, 0), metadata=meta, data=data)
    meta = {"PatientID": "test_patient_001", "key with spaces": "value;with@special=chars"}
    imgs = [image(s, 3, 4, lambda s, r, c: (0.1 * c, -0.2 * r, 0.05 * s), meta) for s in range(2)]
    text = ft.write_field_text(imgs)
    back = ft.read_field_text(text)
    assert [b["data"] for b in back] == [i["data"] for i in imgs] and back[0]["metadata"] == meta
    assert text.startswith(b"2\n3 4 3\n1 1 1\n(0, 0, 0)\n(0, 0, 0)\n(1, 0, 0)\n(0, 1, 0)\nnum_metadata= 2\nUGF0aWVudElE dGVzdF9wYXRpZW50XzAwMQ==\n")
    # :268-289: empty stream, negative / zero image count, truncated data
    for bad in (b"", b"-1\n", b"0\n", b"1\n3 3 3\n"):
        with pytest.raises(ValueError):
            ft.read_field_text(bad)


def test_text_calls_have_no_cpu_fallback():
    if has_gpu():
        pytest.skip("a device is present")
    from dicomautomaton_b200 import capi, field_io
    with pytest.raises(capi.DcmaError) as ei:
        field_io.format_doubles(np.array([0.1, 0.2]))
    assert ei.value.code == capi.ENODEVICE
    with pytest.raises(capi.DcmaError) as ei:
        field_io.parse_doubles(b"1 2 3", 3)
    assert ei.value.code == capi.ENODEVICE
    # empty requests need no device
    assert field_io.format_doubles(np.array([])) == b""


def test_native_checks_build_and_report_no_device(tmp_path):
    """The plain-C and C++ GPU checks compile and link here; without a device they exit 77 (skip)."""
    exes = _build_native(str(tmp_path))
    if has_gpu():
        pytest.skip("a device is present")
    for exe in exes:
        assert subprocess.run([exe, "1000"], capture_output=True).returncode == 77


def _build_native(outdir):
    """[plain-C caller of the two entry points, the patched reference writer/reader driven by the reference's cases].  The
    second is built where /root/reference exists (host/build_host_tests.py::build_field_io) and travels prebuilt."""
    import sys
    sys.path.insert(0, HOST)
    import build_host_tests
    os.makedirs(outdir, exist_ok=True)
    rpath = "-Wl,-rpath," + LIBDIR
    c_exe = os.path.join(outdir, "text_io_gpu_check")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-Wall", "-D_POSIX_C_SOURCE=200809L", "-I", os.path.join(ROOT, "include"),
                           os.path.join(NATIVE, "text_io_gpu_check.c"), "-o", c_exe, "-L", LIBDIR, "-ldcma_b200", "-lm", rpath])
    cc_exe = build_host_tests.build_field_io(verbose=False)
    if cc_exe is None:
        pytest.skip("reference sources not available and no prebuilt host/_build/field_io_check")
    return [c_exe, cc_exe]


def test_patched_reference_writer_reader_host_logic_on_cpu(tmp_path):
    """The reference's own write_to / read_from with patches/Alignment_Field.patch applied, plus the payload codec's host
    logic (bounded read-ahead that has to grow, stream left just behind the last number, failure states), run on CPU: the
    objects of build_field_io are linked against tests/native/fake_text_capi.c (printf / strtod) instead of the GPU
    library, so this checks everything except the conversion kernels, which the GPU tests check."""
    import sys
    sys.path.insert(0, HOST)
    import build_host_tests
    if build_host_tests.build_field_io(verbose=False) is None:
        pytest.skip("reference sources not available")
    objs = [os.path.join(build_host_tests.OUT, o) for o in ("ref_field_io_patched.o", "field_io_check.o", "Field_Payload.o")]
    if not all(os.path.isfile(o) for o in objs):
        pytest.skip("objects of build_field_io not present (prebuilt executable only)")
    fake = os.path.join(str(tmp_path), "fake_text_capi.o")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-D_POSIX_C_SOURCE=200809L", "-c", os.path.join(NATIVE, "fake_text_capi.c"), "-o", fake])
    exe = os.path.join(str(tmp_path), "field_io_check_cpu")
    subprocess.check_call(["g++", "-o", exe, *objs, fake])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr


def test_field_patch_is_current():
    """patches/Alignment_Field.patch is what tools/make_field_patch.py derives from the reference file, and touches only
    the include list and the two per-value loops."""
    ref = os.environ.get("DCMA_REFERENCE_SRC", "/root/reference/src")
    if not os.path.isfile(os.path.join(ref, "Alignment_Field.cc")):
        pytest.skip("reference sources not available")
    patch = open(os.path.join(ROOT, "patches", "Alignment_Field.patch")).read()
    removed = [l[1:].strip() for l in patch.splitlines() if l.startswith("-") and not l.startswith("---")]
    assert removed == ["for(const auto &val : img.data){", "os << val << '\\n';", "for(auto &val : img.data){", "is >> val;", "}"]
    out = subprocess.run(["patch", "-s", "-o", "-", os.path.join(ref, "Alignment_Field.cc"), os.path.join(ROOT, "patches", "Alignment_Field.patch")],
                         capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.count("dcma_b200_io::") == 2


# ---- GPU ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_format_is_byte_identical_to_printf():
    from dicomautomaton_b200 import field_io
    from oracle import field_text as ft
    v = _mixed_values(300_000, seed=1)
    v[:8] = [0.0, -0.0, 0.1, 100001 / 2**18, 1e23, 5e-324, 1.7976931348623157e308, -123456.0]
    got = field_io.format_doubles(v)
    want = ft.format_doubles(v)
    assert got == want
    # non-finite values print like printf
    assert field_io.format_doubles(np.array([np.inf, -np.inf, np.nan])) == b"inf\n-inf\nnan\n"
    assert field_io.format_doubles(np.array([struct.unpack("<d", struct.pack("<Q", 0xFFF8000000000000))[0]])) == b"-nan\n"
    for n in (1, 255, 256, 257, 1023, 1024, 1025, 4097):  # tile edges
        assert field_io.format_doubles(v[:n]) == ft.format_doubles(v[:n])


@pytest.mark.gpu
def test_parse_is_bit_identical_to_strtod_and_round_trips():
    from dicomautomaton_b200 import field_io
    v = _mixed_values(300_000, seed=2)
    text = field_io.format_doubles(v)
    back, used = field_io.parse_doubles(text, v.size)
    assert used == len(text) - 1
    assert np.array_equal(back.view(np.uint64), v.view(np.uint64))
    # literals the writer never produces: long digit strings, odd spellings, arbitrary white space; float() is strtod
    rng = np.random.default_rng(3)
    lits = []
    for _ in range(20000):
        a = "".join(rng.choice(list("0123456789"), size=rng.integers(1, 23)))
        b = "".join(rng.choice(list("0123456789"), size=rng.integers(0, 25)))
        e = ("e%+d" % rng.integers(-340, 300)) if rng.random() < 0.5 else ""
        lits.append(("-" if rng.random() < 0.3 else "") + a + ("." + b if b or rng.random() < 0.2 else "") + e)
    lits += ["2.4703282292062327e-324", "2.4703282292062328e-324", "9007199254740993", "9007199254740992.5",
             "9007199254740993.0000000000000000001", "1e-400", "0e999", ".5", "5.", "+7", "00012.5000e-00"]
    want = np.array([float(l) for l in lits])
    keep = np.isfinite(want)
    lits = [l for l, k in zip(lits, keep) if k]
    want = want[keep]
    seps = [" ", "\n", "\t", "\r\n", "  \n ", "\v", "\f"]
    blob = "".join(l + seps[i % len(seps)] for i, l in enumerate(lits)).encode()
    got, used = field_io.parse_doubles(b" \n" + blob, len(lits))
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    assert used == 2 + len(blob) - len(seps[(len(lits) - 1) % len(seps)])
    # a prefix of the numbers only
    got, used = field_io.parse_doubles(b"1.5 2.5 3.5 rest", 2)
    assert got.tolist() == [1.5, 2.5] and used == 7


@pytest.mark.gpu
def test_parse_errors_are_the_reference_read_failures():
    from dicomautomaton_b200 import capi, field_io
    for bad, n in ((b"1 2 3", 4), (b"", 1), (b"   \n", 1), (b"1 2x 3", 3), (b"1 inf 3", 3), (b"nan", 1), (b"1e999", 1),
                   (b"1.2.3", 1), (b"--1", 1), (b"1e", 1), (b"0x10", 1)):
        with pytest.raises(capi.DcmaError) as ei:
            field_io.parse_doubles(bad, n)
        assert ei.value.code == capi.EINVAL and "could not be read" in str(ei.value)
    with pytest.raises(capi.DcmaError):  # capacity too small is an argument error, not a truncated result
        buf = np.empty(10, dtype=np.uint8)
        import ctypes as C
        n_out = C.c_int64(0)
        v = np.arange(100, dtype=np.float64)
        capi.check(capi.lib().dcma_b200_format_doubles(v.ctypes.data, v.size, buf.ctypes.data, buf.size, C.byref(n_out), -1, None, 0))


@pytest.mark.gpu
def test_field_text_equals_the_oracle_and_round_trips():
    from dicomautomaton_b200 import demons, field_io
    from oracle import field_text as ft
    from _helpers import random_smooth_field
    shape, spacing, origin = (5, 24, 20), (0.9, 1.1, 2.5), (-12.5, 3.25, 100.0)
    D = random_smooth_field(shape, max_mm=4.0, seed=5)
    D[0, 0, 0] = (0.0, -0.0, 1e-300)
    grid = demons.make_grid(shape, spacing, origin)
    meta = {"PatientID": "test_patient_001", "key with spaces": "value;with@special=chars"}
    f = io.BytesIO()
    field_io.write_field_text(f, D, grid, meta)
    imgs = [dict(rows=shape[1], cols=shape[2], pxl_dx=spacing[0], pxl_dy=spacing[1], pxl_dz=spacing[2], anchor=(0, 0, 0),
                 offset=(origin[0], origin[1], origin[2] + z * spacing[2]), row_unit=(1, 0, 0), col_unit=(0, 1, 0),
                 metadata=meta, data=D[z].reshape(-1)) for z in range(shape[0])]
    assert f.getvalue() == ft.write_field_text(imgs)
    back, images = field_io.read_field_text(io.BytesIO(f.getvalue()))
    assert np.array_equal(back.view(np.uint64), D.view(np.uint64))
    assert images[3]["metadata"] == meta and images[3]["offset"] == (origin[0], origin[1], origin[2] + 3 * spacing[2])
    # and the oracle reads what the GPU wrote
    assert [list(i["data"]) for i in ft.read_field_text(f.getvalue())] == [list(D[z].reshape(-1)) for z in range(shape[0])]


@pytest.mark.gpu
@pytest.mark.slow
def test_sizes_that_cross_the_internal_chunks_round_trip():
    """11.5 million numbers (~225 MB of text): two formatting passes (8 Mi numbers each) and two parsing passes (192 MiB of text each).
    Size-independent checks: the text reads back bit-identically, and its first / last / chunk-boundary lines are the
    oracle's."""
    from dicomautomaton_b200 import field_io
    from oracle import field_text as ft
    v = _mixed_values(11_500_000, seed=4)  # ~19.6 bytes per line measured
    v[v == np.round(v)] += 0.123456789  # keep the lines long enough to pass 192 MiB
    text = field_io.format_doubles(v)
    assert len(text) > (192 << 20)
    assert text.count(b"\n") == v.size
    back, used = field_io.parse_doubles(text, v.size)
    assert used == len(text) - 1 and np.array_equal(back.view(np.uint64), v.view(np.uint64))
    assert text.startswith(ft.format_doubles(v[:1000])) and text.endswith(ft.format_doubles(v[-1000:]))
    b = 8 << 20
    edge = ft.format_doubles(v[b - 500:b + 500])
    assert edge in text


@pytest.mark.gpu
def test_native_c_and_cpp_callers(tmp_path):
    """The plain-C caller of the two entry points (12 million numbers against printf / bit-identical read back, error
    codes) and the C++ host writer/reader on the reference's own round-trip cases (src/Alignment_Field_Tests.cc:196-289)."""
    c_exe, cc_exe = _build_native(str(tmp_path))
    r = subprocess.run([c_exe, "3000000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
    r = subprocess.run([cc_exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr


def test_trans_file_wrapper_and_header_logic_on_cpu(tmp_path, monkeypatch):
    """Host logic of field_io (headers, metadata, `.trans` wrapper, src/Transformation_File_Loader.cc:23-149) with the two
    GPU calls replaced by the oracle's conversions: the file equals the oracle's text and reads back."""
    from dicomautomaton_b200 import demons, field_io
    from oracle import field_text as ft

    def fake_parse(text, n, device=-1):
        toks, pos, vals = 0, 0, []
        while toks < n:
            while text[pos:pos + 1].isspace():
                pos += 1
            b = pos
            while pos < len(text) and not text[pos:pos + 1].isspace():
                pos += 1
            vals.append(float(text[b:pos]))
            toks += 1
        return np.array(vals), pos

    monkeypatch.setattr(field_io, "format_doubles", lambda v, device=-1: ft.format_doubles(np.asarray(v).reshape(-1)))
    monkeypatch.setattr(field_io, "parse_doubles", fake_parse)
    shape, spacing, origin = (2, 3, 4), (1.0, 1.0, 1.0), (0.0, 0.0, 0.0)
    D = np.zeros(shape + (3,))
    for s in range(2):
        for r in range(3):
            for c in range(4):
                D[s, r, c] = (0.1 * c, -0.2 * r, 0.05 * s)   # src/Alignment_Field_Tests.cc:197-200
    meta = {"PatientID": "test_patient_001", "key with spaces": "value;with@special=chars"}
    path = tmp_path / "field.trans"
    field_io.write_trans_file(str(path), D, demons.make_grid(shape, spacing, origin), meta)
    raw = path.read_bytes()
    imgs = [dict(rows=3, cols=4, pxl_dx=1.0, pxl_dy=1.0, pxl_dz=1.0, anchor=(0, 0, 0), offset=(0, 0, float(z)), row_unit=(1, 0, 0),
                 col_unit=(0, 1, 0), metadata=meta, data=D[z].reshape(-1)) for z in range(2)]
    assert raw == b"DCMA_TRANSFORM\nTRANSFORM_VARIANT_DEFORMATION_FIELD_3\n" + ft.write_field_text(imgs)
    back, images = field_io.read_trans_file(str(path))
    assert np.array_equal(back, D) and images[1]["metadata"] == meta and images[1]["offset"] == (0.0, 0.0, 1.0)
    # comment and blank lines before the magic / variant lines are skipped; anything else is rejected
    path.write_bytes(b"# a = b\n\n" + raw)
    assert np.array_equal(field_io.read_trans_file(str(path))[0], D)
    path.write_bytes(b"SOMETHING_ELSE\n" + raw)
    with pytest.raises(ValueError):
        field_io.read_trans_file(str(path))
