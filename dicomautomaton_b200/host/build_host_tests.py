#!/usr/bin/env python3
"""Build `host/_build/demons_host_tests`: the reference's OWN unit-test file for this path, compiled IN PLACE from
/root/reference/src/Alignment_Demons_Tests.cc (read through stdin so that its quoted includes resolve to our drop-in
header and the shims instead of the reference's Ygor-dependent headers), linked against our AlignViaDemons drop-in and
libdcma_b200.so.  Nothing of the reference is copied into the repository; the binary is git-ignored and travels to the
GPU box with the snapshot.  No-op when /root/reference is absent.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = os.environ.get("DCMA_REFERENCE_SRC", "/root/reference/src")
OUT = os.path.join(HERE, "_build")
EXE = os.path.join(OUT, "demons_host_tests")


def build(verbose=True):
    test_src = os.path.join(REF_SRC, "Alignment_Demons_Tests.cc")
    if not os.path.isfile(test_src):
        if verbose:
            print(f"[host tests] {test_src} not found; {'using prebuilt ' + EXE if os.path.isfile(EXE) else 'nothing to do'}")
        return EXE if os.path.isfile(EXE) else None
    os.makedirs(OUT, exist_ok=True)
    libdir = os.path.join(os.path.dirname(HERE), "lib")
    inc = ["-I", os.path.join(HERE, "shim", "dcma_min"), "-I", os.path.join(HERE, "shim", "ygor_min"), "-I", HERE,
           "-I", os.path.join(ROOT, "include"), "-I", REF_SRC]
    cxx = ["g++", "-std=c++17", "-O2", "-Wno-unused-function", "-Wno-unused-variable"]
    objs = []
    # 1. the reference's test file, in place, via stdin (so "Structs.h" etc. resolve through -I, not its own directory)
    o = os.path.join(OUT, "ref_tests.o")
    cmd = cxx + inc + ["-x", "c++", "-c", "-", "-o", o]
    if verbose:
        print("[host tests]", " ".join(cmd), "<", test_src)
    with open(test_src, "rb") as f:
        subprocess.check_call(cmd, stdin=f, cwd=OUT)
    objs.append(o)
    # 2. doctest main (header used in place from the reference's vendored copy)
    main_cc = os.path.join(OUT, "doctest_main.cc")
    with open(main_cc, "w") as f:
        f.write('#define DOCTEST_CONFIG_IMPLEMENT_WITH_MAIN\n#include "doctest20251212/doctest.h"\n')
    o = os.path.join(OUT, "doctest_main.o")
    subprocess.check_call(cxx + ["-I", REF_SRC, "-c", main_cc, "-o", o])
    objs.append(o)
    # 3. our drop-in
    o = os.path.join(OUT, "Alignment_Demons.o")
    subprocess.check_call(cxx + inc + ["-c", os.path.join(HERE, "Alignment_Demons.cc"), "-o", o])
    objs.append(o)
    cmd = ["g++", "-o", EXE, *objs, "-L", libdir, "-ldcma_b200", "-Wl,-rpath,$ORIGIN/../../lib"]
    if verbose:
        print("[host tests]", " ".join(cmd))
    subprocess.check_call(cmd)
    return EXE


OP_EXE = os.path.join(OUT, "demons_operation_check")
OP_PATCHED_EXE = os.path.join(OUT, "demons_operation_check_patched")
OP_PATCH = os.path.join(ROOT, "patches", "RegisterImagesDemons.patch")


def build_operation(verbose=True, patched=False):
    """`host/_build/demons_operation_check`: the reference's OWN operation file, src/Operations/RegisterImagesDemons.cc,
    compiled in place and unmodified (read through stdin from a scratch directory whose parent holds forwarding headers,
    so that its `../Structs.h`-style includes resolve to the stand-ins and to our Alignment_Demons.h), linked with the
    AlignViaDemons drop-in, libdcma_b200.so and the driver tests/native/operation_check.cc."""
    op_src = os.path.join(REF_SRC, "Operations", "RegisterImagesDemons.cc")
    exe = OP_PATCHED_EXE if patched else OP_EXE
    if not os.path.isfile(op_src):
        if verbose:
            print(f"[host tests] {op_src} not found; {'using prebuilt ' + exe if os.path.isfile(exe) else 'nothing to do'}")
        return exe if os.path.isfile(exe) else None
    ops = os.path.join(OUT, "ops")
    os.makedirs(ops, exist_ok=True)
    libdir = os.path.join(os.path.dirname(HERE), "lib")
    shim = os.path.join(HERE, "shim")
    inc = ["-I", os.path.join(shim, "dcma_min"), "-I", os.path.join(shim, "ygor_min"), "-I", HERE, "-I", os.path.join(ROOT, "include")]
    cxx = ["g++", "-std=c++17", "-O2", "-Wno-unused-function", "-Wno-unused-variable"]
    # forwarding headers: "../X.h" as seen from the scratch directory -> <X.h> on the include path (angle brackets, so that
    # a forwarder never finds itself)
    for h in ("Structs.h", "Regex_Selectors.h", "Thread_Pool.h", "Alignment_Rigid.h", "Alignment_Field.h", "Alignment_Demons.h"):
        with open(os.path.join(OUT, h), "w") as f:
            f.write(f"#pragma once\n#include <{h}>\n")
    # the operation's own header declares exactly these two functions (src/Operations/RegisterImagesDemons.h:13-18); the
    # reference's copy cannot be used because its relative includes would pull in the real Structs.h
    with open(os.path.join(ops, "RegisterImagesDemons.h"), "w") as f:
        f.write("#pragma once\n#include <map>\n#include <string>\n#include <Structs.h>\n#include <Regex_Selectors.h>\n"
                "OperationDoc OpArgDocRegisterImagesDemons();\n"
                "bool RegisterImagesDemons(Drover &, const OperationArgPkg &, std::map<std::string, std::string> &, const std::string &);\n")
    objs = []
    o = os.path.join(OUT, "ref_operation_patched.o" if patched else "ref_operation.o")
    cmd = cxx + inc + ["-x", "c++", "-c", "-", "-o", o]
    if verbose:
        print("[host tests]", " ".join(cmd), "<", op_src, "(+ patches/RegisterImagesDemons.patch)" if patched else "")
    if patched:
        # the reference file with patches/RegisterImagesDemons.patch applied, streamed: `patch -o -` writes the result
        # to stdout, nothing of the reference is written into the repository
        text = subprocess.run(["patch", "-s", "-o", "-", op_src, OP_PATCH], check=True, capture_output=True).stdout
        subprocess.run(cmd, input=text, check=True, cwd=ops)
    else:
        with open(op_src, "rb") as f:
            subprocess.check_call(cmd, stdin=f, cwd=ops)
    objs.append(o)
    for src, name in ((os.path.join(ROOT, "tests", "native", "operation_check.cc"), "operation_check.o"),
                      (os.path.join(HERE, "Alignment_Demons.cc"), "Alignment_Demons.o")):
        o = os.path.join(OUT, name)
        subprocess.check_call(cxx + inc + ["-c", src, "-o", o])
        objs.append(o)
    cmd = ["g++", "-pthread", "-o", exe, *objs, "-L", libdir, "-ldcma_b200", "-Wl,-rpath,$ORIGIN/../../lib"]
    if verbose:
        print("[host tests]", " ".join(cmd))
    subprocess.check_call(cmd)
    return exe


WARP_EXE = os.path.join(OUT, "warp_operation_check")
WARP_PATCHED_EXE = os.path.join(OUT, "warp_operation_check_patched")
WARP_PATCH = os.path.join(ROOT, "patches", "WarpImages.patch")


def build_warp_operation(verbose=True, patched=False):
    """`host/_build/warp_operation_check[_patched]`: the reference's OWN operation file src/Operations/WarpImages.cc,
    compiled in place -- unmodified, or with patches/WarpImages.patch applied (streamed through `patch -o -`, nothing of
    the reference is written into the repository) -- against the stand-in headers, linked with host/Warp_Images_Field.cc,
    libdcma_b200.so and the driver tests/native/warp_operation_check.cc."""
    op_src = os.path.join(REF_SRC, "Operations", "WarpImages.cc")
    exe = WARP_PATCHED_EXE if patched else WARP_EXE
    if not os.path.isfile(op_src):
        if verbose:
            print(f"[host tests] {op_src} not found; {'using prebuilt ' + exe if os.path.isfile(exe) else 'nothing to do'}")
        return exe if os.path.isfile(exe) else None
    top = os.path.join(OUT, "warp")          # plays the reference's src/ ; the operation is compiled from warp/ops
    ops = os.path.join(top, "ops")
    os.makedirs(ops, exist_ok=True)
    libdir = os.path.join(os.path.dirname(HERE), "lib")
    shim = os.path.join(HERE, "shim")
    inc = ["-I", os.path.join(shim, "dcma_min"), "-I", os.path.join(shim, "ygor_min"), "-I", HERE, "-I", os.path.join(ROOT, "include")]
    cxx = ["g++", "-std=c++17", "-O2", "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unused-but-set-variable"]
    for h in ("Structs.h", "Regex_Selectors.h", "Thread_Pool.h", "Metadata.h", "Alignment_Rigid.h", "Alignment_TPSRPM.h",
              "Alignment_Field.h", "Warp_Images_Field.h"):
        with open(os.path.join(top, h), "w") as f:
            f.write(f"#pragma once\n#include <{h}>\n")
    for sub, h in (("YgorImages_Functors/Grouping", "Misc_Functors.h"),
                   ("YgorImages_Functors/Processing", "Partitioned_Image_Voxel_Visitor_Mutator.h")):
        os.makedirs(os.path.join(top, sub), exist_ok=True)
        with open(os.path.join(top, sub, h), "w") as f:
            f.write(f"#pragma once\n#include <{h}>\n")
    with open(os.path.join(ops, "WarpImages.h"), "w") as f:  # the operation's own header declares exactly these two
        f.write("#pragma once\n#include <map>\n#include <string>\n#include <Structs.h>\n#include <Regex_Selectors.h>\n"
                "OperationDoc OpArgDocWarpImages();\n"
                "bool WarpImages(Drover &, const OperationArgPkg &, std::map<std::string, std::string> &, const std::string &);\n")
    objs = []
    o = os.path.join(OUT, "ref_warp_operation_patched.o" if patched else "ref_warp_operation.o")
    cmd = cxx + inc + ["-x", "c++", "-c", "-", "-o", o]
    if verbose:
        print("[host tests]", " ".join(cmd), "<", op_src, "(+ patches/WarpImages.patch)" if patched else "")
    if patched:
        text = subprocess.run(["patch", "-s", "-o", "-", op_src, WARP_PATCH], check=True, capture_output=True).stdout
        subprocess.run(cmd, input=text, check=True, cwd=ops)
    else:
        with open(op_src, "rb") as f:
            subprocess.check_call(cmd, stdin=f, cwd=ops)
    objs.append(o)
    for src, name in ((os.path.join(ROOT, "tests", "native", "warp_operation_check.cc"), "warp_operation_check.o"),
                      (os.path.join(HERE, "Warp_Images_Field.cc"), "Warp_Images_Field.o")):
        o = os.path.join(OUT, name)
        subprocess.check_call(cxx + inc + ["-c", src, "-o", o])
        objs.append(o)
    subprocess.check_call(["g++", "-pthread", "-o", exe, *objs, "-L", libdir, "-ldcma_b200", "-Wl,-rpath,$ORIGIN/../../lib"])
    return exe


HARNESS = os.path.join(OUT, "libdcma_b200_dropin.so")


def build_harness(verbose=True):
    """`host/_build/libdcma_b200_dropin.so`: the C++ drop-in `AlignViaDemons` (host/Alignment_Demons.cc, against the minimal
    Ygor stand-in) behind a C entry, so that bench.py can time the call an operation makes -- both marshalling steps,
    the grid resampling and the pinned-buffer pool inside the timed region.  Needs no reference sources."""
    os.makedirs(OUT, exist_ok=True)
    libdir = os.path.join(os.path.dirname(HERE), "lib")
    shim = os.path.join(HERE, "shim")
    inc = ["-I", os.path.join(shim, "dcma_min"), "-I", os.path.join(shim, "ygor_min"), "-I", HERE, "-I", os.path.join(ROOT, "include")]
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-fvisibility=hidden", "-Wno-unused-function", *inc,
           os.path.join(HERE, "dropin_harness.cc"), os.path.join(HERE, "Alignment_Demons.cc"),
           "-o", HARNESS, "-L", libdir, "-ldcma_b200", "-Wl,-rpath,$ORIGIN/../../lib"]
    if verbose:
        print("[host tests]", " ".join(cmd))
    subprocess.check_call(cmd)
    return HARNESS


FIELD_IO_EXE = os.path.join(OUT, "field_io_check")
FIELD_PATCH = os.path.join(ROOT, "patches", "Alignment_Field.patch")


def build_field_io(verbose=True):
    """`host/_build/field_io_check`: the reference's OWN `deformation_field::write_to` / `read_from`
    (src/Alignment_Field.cc:322-436) with patches/Alignment_Field.patch applied -- the two per-value loops call the GPU
    payload codec host/Field_Payload.cc -- driven by the reference's round-trip cases (tests/native/field_io_check.cc).
    The patched text is streamed (`patch -o -`), cut at the textual anchor of write_to (the members before it need the
    whole of Ygor) and compiled from stdin against the stand-in class declaration; nothing of it is written to the repo."""
    src = os.path.join(REF_SRC, "Alignment_Field.cc")
    if not os.path.isfile(src):
        if verbose:
            print(f"[host tests] {src} not found; {'using prebuilt ' + FIELD_IO_EXE if os.path.isfile(FIELD_IO_EXE) else 'nothing to do'}")
        return FIELD_IO_EXE if os.path.isfile(FIELD_IO_EXE) else None
    os.makedirs(OUT, exist_ok=True)
    libdir = os.path.join(os.path.dirname(HERE), "lib")
    shim = os.path.join(HERE, "shim")
    inc = ["-I", os.path.join(shim, "dcma_min"), "-I", os.path.join(shim, "ygor_min"), "-I", HERE, "-I", os.path.join(ROOT, "include")]
    cxx = ["g++", "-std=c++17", "-O2", "-Wno-unused-function", "-Wno-unused-variable"]
    text = subprocess.run(["patch", "-s", "-o", "-", src, FIELD_PATCH], check=True, capture_output=True).stdout.decode()
    anchor = "bool\ndeformation_field::write_to("
    if text.count(anchor) != 1 or "dcma_b200_io::write_payload" not in text or "dcma_b200_io::read_payload" not in text:
        raise RuntimeError("patched Alignment_Field.cc does not look as expected")
    prologue = ("#include <cstdint>\n#include <iostream>\n#include <limits>\n#include <string>\n#include \"YgorImages.h\"\n"
                "#include \"YgorMath.h\"\n#include \"YgorMisc.h\"\n#include \"YgorLog.h\"\n#include \"YgorBase64.h\"\n"
                "#include \"Alignment_Field.h\"\n#include \"Field_Payload.h\"\n")
    o_ref = os.path.join(OUT, "ref_field_io_patched.o")
    cmd = cxx + inc + ["-x", "c++", "-c", "-", "-o", o_ref]
    if verbose:
        print("[host tests]", " ".join(cmd), "<", src, "(+ patches/Alignment_Field.patch, from write_to on)")
    subprocess.run(cmd, input=(prologue + text[text.index(anchor):]).encode(), check=True, cwd=OUT)
    objs = [o_ref]
    for s, name in ((os.path.join(ROOT, "tests", "native", "field_io_check.cc"), "field_io_check.o"),
                    (os.path.join(HERE, "Field_Payload.cc"), "Field_Payload.o")):
        o = os.path.join(OUT, name)
        subprocess.check_call(cxx + inc + ["-c", s, "-o", o])
        objs.append(o)
    subprocess.check_call(["g++", "-o", FIELD_IO_EXE, *objs, "-L", libdir, "-ldcma_b200", "-Wl,-rpath,$ORIGIN/../../lib"])
    return FIELD_IO_EXE


if __name__ == "__main__":
    build_harness()
    build_field_io()
    build_warp_operation()
    build_warp_operation(patched=True)
    p = build()
    q = build_operation()
    build_operation(patched=True)
    print(p if p else "unavailable")
    print(q if q else "unavailable")
    sys.exit(0 if p else 1)
