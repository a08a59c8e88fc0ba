"""Build `dicomautomaton_b200/lib/libdcma_b200.so` (C ABI + sm_100a kernels) IN-TREE with nvcc.

    python -m dicomautomaton_b200.build [--force] [--verbose]

The library is built for sm_100a only (`-gencode arch=compute_100a,code=sm_100a -lineinfo`); the strict fp64
translation unit gets `-fmad=false` so nothing in it is contracted into FMAs (bit-parity with the reference's x86-64
build).  nvcc cross-compiles without a GPU.  The cudart is linked statically so the .so has no CUDA runtime dependency
beyond the driver.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libdcma_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "-I", INCLUDE, "-I", CSRC]

# translation unit -> extra flags
UNITS = {
    "engine_f64.cu": [],
    "engine_f64_strict.cu": ["-fmad=false"],
    "engine_f32.cu": [],
    "engine_mixed.cu": [],
    "slab.cu": [],
    "pyramid.cu": [],
    "warp_grid.cu": [],
    "virtual_field.cu": [],
    "histmatch.cu": ["-fmad=false"],
    "text_io.cu": [],
    "c_api.cu": [],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(INCLUDE, "dcma_b200")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h", ".inc")):
                h.update(name.encode())
                h.update(open(os.path.join(root, name), "rb").read())
    h.update(json.dumps([ARCH, COMMON, UNITS], sort_keys=True).encode())
    return h.hexdigest()


def build_variant(tag: str, defines, verbose: bool = False) -> str:
    """Kernel-tuning builds: the same library with extra -D flags, as lib/libdcma_b200_<tag>.so (selected at run time
    with DCMA_B200_LIB=<path>).  Never used by the product path or the tests."""
    nvcc = _nvcc()
    objdir = os.path.join(OBJDIR, "variant_" + tag)
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, f"libdcma_b200_{tag}.so")

    def one(item):
        src, extra = item
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(one, UNITS.items()))
    subprocess.check_call([nvcc, *ARCH, "-shared", "-cudart", "static", "-Xcompiler", "-fPIC", *objs, "-o", out])
    if verbose:
        print("[dcma_b200.build] variant", tag, defines, "->", out)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "BUILD_INFO.json")
    digest = _digest()
    if not force and os.path.isfile(LIB) and os.path.isfile(stamp):
        try:
            if json.load(open(stamp)).get("sha256") == digest:
                return LIB
        except Exception:
            pass
    nvcc = _nvcc()

    def compile_one(item):
        src, extra = item
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print("[dcma_b200.build]", " ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".ptxas.log", "w") as f:
            f.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        objs = list(ex.map(compile_one, UNITS.items()))
    link = [nvcc, *ARCH, "-shared", "-cudart", "static", "-Xcompiler", "-fPIC", *objs, "-o", LIB]
    if verbose:
        print("[dcma_b200.build]", " ".join(link), flush=True)
    subprocess.check_call(link)
    json.dump({"sha256": digest, "arch": "sm_100a", "units": UNITS}, open(stamp, "w"), indent=1)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv))
