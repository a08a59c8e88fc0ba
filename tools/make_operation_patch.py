#!/usr/bin/env python3
"""Write patches/RegisterImagesDemons.patch: the additive operation arguments of the B200 drop-in (FieldMode,
PyramidLevels, PyramidIterations, UseSymmetricForce, Devices) as a unified diff against the reference's
src/Operations/RegisterImagesDemons.cc.  Needs /root/reference; the patch itself is what is committed.

    python tools/make_operation_patch.py            (re-run when the reference file changes)

The insertions are made next to textual anchors, not line numbers.  Defaults keep the reference's behaviour: fp64 fields,
one resolution level, fixed-image gradient only, one device."""
import difflib
import os
import sys

REF = os.environ.get("DCMA_REFERENCE_SRC", "/root/reference/src")
SRC = os.path.join(REF, "Operations", "RegisterImagesDemons.cc")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "patches", "RegisterImagesDemons.patch")

DOC = '''
    out.args.emplace_back();
    out.args.back().name = "FieldMode";
    out.args.back().desc = "Storage and arithmetic of the update and deformation fields on the GPU."
                           " 'fp64' keeps the reference's double fields (FMA contraction allowed),"
                           " 'fp64_strict' reproduces the CPU build bit for bit,"
                           " 'mixed' keeps the deformation field and all arithmetic in double and stores only the"
                           " per-iteration update field as float (within 1e-6 mm of the reference after 200 iterations),"
                           " 'fp32' stores float fields (fastest; ~1e-3 mm from the reference after 200 iterations).";
    out.args.back().default_val = "fp64";
    out.args.back().expected = true;
    out.args.back().examples = { "fp64", "fp64_strict", "mixed", "fp32" };
    out.args.back().samples = OpArgSamples::Exhaustive;

    out.args.emplace_back();
    out.args.back().name = "PyramidLevels";
    out.args.back().desc = "Number of multi-resolution levels (coarse to fine, factor 2 per axis)."
                           " 1 runs the single-resolution algorithm.";
    out.args.back().default_val = "1";
    out.args.back().expected = true;
    out.args.back().examples = { "1", "2", "3" };

    out.args.emplace_back();
    out.args.back().name = "PyramidIterations";
    out.args.back().desc = "Comma-separated iteration counts for the COARSER pyramid levels, from the second-finest"
                           " to the coarsest; the finest level runs MaxIterations."
                           " Levels without an entry (or with 0) use MaxIterations too.";
    out.args.back().default_val = "";
    out.args.back().expected = false;
    out.args.back().examples = { "50,100" };

    out.args.emplace_back();
    out.args.back().name = "UseSymmetricForce";
    out.args.back().desc = "Whether to use the symmetric demons force, which averages the gradients of the fixed and the"
                           " warped moving image, instead of the fixed-image gradient alone.";
    out.args.back().default_val = "false";
    out.args.back().expected = true;
    out.args.back().examples = { "true", "false" };
    out.args.back().samples = OpArgSamples::Exhaustive;

    out.args.emplace_back();
    out.args.back().name = "Devices";
    out.args.back().desc = "Number of GPUs to shard the volume over (contiguous slabs along the slice direction,"
                           " halo planes exchanged every iteration). 1 keeps the whole volume on one GPU.";
    out.args.back().default_val = "1";
    out.args.back().expected = true;
    out.args.back().examples = { "1", "2", "4", "8" };
'''

PARSE = '''    const auto FieldModeStr = OptArgs.getValueStr("FieldMode").value_or("fp64");
    const auto PyramidLevels = std::stol(OptArgs.getValueStr("PyramidLevels").value_or("1"));
    const auto PyramidIterationsOpt = OptArgs.getValueStr("PyramidIterations");
    const auto UseSymmetricForceStr = OptArgs.getValueStr("UseSymmetricForce").value_or("false");
    const auto Devices = std::stol(OptArgs.getValueStr("Devices").value_or("1"));
'''

FILL = '''    if(FieldModeStr == "fp64"){ params.field_mode = 0;
    }else if(FieldModeStr == "fp64_strict"){ params.field_mode = 1;
    }else if(FieldModeStr == "fp32"){ params.field_mode = 2;
    }else if(FieldModeStr == "mixed"){ params.field_mode = 3;
    }else{ throw std::invalid_argument("FieldMode not understood. Cannot continue."); }
    if((PyramidLevels < 1) || (3 < PyramidLevels)) throw std::invalid_argument("PyramidLevels must be 1, 2 or 3.");
    params.pyramid_levels = static_cast<int32_t>(PyramidLevels);
    if(PyramidIterationsOpt){
        const auto toks = SplitStringToVector(PyramidIterationsOpt.value(), ',', 'd');
        for(size_t i = 0; (i < toks.size()) && (i < 3); ++i){
            if(!toks[i].empty()) params.pyramid_iterations[i] = static_cast<int32_t>(std::stol(toks[i]));
        }
    }
    params.use_symmetric_force = std::regex_match(UseSymmetricForceStr, regex_true);
    if(Devices < 1) throw std::invalid_argument("Devices must be at least 1.");
    params.n_devices = static_cast<int32_t>(Devices);
'''


def patched(text: str) -> str:
    a1 = '    out.args.emplace_back();\n    out.args.back().name = "TransformName";'
    a2 = '    const auto TransformName = OptArgs.getValueStr("TransformName").value();'
    a3 = '    params.verbosity = Verbosity;\n'
    for a in (a1, a2, a3):
        if text.count(a) != 1:
            raise SystemExit(f"anchor not found exactly once: {a!r}")
    text = text.replace(a1, DOC.lstrip("\n") + "\n" + a1)
    text = text.replace(a2, PARSE + a2)
    text = text.replace(a3, a3 + FILL)
    return text


def main():
    if not os.path.isfile(SRC):
        raise SystemExit(f"{SRC} not found")
    old = open(SRC).read()
    new = patched(old)
    diff = difflib.unified_diff(old.splitlines(True), new.splitlines(True), "a/src/Operations/RegisterImagesDemons.cc",
                                "b/src/Operations/RegisterImagesDemons.cc", n=3)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.writelines(diff)
    print(OUT)


if __name__ == "__main__":
    main()
