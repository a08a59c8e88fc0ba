#!/usr/bin/env python3
"""Write patches/Alignment_Field.patch: the two per-value loops of `deformation_field::write_to` / `read_from` in the
reference's src/Alignment_Field.cc replaced by calls into the payload codec (dicomautomaton_b200/host/Field_Payload.h),
as a unified diff.  Needs /root/reference; the patch itself is what is committed.

    python tools/make_field_patch.py            (re-run when the reference file changes)

Edits are anchored on text, not line numbers.  Everything else in the file -- header lines, metadata, warnings, return
values -- stays the reference's own code."""
import difflib
import os

REF = os.environ.get("DCMA_REFERENCE_SRC", "/root/reference/src")
SRC = os.path.join(REF, "Alignment_Field.cc")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "patches", "Alignment_Field.patch")

EDITS = [
    ('#include "Alignment_Field.h"\n',
     '#include "Alignment_Field.h"\n#include "Field_Payload.h"     //B200 codec for the pixel values of a stored field.\n'),
    ('        for(const auto &val : img.data){\n            os << val << \'\\n\';\n        }\n',
     '        if(!dcma_b200_io::write_payload(os, img.data.data(), static_cast<int64_t>(img.data.size()))){\n'
     '            os.precision( original_precision );\n            return false;\n        }\n'),
    ('            for(auto &val : img.data){\n                is >> val;\n            }\n',
     '            dcma_b200_io::read_payload(is, img.data.data(), static_cast<int64_t>(img.data.size()));\n'),
]


def patched(text: str) -> str:
    for old, new in EDITS:
        if text.count(old) != 1:
            raise SystemExit(f"anchor not found exactly once: {old!r}")
        text = text.replace(old, new)
    return text


def main():
    if not os.path.isfile(SRC):
        raise SystemExit(f"{SRC} not found")
    old = open(SRC).read()
    diff = difflib.unified_diff(old.splitlines(True), patched(old).splitlines(True), "a/src/Alignment_Field.cc",
                                "b/src/Alignment_Field.cc", n=3)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.writelines(diff)
    print(OUT)


if __name__ == "__main__":
    main()
