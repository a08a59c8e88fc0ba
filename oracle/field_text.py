"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's deformation-field text format.

Follows `deformation_field::write_to` / `read_from` (reference src/Alignment_Field.cc:322-436) line by line in plain
Python.  `os << val` at `precision(max_digits10)` is printf("%.17g") (libstdc++'s num_put calls vsnprintf with "%.*g");
`is >> val` is strtod.  CPython's `'%.17g' % x` and `float(s)` are the same correctly rounded conversions (David Gay's
dtoa), and `tests/native/text_format_check.cc` pins the product's conversion code directly against this C library's
printf/strtod, which is what the reference executes.  Pinned by the reference's own round-trip cases
(src/Alignment_Field_Tests.cc:196-289, restated in tests/test_field_text.py).

Only tests/ may import this module.
"""
from __future__ import annotations

import base64


def format_double(x: float) -> str:
    return "%.17g" % x


def format_doubles(values) -> bytes:
    return "".join("%.17g\n" % float(v) for v in values).encode()


def _vec(v) -> str:
    return "(%.17g, %.17g, %.17g)" % tuple(float(c) for c in v)


def write_field_text(images) -> bytes:
    """images: list of dicts with rows, cols, pxl_dx, pxl_dy, pxl_dz, anchor, offset, row_unit, col_unit, metadata (dict),
    data (flat sequence of rows*cols*3 doubles in (row, col, channel) order).  cc:322-357."""
    out = ["%d\n" % len(images)]
    for im in images:
        out.append("%d %d %d\n" % (im["rows"], im["cols"], 3))
        out.append("%.17g %.17g %.17g\n" % (im["pxl_dx"], im["pxl_dy"], im["pxl_dz"]))
        for key in ("anchor", "offset", "row_unit", "col_unit"):
            out.append(_vec(im[key]) + "\n")
        meta = sorted(im.get("metadata", {}).items())
        out.append("num_metadata= %d\n" % len(meta))
        for k, v in meta:
            out.append(base64.b64encode(k.encode()).decode() + " " + base64.b64encode(v.encode()).decode() + "\n")
        out.extend("%.17g\n" % float(v) for v in im["data"])
    return "".join(out).encode()


def read_field_text(text: bytes):
    """cc:359-436; raises ValueError where the reference returns false."""
    toks = text.split()
    pos = 0

    def take():
        nonlocal pos
        if pos >= len(toks):
            raise ValueError("unexpected end of field text")
        pos += 1
        return toks[pos - 1]

    def vec():
        parts = b" ".join(take() for _ in range(3)).strip(b"()").replace(b",", b" ").split()
        return tuple(float(p) for p in parts)

    n = int(take())
    if not 1 <= n <= 10000:
        raise ValueError("Number of images could not be read, or is invalid.")
    images = []
    for _ in range(n):
        rows, cols, ch = int(take()), int(take()), int(take())
        if rows <= 0 or cols <= 0 or ch != 3:
            raise ValueError("Image dimensions could not be read, or are invalid.")
        dx, dy, dz = float(take()), float(take()), float(take())
        anchor, offset, row_unit, col_unit = vec(), vec(), vec(), vec()
        if take() != b"num_metadata=":
            raise ValueError("Metadata count could not be read.")
        meta = {}
        for _m in range(int(take())):
            k, v = take(), take()
            meta[base64.b64decode(k).decode()] = base64.b64decode(v).decode()
        data = []
        for _v in range(rows * cols * 3):
            t = take()
            if not t.replace(b"+", b"").replace(b"-", b"").replace(b".", b"").replace(b"e", b"").replace(b"E", b"").isdigit():
                raise ValueError("Pixel data could not be read.")  # the stream extraction takes no inf/nan/hex
            data.append(float(t))
        images.append(dict(rows=rows, cols=cols, pxl_dx=dx, pxl_dy=dy, pxl_dz=dz, anchor=anchor, offset=offset,
                           row_unit=row_unit, col_unit=col_unit, metadata=meta, data=data))
    return images
