"""The `.trans` text form of a deformation field with its numbers converted on a B200.

Mirrors `deformation_field::write_to` / `read_from` (reference src/Alignment_Field.cc:322-436): per image a few header
lines (extent, spacing, anchor, offset, row/column unit vectors, base64 metadata) written here on the host, then every
double of the image on its own line at max_digits10 -- that part, all of the volume of the file, goes through the C ABI
(`dcma_b200_format_doubles` / `dcma_b200_parse_doubles`).  No CPU fallback: without a device these raise `DcmaError`.
"""
from __future__ import annotations

import base64
import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import capi


def format_doubles(values: np.ndarray, device: int = -1) -> bytes:
    """printf("%.17g\\n") of every element, in order (what `os << val << '\\n'` at max_digits10 writes, cc:346-348)."""
    v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
    buf = np.empty(max(v.size * 25, 1), dtype=np.uint8)
    n_out = C.c_int64(0)
    err = C.create_string_buffer(512)
    capi.check(capi.lib().dcma_b200_format_doubles(v.ctypes.data, v.size, buf.ctypes.data, buf.size, C.byref(n_out), device,
                                                   err, len(err)), err)
    return buf[:n_out.value].tobytes()


def parse_doubles(text: bytes, n: int, device: int = -1) -> Tuple[np.ndarray, int]:
    """The first n white-space separated numbers of `text` as n stream extractions read them (cc:418-420), and the offset
    just past the last one."""
    src = np.frombuffer(text, dtype=np.uint8)
    out = np.empty(int(n), dtype=np.float64)
    used = C.c_int64(0)
    err = C.create_string_buffer(512)
    capi.check(capi.lib().dcma_b200_parse_doubles(src.ctypes.data if src.size else None, src.size, int(n), out.ctypes.data,
                                                  C.byref(used), device, err, len(err)), err)
    return out, int(used.value)


def _g(x: float) -> bytes:
    return b"%.17g" % float(x)


def _vec(v) -> bytes:
    return b"(" + _g(v[0]) + b", " + _g(v[1]) + b", " + _g(v[2]) + b")"


def write_field_text(fileobj, deformation: np.ndarray, grid: capi.Grid, metadata: Optional[Dict[str, str]] = None,
                     device: int = -1) -> None:
    """Write `deformation` (slices, rows, cols, 3; mm along the grid axes) as the reference's field text: one image per
    slice with anchor 0, offset = origin + z * pxl_dz * ez, row_unit = ex, col_unit = ey, and `metadata` on every image."""
    D = np.ascontiguousarray(deformation, dtype=np.float64)
    if D.ndim != 4 or D.shape[3] != 3:
        raise ValueError("deformation must be (slices, rows, cols, 3)")
    S, R, Cc, _ = D.shape
    g = grid.geom
    if (S, R, Cc) != (g.slices, g.rows, g.cols):
        raise ValueError("deformation and grid extents differ")
    origin, ex, ey, ez = (np.array(list(a)) for a in (grid.origin, grid.ex, grid.ey, grid.ez))
    meta = sorted((metadata or {}).items())  # std::map order
    fileobj.write(b"%d\n" % S)
    for z in range(S):
        off = origin + z * g.pxl_dz * ez
        head = [b"%d %d %d\n" % (R, Cc, 3), _g(g.pxl_dx) + b" " + _g(g.pxl_dy) + b" " + _g(g.pxl_dz) + b"\n",
                _vec((0.0, 0.0, 0.0)) + b"\n", _vec(off) + b"\n", _vec(ex) + b"\n", _vec(ey) + b"\n",
                b"num_metadata= %d\n" % len(meta)]
        for k, v in meta:
            head.append(base64.b64encode(k.encode()) + b" " + base64.b64encode(v.encode()) + b"\n")
        fileobj.write(b"".join(head))
        fileobj.write(format_doubles(D[z], device))


def read_field_text(fileobj, device: int = -1):
    """Inverse of `write_field_text` for any field text the reference writes.  Returns (deformation (slices, rows, cols, 3),
    images) where images[z] is a dict with rows, cols, pxl_dx/dy/dz, anchor, offset, row_unit, col_unit, metadata."""
    text = fileobj.read()
    pos = 0

    def token() -> bytes:
        nonlocal pos
        n = len(text)
        while pos < n and text[pos:pos + 1].isspace():
            pos += 1
        b = pos
        while pos < n and not text[pos:pos + 1].isspace():
            pos += 1
        if b == pos:
            raise ValueError("unexpected end of field text")
        return text[b:pos]

    def vec():
        t = b" ".join(token() for _ in range(3)).strip(b"()").replace(b",", b" ").split()
        return tuple(float(x) for x in t)

    n_imgs = int(token())
    if not 1 <= n_imgs <= 10000:
        raise ValueError("Number of images could not be read, or is invalid.")
    images: List[dict] = []
    planes = []
    for _ in range(n_imgs):
        rows, cols, ch = int(token()), int(token()), int(token())
        if rows <= 0 or cols <= 0 or ch != 3:
            raise ValueError("Image dimensions could not be read, or are invalid.")
        dx, dy, dz = float(token()), float(token()), float(token())
        anchor, offset, row_unit, col_unit = vec(), vec(), vec(), vec()
        if token() != b"num_metadata=":
            raise ValueError("Metadata count could not be read.")
        meta = {}
        for _m in range(int(token())):
            k, v = token(), token()
            meta[base64.b64decode(k).decode()] = base64.b64decode(v).decode()
        vals, used = parse_doubles(text[pos:], rows * cols * 3, device)
        pos += used
        planes.append(vals.reshape(rows, cols, 3))
        images.append(dict(rows=rows, cols=cols, pxl_dx=dx, pxl_dy=dy, pxl_dz=dz, anchor=anchor, offset=offset,
                           row_unit=row_unit, col_unit=col_unit, metadata=meta))
    same = all(p.shape == planes[0].shape for p in planes)
    return (np.stack(planes) if same else planes), images


_MAGIC = b"DCMA_TRANSFORM"
_VARIANT = b"TRANSFORM_VARIANT_DEFORMATION_FIELD_3"


def write_trans_file(path, deformation: np.ndarray, grid: capi.Grid, metadata: Optional[Dict[str, str]] = None,
                     device: int = -1) -> None:
    """A `.trans` file as `ExportWarps` -> `WriteTransform3` writes it for a deformation field (reference
    src/Transformation_File_Loader.cc:112-149): magic line, variant line, then the field text.  The transform-level
    `# key = value` comment lines are omitted (their encoding is Ygor's; `ReadTransform3` treats them as optional)."""
    with open(path, "wb") as f:
        f.write(_MAGIC + b"\n" + _VARIANT + b"\n")
        write_field_text(f, deformation, grid, metadata, device)


def read_trans_file(path, device: int = -1):
    """Inverse of `write_trans_file`; follows `ReadTransform3` (src/Transformation_File_Loader.cc:23-100): comment lines
    and blank lines are skipped, the first other line must be the magic phrase, the second names the variant."""
    with open(path, "rb") as f:
        magic = variant = None
        while variant is None:
            line = f.readline()
            if not line:
                raise ValueError("not a DCMA_TRANSFORM file")
            if b"#" in line or not line.strip():
                continue
            if magic is None:
                magic = line.strip()
                if magic != _MAGIC:
                    raise ValueError("not a DCMA_TRANSFORM file")
            else:
                variant = line.strip()
        if variant != _VARIANT:
            raise ValueError("transform variant %r is not a deformation field" % variant.decode(errors="replace"))
        return read_field_text(f, device)
