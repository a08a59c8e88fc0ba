"""Helpers shared by the tests (inputs only; no checking logic)."""
import numpy as np


def phantom_pair(shape, spacing=(1.0, 1.0, 1.0), case=0, amplitude_mm=2.0):
    from dicomautomaton_b200 import phantom
    f, m = phantom.make_pair(shape, spacing=spacing, case=case, amplitude_mm=amplitude_mm)
    return f.numpy(), m.numpy()


def random_smooth_field(shape, max_mm=3.0, seed=0):
    """Smooth random displacement field (slices, rows, cols, 3) in mm with |D| <= max_mm."""
    rng = np.random.default_rng(seed)
    s, r, c = shape
    zz, yy, xx = np.meshgrid(np.arange(s), np.arange(r), np.arange(c), indexing="ij")
    out = np.zeros(shape + (3,), dtype=np.float64)
    for comp in range(3):
        acc = np.zeros(shape, dtype=np.float64)
        for _ in range(4):
            k = rng.uniform(0.05, 0.6, size=3)
            ph = rng.uniform(0, 2 * np.pi, size=3)
            acc += rng.uniform(-1, 1) * np.sin(k[0] * xx + ph[0]) * np.cos(k[1] * yy + ph[1]) * np.cos(k[2] * zz + ph[2])
        out[..., comp] = acc
    out *= max_mm / max(np.abs(out).max(), 1e-12)
    return out


def ref_kind(oracle_mod):
    """The compiled reference engine (oracle/_ref, built from /root/reference by oracle/build_ref.py and shipped to the
    GPU box as a .so) whenever it is present, else the C restatement that tests/test_oracle.py pins to it bit for bit."""
    return "reference" if oracle_mod.available("reference") else "port"
