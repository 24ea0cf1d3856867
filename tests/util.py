"""Shared helpers for the parity tests."""
import numpy as np

from mimc3_b200 import synth

VEC_OCW = (7, 15, 30, 40)


def small_scene(**kw):
    args = dict(H=640, W=640, dtype="u8", spacing=23, seed=5, peak_px=6.3, null_wedge=True)
    args.update(kw)
    return synth.make_scene(**args)


def same_bits_nan_aware(a, b):
    """Bit-exact comparison where any NaN equals any NaN (x86 sqrt(-1) gives 0xFFC00000, CUDA 0x7FFFFFFF)."""
    a = np.asarray(a); b = np.asarray(b)
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return np.array_equal(a[~na].view(np.uint32), b[~nb].view(np.uint32))


def mismatch_report(a, b, name=""):
    a = np.asarray(a); b = np.asarray(b)
    na, nb = np.isnan(a), np.isnan(b)
    bad = (na != nb) | (~na & ~nb & (a != b))
    idx = np.argwhere(bad)
    return f"{name}: {bad.sum()} mismatching entries, first {idx[:5].tolist()}: {a[bad][:5]} vs {b[bad][:5]}"
