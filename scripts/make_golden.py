"""Generates tests/golden/*.npz from the UNMODIFIED reference build (oracle/_ref/libmimc3ref.so,
compiled from /root/reference by oracle/Makefile with the zeroing allocator).

Run in the container that has /root/reference:   python scripts/make_golden.py
The reference ships no golden vectors of its own, so these files pin the
oracle (tests/test_oracle_cpu.py) and the CUDA path (tests/test_golden_gpu.py) to outputs of
the reference itself.  Inputs are stored in the fixture (uint8/uint16 images, xyuvav), so the
tests do not depend on the generator's RNG.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from mimc3_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
VEC_OCW = (7, 15, 30, 40)


def multimatch(R, i0, i1, xy, dt, offset, H, W):
    pivs = [R.get_uv_pivot(xy, dt, ocw, H, W) for ocw in VEC_OCW]
    dps = []

    def attempts(a, b):
        for (off, piv), ocw in zip(pivs, VEC_OCW):
            o1, _ = R.match(a, b, xy, offset, off, piv, +1, ocw)
            o2, _ = R.match(b, a, xy, -offset, off, -piv, +1, ocw)   # main negates the pivots in place (MIMC_main.c:272-279)
            o2 = o2.copy(); o2[:, :2] = -o2[:, :2]                   # ... and the result (:289-293)
            dps.extend([o1, o2])
    attempts(i0, i1)
    filt = []
    c0 = np.zeros_like(i0); c1 = np.zeros_like(i1)
    for k in range(3):
        R.conv2(i0, k, c0); R.conv2(i1, k, c1)
        filt.append(c0.copy())
        attempts(c0, c1)
    return np.stack(dps), pivs, filt


def make(name, **scene_kw):
    sc = synth.make_scene(**scene_kw)
    store_dt = np.uint8 if sc.dtype == "u8" else np.uint16
    i0 = sc.i0.numpy(); i1 = sc.i1.numpy()
    assert np.array_equal(i0, i0.astype(store_dt)) and np.array_equal(i1, i1.astype(store_dt))
    H, W = i0.shape
    R = oracle.Reference()
    R.set_globals(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    offset = np.array(sc.offset, np.int32)
    dp, pivs, filt = multimatch(R, i0, i1, sc.xyuvav, sc.dt, offset, H, W)
    mvn, ncl = R.cluster(dp)
    stages = R.postprocess_stages(dp, sc.xyuvav, sc.dimx, sc.dimy)
    planes = R.postprocess(dp, sc.xyuvav, sc.dimx, sc.dimy)
    out = dict(i0=i0.astype(store_dt), i1=i1.astype(store_dt), xyuvav=sc.xyuvav, dimx=sc.dimx, dimy=sc.dimy, dt=np.float32(sc.dt),
               offset=offset, dp=dp, mvn=mvn, ncl=ncl, planes=planes,
               # filtered images: SHA-256 of the float32 bytes (no NaN survives the shift) + the top-left 96x96 crop
               conv2_i0_sha256=np.array([hashlib.sha256(f.tobytes()).hexdigest() for f in filt]),
               conv2_i0_crop=np.stack([f[:96, :96] for f in filt]))
    for k, v in stages.items():
        out["stage_" + k] = v
    for (off, piv), ocw in zip(pivs, VEC_OCW):
        out[f"piv_off_{ocw}"] = off
        out[f"piv_{ocw}"] = piv
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    fin = np.isfinite(dp[:, :, 0])
    print(f"{name}: {sc.n} nodes ({sc.dimy}x{sc.dimx}), {H}x{W} {sc.dtype}; finite dp {fin.mean():.3f}; "
          f"invalid(-3) {(dp[:, :, 2] == -3).mean():.3f}; nan planes {np.isnan(planes[0]).mean():.3f}; "
          f"{os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    oracle.build()
    # u8 pair with a no-data wedge and decorrelated patches (exercises null exclusion, invalid
    # nodes, hole filling and pseudosmoothing)
    make("ref_u8_wedge", H=448, W=448, dtype="u8", spacing=20, seed=7, peak_px=6.3, null_wedge=True, decorrelated_patches=6)
    # u16 pair, faster band (longer pivot lines, float-rounded products)
    make("ref_u16_fast", H=420, W=420, dtype="u16", spacing=29, seed=11, peak_px=14.0, apriori_gain=0.9, null_wedge=False,
         band_width_frac=0.2)
