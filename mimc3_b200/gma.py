"""GMA matrix files: the reference CLI's on-disk format (GMA.c:166-244 loaders, :319-424 savers;
reader tooling gma.py:3-21).  Layout: int32 nrows, int32 ncols, then the row-major payload in
the element type the caller names (the file does not record it).

    read(src, dtype)   src = path, open binary file, tarfile member object or bytes
    write(dst, arr)    dst = path or open binary file; 1-D arrays are written as one row
"""
from __future__ import annotations

import io
import os

import numpy as np

# element types of the files the driver writes (MIMC_main.c:428-435) and reads (:229-236)
FIELD_DTYPE = {
    "x": "float64", "y": "float64", "vx": "float32", "vy": "float32", "ex": "float32", "ey": "float32",
    "qual": "float32", "flagcp": "uint8", "xyuvav": "float64", "dp": "float32",
}


def _payload(src) -> bytes:
    if isinstance(src, (bytes, bytearray, memoryview)):
        return bytes(src)
    if isinstance(src, (str, os.PathLike)):
        with open(src, "rb") as f:
            return f.read()
    if hasattr(src, "read"):
        return src.read()
    raise TypeError(f"gma.read: cannot read a GMA matrix from {type(src).__name__}")


def read(src, dtype="float32") -> np.ndarray:
    """-> writable (nrows, ncols) array.  Raises ValueError when the payload size does not match the
    header for this dtype (the reference's reader reshapes blindly)."""
    raw = _payload(src)
    if len(raw) < 8:
        raise ValueError("gma.read: shorter than the 8-byte header")
    nrows, ncols = (int(v) for v in np.frombuffer(raw, dtype="<i4", count=2))
    dt = np.dtype(dtype).newbyteorder("<")
    if nrows < 0 or ncols < 0 or len(raw) - 8 != nrows * ncols * dt.itemsize:
        raise ValueError(f"gma.read: header says {nrows} x {ncols} {dt.name} but the payload is {len(raw) - 8} bytes")
    return np.frombuffer(raw, dtype=dt, offset=8).reshape(nrows, ncols).astype(np.dtype(dtype), copy=True)


def write(dst, arr) -> None:
    a = np.asarray(arr)
    if a.ndim == 1:
        a = a[None, :]
    if a.ndim != 2:
        raise ValueError("gma.write: GMA files hold 2-D matrices")
    a = np.ascontiguousarray(a.astype(a.dtype.newbyteorder("<"), copy=False))
    head = np.array(a.shape, dtype="<i4").tobytes()
    if isinstance(dst, (str, os.PathLike)):
        with open(dst, "wb") as f:
            f.write(head); f.write(a.tobytes())
    else:
        dst.write(head); dst.write(a.tobytes())


def dumps(arr) -> bytes:
    b = io.BytesIO()
    write(b, arr)
    return b.getvalue()


# ---- the dp dump / reload mode of MIMC_main_test_postprocessing.c:262-284 ---------------------------

def save_dp(directory: str, dp, flag_cp=None) -> None:
    """dp (num_dp, n, 3) -> <directory>/FT_result/dp_%02d.gma (+ flag_cp.gma), the layout the reference's
    postprocess-only driver reloads."""
    d = os.path.join(directory, "FT_result")
    os.makedirs(d, exist_ok=True)
    dp = np.asarray(dp, dtype=np.float32)
    for k in range(dp.shape[0]):
        write(os.path.join(d, f"dp_{k:02d}.gma"), dp[k])
    if flag_cp is not None:
        write(os.path.join(d, "flag_cp.gma"), np.asarray(flag_cp, dtype=np.uint8).reshape(-1, 1))


def load_dp(directory: str, num_dp: int = 32):
    """-> (dp (num_dp, n, 3) float32, flag_cp (n,) uint8 or None)."""
    d = os.path.join(directory, "FT_result")
    mats = [read(os.path.join(d, f"dp_{k:02d}.gma"), "float32") for k in range(num_dp)]
    shape = mats[0].shape
    if shape[1] != 3 or any(m.shape != shape for m in mats):
        raise ValueError("load_dp: every dp_NN.gma must be n x 3 with the same n")
    fc = os.path.join(d, "flag_cp.gma")
    flag = read(fc, "uint8").ravel() if os.path.exists(fc) else None
    return np.stack(mats), flag
