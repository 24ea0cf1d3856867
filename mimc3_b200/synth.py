"""Seeded synthetic image pairs with a known shift field, plus the matching xyuvav grid.

The reference ships no sample data (its README.md:32 points at an unreachable Google
Drive link), so every test and benchmark input is generated here, following the recipe
in SURVEY.md section 8(d):

* texture = Gaussian-blurred white noise at sigma 1.5 / 5 / 15 px (weights 1, 0.7, 0.5),
  min-max scaled to DN [3, 253] (uint8) or [64, 16383] (uint16) so that there are no
  accidental zero ("null", MIMC_module.c:723) pixels;
* image 1 = image 0 resampled (bicubic) through a smooth shift field
  ``(du, dv)(x, y)`` = rigid integer offset + Gaussian-profile "glacier" band, plus
  N(0, sigma_noise) noise;
* optional zero-valued no-data wedge (null exclusion) and decorrelated patches
  (to force low-support clusters so that pseudosmoothing has work to do);
* xyuvav rows = ``[x, y, u, v, vx_apriori, vy_apriori]`` (reference README.md:36,
  MIMC_main.c:203-223), row-major with u fastest, a-priori = ``apriori_gain`` x truth.

Input preconditions of the reference that the generator respects:
nodes >= 64 px from every image edge, >= 50 slow (< 10 m/yr) nodes, grid width >= 2.

All heavy lifting is torch so that the 16384^2 benchmark pair can be produced on the GPU
in seconds; on CPU the same code serves the small test cases.
"""
from __future__ import annotations

import dataclasses
import math
import struct

import numpy as np
import torch
import torch.nn.functional as F

from . import gma


@dataclasses.dataclass
class Scene:
    """One synthetic image pair + node grid. Images are float32 (integer valued)."""

    i0: torch.Tensor            # (H, W) float32, DN values
    i1: torch.Tensor            # (H, W) float32
    xyuvav: np.ndarray          # (n, 6) float64
    dimx: int                   # nodes per grid row
    dimy: int
    dt: float                   # days
    mpp: float                  # metres per pixel
    spacing: int                # node spacing, pixels
    offset: tuple               # rigid integer (du, dv) between the images
    truth_du: np.ndarray        # (dimy, dimx) true total displacement at the nodes, px
    truth_dv: np.ndarray
    dtype: str                  # "u8" | "u16"

    @property
    def n(self) -> int:
        return self.dimx * self.dimy

    @property
    def shape(self):
        return tuple(self.i0.shape)


def _gauss_kernel1d(sigma: float, device) -> torch.Tensor:
    r = max(1, int(math.ceil(4.0 * sigma)))
    x = torch.arange(-r, r + 1, dtype=torch.float32, device=device)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _blur(img: torch.Tensor, sigma: float) -> torch.Tensor:
    """Separable Gaussian blur with reflect padding; img is (H, W)."""
    k = _gauss_kernel1d(sigma, img.device)
    r = (k.numel() - 1) // 2
    x = img[None, None]
    x = F.conv2d(F.pad(x, (r, r, 0, 0), mode="reflect"), k.view(1, 1, 1, -1))
    x = F.conv2d(F.pad(x, (0, 0, r, r), mode="reflect"), k.view(1, 1, -1, 1))
    return x[0, 0]


def make_texture(H: int, W: int, seed: int, device="cpu") -> torch.Tensor:
    """Unit-range texture in [0, 1], (H, W) float32."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)   # the stream depends on the device type
    out = torch.zeros(H, W, dtype=torch.float32, device=dev)
    for sigma, wgt in ((1.5, 1.0), (5.0, 0.7), (15.0, 0.5)):
        noise = torch.randn(H, W, generator=g, dtype=torch.float32, device=dev)
        b = _blur(noise, sigma)
        b = (b - b.mean()) / b.std()
        out += wgt * b
        del noise, b
    lo, hi = out.min(), out.max()
    return (out - lo) / (hi - lo)


def _band_field(H, W, device, centre_frac, width_px, peak_px, angle_deg, rows=None):
    """Gaussian-profile flow band: displacement along `angle_deg`, magnitude
    peak_px * exp(-(d/width)^2), d = distance from the band's centre line (which runs
    along the flow direction through (centre_frac*W, centre_frac*H))."""
    th = math.radians(angle_deg)
    ys = torch.arange(H, dtype=torch.float32, device=device) if rows is None else rows
    xs = torch.arange(W, dtype=torch.float32, device=device)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    cx, cy = centre_frac[0] * W, centre_frac[1] * H
    d = -(xx - cx) * math.sin(th) + (yy - cy) * math.cos(th)
    mag = peak_px * torch.exp(-((d / width_px) ** 2))
    return mag * math.cos(th), mag * math.sin(th)


def make_scene(
    H: int = 2048,
    W: int = 2048,
    dtype: str = "u8",
    spacing: int = 19,
    margin: int = 64,
    seed: int = 1234,
    peak_px: float = 6.3,
    band_angle_deg: float = 30.0,
    band_width_frac: float = 0.12,
    offset=(2, -1),
    noise_dn: float = 1.0,
    apriori_gain: float = 0.8,
    background_mpy: float = 3.0,
    null_wedge: bool = False,
    decorrelated_patches: int = 0,
    mpp: float = 15.0,
    dt: float = 16.0,
    device="cpu",
    max_nodes_xy=None,
) -> Scene:
    dev = torch.device(device)
    tex = make_texture(H, W, seed, dev)
    lo, hi = (3.0, 253.0) if dtype == "u8" else (64.0, 16383.0)
    i0 = torch.round(tex * (hi - lo) + lo)
    del tex

    # shift field at every pixel (band) + rigid offset
    du, dv = _band_field(H, W, dev, (0.5, 0.5), band_width_frac * min(H, W), peak_px, band_angle_deg)
    # i1(q) = i0(q - d(q)): bicubic resampling through grid_sample
    ys = torch.arange(H, dtype=torch.float32, device=dev)
    xs = torch.arange(W, dtype=torch.float32, device=dev)
    yy, xx = torch.meshgrid(ys, xs, indexing="ij")
    sx = xx - (du + offset[0])
    sy = yy - (dv + offset[1])
    del xx, yy
    grid = torch.stack(((sx + 0.5) / W * 2 - 1, (sy + 0.5) / H * 2 - 1), dim=-1)[None]
    del sx, sy
    i1 = F.grid_sample(i0[None, None], grid, mode="bicubic", padding_mode="reflection", align_corners=False)[0, 0]
    del grid
    g = torch.Generator(device=dev).manual_seed(seed + 1)
    if noise_dn > 0:
        # noise generated in row blocks to bound memory on the big configs
        blk = 2048
        for r0 in range(0, H, blk):
            r1 = min(H, r0 + blk)
            i1[r0:r1] += noise_dn * torch.randn(r1 - r0, W, generator=g, dtype=torch.float32, device=dev)
    i1 = torch.clamp(torch.round(i1), lo, hi if dtype == "u8" else 65535.0)

    rng = np.random.default_rng(seed + 2)
    if decorrelated_patches > 0:
        # independent texture pasted into i1 so that matches there decorrelate
        alt = torch.round(make_texture(256, 256, seed + 77, dev) * (hi - lo) + lo)
        for _ in range(decorrelated_patches):
            ph, pw = int(rng.integers(40, 160)), int(rng.integers(40, 160))
            py, px = int(rng.integers(margin, H - margin - ph)), int(rng.integers(margin, W - margin - pw))
            i1[py:py + ph, px:px + pw] = alt[:ph, :pw]

    # node grid
    us = np.arange(margin, W - margin, spacing, dtype=np.int64)
    vs = np.arange(margin, H - margin, spacing, dtype=np.int64)
    if max_nodes_xy is not None:
        us, vs = us[: max_nodes_xy[0]], vs[: max_nodes_xy[1]]
    dimx, dimy = len(us), len(vs)
    uu, vv = np.meshgrid(us, vs)  # (dimy, dimx), u fastest
    vi, ui = torch.from_numpy(vv).to(dev), torch.from_numpy(uu).to(dev)
    du_n = du[vi, ui].cpu().numpy().astype(np.float64)
    dv_n = dv[vi, ui].cpu().numpy().astype(np.float64)
    del du, dv

    if null_wedge:
        # zero wedge (no-data) inside the fast band, kept > 45 px away from slow nodes
        # (SURVEY.md H11): rows/cols around the band centre only.
        cy, cx = H // 2, W // 2
        hw = max(24, min(H, W) // 24)
        i0[cy - hw:cy + hw, cx - 3 * hw:cx + 3 * hw] = 0.0
        i1[cy - hw:cy + hw, cx - 3 * hw:cx + 3 * hw] = 0.0

    to_mpy = mpp / dt * 365.0
    bg = background_mpy
    th = math.radians(band_angle_deg)
    vx = apriori_gain * du_n * to_mpy + bg * math.cos(th)
    vy = -(apriori_gain * dv_n * to_mpy + bg * math.sin(th))
    x0, y0 = 500000.0, 7000000.0
    xy = np.empty((dimx * dimy, 6), dtype=np.float64)
    xy[:, 0] = (x0 + uu * mpp).ravel()
    xy[:, 1] = (y0 - vv * mpp).ravel()
    xy[:, 2] = uu.ravel()
    xy[:, 3] = vv.ravel()
    xy[:, 4] = vx.ravel()
    xy[:, 5] = vy.ravel()
    return Scene(i0=i0.contiguous(), i1=i1.contiguous(), xyuvav=xy, dimx=dimx, dimy=dimy, dt=dt, mpp=mpp,
                 spacing=spacing, offset=tuple(offset), truth_du=du_n + offset[0], truth_dv=dv_n + offset[1],
                 dtype=dtype)


# ---------------------------------------------------------------------------------------
# file formats of the reference's CLI: GMA matrices and strip TIFFs
# ---------------------------------------------------------------------------------------

def write_gma(path: str, arr: np.ndarray) -> None:
    gma.write(path, arr)


def read_gma(path: str, dtype="float32") -> np.ndarray:
    return gma.read(path, dtype)


def write_tiff(path: str, img: np.ndarray) -> None:
    """Baseline little-endian single-strip grayscale TIFF, 8 or 16 bit (what
    GMA_float_load_tiff, GMA.c:246-316, reads scan-line by scan-line)."""
    a = np.ascontiguousarray(img)
    assert a.dtype in (np.uint8, np.uint16) and a.ndim == 2
    H, W = a.shape
    bits = a.dtype.itemsize * 8
    data = a.astype("<u%d" % a.dtype.itemsize).tobytes()
    entries = [
        (256, 4, 1, W), (257, 4, 1, H), (258, 3, 1, bits), (259, 3, 1, 1), (262, 3, 1, 1),
        (273, 4, 1, 8), (277, 3, 1, 1), (278, 4, 1, H), (279, 4, 1, len(data)),
    ]
    ifd_off = 8 + len(data) + (len(data) & 1)
    with open(path, "wb") as f:
        f.write(struct.pack("<2sHI", b"II", 42, ifd_off))
        f.write(data)
        if len(data) & 1:
            f.write(b"\0")
        f.write(struct.pack("<H", len(entries)))
        for tag, typ, cnt, val in entries:
            if typ == 3:
                f.write(struct.pack("<HHIHH", tag, typ, cnt, val, 0))
            else:
                f.write(struct.pack("<HHII", tag, typ, cnt, val))
        f.write(struct.pack("<I", 0))
