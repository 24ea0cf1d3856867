"""Host-side mirror of the reference driver's flow over the C ABI (MIMC_main.c:229-402):
load the pair -> (CP offset) -> pivots -> 32-attempt multi-match -> postprocess -> finalize.

PyTorch is used for device memory only; all compute goes through libmimc3cu.so.
"""
from __future__ import annotations

import numpy as np
import torch

from . import lib

VEC_OCW = (7, 15, 30, 40)   # MIMC_main.c:134-137
# the three filter kernels the driver builds at MIMC_main.c:175-196 (d/dx, d/dy, Laplacian)
FILTERS = (
    np.array([[-1, 0, 1]], dtype=np.float32),
    np.array([[-1], [0], [1]], dtype=np.float32),
    np.array([[-0.125, -0.125, -0.125], [-0.125, 1.0, -0.125], [-0.125, -0.125, -0.125]], dtype=np.float32),
)


class Pipeline:
    def __init__(self, device: int = 0):
        self.ctx = lib.Context(device)
        self.device = torch.device("cuda", device)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.device)
        self.handles = {}
        self.n = 0
        self.params = None
        self.H = self.W = 0

    def close(self):
        self.ctx.close()

    # ---- inputs ---------------------------------------------------------------------------
    def _ensure_images(self, H, W):
        if (H, W) != (self.H, self.W):
            for h in self.handles.values():
                self.ctx.image_destroy(h)
            self.handles = {k: self.ctx.image_create(H, W) for k in ("i0", "i1", "i0c", "i1c")}
            self.H, self.W = H, W

    def set_images(self, i0, i1):
        """i0/i1: numpy (float32 / uint8 / uint16, host) or CUDA float32 torch tensors."""
        H, W = i0.shape
        self._ensure_images(H, W)
        for key, img in (("i0", i0), ("i1", i1)):
            if isinstance(img, np.ndarray):
                self.ctx.image_upload(self.handles[key], img)
            elif img.is_cuda:
                torch.cuda.current_stream(img.device).synchronize()
                self.ctx._ck(self.ctx.L.mimc3cu_image_copy_from_device(self.ctx.h, self.handles[key], img.contiguous().data_ptr()))
            else:
                self.ctx.image_upload(self.handles[key], img.numpy())

    def set_grid(self, xyuvav, dimx, dimy, dt):
        """Nodes + the four pivot sets (get_uv_pivot per chip size, MIMC_main.c:264)."""
        x = np.ascontiguousarray(xyuvav, dtype=np.float64)
        self.params = lib.params_for(x, dimx, dimy, dt)
        self.xyuvav = x
        self.n = x.shape[0]
        self.ctx.set_nodes(x)
        self.pivot_bytes = 0
        for slot, ocw in enumerate(VEC_OCW):
            off, piv = lib.get_uv_pivot(x, dt, self.params.mpp, ocw, self.H, self.W, self.params.AW_SF, self.params.AW_CRE)
            self.ctx.set_pivots(slot, off, piv)
            self.pivot_bytes += off.nbytes + piv.nbytes

    def match_all(self, xyuvav, dimx, dimy, dt, offset, want_ncell=False):
        """set_grid + multimatch with the host work hidden behind the GPU: the pivots of chip size k+1
        are generated (host threads) and uploaded while the GPU runs the raw-pair attempts of chip size k;
        the filtered variants follow.  Same launches in the same stream order, hence the same results, as
        ``set_grid(); multimatch()`` (MIMC_main.c:261-350) -> dp (32, n, 3) on the device."""
        x = np.ascontiguousarray(xyuvav, dtype=np.float64)
        self.params = lib.params_for(x, dimx, dimy, dt)
        self.xyuvav = x
        self.n = n = x.shape[0]
        self.ctx.set_nodes(x)
        off_f = np.ascontiguousarray(offset, dtype=np.int32)
        off_r = -off_f
        dp = torch.empty((32, n, 3), dtype=torch.float32, device=self.device)
        ncell = torch.empty((32, n), dtype=torch.int32, device=self.device) if want_ncell else None
        h = self.handles
        self.pivot_bytes = 0

        def attempts(variant, slot, a, b):
            ocw = VEC_OCW[slot]
            idx = variant * 8 + slot * 2        # dp[cnt*2] / dp[cnt*8+cntc*2+8], MIMC_main.c:267-349
            self.ctx.match_async(a, b, off_f, slot, +1, ocw, False, dp[idx], None, ncell[idx] if want_ncell else None)
            self.ctx.match_async(b, a, off_r, slot, -1, ocw, True, dp[idx + 1], None, ncell[idx + 1] if want_ncell else None)

        for slot, ocw in enumerate(VEC_OCW):     # raw pair; pivots generated just in time
            off, piv = lib.get_uv_pivot(x, dt, self.params.mpp, ocw, self.H, self.W, self.params.AW_SF, self.params.AW_CRE)
            self.ctx.set_pivots(slot, off, piv)
            self.pivot_bytes += off.nbytes + piv.nbytes
            attempts(0, slot, h["i0"], h["i1"])
        # main allocates i0c/i1c once (zeros under the zero-initialised-allocation semantics)
        self.ctx.image_fill_zero(h["i0c"])
        self.ctx.image_fill_zero(h["i1c"])
        for k in range(3):
            self.ctx.conv2(h["i0"], FILTERS[k], h["i0c"])
            self.ctx.conv2(h["i1"], FILTERS[k], h["i1c"])
            for slot in range(4):
                attempts(k + 1, slot, h["i0c"], h["i1c"])
        return dp, ncell

    def multimatch(self, offset, want_ncell=False):
        """All 32 attempts (MIMC_main.c:261-350) -> dp (32, n, 3) on the device."""
        dp = torch.empty((32, self.n, 3), dtype=torch.float32, device=self.device)
        ncell = torch.empty((32, self.n), dtype=torch.int32, device=self.device) if want_ncell else None
        h = self.handles
        self.ctx.multimatch_async(h["i0"], h["i1"], h["i0c"], h["i1c"], offset, self.params, dp, ncell)
        return dp, ncell

    def postprocess(self, dp):
        planes = torch.empty((5, self.params.dimy, self.params.dimx), dtype=torch.float32, device=self.device)
        stats = self.ctx.postprocess(dp, self.xyuvav, self.params, planes)
        return planes, stats

    def postprocess_from_dump(self, directory, xyuvav, dimx, dimy, dt, finalize=True):
        """Postprocess-only re-run from the dp_NN.gma dump of an earlier run -- the flow of the reference's
        MIMC_main_test_postprocessing.c:262-300.  Returns (planes (5, dimy, dimx) host, stats, bias)."""
        from . import gma
        dp_host, _ = gma.load_dp(directory)
        x = np.ascontiguousarray(xyuvav, dtype=np.float64)
        if dp_host.shape[1] != x.shape[0]:
            raise ValueError(f"dump holds {dp_host.shape[1]} nodes, xyuvav {x.shape[0]}")
        self.params = lib.params_for(x, dimx, dimy, dt)
        self.xyuvav = x
        self.n = x.shape[0]
        dp = torch.from_numpy(dp_host).to(self.device)
        planes, stats = self.postprocess(dp)
        bias = self.ctx.finalize(planes, self.params) if finalize else (0.0, 0.0)
        self.ctx.sync()
        return planes.cpu().numpy(), stats, bias

    def postprocess_band(self, dp, xyuvav_global, params_global, own_row0, own_rows, transport):
        """This rank's band of the postprocess; collective over all ranks of ``transport``
        (halo exchange + counter all-reduce per sweep, see bands.py).  Returns (planes (5, own_rows,
        dimx) on the device, stats, BandComm)."""
        from . import bands
        geo = bands.BandGeometry(params_global.dimx, params_global.dimy, own_row0, own_rows, lib.band_halo(params_global))
        comm = bands.BandComm(transport, geo, self.device, stream=self.stream)
        planes = torch.empty((5, own_rows, params_global.dimx), dtype=torch.float32, device=self.device)
        stats = self.ctx.postprocess_band(dp, xyuvav_global, params_global, own_row0, own_rows, comm, planes)
        return planes, stats, comm

    def run(self, i0, i1, xyuvav, dimx, dimy, dt, offset, finalize=True):
        """End to end from host buffers to the five host planes (+ CP sub-pixel bias)."""
        self.set_images(i0, i1)
        dp, _ = self.match_all(xyuvav, dimx, dimy, dt, offset)
        planes, stats = self.postprocess(dp)
        bias = (0.0, 0.0)
        if finalize:
            bias = self.ctx.finalize(planes, self.params)
        self.ctx.sync()
        return planes.cpu().numpy(), stats, bias
