"""ctypes front-end of the parity oracle -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py``
(``cpu_baseline`` and ``--impl reference``) may import this package.  The shipped
CUDA path (``mimc3_b200``) never does.

Two libraries:

* ``Oracle``    -> ``oracle/_build/libmimc3oracle.so``: the restated C oracle
  (``mimc3_oracle.c``), buildable anywhere with gcc.
* ``Reference`` -> ``oracle/_ref/libmimc3ref.so``: the UNMODIFIED reference
  ``MIMC_module.c``/``GMA.c`` behind ``ref_harness.c``; built in the container that has
  ``/root/reference`` and shipped to the GPU box as a prebuilt file.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_build", "libmimc3oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libmimc3ref.so")
REF_CLI = os.path.join(_HERE, "_ref", "MIMC3_ref")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

# defaults hard-coded in the reference driver, MIMC_main.c:134-168
AW_CRE = 10.0
AW_SF = 1.8
VEC_OCW = (7, 15, 30, 40)
# the three filter kernels, MIMC_main.c:175-196
KERNELS = (
    np.array([[-1, 0, 1]], dtype=np.float32),
    np.array([[-1], [0], [1]], dtype=np.float32),
    np.array([[-0.125, -0.125, -0.125], [-0.125, 1.0, -0.125], [-0.125, -0.125, -0.125]], dtype=np.float32),
)


def build(force: bool = False) -> None:
    """Compile the oracle (and, where /root/reference exists, oracle/_ref)."""
    args = ["make", "-C", _HERE, "all"] + (["-B"] if force else [])
    subprocess.run(args, check=True, stdout=subprocess.DEVNULL)


class CpParams(C.Structure):
    """orc_cp_params: the control-point constants of MIMC_main.c:134-168."""
    _fields_ = [("vec_ocw", C.c_int32 * 4), ("AW_CRE", C.c_float), ("num_cp_max", C.c_int32), ("num_cp_min", C.c_int32),
                ("ratio_cp", C.c_float), ("thres_spd_cp", C.c_float)]


class PostParams(C.Structure):
    _fields_ = [("dt", C.c_float), ("mpp", C.c_float), ("meter_per_spacing", C.c_float),
                ("radius_neighbor_dpf1", C.c_float), ("radius_neighbor_ps", C.c_float),
                ("dimx", C.c_int32), ("dimy", C.c_int32), ("num_dp", C.c_int32)]


def post_params(xyuvav: np.ndarray, dimx: int, dimy: int, dt: float, num_dp: int = 32) -> PostParams:
    """The values main derives from xyuvav rows 0-1 (MIMC_main.c:221-223) + defaults :160-162."""
    mpp = np.float32((xyuvav[1, 0] - xyuvav[0, 0]) / (xyuvav[1, 2] - xyuvav[0, 2]))
    mps = np.float32(xyuvav[1, 0] - xyuvav[0, 0])
    return PostParams(dt=np.float32(dt), mpp=mpp, meter_per_spacing=mps, radius_neighbor_dpf1=float(1000 // 300),
                      radius_neighbor_ps=5.0, dimx=dimx, dimy=dimy, num_dp=num_dp)


def _as(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Oracle:
    """The restated C oracle (mimc3_oracle.h)."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.orc_get_uv_pivot.restype = C.c_int64
        L.orc_get_uv_pivot.argtypes = [_f64p, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                                       C.c_int32, C.c_int32, _i32p, _i32p, C.c_int64]
        L.orc_match.restype = None
        L.orc_match.argtypes = [_f32p, _f32p, C.c_int32, C.c_int32, _f64p, C.c_int32, _i32p, _i32p, _i32p,
                                C.c_int32, C.c_int32, _f32p, _i32p, _i32p]
        L.orc_model_match.restype = None
        L.orc_model_match.argtypes = [_f32p, _f32p, C.c_int32, C.c_int32, _f64p, C.c_int32, _i32p, _i32p, _i32p,
                                      C.c_int32, C.c_int32, C.c_int32, _f32p, _i32p, _i32p, _i32p]
        L.orc_find_ncc_peak.restype = None
        L.orc_find_ncc_peak.argtypes = [_f32p, C.c_int32, _f32p, C.c_int32, C.c_int32, _i32p, C.c_int32,
                                        _f32p, _i32p, _i32p]
        L.orc_conv2.restype = None
        L.orc_conv2.argtypes = [_f32p, C.c_int32, C.c_int32, _f32p, C.c_int32, C.c_int32, _f32p]
        L.orc_cluster.restype = None
        L.orc_cluster.argtypes = [_f32p, C.c_int32, C.c_int32, _f32p, _i32p]
        L.orc_ruv_neighbor.restype = C.c_int32
        L.orc_ruv_neighbor.argtypes = [_f64p, C.POINTER(PostParams), C.c_float, _i32p, C.c_int32]
        L.orc_dpf0.restype = None
        L.orc_dpf0.argtypes = [_f32p, _i32p, C.POINTER(PostParams), C.c_float, _i32p]
        L.orc_dpf1.restype = C.c_int32
        L.orc_dpf1.argtypes = [_i32p, _f32p, _f32p, _i32p, C.c_int32, _f32p, _i32p, _f64p, C.POINTER(PostParams)]
        L.orc_pseudosmooth.restype = C.c_int32
        L.orc_pseudosmooth.argtypes = L.orc_dpf1.argtypes
        L.orc_postprocess.restype = None
        L.orc_postprocess.argtypes = [_f32p, _f64p, C.POINTER(PostParams), _f32p]
        L.orc_finalize.restype = None
        L.orc_finalize.argtypes = [_f32p, C.POINTER(PostParams), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_num_threads.restype = C.c_int
        L.orc_get_offset_image.restype = C.c_int
        L.orc_get_offset_image.argtypes = [_f32p, _f32p, C.c_int32, C.c_int32, _f64p, C.c_int32, C.POINTER(CpParams), C.c_uint,
                                           _i32p, _u8p]

    def num_threads(self) -> int:
        return int(self.lib.orc_num_threads())

    def get_uv_pivot(self, xyuvav, dt, mpp, ocw, H, W, aw_sf=AW_SF, aw_cre=AW_CRE):
        x = _as(xyuvav, np.float64)
        n = x.shape[0]
        off = np.zeros(n + 1, np.int32)
        cap = 16 * n + 1024
        while True:
            piv = np.zeros((cap, 2), np.int32)
            tot = self.lib.orc_get_uv_pivot(x, n, dt, mpp, aw_sf, aw_cre, ocw, H, W, off, piv, cap)
            if tot >= 0:
                return off, piv[:tot].copy()
            cap = -tot

    def match(self, i0, i1, xyuvav, offset, csr_off, piv, sign, ocw):
        i0 = _as(i0, np.float32); i1 = _as(i1, np.float32); x = _as(xyuvav, np.float64)
        H, W = i0.shape
        n = x.shape[0]
        out = np.zeros((n, 3), np.float32); peak = np.zeros((n, 2), np.int32); ncell = np.zeros(n, np.int32)
        piv = _as(piv, np.int32).reshape(-1, 2)
        if piv.shape[0] == 0:
            piv = np.zeros((1, 2), np.int32)
        self.lib.orc_match(i0, i1, H, W, x, n, _as(offset, np.int32), _as(csr_off, np.int32), piv,
                           sign, ocw, out, peak, ncell)
        return out, peak, ncell

    def model_match(self, i0, i1, xyuvav, offset, csr_off, piv, sign, ocw, maxj=64):
        """orc_match through the CUDA kernel's explore/replay schedule (leader_model.c) -> (dp, peak, ncell,
        stats (n,4): rounds, cells computed, explore steps, replay steps)."""
        i0 = _as(i0, np.float32); i1 = _as(i1, np.float32); x = _as(xyuvav, np.float64)
        H, W = i0.shape
        n = x.shape[0]
        out = np.zeros((n, 3), np.float32); peak = np.zeros((n, 2), np.int32); ncell = np.zeros(n, np.int32)
        stats = np.zeros((n, 4), np.int32)
        piv = _as(piv, np.int32).reshape(-1, 2)
        if piv.shape[0] == 0:
            piv = np.zeros((1, 2), np.int32)
        self.lib.orc_model_match(i0, i1, H, W, x, n, _as(offset, np.int32), _as(csr_off, np.int32), piv,
                                 sign, ocw, maxj, out, peak, ncell, stats)
        return out, peak, ncell, stats

    def find_ncc_peak(self, refchip, sarea, piv):
        r = _as(refchip, np.float32); s = _as(sarea, np.float32); p = _as(piv, np.int32).reshape(-1, 2)
        uv = np.zeros(3, np.float32); pk = np.zeros(2, np.int32); nc = np.zeros(1, np.int32)
        self.lib.orc_find_ncc_peak(r, r.shape[0], s, s.shape[0], s.shape[1], p, p.shape[0], uv, pk, nc)
        return uv, pk, int(nc[0])

    def get_offset_image(self, i0, i1, xyuvav, seed):
        """get_offset_image (MIMC_module.c:33-492) with the reference's constants -> (rc, offset, flag_cp)."""
        i0 = _as(i0, np.float32); i1 = _as(i1, np.float32); x = _as(xyuvav, np.float64)
        p = CpParams((C.c_int32 * 4)(*VEC_OCW), AW_CRE, 500, 50, 0.03, 10.0)
        off = np.zeros(2, np.int32); flag = np.zeros(x.shape[0], np.uint8)
        rc = self.lib.orc_get_offset_image(i0, i1, i0.shape[0], i0.shape[1], x, x.shape[0], C.byref(p), int(seed) & 0xffffffff, off, flag)
        return rc, off, flag

    def conv2(self, img, kernel_id, out):
        """In place on `out` (float32 (H,W)), like main's reused i0c/i1c buffers."""
        img = _as(img, np.float32)
        k = KERNELS[kernel_id]
        assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == img.shape
        self.lib.orc_conv2(img, img.shape[0], img.shape[1], _as(k, np.float32), k.shape[0], k.shape[1], out)
        return out

    def cluster(self, dp):
        """dp: (num_dpoi, n, 3) -> mvn (n, num_dpoi, 5), ncl (n)."""
        dp = _as(dp, np.float32)
        K, n, _ = dp.shape
        mvn = np.zeros((n, K, 5), np.float32); ncl = np.zeros(n, np.int32)
        self.lib.orc_cluster(dp, n, K, mvn, ncl)
        return mvn, ncl

    def ruv_neighbor(self, xyuvav, pp, radius):
        x = _as(xyuvav, np.float64)
        cap = 4096
        ruv = np.zeros((cap, 2), np.int32)
        k = self.lib.orc_ruv_neighbor(x, C.byref(pp), radius, ruv, cap)
        assert k <= cap
        return ruv[:k].copy()

    def postprocess_stages(self, dp, xyuvav, pp):
        """Returns dict of every intermediate field of mimc2_postprocess."""
        x = _as(xyuvav, np.float64)
        n = pp.dimx * pp.dimy
        mvn, ncl = self.cluster(dp)
        dpf0 = np.zeros(n, np.int32)
        self.lib.orc_dpf0(mvn, ncl, C.byref(pp), 0.6, dpf0)
        res = {"mvn": mvn, "ncl": ncl, "dpf0": dpf0.copy()}
        ruv = self.ruv_neighbor(x, pp, pp.radius_neighbor_dpf1)
        dx = np.zeros(n, np.float32); dy = np.zeros(n, np.float32)
        ids = dpf0.copy()
        res["dpf1_sweeps"] = self.lib.orc_dpf1(ids, dx, dy, ruv, ruv.shape[0], mvn, ncl, x, C.byref(pp))
        res.update(dpf1_id=ids.copy(), dpf1_dx=dx.copy(), dpf1_dy=dy.copy(), ruv_dpf1=ruv)
        ruv = self.ruv_neighbor(x, pp, pp.radius_neighbor_ps)
        res["ps_sweeps"] = self.lib.orc_pseudosmooth(ids, dx, dy, ruv, ruv.shape[0], mvn, ncl, x, C.byref(pp))
        res.update(ps_id=ids, ps_dx=dx, ps_dy=dy, ruv_ps=ruv)
        return res

    def postprocess(self, dp, xyuvav, pp):
        dp = _as(dp, np.float32)
        n = pp.dimx * pp.dimy
        planes = np.zeros((5, pp.dimy, pp.dimx), np.float32)
        self.lib.orc_postprocess(dp, _as(xyuvav, np.float64), C.byref(pp), planes)
        return planes

    def finalize(self, planes, pp):
        planes = np.ascontiguousarray(planes, dtype=np.float32).copy()
        a = C.c_float(); b = C.c_float()
        self.lib.orc_finalize(planes, C.byref(pp), C.byref(a), C.byref(b))
        return planes, a.value, b.value


class Reference:
    """The unmodified reference behind ref_harness.c (prebuilt oracle/_ref/libmimc3ref.so)."""

    def __init__(self, path: str = REF_SO, quiet: bool = True):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build it with `make -C oracle ref` where /root/reference exists)")
        L = self.lib = C.CDLL(path)
        L.ref_set_globals.restype = None
        L.ref_set_globals.argtypes = [C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float]
        L.ref_get_uv_pivot.restype = C.c_int64
        L.ref_get_uv_pivot.argtypes = [_f64p, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p, C.c_int64]
        L.ref_match.restype = C.c_double
        L.ref_match.argtypes = [_f32p, _f32p, C.c_int32, C.c_int32, _f64p, C.c_int32, _i32p, _i32p, _i32p,
                                C.c_int32, C.c_int32, _f32p]
        L.ref_conv2.restype = None
        L.ref_conv2.argtypes = [_f32p, C.c_int32, C.c_int32, C.c_int32, _f32p]
        L.ref_cluster.restype = None
        L.ref_cluster.argtypes = [_f32p, C.c_int32, C.c_int32, _f32p, _i32p]
        L.ref_postprocess.restype = None
        L.ref_postprocess.argtypes = [_f32p, _f64p, C.c_int32, _f32p]
        L.ref_postprocess_stages.restype = None
        L.ref_postprocess_stages.argtypes = [_f32p, _f64p, C.c_int32, _i32p, _i32p, _f32p, _f32p, _i32p, _f32p, _f32p]
        L.ref_get_offset_image.restype = C.c_int
        L.ref_get_offset_image.argtypes = [_f32p, _f32p, C.c_int32, C.c_int32, _f64p, C.c_int32, _i32p, _u8p]
        L.ref_set_fake_time.restype = None
        L.ref_set_fake_time.argtypes = [C.c_int64, C.c_int]
        L.ref_quiet.restype = None
        L.ref_quiet.argtypes = [C.c_int]
        L.ref_num_threads.restype = C.c_int
        if hasattr(L, "ref_set_num_threads"):
            L.ref_set_num_threads.restype = None
            L.ref_set_num_threads.argtypes = [C.c_int]
        self.quiet = quiet

    def _q(self, on):
        if self.quiet:
            self.lib.ref_quiet(1 if on else 0)

    def num_threads(self) -> int:
        return int(self.lib.ref_num_threads())

    def set_num_threads(self, n: int) -> int:
        """OpenMP team size of the reference's parallel regions (returns what is in effect)."""
        if hasattr(self.lib, "ref_set_num_threads"):
            self.lib.ref_set_num_threads(int(n))
        return self.num_threads()

    def set_globals(self, xyuvav, dimx, dimy, dt, num_dp=32):
        x = np.asarray(xyuvav)
        mpp = np.float32((x[1, 0] - x[0, 0]) / (x[1, 2] - x[0, 2]))
        spacing = np.float32(x[1, 2] - x[0, 2])
        mps = np.float32(x[1, 0] - x[0, 0])
        self.lib.ref_set_globals(dt, num_dp, dimx * dimy, dimx, dimy, mpp, spacing, mps)
        return float(mpp)

    def get_uv_pivot(self, xyuvav, dt, ocw, H, W):
        x = _as(xyuvav, np.float64)
        n = x.shape[0]
        off = np.zeros(n + 1, np.int32)
        cap = 16 * n + 1024
        while True:
            piv = np.zeros((cap, 2), np.int32)
            self._q(True)
            tot = self.lib.ref_get_uv_pivot(x, n, dt, ocw, H, W, off, piv, cap)
            self._q(False)
            if tot >= 0:
                return off, piv[:tot].copy()
            cap = -tot

    def match(self, i0, i1, xyuvav, offset, csr_off, piv, sign, ocw):
        i0 = _as(i0, np.float32); i1 = _as(i1, np.float32); x = _as(xyuvav, np.float64)
        H, W = i0.shape
        out = np.zeros((x.shape[0], 3), np.float32)
        secs = self.lib.ref_match(i0, i1, H, W, x, x.shape[0], _as(offset, np.int32), _as(csr_off, np.int32),
                                  _as(piv, np.int32).reshape(-1, 2), sign, ocw, out)
        return out, secs

    def conv2(self, img, kernel_id, out):
        img = _as(img, np.float32)
        assert out.dtype == np.float32 and out.flags.c_contiguous
        self.lib.ref_conv2(img, img.shape[0], img.shape[1], kernel_id, out)
        return out

    def cluster(self, dp):
        dp = _as(dp, np.float32)
        K, n, _ = dp.shape
        mvn = np.zeros((n, K, 5), np.float32); ncl = np.zeros(n, np.int32)
        self.lib.ref_cluster(dp, n, K, mvn, ncl)
        return mvn, ncl

    def postprocess(self, dp, xyuvav, dimx, dimy):
        dp = _as(dp, np.float32)
        planes = np.zeros((5, dimy, dimx), np.float32)
        self._q(True)
        self.lib.ref_postprocess(dp, _as(xyuvav, np.float64), dimx * dimy, planes)
        self._q(False)
        return planes

    def postprocess_stages(self, dp, xyuvav, dimx, dimy):
        dp = _as(dp, np.float32)
        n = dimx * dimy
        r = {k: np.zeros(n, np.int32) for k in ("dpf0", "dpf1_id", "ps_id")}
        r.update({k: np.zeros(n, np.float32) for k in ("dpf1_dx", "dpf1_dy", "ps_dx", "ps_dy")})
        self._q(True)
        self.lib.ref_postprocess_stages(dp, _as(xyuvav, np.float64), n, r["dpf0"], r["dpf1_id"], r["dpf1_dx"],
                                        r["dpf1_dy"], r["ps_id"], r["ps_dx"], r["ps_dy"])
        self._q(False)
        return r

    def get_offset_image(self, i0, i1, xyuvav, fake_time=None):
        i0 = _as(i0, np.float32); i1 = _as(i1, np.float32); x = _as(xyuvav, np.float64)
        if fake_time is not None:
            self.lib.ref_set_fake_time(int(fake_time), 1)
        off = np.zeros(2, np.int32); flag = np.zeros(x.shape[0], np.uint8)
        self._q(True)
        rc = self.lib.ref_get_offset_image(i0, i1, i0.shape[0], i0.shape[1], x, x.shape[0], off, flag)
        self._q(False)
        return rc, off, flag
