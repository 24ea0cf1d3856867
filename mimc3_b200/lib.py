"""ctypes binding of libmimc3cu.so (include/mimc3cu.h) -- the product's Python host layer.

The library is the product; this module only marshals numpy / torch buffers into the
C ABI.  There is no CPU fallback: `Context()` raises if the shared library is missing or
no sm_100 device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MIMC3CU_LIB selects a development variant of the library (mimc3_b200/build.py --variant); there is still no fallback
LIB_PATH = os.environ.get("MIMC3CU_LIB") or os.path.join(HERE, "libmimc3cu.so")


class Mimc3CuError(RuntimeError):
    pass


class Params(C.Structure):
    """mimc3cu_params (include/mimc3cu.h) == the reference's hard-coded MIMC_main.c:134-168."""
    _fields_ = [("vec_ocw", C.c_int32 * 4), ("AW_CRE", C.c_float), ("AW_SF", C.c_float), ("mpp", C.c_float),
                ("meter_per_spacing", C.c_float), ("radius_neighbor_dpf1", C.c_float),
                ("radius_neighbor_ps", C.c_float), ("dt", C.c_float), ("dimx", C.c_int32), ("dimy", C.c_int32),
                ("num_dp", C.c_int32), ("num_cp_max", C.c_int32), ("num_cp_min", C.c_int32),
                ("ratio_cp", C.c_float), ("thres_spd_cp", C.c_float)]


_lib = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """Load libmimc3cu.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise Mimc3CuError(f"{path} is missing: run `python -m mimc3_b200.build` (nvcc, sm_100a). "
                           "There is no CPU fallback.")
    L = C.CDLL(path)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    sigs = {
        "mimc3cu_default_params": (None, [C.POINTER(Params)]),
        "mimc3cu_version": (C.c_int, []),
        "mimc3cu_device_count": (C.c_int, []),
        "mimc3cu_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "mimc3cu_destroy": (None, [vp]),
        "mimc3cu_last_error": (C.c_char_p, [vp]),
        "mimc3cu_stream": (vp, [vp]),
        "mimc3cu_sync": (C.c_int, [vp]),
        "mimc3cu_launch_count": (i64, [vp]),
        "mimc3cu_image_create": (C.c_int, [vp, i32, i32, C.POINTER(i32)]),
        "mimc3cu_image_destroy": (C.c_int, [vp, i32]),
        "mimc3cu_image_upload": (C.c_int, [vp, i32, vp]),
        "mimc3cu_image_upload_u8": (C.c_int, [vp, i32, vp]),
        "mimc3cu_image_upload_u16": (C.c_int, [vp, i32, vp]),
        "mimc3cu_image_copy_from_device": (C.c_int, [vp, i32, vp]),
        "mimc3cu_image_download": (C.c_int, [vp, i32, vp]),
        "mimc3cu_image_fill_zero": (C.c_int, [vp, i32]),
        "mimc3cu_image_ptr": (vp, [vp, i32]),
        "mimc3cu_image_invalidate": (C.c_int, [vp, i32]),
        "mimc3cu_conv2": (C.c_int, [vp, i32, vp, i32, i32, i32]),
        "mimc3cu_get_uv_pivot": (i64, [vp, i32, f32, f32, f32, f32, i32, i32, i32, vp, vp]),
        "mimc3cu_set_nodes": (C.c_int, [vp, vp, i32]),
        "mimc3cu_set_pivots": (C.c_int, [vp, i32, vp, vp, i32]),
        "mimc3cu_match_async": (C.c_int, [vp, i32, i32, vp, i32, i32, i32, i32, vp, vp, vp]),
        "mimc3cu_match": (C.c_int, [vp, i32, i32, vp, i32, i32, i32, i32, vp, vp, vp]),
        "mimc3cu_find_ncc_peak_batch": (C.c_int, [vp, vp, i32, vp, i32, i32, vp, i32, vp, vp, vp]),
        "mimc3cu_multimatch_async": (C.c_int, [vp, i32, i32, i32, i32, vp, C.POINTER(Params), vp, vp]),
        "mimc3cu_multimatch_diag_async": (C.c_int, [vp, i32, i32, i32, i32, vp, C.POINTER(Params), vp, vp, vp]),
        "mimc3cu_cluster_async": (C.c_int, [vp, vp, i32, i32, vp, vp]),
        "mimc3cu_postprocess": (C.c_int, [vp, vp, vp, C.POINTER(Params), vp, vp]),
        "mimc3cu_postprocess_stage": (C.c_int, [vp, i32, vp]),
        "mimc3cu_finalize": (C.c_int, [vp, vp, C.POINTER(Params), C.POINTER(f32), C.POINTER(f32)]),
        "mimc3cu_fp32_peak": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "mimc3cu_timing_enable": (C.c_int, [vp, C.c_int]),
        "mimc3cu_timing_read": (C.c_int, [vp, vp, vp]),
        "mimc3cu_malloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
        "mimc3cu_free": (C.c_int, [vp, vp]),
        "mimc3cu_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "mimc3cu_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
        "mimc3cu_get_offset_image": (C.c_int, [vp, i32, i32, vp, i32, C.POINTER(Params), vp, vp, vp, C.c_uint32, vp, vp,
                                               C.POINTER(i32), C.POINTER(i32)]),
        "mimc3cu_band_halo": (C.c_int, [C.POINTER(Params)]),
        "mimc3cu_postprocess_band": (C.c_int, [vp, vp, vp, C.POINTER(Params), i32, i32, vp, vp, vp]),
        "mimc3cu_comm_unique_id": (C.c_int, [vp]),
        "mimc3cu_comm_init_rank": (C.c_int, [vp, vp, i32, i32]),
        "mimc3cu_comm_init_all": (C.c_int, [C.POINTER(vp), i32]),
        "mimc3cu_comm_destroy": (None, [vp]),
        "mimc3cu_comm_info": (C.c_int, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]),
        "mimc3cu_comm_gather": (C.c_int, [vp, vp, vp, vp, i32]),
        "mimc3cu_comm_timing": (C.c_int, [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "mimc3cu_dp_negate_uv_async": (C.c_int, [vp, vp, i32]),
        "mimc3cu_set_matcher": (C.c_int, [vp, i32]),
        "mimc3cu_last_matcher": (C.c_int, [vp]),
        "mimc3cu_image_class": (C.c_int, [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(f32)]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._mimc3cu_symbols = tuple(sigs)
    _lib = L
    return L


EXPORTED_SYMBOLS = (
    "mimc3cu_default_params", "mimc3cu_version", "mimc3cu_device_count", "mimc3cu_create", "mimc3cu_destroy",
    "mimc3cu_last_error", "mimc3cu_stream", "mimc3cu_sync", "mimc3cu_launch_count", "mimc3cu_image_create",
    "mimc3cu_image_destroy", "mimc3cu_image_upload", "mimc3cu_image_upload_u8", "mimc3cu_image_upload_u16",
    "mimc3cu_image_copy_from_device", "mimc3cu_image_download", "mimc3cu_image_ptr", "mimc3cu_conv2",
    "mimc3cu_get_uv_pivot", "mimc3cu_set_nodes", "mimc3cu_set_pivots", "mimc3cu_match_async", "mimc3cu_match",
    "mimc3cu_find_ncc_peak_batch", "mimc3cu_multimatch_async", "mimc3cu_cluster_async", "mimc3cu_postprocess",
    "mimc3cu_postprocess_stage", "mimc3cu_finalize", "mimc3cu_fp32_peak", "mimc3cu_timing_enable",
    "mimc3cu_timing_read", "mimc3cu_malloc", "mimc3cu_free", "mimc3cu_memcpy_d2h",
    "mimc3cu_memcpy_h2d", "mimc3cu_set_matcher", "mimc3cu_last_matcher", "mimc3cu_image_class",
    "mimc3cu_get_offset_image", "mimc3cu_band_halo", "mimc3cu_postprocess_band", "mimc3cu_image_fill_zero",
    "mimc3cu_multimatch_diag_async", "mimc3cu_comm_unique_id", "mimc3cu_comm_init_rank", "mimc3cu_comm_init_all",
    "mimc3cu_comm_destroy", "mimc3cu_comm_info", "mimc3cu_comm_gather", "mimc3cu_dp_negate_uv_async",
    "mimc3cu_image_invalidate", "mimc3cu_comm_timing",
)


def _ptr(a):
    """Address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags.c_contiguous
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr()
    raise TypeError(type(a))


def default_params() -> Params:
    p = Params()
    load_library().mimc3cu_default_params(C.byref(p))
    return p


def params_for(xyuvav: np.ndarray, dimx: int, dimy: int, dt: float) -> Params:
    """What main derives from the xyuvav matrix, MIMC_main.c:211-223."""
    p = default_params()
    p.mpp = np.float32((xyuvav[1, 0] - xyuvav[0, 0]) / (xyuvav[1, 2] - xyuvav[0, 2]))
    p.meter_per_spacing = np.float32(xyuvav[1, 0] - xyuvav[0, 0])
    p.dt = np.float32(dt)
    p.dimx, p.dimy = dimx, dimy
    return p


def band_halo(params) -> int:
    """Node rows of halo each band keeps of its neighbours (mimc3cu_band_halo)."""
    return int(load_library().mimc3cu_band_halo(C.byref(params)))


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id (call on one rank, distribute to the others, then Context.comm_init_rank)."""
    L = load_library()
    buf = C.create_string_buffer(128)
    if L.mimc3cu_comm_unique_id(C.cast(buf, C.c_void_p)):
        raise Mimc3CuError(L.mimc3cu_last_error(None).decode())
    return buf.raw


def comm_init_all(contexts) -> None:
    """One process driving several GPUs: attach an NCCL communicator to every context (rank = list position).
    The bands' collective calls must then be issued from one host thread per context."""
    L = load_library()
    arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    if L.mimc3cu_comm_init_all(arr, len(contexts)):
        raise Mimc3CuError(L.mimc3cu_last_error(contexts[0].h).decode())


def device_count() -> int:
    return int(load_library().mimc3cu_device_count())


def get_uv_pivot(xyuvav, dt, mpp, ocw, H, W, aw_sf=1.8, aw_cre=10.0):
    """get_uv_pivot (MIMC_module.c:543-602) -> CSR (off[n+1], piv[total,2])."""
    L = load_library()
    x = np.ascontiguousarray(xyuvav, dtype=np.float64)
    n = x.shape[0]
    off = np.zeros(n + 1, np.int32)
    tot = L.mimc3cu_get_uv_pivot(_ptr(x), n, dt, mpp, aw_sf, aw_cre, ocw, H, W, _ptr(off), None)
    if tot < 0:
        raise Mimc3CuError(L.mimc3cu_last_error(None).decode())
    piv = np.zeros((max(tot, 1), 2), np.int32)
    L.mimc3cu_get_uv_pivot(_ptr(x), n, dt, mpp, aw_sf, aw_cre, ocw, H, W, _ptr(off), _ptr(piv))
    return off, piv[:tot]


class Context:
    """One mimc3cu context (one GPU, one stream)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        if self.L.mimc3cu_create(device, C.byref(h)):
            raise Mimc3CuError(self.L.mimc3cu_last_error(None).decode())
        self.h = h
        self.device = device
        self.n = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.mimc3cu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise Mimc3CuError(self.L.mimc3cu_last_error(self.h).decode())

    @property
    def stream(self) -> int:
        return self.L.mimc3cu_stream(self.h) or 0

    def sync(self):
        self._ck(self.L.mimc3cu_sync(self.h))

    def launch_count(self) -> int:
        return int(self.L.mimc3cu_launch_count(self.h))

    # images ------------------------------------------------------------------------------
    def image_create(self, H, W) -> int:
        h = C.c_int32()
        self._ck(self.L.mimc3cu_image_create(self.h, H, W, C.byref(h)))
        return h.value

    def image_destroy(self, handle):
        self._ck(self.L.mimc3cu_image_destroy(self.h, handle))

    def image_upload(self, handle, host):
        """host: numpy float32 / uint8 / uint16 (H, W)."""
        host = np.ascontiguousarray(host)
        fn = {np.dtype(np.float32): self.L.mimc3cu_image_upload, np.dtype(np.uint8): self.L.mimc3cu_image_upload_u8,
              np.dtype(np.uint16): self.L.mimc3cu_image_upload_u16}[host.dtype]
        self._ck(fn(self.h, handle, _ptr(host)))

    def image_from(self, arr) -> int:
        """Create + fill an image from a numpy array (host) or a CUDA torch tensor (device copy)."""
        H, W = arr.shape
        h = self.image_create(H, W)
        if isinstance(arr, np.ndarray):
            self.image_upload(h, arr)
        elif arr.is_cuda:
            import torch
            assert arr.dtype == torch.float32
            torch.cuda.current_stream(arr.device).synchronize()
            self._ck(self.L.mimc3cu_image_copy_from_device(self.h, h, _ptr(arr.contiguous())))
        else:
            self.image_upload(h, arr.numpy())
        return h

    def image_fill_zero(self, handle):
        self._ck(self.L.mimc3cu_image_fill_zero(self.h, handle))

    def image_download(self, handle, H, W) -> np.ndarray:
        out = np.empty((H, W), np.float32)
        self._ck(self.L.mimc3cu_image_download(self.h, handle, _ptr(out)))
        return out

    def conv2(self, src, kernel, dst):
        k = np.ascontiguousarray(kernel, dtype=np.float32)
        self._ck(self.L.mimc3cu_conv2(self.h, src, _ptr(k), k.shape[0], k.shape[1], dst))

    # nodes / pivots ---------------------------------------------------------------------------
    def set_nodes(self, xyuvav):
        x = np.ascontiguousarray(xyuvav, dtype=np.float64)
        self._ck(self.L.mimc3cu_set_nodes(self.h, _ptr(x), x.shape[0]))
        self.n = x.shape[0]

    def set_pivots(self, slot, off, piv):
        off = np.ascontiguousarray(off, dtype=np.int32)
        piv = np.ascontiguousarray(piv, dtype=np.int32).reshape(-1, 2)
        if piv.shape[0] == 0:
            piv = np.zeros((1, 2), np.int32)
        self._ck(self.L.mimc3cu_set_pivots(self.h, slot, _ptr(off), _ptr(piv), off.shape[0] - 1))

    # matcher -------------------------------------------------------------------------------------
    def set_matcher(self, mode):
        """0 auto, 1 general FP64 kernel, 2 require the exact-FP32 kernel (also 'auto'/'v1'/'v2')."""
        mode = {"auto": 0, "v1": 1, "general": 1, "v2": 2}.get(mode, mode)
        self._ck(self.L.mimc3cu_set_matcher(self.h, int(mode)))

    def last_matcher(self) -> int:
        return int(self.L.mimc3cu_last_matcher(self.h))

    def image_class(self, handle):
        """(exact_class, frac_bits, max_value) of an image as the matcher sees it."""
        a = C.c_int32(); b = C.c_int32(); m = C.c_float()
        self._ck(self.L.mimc3cu_image_class(self.h, handle, C.byref(a), C.byref(b), C.byref(m)))
        return bool(a.value), b.value, m.value

    def match(self, ref_img, search_img, offset, slot, sign, ocw, negate=False):
        """Synchronous; returns host arrays dp (n,3), peak (n,2), ncell (n)."""
        n = self.n
        dp = np.empty((n, 3), np.float32); peak = np.empty((n, 2), np.int32); ncell = np.empty(n, np.int32)
        off = np.ascontiguousarray(offset, dtype=np.int32)
        self._ck(self.L.mimc3cu_match(self.h, ref_img, search_img, _ptr(off), slot, sign, ocw, int(negate),
                                      _ptr(dp), _ptr(peak), _ptr(ncell)))
        return dp, peak, ncell

    def match_async(self, ref_img, search_img, offset, slot, sign, ocw, negate, dp_dev, peak_dev=None, ncell_dev=None):
        off = np.ascontiguousarray(offset, dtype=np.int32)
        self._ck(self.L.mimc3cu_match_async(self.h, ref_img, search_img, _ptr(off), slot, sign, ocw, int(negate),
                                            _ptr(dp_dev), _ptr(peak_dev), _ptr(ncell_dev)))

    def find_ncc_peak_batch(self, refchips, sareas, piv):
        r = np.ascontiguousarray(refchips, dtype=np.float32); s = np.ascontiguousarray(sareas, dtype=np.float32)
        p = np.ascontiguousarray(piv, dtype=np.int32).reshape(-1, 2)
        cnt = r.shape[0]
        uv = np.empty((cnt, 3), np.float32); pk = np.empty((cnt, 2), np.int32); nc = np.empty(cnt, np.int32)
        self._ck(self.L.mimc3cu_find_ncc_peak_batch(self.h, _ptr(r), r.shape[1], _ptr(s), s.shape[1], cnt, _ptr(p),
                                                    p.shape[0], _ptr(uv), _ptr(pk), _ptr(nc)))
        return uv, pk, nc

    def multimatch_async(self, i0, i1, i0c, i1c, offset, params, dp_dev, ncell_dev=None, peak_dev=None):
        off = np.ascontiguousarray(offset, dtype=np.int32)
        self._ck(self.L.mimc3cu_multimatch_diag_async(self.h, i0, i1, i0c, i1c, _ptr(off), C.byref(params), _ptr(dp_dev),
                                                      _ptr(ncell_dev), _ptr(peak_dev)))

    # control points ---------------------------------------------------------------------------------
    def get_offset_image(self, i0, i1, xyuvav, params, seed):
        """get_offset_image (MIMC_module.c:33-492) -> (result 1/-1, offset int32[2], flag_cp uint8[n], #CP)."""
        x = np.ascontiguousarray(xyuvav, dtype=np.float64)
        off = np.zeros(2, np.int32); flag = np.zeros(x.shape[0], np.uint8)
        res = C.c_int32(); ncp = C.c_int32()
        self._ck(self.L.mimc3cu_get_offset_image(self.h, i0, i1, _ptr(x), x.shape[0], C.byref(params), None, None, None,
                                                 int(seed) & 0xffffffff, _ptr(off), _ptr(flag), C.byref(res), C.byref(ncp)))
        return res.value, off, flag, ncp.value

    # postprocess ------------------------------------------------------------------------------------
    def cluster_async(self, dp_dev, n, num_dp, mvn_dev, ncl_dev):
        self._ck(self.L.mimc3cu_cluster_async(self.h, _ptr(dp_dev), n, num_dp, _ptr(mvn_dev), _ptr(ncl_dev)))

    def postprocess(self, dp_dev, xyuvav, params, planes_dev):
        x = np.ascontiguousarray(xyuvav, dtype=np.float64)
        stats = np.zeros(4, np.int32)
        self._ck(self.L.mimc3cu_postprocess(self.h, _ptr(dp_dev), _ptr(x), C.byref(params), _ptr(planes_dev), _ptr(stats)))
        return stats

    def postprocess_band(self, dp_dev, xyuvav_global, params_global, own_row0, own_rows, comm, planes_dev):
        """One band of node rows of mimc2_postprocess (collective over the bands).  ``comm`` is a
        bands.BandComm, or None: the single band [0, dimy), or -- with comm_init_rank done -- the library's
        own NCCL communicator."""
        x = np.ascontiguousarray(xyuvav_global, dtype=np.float64)
        stats = np.zeros(4, np.int32)
        cptr = C.cast(C.pointer(comm.struct), C.c_void_p) if comm is not None else None
        rc = self.L.mimc3cu_postprocess_band(self.h, _ptr(dp_dev), _ptr(x), C.byref(params_global), own_row0, own_rows, cptr,
                                             _ptr(planes_dev), _ptr(stats))
        if rc and comm is not None and comm.error is not None:
            raise comm.error
        self._ck(rc)
        return stats

    # the library's own NCCL band communicator ------------------------------------------------------------------
    def comm_init_rank(self, unique_id: bytes, rank: int, world: int):
        """Attach an NCCL communicator (one rank per node-row band); ``unique_id`` = comm_unique_id() of rank 0."""
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._ck(self.L.mimc3cu_comm_init_rank(self.h, C.cast(buf, C.c_void_p), rank, world))

    def comm_info(self):
        r = C.c_int32(); w = C.c_int32(); e = C.c_int64(); a = C.c_int64()
        if self.L.mimc3cu_comm_info(self.h, C.byref(r), C.byref(w), C.byref(e), C.byref(a)):
            return None
        return {"rank": r.value, "world": w.value, "halo_exchanges": e.value, "allreduces": a.value}

    def comm_timing(self, on=True):
        """(exchange_ms, allreduce_ms): device time of the band collectives since the last call; switches the timing on/off."""
        a = C.c_double(); b = C.c_double()
        self._ck(self.L.mimc3cu_comm_timing(self.h, int(on), C.byref(a), C.byref(b)))
        return a.value, b.value

    def comm_gather(self, send_dev, bytes_per_rank, recv_dev, root=0):
        b = np.ascontiguousarray(bytes_per_rank, dtype=np.int64)
        self._ck(self.L.mimc3cu_comm_gather(self.h, _ptr(send_dev), _ptr(b), _ptr(recv_dev), root))

    def postprocess_stage(self, which, n):
        out = np.empty(n, np.float32 if which in (2, 3, 5, 6) else np.int32)
        self._ck(self.L.mimc3cu_postprocess_stage(self.h, which, _ptr(out)))
        return out

    def fp32_peak(self):
        """(TFLOP/s, ms) of the FMA micro-benchmark."""
        a = C.c_double(); b = C.c_double()
        self._ck(self.L.mimc3cu_fp32_peak(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def timing_enable(self, on=True):
        self._ck(self.L.mimc3cu_timing_enable(self.h, int(on)))

    def timing_read(self):
        """(ms[3], counts[3]) for the families match / conv2 / postprocess since the last read."""
        ms = np.zeros(3, np.float64); cnt = np.zeros(3, np.int64)
        self._ck(self.L.mimc3cu_timing_read(self.h, _ptr(ms), _ptr(cnt)))
        return ms, cnt

    def finalize(self, planes_dev, params):
        a = C.c_float(); b = C.c_float()
        self._ck(self.L.mimc3cu_finalize(self.h, _ptr(planes_dev), C.byref(params), C.byref(a), C.byref(b)))
        return a.value, b.value
