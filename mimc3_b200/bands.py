"""Node-row bands: how the hot path shards over the GPUs of one box.

Matching, clustering and the prominent-cluster pass are independent per node, so every rank (one
process per GPU) owns a contiguous band of node rows and runs them with no data-path collective.
The iterative stages of the postprocess (get_dpf1, MIMC_module.c:1387-1612; pseudosmoothing,
:2077-2288) read neighbours up to ``halo`` rows away, so each band keeps that many rows of its
neighbours and refreshes them after every committed sweep.  The sweep control lives in the
library (mimc3cu_postprocess_band); this module implements the three communication callbacks it
calls -- halo exchange, OR-reduction of the scattered dirty flags, and the all-reduce of the
per-sweep counters -- on top of a small *transport*:

* ``DistTransport``   torch.distributed (NCCL over NVLink on GPUs, gloo on CPU for the tests);
* ``ThreadTransport`` ranks emulated as threads of one process (one GPU, tests only).

Everything here is device-agnostic tensor code, so the N > 1 logic is covered on CPU with gloo.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch


def split_rows(dimy: int, world: int, min_rows: int = 1, weights=None):
    """Contiguous bands [(row0, rows)] covering [0, dimy).  ``weights`` (per node row, e.g. the
    number of DLC pivots of the row: matching cost is proportional to it) balances the bands by
    work instead of by row count; every band keeps at least ``min_rows`` rows."""
    if world < 1 or dimy < world * min_rows:
        raise ValueError(f"cannot split {dimy} node rows into {world} bands of >= {min_rows} rows")
    w = np.ones(dimy, np.float64) if weights is None else np.asarray(weights, np.float64)
    assert w.shape == (dimy,) and (w >= 0).all()
    cum = np.concatenate(([0.0], np.cumsum(w)))
    total = cum[-1] if cum[-1] > 0 else 1.0
    cuts = [0]
    for r in range(1, world):
        c = int(np.searchsorted(cum, total * r / world, side="left"))
        c = max(c, cuts[-1] + min_rows)                  # at least min_rows for the band that ends here
        c = min(c, dimy - (world - r) * min_rows)        # ... and for every band still to come
        cuts.append(c)
    cuts.append(dimy)
    return [(cuts[r], cuts[r + 1] - cuts[r]) for r in range(world)]


def row_work(csr_offsets, dimx: int, dimy: int) -> np.ndarray:
    """Matching cost estimate per node row, known before any matching: a node-attempt evaluates about
    3 * (P + 2) correlation cells for P DLC pivots, so fast-glacier rows are several times
    dearer than static ones.  The weight is P itself, calibrated on the 8-GPU run of one 32768^2 scene
    (profiles/): every node-attempt also has a fixed cost (staging, the serial replay), which argues for P + c
    with c > 0, but nodes with long pivot lines run in the less efficient wide-search-area bins and need more
    evaluation rounds, which outweighs it -- with P + 2 / P + 5 / P + 12 the fast-ice bands finished
    7 / 10 / 15 % after the slow-ice ones.
    ``csr_offsets``: one int array (n + 1) per chip size, as get_uv_pivot returns them.  Feed the result to
    ``split_rows(weights=...)``."""
    w = np.zeros(dimy, np.float64)
    for off in csr_offsets:
        P = np.diff(np.asarray(off, dtype=np.int64))
        if P.shape[0] != dimx * dimy:
            raise ValueError(f"{P.shape[0]} nodes do not form a {dimy} x {dimx} grid")
        w += P.reshape(dimy, dimx).sum(axis=1)
    return w


class BandGeometry:
    """Local array layout of one band (mirrors the library's ``Band`` struct, csrc/post.cu)."""

    def __init__(self, dimx, gdimy, own_row0, own_rows, halo, single=False):
        self.dimx, self.gdimy, self.own_row0, self.own_rows, self.halo = dimx, gdimy, own_row0, own_rows, halo
        self.ht = 0 if single else min(halo, own_row0)
        self.hb = 0 if single else min(halo, gdimy - own_row0 - own_rows)
        self.rows = own_rows + self.ht + self.hb
        self.own0 = self.ht


# ------------------------------------------------------------------------------------------------
# transports
# ------------------------------------------------------------------------------------------------
class DistTransport:
    """Neighbour exchange and all-reduce over a torch.distributed process group."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def sendrecv(self, to_up, to_down, from_up_like, from_down_like):
        """Send ``to_up`` to rank-1 and ``to_down`` to rank+1; receive into the ``*_like`` buffers
        (any of the four may be None at the grid edges).  Returns (from_up, from_down)."""
        dist = self.dist
        ops = []
        up, down = self.rank - 1, self.rank + 1
        if to_up is not None:
            ops.append(dist.P2POp(dist.isend, to_up, up, self.group))
        if from_up_like is not None:
            ops.append(dist.P2POp(dist.irecv, from_up_like, up, self.group))
        if to_down is not None:
            ops.append(dist.P2POp(dist.isend, to_down, down, self.group))
        if from_down_like is not None:
            ops.append(dist.P2POp(dist.irecv, from_down_like, down, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return from_up_like, from_down_like

    def allreduce_sum(self, t):
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t

    def gather_rows(self, t, dst=0):
        """Gather equally-shaped tensors to ``dst`` (list on dst, None elsewhere)."""
        parts = [torch.empty_like(t) for _ in range(self.world)] if self.rank == dst else None
        self.dist.gather(t, parts, dst=dst, group=self.group)
        return parts


class ThreadTransport:
    """Ranks as threads of one process sharing a mailbox (tests on a single GPU / CPU)."""

    class Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.box = {}
            self.acc = None
            self.lock = threading.Lock()

    def __init__(self, shared, rank):
        self.s, self.rank, self.world = shared, rank, shared.world

    def sendrecv(self, to_up, to_down, from_up_like, from_down_like):
        s = self.s
        if to_up is not None:
            s.box[(self.rank, self.rank - 1)] = to_up.clone()
        if to_down is not None:
            s.box[(self.rank, self.rank + 1)] = to_down.clone()
        s.barrier.wait()
        if from_up_like is not None:
            from_up_like.copy_(s.box[(self.rank - 1, self.rank)])
        if from_down_like is not None:
            from_down_like.copy_(s.box[(self.rank + 1, self.rank)])
        s.barrier.wait()
        return from_up_like, from_down_like

    def allreduce_sum(self, t):
        s = self.s
        with s.lock:
            s.acc = t.clone() if s.acc is None else s.acc + t.to(s.acc.device)
        s.barrier.wait()
        t.copy_(s.acc)
        s.barrier.wait()
        if self.rank == 0:
            s.acc = None
        s.barrier.wait()
        return t


# ------------------------------------------------------------------------------------------------
# the three communication steps on tensors (rows x row_bytes, uint8 views of the local arrays)
# ------------------------------------------------------------------------------------------------
def halo_exchange(transport, geo: BandGeometry, arrays):
    """``arrays``: list of 2-D uint8 tensors (geo.rows, row_bytes) = the band's local arrays.
    Overwrites the halo rows with the neighbours' owned rows; one packed message per neighbour."""
    h = geo.halo
    up_ok, down_ok = geo.ht > 0, geo.hb > 0
    if not (up_ok or down_ok):
        return
    assert geo.own_rows >= h, "a band needs at least `halo` rows"
    own0, own1 = geo.own0, geo.own0 + geo.own_rows
    to_up = torch.cat([a[own0:own0 + h].reshape(-1) for a in arrays]) if up_ok else None
    to_down = torch.cat([a[own1 - h:own1].reshape(-1) for a in arrays]) if down_ok else None
    from_up = torch.empty_like(to_up) if up_ok else None
    from_down = torch.empty_like(to_down) if down_ok else None
    transport.sendrecv(to_up, to_down, from_up, from_down)
    pos_u = pos_d = 0
    for a in arrays:
        nb = h * a.shape[1]
        if up_ok:
            a[0:h] = from_up[pos_u:pos_u + nb].view(h, a.shape[1]); pos_u += nb
        if down_ok:
            a[own1:own1 + h] = from_down[pos_d:pos_d + nb].view(h, a.shape[1]); pos_d += nb


def halo_or_reduce(transport, geo: BandGeometry, flags):
    """``flags``: uint8 (geo.rows, dimx).  Flags this band scattered into its halo rows belong to the
    neighbours: send them there and OR what the neighbours scattered into our rows."""
    h = geo.halo
    up_ok, down_ok = geo.ht > 0, geo.hb > 0
    if not (up_ok or down_ok):
        return
    own0, own1 = geo.own0, geo.own0 + geo.own_rows
    to_up = flags[0:h].contiguous() if up_ok else None
    to_down = flags[own1:own1 + h].contiguous() if down_ok else None
    from_up = torch.empty_like(to_up) if up_ok else None
    from_down = torch.empty_like(to_down) if down_ok else None
    transport.sendrecv(to_up, to_down, from_up, from_down)
    if up_ok:
        flags[own0:own0 + h] |= from_up
    if down_ok:
        flags[own1 - h:own1] |= from_down


# ------------------------------------------------------------------------------------------------
# ctypes glue: mimc3cu_band_comm callbacks over device pointers
# ------------------------------------------------------------------------------------------------
_HALO_EX = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32)
_OR_RED = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)
_ALLRED = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_int32), C.c_int32)


class BandCommStruct(C.Structure):
    _fields_ = [("user", C.c_void_p), ("halo_exchange", _HALO_EX), ("halo_or_reduce", _OR_RED), ("allreduce_sum", _ALLRED)]


class _DevPtr:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_bytes(ptr: int, nbytes: int, device) -> torch.Tensor:
    """Zero-copy uint8 view of library-owned device memory."""
    return torch.as_tensor(_DevPtr(ptr, nbytes), device=device)


class BandComm:
    """Owns the ctypes callbacks handed to mimc3cu_postprocess_band for one band."""

    def __init__(self, transport, geo: BandGeometry, device, stream=None):
        self.t, self.geo, self.device, self.stream = transport, geo, device, stream
        self.error = None
        self.n_exchanges = self.n_allreduce = 0

        def guard(fn):
            def wrapped(*a):
                try:
                    if self.stream is not None:
                        with torch.cuda.stream(self.stream):
                            fn(*a)
                            self.stream.synchronize()
                    else:
                        fn(*a)
                    return 0
                except Exception as e:   # noqa: BLE001 -- reported through the C return code
                    self.error = e
                    return 1
            return wrapped

        def ex(_user, arrays, elem_bytes, count):
            g = self.geo
            views = [device_bytes(arrays[k], g.rows * g.dimx * elem_bytes[k], self.device).view(g.rows, g.dimx * elem_bytes[k])
                     for k in range(count)]
            halo_exchange(self.t, g, views)
            self.n_exchanges += 1

        def orr(_user, ptr):
            g = self.geo
            halo_or_reduce(self.t, g, device_bytes(ptr, g.rows * g.dimx, self.device).view(g.rows, g.dimx))
            self.n_exchanges += 1

        def allred(_user, vals, count):
            host = np.ctypeslib.as_array(vals, shape=(count,))
            t = torch.from_numpy(host.copy()).to(self.device)
            self.t.allreduce_sum(t)
            host[:] = t.cpu().numpy()
            self.n_allreduce += 1

        self._cbs = (_HALO_EX(guard(ex)), _OR_RED(guard(orr)), _ALLRED(guard(allred)))   # keep alive
        self.struct = BandCommStruct(None, *self._cbs)
