"""GPU test of the banded postprocess (mimc3cu_postprocess_band + mimc3_b200/bands.py): the node
grid split into 2, 3 and 4 bands -- ranks emulated as threads with their own contexts on one GPU --
must give exactly the single-band result (same ids, same sweeps, same planes), including the
cross-band hole filling and pseudosmoothing."""
import threading

import numpy as np
import pytest
import torch

from mimc3_b200 import bands, lib
from tests.test_oracle_cpu import load
from tests.test_post_gpu import oracle_multimatch
from tests.util import small_scene

pytestmark = pytest.mark.gpu


def run_banded(dp, xyuvav, dimx, dimy, dt, world):
    p = lib.params_for(xyuvav, dimx, dimy, dt)
    halo = lib.band_halo(p)
    parts = bands.split_rows(dimy, world, min_rows=halo)
    shared = bands.ThreadTransport.Shared(world)
    dev = torch.device("cuda", 0)
    out, stats, stages, errs = [None] * world, [None] * world, [None] * world, []

    def rank_main(r):
        try:
            ctx = lib.Context(0)
            row0, rows = parts[r]
            geo = bands.BandGeometry(dimx, dimy, row0, rows, halo)
            comm = bands.BandComm(bands.ThreadTransport(shared, r), geo, dev)
            d = torch.from_numpy(np.ascontiguousarray(dp[:, row0 * dimx:(row0 + rows) * dimx])).to(dev)
            planes = torch.empty((5, rows, dimx), dtype=torch.float32, device=dev)
            stats[r] = ctx.postprocess_band(d, xyuvav, p, row0, rows, comm, planes)
            out[r] = planes.cpu().numpy()
            stages[r] = {w: ctx.postprocess_stage(w, rows * dimx) for w in (0, 1, 4)}
            ctx.close()
        except Exception as e:   # noqa: BLE001
            errs.append((r, e))
            shared.barrier.abort()
    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not errs, errs
    return np.concatenate(out, axis=1), stats, {w: np.concatenate([s[w] for s in stages]) for w in (0, 1, 4)}


@pytest.mark.parametrize("world", (2, 3))
def test_banded_postprocess_equals_reference_golden(world):
    g = load("ref_u8_wedge")
    planes, stats, st = run_banded(g["dp"], g["xyuvav"], g["dimx"], g["dimy"], g["dt"], world)
    assert np.array_equal(st[0], g["stage_dpf0"]) and np.array_equal(st[1], g["stage_dpf1_id"]) and np.array_equal(st[4], g["stage_ps_id"])
    a, b = planes, g["planes"]
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    assert all((s == stats[0]).all() for s in stats)          # every band saw the same global sweep counts


@pytest.mark.parametrize("world", (2, 4))
def test_banded_postprocess_equals_single_band(orc, gpu_ctx, world):
    sc = small_scene(H=700, W=700, seed=17, spacing=21, decorrelated_patches=8, null_wedge=True)
    offset = np.array(sc.offset, np.int32)
    dp = oracle_multimatch(orc, sc, offset)
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    single = torch.empty((5, sc.dimy, sc.dimx), dtype=torch.float32, device="cuda")
    stats1 = gpu_ctx.postprocess(torch.from_numpy(dp).cuda(), sc.xyuvav, p, single)
    ids1 = {w: gpu_ctx.postprocess_stage(w, sc.n) for w in (0, 1, 4)}
    assert stats1[0] > 5 and stats1[1] >= 2 and stats1[2] > 20   # hole-filling sweeps, pseudosmoothing sweeps, holes
    planes, stats, st = run_banded(dp, sc.xyuvav, sc.dimx, sc.dimy, sc.dt, world)
    for w in (0, 1, 4):
        assert np.array_equal(st[w], ids1[w]), w
    assert np.array_equal(stats[0][:3], stats1[:3]), (stats[0], stats1)
    a, b = planes, single.cpu().numpy()
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def run_banded_nccl(dp, xyuvav, dimx, dimy, dt, world):
    """The library's own communicator (csrc/comm.cu): one context per GPU, NCCL between them, one host thread per band."""
    p = lib.params_for(xyuvav, dimx, dimy, dt)
    parts = bands.split_rows(dimy, world, min_rows=lib.band_halo(p))
    ctxs = [lib.Context(r) for r in range(world)]
    lib.comm_init_all(ctxs)
    out, stats, stages, errs = [None] * world, [None] * world, [None] * world, []

    def rank_main(r):
        try:
            dev = torch.device("cuda", r)
            row0, rows = parts[r]
            d = torch.from_numpy(np.ascontiguousarray(dp[:, row0 * dimx:(row0 + rows) * dimx])).to(dev)
            planes = torch.empty((5, rows, dimx), dtype=torch.float32, device=dev)
            torch.cuda.synchronize(dev)
            stats[r] = ctxs[r].postprocess_band(d, xyuvav, p, row0, rows, None, planes)
            out[r] = planes.cpu().numpy()
            stages[r] = {w: ctxs[r].postprocess_stage(w, rows * dimx) for w in (0, 1, 4)}
        except Exception as e:   # noqa: BLE001
            errs.append((r, e))
    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    info = [c.comm_info() for c in ctxs]
    for c in ctxs:
        c.close()
    assert not errs, errs
    return np.concatenate(out, axis=1), stats, {w: np.concatenate([s[w] for s in stages]) for w in (0, 1, 4)}, info


@pytest.mark.parametrize("world", (2, 4, 8))
def test_library_nccl_bands_equal_single_band(orc, gpu_ctx, world):
    """Node-row bands on `world` GPUs with the halo exchange inside the library (grouped ncclSend/ncclRecv, OR of the
    scattered dirty flags, ncclAllReduce of the sweep counters) == the single-GPU postprocess, bit for bit."""
    if lib.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    sc = small_scene(H=900, W=700, seed=17, spacing=17, decorrelated_patches=10, null_wedge=True)
    offset = np.array(sc.offset, np.int32)
    dp = oracle_multimatch(orc, sc, offset)
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    single = torch.empty((5, sc.dimy, sc.dimx), dtype=torch.float32, device="cuda:0")
    stats1 = gpu_ctx.postprocess(torch.from_numpy(dp).cuda(0), sc.xyuvav, p, single)
    ids1 = {w: gpu_ctx.postprocess_stage(w, sc.n) for w in (0, 1, 4)}
    assert stats1[0] > 5 and stats1[1] >= 2
    planes, stats, st, info = run_banded_nccl(dp, sc.xyuvav, sc.dimx, sc.dimy, sc.dt, world)
    for w in (0, 1, 4):
        assert np.array_equal(st[w], ids1[w]), w
    assert np.array_equal(stats[0][:3], stats1[:3]), (stats[0], stats1)
    a, b = planes, single.cpu().numpy()
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    assert all(i["world"] == world and i["halo_exchanges"] > 0 and i["allreduces"] > 0 for i in info)
