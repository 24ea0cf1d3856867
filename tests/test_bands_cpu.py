"""CPU tests of the N > 1 host logic (mimc3_b200/bands.py): band partition, halo exchange,
OR-reduction of scattered flags and the counter all-reduce, over gloo with world_size 2 and 3
(one process per rank, 127.0.0.1 rendezvous) and over the in-process thread transport."""
import os
import socket
import threading

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from mimc3_b200 import bands


def test_bands_balanced_by_pivot_counts():
    """Rows crossing the fast band carry more DLC pivots, hence more NCC cells: bands cut by that weight are
    narrower there and their estimated work agrees to within one row."""
    from mimc3_b200 import lib, synth
    sc = synth.make_scene(H=1024, W=1024, dtype="u8", spacing=16, seed=11, peak_px=30.0, band_angle_deg=0.0, apriori_gain=0.9)
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    offs = [lib.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, 1024, 1024)[0] for ocw in (7, 15, 30, 40)]
    w = bands.row_work(offs, sc.dimx, sc.dimy)
    assert w.shape == (sc.dimy,) and w.max() > 2.0 * w.min()
    even = bands.split_rows(sc.dimy, 4, min_rows=5)
    bal = bands.split_rows(sc.dimy, 4, min_rows=5, weights=w)
    load = lambda parts: np.array([w[r0:r0 + n].sum() for r0, n in parts])
    assert sum(n for _, n in bal) == sc.dimy
    assert load(bal).max() - load(bal).min() <= 2 * w.max()
    assert load(bal).max() < load(even).max()
    with pytest.raises(ValueError):
        bands.row_work([offs[0][:-1]], sc.dimx, sc.dimy)


def test_split_rows_covers_and_balances():
    for dimy, world in ((813, 8), (100, 2), (40, 8), (17, 3)):
        parts = bands.split_rows(dimy, world, min_rows=5)
        assert parts[0][0] == 0 and sum(r for _, r in parts) == dimy
        assert all(parts[k][0] + parts[k][1] == parts[k + 1][0] for k in range(world - 1))
        assert min(r for _, r in parts) >= 5
    w = np.ones(100); w[40:60] = 10.0                   # fast-glacier rows are 10x dearer
    parts = bands.split_rows(100, 4, min_rows=5, weights=w)
    loads = [w[a:a + r].sum() for a, r in parts]
    assert max(loads) / (sum(loads) / 4) < 1.35
    with pytest.raises(ValueError):
        bands.split_rows(9, 2, min_rows=5)


def _global_fields(gdimy, dimx, seed=0):
    rng = np.random.default_rng(seed)
    f32 = rng.normal(size=(gdimy, dimx)).astype(np.float32)
    i32 = rng.integers(-5, 50, size=(gdimy, dimx)).astype(np.int32)
    u8 = (rng.random((gdimy, dimx)) < 0.3).astype(np.uint8)
    return f32, i32, u8


def _rank_body(transport, rank, world, gdimy, dimx, halo, parts):
    """What every rank checks; returns a list of failure strings."""
    fails = []
    f32, i32, u8 = _global_fields(gdimy, dimx)
    row0, rows = parts[rank]
    geo = bands.BandGeometry(dimx, gdimy, row0, rows, halo)
    lo = row0 - geo.ht

    def local(a):   # owned rows right, halo rows poisoned
        t = torch.from_numpy(a[lo:lo + geo.rows].copy())
        poison = torch.full_like(t, 77)
        poison[geo.own0:geo.own0 + rows] = t[geo.own0:geo.own0 + rows]
        return poison
    lf, li, lu = local(f32), local(i32), local(u8)
    views = [lf.view(torch.uint8).view(geo.rows, dimx * 4), li.view(torch.uint8).view(geo.rows, dimx * 4), lu.view(geo.rows, dimx)]
    bands.halo_exchange(transport, geo, views)
    for name, got, want in (("f32", lf, f32), ("i32", li, i32), ("u8", lu, u8)):
        if not np.array_equal(got.numpy(), want[lo:lo + geo.rows]):
            fails.append(f"rank {rank}: halo rows of {name} wrong after the exchange")
    # scatter: every rank marks flags in its whole local array; after the OR-reduce the owned rows
    # must equal the OR over all ranks that can see them
    rng = np.random.default_rng(100 + rank)
    mine = (rng.random((geo.rows, dimx)) < 0.2).astype(np.uint8)
    want = np.zeros((gdimy, dimx), np.uint8)
    for r in range(world):
        g2 = bands.BandGeometry(dimx, gdimy, parts[r][0], parts[r][1], halo)
        m2 = (np.random.default_rng(100 + r).random((g2.rows, dimx)) < 0.2).astype(np.uint8)
        l2 = parts[r][0] - g2.ht
        want[l2:l2 + g2.rows] |= m2
    t = torch.from_numpy(mine.copy())
    bands.halo_or_reduce(transport, geo, t)
    if not np.array_equal(t.numpy()[geo.own0:geo.own0 + rows], want[row0:row0 + rows]):
        fails.append(f"rank {rank}: OR-reduce of scattered flags wrong")
    c = torch.tensor([rank + 1, 10 * (rank + 1), 0], dtype=torch.int32)
    transport.allreduce_sum(c)
    s = world * (world + 1) // 2
    if c.tolist() != [s, 10 * s, 0]:
        fails.append(f"rank {rank}: all-reduce gave {c.tolist()}")
    return fails


def _gloo_worker(rank, world, port, gdimy, dimx, halo, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parts = bands.split_rows(gdimy, world, min_rows=halo)
        q.put(_rank_body(bands.DistTransport(), rank, world, gdimy, dimx, halo, parts))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", (2, 3))
def test_halo_exchange_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, 37, 11, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    fails = []
    for _ in procs:
        fails += q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert not fails, fails


def test_halo_exchange_over_threads():
    world, gdimy, dimx, halo = 4, 41, 7, 5
    shared = bands.ThreadTransport.Shared(world)
    parts = bands.split_rows(gdimy, world, min_rows=halo)
    out = [None] * world

    def run(r):
        out[r] = _rank_body(bands.ThreadTransport(shared, r), r, world, gdimy, dimx, halo, parts)
    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=60)
    assert all(o == [] for o in out), out


def test_split_rows_property():
    """Bands are contiguous, cover every row once and respect the minimum height for any weights."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=100, deadline=None)
    @hyp.given(st.integers(1, 8), st.integers(1, 6), st.integers(0, 40), st.integers(0, 2**31 - 1), st.booleans())
    def check(world, min_rows, extra, seed, weighted):
        dimy = world * min_rows + extra
        w = np.random.default_rng(seed).random(dimy) ** 4 * 100 if weighted else None
        parts = bands.split_rows(dimy, world, min_rows=min_rows, weights=w)
        assert len(parts) == world and parts[0][0] == 0
        assert all(n >= min_rows for _, n in parts)
        assert all(parts[k][0] + parts[k][1] == parts[k + 1][0] for k in range(world - 1))
        assert parts[-1][0] + parts[-1][1] == dimy

    check()
