"""GPU parity of the DLC-NCC matcher (through the C ABI) against the oracle.

Bar (BASELINE.json north_star): integer peaks and valid/null flags bit-exact, sub-pixel
displacements within 0.01 px, NCC within 1e-5 relative.  The CUDA path is built to be
bit-identical in all three, so the tests assert bit equality (NaN-aware) and would also
report the tolerance-level numbers on failure.
"""
import numpy as np
import pytest

from mimc3_b200 import lib, synth
from tests.util import VEC_OCW, mismatch_report, same_bits_nan_aware, small_scene

pytestmark = pytest.mark.gpu


MATCHERS = ("v1", "v2")   # general FP64 kernel / exact-FP32 kernel with summed-area tables


def _run_both(gpu_ctx, orc, sc, a, b, offset, sign, ocw, slot=0, matcher="v2"):
    H, W = a.shape
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    off, piv = lib.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, H, W)
    off_o, piv_o = orc.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, H, W)
    assert np.array_equal(off, off_o) and np.array_equal(piv, piv_o)
    gpu_ctx.set_nodes(sc.xyuvav)
    gpu_ctx.set_pivots(slot, off, piv)
    ia, ib = gpu_ctx.image_from(a), gpu_ctx.image_from(b)
    gpu_ctx.set_matcher(matcher)
    try:
        dp, peak, ncell = gpu_ctx.match(ia, ib, offset, slot, sign, ocw)
        assert gpu_ctx.last_matcher() == (2 if matcher == "v2" else 1)
    finally:
        gpu_ctx.set_matcher("auto")
        gpu_ctx.image_destroy(ia); gpu_ctx.image_destroy(ib)
    dpo, peako, ncello = orc.match(a, b, sc.xyuvav, offset, off, piv, sign, ocw)
    return (dp, peak, ncell), (dpo, peako, ncello)


def _assert_parity(got, want, tag):
    (dp, peak, ncell), (dpo, peako, ncello) = got, want
    assert np.array_equal(dpo[:, 2] == -3, dp[:, 2] == -3), tag + " validity flags differ"
    assert np.array_equal(peak, peako), tag + " integer peaks differ: " + mismatch_report(peak, peako)
    assert np.array_equal(ncell, ncello), tag + " evaluated-cell counts differ"
    fin = np.isfinite(dpo[:, 0]) & np.isfinite(dpo[:, 1])
    assert np.array_equal(fin, np.isfinite(dp[:, 0]) & np.isfinite(dp[:, 1]))
    assert np.max(np.abs(dp[fin, :2] - dpo[fin, :2]), initial=0) <= 0.01, tag       # sub-pixel tolerance
    assert np.allclose(dp[:, 2], dpo[:, 2], rtol=1e-5, atol=0), tag                 # NCC tolerance
    assert same_bits_nan_aware(dp, dpo), tag + " not bit-identical: " + mismatch_report(dp, dpo)


@pytest.mark.parametrize("matcher", MATCHERS)
@pytest.mark.parametrize("ocw", VEC_OCW)
@pytest.mark.parametrize("direction", ["fwd", "swapped"])
def test_match_u8_with_null_wedge(gpu_ctx, orc, ocw, direction, matcher):
    sc = small_scene()
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    offset = np.array(sc.offset, np.int32)
    if direction == "fwd":
        got, want = _run_both(gpu_ctx, orc, sc, i0, i1, offset, +1, ocw, matcher=matcher)
    else:
        got, want = _run_both(gpu_ctx, orc, sc, i1, i0, -offset, -1, ocw, matcher=matcher)
    _assert_parity(got, want, f"u8 ocw={ocw} {direction} {matcher}")
    assert (want[0][:, 2] == -3).sum() > 0 or ocw >= 30      # the wedge invalidates some small-chip nodes


@pytest.mark.parametrize("matcher", MATCHERS)
@pytest.mark.parametrize("ocw", (7, 15, 30, 40))
def test_match_u16(gpu_ctx, orc, ocw, matcher):
    """uint16-range data: float products exceed 2^24 and are rounded like the reference's."""
    sc = small_scene(dtype="u16", seed=9, null_wedge=False)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    i0 = np.minimum(i0 * 3.9, 65535).round().astype(np.float32)    # reach DN > 60000
    i1 = np.minimum(i1 * 3.9, 65535).round().astype(np.float32)
    assert i0.max() > 60000
    got, want = _run_both(gpu_ctx, orc, sc, i0, i1, np.array(sc.offset, np.int32), +1, ocw, matcher=matcher)
    _assert_parity(got, want, f"u16 ocw={ocw} {matcher}")


@pytest.mark.parametrize("matcher", MATCHERS)
@pytest.mark.parametrize("ocw", (15, 40))
@pytest.mark.parametrize("kid", (0, 1, 2))
def test_match_on_filtered_images(gpu_ctx, orc, kid, ocw, matcher):
    """conv2-filtered inputs (multiples of 1/8 for the Laplacian), CPU-filtered so that only
    the matcher is under test."""
    import oracle
    sc = small_scene(seed=21)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    c0 = np.zeros_like(i0); c1 = np.zeros_like(i1)
    orc.conv2(i0, kid, c0); orc.conv2(i1, kid, c1)
    got, want = _run_both(gpu_ctx, orc, sc, c0, c1, np.array(sc.offset, np.int32), +1, ocw, matcher=matcher)
    _assert_parity(got, want, f"filtered kernel {kid} ocw={ocw} {matcher}")


@pytest.mark.parametrize("matcher", MATCHERS)
def test_match_sentinel2_like_ragged_image(gpu_ctx, orc, matcher):
    """BASELINE.json configs[2] in small: 10 m pixels, dense 100 m (10 px) node spacing, u16, and an image
    whose width and height are odd (row pitch not a multiple of 16 bytes, like the 10980-px tiles are not a
    multiple of 128) -- staging and the summed-area tables must not assume alignment."""
    sc = small_scene(H=549, W=613, dtype="u16", spacing=10, mpp=10.0, seed=31, null_wedge=True)
    assert sc.n > 2000
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    for ocw in VEC_OCW:
        got, want = _run_both(gpu_ctx, orc, sc, i0, i1, np.array(sc.offset, np.int32), +1, ocw, matcher=matcher)
        _assert_parity(got, want, f"sentinel2-like ocw={ocw} {matcher}")


@pytest.mark.parametrize("kid", (0, 2))
def test_match_on_filtered_u16_images(gpu_ctx, orc, kid):
    """14-bit DN filtered by d/dx and by the Laplacian (values up to 2^18 in units of 1/8):
    still inside the exact-FP32 class (product bits + fraction bits <= 37)."""
    sc = small_scene(dtype="u16", seed=23, null_wedge=True)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    c0 = np.zeros_like(i0); c1 = np.zeros_like(i1)
    orc.conv2(i0, kid, c0); orc.conv2(i1, kid, c1)
    for ocw in (15, 40):
        got, want = _run_both(gpu_ctx, orc, sc, c0, c1, np.array(sc.offset, np.int32), +1, ocw, matcher="v2")
        _assert_parity(got, want, f"filtered u16 kernel {kid} ocw={ocw}")


def test_general_float_images_use_the_general_kernel(gpu_ctx, orc):
    """Arbitrary float pixels are outside the exact class: auto mode must pick the FP64 kernel,
    and requiring v2 must fail loudly rather than return inexact numbers."""
    from mimc3_b200.lib import Mimc3CuError
    sc = small_scene(seed=29, null_wedge=False)
    i0 = (sc.i0.numpy() * np.float32(1.37)).astype(np.float32); i1 = (sc.i1.numpy() * np.float32(1.37)).astype(np.float32)
    H, W = i0.shape
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    off, piv = lib.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, 15, H, W)
    gpu_ctx.set_nodes(sc.xyuvav); gpu_ctx.set_pivots(0, off, piv)
    ia, ib = gpu_ctx.image_from(i0), gpu_ctx.image_from(i1)
    try:
        assert gpu_ctx.image_class(ia)[0] is False
        gpu_ctx.match(ia, ib, np.zeros(2, np.int32), 0, +1, 15)
        assert gpu_ctx.last_matcher() == 1
        gpu_ctx.set_matcher("v2")
        with pytest.raises(Mimc3CuError):
            gpu_ctx.match(ia, ib, np.zeros(2, np.int32), 0, +1, 15)
    finally:
        gpu_ctx.set_matcher("auto")
        gpu_ctx.image_destroy(ia); gpu_ctx.image_destroy(ib)


@pytest.mark.parametrize("matcher", MATCHERS)
def test_match_fast_glacier_wide_windows(gpu_ctx, orc, matcher):
    """Config-4-like: ~40 px a-priori displacement => ~80 pivots and search areas that do not
    fit the small-window fast sizes."""
    sc = small_scene(H=900, W=900, seed=33, peak_px=43.0, apriori_gain=0.9, spacing=41, null_wedge=False,
                     band_width_frac=0.2)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    for ocw in (40, 15):
        got, want = _run_both(gpu_ctx, orc, sc, i0, i1, np.array(sc.offset, np.int32), +1, ocw, matcher=matcher)
        assert want[2].max() > 200     # many evaluated cells
        _assert_parity(got, want, f"fast glacier ocw={ocw} {matcher}")


@pytest.mark.parametrize("angle", (30.0, 120.0, 200.0, 290.0))
def test_match_flow_into_all_quadrants(gpu_ctx, orc, angle):
    """Fast flow into each of the four quadrants, forward and swapped pass (pivots negated): the pivot line, hence
    the part of the search area that is actually visited, sits in every corner of the (symmetric) area."""
    sc = small_scene(H=760, W=760, seed=37, peak_px=30.0, apriori_gain=0.9, spacing=57, null_wedge=False,
                     band_width_frac=0.3, band_angle_deg=angle, margin=90)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    offset = np.array(sc.offset, np.int32)
    for ocw in (30, 7):
        got, want = _run_both(gpu_ctx, orc, sc, i0, i1, offset, +1, ocw)
        assert want[2].max() > 100
        _assert_parity(got, want, f"quadrant {angle} fwd ocw={ocw}")
        got, want = _run_both(gpu_ctx, orc, sc, i1, i0, -offset, -1, ocw)
        _assert_parity(got, want, f"quadrant {angle} swapped ocw={ocw}")


def test_match_long_climbs_from_a_poor_apriori(gpu_ctx, orc):
    """A-priori velocity far off the truth (half of it, 15 px short): climbs of many steps away from the pivot line."""
    sc = small_scene(H=760, W=760, seed=39, peak_px=30.0, apriori_gain=0.5, spacing=57, null_wedge=False,
                     band_width_frac=0.3, band_angle_deg=60.0, margin=90)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    for ocw in (40, 15):
        got, want = _run_both(gpu_ctx, orc, sc, i0, i1, np.array(sc.offset, np.int32), +1, ocw)
        _assert_parity(got, want, f"long climbs ocw={ocw}")


@pytest.mark.parametrize("matcher", MATCHERS)
def test_match_nodes_at_image_border(gpu_ctx, orc, matcher):
    """Search areas hanging over the image edge are zero-filled (extract_sarea boundary check)."""
    sc = small_scene(H=400, W=400, seed=41, null_wedge=False, margin=44, spacing=39, peak_px=9.0)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    for ocw in (40, 7):
        got, want = _run_both(gpu_ctx, orc, sc, i0, i1, np.array((7, -6), np.int32), +1, ocw, matcher=matcher)
        _assert_parity(got, want, f"border nodes ocw={ocw} {matcher}")


def test_find_ncc_peak_batch_cp_shape(gpu_ctx, orc):
    """The CP stage's call shape: 85x85 search chips, 61x61 / 31x31 reference chips, 21x21
    rectangular pivot set (MIMC_module.c:165-176, 351)."""
    sc = small_scene(seed=55, null_wedge=False)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    piv = np.array([(a, b) for a in range(-10, 11) for b in range(-10, 11)], np.int32)
    rng = np.random.default_rng(3)
    cs = rng.integers(100, 540, size=(6, 2))
    for ocw in (15, 30):
        S = 2 * ocw + 1
        chips = np.stack([i0[v - ocw:v + ocw + 1, u - ocw:u + ocw + 1] for u, v in cs])
        sas = np.stack([i1[v - 42:v + 43, u - 42:u + 43] for u, v in cs])
        uv, pk, nc = gpu_ctx.find_ncc_peak_batch(chips, sas, piv)
        for k in range(len(cs)):
            uvo, pko, nco = orc.find_ncc_peak(chips[k], sas[k], piv)
            assert np.array_equal(pk[k], pko) and nc[k] == nco
            assert same_bits_nan_aware(uv[k], uvo), (ocw, k, uv[k], uvo)


@pytest.mark.parametrize("matcher", MATCHERS)
def test_ragged_and_empty_pivot_lists(gpu_ctx, orc, matcher):
    """Ragged CSR: nodes with 0 pivots (undefined behaviour in the reference, MIMC_module.c:589-591;
    defined as "nothing evaluable": NaN, NaN, -2), with a single pivot, and with long lists, in one call;
    also a one-node grid."""
    sc = small_scene(seed=47, null_wedge=False)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    H, W = i0.shape
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    off, piv = lib.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, 15, H, W)
    # rebuild the CSR: every 5th node loses all pivots, every 7th keeps only the first
    lists = [piv[off[g]:off[g + 1]] for g in range(sc.n)]
    for g in range(sc.n):
        if g % 5 == 0:
            lists[g] = lists[g][:0]
        elif g % 7 == 0:
            lists[g] = lists[g][:1]
    off2 = np.zeros(sc.n + 1, np.int32); off2[1:] = np.cumsum([len(l) for l in lists])
    piv2 = np.concatenate([l for l in lists if len(l)]).astype(np.int32)
    for xy, o, pv in ((sc.xyuvav, off2, piv2), (sc.xyuvav[37:38], np.array([0, len(lists[37])], np.int32), lists[37].astype(np.int32))):
        gpu_ctx.set_nodes(xy); gpu_ctx.set_pivots(0, o, pv)
        a, b = gpu_ctx.image_from(i0), gpu_ctx.image_from(i1)
        gpu_ctx.set_matcher(matcher)
        try:
            dp, peak, ncell = gpu_ctx.match(a, b, np.array(sc.offset, np.int32), 0, +1, 15)
        finally:
            gpu_ctx.set_matcher("auto")
            gpu_ctx.image_destroy(a); gpu_ctx.image_destroy(b)
        dpo, peako, ncello = orc.match(i0, i1, xy, np.array(sc.offset, np.int32), o, pv, +1, 15)
        _assert_parity((dp, peak, ncell), (dpo, peako, ncello), f"ragged {matcher} n={len(xy)}")
        empty = np.diff(o) == 0
        assert (dp[empty, 2] == -2).all() and np.isnan(dp[empty, 0]).all()
