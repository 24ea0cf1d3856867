"""CPU tests (no GPU): the restated C oracle against (a) the committed golden vectors, which are
outputs of the UNMODIFIED reference (scripts/make_golden.py), and (b) the reference build itself
where oracle/_ref exists.  The reference has no tests or golden files of its own,
so these runs of the reference are what pins the oracle."""
import hashlib
import os

import numpy as np
import pytest

import oracle
from tests.util import VEC_OCW, mismatch_report, same_bits_nan_aware, small_scene

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIXTURES = ("ref_u8_wedge", "ref_u16_fast")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["i0f"] = d["i0"].astype(np.float32); d["i1f"] = d["i1"].astype(np.float32)
    d["dimx"] = int(d["dimx"]); d["dimy"] = int(d["dimy"]); d["dt"] = float(d["dt"])
    return d


def oracle_multimatch(orc, g):
    i0, i1, xy, offset = g["i0f"], g["i1f"], g["xyuvav"], g["offset"]
    H, W = i0.shape
    dps, filt = [], []

    def attempts(a, b):
        for ocw in VEC_OCW:
            off, piv = g[f"piv_off_{ocw}"], g[f"piv_{ocw}"]
            o1, _, _ = orc.match(a, b, xy, offset, off, piv, +1, ocw)
            o2, _, _ = orc.match(b, a, xy, -offset, off, piv, -1, ocw)
            o2 = o2.copy(); o2[:, :2] = -o2[:, :2]
            dps.extend([o1, o2])
    attempts(i0, i1)
    c0 = np.zeros_like(i0); c1 = np.zeros_like(i1)
    for k in range(3):
        orc.conv2(i0, k, c0); orc.conv2(i1, k, c1)
        filt.append(c0.copy())
        attempts(c0, c1)
    return np.stack(dps), filt


@pytest.mark.parametrize("name", FIXTURES)
def test_pivots_match_reference(orc, name):
    g = load(name)
    pp = oracle.post_params(g["xyuvav"], g["dimx"], g["dimy"], g["dt"])
    H, W = g["i0"].shape
    for ocw in VEC_OCW:
        off, piv = orc.get_uv_pivot(g["xyuvav"], g["dt"], float(pp.mpp), ocw, H, W)
        assert np.array_equal(off, g[f"piv_off_{ocw}"]) and np.array_equal(piv, g[f"piv_{ocw}"]), ocw


@pytest.mark.parametrize("name", FIXTURES)
def test_matcher_and_conv2_match_reference(orc, name):
    """All 32 attempts (4 chip sizes x 2 directions x raw + 3 filtered pairs) bit-identical."""
    g = load(name)
    dp, filt = oracle_multimatch(orc, g)
    for k in range(3):
        assert hashlib.sha256(filt[k].tobytes()).hexdigest() == str(g["conv2_i0_sha256"][k]), f"conv2 kernel {k}"
        assert np.array_equal(filt[k][:96, :96], g["conv2_i0_crop"][k])
    for a in range(32):
        assert same_bits_nan_aware(dp[a], g["dp"][a]), f"attempt {a}: " + mismatch_report(dp[a], g["dp"][a])
    assert (g["dp"][:, :, 2] == -3).any() or name != "ref_u8_wedge"     # the wedge produces invalid nodes


@pytest.mark.parametrize("name", FIXTURES)
def test_postprocess_matches_reference(orc, name):
    g = load(name)
    pp = oracle.post_params(g["xyuvav"], g["dimx"], g["dimy"], g["dt"])
    mvn, ncl = orc.cluster(g["dp"])
    assert np.array_equal(ncl, g["ncl"])
    assert same_bits_nan_aware(mvn, g["mvn"]), mismatch_report(mvn, g["mvn"])
    st = orc.postprocess_stages(g["dp"], g["xyuvav"], pp)
    for key in ("dpf0", "dpf1_id", "ps_id"):
        assert np.array_equal(st[key], g["stage_" + key]), key
    for key in ("dpf1_dx", "dpf1_dy", "ps_dx", "ps_dy"):
        assert same_bits_nan_aware(st[key], g["stage_" + key]), key + ": " + mismatch_report(st[key], g["stage_" + key])
    planes = orc.postprocess(g["dp"], g["xyuvav"], pp)
    assert same_bits_nan_aware(planes, g["planes"]), mismatch_report(planes, g["planes"])


def test_golden_fixture_exercises_hole_filling():
    g = load("ref_u8_wedge")
    assert (g["stage_dpf0"] < 0).sum() > 0           # holes after the prominent-cluster pass
    assert (g["stage_dpf1_id"] >= 0).all() or (g["ncl"] == 0).any()


def test_find_ncc_peak_edge_cases(orc):
    """Degenerate inputs of find_ncc_peak: all-null chip (invalid, -3), constant chip (0/0 NCC)."""
    rng = np.random.default_rng(0)
    piv = np.array([(k, 0) for k in range(5)], np.int32)
    sa = rng.integers(1, 200, size=(2 * (0 + 7 + 2) + 1, 2 * (4 + 7 + 2) + 1)).astype(np.float32)
    chip = np.zeros((15, 15), np.float32)
    uv, pk, nc = orc.find_ncc_peak(chip, sa, piv)
    assert uv[2] == -3 and np.isnan(uv[0]) and np.isnan(uv[1])
    chip[:] = 7.0
    uv, pk, nc = orc.find_ncc_peak(chip, sa, piv)
    assert np.isnan(uv[2]) or uv[2] <= 1.0     # constant chip: zero variance => NaN cells never beat -2


def test_oracle_equals_reference_build_live(orc, ref):
    """Where the reference build is present: a fresh seeded scene, every chip size, both directions."""
    sc = small_scene(H=360, W=360, seed=77, spacing=31, null_wedge=True)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    H, W = i0.shape
    mpp = ref.set_globals(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    offset = np.array(sc.offset, np.int32)
    for ocw in VEC_OCW:
        off, piv = orc.get_uv_pivot(sc.xyuvav, sc.dt, mpp, ocw, H, W)
        off_r, piv_r = ref.get_uv_pivot(sc.xyuvav, sc.dt, ocw, H, W)
        assert np.array_equal(off, off_r) and np.array_equal(piv, piv_r)
        for sign, (a, b, o) in ((+1, (i0, i1, offset)), (-1, (i1, i0, -offset))):
            dpo, _, _ = orc.match(a, b, sc.xyuvav, o, off, piv, sign, ocw)
            dpr, _ = ref.match(a, b, sc.xyuvav, o, off, sign * piv, +1, ocw)
            assert same_bits_nan_aware(dpo, dpr), f"ocw {ocw} sign {sign}: " + mismatch_report(dpo, dpr)


@pytest.mark.parametrize("seed_time", (1700000000, 99))
def test_control_point_stage_equals_reference(orc, ref, seed_time):
    """get_offset_image (candidates, glibc-rand permutation, segments, tile-local conv2 on a reused
    output buffer, 16 attempts, clusters): same return code, offset and CP flags as the reference
    build with time() pinned to the seed."""
    sc = small_scene(H=600, W=600, seed=81, spacing=22, null_wedge=True, offset=(-2, 3))
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    ref.set_globals(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    rc_r, off_r, flag_r = ref.get_offset_image(i0, i1, sc.xyuvav, fake_time=seed_time)
    rc, off, flag = orc.get_offset_image(i0, i1, sc.xyuvav, seed_time)
    assert rc == rc_r == 1
    assert np.array_equal(off, off_r) and np.array_equal(off, np.array(sc.offset, np.int32))
    assert np.array_equal(flag, flag_r) and flag.sum() > 0


@pytest.mark.parametrize("maxj", (32, 64))
@pytest.mark.parametrize("case", ("wedge", "fast"))
def test_explore_replay_schedule_equals_reference_algorithm(orc, case, maxj):
    """The CUDA matcher evaluates cells in "explore" rounds and then replays the reference's state machine
    (csrc/match2.cu); oracle/leader_model.c restates that schedule on the CPU.  It must give the oracle's
    results bit for bit -- peaks, evaluated-cell counts, dp -- and never leave a needed cell unevaluated."""
    from mimc3_b200 import synth
    if case == "wedge":
        sc = synth.make_scene(H=448, W=448, dtype="u8", spacing=29, seed=5, peak_px=6.3, null_wedge=True)
    else:
        sc = synth.make_scene(H=512, W=512, dtype="u16", spacing=53, seed=3, peak_px=30.0, apriori_gain=0.9,
                              band_width_frac=0.2, decorrelated_patches=4)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    H, W = i0.shape
    mpp = float(np.float32((sc.xyuvav[1, 0] - sc.xyuvav[0, 0]) / (sc.xyuvav[1, 2] - sc.xyuvav[0, 2])))
    offset = np.array(sc.offset, np.int32)
    for ocw in (7, 30):
        off, piv = orc.get_uv_pivot(sc.xyuvav, sc.dt, mpp, ocw, H, W)
        for sign, a, b, o in ((1, i0, i1, offset), (-1, i1, i0, -offset)):
            dp, pk, nc = orc.match(a, b, sc.xyuvav, o, off, piv, sign, ocw)
            dm, pm, nm, st = orc.model_match(a, b, sc.xyuvav, o, off, piv, sign, ocw, maxj)
            assert (st[:, 0] >= 0).all(), "the schedule left a cell the replay needs unevaluated"
            assert np.array_equal(pk, pm) and np.array_equal(nc, nm)
            assert np.array_equal(np.isnan(dp), np.isnan(dm)) and np.array_equal(dp[~np.isnan(dp)], dm[~np.isnan(dm)])
            assert st[:, 1].sum() <= 1.02 * nc.sum() + 9 * len(nc)      # explore evaluates (almost) only what the reference does
