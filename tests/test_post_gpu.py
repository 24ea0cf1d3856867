"""GPU parity of the multi-match + postprocess chain against the oracle.

dp (32 attempts) must be bit-identical.  The postprocess stages are bit-identical except
where glibc's expf/exp and CUDA's differ in the last ulp (see mimc3_b200/csrc/post.cu);
the discrete outputs (cluster ids) must agree and float fields agree to 1e-5.
"""
import numpy as np
import pytest
import torch

import oracle
from mimc3_b200 import lib, pipeline
from tests.util import VEC_OCW, mismatch_report, same_bits_nan_aware, small_scene

pytestmark = pytest.mark.gpu


def oracle_multimatch(orc, sc, offset):
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    H, W = i0.shape
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    pivs = [orc.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, H, W) for ocw in VEC_OCW]
    dps = []

    def attempts(a, b):
        for (off, piv), ocw in zip(pivs, VEC_OCW):
            o1, _, _ = orc.match(a, b, sc.xyuvav, offset, off, piv, +1, ocw)
            o2, _, _ = orc.match(b, a, sc.xyuvav, -offset, off, piv, -1, ocw)
            o2 = o2.copy(); o2[:, :2] = -o2[:, :2]
            dps.extend([o1, o2])
    attempts(i0, i1)
    c0 = np.zeros_like(i0); c1 = np.zeros_like(i1)
    for k in range(3):
        orc.conv2(i0, k, c0); orc.conv2(i1, k, c1)
        attempts(c0, c1)
    return np.stack(dps)


@pytest.fixture(scope="module")
def scene_and_dp(orc):
    sc = small_scene(H=700, W=700, seed=17, spacing=21, decorrelated_patches=8, null_wedge=True)
    offset = np.array(sc.offset, np.int32)
    return sc, offset, oracle_multimatch(orc, sc, offset)


def test_multimatch_32_attempts_bitwise(scene_and_dp):
    sc, offset, dpo = scene_and_dp
    pl = pipeline.Pipeline(0)
    try:
        pl.set_images(sc.i0.numpy().astype(np.uint8), sc.i1.numpy().astype(np.uint8))   # u8 ingest path
        pl.set_grid(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
        dp, ncell = pl.multimatch(offset, want_ncell=True)
        pl.ctx.sync()
        dp = dp.cpu().numpy()
        for a in range(32):
            assert same_bits_nan_aware(dp[a], dpo[a]), f"attempt {a}: " + mismatch_report(dp[a], dpo[a])
        assert int(ncell.min()) >= 0
    finally:
        pl.close()


def test_cluster_bitwise(scene_and_dp, orc, gpu_ctx):
    sc, offset, dpo = scene_and_dp
    n = sc.n
    mvn_o, ncl_o = orc.cluster(dpo)
    dp_d = torch.from_numpy(dpo).cuda()
    mvn = torch.empty((n, 32, 5), dtype=torch.float32, device="cuda"); ncl = torch.empty(n, dtype=torch.int32, device="cuda")
    gpu_ctx.cluster_async(dp_d, n, 32, mvn, ncl); gpu_ctx.sync()
    assert np.array_equal(ncl.cpu().numpy(), ncl_o)
    assert same_bits_nan_aware(mvn.cpu().numpy(), mvn_o), mismatch_report(mvn.cpu().numpy(), mvn_o)


def test_cluster_nan_candidate_defined(orc, gpu_ctx):
    """H11: a candidate with NaN displacement but ncc > 0.1 consumes an id and forms no cluster."""
    rng = np.random.default_rng(0)
    n = 64
    dp = np.zeros((32, n, 3), np.float32)
    dp[:, :, :2] = rng.normal(0, 0.2, size=(32, n, 2)); dp[:, :, 2] = rng.uniform(0, 1, size=(32, n))
    dp[3, ::2, :2] = np.nan; dp[31, ::3, 0] = np.nan; dp[0, 5, :2] = np.nan
    mvn_o, ncl_o = orc.cluster(dp)
    d = torch.from_numpy(dp).cuda()
    mvn = torch.empty((n, 32, 5), dtype=torch.float32, device="cuda"); ncl = torch.empty(n, dtype=torch.int32, device="cuda")
    gpu_ctx.cluster_async(d, n, 32, mvn, ncl); gpu_ctx.sync()
    assert np.array_equal(ncl.cpu().numpy(), ncl_o)
    assert same_bits_nan_aware(mvn.cpu().numpy(), mvn_o)


def test_postprocess_stages(scene_and_dp, orc, gpu_ctx):
    sc, offset, dpo = scene_and_dp
    n = sc.n
    pp = oracle.post_params(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    want = orc.postprocess_stages(dpo, sc.xyuvav, pp)
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    planes = torch.empty((5, sc.dimy, sc.dimx), dtype=torch.float32, device="cuda")
    stats = gpu_ctx.postprocess(torch.from_numpy(dpo).cuda(), sc.xyuvav, p, planes)
    assert (want["dpf0"] < 0).sum() > 20, "scene must have holes to fill"
    assert (want["ps_id"] != want["dpf1_id"]).sum() > 0, "scene must exercise pseudosmoothing"
    assert np.array_equal(gpu_ctx.postprocess_stage(0, n), want["dpf0"])
    assert stats[0] == want["dpf1_sweeps"] and stats[1] == want["ps_sweeps"], (stats, want["dpf1_sweeps"], want["ps_sweeps"])
    for which, key in ((1, "dpf1_id"), (4, "ps_id")):
        got = gpu_ctx.postprocess_stage(which, n)
        assert np.array_equal(got, want[key]), f"{key}: {(got != want[key]).sum()} cluster choices differ"
    for which, key in ((2, "dpf1_dx"), (3, "dpf1_dy"), (5, "ps_dx"), (6, "ps_dy")):
        got = gpu_ctx.postprocess_stage(which, n)
        assert np.array_equal(np.isnan(got), np.isnan(want[key]))
        assert np.allclose(got, want[key], rtol=1e-5, atol=1e-6, equal_nan=True), key
    planes_o = orc.postprocess(dpo, sc.xyuvav, pp)
    assert same_bits_nan_aware(planes.cpu().numpy(), planes_o), mismatch_report(planes.cpu().numpy(), planes_o)
    fin_o, a_o, b_o = orc.finalize(planes_o, pp)
    a, b = gpu_ctx.finalize(planes, p)
    assert (a, b) == (a_o, b_o)
    assert same_bits_nan_aware(planes.cpu().numpy()[:4], fin_o[:4])


def test_pipeline_end_to_end_accuracy(orc):
    """Host buffers in, five planes out; checks absolute accuracy against the known shift field."""
    sc = small_scene(H=700, W=700, seed=29, spacing=21, null_wedge=False)
    pl = pipeline.Pipeline(0)
    try:
        planes, stats, bias = pl.run(sc.i0.numpy().astype(np.uint8), sc.i1.numpy().astype(np.uint8), sc.xyuvav,
                                     sc.dimx, sc.dimy, sc.dt, np.array(sc.offset, np.int32), finalize=False)
    finally:
        pl.close()
    err_u = planes[0] - (sc.truth_du - sc.offset[0])
    err_v = planes[1] - (sc.truth_dv - sc.offset[1])
    err = np.hypot(err_u, err_v)
    assert np.nanmedian(err) < 0.05 and np.nanpercentile(err, 95) < 0.25, (np.nanmedian(err), np.nanpercentile(err, 95))
    assert np.isnan(planes[0]).mean() < 0.02


def test_postprocess_only_rerun_from_a_dp_dump(scene_and_dp, tmp_path):
    """dp_NN.gma dump -> reload -> postprocess (the reference's MIMC_main_test_postprocessing.c flow)
    gives the planes of the in-memory run, bit for bit."""
    from mimc3_b200 import gma
    sc, offset, dpo = scene_and_dp
    pl = pipeline.Pipeline(0)
    try:
        pl.set_images(sc.i0.numpy(), sc.i1.numpy())
        dp, _ = pl.match_all(sc.xyuvav, sc.dimx, sc.dimy, sc.dt, offset)
        planes, stats = pl.postprocess(dp)
        bias = pl.ctx.finalize(planes, pl.params)
        pl.ctx.sync()
        gma.save_dp(str(tmp_path), dp.cpu().numpy())
        planes2, stats2, bias2 = pl.postprocess_from_dump(str(tmp_path), sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
        assert same_bits_nan_aware(planes.cpu().numpy(), planes2) and bias == bias2
        with pytest.raises(ValueError):
            pl.postprocess_from_dump(str(tmp_path), sc.xyuvav[:-1], sc.dimx, sc.dimy, sc.dt)
    finally:
        pl.close()
