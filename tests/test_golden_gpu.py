"""GPU parity against the committed golden vectors = outputs of the UNMODIFIED reference
(scripts/make_golden.py; /root/reference itself is not available on the GPU box)."""
import hashlib

import numpy as np
import pytest
import torch

from mimc3_b200 import lib, pipeline
from tests.test_oracle_cpu import FIXTURES, load
from tests.util import VEC_OCW, mismatch_report, same_bits_nan_aware

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("matcher", ("v1", "v2"))
@pytest.mark.parametrize("name", FIXTURES)
def test_multimatch_and_postprocess_match_the_reference(name, matcher):
    g = load(name)
    pl = pipeline.Pipeline(0)
    try:
        pl.ctx.set_matcher(matcher)
        pl.set_images(g["i0"], g["i1"])                         # u8 / u16 ingest path
        pl.set_grid(g["xyuvav"], g["dimx"], g["dimy"], g["dt"])
        # pivots: library host code == reference get_uv_pivot
        H, W = g["i0"].shape
        for ocw in VEC_OCW:
            off, piv = lib.get_uv_pivot(g["xyuvav"], g["dt"], pl.params.mpp, ocw, H, W)
            assert np.array_equal(off, g[f"piv_off_{ocw}"]) and np.array_equal(piv, g[f"piv_{ocw}"])
        dp, _ = pl.multimatch(g["offset"])
        pl.ctx.sync()
        dp_h = dp.cpu().numpy()
        for a in range(32):
            assert same_bits_nan_aware(dp_h[a], g["dp"][a]), f"attempt {a}: " + mismatch_report(dp_h[a], g["dp"][a])
        # the last conv2 pass (Laplacian) is still in i0c
        lap = pl.ctx.image_download(pl.handles["i0c"], H, W)
        assert hashlib.sha256(lap.tobytes()).hexdigest() == str(g["conv2_i0_sha256"][2])
        planes, stats = pl.postprocess(dp)
        n = g["dimx"] * g["dimy"]
        for which, key in ((0, "dpf0"), (1, "dpf1_id"), (4, "ps_id")):
            assert np.array_equal(pl.ctx.postprocess_stage(which, n), g["stage_" + key]), key
        for which, key in ((2, "dpf1_dx"), (3, "dpf1_dy"), (5, "ps_dx"), (6, "ps_dy")):
            got = pl.ctx.postprocess_stage(which, n)
            assert np.allclose(got, g["stage_" + key], rtol=1e-5, atol=1e-6, equal_nan=True), key
        assert same_bits_nan_aware(planes.cpu().numpy(), g["planes"]), mismatch_report(planes.cpu().numpy(), g["planes"])
    finally:
        pl.close()
