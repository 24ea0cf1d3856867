"""GPU parity of the control-point stage (get_offset_image, MIMC_module.c:33-492) against the
UNMODIFIED reference build with time() pinned (the reference seeds rand() with time(NULL))."""
import numpy as np
import pytest

from mimc3_b200 import lib
from tests.util import small_scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed_time,scene_kw", [
    (1700000000, dict(H=700, W=700, seed=61, spacing=21, null_wedge=True)),
    (1234567, dict(H=640, W=900, seed=62, spacing=17, null_wedge=False, offset=(-3, 4))),
    (42, dict(H=700, W=700, seed=63, spacing=23, null_wedge=True, dtype="u16", offset=(5, 2))),
])
def test_get_offset_image_matches_reference(gpu_ctx, ref, orc, seed_time, scene_kw):
    sc = small_scene(**scene_kw)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    ref.set_globals(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    rc_r, off_r, flag_r = ref.get_offset_image(i0, i1, sc.xyuvav, fake_time=seed_time)
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    a, b = gpu_ctx.image_from(i0), gpu_ctx.image_from(i1)
    try:
        rc, off, flag, ncp = gpu_ctx.get_offset_image(a, b, sc.xyuvav, p, seed_time)
    finally:
        gpu_ctx.image_destroy(a); gpu_ctx.image_destroy(b)
    assert rc == rc_r == 1
    assert np.array_equal(off, off_r), (off, off_r)
    assert np.array_equal(off, np.array(sc.offset, np.int32))          # and it is the true rigid offset
    assert np.array_equal(flag, flag_r), f"{(flag != flag_r).sum()} control-point flags differ"
    assert ncp >= flag.sum() > 0
    rc_o, off_o, flag_o = orc.get_offset_image(i0, i1, sc.xyuvav, seed_time)          # and the restated oracle
    assert rc_o == 1 and np.array_equal(off_o, off) and np.array_equal(flag_o, flag)


def test_get_offset_image_fails_without_candidates(gpu_ctx):
    """Every node fast (>= 10 m/yr): fewer than num_cp_min candidates => result -1 (MIMC_module.c:130-135)."""
    sc = small_scene(H=500, W=500, seed=64, spacing=21, null_wedge=False, background_mpy=400.0)
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    a, b = gpu_ctx.image_from(sc.i0.numpy()), gpu_ctx.image_from(sc.i1.numpy())
    try:
        rc, off, flag, ncp = gpu_ctx.get_offset_image(a, b, sc.xyuvav, p, 1)
    finally:
        gpu_ctx.image_destroy(a); gpu_ctx.image_destroy(b)
    assert rc == -1 and flag.sum() == 0 and ncp == 0
