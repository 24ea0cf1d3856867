"""GPU parity of GMA_float_conv2 (MIMC_module.c:2517-2585) incl. stale-border semantics."""
import numpy as np
import pytest

import oracle
from tests.util import same_bits_nan_aware, mismatch_report, small_scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["u8", "u16"])
def test_conv2_chain_bitwise(gpu_ctx, orc, dtype):
    sc = small_scene(H=300, W=517, dtype=dtype, seed=3)
    img = sc.i0.numpy()
    H, W = img.shape
    want = np.zeros_like(img)
    src = gpu_ctx.image_from(img)
    dst = gpu_ctx.image_create(H, W)
    try:
        for kid in (0, 1, 2, 0):      # same output buffer reused, like main's i0c
            orc.conv2(img, kid, want)
            gpu_ctx.conv2(src, oracle.KERNELS[kid], dst)
            got = gpu_ctx.image_download(dst, H, W)
            assert same_bits_nan_aware(got, want), mismatch_report(got, want, f"kernel {kid}")
    finally:
        gpu_ctx.image_destroy(src); gpu_ctx.image_destroy(dst)


def test_upload_u8_u16_cast(gpu_ctx):
    rng = np.random.default_rng(0)
    for dt in (np.uint8, np.uint16):
        a = rng.integers(0, np.iinfo(dt).max, size=(77, 131), dtype=dt)
        h = gpu_ctx.image_create(*a.shape)
        gpu_ctx.image_upload(h, a)
        assert np.array_equal(gpu_ctx.image_download(h, *a.shape), a.astype(np.float32))
        gpu_ctx.image_destroy(h)
