"""Reader tooling: GMA files, the vmap reader, the dp dump/reload layout."""
import io
import os
import struct
import tarfile

import numpy as np
import pytest

from mimc3_b200 import gma, vmap


def test_gma_roundtrip_all_element_types(tmp_path):
    rng = np.random.default_rng(1)
    for dt in ("float32", "float64", "int32", "uint8", "uint16", "uint32"):
        a = (rng.random((7, 5)) * 200).astype(dt)
        p = tmp_path / f"m_{dt}.GMA"
        gma.write(p, a)
        raw = p.read_bytes()
        assert struct.unpack("<ii", raw[:8]) == (7, 5) and len(raw) == 8 + a.nbytes     # GMA.c:324-325 header
        for src in (p, str(p), raw, io.BytesIO(raw)):
            b = gma.read(src, dt)
            assert b.dtype == np.dtype(dt) and np.array_equal(a, b) and b.flags.writeable
    gma.write(tmp_path / "row.GMA", np.arange(4, dtype=np.float32))          # 1-D -> one row
    assert gma.read(tmp_path / "row.GMA").shape == (1, 4)
    assert gma.read(gma.dumps(np.zeros((0, 3), np.float32))).shape == (0, 3)  # empty matrix


def test_gma_rejects_a_payload_that_does_not_match_the_header(tmp_path):
    p = tmp_path / "bad.GMA"
    p.write_bytes(struct.pack("<ii", 3, 3) + b"\0" * 20)
    with pytest.raises(ValueError):
        gma.read(p, "float32")
    with pytest.raises(ValueError):
        gma.read(b"\0\0\0")
    with pytest.raises(ValueError):
        gma.write(tmp_path / "x.GMA", np.zeros((2, 2, 2)))


def test_gma_reads_a_file_written_by_the_reference_savers(tmp_path):
    """Byte layout of GMA_float_save (GMA.c:407-424): dims as two 32-bit words, then one element at a time."""
    vals = np.array([[1.5, -2.0, 3.25], [4.0, np.nan, 6.0]], np.float32)
    with open(tmp_path / "ref.GMA", "wb") as f:
        f.write(struct.pack("<I", 2)); f.write(struct.pack("<I", 3))
        for r in range(2):
            for c in range(3):
                f.write(struct.pack("<f", vals[r, c]))
    got = gma.read(tmp_path / "ref.GMA")
    assert np.array_equal(got, vals, equal_nan=True)


def _write_vmap(d, stem="vmap_20200101000000_20200117000000"):
    rng = np.random.default_rng(2)
    f = {k: rng.normal(size=(4, 6)).astype(gma.FIELD_DTYPE[k]) for k in ("x", "y", "vx", "vy", "ex", "ey", "qual")}
    f["vx"][0, 0] = np.nan
    f["flagcp"] = (rng.random((4, 6)) < 0.3).astype(np.uint8)
    for k, v in f.items():
        gma.write(os.path.join(d, f"{stem}_{k}.GMA"), v)
    with open(os.path.join(d, f"{stem}_meta.txt"), "w") as fo:      # MIMC_main.c:439-447
        fo.write("MIMC_version=3.0\nname_i0=/a/b=c.tif\nname_i1=i1.tif\ncp_offset_int_u=2\ncp_offset_int_v=-1\n"
                 "cp_offset_subint_u=0.125000\ncp_offset_subint_v=-0.250000\n")
    return f, stem


def test_vmap_reads_directory_and_tar(tmp_path):
    d = tmp_path / "out"; d.mkdir()
    f, stem = _write_vmap(str(d))
    tar = tmp_path / f"{stem}.tar"
    with tarfile.open(tar, "w:gz") as tf:        # the driver packs with `tar -cvzf`, :462
        for name in os.listdir(d):
            tf.add(d / name, arcname=name)
    for src in (str(d), str(tar)):
        vm = vmap.VMap(src)
        for k in vmap.FIELDS:
            assert np.array_equal(getattr(vm, k), f[k], equal_nan=True), k
        assert vm.meta["cp_offset_int_u"] == 2 and vm.meta["cp_offset_int_v"] == -1
        assert vm.meta["cp_offset_subint_u"] == 0.125 and vm.meta["name_i0"] == "/a/b=c.tif"
        assert np.allclose(vm.spd, np.sqrt(f["vx"] ** 2 + f["vy"] ** 2), equal_nan=True)
        cp = f["flagcp"] != 0
        bx, by = vm.adjust()                     # vmap.py:174-185
        assert np.isclose(bx, np.nanmean(f["vx"][cp])) and np.isclose(by, np.nanmean(f["vy"][cp]))
        assert np.allclose(vm.vx, f["vx"] - np.float32(bx), equal_nan=True)
    with pytest.raises(FileNotFoundError):
        vmap.VMap(str(tmp_path)).vx


def test_dp_dump_layout_and_reload(tmp_path):
    rng = np.random.default_rng(3)
    dp = rng.normal(size=(32, 11, 3)).astype(np.float32)
    dp[3, 4] = np.nan
    flag = (rng.random(11) < 0.5).astype(np.uint8)
    gma.save_dp(str(tmp_path), dp, flag)
    assert sorted(os.listdir(tmp_path / "FT_result"))[:2] == ["dp_00.gma", "dp_01.gma"]   # MIMC_main_test_postprocessing.c:279
    back, fl = gma.load_dp(str(tmp_path))
    assert np.array_equal(back, dp, equal_nan=True) and np.array_equal(fl, flag)
    gma.write(tmp_path / "FT_result" / "dp_05.gma", np.zeros((10, 3), np.float32))
    with pytest.raises(ValueError):
        gma.load_dp(str(tmp_path))


def test_gma_roundtrip_property():
    """Any 2-D shape (including empty ones) and any byte pattern survive write -> read unchanged."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=60, deadline=None)
    @hyp.given(st.integers(0, 9), st.integers(0, 9), st.sampled_from(["float32", "float64", "int32", "uint8"]), st.integers(0, 2**31 - 1))
    def check(nr, nc, dt, seed):
        raw = np.random.default_rng(seed).integers(0, 256, size=nr * nc * np.dtype(dt).itemsize, dtype=np.uint8)
        a = raw.view(dt).reshape(nr, nc)
        b = gma.read(gma.dumps(a), dt)
        assert b.shape == (nr, nc) and a.tobytes() == b.tobytes()      # NaN payloads included

    check()
