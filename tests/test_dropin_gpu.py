"""End-to-end drop-in test: the reference's UNCHANGED driver (MIMC_main.c, GMA.c, MIMC_misc.c,
georefimg.c) linked against libmimc3cu_dropin.a + libmimc3cu.so instead of MIMC_module.c
(oracle/_ref/MIMC3_dropin) must write the same eight GMA files and meta values as the stock
binary (oracle/_ref/MIMC3_ref) for the same four-argument command line.  Both binaries are built
in the container that has the reference sources (oracle/Makefile) and travel to the GPU box."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from mimc3_b200 import synth
from tests.util import small_scene

pytestmark = pytest.mark.gpu

REF_DIR = os.path.dirname(oracle.REF_CLI)
DROPIN_CLI = os.path.join(REF_DIR, "MIMC3_dropin")
FAKETIME = os.path.join(REF_DIR, "libfaketime.so")


def run_cli(binary, workdir, outdir, fake_time, **extra_env):
    os.makedirs(outdir, exist_ok=True)
    env = dict(os.environ, MIMC3_FAKE_TIME=str(fake_time), LD_PRELOAD=FAKETIME, **extra_env)
    args = [binary, os.path.join(workdir, "20200101000000_i0.tif"), os.path.join(workdir, "20200117000000_i1.tif"),
            os.path.join(workdir, "xyuvav.GMA"), outdir]
    r = subprocess.run(args, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, f"{binary} exited {r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"
    return r.stdout


def read_outputs(outdir):
    pre = os.path.join(outdir, "vmap_20200101000000_20200117000000_")
    out = {k: synth.read_gma(pre + k + ".GMA", "float32") for k in ("vx", "vy", "ex", "ey", "qual")}
    out["x"] = synth.read_gma(pre + "x.GMA", "float64"); out["y"] = synth.read_gma(pre + "y.GMA", "float64")
    out["flagcp"] = synth.read_gma(pre + "flagcp.GMA", "uint8")
    out["meta"] = dict(ln.strip().split("=", 1) for ln in open(pre + "meta.txt") if "=" in ln)
    return out


@pytest.mark.parametrize("dtype", ["u8", "u16"])
def test_unchanged_driver_with_the_cuda_module_writes_the_same_files(tmp_path, dtype):
    for f in (oracle.REF_CLI, DROPIN_CLI, FAKETIME):
        if not os.path.exists(f):
            pytest.skip(f"{f} not built (needs the container with /root/reference)")
    sc = small_scene(H=640, W=640, seed=71, spacing=24, dtype=dtype, null_wedge=True, decorrelated_patches=6, offset=(3, -2))
    work = str(tmp_path)
    np_dt = np.uint8 if dtype == "u8" else np.uint16
    synth.write_tiff(os.path.join(work, "20200101000000_i0.tif"), sc.i0.numpy().astype(np_dt))
    synth.write_tiff(os.path.join(work, "20200117000000_i1.tif"), sc.i1.numpy().astype(np_dt))
    synth.write_gma(os.path.join(work, "xyuvav.GMA"), sc.xyuvav)
    run_cli(oracle.REF_CLI, work, os.path.join(work, "out_ref"), 1700000123)
    log = run_cli(DROPIN_CLI, work, os.path.join(work, "out_gpu"), 1700000123)
    assert "Elapsed time" in log                      # the driver's own timing prints around every matcher call
    a, b = read_outputs(os.path.join(work, "out_ref")), read_outputs(os.path.join(work, "out_gpu"))
    assert a["meta"]["cp_offset_int_u"] == b["meta"]["cp_offset_int_u"] == str(sc.offset[0])
    assert a["meta"]["cp_offset_int_v"] == b["meta"]["cp_offset_int_v"] == str(sc.offset[1])
    assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["y"], b["y"])
    assert np.array_equal(a["flagcp"], b["flagcp"])
    # the reader tooling (mimc3_b200/vmap.py) on files written by the reference's own savers
    from mimc3_b200 import vmap
    vm = vmap.VMap(os.path.join(work, "out_ref"))
    assert np.array_equal(vm.vx, a["vx"], equal_nan=True) and np.array_equal(vm.flagcp, a["flagcp"])
    assert vm.meta["cp_offset_int_u"] == sc.offset[0] and vm.meta["name_i0"].endswith("_i0.tif")
    assert a["vx"].shape == (sc.dimy, sc.dimx)
    report = []
    for k in ("vx", "vy", "ex", "ey", "qual"):
        assert np.array_equal(np.isnan(a[k]), np.isnan(b[k])), k
        # bit-identical except where glibc / CUDA exp differ by an ulp inside the hole-filling and pseudosmoothing weights
        # (interpolated nodes only); the count is written to gpurun_out/ so that it can be quoted
        fin = ~np.isnan(a[k])
        differ = int((a[k] != b[k])[fin].sum())
        report.append(f"{dtype} {k}: {differ} of {int(fin.sum())} finite values not bit-identical, max |diff| {float(np.nanmax(np.abs(a[k] - b[k]))):.3g}")
        assert np.allclose(a[k], b[k], rtol=1e-5, atol=1e-4, equal_nan=True), (k, np.nanmax(np.abs(a[k] - b[k])))
        assert differ <= 0.005 * fin.sum(), report[-1]
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"dropin_bit_identity_{dtype}.txt"), "w") as f:
            f.write("\n".join(report) + "\n")
    assert abs(float(a["meta"]["cp_offset_subint_u"]) - float(b["meta"]["cp_offset_subint_u"])) < 1e-4


def test_unchanged_driver_on_two_gpus(tmp_path):
    """MIMC3CU_DEVICES=2: the same unchanged driver, node rows sharded over two GPUs inside the drop-in (one context and
    one host thread per GPU, banded postprocess over the library's NCCL communicator) writes the files of the one-GPU run."""
    from mimc3_b200 import lib
    for f in (DROPIN_CLI, FAKETIME):
        if not os.path.exists(f):
            pytest.skip(f"{f} not built (needs the container with /root/reference)")
    if lib.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sc = small_scene(H=800, W=640, seed=72, spacing=22, dtype="u16", null_wedge=True, decorrelated_patches=8, offset=(1, 2))
    work = str(tmp_path)
    synth.write_tiff(os.path.join(work, "20200101000000_i0.tif"), sc.i0.numpy().astype(np.uint16))
    synth.write_tiff(os.path.join(work, "20200117000000_i1.tif"), sc.i1.numpy().astype(np.uint16))
    synth.write_gma(os.path.join(work, "xyuvav.GMA"), sc.xyuvav)
    run_cli(DROPIN_CLI, work, os.path.join(work, "out_1"), 1700000555)
    run_cli(DROPIN_CLI, work, os.path.join(work, "out_2"), 1700000555, MIMC3CU_DEVICES="2")
    a, b = read_outputs(os.path.join(work, "out_1")), read_outputs(os.path.join(work, "out_2"))
    for k in ("vx", "vy", "ex", "ey", "qual", "flagcp", "x", "y"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert a["meta"] == b["meta"]
