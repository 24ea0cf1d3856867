"""CPU tests (no GPU) of the host side of the product: the C-ABI library loads and exports every
symbol include/mimc3cu.h declares, its host-only entry points (pivot generation, defaults) agree
with the oracle, and every compute entry point fails loudly without a CUDA device."""
import os
import re

import numpy as np
import pytest
import torch

import oracle
from mimc3_b200 import lib
from tests.util import VEC_OCW, small_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mimc3cu_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = lib.load_library()
    syms = declared_symbols("mimc3cu.h")
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert set(lib.EXPORTED_SYMBOLS) <= set(syms), sorted(set(lib.EXPORTED_SYMBOLS) - set(syms))
    assert L.mimc3cu_version() == 100


def test_default_params_are_the_reference_constants():
    p = lib.default_params()
    assert list(p.vec_ocw) == [7, 15, 30, 40]          # MIMC_main.c:134-137
    assert p.AW_CRE == 10.0 and abs(p.AW_SF - 1.8) < 1e-6
    assert p.radius_neighbor_dpf1 == 3.0 and p.radius_neighbor_ps == 5.0 and p.num_dp == 32


@pytest.mark.parametrize("peak_px", (6.3, 43.0))
def test_get_uv_pivot_host_matches_oracle(orc, peak_px):
    """mimc3cu_get_uv_pivot is host code (like the reference's) and needs no GPU."""
    sc = small_scene(H=900, W=700, seed=3, peak_px=peak_px, spacing=17, apriori_gain=0.9)
    H, W = sc.shape
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    for ocw in VEC_OCW:
        off, piv = lib.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, H, W)
        off_o, piv_o = orc.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, H, W)
        assert np.array_equal(off, off_o) and np.array_equal(piv, piv_o)
        assert (np.diff(off) >= 1).all()


def test_get_uv_pivot_near_the_border_truncates(orc):
    """The border test of MIMC_module.c:576-580 shortens pivot lines of nodes next to the edge."""
    sc = small_scene(H=300, W=300, seed=4, peak_px=30.0, spacing=11, margin=44, apriori_gain=1.0)
    H, W = sc.shape
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    off, piv = lib.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, 40, H, W)
    off_o, piv_o = orc.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, 40, H, W)
    assert np.array_equal(off, off_o) and np.array_equal(piv, piv_o)
    assert np.diff(off).min() < np.diff(off).max()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_gpu():
    L = lib.load_library()
    assert L.mimc3cu_device_count() == 0
    with pytest.raises(lib.Mimc3CuError) as e:
        lib.Context(0)
    assert "no CPU fallback" in str(e.value) or "no CUDA device" in str(e.value)


def test_product_package_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under mimc3_b200/ may reference it."""
    pkg = os.path.join(ROOT, "mimc3_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), os.path.join(dirpath, f)
                assert "mimc3_oracle" not in txt, os.path.join(dirpath, f)


def test_get_uv_pivot_multithreaded_grid(orc):
    """Grids of >= 4096 nodes take the library's multi-threaded path (and its sizing/fill cache)."""
    from mimc3_b200 import synth
    sc = synth.make_scene(H=1536, W=1536, dtype="u8", spacing=19, seed=3, peak_px=9.0)
    assert sc.n >= 4096
    p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    for ocw in (7, 40):
        off, piv = lib.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, 1536, 1536)
        off_o, piv_o = orc.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, 1536, 1536)
        assert np.array_equal(off, off_o) and np.array_equal(piv, piv_o)
    # a modified copy of the same array object must not hit the cache of the previous call
    x2 = sc.xyuvav.copy(); x2[:, 4] *= 3.0
    off2, piv2 = lib.get_uv_pivot(x2, sc.dt, p.mpp, 40, 1536, 1536)
    off2_o, piv2_o = orc.get_uv_pivot(x2, sc.dt, p.mpp, 40, 1536, 1536)
    assert np.array_equal(off2, off2_o) and np.array_equal(piv2, piv2_o)


def test_reference_arm_of_the_bench_runs_without_the_cuda_library(tmp_path):
    """`bench.py --impl reference` is the reference's own CPU implementation only: it must work where libmimc3cu.so cannot
    even be found, take all host threads although the launcher exports OMP_NUM_THREADS=1 (torchrun does), and print the
    contract's JSON line."""
    import json
    import subprocess
    import sys
    import oracle
    if not os.path.exists(oracle.REF_SO):
        pytest.skip("oracle/_ref/libmimc3ref.so not present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MIMC3CU_LIB=str(tmp_path / "no_such_library.so"), OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                        "--warmup", "0", "--cpu-sample-nodes", "102"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "nodes/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cp_stage_ms"] > 0 and line["cp_offset"] == [2, -1]


def test_bench_reference_sample_returns_the_references_dp(orc):
    """The parity verdict of the bench line compares the GPU dp with what cpu_reference_sample keeps: it must be the
    reference's matching_ncc_dlc_2 output of the sampled nodes, with the sign flip main applies to the swapped passes."""
    import oracle
    import bench
    from mimc3_b200 import synth
    if not os.path.exists(oracle.REF_SO):
        pytest.skip("oracle/_ref/libmimc3ref.so not present")
    sc = synth.make_scene(H=384, W=384, dtype="u8", spacing=31, seed=8, peak_px=4.0)
    i0, i1 = sc.i0.numpy(), sc.i1.numpy()
    offset = np.array(sc.offset, np.int32)
    res = bench.cpu_reference_sample(i0, i1, sc.xyuvav, sc.dimx, sc.dimy, sc.dt, offset, 2 * sc.dimx, keep_dp=True, with_cp=False)
    idx = res["idx"]
    assert res["dp"].shape == (32, len(idx), 3) and res["kind"] == "reference"
    mpp = float(np.float32((sc.xyuvav[1, 0] - sc.xyuvav[0, 0]) / (sc.xyuvav[1, 2] - sc.xyuvav[0, 2])))
    xs = np.ascontiguousarray(sc.xyuvav[idx])
    off, piv = orc.get_uv_pivot(xs, sc.dt, mpp, 15, 384, 384)
    fwd, _, _ = orc.match(i0, i1, xs, offset, off, piv, +1, 15)
    swp, _, _ = orc.match(i1, i0, xs, -offset, off, piv, -1, 15)
    a, b = res["dp"][2], fwd          # attempt 2 = raw pair, ocw 15, forward
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    a, b = res["dp"][3], swp * np.array([-1, -1, 1], np.float32)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
