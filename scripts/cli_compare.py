"""Wall-clock of the reference's unchanged CLI: stock binary (OpenMP, all host cores) vs the same driver
linked against the CUDA drop-in, on the C1 configuration (2048x2048 u8, ~100x100 nodes).
    python scripts/cli_compare.py [outdir]"""
import os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from mimc3_b200 import synth

work = sys.argv[1] if len(sys.argv) > 1 else tempfile.mkdtemp(prefix="mimc3_cli_")
os.makedirs(work, exist_ok=True)
sc = synth.make_scene(H=2048, W=2048, dtype="u8", spacing=19, seed=1234, peak_px=6.3, null_wedge=True, decorrelated_patches=20)
synth.write_tiff(os.path.join(work, "20200101000000_i0.tif"), sc.i0.numpy().astype(np.uint8))
synth.write_tiff(os.path.join(work, "20200117000000_i1.tif"), sc.i1.numpy().astype(np.uint8))
synth.write_gma(os.path.join(work, "xyuvav.GMA"), sc.xyuvav)
ref_dir = os.path.dirname(oracle.REF_CLI)
env = dict(os.environ, MIMC3_FAKE_TIME="1700000123", LD_PRELOAD=os.path.join(ref_dir, "libfaketime.so"))
res = {}
for name, binary in (("reference (OpenMP)", oracle.REF_CLI), ("CUDA drop-in", os.path.join(ref_dir, "MIMC3_dropin"))):
    out = os.path.join(work, "out_" + name.split()[0])
    os.makedirs(out, exist_ok=True)
    for f in os.listdir(out):
        os.remove(os.path.join(out, f))
    t0 = time.perf_counter()
    r = subprocess.run([binary, os.path.join(work, "20200101000000_i0.tif"), os.path.join(work, "20200117000000_i1.tif"),
                        os.path.join(work, "xyuvav.GMA"), out], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    dt = time.perf_counter() - t0
    match = sum(float(l.split("Elapsed time:")[1].split()[0]) for l in r.stdout.splitlines() if "Elapsed time:" in l)
    res[name] = (dt, match, out)
    print(f"{name:20s} rc={r.returncode} wall {dt:7.2f} s   sum of the driver's own 'Elapsed time' prints around matching_ncc_dlc_2: {match:7.2f} s   ({sc.n} nodes, {os.cpu_count()} host cores)")
a = synth.read_gma(os.path.join(res["reference (OpenMP)"][2], "vmap_20200101000000_20200117000000_vx.GMA"))
b = synth.read_gma(os.path.join(res["CUDA drop-in"][2], "vmap_20200101000000_20200117000000_vx.GMA"))
print("vx identical:", np.array_equal(np.isnan(a), np.isnan(b)) and float((a == b)[~np.isnan(a)].mean()), " max |diff|:", float(np.nanmax(np.abs(a - b))))
print(f"speed-up: wall {res['reference (OpenMP)'][0] / res['CUDA drop-in'][0]:.1f}x, matching {res['reference (OpenMP)'][1] / max(res['CUDA drop-in'][1], 1e-9):.1f}x")
