"""Wall clock of the reference's UNCHANGED four-argument CLI linked against the CUDA drop-in (oracle/_ref/MIMC3_dropin)
on a benchmark-sized scene, with the host share broken down per module entry point (MIMC3CU_DROPIN_TIMING=1) --
SURVEY.md 8f-2 / 8f-3.  The stock OpenMP binary is only run when --with-reference is given (at C2 it needs ~10 min).

    python scripts/cli_time.py c2 [--devices N] [--with-reference] [--workdir DIR]
"""
import argparse, collections, os, re, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from mimc3_b200 import synth
import bench

ap = argparse.ArgumentParser()
ap.add_argument("workload", nargs="?", default="c2")
ap.add_argument("--devices", type=int, default=1)
ap.add_argument("--with-reference", action="store_true")
ap.add_argument("--workdir", default=None)
args = ap.parse_args()

wl = dict(bench.WORKLOADS[args.workload]); wl.pop("desc"); wl.pop("strong", None)
work = args.workdir or tempfile.mkdtemp(prefix="mimc3_cli_")
os.makedirs(work, exist_ok=True)
import torch
t0 = time.perf_counter()
sc = synth.make_scene(seed=1234, device="cuda" if torch.cuda.is_available() else "cpu", **wl)
np_dt = np.uint8 if sc.dtype == "u8" else np.uint16
synth.write_tiff(os.path.join(work, "20200101000000_i0.tif"), sc.i0.cpu().numpy().astype(np_dt))
synth.write_tiff(os.path.join(work, "20200117000000_i1.tif"), sc.i1.cpu().numpy().astype(np_dt))
synth.write_gma(os.path.join(work, "xyuvav.GMA"), sc.xyuvav)
H, W = sc.shape
print(f"scene {args.workload}: {H}x{W} {sc.dtype}, {sc.n} nodes, written in {time.perf_counter() - t0:.1f} s")
del sc
torch.cuda.empty_cache()
ref_dir = os.path.dirname(oracle.REF_CLI)
runs = [("CUDA drop-in", os.path.join(ref_dir, "MIMC3_dropin"), dict(MIMC3CU_DROPIN_TIMING="1", MIMC3CU_DEVICES=str(args.devices)))]
if args.with_reference:
    runs.append(("reference (OpenMP)", oracle.REF_CLI, {}))
for name, binary, extra in runs:
    out = os.path.join(work, "out_" + name.split()[0])
    os.makedirs(out, exist_ok=True)
    env = dict(os.environ, MIMC3_FAKE_TIME="1700000123", LD_PRELOAD=os.path.join(ref_dir, "libfaketime.so"), **extra)
    t0 = time.perf_counter()
    r = subprocess.run([binary, os.path.join(work, "20200101000000_i0.tif"), os.path.join(work, "20200117000000_i1.tif"),
                        os.path.join(work, "xyuvav.GMA"), out], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.perf_counter() - t0
    match = sum(float(l.split("Elapsed time:")[1].split()[0]) for l in r.stdout.splitlines() if "Elapsed time:" in l)
    print(f"{name:20s} rc={r.returncode} wall {wall:8.2f} s; the driver's own 'Elapsed time' prints around matching_ncc_dlc_2 sum to {match:8.2f} s "
          f"({os.cpu_count()} host cores, {args.devices} GPU(s))")
    if r.returncode:
        print(r.stdout[-1500:], r.stderr[-1500:])
    tot = collections.OrderedDict()
    for m in re.finditer(r"\[mimc3cu drop-in\] (\S+)\s+([0-9.]+) ms", r.stderr):
        tot.setdefault(m.group(1), [0, 0.0])
        tot[m.group(1)][0] += 1; tot[m.group(1)][1] += float(m.group(2))
    if tot:
        inside = sum(v[1] for v in tot.values()) / 1e3
        for k, (cnt, ms) in tot.items():
            print(f"    {k:22s} {cnt:3d} calls {ms / 1e3:8.2f} s")
        print(f"    inside the module: {inside:.2f} s; the driver's own share (TIFF decode, pivot negation and per-node frees, GMA writes, "
              f"process start and CUDA context): {wall - inside:.2f} s")
