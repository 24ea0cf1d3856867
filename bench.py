#!/usr/bin/env python
"""Benchmark of the MIMC3 per-grid-node matching hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c1|c3|c4|c5|c5s|c5s_small]

Metric (BASELINE.json): grid nodes matched per second.  One node = all 32 matching
attempts (4 chip sizes x {forward, swapped} x {raw, d/dx, d/dy, Laplacian}) + its share of
the 6 conv2 passes + the postprocess (cluster, dpf0/dpf1, pseudosmoothing).  A "step" is
one pass of that path over every node of the workload.

  value  device-timed (CUDA events on the library's stream, max over ranks), inputs
         (both images, nodes, pivots) already resident in HBM.
  e2e    the same metric through the C-ABI host entry points with HOST buffers, `--steps` timed steps: pinned
         u16/u8 images and xyuvav copied H2D, the control-point stage (N = 1), host pivot generation,
         multi-match, postprocess, finalize, five planes copied D2H -- all inside the timed region.
  roofline      dominant kernel = the matcher (match2_kernel<ocw,G>, plus match_kernel for nodes
                outside its class); achieved = sum over attempts and nodes of 8*S^2*E flop
                (E = NCC cells the reference algorithm evaluates, counted by the kernel and
                cross-checked against the oracle in tests) / summed matcher durations (CUDA
                events recorded by the library around every attempt inside the timed
                region); peak = FP32 FMA throughput measured live by an FMA micro-benchmark
                (MEASURED_PEAKS.json has no FP32 CUDA-core figure); peak_nominal / frac_nominal =
                the same against 148 SMs x 128 lanes x 2 x the maximum SM clock.
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/libmimc3ref.so, OpenMP, all host
                threads) on a bounded sample of the same workload: 32 attempts on four node rows of the
                full images, the six conv2 calls on the full images, the control-point stage once.
  parity        (N = 1) the GPU results of the measured scene against that reference sample (dp of all 32
                attempts bit for bit), against the oracle (integer peaks, evaluated cells) and the postprocess of
                a band of node rows against the oracle's (cluster choice): mismatch counts.
  --impl reference   the same CPU sample as its own arm (rank 0 alone under torchrun); loads nothing of
                libmimc3cu.so.

N > 1 (torchrun), default workloads: weak scaling.  Every rank owns one tile of a vertical mosaic (its own
image pair + node-row band) and matches it with no data-path collective.  The postprocess
runs banded over the whole mosaic grid: each rank sweeps its own band; halo rows, dirty flags and sweep counters
travel over NCCL, issued by the library itself on its stream (csrc/comm.cu; `--band-comm python` routes them
through the callbacks of mimc3_b200/bands.py instead), then the five planes are gathered on rank 0.
`--workload c5s`: strong scaling -- ONE 32768^2 scene, node-row bands (balanced by pivot counts) over the ranks.
Timing = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "c2": dict(H=16384, W=16384, dtype="u16", spacing=20, mpp=15.0, peak_px=6.3,
               desc="Landsat-8-like 15 m pan pair, 16384x16384 u16 synthetic, 300 m (20 px) node spacing, default chip sizes"),
    # configs[0]
    "c1": dict(H=2048, W=2048, dtype="u8", spacing=19, mpp=15.0, peak_px=6.3,
               desc="2048x2048 u8 synthetic pair, 100x100-node grid"),
    # configs[2]: Sentinel-2-like 10 m tile, dense 100 m (10 px) node spacing, all four chip sizes
    "c3": dict(H=10980, W=10980, dtype="u16", spacing=10, mpp=10.0, peak_px=6.3,
               desc="Sentinel-2-like 10 m tile, 10980x10980 u16 synthetic, 100 m (10 px) node spacing, multichip"),
    # configs[4]: the 32768^2 mosaic at 8-px node spacing (16.7 M nodes on ONE GPU here; with --gpus N every
    # rank takes one such tile only if memory allows -- meant as a single-GPU maximum-size run)
    "c5": dict(H=32768, W=32768, dtype="u8", spacing=8, mpp=15.0, peak_px=6.3,
               desc="32768x32768 u8 synthetic mosaic, 8-px node spacing"),
    # configs[3] (fast outlet glacier), reduced image so it stays a quick extra
    "c4": dict(H=8192, W=8192, dtype="u16", spacing=20, mpp=15.0, peak_px=43.0, apriori_gain=0.9, decorrelated_patches=200,
               band_width_frac=0.2, desc="fast-glacier case: ~43 px a-priori displacement, wide DLC windows, 8192x8192 u16"),
}
# configs[4] as ONE scene sharded over the GPUs (strong scaling): every rank holds both images and a contiguous band of
# node rows; decorrelated patches make the iterative stages of the postprocess sweep (halo exchanges every sweep)
WORKLOADS["c5s"] = dict(H=32768, W=32768, dtype="u8", spacing=8, mpp=15.0, peak_px=6.3, decorrelated_patches=3000, strong=True,
                        desc="ONE 32768x32768 u8 synthetic mosaic, 8-px node spacing, node-row bands sharded over the GPUs, decorrelated patches")
# the same layout at a size the CPU-side checks and 2-GPU development runs finish quickly
WORKLOADS["c5s_small"] = dict(H=8192, W=8192, dtype="u8", spacing=8, mpp=15.0, peak_px=6.3, decorrelated_patches=200, strong=True,
                              desc="ONE 8192x8192 u8 synthetic pair, 8-px node spacing, node-row bands sharded over the GPUs, decorrelated patches")
VEC_OCW = (7, 15, 30, 40)


_JSON_FD = 1


def emit_json(line):
    sys.stdout.flush()
    os.write(_JSON_FD, (json.dumps(line) + "\n").encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# -------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# -------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # under load = samples in the upper half of the observed range
        hi = [x for x in sm if x >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(hi)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------
# the reference / CPU arm
# -------------------------------------------------------------------------------------------
def _host_threads():
    """Host threads the CPU arm may use: all cores this process is allowed on (torchrun exports
    OMP_NUM_THREADS=1 to its workers; the reference arm runs on rank 0 alone, so it takes them all)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_sample(i0, i1, xyuvav, dimx, dimy, dt, offset, target_nodes, keep_dp=False, with_cp=True):
    """Times the unmodified reference (oracle/_ref/libmimc3ref.so: MIMC_module.c + GMA.c compiled as they are) on a
    bounded sample of the workload, with all host threads:
      * all 32 matching attempts (MIMC_main.c:261-350) on every k-th node row of the full-size images, scaled
        linearly in nodes;
      * the six GMA_float_conv2 calls on the FULL images (the reference's own conv2, timed, not estimated);
      * get_offset_image (the control-point stage) once, on the full node list.
    Nothing of mimc3_b200's library is loaded or called here.  Returns a dict; with keep_dp also the
    reference's dp (32, sample, 3) and the sampled node indices, for the parity verdict of the bench line."""
    import oracle
    H, W = i0.shape
    kind = "reference"
    nthreads = _host_threads()
    try:
        R = oracle.Reference()
        ncores = R.set_num_threads(nthreads)
        R.set_globals(xyuvav, dimx, dimy, dt)
        match = lambda a, b, x, offs, off, piv, sign, ocw: R.match(a, b, x, offs, off, piv, sign, ocw)
        pivots = lambda x, ocw: R.get_uv_pivot(x, dt, ocw, H, W)
        conv2 = R.conv2

        def cp_stage():
            t = time.perf_counter(); rc, off, _ = R.get_offset_image(i0, i1, xyuvav, fake_time=1); return time.perf_counter() - t, rc, off
    except (FileNotFoundError, OSError) as e:
        log(f"[bench] reference build not available ({e}); timing the oracle port instead")
        kind = "port"
        O = oracle.Oracle()
        ncores = O.num_threads()
        mpp = float(np.float32((xyuvav[1, 0] - xyuvav[0, 0]) / (xyuvav[1, 2] - xyuvav[0, 2])))

        def match(a, b, x, offs, off, piv, sign, ocw):
            t = time.perf_counter(); out = O.match(a, b, x, offs, off, piv, sign, ocw)[0]; return out, time.perf_counter() - t
        pivots = lambda x, ocw: O.get_uv_pivot(x, dt, mpp, ocw, H, W)
        conv2 = O.conv2

        def cp_stage():
            t = time.perf_counter(); rc, off, _ = O.get_offset_image(i0, i1, xyuvav, 1); return time.perf_counter() - t, rc, off
    n = dimx * dimy
    rows = max(1, min(dimy, int(round(target_nodes / dimx))))
    step = max(1, dimy // rows)
    sel_rows = np.arange(step // 2, dimy, step)[:rows]
    idx = (sel_rows[:, None] * dimx + np.arange(dimx)[None, :]).ravel()
    xs = np.ascontiguousarray(xyuvav[idx])
    piv = [pivots(xs, ocw) for ocw in VEC_OCW]
    dp = np.zeros((32, len(idx), 3), np.float32) if keep_dp else None
    t_match = t_conv = 0.0
    # main allocates i0c / i1c once (zero-initialised under the oracle's allocator shim) and reuses them for the
    # three filters (MIMC_main.c:304-349): the same here, so the stale-border semantics of conv2 are the reference's
    c0 = np.zeros_like(i0); c1 = np.zeros_like(i1)
    for variant in range(4):
        a, b = i0, i1
        if variant > 0:
            t0 = time.perf_counter()
            conv2(i0, variant - 1, c0); conv2(i1, variant - 1, c1)
            t_conv += time.perf_counter() - t0
            a, b = c0, c1
        for c, ocw in enumerate(VEC_OCW):
            off, pv = piv[c]
            k = variant * 8 + c * 2
            d0, s0 = match(a, b, xs, offset, off, pv, +1, ocw)
            d1, s1 = match(b, a, xs, -offset, off, pv, -1, ocw)
            t_match += s0 + s1
            if keep_dp:
                dp[k] = d0
                dp[k + 1] = d1 * np.array([-1, -1, 1], np.float32)    # main negates du, dv of the swapped pass (:289-293)
    del c0, c1
    t_cp, cp_rc, cp_off = (0.0, 0, None)
    if with_cp:
        t_cp, cp_rc, cp_off = cp_stage()
    t_total = t_match * (n / len(idx)) + t_conv + t_cp
    res = {"value": n / t_total, "unit": "nodes/s", "cores": int(ncores), "kind": kind,
           "sample": f"all 32 attempts on {len(idx)} of {n} nodes ({len(sel_rows)} evenly spaced node rows, full-size images; "
                     f"{t_match:.2f} s, scaled linearly in nodes) + the 6 conv2 calls on the full images ({t_conv:.2f} s) + "
                     f"get_offset_image once ({t_cp:.2f} s); postprocess (<0.3 % of CPU time) not included; {ncores} OpenMP threads",
           "t_match_sample_s": t_match, "t_conv2_s": t_conv, "cp_stage_ms": 1e3 * t_cp, "cp_offset": None if cp_off is None else [int(cp_off[0]), int(cp_off[1])],
           "sample_nodes": int(len(idx))}
    if keep_dp:
        res["dp"] = dp; res["idx"] = idx
    return res


def parity_verdict(ref, dp_gpu, peak_gpu, ncell_gpu, i0, i1, xyuvav, dimx, dimy, dt, offset, ctx, params_for):
    """Per-config parity verdict for the bench line: the GPU results of the measured
    workload against the unmodified reference on the sampled nodes (dp of all 32 attempts, bit for bit), against the
    oracle for what the reference keeps internal (integer peaks, evaluated-cell counts; raw-pair attempts), and the
    postprocess of a band of node rows against the oracle's (cluster choice per node)."""
    import oracle
    idx = ref["idx"]
    a = ref["dp"]; b = dp_gpu[:, idx, :]
    na, nb = np.isnan(a), np.isnan(b)
    bad = (na != nb) | (~na & ~nb & (a.view(np.uint32) != b.view(np.uint32)))
    flags_a = np.where(a[..., 2] <= -2.0, a[..., 2], 0.0); flags_b = np.where(b[..., 2] <= -2.0, b[..., 2], 0.0)
    fin = ~na[..., 0] & ~nb[..., 0] & ~na[..., 1] & ~nb[..., 1]
    dsub = float(np.max(np.abs(a[..., :2][fin] - b[..., :2][fin]), initial=0.0))
    okn = np.isfinite(a[..., 2]) & np.isfinite(b[..., 2]) & (a[..., 2] > -2.0)
    drel = float(np.max(np.abs(a[..., 2][okn] - b[..., 2][okn]) / np.maximum(np.abs(a[..., 2][okn]), 1e-30), initial=0.0))
    out = {"reference_nodes": int(len(idx)), "attempts": 32, "dp_values_compared": int(a.size),
           "dp_mismatches": int(bad.sum()), "flag_mismatches": int((flags_a != flags_b).sum()),
           "max_abs_subpixel_diff_px": dsub, "max_rel_ncc_diff": drel}
    # integer peaks / evaluated cells: the oracle on the same sampled nodes, the eight raw-pair attempts
    O = oracle.Oracle()
    H, W = i0.shape
    mpp = float(np.float32((xyuvav[1, 0] - xyuvav[0, 0]) / (xyuvav[1, 2] - xyuvav[0, 2])))
    xs = np.ascontiguousarray(xyuvav[idx])
    pm = nm = 0
    for c, ocw in enumerate(VEC_OCW):
        off, pv = O.get_uv_pivot(xs, dt, mpp, ocw, H, W)
        for d, (x, y, o, sg) in enumerate(((i0, i1, offset, +1), (i1, i0, -offset, -1))):
            _, pk, nc = O.match(x, y, xs, o, off, pv, sg, ocw)
            k = c * 2 + d
            pm += int((pk != peak_gpu[k][idx]).any(axis=1).sum()); nm += int((nc != ncell_gpu[k][idx]).sum())
    out.update(peak_nodes=int(len(idx)), peak_attempts=8, peak_mismatches=pm, ncell_mismatches=nm)
    # cluster choice: postprocess of a contiguous band of node rows taken as a grid of its own, GPU vs oracle
    rows = min(dimy, 24)
    r0 = (dimy - rows) // 2
    band = np.arange(r0 * dimx, (r0 + rows) * dimx)
    xb = np.ascontiguousarray(xyuvav[band]); dpb = np.ascontiguousarray(dp_gpu[:, band, :])
    pp = oracle.post_params(xb, dimx, rows, dt)
    st = O.postprocess_stages(dpb, xb, pp)
    import torch
    pb = params_for(xb, dimx, rows, dt)
    dpd = torch.from_numpy(dpb).cuda(); planes = torch.empty((5, rows, dimx), dtype=torch.float32, device="cuda")
    gstats = ctx.postprocess(dpd, xb, pb, planes)
    gid = ctx.postprocess_stage(4, rows * dimx)
    out.update(cluster_nodes=int(rows * dimx), cluster_mismatches=int((gid != st["ps_id"]).sum()),
               cluster_band_sweeps={"dpf1": int(gstats[0]), "pseudosmoothing": int(gstats[1]),
                                    "oracle_dpf1": int(st["dpf1_sweeps"]), "oracle_pseudosmoothing": int(st["ps_sweeps"])})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-nodes", type=int, default=0, help="nodes in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--band-comm", default="nccl", choices=["nccl", "python"],
                    help="N > 1: halo exchange inside the library (NCCL, default) or through the Python callbacks of bands.py")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("[bench] note: timing rules ask for >= 3 warm-up steps")

    # stdout carries exactly ONE JSON line: libraries that print there (the NCCL version banner, the
    # reference's progress printf's) are sent to stderr; the line itself goes to the saved descriptor
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference" and rank != 0:
        return 0   # rank 0 alone runs the CPU arm
    if world > 1 and "MIMC3CU_HOST_THREADS" not in os.environ:
        # the ranks of one box share its host cores for the pivot generation of the end-to-end leg
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        os.environ["MIMC3CU_HOST_THREADS"] = str(max(1, (os.cpu_count() or 8) // max(1, local_world)))

    import torch
    from mimc3_b200 import synth

    wl = dict(WORKLOADS[args.workload])
    desc = wl.pop("desc")
    strong = bool(wl.pop("strong", False))
    have_gpu = torch.cuda.is_available()
    if args.impl == "ours" and not have_gpu:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dev = torch.device("cuda", local_rank) if have_gpu else torch.device("cpu")
    if have_gpu:
        torch.cuda.set_device(dev)

    dist = None
    if world > 1 and args.impl == "ours":
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=dev)

    t_gen = time.perf_counter()
    sc = synth.make_scene(seed=1234 + (0 if strong else rank), device=dev, **wl)
    H, W = sc.shape
    n = sc.n
    offset = np.array(sc.offset, np.int32)
    log(f"[bench] rank {rank}: scene {H}x{W} {sc.dtype}, grid {sc.dimy}x{sc.dimx} = {n} nodes, generated in {time.perf_counter() - t_gen:.1f} s on {dev}")
    config = {"workload": f"{args.workload}: {desc}", "image": [H, W], "image_dtype": sc.dtype, "nodes_per_gpu": n,
              "node_grid_per_gpu": [sc.dimy, sc.dimx], "attempts_per_node": 32, "chip_half_widths": list(VEC_OCW),
              "l2_policy": "inputs larger than L2 (4 float32 images of %.2f GB each are streamed every step)" % (H * W * 4 / 1e9),
              "parallelism": (f"node-row bands x{world} of one scene" if strong else f"node-row bands x{world} (one tile per GPU)") if world > 1 else "single GPU"}

    # ------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        # the reference's own CPU implementation only: nothing of libmimc3cu.so is loaded by this process
        # (the filtered pairs come from the reference's GMA_float_conv2, timed as part of the unit)
        i0 = sc.i0.cpu().numpy(); i1 = sc.i1.cpu().numpy()
        del sc.i0, sc.i1
        target = args.cpu_sample_nodes or 4 * sc.dimx
        reps = max(1, min(args.steps, 2))        # a bounded sample, not `steps` repetitions of it
        if args.warmup > 0:                       # warm the OpenMP pool on a handful of nodes
            cpu_reference_sample(i0[:1024, :1024].copy(), i1[:1024, :1024].copy(), sc.xyuvav[:8], 8, 1, sc.dt, offset, 8, with_cp=False)
        vals, res = [], None
        for it in range(reps):
            res = cpu_reference_sample(i0, i1, sc.xyuvav, sc.dimx, sc.dimy, sc.dt, offset, target)
            vals.append(res["value"])
        value = float(np.mean(vals))
        line = {"impl": "reference", "metric": "grid nodes matched per second", "value": value, "unit": "nodes/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / value, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 products, f64 accumulation", "data": "synthetic",
                "config": config, "sample_repetitions": reps, "cp_stage_ms": res["cp_stage_ms"], "cp_offset": res["cp_offset"],
                "conv2_s": res["t_conv2_s"],
                "cpu_baseline": {"value": value, "unit": "nodes/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
                "e2e": {"value": value, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit_json(line)
        return 0

    # ------------------------------------------------------------------------------- our arm
    from mimc3_b200 import bands, lib, pipeline
    pl = pipeline.Pipeline(local_rank)
    ctx = pl.ctx
    pl.set_images(sc.i0, sc.i1)
    # ---- the node grid and this rank's band of node rows -------------------------------------------------
    #   N = 1   : the whole grid
    #   weak    : the ranks' grids stacked along y are the bands of one mosaic grid (one tile per GPU)
    #   strong  : one scene; contiguous bands balanced by the matching cost of their rows (sum of P + 2, known
    #             before matching); every rank holds both images
    dimx = sc.dimx
    if world == 1:
        gxy, gdimy, row0, rows = sc.xyuvav, sc.dimy, 0, sc.dimy
        band_rows = [(0, sc.dimy)]
    elif strong:
        gxy, gdimy = sc.xyuvav, sc.dimy
        p0 = lib.params_for(gxy, dimx, gdimy, sc.dt)
        halo = lib.band_halo(p0)
        work = bands.row_work([lib.get_uv_pivot(gxy, sc.dt, p0.mpp, ocw, H, W)[0] for ocw in VEC_OCW], dimx, gdimy)
        band_rows = bands.split_rows(gdimy, world, min_rows=halo, weights=work)
        row0, rows = band_rows[rank]
    else:
        xy_all = [None] * world
        dist.all_gather_object(xy_all, sc.xyuvav)
        gxy = np.concatenate(xy_all, axis=0)
        for r in range(world):     # map-y continues down the mosaic so that the global grid is regular
            gxy[r * sc.n:(r + 1) * sc.n, 1] -= r * sc.dimy * sc.spacing * sc.mpp
        gdimy, row0, rows = sc.dimy * world, rank * sc.dimy, sc.dimy
        band_rows = [(r * sc.dimy, sc.dimy) for r in range(world)]
    n_total = dimx * gdimy if (strong or world == 1) else sc.n * world
    n = rows * dimx                                     # this rank's nodes
    # weak scaling: the rank's own tile is matched with its LOCAL node coordinates; the global grid only drives the postprocess
    xy_band = np.ascontiguousarray(gxy[row0 * dimx:(row0 + rows) * dimx]) if (strong or world == 1) else sc.xyuvav
    gparams = lib.params_for(gxy, dimx, gdimy, sc.dt)
    pl.set_grid(xy_band, dimx, rows, sc.dt)
    params = pl.params
    config["nodes_per_gpu"] = n if world == 1 or not strong else [r * dimx for _, r in band_rows]
    config["nodes_total"] = n_total
    config["node_grid"] = [gdimy, dimx]
    # host copies for the e2e leg (pinned) and the CPU baseline
    np_dt = np.uint8 if sc.dtype == "u8" else np.uint16
    t_dt = torch.uint8 if sc.dtype == "u8" else torch.int16          # int16 storage viewed as uint16 by numpy
    h_i0 = torch.empty((H, W), dtype=t_dt).pin_memory(); h_i1 = torch.empty((H, W), dtype=t_dt).pin_memory()
    i0_host = h_i0.numpy().view(np_dt); i1_host = h_i1.numpy().view(np_dt)
    i0_host[...] = sc.i0.cpu().numpy().astype(np_dt); i1_host[...] = sc.i1.cpu().numpy().astype(np_dt)
    h_xy = torch.from_numpy(xy_band.copy()).pin_memory()
    del sc.i0, sc.i1
    torch.cuda.empty_cache()

    transport = None
    comm_stats = {"halo_exchanges": 0, "allreduces": 0, "steps_counted": 0}
    if world > 1:
        if args.band_comm == "nccl":
            # the library's own communicator: rank 0's NCCL id travels over torch.distributed, every rank attaches
            ids = [lib.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            ctx.comm_init_rank(ids[0], rank, world)
        else:
            transport = bands.DistTransport()

    dp = torch.empty((32, n, 3), dtype=torch.float32, device=dev)
    ncell = torch.empty((32, n), dtype=torch.int32, device=dev)
    want_parity = rank == 0 and world == 1 and not args.no_cpu_baseline
    peaks = torch.empty((32, n, 2), dtype=torch.int32, device=dev) if want_parity else None
    planes = torch.empty((5, gdimy, dimx), dtype=torch.float32, device=dev) if rank == 0 else None
    band_planes = torch.empty((5, rows, dimx), dtype=torch.float32, device=dev) if world > 1 else None
    plane_bytes = np.array([r * dimx * 4 for _, r in band_rows], np.int64)
    hd = pl.handles
    stream = pl.stream

    def postprocess_all(d):
        """mimc2_postprocess over the whole grid -> stats; the five planes end up in `planes` on rank 0."""
        if world == 1:
            return ctx.postprocess(d, gxy, gparams, planes)
        if transport is None:
            # banded postprocess inside the library: halo rows / dirty flags / sweep counters over NCCL on the context's
            # stream, then the final gather of the five planes on rank 0
            st = ctx.postprocess_band(d, gxy, gparams, row0, rows, None, band_planes)
            for k in range(5):
                ctx.comm_gather(band_planes[k], plane_bytes, planes[k] if rank == 0 else None, 0)
            info = ctx.comm_info()
            comm_stats["halo_exchanges"], comm_stats["allreduces"] = info["halo_exchanges"], info["allreduces"]
            comm_stats["steps_counted"] += 1
            return st
        bp, st, comm = pl.postprocess_band(d, gxy, gparams, row0, rows, transport)
        comm_stats["halo_exchanges"] += comm.n_exchanges; comm_stats["allreduces"] += comm.n_allreduce; comm_stats["steps_counted"] += 1
        with torch.cuda.stream(stream):
            # final gather of the five planes (bands may differ in size: point-to-point)
            if rank == 0:
                planes[:, row0:row0 + rows].copy_(bp)
                for r in range(1, world):
                    r0, rr = band_rows[r]
                    for k in range(5):
                        dist.recv(planes[k, r0:r0 + rr], src=r)
            else:
                for k in range(5):
                    dist.send(bp[k].contiguous(), dst=0)
        return st

    def step(collect_ncell=False):
        ctx.multimatch_async(hd["i0"], hd["i1"], hd["i0c"], hd["i1c"], offset, params, dp, ncell if collect_ncell else None,
                             peaks if collect_ncell else None)
        return postprocess_all(dp)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(max(args.warmup, 1)):
        stats = step(collect_ncell=(w == 0))
    ctx.sync()
    E_sum = ncell.to(torch.float64).sum(dim=1).cpu().numpy()          # per attempt
    S2 = np.array([(2 * VEC_OCW[(a % 8) // 2] + 1) ** 2 for a in range(32)], np.float64)
    alg_flop_step = float((8.0 * S2 * E_sum).sum())                    # W = sum E * 8 * S^2
    launches0 = ctx.launch_count()

    sampler = ClockSampler(local_rank)
    ctx.timing_enable(True); ctx.timing_read()
    if world > 1 and transport is None:
        ctx.comm_timing(True)
    barrier()
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        stats = step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    fam_ms, fam_cnt = ctx.timing_read()
    ctx.timing_enable(False)
    comm_ms = ctx.comm_timing(False) if (world > 1 and transport is None) else None
    comm_calls = dict(comm_stats)
    launches = ctx.launch_count() - launches0
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = n_total / (ms_per_step * 1e-3)
    # fingerprint of the five planes of the last timed step (NaNs canonicalised): identical for every N of a
    # strong-scaling workload, i.e. the banded postprocess reproduces the single-GPU result bit for bit
    planes_digest = None
    if rank == 0:
        import hashlib
        ph = planes.cpu().numpy().copy()
        ph[np.isnan(ph)] = np.float32(np.nan)
        planes_digest = hashlib.sha256(ph.tobytes()).hexdigest()
        del ph

    # ---- roofline of the dominant kernel -------------------------------------------------------
    peak_tf, _ = ctx.fp32_peak()
    match_ms_per_launch = fam_ms[0] / max(1, fam_cnt[0])
    achieved_tf = alg_flop_step * args.steps / (fam_ms[0] * 1e-3) / 1e12
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "match_traffic.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get(args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "fp32", "kernel": "match2_kernel<ocw,G> (exact-FP32 DLC-NCC matcher; match_kernel for nodes outside its class)", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved_tf / peak_tf, "traffic": traffic,
                "peak_source": "FP32 FMA micro-benchmark measured live on this GPU (mimc3cu_fp32_peak); MEASURED_PEAKS.json has no CUDA-core FP32 figure",
                # driver-independent denominator: SMs x 128 FP32 lanes x 2 flop x the maximum SM clock nvidia-smi reports
                "peak_nominal": 148 * 128 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12,
                "frac_nominal": achieved_tf / (148 * 128 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12),
                "algorithmic_flop_per_step": alg_flop_step, "algorithmic_flop_per_launch": alg_flop_step / 32.0,
                "launches_per_step": 32, "avg_launch_ms": match_ms_per_launch,
                "kernel_share_of_step": fam_ms[0] / ms_total, "conv2_share_of_step": fam_ms[1] / ms_total,
                "postprocess_share_of_step": fam_ms[2] / ms_total,
                "mean_cells_per_node_attempt": float(E_sum.sum() / (32.0 * n))}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        # secondary view: compulsory bytes of the matcher vs measured HBM copy bandwidth
        cells = E_sum.sum() / 32.0 / n
        bytes_step = sum(4.0 * ((2 * o + 1) ** 2 + (2 * (o + 12) + 1) * (2 * (o + 12) + 1)) * n * 8 for o in VEC_OCW)
        roofline["hbm_view"] = {"compulsory_GB_per_step": bytes_step / 1e9, "achieved_GBps": bytes_step * args.steps / (fam_ms[0] * 1e-3) / 1e9,
                                "peak_GBps": mp["hbm_gbs"], "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)"}
    except Exception:
        pass

    # ---- end-to-end leg: host buffers through the C ABI ---------------------------------------------
    e2e = None
    if not args.no_e2e:
        xy_host = h_xy.numpy()
        planes_host = torch.empty((5, gdimy, dimx), dtype=torch.float32).pin_memory().numpy() if rank == 0 else None

        cp_ms = []

        def e2e_step():
            pl.set_images(i0_host, i1_host)                              # H2D + on-device cast
            off_e2e = offset
            if world == 1:
                # control-point stage (get_offset_image, MIMC_module.c:33-492) on the GPU: its offset feeds the matcher
                tcp = time.perf_counter()
                rc_cp, off_cp, _, _ = ctx.get_offset_image(hd["i0"], hd["i1"], xy_host, gparams, 1)
                cp_ms.append(1e3 * (time.perf_counter() - tcp))
                if rc_cp != 1 or tuple(off_cp) != tuple(offset):
                    raise SystemExit(f"bench.py: control-point stage returned {rc_cp}, offset {off_cp} (scene offset {offset})")
                off_e2e = off_cp
            d, _ = pl.match_all(xy_host, dimx, rows, sc.dt, off_e2e)   # nodes + host pivots + H2D overlapped with the attempts
            postprocess_all(d)
            if rank == 0:
                ctx.finalize(planes, gparams)
                ctx._ck(ctx.L.mimc3cu_memcpy_d2h(ctx.h, planes_host.ctypes.data, planes.data_ptr(), planes_host.nbytes))
                return float(np.nanmean(planes_host[4]))
            ctx.sync()
            return 0.0
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, args.steps)
        cp_ms.clear()
        for _ in range(e2e_steps):
            qual = e2e_step()
        barrier()
        dt_e2e = (time.perf_counter() - t0) / e2e_steps
        if dist is not None:
            tt = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt_e2e = float(tt.item())
        e2e = {"value": n_total / dt_e2e, "unit": "nodes/s",
               "h2d_bytes_per_step": int(world * (i0_host.nbytes + i1_host.nbytes) + 6 * 8 * n_total + world * pl.pivot_bytes),
               "d2h_bytes_per_step": int(5 * 4 * n_total), "ms_per_step": dt_e2e * 1e3, "steps": e2e_steps,
               "timing": "host wall clock between device synchronisations, max over ranks (the leg includes the control-point stage at N=1 and host pivot generation)",
               "cp_stage_ms": float(np.mean(cp_ms)) if cp_ms else None, "cp_stage_ms_per_step": [round(x, 2) for x in cp_ms],
               "mean_support": qual}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1) + parity verdict on the measured scene ---------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        i0f = ctx.image_download(hd["i0"], H, W); i1f = ctx.image_download(hd["i1"], H, W)
        target = args.cpu_sample_nodes or 4 * sc.dimx
        res = cpu_reference_sample(i0f, i1f, sc.xyuvav, sc.dimx, sc.dimy, sc.dt, offset, target, keep_dp=True)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "cp_stage_ms", "cp_offset")}
        # the GPU results of the measured workload against the reference's on the sampled nodes
        ctx.multimatch_async(hd["i0"], hd["i1"], hd["i0c"], hd["i1c"], offset, params, dp, ncell, peaks)
        ctx.sync()
        parity = parity_verdict(res, dp.cpu().numpy(), peaks.cpu().numpy(), ncell.cpu().numpy(), i0f, i1f, sc.xyuvav, sc.dimx,
                                sc.dimy, sc.dt, offset, ctx, lib.params_for)
        parity["workload"] = args.workload
        del i0f, i1f

    if rank == 0:
        line = {"metric": "grid nodes matched per second", "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "f32 (exact two-float accumulation of the float products, f64 normalisation; bit-exact vs the reference)", "data": "synthetic", "config": config,
                "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "collectives": (dict(comm_stats, backend="nccl", issued_by=("libmimc3cu.so (ncclSend/ncclRecv/ncclAllReduce on the context stream)" if transport is None else "bands.py callbacks over torch.distributed"),
                                     pattern="neighbour halo rows + int32 counter all-reduce per sweep, final gather of 5 planes",
                                     postprocess_ms_per_step=fam_ms[2] / max(1, args.steps),
                                     halo_exchange_ms_per_step=(comm_ms[0] / args.steps if comm_ms else None),
                                     allreduce_ms_per_step=(comm_ms[1] / args.steps if comm_ms else None)) if world > 1 else None),
                "planes_sha256": planes_digest,
                "postprocess_stats": {"dpf1_sweeps": int(stats[0]), "pseudosmoothing_sweeps": int(stats[1]), "holes_after_dpf0": int(stats[2])} if stats is not None else None}
        emit_json(line)
    pl.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
