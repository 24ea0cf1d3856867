#!/usr/bin/env python
"""Benchmark of the MIMC3 per-grid-node matching hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c2|c1|c4]

Metric (BASELINE.json): grid nodes matched per second.  One node = all 32 matching
attempts (4 chip sizes x {forward, swapped} x {raw, d/dx, d/dy, Laplacian}) + its share of
the 6 conv2 passes + the postprocess (cluster, dpf0/dpf1, pseudosmoothing).  A "step" is
one pass of that path over every node of the workload.

  value  device-timed (CUDA events on the library's stream, max over ranks), inputs
         (both images, nodes, pivots) already resident in HBM.
  e2e    the same metric through the C-ABI host entry points with HOST buffers: pinned
         u16/u8 images and xyuvav copied H2D, host pivot generation, multi-match,
         postprocess, finalize, five planes copied D2H -- all inside the timed region.
  roofline      dominant kernel = the matcher (match2_kernel<ocw,G>, plus match_kernel for nodes
                outside its class); achieved = sum over attempts and nodes of 8*S^2*E flop
                (E = NCC cells the reference algorithm evaluates, counted by the kernel and
                cross-checked against the oracle in tests) / summed matcher durations (CUDA
                events recorded by the library around every attempt inside the timed
                region); peak = FP32 FMA throughput measured live by an FMA micro-benchmark
                (MEASURED_PEAKS.json has no FP32 CUDA-core figure).
  cpu_baseline  the UNMODIFIED reference (oracle/_ref/libmimc3ref.so, OpenMP, all host
                threads) on a bounded node sample of the same workload.

N > 1 (torchrun): weak scaling.  Every rank owns one tile of a vertical mosaic (its own
image pair + node-row band) and matches it with no data-path collective.  The postprocess
runs banded over the whole mosaic grid: each rank sweeps its own band and exchanges halo
rows / dirty flags / sweep counters with its neighbours over NCCL (mimc3_b200/bands.py),
then the five planes are gathered on rank 0.  Timing = max over ranks.
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
def cpu_reference_sample(i0, i1, filtered, xyuvav, dimx, dimy, dt, offset, target_nodes):
    """Times the unmodified reference on a bounded sample: all 32 matching attempts on every
    k-th node row (full-size images), plus conv2 on a row band scaled to the full image.
    `filtered` = list of three (i0c, i1c) host pairs (bit-identical to the reference's conv2,
    asserted by tests/test_conv2_gpu.py).  Returns dict with nodes/s for the whole workload."""
    import oracle
    H, W = i0.shape
    kind = "reference"
    try:
        R = oracle.Reference()
        ncores = R.num_threads()
        R.set_globals(xyuvav, dimx, dimy, dt)
        mpp = float(np.float32((xyuvav[1, 0] - xyuvav[0, 0]) / (xyuvav[1, 2] - xyuvav[0, 2])))
        match = lambda a, b, x, offs, off, piv, sign, ocw: R.match(a, b, x, offs, off, piv, sign, ocw)[1]
        pivots = lambda x, ocw: R.get_uv_pivot(x, dt, ocw, H, W)
        conv2 = R.conv2
    except (FileNotFoundError, OSError) as e:
        log(f"[bench] reference build not available ({e}); timing the oracle port instead")
        kind = "port"
        O = oracle.Oracle()
        ncores = O.num_threads()
        mpp = float(np.float32((xyuvav[1, 0] - xyuvav[0, 0]) / (xyuvav[1, 2] - xyuvav[0, 2])))

        def match(a, b, x, offs, off, piv, sign, ocw):
            t = time.perf_counter(); O.match(a, b, x, offs, off, piv, sign, ocw); return time.perf_counter() - t
        pivots = lambda x, ocw: O.get_uv_pivot(x, dt, mpp, ocw, H, W)
        conv2 = O.conv2
    n = dimx * dimy
    rows = max(1, min(dimy, int(round(target_nodes / dimx))))
    step = max(1, dimy // rows)
    sel_rows = np.arange(step // 2, dimy, step)[:rows]
    idx = (sel_rows[:, None] * dimx + np.arange(dimx)[None, :]).ravel()
    xs = np.ascontiguousarray(xyuvav[idx])
    t_match = 0.0
    pairs = [(i0, i1)] + list(filtered)
    for (a, b) in pairs:
        for ocw in VEC_OCW:
            off, piv = pivots(xs, ocw)
            t_match += match(a, b, xs, offset, off, piv, +1, ocw)
            t_match += match(b, a, xs, -offset, off, piv, -1, ocw)
    # conv2: 6 calls on a band of rows, scaled by the pixel ratio (the loop is O(H*W), single-threaded)
    band = min(H, 1024)
    src = np.ascontiguousarray(i0[:band]); dst = np.zeros_like(src)
    t0 = time.perf_counter()
    for k in range(3):
        conv2(src, k, dst)
    t_conv_band = (time.perf_counter() - t0) * 2.0   # two images
    t_conv_full = t_conv_band * (H / band)
    t_total = t_match * (n / len(idx)) + t_conv_full
    return {"value": n / t_total, "unit": "nodes/s", "cores": int(ncores), "kind": kind,
            "sample": f"all 32 attempts on {len(idx)} of {n} nodes ({len(sel_rows)} evenly spaced node rows, full-size images; "
                      f"{t_match:.2f} s, scaled linearly in nodes) + 6 conv2 calls on a {band}-row band scaled by H/{band} "
                      f"({t_conv_full:.2f} s est.); postprocess (<0.3 % of CPU time) not included",
            "t_match_sample_s": t_match, "t_conv2_full_est_s": t_conv_full, "sample_nodes": int(len(idx))}


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
    sc = synth.make_scene(seed=1234 + rank, device=dev, **wl)
    H, W = sc.shape
    n = sc.n
    offset = np.array(sc.offset, np.int32)
    log(f"[bench] rank {rank}: scene {H}x{W} {sc.dtype}, grid {sc.dimy}x{sc.dimx} = {n} nodes, generated in {time.perf_counter() - t_gen:.1f} s on {dev}")
    config = {"workload": f"{args.workload}: {desc}", "image": [H, W], "image_dtype": sc.dtype, "nodes_per_gpu": n,
              "node_grid_per_gpu": [sc.dimy, sc.dimx], "attempts_per_node": 32, "chip_half_widths": list(VEC_OCW),
              "l2_policy": "inputs larger than L2 (4 float32 images of %.2f GB each are streamed every step)" % (H * W * 4 / 1e9),
              "parallelism": f"node-row bands x{world}" if world > 1 else "single GPU"}

    # ------------------------------------------------------------------------------- reference arm
    if args.impl == "reference":
        i0 = sc.i0.cpu().numpy(); i1 = sc.i1.cpu().numpy()
        filtered = []
        if have_gpu:
            from mimc3_b200 import pipeline
            import oracle
            pl = pipeline.Pipeline(local_rank)
            pl.set_images(sc.i0, sc.i1)
            for k in range(3):
                pl.ctx.conv2(pl.handles["i0"], oracle.KERNELS[k], pl.handles["i0c"])
                pl.ctx.conv2(pl.handles["i1"], oracle.KERNELS[k], pl.handles["i1c"])
                filtered.append((pl.ctx.image_download(pl.handles["i0c"], H, W), pl.ctx.image_download(pl.handles["i1c"], H, W)))
            pl.close()
        else:
            import oracle
            O = oracle.Oracle()
            c0 = np.zeros_like(i0); c1 = np.zeros_like(i1)
            for k in range(3):
                O.conv2(i0, k, c0); O.conv2(i1, k, c1)
                filtered.append((c0.copy(), c1.copy()))
        del sc.i0, sc.i1
        target = args.cpu_sample_nodes or 4 * sc.dimx
        res = None
        times = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = cpu_reference_sample(i0, i1, filtered, sc.xyuvav, sc.dimx, sc.dimy, sc.dt, offset, target)
            if it >= args.warmup:
                times.append((time.perf_counter() - t0, res["value"]))
        value = float(np.mean([v for _, v in times]))
        line = {"impl": "reference", "metric": "grid nodes matched per second", "value": value, "unit": "nodes/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n / value, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 products, f64 accumulation", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": value, "unit": "nodes/s", "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
                "e2e": {"value": value, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit_json(line)
        return 0

    # ------------------------------------------------------------------------------- our arm
    from mimc3_b200 import lib, pipeline
    pl = pipeline.Pipeline(local_rank)
    ctx = pl.ctx
    pl.set_images(sc.i0, sc.i1)
    pl.set_grid(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
    params = pl.params
    # host copies for the e2e leg (pinned) and the CPU baseline
    np_dt = np.uint8 if sc.dtype == "u8" else np.uint16
    t_dt = torch.uint8 if sc.dtype == "u8" else torch.int16          # int16 storage viewed as uint16 by numpy
    h_i0 = torch.empty((H, W), dtype=t_dt).pin_memory(); h_i1 = torch.empty((H, W), dtype=t_dt).pin_memory()
    i0_host = h_i0.numpy().view(np_dt); i1_host = h_i1.numpy().view(np_dt)
    i0_host[...] = sc.i0.cpu().numpy().astype(np_dt); i1_host[...] = sc.i1.cpu().numpy().astype(np_dt)
    h_xy = torch.from_numpy(sc.xyuvav.copy()).pin_memory()
    h_planes = torch.empty((5, sc.dimy, sc.dimx), dtype=torch.float32).pin_memory()
    del sc.i0, sc.i1
    torch.cuda.empty_cache()

    # global (mosaic) grid for N > 1: the ranks' grids stacked along y = the bands of one grid
    if world > 1:
        from mimc3_b200 import bands
        gl_dimy = sc.dimy * world
        xy_all = [None] * world
        dist.all_gather_object(xy_all, sc.xyuvav)
        xy_glob = np.concatenate(xy_all, axis=0)
        # make map-y continue down the mosaic so the global grid is regular
        for r in range(world):
            xy_glob[r * n:(r + 1) * n, 1] -= r * sc.dimy * sc.spacing * sc.mpp
        gparams = lib.params_for(xy_glob, sc.dimx, gl_dimy, sc.dt)
        transport = bands.DistTransport()
        comm_stats = {"halo_exchanges": 0, "allreduces": 0, "steps_counted": 0}

    dp = torch.empty((32, n, 3), dtype=torch.float32, device=dev)
    ncell = torch.empty((32, n), dtype=torch.int32, device=dev)
    planes = torch.empty((5, sc.dimy * (world if (world > 1 and rank == 0) else 1), sc.dimx), dtype=torch.float32, device=dev)
    hd = pl.handles
    stream = pl.stream

    def step(collect_ncell=False):
        ctx.multimatch_async(hd["i0"], hd["i1"], hd["i0c"], hd["i1c"], offset, params, dp, ncell if collect_ncell else None)
        if world == 1:
            return ctx.postprocess(dp, sc.xyuvav, params, planes)
        # banded postprocess: halo exchange + counter all-reduce per sweep over NCCL, then the final gather
        band_planes, st, comm = pl.postprocess_band(dp, xy_glob, gparams, rank * sc.dimy, sc.dimy, transport)
        comm_stats["halo_exchanges"] += comm.n_exchanges; comm_stats["allreduces"] += comm.n_allreduce; comm_stats["steps_counted"] += 1
        with torch.cuda.stream(stream):
            parts = transport.gather_rows(band_planes, dst=0)
            if rank == 0:
                torch.cat(parts, dim=1, out=planes)
        return st

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
    launches = ctx.launch_count() - launches0
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = n * world / (ms_per_step * 1e-3)

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
        planes_host = (torch.empty((5, sc.dimy * world, sc.dimx), dtype=torch.float32).pin_memory().numpy()
                       if (world > 1 and rank == 0) else h_planes.numpy())

        def e2e_step():
            pl.set_images(i0_host, i1_host)                              # H2D + on-device cast
            d, _ = pl.match_all(xy_host, sc.dimx, sc.dimy, sc.dt, offset)   # nodes + host pivots + H2D overlapped with the attempts
            if world == 1:
                pln, _ = pl.postprocess(d)
                ctx.finalize(pln, pl.params)
            else:
                band, _, _ = pl.postprocess_band(d, xy_glob, gparams, rank * sc.dimy, sc.dimy, transport)
                with torch.cuda.stream(stream):
                    parts = transport.gather_rows(band, dst=0)
                    pln = torch.cat(parts, dim=1) if rank == 0 else None
                stream.synchronize()
                if rank == 0:
                    ctx.finalize(pln, gparams)
            if rank == 0:
                ctx._ck(ctx.L.mimc3cu_memcpy_d2h(ctx.h, planes_host.ctypes.data, pln.data_ptr(), planes_host.nbytes))
                return float(np.nanmean(planes_host[4]))
            return 0.0
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 2))
        for _ in range(e2e_steps):
            qual = e2e_step()
        barrier()
        dt_e2e = (time.perf_counter() - t0) / e2e_steps
        if dist is not None:
            tt = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt_e2e = float(tt.item())
        e2e = {"value": n * world / dt_e2e, "unit": "nodes/s",
               "h2d_bytes_per_step": int(world * (i0_host.nbytes + i1_host.nbytes + xy_host.nbytes + pl.pivot_bytes)),
               "d2h_bytes_per_step": int(5 * 4 * n * world), "ms_per_step": dt_e2e * 1e3, "steps": e2e_steps,
               "timing": "host wall clock between device synchronisations, max over ranks (the leg includes host pivot generation)",
               "mean_support": qual}

    # ---- CPU baseline on this box's host cores (rank 0, N = 1) ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        i0f = ctx.image_download(hd["i0"], H, W); i1f = ctx.image_download(hd["i1"], H, W)
        filtered = []
        zero = torch.zeros((H, W), dtype=torch.float32, device=dev)
        for h in (hd["i0c"], hd["i1c"]):
            ctx._ck(ctx.L.mimc3cu_image_copy_from_device(ctx.h, h, zero.data_ptr()))
        for k in range(3):
            ctx.conv2(hd["i0"], oracle.KERNELS[k], hd["i0c"]); ctx.conv2(hd["i1"], oracle.KERNELS[k], hd["i1c"])
            filtered.append((ctx.image_download(hd["i0c"], H, W), ctx.image_download(hd["i1c"], H, W)))
        target = args.cpu_sample_nodes or 4 * sc.dimx
        res = cpu_reference_sample(i0f, i1f, filtered, sc.xyuvav, sc.dimx, sc.dimy, sc.dt, offset, target)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "grid nodes matched per second", "value": value, "unit": "nodes/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (exact two-float accumulation of the float products, f64 normalisation; bit-exact vs the reference)", "data": "synthetic", "config": config,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "collectives": (dict(comm_stats, backend="nccl", pattern="neighbour halo rows + int32 counter all-reduce per sweep, final gather of 5 planes") if world > 1 else None),
                "postprocess_stats": {"dpf1_sweeps": int(stats[0]), "pseudosmoothing_sweeps": int(stats[1]), "holes_after_dpf0": int(stats[2])} if stats is not None else None}
        emit_json(line)
    pl.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
