# matcher time of one chip half-width on the C2 scene under several shared-memory bin tables (MIMC3CU_BINS), one process
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mimc3_b200 import lib, synth
import bench
wl = dict(bench.WORKLOADS["c2"]); wl.pop("desc")
sc = synth.make_scene(seed=1234, device="cuda", **wl)
ctx = lib.Context(0)
p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
H, W = sc.shape
ctx.set_nodes(sc.xyuvav)
a, b = ctx.image_from(sc.i0), ctx.image_from(sc.i1)
n = sc.n
dp = torch.empty((n, 3), device="cuda"); nc = torch.empty(n, dtype=torch.int32, device="cuda")
st = torch.cuda.ExternalStream(ctx.stream)
os.environ["MIMC3CU_DEBUG_BINS"] = "1"
ref = {}
for spec in sys.argv[1:]:
    ocw = int(spec.split(":")[0]); bins = spec.partition("=")[2]
    os.environ["MIMC3CU_BINS"] = bins
    off, piv = lib.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, H, W)
    ctx.set_pivots(0, off, piv)
    best = 1e9
    for rep in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); ctx.match_async(a, b, np.array(sc.offset, np.int32), 0, 1, ocw, False, dp, None, nc); e1.record(st); ctx.sync()
        if rep: best = min(best, e0.elapsed_time(e1))
    h = (float(torch.nan_to_num(dp).double().sum()), int(nc.sum()))
    ref.setdefault(ocw, h)
    print(f"C2 ocw {ocw} bins '{bins}': {best:.3f} ms  same results: {h == ref[ocw]}", flush=True)
