# per-ocw timing on the C2 scene itself (the bins matter for the band nodes, which the 4096^2 scene barely has)
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
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
ocws = [int(x) for x in sys.argv[1].split(",")]
for slot, ocw in enumerate(ocws):
    off, piv = lib.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, H, W)
    ctx.set_pivots(slot, off, piv)
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); ctx.match_async(a, b, np.array(sc.offset, np.int32), slot, 1, ocw, False, dp, None, nc); e1.record(st); ctx.sync()
    print(f"C2 ocw {ocw}: {e0.elapsed_time(e1):.3f} ms")
