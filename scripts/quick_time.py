"""Per-chip-size matcher timing on one GPU (development aid): python scripts/quick_time.py [c1|c2s|c4s] [v1|v2]"""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mimc3_b200 import lib, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "c1"
mode = sys.argv[2] if len(sys.argv) > 2 else "auto"
ocws = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [7, 15, 30, 40]
cfg = dict(c1=dict(H=2048, W=2048, dtype="u8", spacing=19), c2s=dict(H=4096, W=4096, dtype="u16", spacing=20),
           c4s=dict(H=4096, W=4096, dtype="u16", spacing=20, peak_px=43.0, apriori_gain=0.9, band_width_frac=0.2))[wl]
sc = synth.make_scene(seed=1, device="cuda", **cfg)
ctx = lib.Context(0)
ctx.set_matcher(mode)
p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
H, W = sc.shape
ctx.set_nodes(sc.xyuvav)
a, b = ctx.image_from(sc.i0), ctx.image_from(sc.i1)
n = sc.n
dp = torch.empty((n, 3), device="cuda"); nc = torch.empty(n, dtype=torch.int32, device="cuda")
st = torch.cuda.ExternalStream(ctx.stream)
tot = 0.0; flop = 0.0
for slot, ocw in enumerate(ocws):
    t = time.time(); off, piv = lib.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, H, W); tp = time.time() - t
    ctx.set_pivots(slot, off, piv)
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(st); ctx.match_async(a, b, np.array(sc.offset, np.int32), slot, 1, ocw, False, dp, None, nc); e1.record(st); ctx.sync()
    ms = e0.elapsed_time(e1); E = nc.float().mean().item()
    S = 2 * ocw + 1
    tot += ms; flop += 8 * S * S * E * n
    print(f"{wl} {mode} ocw {ocw}: {ms:.3f} ms  {n/ms*1e3:.0f} node-attempts/s  E={E:.1f}  alg={8*S*S*E*n/ms*1e3/1e12:.2f} TF/s  matcher={ctx.last_matcher()} pivots_host={tp*1e3:.1f}ms")
print(f"{wl} {mode} total {tot:.3f} ms  alg={flop/tot*1e3/1e12:.2f} TF/s")
