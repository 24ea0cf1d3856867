import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mimc3_b200 import lib, synth
sc = synth.make_scene(H=2048, W=2048, dtype="u8", spacing=19, seed=1, device="cuda")
ctx = lib.Context(0)
p = lib.params_for(sc.xyuvav, sc.dimx, sc.dimy, sc.dt)
H,W = sc.shape
ctx.set_nodes(sc.xyuvav)
a, b = ctx.image_from(sc.i0), ctx.image_from(sc.i1)
n = sc.n
dp = torch.empty((n,3), device="cuda"); nc = torch.empty(n, dtype=torch.int32, device="cuda")
st = torch.cuda.ExternalStream(ctx.stream)
for slot, ocw in enumerate((7,15,30,40)):
    t=time.time(); off, piv = lib.get_uv_pivot(sc.xyuvav, sc.dt, p.mpp, ocw, H, W); tp=time.time()-t
    ctx.set_pivots(slot, off, piv)
    for rep in range(2):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(st); ctx.match_async(a,b,np.array(sc.offset,np.int32),slot,1,ocw,False,dp,None,nc); e1.record(st); ctx.sync()
    ms=e0.elapsed_time(e1); E=nc.float().mean().item()
    S=2*ocw+1
    print(f"ocw {ocw}: {ms:.2f} ms  {n/ms*1e3:.0f} node-attempts/s  E={E:.1f}  algflops={8*S*S*E*n/ms*1e3/1e12:.2f} TF/s pivots_host={tp*1e3:.1f}ms")
