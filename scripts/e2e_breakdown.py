"""Wall-clock breakdown of the end-to-end leg (host buffers -> planes) at the C2 size (development aid)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from mimc3_b200 import lib, synth, pipeline
wl = dict(H=16384, W=16384, dtype="u16", spacing=20, mpp=15.0, peak_px=6.3)
if len(sys.argv) > 1 and sys.argv[1] == "small":
    wl.update(H=4096, W=4096)
sc = synth.make_scene(seed=1234, device="cuda", **wl)
H, W = sc.shape
i0 = torch.empty((H, W), dtype=torch.int16).pin_memory(); i1 = torch.empty((H, W), dtype=torch.int16).pin_memory()
i0h = i0.numpy().view(np.uint16); i1h = i1.numpy().view(np.uint16)
i0h[...] = sc.i0.cpu().numpy().astype(np.uint16); i1h[...] = sc.i1.cpu().numpy().astype(np.uint16)
xy = sc.xyuvav.copy()
del sc.i0, sc.i1
torch.cuda.empty_cache()
pl = pipeline.Pipeline(0)
offset = np.array(sc.offset, np.int32)
planes_host = torch.empty((5, sc.dimy, sc.dimx), dtype=torch.float32).pin_memory().numpy()
def T():
    pl.ctx.sync(); torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t = [T()]
    pl.set_images(i0h, i1h); t.append(T())
    pl.set_grid(xy, sc.dimx, sc.dimy, sc.dt); t.append(T())
    d, _ = pl.multimatch(offset); t.append(T())
    pln, _ = pl.postprocess(d); t.append(T())
    pl.ctx.finalize(pln, pl.params); t.append(T())
    pl.ctx._ck(pl.ctx.L.mimc3cu_memcpy_d2h(pl.ctx.h, planes_host.ctypes.data, pln.data_ptr(), planes_host.nbytes)); t.append(T())
    names = ["set_images", "set_grid", "multimatch", "postprocess", "finalize", "d2h"]
    print(f"rep {rep}: total {t[-1]-t[0]:.3f} s  " + "  ".join(f"{n}={(b-a)*1e3:.0f}ms" for n, a, b in zip(names, t[:-1], t[1:])))
# finer: set_grid parts
t0 = T(); pl.ctx.set_nodes(xy); t1 = T()
print(f"set_nodes {1e3*(t1-t0):.0f} ms")
for slot, ocw in enumerate((7, 15, 30, 40)):
    t0 = T(); off, piv = lib.get_uv_pivot(xy, sc.dt, pl.params.mpp, ocw, H, W); t1 = T(); pl.ctx.set_pivots(slot, off, piv); t2 = T()
    print(f"ocw {ocw}: get_uv_pivot {1e3*(t1-t0):.0f} ms, set_pivots {1e3*(t2-t1):.0f} ms, total pivots {len(piv)}")
