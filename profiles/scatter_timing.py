"""Device time of the M-Mbulge scatter stages (K6) at the named grid:  python profiles/scatter_timing.py"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch, scipy.stats
import holodeck_b200 as holo
from holodeck_b200 import _lib
from holodeck_b200.sams import scatter
from holodeck_b200.constants import MSOL

M, Q, Z = 91, 81, 101
mtot = np.logspace(*np.log10([1e4*MSOL, 1e12*MSOL]), M)
mrat = np.logspace(-3, 0, Q)
rng = np.random.default_rng(0)
lm = np.log10(mtot/MSOL)[:, None, None]
dens = 1e-3*np.exp(-0.5*((lm - 8.0)/1.0)**2) * np.ones((M, Q, Z)) * rng.uniform(0.8, 1.2, (M, Q, Z))
lib = _lib.require_gpu()
t0 = time.perf_counter(); gg = scatter._device_geometry(mtot, mrat, 4); print("geometry (host, once per grid): %.0f ms" % ((time.perf_counter()-t0)*1e3))
d = _lib.to_dev(dens)
for rep in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(); out = scatter.add_scatter_to_masses(mtot, mrat, d, 0.3); ev[1].record(); torch.cuda.synchronize()
    print("add_scatter_to_masses total: %.2f ms" % ev[0].elapsed_time(ev[1]))
npts, G = gg["npts"], gg["G"]
data = d.reshape(npts, Z)
def tm(fn, n=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best, r
grad = _lib.empty((npts, 2, Z)); niter = torch.empty(Z, dtype=torch.int32, device="cuda")
t, _ = tm(lambda: lib.holo_scatter_gradients(npts, Z, _lib.ptr(gg["program"]), gg["nsteps"], _lib.ptr(data), 400, 1e-6, _lib.ptr(grad), _lib.ptr(niter), _lib.stream()))
print("K6a gradients: %.2f ms  (steps %d, sweeps max %d)" % (t, gg["nsteps"], int(niter.max())))
grid = _lib.empty((G, G, Z)); flags = torch.zeros(1, dtype=torch.int32, device="cuda")
t, _ = tm(lambda: lib.holo_scatter_ct_eval(G*G, Z, _lib.ptr(gg["geo"]), _lib.ptr(data), _lib.ptr(grad), _lib.ptr(grid), _lib.ptr(flags), _lib.stream()))
print("K6b CT eval + fill: %.2f ms" % t)
w2t = gg[("weights", 0.3)]
t, g2 = tm(lambda: torch.matmul(w2t, grid.reshape(G, G*Z)))
print("full DGEMM (%d x %d) x (%d x %d): %.2f ms = %.1f TFLOP/s" % (G, G, G, G*Z, t, 2*G*G*G*Z/t/1e9))
for nblk in (2, 3, 4, 6, 8, 16):
    blocks = scatter._gemm_blocks(gg["i0"].cpu().numpy(), gg["i1"].cpu().numpy(), G, nblk)
    a2 = grid.reshape(G, G*Z); sc = _lib.empty((G, G*Z))
    def blocked():
        for k0, k1, b0, b1 in blocks:
            torch.mm(w2t[k0:k1], a2[:, b0*Z:b1*Z], out=sc[k0:k1, b0*Z:b1*Z])
        return sc
    t, g3 = tm(blocked)
    # device-only time: queue the calls behind a long kernel so that the host is never the bottleneck
    big = torch.empty((8192, 8192), device="cuda"); tq = []
    for rep in range(3):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.mm(big, big); e0.record(); blocked(); e1.record(); torch.cuda.synchronize(); tq.append(e0.elapsed_time(e1))
    tdev = min(tq)
    fl = sum(2.0*(k1-k0)*G*(b1-b0)*Z for k0, k1, b0, b1 in blocks)
    err = max(float((g3[k0:k1, b0*Z:b1*Z] - g2[k0:k1, b0*Z:b1*Z]).abs().max()) for k0, k1, b0, b1 in blocks)
    print("blocked DGEMM, %2d column blocks: %.3f ms launched live, %.3f ms of device time (%.0f %% of the flops, %.1f TFLOP/s), max |diff| to the full product %.1e" % (len(blocks), t, tdev, 100*fl/(2.0*G*G*G*Z), fl/tdev/1e9, err))
outp = _lib.empty((npts, Z))
t, _ = tm(lambda: lib.holo_scatter_bilinear(npts, G, Z, _lib.ptr(gg["i0"]), _lib.ptr(gg["i1"]), _lib.ptr(gg["y0"]), _lib.ptr(gg["y1"]), _lib.ptr(g2), _lib.ptr(outp), _lib.stream()))
print("K6c bilinear: %.2f ms" % t)
