"""Where a step of the K6 gradients kernel spends its cycles (thread 0 of every CTA, clock64 between the segments).
Needs the profiling build:  make -C holodeck_b200/csrc phase ;  python profiles/gs_clocks.py"""
import os, sys, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("HOLO_B200_LIB", str(ROOT / "build" / "libholo_b200_phase.so"))
import numpy as np, torch
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
lib.holo_debug_gs_clocks.argtypes = [C.c_void_p, C.c_int]
gg = scatter._device_geometry(mtot, mrat, 4)
npts = gg["npts"]
data = _lib.to_dev(dens).reshape(npts, Z)
grad = _lib.empty((npts, 2, Z)); niter = torch.empty(Z, dtype=torch.int32, device="cuda")
run = lambda: lib.holo_scatter_gradients(npts, Z, _lib.ptr(gg["program"]), gg["nsteps"], _lib.ptr(data), 400, 1e-6, _lib.ptr(grad), _lib.ptr(niter), _lib.stream())
run(); torch.cuda.synchronize()
buf = (C.c_ulonglong * 8)()
lib.holo_debug_gs_clocks(buf, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(); b.record(); torch.cuda.synchronize()
lib.holo_debug_gs_clocks(buf, 1)
steps = int(niter.sum().item()) * gg["nsteps"]
names = ["gather + poll", "edge terms", "fetch next record", "reduce + update", "barrier", "sweep end"]
print("instrumented launch %.2f ms; %d sweeps over %d slices; cycles per step (thread 0, mean over slices):" % (a.elapsed_time(b), int(niter.sum().item()), Z))
for i, n in enumerate(names):
    print("  %-20s %7.1f" % (n, buf[i] / steps))
print("  %-20s %7.1f" % ("total", sum(buf[:6]) / steps))
print("  steps whose next record had not landed when polled: %.1f %%" % (100.0 * buf[6] / steps))
