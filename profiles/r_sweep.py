"""Device time of the realization kernels against the realization count R (fixed per-pass cost vs marginal
cost per realization): python profiles/r_sweep.py"""
import sys, argparse, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, numpy as np
import bench
from holodeck_b200 import _lib, gravwaves, single_sources, cosmo, utils, cyutils
from holodeck_b200.sams import sam_cyutils
from holodeck_b200.constants import YR

args = argparse.Namespace(shape=[91, 81, 101], nfreqs=40, realize=1000, loudest=1)
fobs_cents, fobs_edges = utils.pta_freqs(16.03*YR, 40)
lib = _lib.load()
sam, hard = bench.make_models(args)
rz, dn = sam_cyutils.dynamic_binary_number_at_fobs(fobs_cents / 2.0, sam, hard, cosmo, device=True)
edges = [sam.mtot, sam.mrat, sam.redz, fobs_edges / 2.0]
strain = gravwaves._char_strain_sq(edges, rz, params=True, dnum=dn)
number, h2fdf = strain["number"], strain["h2fdf"]


def dev_ms(fn, n=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(n):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def draw_ms(fn):
    best = 1e30
    for _ in range(3):
        lib.holo_set_profiling(1); fn(); lib.holo_set_profiling(0); torch.cuda.synchronize()
        prof = (C.c_double * 8)(); lib.holo_get_profile(prof, 8)
        best = min(best, prof[1])
    return best


Rs = [int(x) for x in sys.argv[1:]] or [1, 32, 100, 256, 500, 1000, 2000]
print("%6s %10s %12s %12s" % ("R", "gwb", "loud L=5", "loud+par L=5"))
for R in Rs:
    t_g = dev_ms(lambda: cyutils.sam_poisson_gwb(number, h2fdf, R, seed=1, device=True))
    t_l = draw_ms(lambda: single_sources.ss_gws_redz(edges, rz, number, realize=R, loudest=5, params=False, seed=1, device=True, _precomputed=strain))
    t_p = draw_ms(lambda: single_sources.ss_gws_redz(edges, rz, number, realize=R, loudest=5, params=True, seed=1, device=True, _precomputed=strain))
    print("%6d %10.3f %12.3f %12.3f" % (R, t_g, t_l, t_p), flush=True)
