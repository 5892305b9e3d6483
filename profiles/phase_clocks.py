"""Where the realization kernel's CTAs spend their cycles (thread 0 of every CTA, clock64 between the phase
barriers).  Needs the profiling build: make -C holodeck_b200/csrc phase ; then
HOLO_B200_LIB=build/libholo_b200_phase.so python profiles/phase_clocks.py [R ...]"""
import os, sys, argparse, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("HOLO_B200_LIB", str(ROOT / "build" / "libholo_b200_phase.so"))
import torch, numpy as np
import bench
from holodeck_b200 import _lib, gravwaves, single_sources, cosmo, utils, cyutils
from holodeck_b200.sams import sam_cyutils
from holodeck_b200.constants import YR

args = argparse.Namespace(shape=[91, 81, 101], nfreqs=40, realize=1000, loudest=1)
fobs_cents, fobs_edges = utils.pta_freqs(16.03*YR, 40)
lib = _lib.load()
lib.holo_debug_phase_clocks.argtypes = [C.c_void_p, C.c_int]
sam, hard = bench.make_models(args)
rz, dn = sam_cyutils.dynamic_binary_number_at_fobs(fobs_cents / 2.0, sam, hard, cosmo, device=True)
edges = [sam.mtot, sam.mrat, sam.redz, fobs_edges / 2.0]
strain = gravwaves._char_strain_sq(edges, rz, params=True, dnum=dn)
number, h2fdf = strain["number"], strain["h2fdf"]
names = ["staging", "table builds", "phase A (tables)", "group + PTRS", "barrier wait", "flush (quad kernel)", "", ""]


def report(tag, fn):
    fn(); torch.cuda.synchronize()
    buf = (C.c_ulonglong * 8)()
    lib.holo_debug_phase_clocks(buf, 1)
    fn(); torch.cuda.synchronize()
    lib.holo_debug_phase_clocks(buf, 1)
    tot = float(sum(buf))
    print(tag, " ".join("%s %.1f%%" % (names[i], 100.0 * buf[i] / tot) for i in range(6)), " total CTA-cycles %.3e" % tot, flush=True)


for R in [int(x) for x in sys.argv[1:]] or [1, 100, 1000]:
    report("gwb      R=%4d:" % R, lambda: cyutils.sam_poisson_gwb(number, h2fdf, R, seed=1, device=True))
    report("loud L=5 R=%4d:" % R, lambda: single_sources.ss_gws_redz(edges, rz, number, realize=R, loudest=5, params=False, seed=1, device=True, _precomputed=strain))
    report("loud+par R=%4d:" % R, lambda: single_sources.ss_gws_redz(edges, rz, number, realize=R, loudest=5, params=True, seed=1, device=True, _precomputed=strain))
