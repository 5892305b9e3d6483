"""Realised outputs of the draw kernels for fixed seeds, for bit-for-bit comparison between two builds of the library:
HOLO_B200_LIB=<lib.so> python profiles/dump_outputs.py out.npz ; python profiles/dump_outputs.py --compare a.npz b.npz"""
import sys, argparse
from pathlib import Path
import numpy as np
if sys.argv[1] == "--compare":
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    bad = [k for k in a.files if not np.array_equal(a[k], b[k], equal_nan=True)]
    print("identical" if not bad else "DIFFERENT: %s" % bad, "(%d arrays)" % len(a.files))
    sys.exit(1 if bad else 0)
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from holodeck_b200 import gravwaves, single_sources, cosmo, utils, cyutils
from holodeck_b200.sams import sam_cyutils
from holodeck_b200.constants import YR
args = argparse.Namespace(shape=[91, 81, 101], nfreqs=40, realize=1000, loudest=1)
fobs_cents, fobs_edges = utils.pta_freqs(16.03*YR, 40)
sam, hard = bench.make_models(args)
rz, dn = sam_cyutils.dynamic_binary_number_at_fobs(fobs_cents / 2.0, sam, hard, cosmo, device=True)
edges = [sam.mtot, sam.mrat, sam.redz, fobs_edges / 2.0]
strain = gravwaves._char_strain_sq(edges, rz, params=True, dnum=dn)
out = {}
for R in (100, 1000):
    out["gwb%d" % R] = cyutils.sam_poisson_gwb(strain["number"], strain["h2fdf"], R, seed=3)
    for params in (False, True):
        res = single_sources.ss_gws_redz(edges, rz, strain["number"], realize=R, loudest=5, params=params, seed=4, _precomputed=strain,
                                         _gwb=(R, 5) if R == 100 else None)
        for ii, rr in enumerate(res):
            out["ss%d_%d_%d" % (R, int(params), ii)] = rr
np.savez(sys.argv[1], **out)
print("wrote", sys.argv[1], len(out), "arrays")
