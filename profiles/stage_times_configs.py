"""Per-stage device times (CUDA events) of BASELINE configs[2] with params and configs[4]'s inner call.
python profiles/stage_times_configs.py"""
import sys, argparse, time, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, numpy as np
import bench
from holodeck_b200 import _lib, gravwaves, single_sources, cosmo, utils
from holodeck_b200.sams import sam_cyutils
from holodeck_b200.constants import YR

args = argparse.Namespace(shape=[91, 81, 101], nfreqs=40, realize=1000, loudest=1)
fobs_cents, fobs_edges = utils.pta_freqs(16.03*YR, 40)
lib = _lib.load()


def ev():
    ee = torch.cuda.Event(enable_timing=True)
    ee.record()
    return ee


def one(R, L, params, gwb):
    best = {}
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        marks = [("start", ev())]
        sam, hard = bench.make_models(args)
        marks.append(("norm_2pwl", ev()))
        sam._static_binary_density_device()
        marks.append(("density", ev()))
        rz, dn = sam_cyutils.dynamic_binary_number_at_fobs(fobs_cents / 2.0, sam, hard, cosmo, device=True)
        marks.append(("dbn_2pwl", ev()))
        edges = [sam.mtot, sam.mrat, sam.redz, fobs_edges / 2.0]
        strain = gravwaves._char_strain_sq(edges, rz, params=params, dnum=dn)
        marks.append(("integrate_strain", ev()))
        lib.holo_set_profiling(1)
        out = single_sources.ss_gws_redz(edges, rz, strain["number"], realize=R, loudest=L, params=params, seed=1,
                                         _precomputed=strain)
        lib.holo_set_profiling(0)
        marks.append(("ss_gws_redz_total", ev()))
        torch.cuda.synchronize()
        prof = (C.c_double * 8)()
        nn = lib.holo_get_profile(prof, 8)
        if gwb:
            marks.append(("mark", ev()))
            g = gravwaves._gws_from_hc2(strain["h2fdf"], strain["number"], R, True, 2, 0, False)
            marks.append(("gwb_poisson", ev()))
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        cur = {marks[ii][0]: marks[ii - 1][1].elapsed_time(marks[ii][1]) for ii in range(1, len(marks))}
        cur.pop("mark", None)
        if nn >= 4:
            cur.update({"  head_prep": prof[0], "  draw": prof[1], "  resolve": prof[2], "  final": prof[3]})
        cur["wall_total"] = wall
        for kk, vv in cur.items():
            best[kk] = min(best.get(kk, 1e30), vv)
    return best


for name, kw in [("config2  R=1000 L=1", dict(R=1000, L=1, params=False, gwb=False)),
                 ("config3  R=1000 L=10 params", dict(R=1000, L=10, params=True, gwb=False)),
                 ("config5  R=100 L=5 params + gwb", dict(R=100, L=5, params=True, gwb=True))]:
    res = one(**kw)
    print(name)
    for kk, vv in res.items():
        print("   %-20s %8.3f ms" % (kk, vv))
