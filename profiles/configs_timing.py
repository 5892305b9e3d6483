"""Wall/device time of the other BASELINE configs (after warm-up): python profiles/configs_timing.py"""
import sys, time, argparse
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, numpy as np
import bench
import holodeck_b200 as holo
from holodeck_b200 import utils, cyutils, librarian
from holodeck_b200.constants import YR, PC

args = argparse.Namespace(shape=[91, 81, 101], nfreqs=40, realize=1000, loudest=1)
fobs_cents, fobs_edges = utils.pta_freqs(16.03*YR, 40)

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)

def gwb(L, params):
    sam, hard = bench.make_models(args)
    return sam.gwb(fobs_edges, hard, realize=1000, loudest=L, params=params, seed=1)

print("config1/2 sam.gwb L=1            %.1f ms" % timeit(lambda: gwb(1, False)))
print("config3  sam.gwb L=10           %.1f ms" % timeit(lambda: gwb(10, False)))
print("config3' sam.gwb L=10 params    %.1f ms" % timeit(lambda: gwb(10, True)))
print("retries", cyutils.STATS)

def lib_sample():
    sam, hard = bench.make_models(args)
    return librarian.run_model(sam, hard, nreals=100, nloudest=5, params_flag=True, seed=1)
print("config5  run_model R=100 L=5 params+gwb  %.1f ms/sample" % timeit(lambda: lib_sample()))
print("retries", cyutils.STATS)

sam0 = holo.sams.Semi_Analytic_Model(shape=30, mmbulge=holo.host_relations.MMBulge_KH2013(scatter_dex=0.0))
fc0, fe0 = utils.pta_freqs(10.0*YR, 20)
print("config0  default SAM(30) Hard_GW R=10  %.2f ms" % timeit(lambda: holo.sams.Semi_Analytic_Model(shape=30, mmbulge=holo.host_relations.MMBulge_KH2013(scatter_dex=0.0)).gwb(fe0, holo.hardening.Hard_GW(), realize=10, seed=1)))

sam, hard = bench.make_models(args)
sepa, ecc = holo.sams.evolve_eccen_uniform_single(sam, 0.9, 10*PC, 123)
print("config4  eccentric continuous 91x81x101, F=40, H=100  %.1f ms" % timeit(lambda: holo.gravwaves.sam_calc_gwb_single_eccen(fobs_cents, sam, sepa, ecc, nharms=100)))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = holo.gravwaves.sam_calc_gwb_single_eccen_discrete(fobs_cents, sam, sepa, ecc, nharms=100, nreals=500, seed=1)
    torch.cuda.synchronize()
    print("config4' eccentric discrete, F=40, H=100, R=500 (full grid)  %.1f ms" % ((time.perf_counter() - t0) * 1e3), tuple(out.shape))
cont = holo.gravwaves.sam_calc_gwb_single_eccen(fobs_cents, sam, sepa, ecc, nharms=100)
mean = out.mean(axis=2)
print("   mean over realizations / continuous (sum over harmonics, first 5 freqs):", (mean.sum(axis=1) / cont.sum(axis=1))[:5])
