import sys, argparse
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import bench
from holodeck_b200 import gravwaves, cosmo, utils
from holodeck_b200.sams import sam_cyutils
from holodeck_b200.constants import YR
args = argparse.Namespace(shape=[91, 81, 101], nfreqs=40, realize=1000, loudest=1)
fobs_cents, fobs_edges = utils.pta_freqs(16.03*YR, 40)
sam, hard = bench.make_models(args)
rz, dn = sam_cyutils.dynamic_binary_number_at_fobs(fobs_cents / 2.0, sam, hard, cosmo, device=True)
edges = [sam.mtot, sam.mrat, sam.redz, fobs_edges / 2.0]
strain = gravwaves._char_strain_sq(edges, rz, params=False, dnum=dn)
number, h2fdf = strain["number"], strain["h2fdf"]
mean_exp = (number * h2fdf).sum(dim=(0, 1, 2)).cpu().numpy()
var_exp = (number * h2fdf * h2fdf).sum(dim=(0, 1, 2)).cpu().numpy()
m3 = (number * h2fdf**3).sum(dim=(0, 1, 2)).cpu().numpy()
R, L = 256, 10
print("skewness of the mean over R=256 (Gaussian if << 1):", np.round(m3 / var_exp**1.5 / np.sqrt(R), 2)[:8], "...")
for seed in (77, 1, 2, 3, 4, 5, 6, 7, 8, 9):
    hc_ss, hc_bg = sam.gwb(fobs_edges, hard, realize=R, loudest=L, seed=seed)
    tot = hc_bg**2 + np.sum(hc_ss**2, axis=-1)
    zz = (tot.mean(axis=1) - mean_exp) / np.sqrt(var_exp / R)
    k = np.argmax(np.abs(zz))
    print("seed", seed, "max |z| %.2f at f index %d (z = %+.2f); rms z %.2f" % (abs(zz[k]), k, zz[k], np.sqrt(np.mean(zz**2))))
