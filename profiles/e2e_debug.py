import sys, time, argparse
sys.path.insert(0, '/root/repo')
import torch, numpy as np
import bench
from holodeck_b200 import utils, _lib, cyutils
from holodeck_b200.constants import YR
args = argparse.Namespace(shape=[91, 81, 101], nfreqs=40, realize=1000, loudest=1)
fobs_cents, fobs_edges = utils.pta_freqs(16.03*YR, 40)
def run(device):
    torch.cuda.synchronize(); t0=time.perf_counter()
    sam, hard = bench.make_models(args)
    torch.cuda.synchronize(); t1=time.perf_counter()
    out = sam.gwb(fobs_edges, hard, realize=1000, loudest=1, seed=12345, device=device)
    torch.cuda.synchronize(); t2=time.perf_counter()
    return round((t1-t0)*1e3,2), round((t2-t1)*1e3,2), dict(cyutils.STATS), _lib.load().holo_last_error()[:60]
for dev in [True, True, True, False, False, False, True, False, True, False]:
    print(dev, run(dev))
