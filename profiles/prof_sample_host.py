"""Host-side profile of library samples (what the Python driver spends per sample around the kernels):
OMP_NUM_THREADS=1 python profiles/prof_sample_host.py [nsamples]   (torchrun sets OMP_NUM_THREADS=1 for its ranks)"""
import cProfile, pstats, sys, time, io
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
import holodeck_b200 as holo
from holodeck_b200.librarian import lib_tools

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
space = holo.librarian.PS_Classic_Phenom_Uniform(nsamples=n + 2, seed=1)
def one(ii):
    sam, hard = space.model_for_params(space.param_dict(ii))
    return lib_tools.run_model(sam, hard, nreals=100, nloudest=5, params_flag=True, seed=ii, device=True)
one(0); one(1); torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for ii in range(2, n + 2):
    one(ii)
torch.cuda.synchronize()
pr.disable()
dt = time.perf_counter() - t0
print("%.2f ms per sample (wall, %d samples)" % (1e3 * dt / n, n))
ss = io.StringIO()
pstats.Stats(pr, stream=ss).sort_stats("tottime").print_stats(28)
print(ss.getvalue()[:6000])
