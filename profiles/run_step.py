"""One bench step (BASELINE configs[1]) for use under ncu:
python profiles/run_step.py [nsteps] [realize] [loudest] [scatter_dex] [params 0|1]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import argparse
import torch
import bench

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1
args = argparse.Namespace(shape=[91, 81, 101], nfreqs=40, realize=int(sys.argv[2]) if len(sys.argv) > 2 else 1000,
                          loudest=int(sys.argv[3]) if len(sys.argv) > 3 else 1)
from holodeck_b200 import utils
from holodeck_b200.constants import YR
fobs_cents, fobs_edges = utils.pta_freqs(16.03*YR, args.nfreqs)
for _ in range(ns):
    sam, hard = bench.make_models(args, scatter_dex=float(sys.argv[4]) if len(sys.argv) > 4 else 0.0)
    out = sam.gwb(fobs_edges, hard, realize=args.realize, loudest=args.loudest, seed=12345, device=True,
                  params=bool(int(sys.argv[5])) if len(sys.argv) > 5 else False)
torch.cuda.synchronize()
print("done", out[1].shape)
