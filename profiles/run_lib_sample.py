"""One librarian sample (BASELINE configs[4] inner call: run_model R=100, L=5, params + gwb) for use under ncu."""
import sys, argparse
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import bench
from holodeck_b200 import librarian
args = argparse.Namespace(shape=[91, 81, 101], nfreqs=40, realize=100, loudest=5)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    sam, hard = bench.make_models(args, scatter_dex=float(sys.argv[2]) if len(sys.argv) > 2 else 0.3)
    out = librarian.run_model(sam, hard, nreals=100, nloudest=5, params_flag=True, seed=1)
torch.cuda.synchronize()
print("done", sorted(out.keys()))
