"""Start / end time (globaltimer) and SM of every CTA of one realization-kernel launch (profiling build):
HOLO_B200_LIB=build/libholo_b200_phase.so python profiles/cta_spans.py [R]"""
import os, sys, argparse, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
os.environ.setdefault("HOLO_B200_LIB", str(ROOT / "build" / "libholo_b200_phase.so"))
import torch, numpy as np
import bench
from holodeck_b200 import _lib, gravwaves, cosmo, utils, cyutils
from holodeck_b200.sams import sam_cyutils
from holodeck_b200.constants import YR

args = argparse.Namespace(shape=[91, 81, 101], nfreqs=40, realize=1000, loudest=1)
fobs_cents, fobs_edges = utils.pta_freqs(16.03*YR, 40)
lib = _lib.load()
lib.holo_debug_cta_spans.argtypes = [C.c_void_p]
sam, hard = bench.make_models(args)
rz, dn = sam_cyutils.dynamic_binary_number_at_fobs(fobs_cents / 2.0, sam, hard, cosmo, device=True)
edges = [sam.mtot, sam.mrat, sam.redz, fobs_edges / 2.0]
strain = gravwaves._char_strain_sq(edges, rz, params=False, dnum=dn)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
lib.holo_debug_item_phases.argtypes = [C.c_void_p, C.c_int]
LOUD = len(sys.argv) > 2 and sys.argv[2] == "loud"      # the loudest-split variant instead of the plain GWB
for _ in range(2):
    if LOUD:
        from holodeck_b200 import single_sources
        try:
            single_sources.ss_gws_redz(edges, rz, strain["number"], realize=R, loudest=5, seed=1, device=True, _precomputed=strain)
        except Exception as err:      # (the HOLO_NO_PUSH experiment build cannot fill the loudest slots)
            print("loudest call failed:", str(err)[:80])
    else:
        cyutils.sam_poisson_gwb(strain["number"], strain["h2fdf"], R, seed=1, device=True)
torch.cuda.synchronize()
ph = np.zeros((16384, 8), dtype=np.uint64)
lib.holo_debug_item_phases(ph.ctypes.data, 0)
np.save(ROOT / "gpurun_out" / ("item_phase_R%d%s.npy" % (R, "_loud" if LOUD else "")), ph[:5120, :5] / 2.0)   # (two launches)
buf = np.zeros((4, 16384), dtype=np.uint64)
lib.holo_debug_cta_spans(buf.ctypes.data)
n = 5120
t0, t1, sm = buf[0, :n].astype(np.int64), buf[1, :n].astype(np.int64), buf[2, :n].astype(int)
base = t0.min()
t0, t1 = (t0 - base) / 1e6, (t1 - base) / 1e6
dur = t1 - t0
print("kernel span %.3f ms; sum of CTA durations %.1f ms -> mean residency %.2f CTAs/SM" % (t1.max(), dur.sum(), dur.sum() / t1.max() / 148))
print("CTA duration ms: min %.3f median %.3f mean %.3f p90 %.3f max %.3f" % (dur.min(), np.median(dur), dur.mean(), np.quantile(dur, 0.9), dur.max()))
print("last CTA start %.3f ms; CTAs still running at 80/90/95%% of the span: %d %d %d" % (
    t0.max(), *[int(np.sum((t0 <= ff * t1.max()) & (t1 > ff * t1.max()))) for ff in (0.8, 0.9, 0.95)]))
lin = np.arange(n)
item = buf[3, :n].astype(int)          # (chunk, frequency group) work item each launch slot was handed
chunk = item // 10
late = np.argsort(-t1)[:10]
print("the 10 CTAs finishing last: (launch slot, chunk, fg, start, dur)", [(int(i), int(chunk[i]), int(item[i] % 10), round(float(t0[i]), 2), round(float(dur[i]), 2)) for i in late])
per_chunk = np.array([dur[chunk == c].mean() for c in range(512)])
print("mean CTA duration by 1/16 of the chunk index:", np.round(per_chunk.reshape(16, -1).mean(1), 3))
busy = np.array([dur[sm == s].sum() for s in range(sm.max() + 1)])
print("per-SM busy CTA-ms: min %.1f mean %.1f max %.1f (SMs: %d)" % (busy.min(), busy.mean(), busy.max(), len(busy)))

# list scheduling of the measured CTA durations on 2 x 148 slots, in different launch orders
import heapq
def makespan(order):
    slots = [0.0] * 296
    heapq.heapify(slots)
    end = 0.0
    for i in order:
        s = heapq.heappop(slots)
        heapq.heappush(slots, s + dur[i])
        end = max(end, s + dur[i])
    return end
fg = item % 10
print("simulated makespan [ms]: launch order %.2f | frequency-group-major %.2f | longest first %.2f | ideal %.2f | longest CTA %.2f" % (
    makespan(lin), makespan(np.lexsort((chunk, fg))), makespan(np.argsort(-dur)), dur.sum() / 296, dur.max()))
for split in (2, 4):
    d2 = np.repeat(dur / split, split)
    dur_save, dur = dur, d2
    print("  chunks / %d: frequency-group-major %.2f  longest first %.2f" % (split, makespan(np.lexsort((np.repeat(chunk, split), np.repeat(fg, split)))), makespan(np.argsort(-d2))))
    dur = dur_save
os.makedirs(ROOT / "gpurun_out", exist_ok=True)
by_item = np.zeros(n)
by_item[item] = dur
np.save(ROOT / "gpurun_out" / ("cta_dur_R%d%s.npy" % (R, "_loud" if LOUD else "")), by_item)
