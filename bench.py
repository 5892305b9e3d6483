#!/usr/bin/env python
"""bench.py -- headline benchmark: cell-realizations/s of `sam.gwb` on the (91x81x101x40, realize=1000) grid.

Contract (see the task statement):  ``python bench.py --gpus N --steps K --warmup W [--impl reference]``
prints ONE JSON line on rank 0.  For N>1 it is launched under torchrun (one rank per GPU).

Step (this arm)   one full pass of the hot path for BASELINE.json configs[1]: a fresh
                  ``Semi_Analytic_Model`` (PS_Classic defaults INCLUDING its 0.3 dex M-Mbulge scatter,
                  param_spaces_classic.py:41: density K0 -> scatter K6 -> stalled-bin zeroing) ->
                  ``Fixed_Time_2PL_SAM`` (K1a) -> ``sam.gwb(fobs_edges, hard, realize=R, loudest=L)`` (K1b, K2+K2b, rank
                  sort, K4).  `value` keeps every array on the device; `e2e` is the same call through the public
                  numpy API (host edge arrays in, hc_ss / hc_bg numpy out, PCIe copies inside the timing).
N > 1             realizations shard.  ``--scaling weak`` (default): every rank runs R realizations (global index
                  r0 = rank*R: the union is one R*N-realization run); ``--scaling strong``: the R realizations of the
                  named metric are split over the ranks.  The per-rank hc tables are gathered with ONE
                  all_gather_into_tensor (NCCL).  value = cells * (realizations of all ranks) / max-over-ranks time.
--impl reference  the reference's own CPU path (compiled reference Cython from oracle/_ref + the numpy/scipy glue of
                  oracle/, driven by oracle/chain.py), all host cores, on a bounded sample of the same workload.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "cell-realizations/s for sam.gwb (91x81x101x40, realize=1000)"
UNIT = "cell-realizations/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", type=int, nargs=3, default=[91, 81, 101], help="(debug) grid edges M Q Z")
    ap.add_argument("--nfreqs", type=int, default=40)
    ap.add_argument("--realize", type=int, default=1000)
    ap.add_argument("--loudest", type=int, default=1, help="sam.gwb default")
    ap.add_argument("--settle-steps", type=int, default=60, help="untimed steps after the W warm-up steps (clock/allocator settle)")
    ap.add_argument("--scatter-dex", type=float, default=0.3,
                    help="M-Mbulge scatter of the SAM: 0.3 is PS_Classic's own default (param_spaces_classic.py:41); 0 = off")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: R realizations PER RANK (default); strong: the R realizations are split over the ranks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-reals", type=int, default=2, help="realizations in the bounded CPU sample")
    return ap.parse_args()


def kernel_source_hash():
    """short hash of the draw kernel's sources: `profiles/ncu_metrics.json` records the hash it was captured at, so a
    capture of an older kernel version is flagged instead of being re-printed as if live"""
    import hashlib
    hh = hashlib.sha1()
    for name in ("holo_realize.cu", "holo_rng.cuh"):
        hh.update((ROOT / "holodeck_b200" / "csrc" / name).read_bytes())
    return hh.hexdigest()[:12]


def peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        dd = json.loads(path.read_text())
        return dict(hbm_gbs=float(dd["hbm_gbs"]), sm_max_mhz=float(dd.get("sm_max_mhz", 1965.0)), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


# ==================================================================================================
# clocks sampler (nvidia-smi, recipe of /opt/skills/guides/B200_PROFILING.md)
# ==================================================================================================

class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    # NVML polled every 10 ms from a CHILD process (a sampling thread in this process would fight the step loop for
    # the GIL and slow the very thing being timed); prints "unix_time,sm,max_sm,reasons" lines
    NVML_CHILD = (
        "import sys,time,pynvml as nv\n"
        "nv.nvmlInit(); hh=nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))\n"
        "bits={'hw_slowdown':nv.nvmlClocksThrottleReasonHwSlowdown,'hw_thermal_slowdown':nv.nvmlClocksThrottleReasonHwThermalSlowdown,"
        "'sw_thermal_slowdown':nv.nvmlClocksThrottleReasonSwThermalSlowdown,'sw_power_cap':nv.nvmlClocksThrottleReasonSwPowerCap}\n"
        "cmax=nv.nvmlDeviceGetMaxClockInfo(hh,nv.NVML_CLOCK_SM)\n"
        "print('ready',flush=True)\n"
        "while True:\n"
        "    rr=int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(hh))\n"
        "    print(time.time(),nv.nvmlDeviceGetClockInfo(hh,nv.NVML_CLOCK_SM),cmax,'|'.join(k for k,b in bits.items() if rr&b),sep=',',flush=True)\n"
        "    time.sleep(0.01)\n")

    def start(self):
        self.samples, self.nvml = [], False
        try:
            import pynvml  # noqa: F401
            self.proc = subprocess.Popen([sys.executable, "-c", self.NVML_CHILD, str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            if self.proc.stdout.readline().strip() == "ready":
                self.nvml = True      # its output waits in the pipe (a few KB) until stop(): no reader thread either
                return
            self.proc.kill()
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if getattr(self, "nvml", False):
            time.sleep(0.03)
            self.proc.terminate()
            for line in self.proc.stdout.read().splitlines():
                parts = line.strip().split(",")
                if len(parts) == 4:
                    self.samples.append((float(parts[0]), float(parts[1]), float(parts[2]), [rr for rr in parts[3].split("|") if rr]))
            # t0, t1 are perf_counter values: map the child's unix times onto them
            off = time.time() - time.perf_counter()
            inside = [ss for ss in self.samples if t0 <= ss[0] - off <= t1] or self.samples
            reasons = sorted({nn for ss in inside for nn in ss[3]})
            return {"sm_mhz": float(np.median([ss[1] for ss in inside])) if inside else None,
                    "sm_max_mhz": inside[0][2] if inside else None, "reasons": reasons, "samples": len(inside),
                    "source": "nvml polled every 10 ms by a child process; samples inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            parts = [pp.strip() for pp in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                clk, cmax = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            if t0 <= ts <= t1 + 0.1:
                sm.append(clk)
                smax.append(cmax)
                for nn, vv in zip(names, parts[5:9]):
                    if vv.lower().startswith("active"):
                        reasons.add(nn)
        if not sm:   # timed region shorter than the sampling period: use everything we saw
            for ts, line in self.lines:
                parts = [pp.strip() for pp in line.split(",")]
                try:
                    sm.append(float(parts[1]))
                    smax.append(float(parts[2]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(smax)) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ==================================================================================================
# reference arm / CPU baseline
# ==================================================================================================

_STATE = None   # reference grids, inherited copy-on-write by forked workers (never pickled)


def _cpu_worker(args):
    """One process = one independent run of the reference's realised stage on its own view of the grids."""
    nreals, loudest, seed = args
    from oracle import chain
    _, _, dt = chain.reference_realize(_STATE, nreals, loudest, seed=seed)
    return dt


def cpu_reference(args, nproc, nreals_each, state=None, tdet=None, scatter_sample=None):
    """Bounded sample of the reference CPU path.  Returns (cell_real_per_s, info dict, state, tdet).

    The M-Mbulge scatter (`add_scatter_to_masses`, scipy, once per SAM) is part of the workload when
    `--scatter-dex` > 0: the all-cores arm runs all 101 redshift slices once, spread over the processes (they are
    independent), and charges every one of the `nproc` side-by-side jobs the full core-seconds; the one-core
    `cpu_baseline` of the GPU arm times a bounded sample of slices (`scatter_sample`) and extrapolates linearly."""
    from oracle import chain
    wl = chain.classic_workload(shape=tuple(args.shape), nfreqs=args.nfreqs, scatter_dex=args.scatter_dex)
    if state is None:
        state, tdet = chain.reference_deterministic(wl, nproc=nproc, scatter_sample=scatter_sample)
        if nproc > 1 and "scatter" in tdet and scatter_sample is None:
            tdet["scatter"] = tdet["scatter"] * nproc          # wall on nproc processes -> core-seconds of one job
    global _STATE
    _STATE = state
    ncell = int(np.prod(state["number"].shape))
    t_det = float(sum(tdet.values()))
    if nproc <= 1:
        t0 = time.perf_counter()
        _cpu_worker((nreals_each, args.loudest, 1))
        wall = time.perf_counter() - t0
    else:
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        with ctx.Pool(nproc) as pool:
            pool.map(_cpu_worker, [(0, args.loudest, 0)] * nproc)     # spin the workers up (imports) untimed
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(nreals_each, args.loudest, 1 + ii) for ii in range(nproc)], chunksize=1)
            wall = time.perf_counter() - t0
    # the realised stage is exactly linear in R (cyutils.pyx:1318); nproc processes each did nreals_each
    # realizations concurrently in `wall` seconds -> seconds per realization per process = wall/nreals_each
    t_real = wall / nreals_each
    # a full job per process: deterministic stages once + R realizations; nproc jobs run side by side
    t_job = t_det + args.realize * t_real
    value = nproc * ncell * args.realize / t_job
    info = dict(t_deterministic_s=round(t_det, 3), t_per_realization_s=round(t_real, 4), stage_s={kk: round(vv, 3) for kk, vv in tdet.items()})
    if state.get("scatter_info") is not None:
        info["scatter_sample"] = state["scatter_info"]
    return value, info, state, tdet


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nproc = os.cpu_count() or 1
    state, tdet = None, None
    vals = []
    t_begin = time.perf_counter()
    for ii in range(args.warmup + args.steps):
        val, info, state, tdet = cpu_reference(args, nproc, 1, state, tdet)
        if ii >= args.warmup:
            vals.append(val)
    wall = time.perf_counter() - t_begin
    value = float(np.mean(vals))
    ncell = int(np.prod(state["number"].shape))
    sample = (f"{nproc} processes x 1 realization of loudest_hc_from_sorted (L={args.loudest}) on the full "
              f"{'x'.join(map(str, state['number'].shape))} grid per step, extrapolated linearly to R={args.realize}; "
              f"deterministic stages (density, M-Mbulge scatter {args.scatter_dex} dex over all slices, 2PL norm, dbn, integrate, "
              f"strain+argsort) timed once: {info['t_deterministic_s']} core-s per job")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * ncell * args.realize / value, "higher_is_better": True,   # ms: one step on all host cores
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": "reference", "sample": sample, **info},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(wall, 1),
    }
    emit(line)


def workload_config(args):
    M, Q, Z = args.shape
    sc = f"M-Mbulge scatter {args.scatter_dex} dex" if args.scatter_dex > 0 else "scatter off"
    return {"workload": f"sam.gwb PS_Classic ({sc}) + Fixed_Time_2PL_SAM(3 Gyr), grid {M}x{Q}x{Z}x{args.nfreqs}, "
                        f"realize={args.realize}, loudest={args.loudest}",
            "grid": [M, Q, Z, args.nfreqs], "realize": args.realize, "loudest": args.loudest,
            "mmb_scatter_dex": args.scatter_dex,
            "parallelism": f"realization-sharded x{args.gpus} ({getattr(args, 'scaling', 'weak')} scaling)",
            "cache": "inputs larger than L2 (2 x 238 MB grids re-streamed every step; no explicit flush)"}


# ==================================================================================================
# B200 arm
# ==================================================================================================

def make_models(args, scatter_dex=None):
    """Fresh SAM + hardening for configs[1] (librarian/param_spaces_classic.py:13-89); the M-Mbulge scatter is
    `args.scatter_dex` (PS_Classic's 0.3 dex by default) unless given."""
    if scatter_dex is None:
        scatter_dex = float(getattr(args, "scatter_dex", 0.0))
    import holodeck_b200 as holo
    from holodeck_b200 import sams, host_relations
    from holodeck_b200.constants import GYR, PC
    gsmf = sams.GSMF_Schechter(phi0=-2.77, phiz=-0.6, mchar0_log10=11.24, mcharz=0.11, alpha0=-1.21, alphaz=-0.03)
    gpf = sams.GPF_Power_Law(frac_norm_allq=0.025, malpha=0.0, qgamma=0.0, zbeta=1.0, max_frac=1.0)
    gmt = sams.GMT_Power_Law(time_norm=0.5*GYR, malpha=0.0, qgamma=-1.0, zbeta=-0.5)
    mmb = host_relations.MMBulge_KH2013(mamp_log10=8.69, mplaw=1.10, scatter_dex=scatter_dex)
    sam = sams.Semi_Analytic_Model(gsmf=gsmf, gpf=gpf, gmt=gmt, mmbulge=mmb, shape=tuple(args.shape))
    hard = holo.hardening.Fixed_Time_2PL_SAM(sam, 3.0*GYR, sepa_init=1e4*PC, rchar=100.0*PC, gamma_inner=-1.0,
                                             gamma_outer=+2.5)
    return sam, hard


def deterministic_parity(sam, hard, fobs_edges, state):
    """Parity of the very arrays a timed step produces (the product's own density, Brent roots, cosmology tables,
    K1b, K2+K2b) against the oracle chain's `state` (oracle/chain.reference_deterministic) on the same workload:
    per array the zero/sentinel-pattern mismatch count, the max relative error and the number of elements above
    1e-10.  Used by the `parity` key of the JSON line and by tests/test_gpu_fullsize.py."""
    from holodeck_b200 import _lib

    def cmp(got, want, tol=1e-10):
        got = np.asarray(got, dtype=float)
        sel = (want != 0) & np.isfinite(want)
        err = np.abs(got[sel] - want[sel]) / np.abs(want[sel])
        return {"n": int(want.size), "zero_mismatch": int(np.count_nonzero((got == 0) != (want == 0))),
                "max_rel": float(err.max()) if err.size else 0.0, f"n_above_{tol:g}": int(np.count_nonzero(err > tol))}

    edges, redz_final, strain = sam._number_and_strain(fobs_edges, hard, params=False)
    rep = {"dens": cmp(sam.static_binary_density, state["dens"])}
    dl = np.abs(np.log10(hard._norm) - state["norm_log10"])
    rep["norm_log10"] = {"n": int(dl.size), "max_abs": float(dl.max()), "n_above_1e-10": int(np.count_nonzero(dl > 1e-10))}
    rz = _lib.to_host(redz_final)
    rep["redz_final"] = cmp(rz, state["redz_final"])
    rep["redz_final"]["sentinel_mismatch"] = int(np.count_nonzero((rz == -1.0) != (state["redz_final"] == -1.0)))
    del rz
    number, h2fdf = _lib.to_host(strain["number"]), _lib.to_host(strain["h2fdf"])
    rep["number"] = cmp(number, state["number"])
    rep["h2fdf"] = cmp(h2fdf, state["h2fdf"])
    # the expectation-value spectrum sum(number * h2fdf) per frequency: what the realised spectra scatter about
    got = np.sum(number * h2fdf, axis=(0, 1, 2))
    want = np.sum(state["number"] * state["h2fdf"], axis=(0, 1, 2))
    rep["hc2_expect_max_rel"] = float(np.max(np.abs(got - want) / want))
    return rep


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import holodeck_b200 as holo   # noqa: F401
    from holodeck_b200 import _lib, utils
    from holodeck_b200.constants import YR
    lib = _lib.require_gpu()

    fobs_cents, fobs_edges = utils.pta_freqs(16.03*YR, args.nfreqs)
    L = args.loudest
    from holodeck_b200 import dist as hdist
    if args.scaling == "strong":        # the named R realizations are split over the ranks
        r0, R = hdist.realization_slice(args.realize, rank, world)
        Rtot = args.realize
        even = (args.realize % world == 0)
    else:                               # every rank draws R realizations of its own: one R*N-realization run
        R = args.realize
        r0 = rank * R
        Rtot = R * world
        even = True
    M, Q, Z = args.shape
    ncell = (M - 1) * (Q - 1) * (Z - 1) * args.nfreqs
    seed = 12345

    gbuf = {"buf": None, "pin": None}

    def gather(hc_ss, hc_bg):
        """ONE all_gather_into_tensor per step for both tables, into a buffer that lives across steps"""
        if not even:    # ragged strong-scaling split: per-table gather with padding
            return (hdist.gather_realizations(hc_ss, axis=1, nreals=Rtot), hdist.gather_realizations(hc_bg, axis=1, nreals=Rtot))
        (g_ss, g_bg), gbuf["buf"] = hdist.gather_tables([hc_ss, hc_bg], out=gbuf["buf"])
        return g_ss, g_bg

    def step_device():
        sam, hard = make_models(args)
        hc_ss, hc_bg = sam.gwb(fobs_edges, hard, realize=R, loudest=L, seed=seed, r0=r0, device=True)
        return gather(hc_ss, hc_bg)

    def step_e2e():
        sam, hard = make_models(args)
        if world == 1:
            return sam.gwb(fobs_edges, hard, realize=R, loudest=L, seed=seed, r0=r0)
        hc_ss, hc_bg = sam.gwb(fobs_edges, hard, realize=R, loudest=L, seed=seed, r0=r0, device=True)
        g_ss, g_bg = gather(hc_ss, hc_bg)
        if not even:
            return (_lib.to_host(g_ss), _lib.to_host(g_bg)) if rank == 0 else (g_ss, g_bg)
        if rank != 0:
            return g_ss, g_bg
        # the job's result is read on rank 0 only: one copy of the packed gather buffer into pinned memory
        buf = gbuf["buf"]
        if gbuf["pin"] is None or gbuf["pin"].shape != buf.shape:
            gbuf["pin"] = torch.empty(buf.shape, dtype=buf.dtype).pin_memory()
        gbuf["pin"].copy_(buf, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        _lib.TRAFFIC["d2h"] += buf.numel() * buf.element_size()
        full = gbuf["pin"].numpy().transpose(1, 0, 2)                       # (F, N*R, L + 1)
        return full[:, :, :L], full[:, :, L]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, nsteps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # a full (generation 2) pass of Python's cycle collector over the torch/numpy/scipy heap takes 30-60 ms --
        # several steps -- and lands wherever the allocation counter says: collect now, keep the collector off for
        # the K timed steps (device buffers are freed by reference counting, not by the collector)
        gc.collect()
        gc.disable()
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        marks = [ev0]
        for _ in range(nsteps):
            out = fn()
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
        ev1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        gc.enable()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # per-step spread on this rank (diagnostic only: the reported time is the bracket above)
        timed.last_steps = [marks[ii].elapsed_time(marks[ii + 1]) for ii in range(nsteps)]
        return float(ms.item()), out, t0, t1

    # the clock sampler (nvidia-smi) is started BEFORE the warm-up: its NVML start-up stalls kernel
    # submission for tens of ms and must not land in the timed region
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("HOLO_BENCH_NO_SAMPLER"):
        sampler.start()
    # Warm-up runs through the SAME bracket as the timed region (same events, same lifetime of the previous step's
    # outputs, hence the same caching-allocator pattern): on a fresh box the first process otherwise paid a one-off
    # 35-100 ms stall in the 4th timed step (a first-time segment allocation while the NVML child polls the driver).
    # After the W requested steps a fixed number of further untimed steps (the same on every rank: each step ends in
    # a collective), about a second of device work, lets clocks and allocator settle.
    timed(step_device, max(1, args.warmup))
    extra_warmup = max(0, args.settle_steps)
    if extra_warmup:
        timed(step_device, extra_warmup)
    n_launch0 = lib.holo_launch_count()
    ms_total, out, t0, t1 = timed(step_device, args.steps)
    slowest = int(np.argmax(timed.last_steps))
    if os.environ.get("HOLO_BENCH_DUMP_STEPS"):
        print("step_ms", [round(xx, 2) for xx in timed.last_steps], file=sys.stderr)
    steps_ms = sorted(timed.last_steps)
    n_launch = lib.holo_launch_count() - n_launch0
    value = ncell * Rtot * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public numpy API
    timed(step_e2e, max(2, min(args.warmup, 4)))
    _lib.TRAFFIC["h2d"] = _lib.TRAFFIC["d2h"] = 0
    ms_e2e, out_e2e, _, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    h2d = _lib.TRAFFIC["h2d"] // args.steps
    d2h = _lib.TRAFFIC["d2h"] // args.steps
    e2e_value = ncell * Rtot * args.steps / (ms_e2e * 1e-3)
    assert out_e2e[0].shape == (args.nfreqs, Rtot, L) and out_e2e[1].shape == (args.nfreqs, Rtot)
    if rank == 0:
        assert np.all(np.isfinite(out_e2e[1])) and np.all(out_e2e[1] > 0)

    # ---- the same step WITHOUT the M-Mbulge scatter (round 1's headline configuration; K6 skipped), and with
    #      loudest = 10 (BASELINE configs[2]: the single-source split), both beside the headline
    def step_noscatter():
        sam, hard = make_models(args, scatter_dex=0.0)
        hc_ss, hc_bg = sam.gwb(fobs_edges, hard, realize=R, loudest=L, seed=seed, r0=r0, device=True)
        return gather(hc_ss, hc_bg)
    timed(step_noscatter, 4)
    ms_noscatter, _, _, _ = timed(step_noscatter, args.steps)
    gbuf["buf"] = None

    def step_loud10():
        sam, hard = make_models(args)
        hc_ss, hc_bg = sam.gwb(fobs_edges, hard, realize=R, loudest=10, seed=seed, r0=r0, device=True)
        return gather(hc_ss, hc_bg)
    timed(step_loud10, 4)
    ms_loud10, _, _, _ = timed(step_loud10, args.steps)
    gbuf["buf"] = None

    def step_loud10_params():
        sam, hard = make_models(args)
        return sam.gwb(fobs_edges, hard, realize=R, loudest=10, params=True, seed=seed, r0=r0, device=True)
    timed(step_loud10_params, 3)
    ms_loud10p, _, _, _ = timed(step_loud10_params, args.steps)

    # ---- one librarian sample (BASELINE configs[4] inner call, lib_tools.run_model: R=100, 5 loudest, parameters
    #      and an independently drawn GWB -- both from one fused pass of the realization kernel), through the numpy API
    def step_library_sample():
        from holodeck_b200 import librarian
        sam, hard = make_models(args)
        return librarian.run_model(sam, hard, nreals=100, nloudest=5, params_flag=True, seed=seed + rank)
    timed(step_library_sample, 4)
    ms_lib, _, _, _ = timed(step_library_sample, args.steps)

    # ---- per-stage device times (CUDA events on the launching stream), one extra profiled pass
    stages = stage_times(args, fobs_edges, R, L, seed, r0)
    shim = shim_boundary_times(args, fobs_edges, R, L, seed) if (rank == 0 and world == 1) else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    nedge = M * Q * Z * args.nfreqs
    alg_bytes = {   # algorithmic bytes per launch (DESIGN.md "Kernels and rooflines")
        "density": 3 * M * Q * Z * 8,
        "norm_2pwl": 3 * M * Q * 8,
        "dbn_2pwl": 2 * nedge * 8 + 2 * M * Q * Z * 8,
        "integrate_strain": 2 * nedge * 8 + 2 * ncell * 8,
        "loudest_draw": 2 * ncell * 8,
    }
    dom = max((kk for kk in stages if kk in alg_bytes), key=lambda kk: stages[kk])
    ach = alg_bytes[dom] / (stages[dom] * 1e-3) / 1e9
    # per-launch DRAM traffic / instruction counts of the same kernels from the committed `ncu --set full` captures
    ncu = {}
    try:
        ncu = json.loads((ROOT / "profiles" / "ncu_metrics.json").read_text())
    except Exception:
        pass
    hbm_view = {"achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                "algorithmic_bytes": alg_bytes[dom]}
    clk_mhz = (clocks or {}).get("sm_mhz") or pk["sm_max_mhz"]
    cap = ncu.get(dom, {})
    stale = cap.get("source_hash") not in (None, kernel_source_hash())
    if dom == "loudest_draw" and cap.get("warp_inst_per_launch"):
        # The dominant kernel draws R Poisson counts per grid element out of shared memory (0.016 B of HBM per
        # cell-realization): it is bound by INSTRUCTION ISSUE.  achieved = warp instructions of one launch (ncu
        # `smsp__inst_executed.sum` of the same kernel, profiles/) / its live CUDA-event duration in this run;
        # peak = 148 SMs x 4 schedulers x 1 warp instruction per clock at the SM clock sampled during the run.
        wi = float(cap["warp_inst_per_launch"])
        ach_i = wi / (stages[dom] * 1e-3) / 1e9
        peak_i = 148 * 4 * clk_mhz * 1e6 / 1e9
        roofline = {"kernel": dom, "bound": "issue", "achieved": ach_i, "peak": peak_i, "unit": "Gwarp-inst/s",
                    "frac": ach_i / peak_i, "traffic": cap.get("dram_bytes_per_launch"), "ms": stages[dom],
                    "warp_inst_per_launch": wi, "inst_source": cap.get("source"), "inst_capture_stale": bool(stale),
                    "peak_source": f"148 SM x 4 issue slots x {clk_mhz:.0f} MHz (sampled)", "hbm": hbm_view,
                    "note": ("issue-bound kernel: `frac` is issue-slot use (ncu sm issue-active agrees: "
                             f"{cap.get('issue_active_pct')} %); `hbm` is the same launch against the measured HBM peak; "
                             "`traffic` = dram read+write bytes per launch from the same ncu capture")}
    else:
        roofline = {"kernel": dom, "bound": "hbm", **hbm_view, "traffic": cap.get("dram_bytes_per_launch"),
                    "peak_source": pk["source"], "ms": stages[dom]}
    per_kernel = {}
    for kk, bb in alg_bytes.items():
        if kk in stages and stages[kk] > 0:
            gbs = bb / (stages[kk] * 1e-3) / 1e9
            per_kernel[kk] = {"ms": round(stages[kk], 4), "GB/s": round(gbs, 1), "frac_hbm": round(gbs / pk["hbm_gbs"], 4)}
            if kk in ncu:
                per_kernel[kk]["ncu"] = {"dram_bytes_per_launch": ncu[kk]["dram_bytes_per_launch"],
                                         "warp_inst_per_launch": ncu[kk]["warp_inst_per_launch"],
                                         "issue_active_pct": ncu[kk]["issue_active_pct"], "source": ncu[kk]["source"]}
    # the draw kernel is instruction-issue bound, not HBM bound: nominal 25 thread-instructions per
    # cell-realization against 148 SM x 4 schedulers x 32 lanes x clock
    clk = (clocks or {}).get("sm_mhz") or pk["sm_max_mhz"]
    issue_peak = 148 * 4 * 32 * clk * 1e6
    if "loudest_draw" in stages:
        rate = ncell * R / (stages["loudest_draw"] * 1e-3)
        per_kernel["loudest_draw"].update({"cell_real_per_s": rate, "issue_frac_nominal25": round(rate * 25 / issue_peak, 4)})
        if "loudest_draw" in ncu:
            # measured: warp instructions of one launch (ncu) / live duration, against 148 SM x 4 issue slots x clock
            wi = ncu["loudest_draw"]["warp_inst_per_launch"]
            per_kernel["loudest_draw"]["issue_frac_measured"] = round(
                wi / (stages["loudest_draw"] * 1e-3) / (148 * 4 * clk * 1e6), 4)
            per_kernel["loudest_draw"]["thread_inst_per_cell_realization"] = round(wi * 32 / (ncell * R), 2)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "clocks": clocks,
        "step_ms_spread": {"min": round(steps_ms[0], 3), "median": round(steps_ms[len(steps_ms) // 2], 3),
                           "max": round(steps_ms[-1], 3), "slowest_step": slowest, "extra_warmup_steps": extra_warmup,
                           "note": "per-step CUDA-event times of the timed region on rank 0 (diagnostic; `ms_per_step` is the "
                                   "bracket over all K steps); extra untimed warm-up steps run after the W requested ones"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(n_launch),
        "loudest_retries": int(__import__("holodeck_b200").cyutils.STATS["loudest_retries"]),
        "mmbulge_scatter_off": {"ms_per_step": ms_noscatter / args.steps, "value": ncell * Rtot * args.steps / (ms_noscatter * 1e-3),
                                "note": "same step with mmb_scatter_dex=0 (round 1's headline configuration: K6 skipped)"},
        "config3_loudest10": {"ms_per_step": ms_loud10 / args.steps, "value": ncell * Rtot * args.steps / (ms_loud10 * 1e-3),
                              "note": "BASELINE configs[2]: same step with loudest=10 (ss_gws_redz nloudest=10)"},
        "config3_loudest10_params": {"ms_per_step": ms_loud10p / args.steps, "value": ncell * Rtot * args.steps / (ms_loud10p * 1e-3),
                                     "note": "same with params=True (sspar, bgpar): the variant run_model uses; no gather"},
        "library_sample": {"ms_per_sample": ms_lib / args.steps, "samples_per_s": world * args.steps / (ms_lib * 1e-3),
                           "note": "librarian.run_model on the same grid: nreals=100, nloudest=5, params + gwb (one sample "
                                   "per rank at a time; BASELINE configs[4] shards 2000 such samples over the ranks)"},
        "roofline": roofline,
        "e2e_shim": shim,
        "stages_ms": {kk: round(vv, 4) for kk, vv in stages.items()},
        "kernels": per_kernel,
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            val, info, _, _ = cpu_reference(args, 1, args.cpu_reals, scatter_sample=11 if args.scatter_dex > 0 else None)
            line["cpu_baseline"] = {
                "value": val, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": (f"reference chain on the same workload on ONE core: deterministic stages once (M-Mbulge scatter: 11 of "
                           f"{Z} redshift slices timed, linear in the slice count; the later stages run on the unscattered grid) + "
                           f"{args.cpu_reals} realizations of loudest_hc_from_sorted (compiled reference, oracle/_ref), "
                           f"extrapolated linearly to R={R}"),
                **info}
            # parity of the very arrays the timed step produces (scatter on) against the oracle chain; the oracle's
            # scatter (scipy, 101 slices) is spread over the host cores here -- it is the checker, not the baseline
            from oracle import chain
            wl = chain.classic_workload(shape=tuple(args.shape), nfreqs=args.nfreqs, scatter_dex=args.scatter_dex)
            st_full, _ = chain.reference_deterministic(wl, nproc=min(os.cpu_count() or 1, 16))
            sam, hard = make_models(args)
            par = deterministic_parity(sam, hard, fobs_edges, st_full)
            line["parity"] = {"against": "oracle/chain.reference_deterministic on the timed workload (compiled reference + numpy glue)",
                              **par}
            worst = max(par[kk]["max_rel"] for kk in ("dens", "number", "h2fdf"))
            assert par["number"]["zero_mismatch"] == 0 and par["redz_final"]["sentinel_mismatch"] == 0 and worst < 1e-6 \
                and par["hc2_expect_max_rel"] < 1e-9, f"deterministic parity lost: {par}"
        except Exception as err:   # the oracle did not travel: say so instead of inventing a number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {err}"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def shim_boundary_times(args, fobs_edges, R, L, seed, reps=3):
    """The drop-in boundary SURVEY section 8(b) defines, timed as stock holodeck would use it after the
    `sys.modules` aliasing of INTEGRATION.md section 3: the three native calls of `sam.gwb` made through the shim
    modules with HOST numpy arrays in and out -- every call uploads its inputs and reads its full-size outputs back
    (2 x 238 MB out of `dynamic_binary_number_at_fobs`, 238 MB in / 230 MB out of `integrate_differential_number_3dx1d`,
    2 x 230 MB into `loudest_hc_from_sorted`), pinned read-backs, wall clock around each call.  The Python glue between
    the calls (strain, argsort: numpy in the reference) is NOT timed here; it is prepared on the device, untimed."""
    import torch
    from holodeck_b200 import _lib, gravwaves, cosmo, utils, cyutils
    from holodeck_b200.sams import sam_cyutils
    sam, hard = make_models(args)
    sam._static_binary_density_device()
    fobs_cents = utils.midpoints(fobs_edges)
    edges = [sam.mtot, sam.mrat, sam.redz, fobs_edges / 2.0]
    best = {}
    traffic = {}

    def timed_call(name, cur, moved, fn):
        h0, d0 = _lib.TRAFFIC["h2d"], _lib.TRAFFIC["d2h"]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        cur[name] = time.perf_counter() - t0
        moved["h2d_bytes"] += _lib.TRAFFIC["h2d"] - h0
        moved["d2h_bytes"] += _lib.TRAFFIC["d2h"] - d0
        return out

    for rep in range(reps + 1):
        cur, moved = {}, {"h2d_bytes": 0, "d2h_bytes": 0}
        redz_final, diff_num = timed_call("dynamic_binary_number_at_fobs", cur, moved, lambda: sam_cyutils.dynamic_binary_number_at_fobs(
            fobs_cents / 2.0, sam, hard, cosmo))
        number = timed_call("integrate_differential_number_3dx1d", cur, moved,
                            lambda: sam_cyutils.integrate_differential_number_3dx1d(edges, diff_num))
        assert isinstance(redz_final, np.ndarray) and isinstance(number, np.ndarray)
        # glue (untimed): strain and rank order, as host arrays -- what the reference's numpy code hands to cyutils
        strain = gravwaves._char_strain_sq(edges, _lib.to_dev(redz_final), params=False)
        h2fdf = _lib.to_host(strain["h2fdf"])
        order = torch.sort(-strain["h2fdf"][..., 0].reshape(-1), stable=True).indices.cpu().numpy()
        msort, qsort, zsort = np.unravel_index(order, number.shape[:3])
        del strain
        hc2ss, hc2bg = timed_call("loudest_hc_from_sorted", cur, moved,
                                  lambda: cyutils.loudest_hc_from_sorted(number, h2fdf, R, L, msort, qsort, zsort, seed=seed))
        assert isinstance(hc2ss, np.ndarray) and hc2ss.shape == (args.nfreqs, R, L) and np.all(hc2bg > 0)
        del redz_final, diff_num, number, h2fdf
        if rep == 0:
            continue                         # first pass: pinned blocks are allocated, caches warm up
        traffic = {kk: int(vv) for kk, vv in moved.items()}
        for kk, vv in cur.items():
            best[kk] = min(best.get(kk, 1e30), vv)
    total = sum(best.values())
    M, Q, Z = args.shape
    ncell = (M - 1) * (Q - 1) * (Z - 1) * args.nfreqs
    return {"ms": {kk: round(1e3 * vv, 3) for kk, vv in best.items()}, "ms_total": round(1e3 * total, 3),
            "value": ncell * R / total, "unit": UNIT, **traffic,
            "note": "the three native calls of sam.gwb through the shim modules, numpy in / numpy out at full size "
                    "(INTEGRATION.md section 3); wall clock per call, best of %d" % reps}


def stage_times(args, fobs_edges, R, L, seed, r0):
    """Device time of each stage of one step, CUDA events on the launching (torch current) stream."""
    import ctypes as C
    import torch
    from holodeck_b200 import _lib, gravwaves, single_sources, cosmo, utils
    from holodeck_b200.sams import sam_cyutils
    lib = _lib.load()
    res = {}

    def ev():
        ee = torch.cuda.Event(enable_timing=True)
        ee.record()
        return ee

    best = {}
    for rep in range(3):
        torch.cuda.synchronize()
        marks = [("start", ev())]
        sam, hard = make_models(args)     # K1a runs in the hardening constructor; density is lazy
        marks.append(("norm_2pwl", ev()))
        if getattr(args, "scatter_dex", 0.0) > 0.0:
            # K0 alone on a scatter-free twin, then K0 + K6 (+ stalled-bin zeroing) on the timed model
            twin, _ = make_models(args, scatter_dex=0.0)
            marks.append(("norm_2pwl_twin", ev()))
            twin._static_binary_density_device()
            marks.append(("density", ev()))
            sam._static_binary_density_device()
            marks.append(("density_and_scatter", ev()))
        else:
            sam._static_binary_density_device()
            marks.append(("density", ev()))
        fobs_gw_cents = utils.midpoints(fobs_edges)
        redz_final, diff_num = sam_cyutils.dynamic_binary_number_at_fobs(fobs_gw_cents / 2.0, sam, hard, cosmo, device=True)
        marks.append(("dbn_2pwl", ev()))
        edges = [sam.mtot, sam.mrat, sam.redz, fobs_edges / 2.0]
        strain = gravwaves._char_strain_sq(edges, redz_final, params=False, dnum=diff_num)
        marks.append(("integrate_strain", ev()))
        lib.holo_set_profiling(1)
        single_sources.ss_gws_redz(edges, redz_final, strain["number"], realize=R, loudest=L, seed=seed, r0=r0,
                                   device=True, _precomputed=strain)
        lib.holo_set_profiling(0)
        marks.append(("ss_gws_redz_total", ev()))
        torch.cuda.synchronize()
        prof = (C.c_double * 8)()
        nn = lib.holo_get_profile(prof, 8)
        cur = {marks[ii][0]: marks[ii - 1][1].elapsed_time(marks[ii][1]) for ii in range(1, len(marks))}
        cur.pop("norm_2pwl_twin", None)
        if "density_and_scatter" in cur:
            cur["mmbulge_scatter"] = cur["density_and_scatter"] - cur["density"]
        if nn >= 4:
            cur.update({"loudest_head_prep": prof[0], "loudest_draw": prof[1], "loudest_resolve": prof[2], "loudest_final": prof[3]})
            cur["rank_sort_and_glue"] = cur["ss_gws_redz_total"] - sum(prof[ii] for ii in range(4))
        for kk, vv in cur.items():
            best[kk] = min(best.get(kk, 1e30), vv)
        del sam, hard, redz_final, diff_num, strain, marks
    res.update(best)
    return res


_JSON_OUT = None


def claim_stdout():
    """Keep the real stdout for the ONE JSON line: every other writer to fd 1 -- NCCL's version banner (printed by
    the C library under torchrun, NCCL_DEBUG_FILE notwithstanding), child processes -- lands on stderr instead."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
