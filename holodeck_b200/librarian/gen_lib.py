"""GPU-sharded library generation (``holodeck/librarian/gen_lib.py:50-337`` with ``torch.distributed``
instead of mpi4py).

Run under torchrun, one rank per GPU::

    torchrun --nnodes=1 --nproc-per-node 8 -m holodeck_b200.librarian.gen_lib PS_Classic_Phenom_Uniform OUT -n 2000 -r 100

As in the reference, rank 0 decides everything that must be common -- the seed of the Latin hypercube and of the
sample permutation (drawn on rank 0 when ``--seed`` is not given, then broadcast; the reference broadcasts the
``space`` and scatters the index lists, ``gen_lib.py:99-169``) -- every rank then builds the identical parameter
space and takes its share of the permuted sample indices.  Per sample a rank runs ``run_model`` and hands the
device-resident products to the streaming file plane (``librarian/stream.py``): rows of the combined layout are
written by a background thread, off the critical path.  ``--sim-files`` additionally writes the reference's
``library_sims/library__pNNNNNN.npz`` per sample.  A failing sample is recorded (NaN rows / a ``fail`` file) and
re-attempted on the next run (``gen_lib.py:287-319``); more than ``MAX_FAILURES`` failures abort the rank.  After the
final barrier rank 0 combines the library into the reference's single file (``gen_lib.py:228-231``).
"""
import argparse
import concurrent.futures
import json
import queue
import threading
from datetime import datetime
from pathlib import Path

import numpy as np

import holodeck_b200 as holo
from holodeck_b200 import dist
from holodeck_b200.librarian import (
    DEF_NUM_REALS, DEF_NUM_FBINS, DEF_NUM_LOUDEST, DEF_PTA_DUR, DIRNAME_LIBRARY_SIMS, ARGS_CONFIG_FNAME, lib_tools,
    stream,
)

MAX_FAILURES = 5

#: run settings that a resumed run must share with the files already on disk (``gen_lib.py:109-121`` reloads them)
_CONFIG_KEYS = ("param_space", "nsamples", "nreals", "nfreqs", "nloudest", "pta_dur", "sam_shape", "gwb_flag", "ss_flag",
                "params_flag", "seed")


def _sample_seed(seed, pnum):
    return None if seed is None else (int(seed) * 1000003 + int(pnum))


def _save_npz(args, space, pnum, params, data):
    """The reference's per-sample file (``gen_lib.py:321-323``)."""
    out = {kk: np.asarray(vv) for kk, vv in data.items()}
    out['params'] = np.array([params[pn] for pn in space.param_names])
    out['param_names'] = space.param_names
    np.savez(lib_tools._get_sim_fname(args.output_sims, pnum), **out)


def run_sam_at_pspace_params(args, space, pnum, params, writer=None):
    """Run sample ``pnum`` of ``space``; returns ``(ok, sim_fname)`` (``gen_lib.py:236-337``).

    With a ``writer`` (:class:`stream.AsyncSampleWriter`) the products go to the streaming store; without one the
    reference's per-sample ``.npz`` file is written synchronously."""
    log = args.log
    sim_fname = lib_tools._get_sim_fname(args.output_sims, pnum)
    if writer is not None:
        if writer.store.is_done(pnum) and not args.recreate:
            return True, sim_fname
    elif sim_fname.exists():
        temp = np.load(sim_fname)
        if 'fail' in list(temp.keys()):
            log.info("Existing file was a failure, re-attempting...")
        elif not args.recreate:
            return True, sim_fname
    try:
        data = _run_model(args, space, pnum, params, device=writer is not None)
        rv = True
    except Exception as err:   # noqa: BLE001  (same catch-all as the reference)
        log.exception(f"`run_model` FAILED on {pnum=}\n")
        log.exception(err)
        rv = False
        data = dict(fail=str(err))
    return _record(args, space, pnum, params, rv, data, writer, sim_fname)


def _run_model(args, space, pnum, params, device, deferred=None):
    sam, hard = space.model_for_params(params)
    return lib_tools.run_model(
        sam, hard,
        pta_dur=args.pta_dur, nfreqs=args.nfreqs, nreals=args.nreals, nloudest=args.nloudest,
        gwb_flag=args.gwb_flag, singles_flag=args.ss_flag, details_flag=False, params_flag=args.params_flag,
        log=args.log, seed=_sample_seed(args.seed, pnum), device=device, deferred=deferred,
    )


def _record(args, space, pnum, params, rv, data, writer, sim_fname):
    """hand a finished (or failed) sample to the file plane"""
    if writer is not None:
        if rv:
            writer.submit(pnum, data)
        else:
            writer.submit_failure(pnum, data['fail'])
            if args.sim_files:
                np.savez(sim_fname, **data)
        return rv, sim_fname
    if rv:
        _save_npz(args, space, pnum, params, data)
    else:
        np.savez(sim_fname, **data)
    return rv, sim_fname


def _check_config(output, config, resume_ok=True):
    """Save the run configuration next to the parameter space, or -- when files of an earlier run are there --
    demand that it is the same (a resumed run must not mix nreals / nfreqs / flags)."""
    fname = Path(output).joinpath(ARGS_CONFIG_FNAME)
    if fname.exists() and resume_ok:
        old = json.loads(fname.read_text())
        diff = {kk: (old.get(kk), config.get(kk)) for kk in _CONFIG_KEYS if old.get(kk) != config.get(kk)}
        if diff:
            raise RuntimeError(f"{fname} was written by a run with different settings {diff}: "
                               "use a new output directory, or `--recreate` to start over")
        return fname
    fname.write_text(json.dumps({**config, "created": str(datetime.now()), "holodeck_b200": holo.__version__}, indent=1))
    return fname


def run_library(space, output, nreals=DEF_NUM_REALS, nfreqs=DEF_NUM_FBINS, nloudest=DEF_NUM_LOUDEST,
                pta_dur=DEF_PTA_DUR, gwb_flag=True, ss_flag=True, params_flag=False, recreate=False, seed=None,
                log=None, indices=None, streaming=True, sim_files=False, param_space_name=None, workers=1, pipeline=True):
    """Generate this rank's share of the library; returns ``(num_done, failures)``.

    ``seed`` must be the same on every rank (``main`` broadcasts it): it fixes the sample permutation.
    ``workers``: host threads per rank, each driving its own CUDA stream.  A sample is ~10 ms of kernels plus a few
    ms of host work (model construction, launches, the overflow check of the loudest split, which synchronises);
    with two samples in flight the host work of one hides behind the kernels of the other and the kernels of the
    two streams fill each other's tails.  Results do not depend on it (every sample has its own seed).  Measured
    on B200: no gain at one rank and a large loss at eight (the threads fight over the interpreter lock), hence the
    default of one; what hides the host work instead is ``pipeline`` (streaming mode only): a sample's kernels are
    enqueued without waiting for its overflow / sanity flags, which are looked at after the NEXT sample has been
    enqueued (``single_sources.DeferredChecks``)."""
    from holodeck_b200 import utils
    from holodeck_b200.constants import YR
    rank, size = dist.world()
    log = holo.log if log is None else log
    output = Path(output)
    output_sims = output.joinpath(DIRNAME_LIBRARY_SIMS)
    if size > 1 and seed is None and indices is None:
        raise ValueError("run_library on several ranks needs a common `seed` (see `dist.shared_seed`)")
    config = dict(param_space=param_space_name or space.name, nsamples=int(space.nsamples), nreals=int(nreals),
                  nfreqs=int(nfreqs), nloudest=int(nloudest), pta_dur=float(pta_dur), sam_shape=space.sam_shape,
                  gwb_flag=bool(gwb_flag), ss_flag=bool(ss_flag), params_flag=bool(params_flag),
                  seed=None if seed is None else int(seed))
    if rank == 0:
        output_sims.mkdir(parents=True, exist_ok=True)
        _check_config(output, config, resume_ok=not recreate)
        space.save(output)
        if streaming:
            fobs_cents, fobs_edges = utils.pta_freqs(dur=pta_dur*YR, num=nfreqs)
            if recreate and stream.LibraryStore.exists(output):
                import shutil
                shutil.rmtree(stream.LibraryStore._dir(output))       # start over, possibly with another layout
            stream.LibraryStore.create(output, space.nsamples, nfreqs, nreals, nloudest, gwb_flag, ss_flag,
                                       params_flag, fobs_cents, fobs_edges)
    dist.barrier()
    args = argparse.Namespace(
        log=log, output=output, output_sims=output_sims, recreate=recreate, pta_dur=pta_dur, nfreqs=nfreqs,
        nreals=nreals, nloudest=nloudest, gwb_flag=gwb_flag, ss_flag=ss_flag, params_flag=params_flag, seed=seed,
        sim_files=sim_files)
    if indices is None:
        indices = dist.sample_indices(space.nsamples, seed=seed, rank=rank, size=size)
    writer = None
    if streaming:
        store = stream.LibraryStore.open(output)
        npz = None
        if sim_files:
            def npz(pnum, host):
                full = dict(fobs_cents=store.fobs[0], fobs_edges=store.fobs[1], **host)
                _save_npz(args, space, pnum, space.param_dict(int(pnum)), full)
        writer = stream.AsyncSampleWriter(store, also_npz=npz)
    beg = datetime.now()
    state = dict(failures=0, done=0)
    lock = threading.Lock()
    todo = queue.SimpleQueue()
    for sim_num in indices:
        todo.put(int(sim_num))

    def count(rv):
        with lock:
            if rv is False:
                state["failures"] += 1
            state["done"] += 1
            if (MAX_FAILURES is not None) and (state["failures"] > MAX_FAILURES):
                err = f"Failed {state['failures']} times on rank:{rank}!"
                log.exception(err)
                raise RuntimeError(err)

    def finalize(pending):
        """second half of a pipelined sample: look at its device flags (this is where the host waits for the GPU),
        redo it synchronously in the rare overflow case, hand it to the writer"""
        sim_num, params, data, checks, err = pending
        rv = err is None
        if rv:
            try:
                if checks.overflow():
                    data = _run_model(args, space, sim_num, params, device=True)      # larger head, synchronous
                else:
                    checks.verify()
            except Exception as ee:   # noqa: BLE001
                err, rv = ee, False
        if not rv:
            log.exception(f"`run_model` FAILED on pnum={sim_num}\n")
            log.exception(err)
            data = dict(fail=str(err))
        _record(args, space, sim_num, params, rv, data, writer, lib_tools._get_sim_fname(args.output_sims, sim_num))
        count(rv)

    def pipelined():
        """One model behind: the kernels of sample i are enqueued without waiting (`DeferredChecks`), then sample
        i-1 is finalised -- so the host work of building and launching a model (a few ms) runs while the GPU is still
        drawing the previous one, instead of after it."""
        from holodeck_b200 import single_sources
        pending = None
        while True:
            try:
                sim_num = todo.get_nowait()
            except queue.Empty:
                break
            if writer.store.is_done(sim_num) and not args.recreate:
                count(True)
                continue
            params = space.param_dict(sim_num)
            checks = single_sources.DeferredChecks()
            try:
                cur = (sim_num, params, _run_model(args, space, sim_num, params, device=True, deferred=checks), checks, None)
            except Exception as ee:   # noqa: BLE001
                cur = (sim_num, params, None, None, ee)
            if pending is not None:
                finalize(pending)
            pending = cur
        if pending is not None:
            finalize(pending)

    def one_sample(sim_num):
        params = space.param_dict(sim_num)
        rv, _ = run_sam_at_pspace_params(args, space, sim_num, params, writer=writer)
        with lock:
            if rv is False:
                state["failures"] += 1
            state["done"] += 1
            if (MAX_FAILURES is not None) and (state["failures"] > MAX_FAILURES):
                err = f"Failed {state['failures']} times on rank:{rank}!"
                log.exception(err)
                raise RuntimeError(err)

    def drain(own_stream):
        import contextlib
        ctx = contextlib.nullcontext()
        if own_stream:
            import torch
            ctx = torch.cuda.stream(torch.cuda.Stream())
        with ctx:
            while True:
                try:
                    sim_num = todo.get_nowait()
                except queue.Empty:
                    return
                one_sample(sim_num)

    nworkers = max(1, int(workers))
    have_cuda = False
    try:
        import torch
        have_cuda = torch.cuda.is_available()
    except Exception:   # noqa: BLE001
        pass
    if not have_cuda:
        nworkers = 1
    t_first = None
    try:
        if not todo.empty():
            one_sample(todo.get_nowait())          # the first sample fills the per-grid caches (geometry, tables) once
            t_first = (datetime.now() - beg).total_seconds()
        if nworkers > 1:
            with concurrent.futures.ThreadPoolExecutor(nworkers) as pool:
                for fut in [pool.submit(drain, True) for _ in range(nworkers)]:
                    fut.result()
        elif pipeline and writer is not None and have_cuda and (ss_flag or params_flag):
            pipelined()
        else:
            drain(False)
    finally:
        if writer is not None:
            writer.close()
    failures, num_done = state["failures"], state["done"]
    run_library.last_first_sample_s = t_first
    run_library.last_loop_s = (datetime.now() - beg).total_seconds()
    log.info(f"\t{rank} done after {run_library.last_loop_s} s")
    dist.barrier()
    return num_done, failures


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('param_space', type=str)
    ap.add_argument('output', type=str)
    ap.add_argument('-n', '--nsamples', type=int, default=1000)
    ap.add_argument('-r', '--nreals', type=int, default=DEF_NUM_REALS)
    ap.add_argument('-d', '--dur', dest='pta_dur', type=float, default=DEF_PTA_DUR)
    ap.add_argument('-f', '--nfreqs', type=int, default=DEF_NUM_FBINS)
    ap.add_argument('-s', '--shape', dest='sam_shape', type=int, default=None)
    ap.add_argument('-l', '--nloudest', type=int, default=DEF_NUM_LOUDEST)
    ap.add_argument('--gwb', dest='gwb_flag', action='store_true', default=False)
    ap.add_argument('--ss', dest='ss_flag', action='store_true', default=False)
    ap.add_argument('--params', dest='params_flag', action='store_true', default=False)
    ap.add_argument('--recreate', action='store_true', default=False)
    ap.add_argument('--seed', type=int, default=None)
    ap.add_argument('--sim-files', action='store_true', default=False,
                    help="also write the reference's per-sample library_sims/*.npz files (from the writer thread)")
    ap.add_argument('--no-streaming', action='store_true', default=False,
                    help="reference file plane only: synchronous per-sample .npz, merged by sam_lib_combine")
    ap.add_argument('--no-combine', action='store_true', default=False)
    ap.add_argument('--workers', type=int, default=1, help="host threads (CUDA streams) per rank; see run_library")
    ap.add_argument('--no-pipeline', action='store_true', default=False, help="wait for every sample before starting the next")
    args = ap.parse_args(argv)
    dist.init()
    rank, size = dist.world()
    # one seed for everybody: the Latin hypercube AND the sample permutation must be identical on all ranks
    seed = dist.shared_seed(args.seed)
    space_class = holo.librarian.param_spaces_dict[args.param_space]
    space = space_class(nsamples=args.nsamples, sam_shape=args.sam_shape, seed=seed)
    done, fails = run_library(space, args.output, nreals=args.nreals, nfreqs=args.nfreqs, nloudest=args.nloudest,
                              pta_dur=args.pta_dur, gwb_flag=args.gwb_flag, ss_flag=args.ss_flag,
                              params_flag=args.params_flag, recreate=args.recreate, seed=seed,
                              streaming=not args.no_streaming, sim_files=args.sim_files or args.no_streaming,
                              param_space_name=args.param_space, workers=args.workers, pipeline=not args.no_pipeline)
    loop_s = dist.max_over_ranks(run_library.last_loop_s)
    first_s = run_library.last_first_sample_s or 0.0
    steady_s = dist.max_over_ranks(run_library.last_loop_s - first_s)
    print(f"rank {rank}: {done} samples, {fails} failures, sample loop {run_library.last_loop_s:.3f} s "
          f"(first sample incl. one-off set-up {first_s:.3f} s)")
    if rank == 0:
        print(f"library: {args.nsamples} samples on {size} rank(s), slowest sample loop {loop_s:.3f} s "
              f"= {args.nsamples / loop_s:.1f} samples/s; after each rank's first sample: "
              f"{(args.nsamples - size) / max(steady_s, 1e-9):.1f} samples/s")
        if not args.no_combine:
            beg = datetime.now()
            fname = holo.librarian.combine.sam_lib_combine(args.output, holo.log, recreate=True)
            print(f"combined library: {fname} ({(datetime.now() - beg).total_seconds():.2f} s)")
    dist.barrier()
    dist.finalize()


if __name__ == "__main__":
    main()
