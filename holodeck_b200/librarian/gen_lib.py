"""GPU-sharded library generation (``holodeck/librarian/gen_lib.py:50-337`` with ``torch.distributed``
instead of mpi4py).

Run under torchrun, one rank per GPU::

    torchrun --nnodes=1 --nproc-per-node 8 -m holodeck_b200.librarian.gen_lib PS_Classic_Phenom_Uniform OUT -n 2000 -r 100

Each rank builds the same parameter space from the seed, takes its share of the permuted sample
indices and, per sample, runs ``run_model`` and writes ``library_sims/library__pNNNNNN.npz`` with the
reference's keys.  A failing sample produces a file holding a single ``fail`` key and is re-attempted
on the next run (``gen_lib.py:287-319``); more than ``MAX_FAILURES`` failures abort the rank.
"""
import argparse
from datetime import datetime
from pathlib import Path

import numpy as np

import holodeck_b200 as holo
from holodeck_b200 import dist
from holodeck_b200.librarian import (
    DEF_NUM_REALS, DEF_NUM_FBINS, DEF_NUM_LOUDEST, DEF_PTA_DUR, DIRNAME_LIBRARY_SIMS, lib_tools,
)

MAX_FAILURES = 5


def run_sam_at_pspace_params(args, space, pnum, params):
    """Run sample ``pnum`` of ``space``; returns ``(ok, sim_fname)`` (``gen_lib.py:236-337``)."""
    log = args.log
    sim_fname = lib_tools._get_sim_fname(args.output_sims, pnum)
    if sim_fname.exists():
        temp = np.load(sim_fname)
        if 'fail' in list(temp.keys()):
            log.info("Existing file was a failure, re-attempting...")
        elif not args.recreate:
            return True, sim_fname
    try:
        sam, hard = space.model_for_params(params)
        data = lib_tools.run_model(
            sam, hard,
            pta_dur=args.pta_dur, nfreqs=args.nfreqs, nreals=args.nreals, nloudest=args.nloudest,
            gwb_flag=args.gwb_flag, singles_flag=args.ss_flag, details_flag=False, params_flag=args.params_flag,
            log=log, seed=None if args.seed is None else (int(args.seed) * 1000003 + int(pnum)),
        )
        data['params'] = np.array([params[pn] for pn in space.param_names])
        data['param_names'] = space.param_names
        rv = True
    except Exception as err:   # noqa: BLE001  (same catch-all as the reference)
        log.exception(f"`run_model` FAILED on {pnum=}\n")
        log.exception(err)
        rv = False
        data = dict(fail=str(err))
    np.savez(sim_fname, **data)
    return rv, sim_fname


def run_library(space, output, nreals=DEF_NUM_REALS, nfreqs=DEF_NUM_FBINS, nloudest=DEF_NUM_LOUDEST,
                pta_dur=DEF_PTA_DUR, gwb_flag=True, ss_flag=True, params_flag=False, recreate=False, seed=None,
                log=None, indices=None):
    """Generate this rank's share of the library; returns ``(num_done, failures)``."""
    rank, size = dist.world()
    log = holo.log if log is None else log
    output = Path(output)
    output_sims = output.joinpath(DIRNAME_LIBRARY_SIMS)
    if rank == 0:
        output_sims.mkdir(parents=True, exist_ok=True)
        space.save(output)
    dist.barrier()
    args = argparse.Namespace(
        log=log, output=output, output_sims=output_sims, recreate=recreate, pta_dur=pta_dur, nfreqs=nfreqs,
        nreals=nreals, nloudest=nloudest, gwb_flag=gwb_flag, ss_flag=ss_flag, params_flag=params_flag, seed=seed)
    if indices is None:
        indices = dist.sample_indices(space.nsamples, seed=seed, rank=rank, size=size)
    beg = datetime.now()
    failures = 0
    num_done = 0
    for sim_num in indices:
        params = space.param_dict(int(sim_num))
        rv, _ = run_sam_at_pspace_params(args, space, int(sim_num), params)
        if rv is False:
            failures += 1
        if (MAX_FAILURES is not None) and (failures > MAX_FAILURES):
            err = f"Failed {failures} times on rank:{rank}!"
            log.exception(err)
            raise RuntimeError(err)
        num_done += 1
    log.info(f"\t{rank} done after {(datetime.now() - beg).total_seconds()} s")
    dist.barrier()
    return num_done, failures


def main():
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
    args = ap.parse_args()
    dist.init()
    space_class = holo.librarian.param_spaces_dict[args.param_space]
    space = space_class(nsamples=args.nsamples, sam_shape=args.sam_shape, seed=args.seed)
    done, fails = run_library(space, args.output, nreals=args.nreals, nfreqs=args.nfreqs, nloudest=args.nloudest,
                              pta_dur=args.pta_dur, gwb_flag=args.gwb_flag, ss_flag=args.ss_flag,
                              params_flag=args.params_flag, recreate=args.recreate, seed=args.seed)
    print(f"rank {dist.world()[0]}: {done} samples, {fails} failures")
    dist.finalize()


if __name__ == "__main__":
    main()
