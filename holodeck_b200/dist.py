"""Multi-GPU sharding of the SAM hot path: one process per GPU, ``torch.distributed`` for the plumbing.

The path shards by *independent units* (SURVEY.md section 8e), with no collective on the data path:

* **realizations** -- every rank holds a replica of the deterministic grids (they are recomputed
  locally: cheaper than broadcasting 1 GB) and draws its slice of the R realizations; the Philox
  counter is keyed on the GLOBAL realization index, so the union over ranks is bit-identical to a
  single-GPU run.  One ``all_gather`` of the small per-rank ``hc`` tables at the end (NCCL).
* **library samples** -- the parameter samples of a library are permuted and split over ranks exactly
  as ``holodeck/librarian/gen_lib.py:139-141,169`` does over MPI ranks.

On CPU-only machines the same code runs on the ``gloo`` backend (tests use world_size 2).
"""
import os

import numpy as np


def is_distributed():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def world():
    """(rank, world_size) of this process (``(0, 1)`` when not launched under torchrun)."""
    if is_distributed():
        import torch.distributed as dist
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """Initialise the default process group from the torchrun environment (no-op for world size 1)."""
    import torch
    import torch.distributed as dist
    size = int(os.environ.get("WORLD_SIZE", "1"))
    if size <= 1 or dist.is_initialized():
        return world()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return world()


def realization_slice(nreals, rank=None, size=None):
    """Contiguous slice ``(r0, count)`` of ``nreals`` realizations owned by ``rank`` (sizes differ by <= 1)."""
    if rank is None or size is None:
        rank, size = world()
    base, extra = divmod(int(nreals), size)
    count = base + (1 if rank < extra else 0)
    r0 = rank * base + min(rank, extra)
    return r0, count


def gather_realizations(tensor, axis=1, nreals=None, out=None):
    """All-gather per-rank result tables along the realization ``axis`` (ragged slices allowed).

    ONE collective: every rank contributes a contiguous ``(cmax, ...)`` block (realization axis first, zero-padded
    when the slices are ragged) to a single ``all_gather_into_tensor`` on a preallocated ``(size*cmax, ...)``
    buffer (``out``, reused by the caller from step to step) -- instead of a list-form ``all_gather`` plus a
    ``torch.cat`` per table.  Returns the gathered table with the realization axis back in place.
    """
    import torch
    import torch.distributed as dist
    if not is_distributed() or dist.get_world_size() == 1:
        return tensor
    size = dist.get_world_size()
    if nreals is None:
        counts = [tensor.shape[axis]] * size
    else:
        counts = [realization_slice(nreals, rr, size)[1] for rr in range(size)]
    cmax = max(counts)
    moved = tensor.movedim(axis, 0).contiguous()
    if moved.shape[0] < cmax:   # pad ragged slices so every rank sends the same shape
        pad = torch.zeros((cmax - moved.shape[0],) + tuple(moved.shape[1:]), dtype=moved.dtype, device=moved.device)
        moved = torch.cat([moved, pad], dim=0)
    want = (size * cmax,) + tuple(moved.shape[1:])           # rank blocks concatenated along dim 0
    if out is None or tuple(out.shape) != want or out.dtype != moved.dtype or out.device != moved.device:
        out = torch.empty(want, dtype=moved.dtype, device=moved.device)
    dist.all_gather_into_tensor(out, moved)
    if all(cc == cmax for cc in counts):
        full = out
    else:
        full = torch.cat([out[rr * cmax:rr * cmax + cc] for rr, cc in enumerate(counts)], dim=0)
    return full.movedim(0, axis)


def gather_tables(tensors, out=None):
    """Gather several per-rank tables that share their leading (F, R) axes -- ``hc_ss (F,R,L)`` and ``hc_bg (F,R)``
    of one ``sam.gwb`` step -- with ONE collective: they are packed side by side into ``(R, F, sum of widths)``,
    gathered into the preallocated ``out`` (``(size*R, F, W)``; pass the previous step's return value ``buf``), and
    handed back as views ``(F, size*R, ...)``.  Returns ``(views, buf)``."""
    import torch
    import torch.distributed as dist
    if not is_distributed() or dist.get_world_size() == 1:
        return list(tensors), out
    size = dist.get_world_size()
    F, R = tensors[0].shape[:2]
    cols = [tt.reshape(F, R, -1) for tt in tensors]
    widths = [cc.shape[2] for cc in cols]
    packed = torch.cat(cols, dim=2).permute(1, 0, 2).contiguous()          # (R, F, W)
    want = (size * R,) + tuple(packed.shape[1:])
    if out is None or tuple(out.shape) != want or out.dtype != packed.dtype or out.device != packed.device:
        out = torch.empty(want, dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed)
    full = out.permute(1, 0, 2)                                              # (F, size*R, W) view
    views, beg = [], 0
    for tt, ww in zip(tensors, widths):
        views.append(full[:, :, beg:beg + ww].reshape((F, size * R) + tuple(tt.shape[2:])))
        beg += ww
    return views, out


def shared_seed(seed=None):
    """The seed every rank uses for things that must be identical everywhere (parameter space, sample permutation).
    ``None`` -> rank 0 draws one from the OS; either way rank 0's value is broadcast (the reference builds the space
    and the permutation on rank 0 and broadcasts / scatters them, ``gen_lib.py:99-169``)."""
    if seed is None:
        seed = int.from_bytes(os.urandom(4), "little")
    if not is_distributed():
        return int(seed)
    import torch.distributed as dist
    box = [int(seed)]
    dist.broadcast_object_list(box, src=0)
    return int(box[0])


def max_over_ranks(value):
    """max of a host float over the ranks (device-side when the backend is NCCL)."""
    if not is_distributed():
        return float(value)
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    tt = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt.item())


def sample_indices(nsamples, seed=None, rank=None, size=None):
    """Indices of the library samples this rank runs: permute, then ``array_split`` over ranks
    (``gen_lib.py:139-141``).  Every rank computes the same permutation from ``seed``."""
    if rank is None or size is None:
        rank, size = world()
    rng = np.random.RandomState(seed)
    indices = rng.permutation(np.arange(nsamples))
    return np.array_split(indices, size)[rank]


def barrier():
    if is_distributed():
        import torch.distributed as dist
        dist.barrier()


def finalize():
    """Tear the process group down (end of ``gen_lib.main``); a no-op for a single-process run."""
    if is_distributed():
        import torch.distributed as dist
        dist.destroy_process_group()
