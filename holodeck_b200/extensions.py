"""Realizer for SAM populations: per-bin Poisson weights for many realizations
(``holodeck/extensions.py:105-213``, SURVEY.md "next" row N3).

``Realizer_SAM(fobs_orb_edges, sam, hard)(nreals)`` returns the bin-centre samples
``[mtot, mrat, redz_final, fobs]`` and an ``(ncell, nreals)`` matrix of Poisson weights.  The number grid
and the bin-centre final redshifts come from the same fused device pass as in ``sam.gwb`` (K1 -> K2+K2b);
the weight matrix is drawn by the bulk Philox sampler (``holo_poisson_as_needed``).
"""
import numpy as np

import holodeck_b200 as holo
from holodeck_b200 import _lib, cosmo, utils, gravwaves
from holodeck_b200.sams import sam_cyutils


def get_samples_from_edges(edges, redz, number_shape, flatten=True):
    """Bin-centre (mtot, mrat, redz_final, fobs_gw) of every (M,Q,Z,F) bin (``extensions.py:172-226``).

    ``redz`` is the (M,Q,Z,F) grid of final redshifts at bin EDGES; bins whose 8-corner mean is <= 0 get -1.
    """
    edges_np = [np.asarray(ee.cpu()) if _lib.is_device_array(ee) else np.asarray(ee, dtype=float) for ee in edges]
    mtot = utils.midpoints(edges_np[0])
    mrat = utils.midpoints(edges_np[1])
    fobs = 2.0 * utils.midpoints(edges_np[3])
    zmid = _lib.to_host(gravwaves._char_strain_sq(edges_np, redz, params=True)["zmid"])
    shape = (mtot.size, mrat.size, zmid.shape[2], fobs.size)
    mtot = np.broadcast_to(mtot[:, None, None, None], shape)
    mrat = np.broadcast_to(mrat[None, :, None, None], shape)
    fobs = np.broadcast_to(fobs[None, None, None, :], shape)
    if np.any([mtot.shape != tuple(number_shape), zmid.shape != tuple(number_shape)]):
        err = f"Sample shapes don't all match number! {mtot.shape=}, {zmid.shape=}, {number_shape=}"
        raise ValueError(err)
    if flatten:
        return [np.ascontiguousarray(mtot).flatten(), np.ascontiguousarray(mrat).flatten(), zmid.flatten(),
                np.ascontiguousarray(fobs).flatten()]
    return [np.array(mtot), np.array(mrat), zmid, np.array(fobs)]


class Realizer_SAM:
    """Draw realizations of the binary population of a SAM (``extensions.py:105-166``)."""

    def __init__(self, fobs_orb_edges, sam=None, hard=None, params=None, pspace=None):
        if params is not None:
            if sam is not None or hard is not None:
                raise ValueError("Only 'params' or ('sam' and 'hard') should be provided.")
            if pspace is None:
                pspace = holo.librarian.PS_Classic_Phenom_Uniform
            pspace = pspace()
            sam, hard = pspace.model_for_params(params=params, sam_shape=pspace.sam_shape)
        elif sam is None or hard is None:
            raise ValueError("'params' or ('sam' and 'hard') must be provided.")
        self._sam = sam
        self._hard = hard
        self._fobs_orb_edges = np.asarray(fobs_orb_edges, dtype=float)

    def __call__(self, nreals=100, clean=False, *, seed=None):
        sam, hard = self._sam, self._hard
        fobs_orb_edges = self._fobs_orb_edges
        fobs_orb_cents = utils.midpoints(fobs_orb_edges)
        redz, diff_num = sam_cyutils.dynamic_binary_number_at_fobs(fobs_orb_cents, sam, hard, cosmo, device=True)
        edges = [sam.mtot, sam.mrat, sam.redz, fobs_orb_edges]
        number = sam_cyutils.integrate_differential_number_3dx1d(edges, diff_num)   # device in -> device out
        samples = get_samples_from_edges(edges, redz, tuple(number.shape), flatten=True)
        names = ['mtot', 'mrat', 'redz', 'fobs']
        flat = number.reshape(-1)
        shape = (flat.numel(), int(nreals))
        weights = gravwaves.poisson_as_needed(flat[:, None].expand(shape).contiguous(), seed=seed, device=False)
        if clean:
            nonzero_samples, nonzero_weights = [], []
            for rr in range(nreals):
                nonzero = weights[:, rr] != 0
                nonzero_samples.append([ss[nonzero] for ss in samples])
                nonzero_weights.append(weights[:, rr][nonzero])
            weights, samples = nonzero_weights, nonzero_samples
        return names, samples, weights
