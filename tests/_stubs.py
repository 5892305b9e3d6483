"""Light-weight stand-ins for `sam` / `hard` objects built from golden fixtures (tests only)."""
import numpy as np


class Log:
    def info(self, *args, **kwargs):
        pass


class GoldenSam:
    """Duck-typed SAM (what `sam_cyutils.dynamic_binary_number_at_fobs` reads, sam_cyutils.pyx:457-497)."""

    def __init__(self, gg):
        self.mtot = gg["mtot"]
        self.mrat = gg["mrat"]
        self.redz = gg["redz"]
        self.shape = (self.mtot.size, self.mrat.size, self.redz.size)
        self.static_binary_density = gg["dens"]
        self._gmt_time = gg.get("gmt_time")
        self._redz_prime = gg.get("redz_prime")
        self._log = Log()


class GoldenCosmo:
    def __init__(self, gg):
        self._grid_z = gg["grid_z"]
        self._grid_dcom = gg["grid_dcom"]
        self._grid_age = gg["grid_age"]


def golden_hard(gg, holo):
    """Hardening instance matching the fixture, with `_norm` taken from the golden file."""
    if str(gg["hard"]) == "gw":
        return holo.hardening.Hard_GW()
    hp = gg["hard_params"]
    hard = holo.hardening.Fixed_Time_2PL_SAM.__new__(holo.hardening.Fixed_Time_2PL_SAM)
    hard._target_time = hp[0]
    hard._sepa_init = hp[1]
    hard._rchar = hp[2]
    hard._gamma_inner = hp[3]
    hard._gamma_outer = hp[4]
    hard._num_steps = int(hp[5])
    hard._norm_host = 10.0 ** gg["norm_log10"]
    hard._norm_dev = None
    hard._norm_device = lambda: hard._norm_host
    return hard


def edges_orb(gg):
    return [gg["mtot"], gg["mrat"], gg["redz"], gg["fobs_edges"] / 2.0]


def sort_indices(gg):
    order = gg["order"].astype(np.int64)
    Mb, Qb, Zb = [nn - 1 for nn in (gg["mtot"].size, gg["mrat"].size, gg["redz"].size)]
    zsort = order % Zb
    mq = order // Zb
    return mq // Qb, mq % Qb, zsort
