"""Marker classes for the compiled reference's ``isinstance`` dispatch (see package docstring).

The attributes read by the reference (`sam_cyutils.pyx:472-476`) are set by `oracle/glue.py`.
"""


class _Hardening:
    pass


class Hard_GW(_Hardening):
    CONSISTENT = False


class Fixed_Time_2PL_SAM(_Hardening):
    CONSISTENT = True

    def __init__(self, norm, sepa_init, rchar, gamma_inner, gamma_outer, num_steps, target_time=None):
        self._norm = norm
        self._sepa_init = sepa_init
        self._rchar = rchar
        self._gamma_inner = gamma_inner
        self._gamma_outer = gamma_outer
        self._num_steps = num_steps
        self._target_time = target_time
