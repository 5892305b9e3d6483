"""holodeck_b200 -- B200-native implementation of holodeck's semi-analytic-model GW-background path.

The package mirrors the part of the ``holodeck`` namespace that lies on the hot path
(SURVEY.md section 8): ``holodeck_b200.sams.Semi_Analytic_Model``, ``holodeck_b200.hardening``,
``holodeck_b200.host_relations``, ``holodeck_b200.gravwaves``, ``holodeck_b200.single_sources``,
``holodeck_b200.librarian`` and the two drop-in native modules ``holodeck_b200.sams.sam_cyutils`` /
``holodeck_b200.cyutils`` whose loops run as hand-written sm_100a CUDA kernels in
``libholo_b200.so`` (C ABI in ``include/holo_b200.h``).  There is no CPU fallback.
"""
import logging

__all__ = ["log", "cosmo"]

__version__ = "0.1.0"


class Parameters:
    """WMAP9 parameters, [WMAP9]_ Table 3, WMAP+BAO+H0 (``holodeck/__init__.py:48-53``)."""
    Omega0 = 0.2880                #: Matter density parameter "Om0"
    OmegaBaryon = 0.0472           #: Baryon density parameter "Ob0"
    HubbleParam = 0.6933           #: Hubble Parameter as H0/[100 km/s/Mpc]


# ---- logger (``holodeck/__init__.py:76``; stdlib logging, WARNING level)
log = logging.getLogger(__name__)
if not log.handlers:
    _handler = logging.StreamHandler()
    _handler.setFormatter(logging.Formatter("%(asctime)s %(levelname)s : %(message)s [%(filename)s:%(funcName)s]"))
    log.addHandler(_handler)
    log.setLevel(logging.WARNING)
    log.propagate = False

# ---- cosmology instance (``holodeck/__init__.py:81-85``); must exist before the submodules load
from holodeck_b200.cosmology import Cosmology   # noqa: E402
cosmo = Cosmology(h=Parameters.HubbleParam, Om0=Parameters.Omega0, Ob0=Parameters.OmegaBaryon, size=200)

from holodeck_b200 import constants       # noqa: E402,F401
from holodeck_b200 import utils           # noqa: E402,F401
from holodeck_b200 import host_relations  # noqa: E402,F401
from holodeck_b200 import hardening       # noqa: E402,F401
from holodeck_b200 import cyutils         # noqa: E402,F401
from holodeck_b200 import gravwaves       # noqa: E402,F401
from holodeck_b200 import single_sources  # noqa: E402,F401
from holodeck_b200 import sams            # noqa: E402,F401
from holodeck_b200.sams import sam        # noqa: E402,F401
from holodeck_b200 import dist            # noqa: E402,F401
from holodeck_b200 import librarian       # noqa: E402,F401
from holodeck_b200 import extensions      # noqa: E402,F401
