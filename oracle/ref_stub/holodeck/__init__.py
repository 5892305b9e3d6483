"""Stub ``holodeck`` package that hosts the reference's two compiled Cython modules.

TEST INFRASTRUCTURE ONLY.  `oracle/build_ref.py` cythonizes the reference's own
``holodeck/cyutils.pyx`` and ``holodeck/sams/sam_cyutils.pyx`` (read in place from
``/root/reference``) and drops the resulting ``.so`` files next to this file under
``oracle/_ref/holodeck``.  The compiled ``sam_cyutils`` does ``import holodeck as holo`` and
dispatches on ``isinstance(hard, holo.hardening.Fixed_Time_2PL_SAM)`` /
``holo.hardening.Hard_GW`` (reference ``holodeck/sams/sam_cyutils.pyx:465,483``), so the only
thing this stub has to provide is those two marker classes.  No reference source lives here.
"""
from . import hardening  # noqa: F401
