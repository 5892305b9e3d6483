"""Physical constants in CGS, mirroring ``holodeck/constants.py:23-57`` of the reference.

The reference pulls these from astropy (CODATA-2018); astropy is not a dependency here, so the
same numbers are hard-coded.  NOTE: the reference's *Cython* kernels use a slightly different
set (``sam_cyutils.pyx:30-42``, ``cyutils.pyx:43-48``: ``MY_NWTG = 6.6742999e-08``,
``MY_MPC = 3.08567758e+24``); those live in ``csrc/holo_constants.cuh`` and are used only by the
kernels that replace Cython code, so both sides of the boundary keep their own rounding.
"""
import numpy as np

# ---- Fundamental Constants
NWTG = 6.6743e-08                  #: Newton's Gravitational Constant [cm^3/g/s^2]
SPLC = 29979245800.0               #: Speed of light [cm/s]
MELC = 9.1093837015e-28            #: Electron Mass [g]
MPRT = 1.67262192369e-24           #: Proton Mass [g]
KBOLTZ = 1.380649e-16              #: Boltzmann constant [erg/K]
HPLANCK = 6.62607015e-27           #: Planck constant [erg/s]
SIGMA_SB = 5.6703744191844314e-05  #: Stefan-Boltzmann constant [erg/cm^2/s/K^4]
SIGMA_T = 6.6524587321000005e-25   #: Thomson cross-section [cm^2]

# ---- Typical astronomy units
MSOL = 1.988409870698051e+33       #: Solar Mass [g]
LSOL = 3.828e+33                   #: Solar Luminosity [erg/s]
RSOL = 69570000000.0               #: Solar Radius [cm]
PC = 3.0856775814913674e+18        #: Parsec [cm]
AU = 14959787070000.0              #: Astronomical Unit [cm]
YR = 31557600.0                    #: year [s]
KMPERSEC = 1.0e5                   #: km/s [cm/s]

# ---- Derived Constants
SCHW = 2*NWTG/(SPLC*SPLC)                        #: Schwarzschild Constant (2*G/c^2) [cm]
EDDT = 4.0*np.pi*NWTG*SPLC*MPRT/SIGMA_T          #: Eddington Luminosity prefactor factor [erg/s/g]

DAY = 86400.0                                   #: Day [s]
MYR = 1.0e6*YR                                  #: Mega-year [s]
GYR = 1.0e9*YR                                  #: Giga-year [s]
KPC = 1.0e3*PC                                  #: Kilo-parsec [cm]
MPC = 1.0e6*PC                                  #: Mega-parsec [cm]
GPC = 1.0e9*PC                                  #: Giga-parsec [cm]
