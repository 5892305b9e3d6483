"""Library generation: parameter spaces and the per-sample driver (hot-path subset of
``holodeck/librarian``): ``_Param_Space`` / ``PD_*`` / ``run_model`` (``lib_tools.py``), the classic
parameter spaces (``param_spaces_classic.py``) and a GPU-sharded ``gen_lib`` (samples partitioned over
ranks exactly as ``gen_lib.py:139-141,169`` does over MPI ranks; ``torch.distributed`` instead of mpi4py).

``combine.sam_lib_combine`` merges the per-sample files.  Out of scope (SURVEY.md section 2a
row 11): ``fit_spectra``, ``posterior_populations``.
"""

__version__ = "1.3"

DEF_NUM_REALS = 100     #: Default number of realizations to construct in libraries.
DEF_NUM_FBINS = 40      #: Default number of frequency bins at which to calculate GW signals.
DEF_NUM_LOUDEST = 5     #: Default number of loudest binaries to calculate in each frequency bin.
DEF_PTA_DUR = 16.03     #: Default PTA duration which determines Nyquist frequency bins [yrs].

FNAME_LIBRARY_SIM_FILE = "library__p{pnum:06d}.npz"
DIRNAME_LIBRARY_SIMS = "library_sims"
PSPACE_FILE_SUFFIX = ".pspace.npz"
ARGS_CONFIG_FNAME = "config.json"

from holodeck_b200.librarian import lib_tools   # noqa: E402
from holodeck_b200.librarian.lib_tools import (   # noqa: E402,F401
    _Param_Space, _Param_Dist, PD_Uniform, PD_Uniform_Log, PD_Normal, run_model,
)
from holodeck_b200.librarian import recipes, param_spaces_classic, combine   # noqa: E402,F401
from holodeck_b200.librarian.param_spaces_classic import (   # noqa: E402,F401
    PS_Classic_Phenom_Uniform, PS_Classic_Phenom_Astro_Extended, PS_Classic_GWOnly_Uniform,
)

from holodeck_b200.librarian import param_spaces   # noqa: E402
from holodeck_b200.librarian.param_spaces import PS_Test, PS_Astro_Strong_All, PS_Astro_Strong_Hard   # noqa: E402,F401

param_spaces_dict = {
    "PS_Classic_Phenom_Uniform": PS_Classic_Phenom_Uniform,
    "PS_Classic_Phenom_Astro_Extended": PS_Classic_Phenom_Astro_Extended,
    "PS_Classic_GWOnly_Uniform": PS_Classic_GWOnly_Uniform,
    "PS_Test": PS_Test,
    "PS_Astro_Strong_All": PS_Astro_Strong_All,
    "PS_Astro_Strong_Hard": PS_Astro_Strong_Hard,
}
