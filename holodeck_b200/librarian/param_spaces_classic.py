"""The 'classic' parameter spaces of the NANOGrav 15 yr analysis (``librarian/param_spaces_classic.py``), declared as
data for ``librarian.recipes``: model components, default settings (``:13-42, 150-176``) and sampled distributions
(``:92-145, 235-253``).  Class names, parameter names, defaults and distributions are the reference's."""
from holodeck_b200.librarian.recipes import define_space, PD_Uniform as U, PD_Normal as N

# default settings shared by the phenomenological spaces: 2-power-law hardening [Gyr, pc], GSMF fit to [Tomczak+2014]
# (`sam-parameters.ipynb`), pair fraction, merger time [Gyr] (qgamma: Boylan-Kolchin+2008), M-Mbulge (mean of MM2013 and
# KH2013) with its 0.3 dex scatter
_GALAXY = dict(
    gsmf_phiz=-0.6, gsmf_mchar0_log10=11.24, gsmf_mcharz=0.11, gsmf_alpha0=-1.21, gsmf_alphaz=-0.03,
    gpf_frac_norm_allq=0.025, gpf_malpha=0.0, gpf_qgamma=0.0, gpf_zbeta=1.0, gpf_max_frac=1.0,
    gmt_norm=0.5, gmt_malpha=0.0, gmt_qgamma=-1.0, gmt_zbeta=-0.5,
    mmb_mamp_log10=8.69, mmb_plaw=1.10, mmb_scatter_dex=0.3,
)
_PHENOM = dict(hard_time=3.0, hard_sepa_init=1e4, hard_rchar=100.0, hard_gamma_inner=-1.0, hard_gamma_outer=+2.5,
               gsmf_phi0_log10=-2.77, **_GALAXY)
_PHENOM_MODEL = dict(sam=dict(gsmf="gsmf_schechter", gpf="gpf_power_law", gmt="gmt_power_law", mmbulge="mmbulge_kh2013"),
                     hard="hard_fixed_time_2pl")

_PS_Classic_Phenom = define_space(
    "_PS_Classic_Phenom", "Base of the classic phenomenological spaces (no sampled parameters).", _PHENOM, lambda: [], **_PHENOM_MODEL)

PS_Classic_Phenom_Uniform = define_space(
    "PS_Classic_Phenom_Uniform", "Classic 6D phenomenological, uniform parameter space ('phenom-uniform').", _PHENOM,
    lambda: [U("gsmf_phi0_log10", -3.5, -1.5), U("gsmf_mchar0_log10", 10.5, 12.5), U("mmb_mamp_log10", +7.6, +9.0),
             U("mmb_scatter_dex", +0.0, +0.9), U("hard_time", 0.1, 11.0), U("hard_gamma_inner", -1.5, +0.0)],
    base=_PS_Classic_Phenom, **_PHENOM_MODEL)

PS_Classic_Phenom_Astro_Extended = define_space(
    "PS_Classic_Phenom_Astro_Extended", "Classic 12D phenomenological parameter space ('phenom-astro+extended'): "
    "normal priors from `sam-parameters.ipynb` fits to [Tomczak+2014] with 4x standard deviations.", _PHENOM,
    lambda: [U("hard_time", 0.1, 11.0), U("hard_gamma_inner", -1.5, +0.5),
             N("gsmf_phi0", -2.56, 0.4), N("gsmf_mchar0_log10", 10.9, 0.4), N("gsmf_alpha0", -1.2, 0.2),
             N("gpf_zbeta", +0.8, 0.4), N("gpf_qgamma", +0.5, 0.3),
             U("gmt_norm", 0.2, 5.0), U("gmt_zbeta", -2.0, +0.0),
             N("mmb_mamp_log10", +8.6, 0.2), N("mmb_plaw", +1.2, 0.2), N("mmb_scatter_dex", +0.32, 0.15)],
    base=_PS_Classic_Phenom, **_PHENOM_MODEL)

# GW-only evolution.  Kept from the reference (`:148-253`): the defaults and the model read `gsmf_phi0`, while a sampled
# parameter of that name is renamed to `gsmf_phi0_log10` (lib_tools.PARAM_NAMES_REPLACE), and the sampled `mmb_scatter`
# is not the `mmb_scatter_dex` the model reads -- those two samples do not reach the model.
_GWONLY_MODEL = dict(sam=dict(gsmf="gsmf_schechter_phi0", gpf="gpf_power_law", gmt="gmt_power_law", mmbulge="mmbulge_kh2013"),
                     hard="hard_gw")
_PS_Classic_GWOnly = define_space(
    "_PS_Classic_GWOnly", "Base of the classic GW-only spaces.", dict(gsmf_phi0=-2.77, **_GALAXY), lambda: [], **_GWONLY_MODEL)

PS_Classic_GWOnly_Uniform = define_space(
    "PS_Classic_GWOnly_Uniform", "Classic 4D GW-only, uniform parameter space ('gw-only').", dict(gsmf_phi0=-2.77, **_GALAXY),
    lambda: [U("gsmf_phi0", -3.5, -1.5), U("gsmf_mchar0_log10", 10.5, 12.5), U("mmb_mamp_log10", +7.5, +9.5),
             U("mmb_scatter", +0.0, +1.2)],
    base=_PS_Classic_GWOnly, **_GWONLY_MODEL)
