"""Declarative model recipes for library parameter spaces.

The reference spells every parameter space out as a class with hand-written ``_init_sam`` / ``_init_hard`` methods
(``librarian/param_spaces_classic.py``, ``param_spaces.py``).  Here a space is DATA: which component classes make up
the model, which setting feeds which constructor argument (with its unit), the default settings, and the sampled
distributions.  ``define_space`` turns such a description into a ``_Param_Space`` subclass with the reference's name
and behaviour; ``build_sam`` / ``build_hard`` are the two generic builders every space shares.
"""
from holodeck_b200 import sams, hardening, host_relations
from holodeck_b200.constants import GYR, PC, MSOL
from holodeck_b200.librarian.lib_tools import _Param_Space, PD_Uniform, PD_Normal, PD_Uniform_Log   # noqa: F401

# component -> (constructor, {constructor argument: setting name or (setting name, unit factor)})
COMPONENTS = {
    "gsmf_schechter": (sams.GSMF_Schechter, dict(
        phi0="gsmf_phi0_log10", phiz="gsmf_phiz", mchar0_log10="gsmf_mchar0_log10", mcharz="gsmf_mcharz",
        alpha0="gsmf_alpha0", alphaz="gsmf_alphaz")),
    # the GW-only classic spaces read the un-renamed key `gsmf_phi0` (kept from the reference, see PS_Classic_GWOnly)
    "gsmf_schechter_phi0": (sams.GSMF_Schechter, dict(
        phi0="gsmf_phi0", phiz="gsmf_phiz", mchar0_log10="gsmf_mchar0_log10", mcharz="gsmf_mcharz",
        alpha0="gsmf_alpha0", alphaz="gsmf_alphaz")),
    "gpf_power_law": (sams.GPF_Power_Law, dict(
        frac_norm_allq="gpf_frac_norm_allq", malpha="gpf_malpha", qgamma="gpf_qgamma", zbeta="gpf_zbeta",
        max_frac="gpf_max_frac")),
    "gmt_power_law": (sams.GMT_Power_Law, dict(
        time_norm=("gmt_norm", GYR), malpha="gmt_malpha", qgamma="gmt_qgamma", zbeta="gmt_zbeta")),
    "gsmf_double_schechter": (sams.GSMF_Double_Schechter, dict(
        log10_phi1=["gsmf_log10_phi_one_z0", "gsmf_log10_phi_one_z1", "gsmf_log10_phi_one_z2"],
        log10_phi2=["gsmf_log10_phi_two_z0", "gsmf_log10_phi_two_z1", "gsmf_log10_phi_two_z2"],
        log10_mstar=["gsmf_log10_mstar_z0", "gsmf_log10_mstar_z1", "gsmf_log10_mstar_z2"],
        alpha1="gsmf_alpha_one", alpha2="gsmf_alpha_two")),
    "gmr_illustris": (sams.GMR_Illustris, dict(
        norm0_log10="gmr_norm0_log10", normz="gmr_normz", malpha0="gmr_malpha0", malphaz="gmr_malphaz",
        mdelta0="gmr_mdelta0", mdeltaz="gmr_mdeltaz", qgamma0="gmr_qgamma0", qgammaz="gmr_qgammaz", qgammam="gmr_qgammam")),
    "mmbulge_kh2013": (host_relations.MMBulge_KH2013, dict(
        mamp_log10="mmb_mamp_log10", mplaw="mmb_plaw", scatter_dex="mmb_scatter_dex")),
    # normalisation given in solar masses (PS_Test, param_spaces.py:126-130: the deprecated `mamp` keyword, in grams)
    "mmbulge_kh2013_mamp": (host_relations.MMBulge_KH2013, dict(
        mamp=("mmb_mamp", MSOL), mplaw="mmb_plaw", scatter_dex="mmb_scatter_dex")),
    "bf_sigmoid": (host_relations.BF_Sigmoid, dict(
        bulge_frac_lo="bf_frac_lo", bulge_frac_hi="bf_frac_hi", mstar_char_log10="bf_mstar_crit", width_dex="bf_width_dex")),
    "hard_fixed_time_2pl": (hardening.Fixed_Time_2PL_SAM, dict(
        sepa_init=("hard_sepa_init", PC), rchar=("hard_rchar", PC), gamma_inner="hard_gamma_inner",
        gamma_outer="hard_gamma_outer")),
}


def _kwargs(component, settings):
    _, table = COMPONENTS[component]
    out = {}
    for arg, src in table.items():
        if isinstance(src, list):                       # a vector-valued argument assembled from several settings
            out[arg] = [settings[key] for key in src]
            continue
        key, unit = src if isinstance(src, tuple) else (src, None)
        out[arg] = settings[key] if unit is None else settings[key] * unit
    return out


def build_sam(sam_shape, settings, gsmf, mmbulge, gpf=None, gmt=None, gmr=None, bulge_frac=None, log=None):
    """`Semi_Analytic_Model` from the named components, each constructed from `settings` through `COMPONENTS`."""
    parts = {}
    for role, component in (("gsmf", gsmf), ("gpf", gpf), ("gmt", gmt), ("gmr", gmr)):
        if component is not None:
            parts[role] = COMPONENTS[component][0](**_kwargs(component, settings))
    extra = {}
    if bulge_frac is not None:
        extra["bulge_frac"] = COMPONENTS[bulge_frac][0](**_kwargs(bulge_frac, settings))
    parts["mmbulge"] = COMPONENTS[mmbulge][0](**_kwargs(mmbulge, settings), **extra)
    return sams.Semi_Analytic_Model(shape=sam_shape, **parts)


def build_hard(sam, settings, hard):
    if hard == "hard_gw":
        return hardening.Hard_GW()
    if hard == "hard_fixed_time_2pl":
        return hardening.Fixed_Time_2PL_SAM(sam, settings["hard_time"] * GYR, **_kwargs(hard, settings))
    raise ValueError(f"unknown hardening recipe {hard!r}")


def define_space(name, doc, defaults, sampled, sam, hard, base=_Param_Space, version=None):
    """A `_Param_Space` subclass called `name`: `sampled` is the list of `PD_*` distributions (a callable returning a
    fresh list), `sam` the keyword arguments of `build_sam`, `hard` the hardening recipe."""

    def __init__(self, log=None, nsamples=None, sam_shape=None, seed=None):
        _Param_Space.__init__(self, sampled(), log=log, nsamples=nsamples, sam_shape=sam_shape, seed=seed)

    def _init_sam(cls, sam_shape, params):
        return build_sam(sam_shape, params, **sam)

    def _init_hard(cls, sam_obj, params):
        return build_hard(sam_obj, params, hard)

    body = dict(__init__=__init__, __doc__=doc, DEFAULTS=dict(defaults), __module__=__name__,
                _init_sam=classmethod(_init_sam), _init_hard=classmethod(_init_hard))
    if version is not None:
        body["__version__"] = version
    return type(name, (base,), body)
