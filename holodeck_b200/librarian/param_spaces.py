"""The astrophysically motivated parameter spaces (``librarian/param_spaces.py``), declared as data for
``librarian.recipes``: double-Schechter GSMF ([Leja2020]_ via `double-schechter.ipynb`), Illustris galaxy merger rate
([Rodriguez-Gomez2015]_), [KH2013]_ M-Mbulge with a sigmoid bulge fraction, 2-power-law hardening.  Class names,
parameter names, defaults (``:25-57, 162-207``) and distributions (``:67-82, 272-340``) are the reference's."""
from holodeck_b200.librarian.recipes import define_space, PD_Uniform as U, PD_Normal as N

_GMR = dict(gmr_norm0_log10=-2.2287, gmr_normz=+2.4644, gmr_malpha0=+0.2241, gmr_malphaz=-1.1759, gmr_mdelta0=+0.7668,
            gmr_mdeltaz=-0.4695, gmr_qgamma0=-1.2595, gmr_qgammaz=+0.0611, gmr_qgammam=-0.0477)
_GMR_SIGMA = dict(gmr_norm0_log10=0.0045, gmr_normz=0.0128, gmr_malpha0=0.0038, gmr_malphaz=0.0316, gmr_mdelta0=0.0202,
                  gmr_mdeltaz=0.0440, gmr_qgamma0=0.0026, gmr_qgammaz=0.0021, gmr_qgammam=0.0013)

PS_Test = define_space(
    "PS_Test", "Simple test space: single-Schechter GSMF, Illustris merger rate, KH2013 (normalisation in Msol), 2PL hardening.",
    dict(hard_time=3.0, hard_sepa_init=1e4, hard_rchar=100.0, hard_gamma_inner=-1.0, hard_gamma_outer=+2.5,
         gsmf_phi0_log10=-2.77, gsmf_phiz=-0.6, gsmf_mchar0_log10=11.24, gsmf_mcharz=0.11, gsmf_alpha0=-1.21, gsmf_alphaz=-0.03,
         mmb_mamp=0.49e9, mmb_plaw=1.17, mmb_scatter_dex=0.28, **_GMR),
    lambda: [U("hard_time", 0.1, 11.0, default=3.0), U("hard_gamma_inner", -1.5, +0.0, default=-1.0), N("mmb_mamp", 0.49e9, 0.055e9)],
    sam=dict(gsmf="gsmf_schechter", gmr="gmr_illustris", mmbulge="mmbulge_kh2013_mamp"), hard="hard_fixed_time_2pl")

_GSMF2 = dict(gsmf_log10_phi_one_z0=-2.383, gsmf_log10_phi_one_z1=-0.264, gsmf_log10_phi_one_z2=-0.107,
              gsmf_log10_phi_two_z0=-2.818, gsmf_log10_phi_two_z1=-0.368, gsmf_log10_phi_two_z2=+0.046,
              gsmf_log10_mstar_z0=+10.767, gsmf_log10_mstar_z1=+0.124, gsmf_log10_mstar_z2=-0.033,
              gsmf_alpha_one=-0.28, gsmf_alpha_two=-1.48)
_GSMF2_SIGMA = dict(gsmf_log10_phi_one_z0=0.028, gsmf_log10_phi_one_z1=0.072, gsmf_log10_phi_one_z2=0.031,
                    gsmf_log10_phi_two_z0=0.050, gsmf_log10_phi_two_z1=0.070, gsmf_log10_phi_two_z2=0.020,
                    gsmf_log10_mstar_z0=0.026, gsmf_log10_mstar_z1=0.045, gsmf_log10_mstar_z2=0.015,
                    gsmf_alpha_one=0.070, gsmf_alpha_two=0.150)
_STRONG = dict(hard_time=3.0, hard_sepa_init=1e4, hard_rchar=10.0, hard_gamma_inner=-1.0, hard_gamma_outer=0.0,
               mmb_mamp_log10=8.69, mmb_plaw=1.17, mmb_scatter_dex=0.28,
               bf_frac_lo=0.4, bf_frac_hi=0.8, bf_mstar_crit=11.0, bf_width_dex=1.0, **_GSMF2, **_GMR)
_STRONG_MODEL = dict(sam=dict(gsmf="gsmf_double_schechter", gmr="gmr_illustris", mmbulge="mmbulge_kh2013", bulge_frac="bf_sigmoid"),
                     hard="hard_fixed_time_2pl")


def _hardening_priors():
    return [U("hard_time", 0.1, 11.0, default=3.0), U("hard_gamma_inner", -2.0, +0.0, default=-1.0),
            U("hard_rchar", 2.0, 20.0, default=10.0)]


_PS_Astro_Strong = define_space("_PS_Astro_Strong", "Base of the strongly astrophysically motivated spaces.", _STRONG,
                                lambda: [], version="0.2", **_STRONG_MODEL)

PS_Astro_Strong_All = define_space(
    "PS_Astro_Strong_All", "All 29 parameters: hardening, GSMF and merger-rate fits (normal, published uncertainties), "
    "M-Mbulge and bulge fraction.", _STRONG,
    lambda: (_hardening_priors() + [N(kk, _GSMF2[kk], ss) for kk, ss in _GSMF2_SIGMA.items()] +
             [N(kk, _GMR[kk], ss) for kk, ss in _GMR_SIGMA.items()] +
             [N("mmb_mamp_log10", 8.69, 0.05), N("mmb_plaw", 1.17, 0.08), N("mmb_scatter_dex", 0.28, 0.05),
              U("bf_frac_lo", 0.1, 0.4), U("bf_frac_hi", 0.6, 1.0), U("bf_width_dex", 0.5, 1.5)]),
    base=_PS_Astro_Strong, version="0.2", **_STRONG_MODEL)

PS_Astro_Strong_Hard = define_space(
    "PS_Astro_Strong_Hard", "Only the three hardening parameters vary.", _STRONG, _hardening_priors,
    base=_PS_Astro_Strong, version="0.2", **_STRONG_MODEL)
