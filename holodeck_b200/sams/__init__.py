"""Semi-analytic models (mirrors ``holodeck/sams/__init__.py``)."""
from holodeck_b200.sams import components   # noqa
from holodeck_b200.sams.components import (   # noqa
    _Galaxy_Pair_Fraction, _Galaxy_Stellar_Mass_Function, _Galaxy_Merger_Time, _Galaxy_Merger_Rate,
    GSMF_Schechter, GSMF_Double_Schechter, GPF_Power_Law, GMT_Power_Law, GMR_Illustris,
)
from holodeck_b200.sams import sam   # noqa
from holodeck_b200.sams.sam import Semi_Analytic_Model, evolve_eccen_uniform_single   # noqa
from holodeck_b200.sams import sam_cyutils   # noqa
