"""Alias so that existing code doing `import nerfacc` picks up nerfacc_b200 unchanged.

Submodules are aliased too, so `from nerfacc.pdf import searchsorted`,
`from nerfacc.estimators.prop_net import PropNetEstimator`, ... keep working.
"""
import importlib as _importlib
import sys as _sys

from nerfacc_b200 import *  # noqa: F401,F403
from nerfacc_b200 import __all__, __version__  # noqa: F401

for _name in ("data_specs", "grid", "pack", "scan", "volrend", "pdf", "losses", "parallel", "estimators", "estimators.base",
              "estimators.occ_grid", "estimators.prop_net"):
    _mod = _importlib.import_module("nerfacc_b200." + _name)
    _sys.modules[__name__ + "." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod
del _name, _mod
