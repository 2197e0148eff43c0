"""Alias so that existing code doing `import nerfacc` picks up nerfacc_b200 unchanged."""
from nerfacc_b200 import *  # noqa: F401,F403
from nerfacc_b200 import __all__, __version__, estimators, grid, pack, scan, volrend  # noqa: F401
