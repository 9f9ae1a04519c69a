"""Moved to ``synthetic_inputs.py`` at the repo root (the seeded weight / input generator is shared
with bench.py, which must not import anything under oracle/).  Kept as an alias for the tests and
``oracle/gen_golden.py``."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from synthetic_inputs import *  # noqa: F401,F403,E402
from synthetic_inputs import VIS_CFG, fill_state_dict, make_inputs, make_param, visibility_case_inputs  # noqa: F401,E402
