"""Makes `genofix_b200` importable when only `src/` is on PYTHONPATH (the reference's convention)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
