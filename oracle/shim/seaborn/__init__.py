"""No-op seaborn stand-in (TEST INFRASTRUCTURE ONLY)."""
from matplotlib.pyplot import _Ax


def distplot(*a, **k):
    return _Ax()
