"""empirical.jalleledist3 -- same module path as the reference."""
import _bootstrap  # noqa: F401
from genofix_b200.empirical.jalleledist3 import JointAllellicDistribution  # noqa: F401
