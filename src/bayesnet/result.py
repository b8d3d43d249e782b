"""bayesnet.result -- same module path as the reference."""
import _bootstrap  # noqa: F401
from genofix_b200.bayesnet.result import ResultGroupInfer, ResultPairInfer, ResultSingleInfer  # noqa: F401
